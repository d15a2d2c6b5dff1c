"""The product's per-path stage functions, compiled for the host by tests/hostsim (test tooling),
against the oracle: same integer outcomes, same radiance, same work counters.  This is how the
device logic is debugged in the GPU-less container; the GPU run of the same checks is
tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests.util import SMALL_SCENES, assert_radiance_close, assert_stats_close, bits_equal

NAMES = list(SMALL_SCENES)


@pytest.mark.parametrize("name", NAMES)
def test_primary_hits_bit_exact(oracle_api, hostsim_api, name):
    sd = SMALL_SCENES[name]()
    a = sd.realize(oracle_api).render_ids(1, msaa=2)
    b = sd.realize(hostsim_api).render_ids(1, msaa=2)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("integrator,depth", [("path", 1), ("direct", 5), ("path", 5)])
def test_per_sample_radiance_and_counters(oracle_api, hostsim_api, name, integrator, depth):
    sd = SMALL_SCENES[name]()
    kw = dict(integrator=integrator, msaa=2, max_depth=depth, crop=(8, 8, 64, 48), flags=1)
    a, sa = sd.realize(oracle_api).render_samples(**kw)
    b, sb = sd.realize(hostsim_api).render_samples(**kw)
    assert_radiance_close(b, a, f"{name} {integrator} depth {depth}", outliers=1e-3)
    assert_stats_close(sb, sa, f"{name} {integrator} depth {depth}")


@pytest.mark.parametrize("name", ["edge_mesh", "edge_single", "cornell"])
def test_axis_aligned_rays_no_jitter(oracle_api, hostsim_api, name):
    """PBRS_FLAG_NO_JITTER: the centre column / row rays have exactly-zero direction components,
    i.e. infinite reciprocals and 0/0 slabs -- the walker's exact-division path."""
    sd = SMALL_SCENES[name]()
    ho, hh = sd.realize(oracle_api), sd.realize(hostsim_api)
    a = ho.render_ids(0, msaa=1, flags=4)
    b = hh.render_ids(0, msaa=1, flags=4)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fa, sa = ho.render_samples(integrator="path", msaa=1, max_depth=4, flags=4 | 1)
    fb, sb = hh.render_samples(integrator="path", msaa=1, max_depth=4, flags=4 | 1)
    assert_radiance_close(fb, fa, name + " no-jitter", outliers=1e-3)
    assert_stats_close(sb, sa, name + " no-jitter")


def test_film_and_splits(oracle_api, hostsim_api):
    sd = SMALL_SCENES["cornell"]()
    ho, hh = sd.realize(oracle_api), sd.realize(hostsim_api)
    full_o, _ = ho.render(integrator="path", msaa=2)
    full_h, _ = hh.render(integrator="path", msaa=2)
    assert_radiance_close(full_h, full_o, "film", outliers=1e-3)
    # tile split: the two ranks' films add up to the full film exactly (x + 0)
    t0, _ = hh.render(integrator="path", msaa=2, rank=0, world_size=2, split="tiles")
    t1, _ = hh.render(integrator="path", msaa=2, rank=1, world_size=2, split="tiles")
    assert ((t0 == 0) | (t1 == 0)).all()
    assert bits_equal(t0 + t1, full_h).all()
    # sample split: partial sums add up to the full sum up to fp32 summation order
    s0, _ = hh.render(integrator="path", msaa=2, rank=0, world_size=2, split="samples", flags=8)
    s1, _ = hh.render(integrator="path", msaa=2, rank=1, world_size=2, split="samples", flags=8)
    np.testing.assert_allclose((s0 + s1) * 0.25, full_h, rtol=1e-5, atol=1e-7)
    so0, _ = ho.render(integrator="path", msaa=2, rank=0, world_size=2, split="samples", flags=8)
    assert_radiance_close(s0, so0, "sample-split partial film", outliers=1e-3)
