"""Parity of the CUDA path (through the C ABI, on a B200) against the CPU oracle.

Bars (BASELINE.json north_star, DESIGN.md "Parity"):
  * primary-hit instance / primitive ids and t: bit-exact;
  * per-sample radiance at depth 1 (and beyond): 1e-4 relative, a bounded fraction of documented
    transcendental-ulp outliers;
  * traversal work counters equal to the oracle's instrumented walk (same boxes, same order);
  * full images: <= 1 % RMSE against the oracle's 4096-spp render;
  * size-independent properties at a full BASELINE config size (tile split = full frame exactly,
    sample split = full frame to summation order, determinism, linearity of the raw sum).
"""
import json
import os

import numpy as np
import pytest

from pbrs_b200 import scenes
from tests.util import SMALL_SCENES, assert_radiance_close, assert_stats_close, bits_equal

pytestmark = pytest.mark.gpu
NAMES = list(SMALL_SCENES)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SAMPLE_OUTLIERS = {}


def _pinned_sample_outliers():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "parity_outliers_samples.json")) as f:
            return json.load(f)
    except FileNotFoundError:
        return {}


@pytest.mark.parametrize("name", NAMES)
def test_primary_hits_bit_exact(oracle_api, gpu_api, name):
    sd = SMALL_SCENES[name]()
    ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
    for sample in (0, 3):
        a = ho.render_ids(sample, msaa=2)
        b = hg.render_ids(sample, msaa=2)
        assert (a[0] == b[0]).all(), f"{name}: {(a[0] != b[0]).sum()} instance ids differ"
        assert (a[1] == b[1]).all(), f"{name}: {(a[1] != b[1]).sum()} primitive ids differ"
        assert bits_equal(a[2], b[2]).all(), f"{name}: hit t differs"


def test_primary_hits_bit_exact_full_c2_frame(oracle_api, gpu_api):
    """BASELINE configs[1]: the Cornell box at 1920x1080, every primary ray."""
    sd = scenes.cornell_box(1920, 1080)
    a = sd.realize(oracle_api).render_ids(0, msaa=1)
    b = sd.realize(gpu_api).render_ids(0, msaa=1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()


def test_primary_hits_no_jitter_crop_and_edges(oracle_api, gpu_api):
    """Ragged inputs: a crop that straddles tile borders, odd frame sizes, jitter off."""
    sd = scenes.spheres500(131, 77, n_small=60)
    ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
    for crop in [(0, 0, 131, 77), (60, 60, 71, 17), (63, 0, 2, 77), (130, 76, 1, 1)]:
        a = ho.render_ids(0, msaa=1, crop=crop, flags=4)
        b = hg.render_ids(0, msaa=1, crop=crop, flags=4)
        assert a[0].shape == (crop[3], crop[2])
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all(), crop


@pytest.mark.parametrize("name", ["edge_mesh", "edge_single", "cornell"])
def test_axis_aligned_rays_no_jitter(oracle_api, gpu_api, name):
    """PBRS_FLAG_NO_JITTER: the centre column / row rays have exactly-zero direction components
    (infinite reciprocals, 0/0 slabs): the walker's exact-division path on the device."""
    sd = SMALL_SCENES[name]()
    ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
    a = ho.render_ids(0, msaa=1, flags=4)
    b = hg.render_ids(0, msaa=1, flags=4)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fa, sa = ho.render_samples(integrator="path", msaa=1, max_depth=4, flags=4 | 1)
    fb, sb = hg.render_samples(integrator="path", msaa=1, max_depth=4, flags=4 | 1)
    assert_radiance_close(fb, fa, name + " no-jitter", outliers=1e-3)
    assert_stats_close(sb, sa, name + " no-jitter")


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("integrator,depth", [("path", 1), ("direct", 5), ("path", 5)])
def test_per_sample_radiance_and_counters(oracle_api, gpu_api, name, integrator, depth):
    sd = SMALL_SCENES[name]()
    kw = dict(integrator=integrator, msaa=2, max_depth=depth, flags=1)
    a, sa = sd.realize(oracle_api).render_samples(**kw)
    b, sb = sd.realize(gpu_api).render_samples(**kw)
    n_bad = assert_radiance_close(b, a, f"{name} {integrator} depth {depth}", outliers=1e-3)
    exact = float(bits_equal(a, b).all(axis=-1).mean())
    print(f"{name} {integrator} d{depth}: {exact * 100:.3f}% of samples bit-identical, {n_bad} beyond 1e-4")
    assert exact > 0.98
    # the allowance is pinned to what was MEASURED on a B200 (tests/golden/parity_outliers_samples.json,
    # rewritten into profiles/ with PBRS_WRITE_OUTLIERS=1): the outliers are last-ulp differences between
    # glibc's float transcendentals and the device's FP64-evaluated ones, so another host libm may move
    # the count by a few -- twice the measured count plus two is still 10-100x tighter than 0.1 %
    key = f"{name}/{integrator}/d{depth}"
    _SAMPLE_OUTLIERS[key] = {"samples": int(a.shape[0] * a.shape[1] * a.shape[2]), "beyond_1e-4": n_bad, "bit_identical_fraction": exact}
    if os.environ.get("PBRS_WRITE_OUTLIERS"):
        with open(os.path.join(ROOT, "profiles", "parity_outliers_samples.json"), "w") as f:
            json.dump(_SAMPLE_OUTLIERS, f, indent=1, sort_keys=True)
    pinned = _pinned_sample_outliers().get(key)
    if pinned is not None:
        assert n_bad <= 2 * pinned["beyond_1e-4"] + 2, f"{key}: {n_bad} samples beyond 1e-4, measured {pinned['beyond_1e-4']} when pinned"
    assert_stats_close(sb, sa, f"{name} {integrator} depth {depth}")
    assert sb["launches"] > 0 and sb["launches_extend"] > 0


def test_depth1_per_pixel_radiance_c1(oracle_api, gpu_api):
    """north_star: per-pixel radiance within 1e-4 relative at depth 1 (C1 scene, 16 spp)."""
    sd = scenes.cornell_box(256, 256)
    a, _ = sd.realize(oracle_api).render(integrator="path", msaa=4, max_depth=1)
    b, _ = sd.realize(gpu_api).render(integrator="path", msaa=4, max_depth=1)
    assert_radiance_close(b, a, "C1 depth-1 film", outliers=1e-4)


def test_full_image_rmse_vs_4096spp_oracle(oracle_api, gpu_api):
    """north_star: full images within 1 % RMSE of a 4096-spp converged reference.  The reference
    is the oracle's 4096-spp render (msaa 64) of a Cornell-box crop; the GPU renders the same
    crop at 4096 spp through the full path integrator (depth 5)."""
    sd = scenes.cornell_box(512, 512)
    crop = (224, 288, 64, 40)
    ref, _ = sd.realize(oracle_api).render(integrator="path", msaa=64, max_depth=5, crop=crop)
    got, st = sd.realize(gpu_api).render(integrator="path", msaa=64, max_depth=5, crop=crop)
    r = ref[crop[1]:crop[1] + crop[3], crop[0]:crop[0] + crop[2]]
    g = got[crop[1]:crop[1] + crop[3], crop[0]:crop[0] + crop[2]]
    assert np.isfinite(g).all()
    rmse = float(np.sqrt(np.mean((g - r) ** 2)) / np.mean(r))
    print(f"relative RMSE vs 4096-spp oracle: {rmse:.3e}")
    assert rmse <= 0.01
    assert (got[:crop[1]] == 0).all()  # pixels outside the crop stay 0


def test_film_equals_oracle_c3_family(oracle_api, gpu_api):
    sd = scenes.spheres500(320, 180)
    a, sa = sd.realize(oracle_api).render(integrator="path", msaa=4, max_depth=5, flags=1)
    b, sb = sd.realize(gpu_api).render(integrator="path", msaa=4, max_depth=5, flags=1)
    assert_radiance_close(b, a, "C3-family film", tol=1e-3, outliers=2e-3)
    assert_stats_close(sb, sa, "C3-family film", rel=1e-3)


def test_properties_at_full_c1_size(gpu_api):
    """Size-independent properties at BASELINE configs[0] size (512x512, 16 spp, depth 5)."""
    h = scenes.cornell_box(512, 512).realize(gpu_api)
    full, st = h.render(integrator="path", msaa=4)
    assert st["n_samples"] == 512 * 512 * 16
    again, _ = h.render(integrator="path", msaa=4)
    assert bits_equal(full, again).all(), "not deterministic"
    # a different number of paths in flight changes batching, not the film
    small, _ = h.render(integrator="path", msaa=4, paths_in_flight=300_000)
    assert bits_equal(full, small).all(), "film depends on the batch size"
    # tile split over 4 ranks: disjoint, and the sum is the full film exactly
    parts = [h.render(integrator="path", msaa=4, rank=r, world_size=4, split="tiles")[0] for r in range(4)]
    assert bits_equal(sum(parts[1:], parts[0]), full).all()
    owned = sum((p != 0).any(axis=-1).astype(np.int32) for p in parts)
    assert owned.max() <= 1
    # sample split over 4 ranks: raw partial sums add up to the raw full sum (fp32 order only)
    raw, _ = h.render(integrator="path", msaa=4, flags=8)
    np.testing.assert_allclose(raw / 16.0, full, rtol=1e-6, atol=1e-7)
    sparts = [h.render(integrator="path", msaa=4, rank=r, world_size=4, split="samples", flags=8)[0] for r in range(4)]
    np.testing.assert_allclose(sum(sparts[1:], sparts[0]), raw, rtol=1e-5, atol=1e-6)
    # a crop renders the same pixels as the full frame
    crop, _ = h.render(integrator="path", msaa=4, crop=(100, 200, 130, 70))
    assert bits_equal(crop[200:270, 100:230], full[200:270, 100:230]).all()
    assert (crop[:200] == 0).all()


def test_graph_replay_equals_direct_enqueue(gpu_api):
    """Frames of a few batches are captured once and replayed as a CUDA graph (kernels.cu): same
    film as enqueuing kernel by kernel (PBRS_FLAG_NO_GRAPH), replay after replay, and a changed
    option set is a new capture, not a stale replay."""
    import torch
    from pbrs_b200 import _capi as K
    h = scenes.cornell_box(256, 192).realize(gpu_api)
    for kw in (dict(integrator="path", msaa=2), dict(integrator="direct", msaa=1), dict(integrator="path", msaa=3, paths_in_flight=200_000)):
        direct, sd = h.render(flags=K.FLAG_NO_GRAPH, **kw)
        for _ in range(3):
            g, sg = h.render(**kw)
            assert bits_equal(g, direct).all(), kw
            assert (sg["n_samples"], sg["n_rays_extend"], sg["n_rays_shadow"], sg["launches"]) == \
                   (sd["n_samples"], sd["n_rays_extend"], sd["n_rays_shadow"], sd["launches"])
    # the same graph on a caller's stream and film
    film = torch.empty((192, 256, 3), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    want, _ = h.render(integrator="path", msaa=2, flags=K.FLAG_NO_GRAPH)
    for _ in range(2):
        film.zero_()
        h.render_device(film.data_ptr(), stream=s.cuda_stream, integrator="path", msaa=2)
        s.synchronize()
        assert bits_equal(film.cpu().numpy(), want).all()


def test_render_device_keeps_the_film_on_the_gpu(gpu_api):
    import torch
    h = scenes.cornell_box(128, 128).realize(gpu_api)
    host, _ = h.render(integrator="direct", msaa=1)
    film = torch.empty((128, 128, 3), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    h.render_device(film.data_ptr(), stream=s.cuda_stream, integrator="direct", msaa=1)
    s.synchronize()
    assert bits_equal(film.cpu().numpy(), host).all()


def test_mesh_scale_walk_matches_oracle(oracle_api, gpu_api):
    """A 131k-triangle height-field + 1280-triangle icospheres at a larger frame: the BLAS walk
    (ordered stack, near child first, leaf runs) against the oracle's, ids and counters."""
    sd = scenes.mesh_terrain(480, 270, grid=256, ico_subdiv=3, tex_size=128)
    ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
    a = ho.render_ids(0, msaa=1)
    b = hg.render_ids(0, msaa=1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fa, sa = ho.render(integrator="path", msaa=1, max_depth=3, flags=1)
    fb, sb = hg.render(integrator="path", msaa=1, max_depth=3, flags=1)
    assert_radiance_close(fb, fa, "terrain film", outliers=1e-3)
    assert_stats_close(sb, sa, "terrain film", rel=1e-3)


def test_full_size_c4_scene_properties_and_oracle_crop(oracle_api, gpu_api):
    """The headline workload's own scene (1,002,528-triangle terrain, 3840x2160) at 1 spp: the
    oracle agrees on a crop (ids bit-exact, per-sample radiance, counters), and the full frame has
    the size-independent properties: deterministic, independent of the batch size, a 2-rank tile
    split sums to it bit-exactly, a crop is its sub-rectangle."""
    sd = scenes.mesh_terrain()
    hg, ho = sd.realize(gpu_api), sd.realize(oracle_api)
    assert hg.info().n_triangles > 1_000_000
    crop = (1700, 1200, 96, 64)
    a, b = ho.render_ids(0, msaa=1, crop=crop), hg.render_ids(0, msaa=1, crop=crop)
    assert (a[0] != 0xFFFFFFFF).mean() > 0.9
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fa, sa = ho.render_samples(integrator="path", msaa=1, max_depth=5, flags=1, crop=crop)
    fb, sb = hg.render_samples(integrator="path", msaa=1, max_depth=5, flags=1, crop=crop)
    assert_radiance_close(fb, fa, "C4 crop", outliers=1e-3)
    assert_stats_close(sb, sa, "C4 crop", rel=1e-3)
    full, st = hg.render(integrator="path", msaa=1)
    assert st["n_samples"] == 3840 * 2160
    again, _ = hg.render(integrator="path", msaa=1, paths_in_flight=1_500_000)
    assert bits_equal(full, again).all()
    t0, _ = hg.render(integrator="path", msaa=1, rank=0, world_size=2, split="tiles")
    t1, _ = hg.render(integrator="path", msaa=1, rank=1, world_size=2, split="tiles")
    assert bits_equal(t0 + t1, full).all() and ((t0 == 0) | (t1 == 0)).all()
    sub, _ = hg.render(integrator="path", msaa=1, crop=crop)
    x, y, w, h = crop
    assert bits_equal(sub[y:y + h, x:x + w], full[y:y + h, x:x + w]).all()
    # the film of the crop equals the mean of the oracle's samples there (1 spp: the sample itself)
    assert_radiance_close(full[y:y + h, x:x + w], fa[:, :, 0, :], "C4 film vs oracle samples", outliers=1e-3)


def test_instanced_walk_matches_oracle(oracle_api, gpu_api):
    """TLAS of 1600 transformed instances: the recursive closest-hit emulation incl. extent quirks."""
    sd = scenes.instanced_field(480, 270, n_side=40, n_meshes=5, ico_subdiv=2, n_lights=8)
    ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
    a = ho.render_ids(0, msaa=1)
    b = hg.render_ids(0, msaa=1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fa, sa = ho.render(integrator="path", msaa=1, max_depth=3, flags=1)
    fb, sb = hg.render(integrator="path", msaa=1, max_depth=3, flags=1)
    assert_radiance_close(fb, fa, "field film", outliers=1e-3)
    assert_stats_close(sb, sa, "field film", rel=1e-3)


@pytest.mark.parametrize("name", ["terrain", "zoo_image", "field", "preset_everything"])
def test_per_scene_kernel_switches_do_not_change_the_film(gpu_api, name, monkeypatch):
    """Commit picks per scene between the one-piece and the split (surface + scatter) shade kernels
    and between the sequential and the cooperative closest-hit leaf phase (DeviceScene::shade_split,
    ::coop_closest: scene-size heuristics).  Both choices are scheduling only: forced either way,
    the film is bit-identical and the counters are equal."""
    sd = SMALL_SCENES[name]()
    ref = None
    for split, coop in ((0, 0), (1, 0), (0, 1), (1, 1)):
        monkeypatch.setenv("PBRS_SHADE_SPLIT", str(split))
        monkeypatch.setenv("PBRS_COOP_CLOSEST", str(coop))
        h = sd.realize(gpu_api)
        film, st = h.render(integrator="path", msaa=2, max_depth=5, flags=1)
        # (would_panic counts threads that saw an assert, per kernel: the KINDS are scheduling-independent, the counts are not)
        key = (st["n_samples"], st["n_rays_extend"], st["n_rays_shadow"], st["n_nodes"], st["n_tris"], st["n_spheres"], st["n_instances"], tuple(sorted(st["would_panic"])))
        if ref is None:
            ref = (film, key)
        else:
            assert bits_equal(film, ref[0]).all(), (name, split, coop, int((~bits_equal(film, ref[0])).sum()))
            assert key == ref[1], (name, split, coop, key, ref[1])
