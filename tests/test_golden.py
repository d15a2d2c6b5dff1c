"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit, and so does the host build of the product's
stage functions.  GPU: the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from tests.util import SMALL_SCENES, assert_radiance_close, bits_equal

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["cornell", "spheres", "terrain", "field", "zoo_image", "shape_zoo", "preset_cornell"]


def _check(api, name, exact):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    crop = tuple(int(v) for v in g["crop"])
    h = SMALL_SCENES[name]().realize(api)
    inst, prim, t = h.render_ids(0, msaa=2, crop=crop)
    assert (inst == g["inst"]).all() and (prim == g["prim"]).all(), name
    assert bits_equal(t, g["t"]).all(), name
    d1, _ = h.render_samples(integrator="path", msaa=2, max_depth=1, crop=crop)
    dl, _ = h.render_samples(integrator="direct", msaa=2, max_depth=5, crop=crop)
    if exact:
        assert bits_equal(d1, g["path_depth1"]).all() and bits_equal(dl, g["direct"]).all(), name
    else:
        assert_radiance_close(d1, g["path_depth1"], name + " path depth 1", outliers=1e-3)
        assert_radiance_close(dl, g["direct"], name + " direct", outliers=1e-3)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle_api, name):
    _check(oracle_api, name, exact=True)


@pytest.mark.parametrize("name", NAMES)
def test_hostsim_reproduces_golden(hostsim_api, name):
    _check(hostsim_api, name, exact=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(gpu_api, name):
    _check(gpu_api, name, exact=False)
