"""The film return paths and the multi-GPU fan-out of pbrs_render (include/pbrs_gpu.h, ABI 2).

  * PBRS_FLAG_OWN_TILES_ONLY: ranks that share one host film assemble the frame by copying only
    their own tiles (emulated on one GPU: the ranks run one after the other into the same film);
  * page-locked films (pbrs_film_alloc, pbrs_host_register) as the target of pbrs_render;
  * num_gpus = N: ONE pbrs_render call over N devices equals the single-device film -- bit-exactly
    for a tile split, to fp32 summation order for a sample split (needs >= 2 GPUs, else skipped).
"""
import ctypes as C

import numpy as np
import pytest

from pbrs_b200 import _capi as K
from pbrs_b200 import scenes
from tests.util import bits_equal

pytestmark = pytest.mark.gpu


def _pinned_film(api, h, w):
    p = api["film_alloc"](w, h)
    assert p, "pbrs_film_alloc failed"
    arr = np.ctypeslib.as_array(C.cast(p, K.c_float_p), shape=(h, w, 3))
    return p, arr


def test_own_tiles_only_assembles_the_frame_in_one_host_film(gpu_api):
    h = scenes.cornell_box(300, 200).realize(gpu_api)  # ragged: 5 x 4 tiles, partial last row / column
    full, _ = h.render(integrator="path", msaa=2)
    shared = np.full((200, 300, 3), -7.0, np.float32)
    for r in range(3):
        h.render(integrator="path", msaa=2, rank=r, world_size=3, split="tiles", flags=K.FLAG_OWN_TILES_ONLY, out=shared, want_stats=False)
        if r < 2:
            assert (shared == -7.0).any(), "a rank wrote tiles it does not own"
    assert bits_equal(shared, full).all()
    # without the flag a rank's film has zeros where it owns nothing (the documented default)
    part, _ = h.render(integrator="path", msaa=2, rank=1, world_size=3, split="tiles")
    assert ((part == 0) | bits_equal(part, full)).all() and (part == 0).any()


def test_page_locked_films(gpu_api):
    h = scenes.cornell_box(256, 192).realize(gpu_api)
    want, _ = h.render(integrator="direct", msaa=1)
    p, arr = _pinned_film(gpu_api, 192, 256)
    try:
        arr[...] = 3.0
        h.render(integrator="direct", msaa=1, out=arr, want_stats=False)
        assert bits_equal(arr, want).all()
    finally:
        arr = None
        gpu_api["film_free"](p)
    own = np.zeros((192, 256, 3), np.float32)
    assert gpu_api["host_register"](own.ctypes.data, own.nbytes) == 0
    try:
        h.render(integrator="direct", msaa=1, out=own, want_stats=False)
        assert bits_equal(own, want).all()
    finally:
        assert gpu_api["host_unregister"](own.ctypes.data) == 0
    assert gpu_api["check_last_frame"](h.ptr) == 0
    assert gpu_api["device_count"]() >= 1


def test_num_gpus_one_call_many_devices(gpu_api):
    n = gpu_api["device_count"]()
    if n < 2:
        pytest.skip("one CUDA device: the multi-device call needs at least two")
    n = min(n, 4)
    h = scenes.mesh_terrain(480, 270, grid=96, ico_subdiv=2, tex_size=128).realize(gpu_api)
    one, s1 = h.render(integrator="path", msaa=2)
    many, sn = h.render(integrator="path", msaa=2, num_gpus=n, split="tiles")
    assert bits_equal(many, one).all(), "tile split over devices differs from the single-device film"
    assert sn["n_samples"] == s1["n_samples"] and sn["n_rays_extend"] == s1["n_rays_extend"] and sn["n_rays_shadow"] == s1["n_rays_shadow"]
    crop = (100, 60, 200, 130)
    c1, _ = h.render(integrator="path", msaa=2, crop=crop)
    cn, _ = h.render(integrator="path", msaa=2, crop=crop, num_gpus=n)
    assert bits_equal(cn, c1).all()
    samp, ss = h.render(integrator="path", msaa=4, num_gpus=n, split="samples")
    ref, _ = h.render(integrator="path", msaa=4)
    np.testing.assert_allclose(samp, ref, rtol=1e-5, atol=1e-6)
    assert ss["n_samples"] == 480 * 270 * 16
    # an impossible device count is an error, not a fallback
    o = h.make_opts(integrator="path", msaa=1, num_gpus=gpu_api["device_count"]() + 1)
    out = np.zeros((270, 480, 3), np.float32)
    assert gpu_api["render"](h.ptr, C.byref(o), out.ctypes.data_as(K.c_float_p), None) == K.ERR_NO_DEVICE
