"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/pbrs_gpu.h declares, mirrors the header's struct layouts, validates arguments like the
reference's constructors assert them -- and refuses to render without a CUDA device (there is no
CPU fallback).  No compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pbrs_b200 import _capi as K
from pbrs_b200 import _ffi, scenes
from pbrs_b200.scene import PbrsError, SceneDesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "pbrs_gpu.h")).read()


def test_library_exports_every_header_symbol():
    code = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    declared = set(re.findall(r"\b(pbrs_[a-z0-9_]+)\s*\(", code))
    assert declared == set(K.HEADER_SYMBOLS), declared ^ set(K.HEADER_SYMBOLS)
    lib = C.CDLL(_ffi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} not exported by libpbrs_gpu.so"


def test_abi_version_and_struct_sizes():
    api = _ffi.load()
    assert api["abi_version"]() == 2
    # sizes implied by the header's field lists on x86-64
    assert C.sizeof(K.MaterialDesc) == 4 * 5 + 12 + 12 + 16 + 4
    assert C.sizeof(K.RenderOpts) == 64
    assert C.sizeof(K.Stats) == 8 * 7 + 8 * 16 + 8 * 6 + 8 * 2 + 8 * 9
    assert C.sizeof(K.SceneInfo) == 4 * 9 + 4 + 8 + 24


def test_sampler_matches_the_oracle(oracle_api):
    api = _ffi.load()
    rng = np.random.default_rng(1)
    for _ in range(200):
        seed, px, s, d = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**32)), int(rng.integers(0, 4096)), int(rng.integers(0, 64))
        assert api["sampler_u32"](seed, px, s, d) == oracle_api["sampler_u32"](seed, px, s, d)


def test_argument_validation_mirrors_reference_asserts():
    api = _ffi.load()
    sd = SceneDesc()
    sd.set_camera(64, 64, 40.0, (0, 0, -5), (0, 0, 0))
    m = sd.lambertian((0.5, 0.5, 0.5))
    sd.add_instance(sd.add_sphere((0, 0, 0), 1.0), m)
    h = sd.realize(api, commit=False)
    # bad ids
    assert api["scene_add_instance"](h.ptr, 7, 0, None, None) == K.ERR_INVALID_ARG
    assert api["scene_add_instance"](h.ptr, 0, 9, None, None) == K.ERR_INVALID_ARG
    assert b"material" in api["last_error"]()
    # geometry/src/transform.rs:277 asserts w == 1: the bottom row must be (0,0,0,1)
    bad = np.eye(4, dtype=np.float32); bad[3, 0] = 0.5
    fa = np.ascontiguousarray(bad.T.reshape(-1)); ia = np.ascontiguousarray(np.eye(4, dtype=np.float32).reshape(-1))
    assert api["scene_add_instance"](h.ptr, 0, 0, fa.ctypes.data_as(K.c_float_p), ia.ctypes.data_as(K.c_float_p)) == K.ERR_INVALID_ARG
    # lambertian without a texture, unknown material kind, NaN mesh position, index out of range
    d = K.MaterialDesc(); d.kind = K.MTL_LAMBERTIAN; d.tex_kd = -1
    assert api["scene_add_material"](h.ptr, C.byref(d)) == K.ERR_INVALID_ARG
    d.kind = 42
    assert api["scene_add_material"](h.ptr, C.byref(d)) == K.ERR_INVALID_ARG
    P = np.array([[0, 0, 0], [1, 0, 0], [0, np.nan, 0]], np.float32); idx = np.array([[0, 1, 2]], np.uint32)
    assert api["scene_add_mesh"](h.ptr, P.ctypes.data_as(K.c_float_p), None, None, 3, idx.ctypes.data_as(K.c_u32_p), 1) == K.ERR_INVALID_ARG
    P[2, 1] = 1.0; idx[0, 2] = 3
    assert api["scene_add_mesh"](h.ptr, P.ctypes.data_as(K.c_float_p), None, None, 3, idx.ctypes.data_as(K.c_u32_p), 1) == K.ERR_INVALID_ARG
    # degenerate look-at: the reference panics in Vec3::hat
    eye = np.zeros(3, np.float32)
    assert api["scene_set_camera"](h.ptr, 64, 64, 0.7, eye.ctypes.data_as(K.c_float_p), eye.ctypes.data_as(K.c_float_p),
                                   eye.ctypes.data_as(K.c_float_p)) == K.ERR_INVALID_ARG
    # render before commit
    o = h.make_opts()
    out = np.zeros((64, 64, 3), np.float32)
    assert api["render"](h.ptr, C.byref(o), out.ctypes.data_as(K.c_float_p), None) == K.ERR_STATE


def test_commit_needs_camera_and_instances():
    api = _ffi.load()
    sd = SceneDesc()
    sd.lambertian((0.5, 0.5, 0.5))
    h = sd.realize(api, commit=False)
    assert api["scene_commit"](h.ptr) == K.ERR_STATE  # no camera
    sd2 = SceneDesc()
    sd2.set_camera(64, 64, 40.0, (0, 0, -5), (0, 0, 0))
    h2 = sd2.realize(api, commit=False)
    assert api["scene_commit"](h2.ptr) == K.ERR_STATE  # tlas/src/bvh.rs:117 "empty instances"


def test_no_cpu_fallback():
    """Without a CUDA device commit fails loudly with PBRS_ERR_NO_DEVICE; nothing renders on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    api = _ffi.load()
    with pytest.raises(PbrsError) as e:
        scenes.cornell_box(32, 32).realize(api)
    assert e.value.code == K.ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under pbrs_b200/ may import, link or name it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pbrs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_ffi" not in text and "liboracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
                assert "hostsim" not in text or f.endswith(".cuh"), os.path.join(dirpath, f)


def test_rust_bindings_are_generated_from_the_header_and_complete():
    """ffi/pbrs_gpu.rs (the Rust `extern "C"` side of the boundary) is generated from
    include/pbrs_gpu.h by tools/gen_rust_ffi.py: it must be up to date, bind every function the header
    declares with the argument count of the ctypes mirror, and carry the structs field for field."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode == 0, \
        "ffi/pbrs_gpu.rs is stale: run python tools/gen_rust_ffi.py"
    rs = open(os.path.join(ROOT, "ffi", "pbrs_gpu.rs")).read()
    header = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    declared = set(re.findall(r"\b(pbrs_[a-z0-9_]+)\s*\(", header))
    bound = dict(re.findall(r"pub fn (pbrs_[a-z0-9_]+)\(([^)]*)\)", rs))
    assert declared == set(bound), (sorted(declared - set(bound)), sorted(set(bound) - declared))
    table = dict(K.SCENE_API)
    table.update(K.PRODUCT_ONLY_API)
    for name, (_, argtypes) in table.items():
        n_rs = len([a for a in bound["pbrs_" + name].split(",") if a.strip()])
        assert n_rs == len(argtypes), f"pbrs_{name}: {n_rs} Rust parameters vs {len(argtypes)} in the ctypes mirror"
    for cname, ctype in (("pbrs_render_opts", K.RenderOpts), ("pbrs_stats", K.Stats), ("pbrs_scene_info", K.SceneInfo), ("pbrs_material_desc", K.MaterialDesc)):
        body = re.search(r"pub struct %s \{(.*?)\}" % cname, rs, flags=re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [f for f, _ in ctype._fields_], cname
    # the recorder and the patches name only functions that exist
    for f in ("gpu_scene.rs", "main_rs.patch", "loader_rs.patch"):
        text = open(os.path.join(ROOT, "ffi", f)).read()
        for used in set(re.findall(r"\b(pbrs_[a-z0-9_]+)\s*\(", text)):
            assert used in declared, f"ffi/{f} calls {used}, which the header does not declare"
    assert "ffi/pbrs_gpu.rs" in open(os.path.join(ROOT, "INTEGRATION.md")).read()
