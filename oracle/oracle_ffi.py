"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from pbrs_b200 import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_math.h", "oracle_geom.h", "oracle_scene.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "pbrs_gpu.h"))
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_api = None
_lib = None


def load():
    global _api, _lib
    if _api is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _api = K.bind(_lib, "oracle_", K.SCENE_API)
        extra = {
            "render_rows": (C.c_int, [K.P, C.POINTER(K.RenderOpts), K.c_float_p, C.POINTER(K.Stats), C.c_uint32]),
            "set_threads": (None, [C.c_int]),
            "get_threads": (C.c_int, []),
            "trace_ray": (C.c_int, [K.P, K.c_float_p, K.c_float_p, C.c_float, K.c_float_p]),
            "occludes_ray": (C.c_int, [K.P, K.c_float_p, K.c_float_p, C.c_float]),
            "kat": (C.c_int, [C.c_int, K.c_float_p, C.c_int, K.c_float_p, C.c_int]),
        }
        _api.update(K.bind(_lib, "oracle_", extra))
    return _api


# KAT opcodes (oracle.cpp)
KAT_FRESNEL_DIELECTRIC, KAT_OMEGA_TRIG, KAT_SPECULAR_DIELECTRIC, KAT_REFLECT, KAT_REFRACT = 1, 2, 3, 4, 5
KAT_SPHERE_INTERSECT, KAT_MAKE_COORD, KAT_LOBE, KAT_BECKMANN, KAT_SPHERE_LIGHT = 6, 7, 8, 9, 10
KAT_ROUGHNESS_TO_ALPHA, KAT_CONCENTRIC, KAT_FRESNEL_CONDUCTOR, KAT_TRIANGLE, KAT_BBOX = 11, 12, 13, 14, 15
KAT_POWI, KAT_LUMINANCE, KAT_CATHETUS = 16, 17, 18


def kat(op, inputs, n_out=16):
    """Runs one known-answer hook; returns (outputs float32[n_out], would_panic_count)."""
    api = load()
    a = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1))
    pad = np.zeros(32, np.float32)
    pad[: a.size] = a
    out = np.zeros(max(n_out, 16), np.float32)
    rc = api["kat"](op, pad.ctypes.data_as(K.c_float_p), a.size, out.ctypes.data_as(K.c_float_p), out.size)
    if rc < 0:
        raise RuntimeError(f"oracle_kat({op}) failed: {api['last_error']().decode()}")
    return out[:n_out], rc


def trace_ray(handle, o, d, t_max=np.inf):
    api = load()
    oa = np.asarray(o, np.float32); da = np.asarray(d, np.float32); out = np.zeros(16, np.float32)
    rc = api["trace_ray"](handle.ptr, oa.ctypes.data_as(K.c_float_p), da.ctypes.data_as(K.c_float_p),
                          float(t_max), out.ctypes.data_as(K.c_float_p))
    assert rc == 0
    return out


def occludes_ray(handle, o, d, t_max=np.inf):
    api = load()
    oa = np.asarray(o, np.float32); da = np.asarray(d, np.float32)
    return api["occludes_ray"](handle.ptr, oa.ctypes.data_as(K.c_float_p), da.ctypes.data_as(K.c_float_p), float(t_max))


def render_rows(handle, row_step, **kw):
    """Bounded sample: renders rows crop_y, crop_y+row_step, ... only. Returns film, stats."""
    from pbrs_b200.scene import SceneHandle
    api = load()
    o = SceneHandle.make_opts(**kw)
    out = np.zeros((handle.height, handle.width, 3), np.float32)
    st = K.Stats()
    rc = api["render_rows"](handle.ptr, C.byref(o), out.ctypes.data_as(K.c_float_p), C.byref(st), int(row_step))
    if rc < 0:
        raise RuntimeError(api["last_error"]().decode())
    return out, st.as_dict()
