"""Loads pbrs_b200/lib/libpbrs_gpu.so (the C ABI of include/pbrs_gpu.h) and binds it with ctypes.

There is no fallback of any kind: if the CUDA library has not been built (see
``__graft_entry__.build()`` / ``pbrs_b200/csrc/Makefile``) loading raises, and every render entry
point returns PBRS_ERR_NO_DEVICE when no CUDA device is usable.
"""
import ctypes as C
import os
import subprocess

from . import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBRS_GPU_LIB: development knob to load an alternative build of the same library (kernel tuning)
LIB_PATH = os.environ.get("PBRS_GPU_LIB") or os.path.join(_HERE, "lib", "libpbrs_gpu.so")
CSRC = os.path.join(_HERE, "csrc")

_api = None
_lib = None


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "-s", "clean"])
    subprocess.check_call(["make", "-C", CSRC, "-s"])
    return LIB_PATH


def load():
    global _api, _lib
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
                "pbrs_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        table = dict(K.SCENE_API)
        table.update(K.PRODUCT_ONLY_API)
        _api = K.bind(_lib, "pbrs_", table)
    return _api
