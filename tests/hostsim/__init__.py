"""TEST TOOLING ONLY: ctypes loader of tests/hostsim/libhostsim.so (see hostsim.cpp)."""
import ctypes as C
import os
import subprocess

from pbrs_b200 import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libhostsim.so")
_api = None
_lib = None


def load():
    global _api, _lib
    if _api is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(_LIB)
        scene = {k: v for k, v in K.SCENE_API.items() if not k.startswith("render")}
        _api = K.bind(_lib, "pbrs_", scene)
        _api.update(K.bind(_lib, "hostsim_", {k: v for k, v in K.SCENE_API.items() if k.startswith("render")}))
    return _api
