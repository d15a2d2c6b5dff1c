"""pbrs_b200: B200 (sm_100a) back end for the pbrs path-tracing inner loop.

Host-side mirror of the reference's scene-construction vocabulary on top of the C ABI in
include/pbrs_gpu.h (library: pbrs_b200/lib/libpbrs_gpu.so, built by __graft_entry__.build()).
There is no CPU fallback: loading fails loudly when the CUDA library is missing.
"""
from . import _capi  # noqa: F401
from .scene import PbrsError, SceneDesc, SceneHandle  # noqa: F401
