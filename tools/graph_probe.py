"""Development probe: per-call wall time of a small frame with and without CUDA-graph replay,
through pbrs_render (host film, legacy stream) and pbrs_render_device (caller stream)."""
import ctypes as C
import time

import numpy as np
import torch

from pbrs_b200 import _capi as K
from pbrs_b200 import _ffi, scenes

api = _ffi.load()
h = scenes.cornell_box_via_parser(1920, 1080).realize(api)
kw = dict(integrator="direct", msaa=1)
host = np.zeros((h.height, h.width, 3), np.float32)
hp = host.ctypes.data_as(K.c_float_p)
film = torch.empty((h.height, h.width, 3), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
for name, flags in (("graph", 0), ("direct", K.FLAG_NO_GRAPH), ("graph", 0), ("direct", K.FLAG_NO_GRAPH)):
    o = h.make_opts(flags=flags, **kw)
    for _ in range(5):
        api["render"](h.ptr, C.byref(o), hp, None)
    t0 = time.perf_counter()
    for _ in range(50):
        api["render"](h.ptr, C.byref(o), hp, None)
    t_host = (time.perf_counter() - t0) / 50
    for _ in range(5):
        h.render_device(film.data_ptr(), stream=s.cuda_stream, want_stats=False, flags=flags, **kw)
    s.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        h.render_device(film.data_ptr(), stream=s.cuda_stream, want_stats=False, flags=flags, **kw)
    s.synchronize()
    t_dev = (time.perf_counter() - t0) / 50
    t0 = time.perf_counter()
    for _ in range(50):
        h.render_device(film.data_ptr(), stream=s.cuda_stream, want_stats=False, flags=flags, **kw)
        s.synchronize()
    t_dev_sync = (time.perf_counter() - t0) / 50
    print(f"{name:7s} pbrs_render(host film) {t_host * 1e3:.3f} ms   render_device back-to-back {t_dev * 1e3:.3f} ms   render_device + sync each {t_dev_sync * 1e3:.3f} ms")

# transient after a key change: per-call wall times of the first calls
for name, flags in (("graph", 0), ("direct", K.FLAG_NO_GRAPH)):
    o = h.make_opts(flags=flags, integrator="direct", msaa=1, seed=1234 + flags)
    ts = []
    for _ in range(14):
        t0 = time.perf_counter()
        api["render"](h.ptr, C.byref(o), hp, None)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "first calls (ms):", " ".join(f"{t:.2f}" for t in ts))
