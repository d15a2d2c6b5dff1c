"""One process, N GPUs: times pbrs_render with num_gpus = 1 and N on a workload (page-locked film).
usage: PYTHONPATH=. python tools/multi_gpu_probe.py c4 1.0 N [reps]"""
import ctypes as C
import sys
import time
import zlib

import numpy as np

from pbrs_b200 import _capi as K
from pbrs_b200 import _ffi, scenes
from pbrs_b200.dist import split_for

name, scale, n = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
api = _ffi.load()
gen, integrator, msaa = scenes.CONFIGS[name]
h = gen(scale).realize(api)
p = api["film_alloc"](h.width, h.height)
film = np.ctypeslib.as_array(C.cast(p, K.c_float_p), shape=(h.height, h.width, 3))
split = split_for(name)
for g in (1, n):
    h.render(integrator=integrator, msaa=msaa, num_gpus=g, split=split, out=film, want_stats=False)  # replicates + warms up
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        h.render(integrator=integrator, msaa=msaa, num_gpus=g, split=split, out=film, want_stats=False)
        best = min(best, time.perf_counter() - t0)
    n_samples = h.width * h.height * msaa * msaa
    print(f"{name}x{scale} num_gpus={g} split={split}: {best * 1e3:.2f} ms/frame, {n_samples / best / 1e6:.1f} Msamples/s e2e (host film), crc {zlib.crc32(film.tobytes()):08x}", flush=True)
