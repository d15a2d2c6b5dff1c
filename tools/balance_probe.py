"""Load balance of the tile split, measured on ONE GPU: every rank's share of an N-way split is rendered
on its own and timed (device time of the frame, events around the launches); the frame time on N GPUs
is the slowest share.  usage: PYTHONPATH=. python tools/balance_probe.py c4 1.0 8 [reps]"""
import sys

from pbrs_b200 import scenes
from pbrs_b200._ffi import load

name, scale, world = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
api = load()
gen, integrator, msaa = scenes.CONFIGS[name]
h = gen(scale).realize(api)
_, whole = h.render(integrator=integrator, msaa=msaa)
_, whole = h.render(integrator=integrator, msaa=msaa)
ms = []
for r in range(world):
    best = None
    for _ in range(reps):
        _, st = h.render(integrator=integrator, msaa=msaa, rank=r, world_size=world, split="tiles")
        best = st["ms_total"] if best is None else min(best, st["ms_total"])
    ms.append(best)
mean = sum(ms) / world
print(f"{name}x{scale} N={world}: whole frame {whole['ms_total']:.1f} ms, shares {' '.join(f'{m:.1f}' for m in ms)} ms")
print(f"  slowest / mean = {max(ms) / mean:.4f}, sum of shares / whole = {sum(ms) / whole['ms_total']:.4f}, efficiency whole / (N * slowest) = {whole['ms_total'] / (world * max(ms)):.4f}")
