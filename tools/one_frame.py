"""One warm-up frame + one frame of a workload through a given build of the library (ncu target).
usage: PYTHONPATH=. python tools/one_frame.py libX.so c4 0.25 [frames]"""
import os
import sys

from pbrs_b200 import scenes
from tools.ab_libs import ROOT, load

lib, name, scale = sys.argv[1], sys.argv[2], float(sys.argv[3])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1
api = load(os.path.join(ROOT, "pbrs_b200", "lib", lib))
gen, integrator, msaa = scenes.CONFIGS[name]
h = gen(scale).realize(api)
for _ in range(frames):
    _, st = h.render(integrator=integrator, msaa=msaa)
print(lib, name, scale, "ms", round(st["ms_total"], 2), "launches", st["launches"])
