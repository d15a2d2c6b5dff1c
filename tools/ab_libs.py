"""A/B harness for alternative builds of the CUDA library (kernel tuning).

    PYTHONPATH=. python tools/ab_libs.py "libA.so libB.so ..." "c3:1.0 c4:0.25 c5:0.125" [reps]

Builds each workload's scene description once, realises it through every library (each .so is an
independent copy of the product, bound with ctypes), renders warm frames with PBRS_FLAG_TIME_STAGES
and prints the per-stage times.  Every library's film must have the same CRC: the variants differ
in scheduling and instruction selection only, never in arithmetic.
Alternative builds: make -C pbrs_b200/csrc OUT=../lib/libX.so EXTRA="-DPBRS_...=v"
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

from pbrs_b200 import _capi as K
from pbrs_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    lib = C.CDLL(path)
    table = dict(K.SCENE_API)
    table.update({k: v for k, v in K.PRODUCT_ONLY_API.items() if hasattr(lib, "pbrs_" + k)})  # older builds lack the newest entry points
    return K.bind(lib, "pbrs_", table)


def main():
    libs = sys.argv[1].split()
    works = [w.split(":") for w in sys.argv[2].split()]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    # "libX.so@VAR=value": the same build with a development knob set while the scene is committed
    apis = {l: load(os.path.join(ROOT, "pbrs_b200", "lib", l.split("@")[0])) for l in libs}
    for name, scale in works:
        gen, integrator, msaa = scenes.CONFIGS[name]
        sd = gen(float(scale))
        ref_crc = None
        for l in libs:
            knob = l.split("@")[1].split("=") if "@" in l else None
            if knob:
                os.environ[knob[0]] = knob[1]
            h = sd.realize(apis[l])
            film, _ = h.render(integrator=integrator, msaa=msaa, want_stats=False)
            crc = zlib.crc32(film.tobytes())
            best = None
            for _ in range(reps):
                _, st = h.render(integrator=integrator, msaa=msaa, flags=2)
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = st
            _, plain = h.render(integrator=integrator, msaa=msaa)
            ok = "" if ref_crc in (None, crc) else "  FILM DIFFERS"
            ref_crc = ref_crc or crc
            if knob:
                del os.environ[knob[0]]
            print(f"{name}x{scale} {l:28s} total {plain['ms_total']:9.2f} ms | staged {best['ms_total']:9.2f}: extend {best['ms_extend']:9.2f} shade {best['ms_shade']:8.2f} "
                  f"shadow {best['ms_shadow']:9.2f} gen {best['ms_generate']:6.2f} acc {best['ms_accumulate']:6.2f} | crc {crc:08x}{ok}", flush=True)
            del h


if __name__ == "__main__":
    main()
