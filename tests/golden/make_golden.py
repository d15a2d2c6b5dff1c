"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference is Rust and cannot run here, so the
fixtures pin the ORACLE (guarding it against drift) and give the GPU tests a target that does not
need the oracle at run time.  Each file: primary-hit ids of sample 0 and per-sample radiance of
the path integrator at depth 1 and of the direct-lighting integrator, msaa 2, seed 0x5EED."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi  # noqa: E402
from tests.util import SMALL_SCENES  # noqa: E402

GOLDEN_SCENES = ["cornell", "spheres", "terrain", "field", "zoo_image", "shape_zoo", "preset_cornell"]
CROP = (16, 12, 48, 36)  # x, y, w, h


def main(only=None):
    api = oracle_ffi.load()
    out = os.path.dirname(os.path.abspath(__file__))
    for name in (only or GOLDEN_SCENES):
        h = SMALL_SCENES[name]().realize(api)
        inst, prim, t = h.render_ids(0, msaa=2, crop=CROP)
        d1, _ = h.render_samples(integrator="path", msaa=2, max_depth=1, crop=CROP)
        dl, _ = h.render_samples(integrator="direct", msaa=2, max_depth=5, crop=CROP)
        np.savez_compressed(os.path.join(out, name + ".npz"), inst=inst, prim=prim, t=t, path_depth1=d1, direct=dl,
                            crop=np.array(CROP, np.int32))
        print(name, "hits", int((inst != 0xFFFFFFFF).sum()), "mean radiance", float(d1.mean()), float(dl.mean()))


if __name__ == "__main__":
    main(sys.argv[1:])  # optional scene names: regenerate only those
