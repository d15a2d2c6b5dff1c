"""`python -m pbrs_b200 --pbrt_file scene.pbrt --integrator path --msaa 4`

The shape of the reference's driver (src/main.rs:56-246 with the knobs of src/cli_options.rs:52-59)
around the B200 back end: scene file -> pbrs_render -> "{scene}-{integrator}-{spp}spp.exr".
A convenience for trying scenes, not part of the hot path.  Needs a CUDA device (no CPU fallback).
"""
import argparse
import os
import time

from . import _ffi, film, scenes
from .pbrt_loader import load_pbrt

PRESETS = {"cornell_box": scenes.cornell_box, "spheres": scenes.spheres500, "terrain": scenes.mesh_terrain, "field": scenes.instanced_field}


def main():
    ap = argparse.ArgumentParser(prog="pbrs_b200")
    ap.add_argument("--pbrt_file")
    ap.add_argument("--scene_name", choices=sorted(PRESETS))
    ap.add_argument("--integrator", default="path", choices=["direct", "path"])
    ap.add_argument("--msaa", type=int, default=2)  # the reference's default, src/cli_options.rs:37-48
    ap.add_argument("--png", action="store_true", help="also write an 8-bit preview")
    args = ap.parse_args()
    if bool(args.pbrt_file) == bool(args.scene_name):
        ap.error("give exactly one of --pbrt_file / --scene_name")
    sd = load_pbrt(args.pbrt_file) if args.pbrt_file else PRESETS[args.scene_name]()
    name = os.path.splitext(os.path.basename(args.pbrt_file))[0] if args.pbrt_file else args.scene_name
    h = sd.realize(_ffi.load())
    t0 = time.time()
    img, st = h.render(integrator=args.integrator, msaa=args.msaa, max_depth=5)
    print(f"whole render time = {time.time() - t0:.3f}s ({st['ms_total']:.1f} ms on the GPU, {st['n_samples']} samples, "
          f"{st['n_rays_extend'] + st['n_rays_shadow']} rays)")
    out = film.exr_file_name(name, args.integrator, args.msaa)
    film.write_exr(out, img)
    if args.png:
        film.write_png(os.path.splitext(out)[0] + ".png", img)
    print(f"Image written to {out}")


if __name__ == "__main__":
    main()
