"""CPU soak: random pbrt-subset scene files through the Python loader (pbrs_b200/pbrt_loader.py)
and the C++ loader (include/pbrs_scene_file.hpp, via tests/hostsim/scene_file_check): same scene
facts and bit-identical primary hits.  usage: python tools/soak_loaders.py <first_seed> <last_seed>"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrs_b200.pbrt_loader import load_pbrt  # noqa: E402
from tests.hostsim import load as hs_load  # noqa: E402

CHECK = os.path.join(ROOT, "tests", "hostsim", "scene_file_check")


def num(rng, lo, hi):
    return f"{rng.uniform(lo, hi):.4f}"   # no exponent floats: the reference's lexer has none


def rgb(rng, lo=0.05, hi=0.95):
    return " ".join(num(rng, lo, hi) for _ in range(3))


def material(rng):
    k = rng.integers(0, 7)
    if k == 0: return f'Material "matte" "rgb Kd" [ {rgb(rng)} ]'
    if k == 1: return f'Material "glass" "float eta" [ {num(rng, 1.2, 1.8)} ]'
    if k == 2: return f'Material "mirror" "rgb Kr" [ {rgb(rng, 0.5, 1.0)} ]'
    if k == 3: return f'Material "metal" "rgb eta" [ {rgb(rng, 0.1, 1.5)} ] "rgb k" [ {rgb(rng, 2.0, 5.0)} ] "float roughness" [ {num(rng, 0.01, 0.4)} ]'
    if k == 4: return f'Material "plastic" "rgb Kd" [ {rgb(rng)} ] "rgb Ks" [ {rgb(rng)} ] "float roughness" {num(rng, 0.02, 0.5)}'
    if k == 5: return f'Material "uber" "rgb Kd" [ {rgb(rng)} ] "rgb Ks" [ {rgb(rng)} ] "float roughness" [ {num(rng, 0.02, 0.5)} ] "float eta" [ {num(rng, 1.2, 1.8)} ]'
    return f'Material "substrate" "rgb Kd" [ {rgb(rng)} ] "rgb Ks" [ {rgb(rng, 0.02, 0.2)} ]'


def transform(rng):
    k = rng.integers(0, 3)
    if k == 0: return f"Translate {num(rng, -2, 2)} {num(rng, -1, 2)} {num(rng, -2, 2)}"
    if k == 1: return f"Rotate {num(rng, -180, 180)} {num(rng, -1, 1)} {num(rng, -1, 1)} {num(rng, 0.1, 1)}"
    return f"Scale {num(rng, 0.5, 1.6)} {num(rng, 0.5, 1.6)} {num(rng, 0.5, 1.6)}"


def ply_shape(rng, tmp, tag):
    """A random binary PLY (either endianness, any index width, polygons, optional normals / uv)."""
    from pbrs_b200.pbrt_loader import write_ply
    nv = int(rng.integers(4, 12))
    P = rng.uniform(-1.5, 1.5, (nv, 3)).astype(np.float32)
    faces = []
    for _ in range(int(rng.integers(1, 6))):
        k = int(rng.integers(3, min(6, nv + 1)))
        faces.append([int(v) for v in rng.choice(nv, size=k, replace=False)])
    used = sorted({v for f in faces for v in f})
    faces.append(list(range(nv))[:3])
    for v in range(nv):   # every vertex in some face: compute_normals panics upstream otherwise
        if v not in used:
            faces.append([v, (v + 1) % nv, (v + 2) % nv])
    N = rng.normal(size=(nv, 3)).astype(np.float32) if rng.random() < 0.4 else None
    UV = rng.random((nv, 2)).astype(np.float32) if rng.random() < 0.5 else None
    name = f"m{tag}.ply"
    write_ply(os.path.join(tmp, name), P, None, N=N, UV=UV, big_endian=bool(rng.random() < 0.5),
              index_type=str(rng.choice(["uchar", "short", "int", "uint"])), polygons=faces)
    return f'Shape "plymesh" "string filename" "{name}"'


def shape(rng, tmp=None, tag=0):
    if tmp is not None and rng.random() < 0.25:
        return ply_shape(rng, tmp, tag)
    if rng.random() < 0.5:
        return f'Shape "sphere" "float radius" [ {num(rng, 0.3, 1.0)} ]'
    n = int(rng.integers(1, 5))
    P = " ".join(num(rng, -1.5, 1.5) for _ in range(9 * n))
    idx = " ".join(str(i) for i in range(3 * n))
    extra = ""
    if rng.random() < 0.5: extra += ' "normal N" [ ' + " ".join(num(rng, -1, 1) for _ in range(9 * n)) + " ]"
    if rng.random() < 0.5: extra += ' "float uv" [ ' + " ".join(num(rng, 0, 1) for _ in range(6 * n)) + " ]"
    return f'Shape "trianglemesh" "point P" [ {P} ] "integer indices" [ {idx} ]{extra}'


def block(rng, depth=0, tmp=None):
    out = []
    for _ in range(int(rng.integers(1, 4))):
        r = rng.random()
        if r < 0.25 and depth < 3:
            kind = "Attribute" if rng.random() < 0.6 else "Transform"
            inner = block(rng, depth + 1, tmp)
            if kind == "Attribute": inner.insert(0, material(rng))
            out += [kind + "Begin"] + ["  " + l for l in inner] + [kind + "End"]
        elif r < 0.5:
            out.append(transform(rng))
        elif r < 0.6:
            out.append(material(rng))
        else:
            out.append(shape(rng, tmp, int(rng.integers(0, 1 << 30))))
    return out


def scene_text(seed, tmp=None):
    rng = np.random.default_rng(seed)
    lines = [f"LookAt {num(rng, -1, 1)} {num(rng, 0.5, 2.5)} -8  0 0.5 0  0 1 0",
             f'Camera "perspective" "float fov" [ {num(rng, 35, 65)} ]',
             'Film "image" "integer xresolution" [ 64 ] "integer yresolution" [ 48 ]']
    if rng.random() < 0.3: lines.append(transform(rng))
    lines.append("WorldBegin")
    lines.append(f'LightSource "point" "point from" [ 3 5 -4 ] "rgb L" [ {rgb(rng, 5, 40)} ]')
    if rng.random() < 0.5: lines.append(f'LightSource "distant" "point from" [ 0 0 0 ] "point to" [ {num(rng, -1, 1)} -1 {num(rng, -1, 1)} ] "rgb L" [ {rgb(rng)} ]')
    if rng.random() < 0.5: lines.append(f'LightSource "infinite" "rgb L" [ {rgb(rng, 0.05, 0.4)} ]')
    lines.append(material(rng))
    lines.append('Shape "trianglemesh" "point P" [ -8 -1 -8  8 -1 -8  8 -1 8  -8 -1 8 ] "integer indices" [ 0 1 2 0 2 3 ]')
    lines += block(rng, 0, tmp)
    if rng.random() < 0.6:
        lamp = ply_shape(rng, tmp, 999) if (tmp is not None and rng.random() < 0.4) else f'Shape "sphere" "float radius" [ {num(rng, 0.2, 0.7)} ]'
        lines += ["AttributeBegin", f'  AreaLightSource "diffuse" "rgb L" [ {rgb(rng, 5, 20)} ]', f"  Translate {num(rng, -2, 2)} 4 {num(rng, -2, 2)}",
                  "  " + lamp, "AttributeEnd"]
    lines.append("WorldEnd")
    return "\n".join(lines) + "\n"


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostsim"), "-s", "scene_file_check"])
    api = hs_load()
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    bad = loaded = 0
    tmp = tempfile.mkdtemp()
    for seed in range(lo, hi):
        path = os.path.join(tmp, f"s{seed}.pbrt")
        open(path, "w").write(scene_text(seed, tmp))
        out = os.path.join(tmp, "ids.bin")
        r = subprocess.run([CHECK, path, out], capture_output=True, text=True)
        try:
            h = load_pbrt(path).realize(api)
            py_ok = True
        except Exception as e:   # both loaders must refuse the same files
            py_ok = False
            py_err = repr(e)[:120]
        if (r.returncode == 0) != py_ok:
            bad += 1
            print("FAIL", seed, "C++ rc", r.returncode, r.stderr.strip()[:120], "| python", "ok" if py_ok else py_err, flush=True)
            continue
        if not py_ok:
            continue
        loaded += 1
        raw = np.fromfile(out, np.uint32)
        n = int(raw[0]) * int(raw[1])
        info = h.info()
        a = h.render_ids(0, msaa=1, flags=4)
        same = (list(raw[:7]) == [info.width, info.height, info.n_instances, info.n_meshes, info.n_spheres, info.n_triangles, info.n_lights]
                and (raw[8:8 + n].reshape(a[0].shape) == a[0]).all() and (raw[8 + n:8 + 2 * n].reshape(a[1].shape) == a[1]).all()
                and (raw[8 + 2 * n:8 + 3 * n].reshape(a[2].shape) == a[2].view(np.uint32)).all())
        if not same:
            bad += 1
            print("FAIL", seed, "scene facts or ids differ", list(raw[:7]), flush=True)
    print("done", lo, hi, "failures", bad, "| loaded by both", loaded, "refused by both", hi - lo - loaded - bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
