"""The C++ host side above the C ABI (include/pbrs_gpu.hpp, include/pbrs_scene_file.hpp,
pbrs_b200/csrc/host/pbrs_main.cpp = the shape of src/main.rs).  CPU: it builds, refuses to run
without a CUDA device, and its scene-file loader agrees bit for bit with the Python loader (both
driven through the host build of the stage functions).  GPU: the driver renders a scene file and
writes the same film as the Python path."""
import os
import subprocess

import numpy as np
import pytest

from pbrs_b200 import film, scenes
from pbrs_b200.pbrt_loader import load_pbrt
from tests.util import bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = os.path.join(ROOT, "pbrs_b200", "lib", "pbrs_main")
CHECK = os.path.join(ROOT, "tests", "hostsim", "scene_file_check")

MIXED = """
# nested transforms, named materials, every light kind the loader accepts
LookAt 0 2 -9  0 1 0  0 1 0
Camera "perspective" "float fov" [ 50 ]
Film "image" "integer xresolution" [ 80 ] "integer yresolution" [ 60 ]
Scale 1 1 1
WorldBegin
LightSource "point" "point from" [ 3 5 -4 ] "rgb L" [ 30 30 25 ]
LightSource "distant" "point from" [ 0 0 0 ] "point to" [ 0.3 -1 0.2 ] "rgb L" [ 0.7 0.7 0.8 ]
LightSource "infinite" "rgb L" [ 0.2 0.25 0.3 ]
MakeNamedMaterial "gold" "string type" "metal" "rgb eta" [ 0.143 0.373 1.444 ] "rgb k" [ 3.98 2.39 1.6 ] "float roughness" [ 0.05 ]
MakeNamedMaterial "glass" "string type" "glass" "float eta" [ 1.45 ]
Material "matte" "rgb Kd" [ 0.5 0.5 0.45 ]
Shape "trianglemesh" "point P" [ -8 0 -8  8 0 -8  8 0 8  -8 0 8 ] "integer indices" [ 0 1 2 0 2 3 ] "float uv" [ 0 0 1 0 1 1 0 1 ]
AttributeBegin
  NamedMaterial "gold"
  Translate -2 1 0
  Rotate 30 0 1 0
  Scale 1 1.5 0.75
  Shape "sphere" "float radius" [ 1 ]
  TransformBegin
    Translate 3 0 1
    Rotate -45 1 0 1
    Shape "trianglemesh" "point P" [ -1 -1 0  1 -1 0  0 1 0 ] "integer indices" [ 0 1 2 ] "normal N" [ 0 0 -1 0 0 -1 0 0 -1 ]
  TransformEnd
AttributeEnd
AttributeBegin
  NamedMaterial "glass"
  Translate 2 1 1
  Shape "sphere"
AttributeEnd
AttributeBegin
  Material "plastic" "rgb Kd" [ 0.2 0.4 0.7 ] "float roughness" 0.2 "string remaproughness" "false"
  Translate 0 0.5 -2
  Shape "sphere" "float radius" 0.5
AttributeEnd
AttributeBegin
  AreaLightSource "diffuse" "rgb L" [ 12 12 10 ]
  Translate 0 5 0
  Scale 0.5 0.5 0.5
  Shape "sphere" "float radius" [ 2 ]
AttributeEnd
WorldEnd
"""


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostsim"), "-s", "scene_file_check"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "pbrs_b200", "csrc"), "-s"])


def _cpp_ids(path, tmp_path):
    out = str(tmp_path / "ids.bin")
    r = subprocess.run([CHECK, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, np.uint32)
    w, h = int(raw[0]), int(raw[1])
    n = w * h
    return raw[:8], raw[8:8 + n].reshape(h, w), raw[8 + n:8 + 2 * n].reshape(h, w), raw[8 + 2 * n:8 + 3 * n].view(np.float32).reshape(h, w)


@pytest.mark.parametrize("which", ["cornell", "mixed", "ply"])
def test_cpp_scene_file_loader_equals_python_loader(tmp_path, hostsim_api, which):
    _build()
    path = str(tmp_path / (which + ".pbrt"))
    if which == "ply":   # plymesh shapes and PLY area lights, scene/src/plyloader.rs
        from tests.test_scene_io import write_ply_scene
        path = write_ply_scene(str(tmp_path))
    else:
        open(path, "w").write(scenes.cornell_box_pbrt(96, 72) if which == "cornell" else MIXED)
    hdr, inst, prim, t = _cpp_ids(path, tmp_path)
    h = load_pbrt(path).realize(hostsim_api)
    info = h.info()
    assert list(hdr[:7]) == [info.width, info.height, info.n_instances, info.n_meshes, info.n_spheres, info.n_triangles, info.n_lights]
    a = h.render_ids(0, msaa=1, flags=4)
    assert (inst == a[0]).all() and (prim == a[1]).all() and bits_equal(t, a[2]).all()
    assert (a[0] != 0xFFFFFFFF).mean() > 0.3


@pytest.mark.parametrize("name,py", [("cornell_box", lambda: scenes.preset_cornell_box()), ("quad", lambda: scenes.preset_quad()),
                                     ("plates", lambda: scenes.preset_plates())])
def test_cpp_presets_equal_python_presets(tmp_path, hostsim_api, name, py):
    """scene/src/preset.rs written twice (include/pbrs_presets.hpp over the C++ constructors,
    pbrs_b200/scenes.py over SceneDesc): same scene facts, bit-identical primary hits.  Covers the
    ParallelQuad::new_* / Cuboid::from_points / with_transform mirrors of the C++ header."""
    _build()
    hdr, inst, prim, t = _cpp_ids("preset:" + name, tmp_path)
    h = py().realize(hostsim_api)
    info = h.info()
    assert list(hdr[:7]) == [info.width, info.height, info.n_instances, info.n_meshes, info.n_spheres, info.n_triangles, info.n_lights]
    a = h.render_ids(0, msaa=1, flags=4)
    assert (inst == a[0]).all() and (prim == a[1]).all() and bits_equal(t, a[2]).all()
    assert (a[0] != 0xFFFFFFFF).mean() > 0.1


@pytest.mark.parametrize("seed", range(8))
def test_random_scene_files_cpp_equals_python(tmp_path, hostsim_api, seed):
    """Random pbrt-subset files (nested blocks, every material, random PLY meshes and PLY area lights;
    generator of tools/soak_loaders.py, which ran 1 400 of them): both loaders accept the file and
    produce bit-identical primary hits."""
    from tools.soak_loaders import scene_text
    _build()
    path = str(tmp_path / "s.pbrt")
    open(path, "w").write(scene_text(500 + seed, str(tmp_path)))
    hdr, inst, prim, t = _cpp_ids(path, tmp_path)
    h = load_pbrt(path).realize(hostsim_api)
    info = h.info()
    assert list(hdr[:7]) == [info.width, info.height, info.n_instances, info.n_meshes, info.n_spheres, info.n_triangles, info.n_lights]
    a = h.render_ids(0, msaa=1, flags=4)
    assert (inst == a[0]).all() and (prim == a[1]).all() and bits_equal(t, a[2]).all()


def test_cpp_loader_rejects_what_the_reference_cannot_load(tmp_path):
    _build()
    path = str(tmp_path / "bad.pbrt")
    open(path, "w").write('Camera "perspective" Film "image" "integer xresolution" [ 8 ] "integer yresolution" [ 8 ] WorldBegin Material "matte" '
                          'Shape "loopsubdiv" WorldEnd')
    r = subprocess.run([CHECK, path, str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 5 and "loopsubdiv" in r.stderr
    # an ascii PLY: bytes_to_f32 panics upstream
    open(str(tmp_path / "a.ply"), "w").write("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                                             "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    open(path, "w").write('Camera "perspective" Film "image" "integer xresolution" [ 8 ] "integer yresolution" [ 8 ] WorldBegin Material "matte" '
                          'Shape "plymesh" "string filename" "a.ply" WorldEnd')
    r = subprocess.run([CHECK, path, str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 5 and "ascii" in r.stderr


def test_cpp_driver_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _build()
    r = subprocess.run([MAIN, "--scene_name", "cornell_box", "--msaa", "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr
    r = subprocess.run([MAIN, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 2 and "Unrecognized key" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_renders_a_scene_file_like_the_python_path(tmp_path, gpu_api):
    path = str(tmp_path / "cornell.pbrt")
    open(path, "w").write(scenes.cornell_box_pbrt(160, 120))
    r = subprocess.run([MAIN, "--pbrt_file", path, "--integrator", "path", "--msaa", "2"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "cornell-path-4spp.exr")  # src/main.rs:238-243
    assert "Image written to cornell-path-4spp.exr" in r.stdout and os.path.exists(out)
    img = film.read_exr(out)
    want, _ = load_pbrt(path).realize(gpu_api).render(integrator="path", msaa=2)
    assert bits_equal(img, want).all()


def test_cpp_ply_loader_rejects_counts_that_would_wrap(tmp_path):
    """A header whose vertex count makes `count * stride * 4` wrap in 64 bits, or whose count has more
    digits than a long holds, is a load error of the C++ loader -- not a crash."""
    import re
    _build()
    from tests.test_scene_io import write_ply_scene
    path = write_ply_scene(str(tmp_path))
    ball = str(tmp_path / "ball.ply")
    raw = open(ball, "rb").read()
    for bad in (re.sub(rb"element vertex \d+", b"element vertex 4611686018427387904", raw),
                re.sub(rb"element face \d+", b"element face 99999999999999999999", raw)):
        open(ball, "wb").write(bad)
        r = subprocess.run([CHECK, path, str(tmp_path / "ids.bin")], capture_output=True, text=True)
        assert r.returncode not in (0, -11, -6, 134, 139), (r.returncode, r.stderr)
        assert "ply" in (r.stderr + r.stdout).lower()
