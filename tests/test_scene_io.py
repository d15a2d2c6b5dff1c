"""Either side of the hot path: pbrt-subset scene files in (pbrs_b200/pbrt_loader.py, mirroring
scene_parser + scene/src/loader.rs) and the EXR film out (pbrs_b200/film.py, src/main.rs:42-53)."""
import os
import re

import numpy as np
import pytest

from pbrs_b200 import _capi as K
from pbrs_b200 import film, pbrt_loader, scenes
from pbrs_b200.pbrt_loader import PbrtError, load_pbrt, load_pbrt_string, tokenize
from tests.util import bits_equal


def test_lexer_token_set_comments_and_floats():
    toks = tokenize('# a comment\nLookAt 0 .5 -1.25 +3 4. [ "x" ] WorldBegin')
    assert toks[0] == ("kw", "LookAt")
    assert [float(t[1]) for t in toks[1:6]] == [0.0, 0.5, -1.25, 3.0, 4.0]
    assert toks[6:9] == [("[",), ("str", "x"), ("]",)]
    # scene_parser/src/token.rs:112-114: no exponent floats -- "1e3" is `1` then an error token
    with pytest.raises(PbrtError):
        tokenize("Translate 1e3 0 0")
    with pytest.raises(PbrtError):
        tokenize("NoSuchDirective 1 2 3")


def test_include_splices_tokens(tmp_path):
    (tmp_path / "geo.pbrt").write_text('Shape "sphere" "float radius" [ 2 ]\n')
    main = ('LookAt 0 0 -5 0 0 0 0 1 0 Camera "perspective" "float fov" [ 45 ] Film "image" "integer xresolution" [ 32 ] '
            '"integer yresolution" [ 24 ] WorldBegin Material "matte" Include "geo.pbrt" WorldEnd')
    (tmp_path / "main.pbrt").write_text(main)
    sd = load_pbrt(str(tmp_path / "main.pbrt"))
    assert (sd.width, sd.height, sd.n_shape, sd.n_inst) == (32, 24, 1, 1)


def test_parameter_lists_and_defaults():
    text = ('Camera "perspective" Film "image" "integer xresolution" [ 16 ] "integer yresolution" [ 8 ] WorldBegin '
            'Material "metal" "float roughness" 0.2 Shape "sphere" '
            'Material "uber" "rgb Kd" [ .1 .2 .3 ] "float uroughness" [ .1 ] "float vroughness" [ .2 ] "float eta" [ 1.3 ] Shape "sphere" "float radius" 3 '
            'Material "glass" Shape "sphere" WorldEnd')
    sd = load_pbrt_string(text)
    mats = [a for n, a in sd.ops if n == "scene_add_material"]
    metal, uber, glass = mats
    assert metal[0] == K.MTL_METAL and metal[4][0] == pytest.approx(0.2)
    assert metal[2] == pytest.approx(pbrt_loader.COPPER_ETA) and metal[3] == pytest.approx(pbrt_loader.COPPER_ETA)  # Q16: k defaults to copper ETA
    assert uber[0] == K.MTL_UBER and uber[4] == pytest.approx((0.1, 0.2, 1.3, 1.0))  # opacity is always 1 (Q16)
    assert glass[0] == K.MTL_DIELECTRIC and glass[4][0] == pytest.approx(1.5)
    cam = [a for n, a in sd.ops if n == "scene_set_camera"][0]
    assert cam[2] == pytest.approx(np.deg2rad(60.0))  # default fov 60 (loader.rs:106)
    spheres = [a for n, a in sd.ops if n == "scene_add_sphere"]
    assert [s[1] for s in spheres] == [1.0, 3.0, 1.0]


def test_xyz_colours_go_through_from_xyz():
    # parse_constant_color "xyz" -> Color::from_xyz (radiometry/src/color.rs:30-36); D65 white -> ~(1,1,1)
    c = pbrt_loader.Loader.constant_color("xyz", [0.95047, 1.0, 1.08883])
    np.testing.assert_allclose(c, (1.0, 1.0, 1.0), atol=2e-4)
    c = pbrt_loader.Loader.constant_color("xyz", [1.0, 0.0, 0.0])
    np.testing.assert_allclose(c, (3.240479, -0.969256, 0.055648), rtol=1e-6)
    with pytest.raises(PbrtError):
        pbrt_loader.Loader.constant_color("blackbody", [6500.0, 1.0])


def test_attribute_blocks_reset_material_and_scope_transforms():
    text = ('Camera "perspective" Film "image" "integer xresolution" [ 16 ] "integer yresolution" [ 8 ] WorldBegin '
            'Material "matte" AttributeBegin Shape "sphere" AttributeEnd '          # material reset inside the block: dropped
            'AttributeBegin Material "mirror" Translate 1 2 3 Scale 2 2 2 Shape "sphere" AttributeEnd '
            'Material "matte" Shape "sphere" WorldEnd')
    sd = load_pbrt_string(text)
    inst = [a for n, a in sd.ops if n == "scene_add_instance"]
    assert len(inst) == 2
    fwd, inv = inst[0][2], inst[0][3]
    np.testing.assert_allclose(fwd, [[2, 0, 0, 1], [0, 2, 0, 2], [0, 0, 2, 3], [0, 0, 0, 1]])
    np.testing.assert_allclose(fwd @ inv, np.eye(4), atol=1e-6)
    assert inst[1][2] is None  # identity transform outside the block


def test_rotate_uses_the_negated_angle():
    # scene/src/loader.rs:792-798
    a = pbrt_loader.Loader.transform_of(("Rotate", [0.0, 1.0, 0.0], pbrt_loader._to_radians(90.0)))
    v = a.apply_vec((1.0, 0.0, 0.0))
    # Mat4::rotater(Y, t): X -> X cos t + (X x Y) sin t = X cos t + Z sin t (math/src/hcm.rs:508-520); t = -90 deg
    np.testing.assert_allclose(v, [0.0, 0.0, -1.0], atol=1e-6)
    np.testing.assert_allclose(a.fwd @ a.inv, np.eye(4), atol=1e-6)


def test_unsupported_directives_fail_loudly():
    base = 'Camera "perspective" Film "image" "integer xresolution" [ 16 ] "integer yresolution" [ 8 ] WorldBegin Material "matte" %s WorldEnd'
    for body in ['Shape "plymesh"', 'Shape "loopsubdiv"', 'ObjectBegin "x" ObjectEnd', 'Material "fourier"',
                 'LightSource "spot"', 'AreaLightSource "diffuse" "rgb L" [ 1 1 1 ] Shape "trianglemesh"']:
        with pytest.raises(PbrtError):
            load_pbrt_string(base % body)
    with pytest.raises(PbrtError):
        load_pbrt_string('WorldBegin WorldEnd')  # no camera / film


PLY_SCENE = """
LookAt 0 1.5 -6  0 0.3 0  0 1 0
Camera "perspective" "float fov" [ 45 ]
Film "image" "integer xresolution" [ 80 ] "integer yresolution" [ 60 ]
WorldBegin
LightSource "point" "point from" [ 3 5 -4 ] "rgb L" [ 30 30 25 ]
Material "matte" "rgb Kd" [ 0.5 0.5 0.45 ]
Shape "plymesh" "string filename" "floor.ply"
AttributeBegin
  Material "plastic" "rgb Kd" [ 0.2 0.4 0.7 ] "float roughness" 0.2
  Translate -1.2 0.9 0
  Rotate 25 0 1 0
  Shape "plymesh" "string filename" "ball.ply"
AttributeEnd
AttributeBegin
  AreaLightSource "diffuse" "rgb L" [ 12 12 10 ]
  Translate 0.5 3 0.5
  Rotate 10 1 0 0
  Shape "plymesh" "string filename" "lamp.ply"
AttributeEnd
WorldEnd
"""


def write_ply_scene(root):
    """floor.ply: a polygon file (fan triangulation, big endian, short indices, uv);
    ball.ply: an icosphere without normals (compute_normals); lamp.ply: two emissive triangles."""
    P, N, UV, idx = scenes.icosphere(1, radius=0.9)
    pbrt_loader.write_ply(os.path.join(root, "ball.ply"), P, idx)
    fl = np.array([[-5, 0, -5], [5, 0, -5], [5, 0, 5], [0, 0, 7], [-5, 0, 5]], np.float32)
    pbrt_loader.write_ply(os.path.join(root, "floor.ply"), fl, None, UV=fl[:, [0, 2]] * 0.1, big_endian=True, index_type="short",
                          polygons=[[0, 1, 2, 3, 4]])
    lamp = np.array([[-0.5, 0, -0.5], [0.5, 0, -0.5], [0.5, 0, 0.5], [-0.5, 0, 0.5]], np.float32)
    pbrt_loader.write_ply(os.path.join(root, "lamp.ply"), lamp, np.array([[0, 1, 2], [0, 2, 3]]), N=np.tile([[0, -1, 0]], (4, 1)), index_type="uchar")
    path = os.path.join(root, "ply_scene.pbrt")
    open(path, "w").write(PLY_SCENE)
    return path


def test_ply_reader_matches_the_reference_reader(tmp_path):
    # scene/src/plyloader.rs:69-256: binary LE/BE, float properties in any order, uchar/short/int
    # index lists, polygons fanned; normals from geometry::compute_normals when the file has none
    P, N, UV, idx = scenes.icosphere(2)
    a = str(tmp_path / "a.ply")
    pbrt_loader.write_ply(a, P, idx)
    P2, N2, UV2, idx2 = pbrt_loader.load_ply(a)
    assert bits_equal(P, P2).all() and (idx == idx2).all() and (UV2 == 0).all()
    # compute_normals restated as the scalar loop of geometry/src/lib.rs:16-32
    acc = np.zeros_like(P)
    for i, j, k in idx:
        e1, e2 = P[j] - P[i], P[k] - P[i]
        n = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]], np.float32)
        acc[i] += n; acc[j] += n; acc[k] += n
    n2 = (acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1]) + acc[:, 2] * acc[:, 2]
    want = acc * (np.float32(1.0) / np.sqrt(n2))[:, None]
    assert bits_equal(N2, want.astype(np.float32)).all()
    b = str(tmp_path / "b.ply")
    pbrt_loader.write_ply(b, P, None, N=N, UV=UV, big_endian=True, index_type="short", polygons=[[0, 1, 2, 3], [4, 5, 6], [1, 2, 3, 4, 5]])
    P3, N3, UV3, idx3 = pbrt_loader.load_ply(b)
    assert bits_equal(P, P3).all() and bits_equal(N.astype(np.float32), N3).all() and bits_equal(UV.astype(np.float32), UV3).all()
    assert idx3.tolist() == [[0, 1, 2], [0, 2, 3], [4, 5, 6], [1, 2, 3], [1, 3, 4], [1, 4, 5]]
    # what the reference cannot read
    raw = open(a, "rb").read()
    for bad in (raw.replace(b"binary_little_endian", b"ascii"), raw.replace(b"ply\n", b"plx\n"), raw[:200],
                raw.replace(b"property float z", b"property uchar z"), raw.replace(b"element face", b"element fac"),
                # counts that would wrap the size arithmetic of a 64-bit loader
                re.sub(rb"element vertex \d+", b"element vertex 4611686018427387904", raw),
                re.sub(rb"element face \d+", b"element face 99999999999999999999", raw)):
        c = str(tmp_path / "c.ply")
        open(c, "wb").write(bad)
        with pytest.raises(PbrtError):
            pbrt_loader.load_ply(c)
    # a vertex no face uses: Vec3::hat(0) panics in compute_normals upstream
    pbrt_loader.write_ply(a, np.concatenate([P, [[9, 9, 9]]]).astype(np.float32), idx)
    with pytest.raises(PbrtError):
        pbrt_loader.load_ply(a)


def test_plymesh_scene_equals_constructor_scene(tmp_path, oracle_api, hostsim_api):
    """`Shape "plymesh"` under a material = TriangleMesh::build_from_raw of the file; under an
    AreaLightSource = one emissive triangle instance + one triangle area light per face
    (scene/src/loader.rs:314-331,408-433)."""
    from pbrs_b200.pbrt_loader import Affine, _to_radians
    path = write_ply_scene(str(tmp_path))
    sd = load_pbrt(path)
    h = sd.realize(hostsim_api)
    info = h.info()
    assert (info.n_instances, info.n_meshes, info.n_triangles, info.n_lights) == (4, 2, 3 + 80, 3)
    # the same scene through the constructors
    from pbrs_b200.scene import SceneDesc
    ref = SceneDesc()
    ref.set_camera(80, 60, 45.0, (0, 1.5, -6), (0, 0.3, 0), (0, 1, 0))
    ref.add_point_light((3, 5, -4), (30, 30, 25))
    P, N, UV, idx = pbrt_loader.load_ply(str(tmp_path / "floor.ply"))
    ref.add_instance(ref.add_mesh(P, idx, N=N, UV=UV), ref.lambertian((0.5, 0.5, 0.45)))
    P, N, UV, idx = pbrt_loader.load_ply(str(tmp_path / "ball.ply"))
    t = Affine.translater((-1.2, 0.9, 0.0)) * Affine.rotater((0, 1, 0), -_to_radians(25.0))
    ref.add_instance(ref.add_mesh(P, idx, N=N, UV=UV), ref.plastic((0.2, 0.4, 0.7), (0.25, 0.25, 0.25), 0.2), fwd=t.fwd, inv=t.inv)
    P, N, UV, idx = pbrt_loader.load_ply(str(tmp_path / "lamp.ply"))
    t = Affine.translater((0.5, 3.0, 0.5)) * Affine.rotater((1, 0, 0), -_to_radians(10.0))
    L = (12.0, 12.0, 10.0)
    lm = ref.diffuse_light(L)
    for tri in idx:
        w = [tuple(float(c) for c in t.apply_point(P[v])) for v in tri]
        ref.add_area_light_triangle(w[0], w[1], w[2], L)
        ref.add_instance(ref.add_triangle(*(tuple(float(c) for c in P[v]) for v in tri)), lm, fwd=t.fwd, inv=t.inv)
    a, b = h.render_ids(0, msaa=1), ref.realize(hostsim_api).render_ids(0, msaa=1)
    assert (a[0] != 0xFFFFFFFF).mean() > 0.4
    # instance numbering differs (the loader appends lights in file order too, so it does not), ids equal
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
    fo, so = sd.realize(oracle_api).render_samples(integrator="path", msaa=2, max_depth=4, flags=1)
    fh, sh = h.render_samples(integrator="path", msaa=2, max_depth=4, flags=1)
    from tests.util import assert_radiance_close, assert_stats_close
    assert_radiance_close(fh, fo, "ply scene", outliers=1e-3)
    assert_stats_close(sh, so, "ply scene")
    assert fo.mean() > 0.01


def test_cornell_via_scene_file_equals_constructor_scene(oracle_api, hostsim_api):
    """BASELINE configs[0]: the Cornell box 'via scene_parser'.  The text route and the direct
    constructor route give the same scene; the product (host build) matches the oracle on it."""
    text_scene = scenes.cornell_box_via_parser(96, 96)
    ctor_scene = scenes.cornell_box(96, 96)
    a = text_scene.realize(oracle_api)
    b = ctor_scene.realize(oracle_api)
    ia, ib = a.render_ids(0, msaa=1, flags=4), b.render_ids(0, msaa=1, flags=4)
    assert (ia[0] == ib[0]).all() and (ia[1] == ib[1]).all()
    fa, _ = a.render(integrator="path", msaa=2)
    fb, _ = b.render(integrator="path", msaa=2)
    assert np.abs(fa - fb).mean() < 1e-6  # FP32 vs FP64 composition of the two box transforms only
    h = text_scene.realize(hostsim_api)
    ih = h.render_ids(0, msaa=1, flags=4)
    assert (ia[0] == ih[0]).all() and (ia[1] == ih[1]).all() and bits_equal(ia[2], ih[2]).all()


def test_exr_roundtrip_and_name(tmp_path):
    rng = np.random.default_rng(3)
    img = rng.random((37, 53, 3), dtype=np.float32) * 100.0
    img[0, 0] = [np.inf, 0.0, -0.0]
    path = str(tmp_path / film.exr_file_name("cornell", "path", 4))
    assert os.path.basename(path) == "cornell-path-16spp.exr"  # src/main.rs:238-243
    film.write_exr(path, img)
    back = film.read_exr(path)
    assert bits_equal(back, img).all()
    raw = open(path, "rb").read()
    assert raw[:4] == bytes([0x76, 0x2F, 0x31, 0x01])
    try:
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        cvimg = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except Exception:
        cvimg = None
    if cvimg is not None:  # an independent reader agrees (OpenCV returns BGR)
        np.testing.assert_array_equal(cvimg[1:, :, ::-1], img[1:])
