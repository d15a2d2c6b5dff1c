"""Either side of the hot path: pbrt-subset scene files in (pbrs_b200/pbrt_loader.py, mirroring
scene_parser + scene/src/loader.rs) and the EXR film out (pbrs_b200/film.py, src/main.rs:42-53)."""
import os

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
    for body in ['Shape "plymesh" "string filename" "a.ply"', 'Shape "loopsubdiv"', 'ObjectBegin "x" ObjectEnd', 'Material "fourier"',
                 'LightSource "spot"', 'AreaLightSource "diffuse" "rgb L" [ 1 1 1 ] Shape "trianglemesh"']:
        with pytest.raises(PbrtError):
            load_pbrt_string(base % body)
    with pytest.raises(PbrtError):
        load_pbrt_string('WorldBegin WorldEnd')  # no camera / film


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
