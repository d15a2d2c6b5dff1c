"""ParallelQuad, Cuboid, Disk and IsoBlas<Sphere> (shape/src/simple.rs:33-182,
291-416, shape/src/blas.rs:36-70,263-292) and the quad / disk area-light shapes
(light/src/sample_shape.rs:257-309).

CPU part: the oracle against known answers -- the reference's own `quad_frame_test`
(shape/tests/frame_test.rs:9-15) and values derived by hand from the cited formulas, including the
quirks (Q10 Disk::occludes ignores the extent, Q11 quad u/v from norms + reciprocal `t` in
occludes) -- then the product's host-compiled stage functions against the oracle.  The GPU part
runs the same comparison through the C ABI (tests/test_gpu_parity.py covers the scene families;
here: single-shape scenes with rays that start inside / graze / miss)."""
import numpy as np
import pytest

from pbrs_b200 import scenes
from pbrs_b200.scene import PbrsError, SceneDesc
from tests.util import assert_radiance_close, assert_stats_close, bits_equal


def _single(build, fwd=None, eye=(0.0, 0.0, -5.0), target=(0.0, 0.0, 0.0), size=(48, 48), fov=45.0):
    sd = SceneDesc()
    sd.set_camera(size[0], size[1], fov, eye, target)
    sd.add_instance(build(sd), sd.lambertian((0.5, 0.5, 0.5)), fwd=fwd)
    sd.add_point_light((2.0, 3.0, -4.0), (30.0, 30.0, 30.0))
    sd.set_env_constant((0.1, 0.1, 0.1))
    return sd


# ------------------------------------------------------------------------------------------------
# oracle known answers
# ------------------------------------------------------------------------------------------------
def test_reference_quad_frame_test(oracle_api):
    # shape/tests/frame_test.rs:9-15: the hit's frame is valid (no with_dpdu assert fires)
    from oracle import oracle_ffi as O
    h = _single(lambda sd: sd.add_quad_xy((-1.0, 1.0), (-1.0, 1.0), 0.0)).realize(oracle_api)
    out = O.trace_ray(h, (0.5, 0.5, -1.0), (-0.2, -0.2, 1.0))
    assert out[0] == 1.0 and out[15] == 0.0
    np.testing.assert_allclose(out[1], 1.0, rtol=1e-6)                  # t
    np.testing.assert_allclose(out[2:5], (0.3, 0.3, 0.0), atol=1e-6)    # accurate hit
    np.testing.assert_allclose(out[5:8], (0.0, 0.0, -1.0), atol=1e-7)   # normal faces the ray
    np.testing.assert_allclose(out[8:10], (0.65, 0.65), atol=1e-6)      # (u, v) = (p - origin) / side
    np.testing.assert_allclose(out[12:15], (1.0, 0.0, 0.0), atol=1e-7)  # tangent = side_u direction


def test_quad_known_answers(oracle_api):
    from oracle import oracle_ffi as O
    h = _single(lambda sd: sd.add_quad((0.0, 0.0, 0.0), (2.0, 0.0, 0.0), (0.0, 1.0, 0.0))).realize(oracle_api)
    # from behind: the normal flips to face the ray (Vec3::facing, hcm.rs:124-130)
    out = O.trace_ray(h, (1.0, 0.5, 3.0), (0.0, 0.0, -1.0))
    assert out[0] == 1.0
    np.testing.assert_allclose(out[1], 3.0)
    np.testing.assert_allclose(out[5:8], (0.0, 0.0, 1.0), atol=1e-7)
    np.testing.assert_allclose(out[8:10], (0.5, 0.5), atol=1e-6)
    # outside the parallelogram
    assert O.trace_ray(h, (2.5, 0.5, 3.0), (0.0, 0.0, -1.0))[0] == 0.0
    # parallel to the plane: t = x / 0 is not inside the extent
    assert O.trace_ray(h, (1.0, 0.5, 3.0), (1.0, 0.0, 0.0))[0] == 0.0
    # the extent cuts the hit off (Ray::truncated_t, ray.rs:40-46)
    assert O.trace_ray(h, (1.0, 0.5, 3.0), (0.0, 0.0, -1.0), t_max=2.5)[0] == 0.0
    # Q11 occludes: t = (d.n) / ((origin - o).n) = 1 / distance.  At distance 1 that is the hit...
    assert O.occludes_ray(h, (1.0, 0.5, -1.0), (0.0, 0.0, 1.0)) == 1
    # ...at distance 4 it is t = 1/4: the point (1, .5, -3.75) is off the plane and the cross-product
    # norms pick that up: v = |a x d| / |a x b| = |(0, 7.5, 1)| / 2 > 1 -> not occluded
    assert O.occludes_ray(h, (1.0, 0.5, -4.0), (0.0, 0.0, 1.0)) == 0
    # ...and a slanted ray that does hit the quad at distance 2 is tested at t = 0.5 instead
    assert O.occludes_ray(h, (-2.5, 0.5, -2.0), (2.0, 0.0, 1.0)) == 0
    assert O.trace_ray(h, (-2.5, 0.5, -2.0), (2.0, 0.0, 1.0))[0] == 1.0


def test_quad_mirrored_extension_is_flagged(oracle_api, hostsim_api):
    # Q11: with the quad as an area-light shape the TLAS box does not cull the mirrored region;
    # u, v from norms pass `inside` and the accurate-vs-coarse assert would fire (simple.rs:140-147).
    sd = SceneDesc()
    sd.set_camera(64, 48, 60.0, (0.0, 2.0, -6.0), (0.0, 0.5, 0.0))
    sd.add_instance(sd.add_quad_xz((-8.0, 8.0), 0.0, (-8.0, 8.0)), sd.lambertian((0.7, 0.7, 0.7)))
    L = (10.0, 10.0, 10.0)
    o, u, v = (0.0, 3.0, 0.0), (1.5, 0.0, 0.0), (0.0, 0.0, 1.5)
    sd.add_instance(sd.add_quad(o, u, v), sd.diffuse_light(L))
    sd.add_area_light_quad(o, u, v, L)
    kw = dict(integrator="path", msaa=2, max_depth=3, flags=1)
    a, sa = sd.realize(oracle_api).render_samples(**kw)
    b, sb = sd.realize(hostsim_api).render_samples(**kw)
    assert sa["would_panic"].get("quad", 0) > 0
    assert_radiance_close(b, a, "quad light", outliers=1e-3)
    assert_stats_close(sb, sa, "quad light")


def test_cuboid_known_answers(oracle_api):
    from oracle import oracle_ffi as O
    # from_points orders the corners per axis (simple.rs:173-181)
    h = _single(lambda sd: sd.add_cuboid((1.0, 2.0, 1.0), (-1.0, -2.0, -1.0))).realize(oracle_api)
    out = O.trace_ray(h, (0.25, 0.5, -5.0), (0.0, 0.0, 1.0))
    assert out[0] == 1.0 and out[15] == 0.0
    np.testing.assert_allclose(out[1], 4.0)
    np.testing.assert_allclose(out[2:5], (0.25, 0.5, -1.0))
    np.testing.assert_allclose(out[5:8], (0.0, 0.0, -1.0))   # -signum(dir[axis]) on the hit axis
    np.testing.assert_allclose(out[8:10], (0.5, 0.5))        # constant uv (:410)
    np.testing.assert_allclose(out[12:15], (1.0, 0.0, 0.0))  # tangent on axis + 1: z -> x
    # from inside: the EXIT face, normal against the ray (:392-396)
    out = O.trace_ray(h, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    assert out[0] == 1.0
    np.testing.assert_allclose(out[1], 2.0)
    np.testing.assert_allclose(out[2:5], (0.0, 2.0, 0.0))
    np.testing.assert_allclose(out[5:8], (0.0, -1.0, 0.0))
    np.testing.assert_allclose(out[12:15], (0.0, 0.0, 1.0))  # y -> z
    # from inside with the exit beyond the extent: hit_max keeps bound = -inf -> None (:397-399)
    assert O.trace_ray(h, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), t_max=1.5)[0] == 0.0
    # occludes is the box test (:412-415): true from inside, false past the extent
    assert O.occludes_ray(h, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), t_max=1.5) == 1
    assert O.occludes_ray(h, (0.25, 0.5, -5.0), (0.0, 0.0, 1.0), t_max=3.5) == 0
    assert O.occludes_ray(h, (0.25, 0.5, -5.0), (0.0, 0.0, 1.0), t_max=4.5) == 1
    # a miss
    assert O.trace_ray(h, (3.0, 0.0, -5.0), (0.0, 0.0, 1.0))[0] == 0.0


def test_disk_known_answers(oracle_api):
    from oracle import oracle_ffi as O
    # the normal is normalised by Disk::new (simple.rs:43)
    h = _single(lambda sd: sd.add_disk((0.0, 0.0, 0.0), (0.0, 0.0, 3.0), (2.0, 0.0, 0.0))).realize(oracle_api)
    out = O.trace_ray(h, (0.0, 1.0, -5.0), (0.0, 0.0, 1.0))
    assert out[0] == 1.0 and out[15] == 0.0
    np.testing.assert_allclose(out[1], 5.0)
    np.testing.assert_allclose(out[2:5], (0.0, 1.0, 0.0), atol=1e-7)
    np.testing.assert_allclose(out[5:8], (0.0, 0.0, -1.0), atol=1e-7)
    # u = fract(atan2((radial x cp).n, radial.cp) / pi + 1): cp = +y, n = -z -> atan2(-2, 0) = -pi/2 -> 0.5
    np.testing.assert_allclose(out[8:10], (0.5, 0.5), atol=1e-6)
    # tangent = hat(n x cp) = (0,0,-1) x (0,1,0) = (1, 0, 0)
    np.testing.assert_allclose(out[12:15], (1.0, 0.0, 0.0), atol=1e-7)
    # outside the radius (inside the bounding square)
    assert O.trace_ray(h, (1.6, 1.6, -5.0), (0.0, 0.0, 1.0))[0] == 0.0
    # Q10: occludes ignores the extent -- but the TLAS leaf box in front of it does not
    assert O.occludes_ray(h, (0.0, 1.0, -5.0), (0.0, 0.0, 1.0)) == 1
    assert O.occludes_ray(h, (0.0, 1.0, -5.0), (0.0, 0.0, 1.0), t_max=4.0) == 0


def test_disk_rejects_what_the_reference_asserts(oracle_api, hostsim_api):
    for api in (oracle_api, hostsim_api):
        with pytest.raises(PbrsError):   # radial not perpendicular to the normal (simple.rs:45)
            _single(lambda sd: sd.add_disk((0, 0, 0), (0, 0, 1), (1.0, 0.0, 0.1))).realize(api)
        with pytest.raises(PbrsError):   # zero normal: Vec3::hat asserts
            _single(lambda sd: sd.add_disk((0, 0, 0), (0, 0, 0), (1.0, 0.0, 0.0))).realize(api)
        sd = _single(lambda sd: sd.add_sphere((0, 0, 0), 1.0))
        sd.add_area_light_disk((0, 3, 0), (0, -1, 0), (0.5, 0.2, 0.0), (1, 1, 1))
        with pytest.raises(PbrsError):
            sd.realize(api)


def test_isolated_triangle_known_answers(oracle_api):
    # IsolatedTriangle::intersect = intersect_triangle + with_dpdu(p1 - p0) (simple.rs:425-427):
    # (u, v) are the barycentrics of p1 and p2, the normal faces the ray
    from oracle import oracle_ffi as O
    h = _single(lambda sd: sd.add_triangle((0.0, 0.0, 0.0), (2.0, 0.0, 0.0), (0.0, 2.0, 0.0))).realize(oracle_api)
    out = O.trace_ray(h, (0.5, 1.0, -3.0), (0.0, 0.0, 1.0))
    assert out[0] == 1.0 and out[15] == 0.0
    np.testing.assert_allclose(out[1], 3.0)
    np.testing.assert_allclose(out[2:5], (0.5, 1.0, 0.0), atol=1e-6)
    np.testing.assert_allclose(out[5:8], (0.0, 0.0, -1.0), atol=1e-7)
    np.testing.assert_allclose(out[8:10], (0.25, 0.5), atol=1e-6)
    np.testing.assert_allclose(out[12:15], (1.0, 0.0, 0.0), atol=1e-6)
    assert O.trace_ray(h, (1.5, 1.5, -3.0), (0.0, 0.0, 1.0))[0] == 0.0            # outside the hypotenuse
    assert O.occludes_ray(h, (0.5, 1.0, -3.0), (0.0, 0.0, 1.0)) == 1
    assert O.occludes_ray(h, (0.5, 1.0, -3.0), (0.0, 0.0, 1.0), t_max=2.0) == 0
    # the same triangle as a one-triangle TriangleMesh listed (0, 2, 1): the mesh swaps (i, k, j) back,
    # so t and the position are bit-identical (the mesh's uv / tangent are its own)
    m = _single(lambda sd: sd.add_mesh(np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0]], np.float32), np.array([[0, 2, 1]], np.uint32))).realize(oracle_api)
    for o, d in [((0.5, 1.0, -3.0), (0.1, -0.05, 1.0)), ((0.3, 0.2, 4.0), (0.02, 0.1, -1.0))]:
        a, b = O.trace_ray(h, o, d), O.trace_ray(m, o, d)
        assert a[0] == b[0] == 1.0 and bits_equal(a[1:8], b[1:8]).all()


def test_sphere_blas_equals_individual_spheres_where_unambiguous(oracle_api):
    # The same spheres as one IsoBlas instance and as separate instances: primary hits agree on t
    # and, through prim / instance id, on the sphere -- except where two spheres overlap along
    # the ray start (tie rules differ) -- so compare only pixels whose hit is far from any other.
    rng = np.random.default_rng(3)
    balls = np.concatenate([rng.uniform(-2.0, 2.0, (40, 3)), rng.uniform(0.15, 0.4, (40, 1))], axis=1).astype(np.float32)
    a = SceneDesc(); b = SceneDesc()
    for sd in (a, b):
        sd.set_camera(64, 64, 50.0, (0.0, 0.0, -7.0), (0.0, 0.0, 0.0))
    m = a.lambertian((0.5, 0.5, 0.5))
    a.add_instance(a.add_sphere_blas(balls), m)
    m = b.lambertian((0.5, 0.5, 0.5))
    for c in balls:
        b.add_instance(b.add_sphere(c[:3], c[3]), m)
    ia, pa, ta = a.realize(oracle_api).render_ids(0, msaa=1)
    ib, pb, tb = b.realize(oracle_api).render_ids(0, msaa=1)
    hit = ia != 0xFFFFFFFF
    assert hit.sum() > 200 and (hit == (ib != 0xFFFFFFFF)).all()
    assert bits_equal(ta[hit], tb[hit]).all()
    assert (pa[hit] == ib[hit]).all()   # prim id of the BLAS = index of the sphere = instance id in scene b


# ------------------------------------------------------------------------------------------------
# product (host-compiled stage functions) against the oracle on single-shape scenes
# ------------------------------------------------------------------------------------------------
def _cases():
    rot = scenes.translate((0.2, -0.1, 0.3)) @ scenes.rotate_axis((1.0, 2.0, 0.5), 0.7) @ scenes.scale(1.3)
    shear = np.array([[1.2, 0.3, 0.0, 0.1], [0.0, 0.9, 0.2, 0.0], [0.1, 0.0, 1.1, -0.2], [0, 0, 0, 1]], np.float64)
    rng = np.random.default_rng(11)
    balls = np.concatenate([rng.uniform(-1.5, 1.5, (33, 3)), rng.uniform(0.1, 0.5, (33, 1))], axis=1)
    return {
        "quad": (lambda sd: sd.add_quad((-1.0, -1.0, 0.0), (2.0, 0.3, 0.4), (-0.2, 1.8, 0.1)), None, (0.0, 0.0, -5.0)),
        "quad_rot": (lambda sd: sd.add_quad_xy((-1.0, 1.0), (-1.0, 1.0), 0.0), rot, (0.0, 0.0, -5.0)),
        "cuboid": (lambda sd: sd.add_cuboid((-1.0, -0.5, -1.0), (1.0, 0.8, 0.5)), rot, (0.0, 1.0, -5.0)),
        "cuboid_inside": (lambda sd: sd.add_cuboid((-3.0, -3.0, -6.0), (3.0, 3.0, 3.0)), None, (0.0, 0.0, -5.0)),
        "cuboid_shear": (lambda sd: sd.add_cuboid((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)), shear, (0.5, 0.5, -5.0)),
        "disk": (lambda sd: sd.add_disk((0.0, 0.0, 0.0), (0.2, 0.3, -1.0), (1.5, 0.0, 0.3)), None, (0.0, 0.0, -5.0)),
        "disk_rot": (lambda sd: sd.add_disk((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0)), rot, (0.0, 0.0, -5.0)),
        "triangle": (lambda sd: sd.add_triangle((-1.5, -1.0, 0.2), (1.6, -0.8, -0.3), (0.1, 1.7, 0.4)), None, (0.0, 0.0, -5.0)),
        "triangle_shear": (lambda sd: sd.add_triangle((-1.5, -1.0, 0.0), (1.5, -1.0, 0.0), (0.0, 1.5, 0.0)), shear, (0.3, 0.2, -5.0)),
        "balls": (lambda sd: sd.add_sphere_blas(balls), rot, (0.0, 0.0, -6.0)),
        "four_balls": (lambda sd: sd.add_sphere_blas(balls[:4]), None, (0.0, 0.0, -6.0)),   # the root is a leaf
    }


CASES = _cases()


def _compare(api_a, api_b, name):
    build, fwd, eye = CASES[name]
    sd = _single(build, fwd=fwd, eye=eye)
    ha, hb = sd.realize(api_a), sd.realize(api_b)
    for flags in (0, 4):   # jittered, and pixel centres (axis-aligned centre rays)
        a = ha.render_ids(1, msaa=2, flags=flags)
        b = hb.render_ids(1, msaa=2, flags=flags)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all(), name
    assert (a[0] != 0xFFFFFFFF).sum() > 50, "the shape is not in view"
    for integrator, depth in (("direct", 5), ("path", 4)):
        fa, sa = ha.render_samples(integrator=integrator, msaa=2, max_depth=depth, flags=1)
        fb, sb = hb.render_samples(integrator=integrator, msaa=2, max_depth=depth, flags=1)
        assert_radiance_close(fb, fa, f"{name} {integrator}", outliers=1e-3)
        assert_stats_close(sb, sa, f"{name} {integrator}")


@pytest.mark.parametrize("name", list(CASES))
def test_single_shape_hostsim(oracle_api, hostsim_api, name):
    _compare(oracle_api, hostsim_api, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_single_shape_gpu(oracle_api, gpu_api, name):
    _compare(oracle_api, gpu_api, name)


def test_area_lights_quad_and_disk_hostsim(oracle_api, hostsim_api):
    sd = scenes.shape_zoo(64, 48)
    for integrator in ("direct", "path"):
        kw = dict(integrator=integrator, msaa=3, max_depth=5, flags=1)
        a, sa = sd.realize(oracle_api).render_samples(**kw)
        b, sb = sd.realize(hostsim_api).render_samples(**kw)
        assert np.isfinite(a).all() and a.mean() > 0.05
        assert_radiance_close(b, a, "shape_zoo " + integrator, outliers=1e-3)
        assert_stats_close(sb, sa, "shape_zoo " + integrator)


@pytest.mark.gpu
def test_presets_full_size_gpu_vs_oracle_crop(oracle_api, gpu_api):
    # the reference's own presets at their native resolution: ids bit-exact on the whole frame,
    # radiance on a crop (the oracle is slow)
    for sd in (scenes.preset_cornell_box(), scenes.preset_everything(400, 400)):
        ho, hg = sd.realize(oracle_api), sd.realize(gpu_api)
        a = ho.render_ids(0, msaa=1)
        b = hg.render_ids(0, msaa=1)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all()
        kw = dict(integrator="path", msaa=2, max_depth=5, flags=1, crop=(150, 150, 96, 64))
        fa, sa = ho.render_samples(**kw)
        fb, sb = hg.render_samples(**kw)
        assert_radiance_close(fb, fa, "preset", outliers=1e-3)
        assert_stats_close(sb, sa, "preset")
