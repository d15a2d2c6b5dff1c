"""Pins the CPU oracle against every known-answer / property test the reference holds for the
hot path.  Constants are the reference's own (file:line cited per test)."""
import numpy as np
import pytest

from oracle import oracle_ffi as O

F32 = np.float32


def test_omega_local_trigonometry(oracle_api):
    # geometry/tests/bxdf_test.rs:16-26
    out, p = O.kat(O.KAT_OMEGA_TRIG, [0.64, 0.48, 0.6], 8)
    assert p == 0
    assert out[0] == F32(0.6)
    assert out[1] == F32(0.36)
    assert out[2] == F32(0.64)
    assert out[3] == F32(0.8)
    close = lambda a, b: 0.999 < b / a < 1.001
    assert close(out[4], 0.8) and close(out[5], 0.6) and close(out[6], 0.64) and close(out[7], 0.36)


def test_fresnel_bit_exact(oracle_api):
    # geometry/tests/bxdf_test.rs:29-49 -- assert_eq! on f32, i.e. bit-exact
    fwd = [F32(0.26872247), F32(0.112083375)]
    inv = [F32(1.0), F32(0.1645631)]
    for i, c in enumerate([0.3, 0.9]):
        a, _ = O.kat(O.KAT_FRESNEL_DIELECTRIC, [1.0, 2.0, c], 1)
        b, _ = O.kat(O.KAT_FRESNEL_DIELECTRIC, [2.0, 1.0, c], 1)
        assert a[0] == fwd[i], (a[0], fwd[i])
        assert b[0] == inv[i], (b[0], inv[i])
        a2, _ = O.kat(O.KAT_FRESNEL_DIELECTRIC, [2.0, 1.0, -c], 1)
        b2, _ = O.kat(O.KAT_FRESNEL_DIELECTRIC, [1.0, 2.0, -c], 1)
        assert a[0] == a2[0] and b[0] == b2[0]


def test_specular_dielectric_sample(oracle_api):
    # geometry/tests/bxdf_test.rs:52-61
    out, _ = O.kat(O.KAT_SPECULAR_DIELECTRIC, [1, 1, 1, 1.0, 2.0, 0.8, 0.0, 0.6, 0.0, 0.0], 8)
    assert out[3] == F32(-0.8)
    assert out[4] == 0.0 and np.signbit(out[4])  # -0.0
    assert out[5] == F32(0.6)
    assert out[6] == 1.0  # Prob::Mass


def test_reflect(oracle_api):
    # math/src/hcm.rs:672-679
    out, _ = O.kat(O.KAT_REFLECT, [0, 1, 0, 2.0, 1.0, 0.5], 3)
    assert np.sum((out - np.array([-2.0, 1.0, -0.5], F32)) ** 2) < np.finfo(F32).eps


def test_refract(oracle_api):
    # math/src/hcm.rs:681-705
    n = [0, 6.0, 0]
    wi = np.array([1, 1, 0], F32) / np.sqrt(F32(2))
    out, _ = O.kat(O.KAT_REFRACT, [*n, *wi, np.sqrt(F32(0.5))], 4)
    assert out[0] == 1.0
    wo = np.array([-0.5, -0.5 * np.sqrt(F32(3.0)), 0.0], F32)
    assert np.sum((out[1:4] - wo) ** 2) < np.finfo(F32).eps
    full = np.array([0.51, np.sqrt(F32(0.75)), 0], F32); full /= np.linalg.norm(full)
    trans = np.array([0.49, np.sqrt(F32(0.75)), 0], F32); trans /= np.linalg.norm(trans)
    assert O.kat(O.KAT_REFRACT, [*n, *full, 2.0], 4)[0][0] == 0.0
    assert O.kat(O.KAT_REFRACT, [*n, *trans, 2.0], 4)[0][0] == 1.0


def test_make_coord_system_orthonormal(oracle_api):
    # math/src/hcm.rs:585-594 doc-test
    v0 = np.array([0.3, 0.4, -0.6], F32); v0 = v0 / np.linalg.norm(v0)
    out, p = O.kat(O.KAT_MAKE_COORD, v0, 6)
    B = np.stack([v0, out[0:3], out[3:6]], axis=1).astype(np.float64)
    assert np.sum((B @ B.T - np.eye(3)) ** 2) < np.finfo(F32).eps


def test_cathetus_and_polynomial_style_doc_tests(oracle_api):
    # math/src/float.rs:73-80
    assert O.kat(O.KAT_CATHETUS, [1.0, 0.6], 1)[0][0] == F32(0.8)
    assert O.kat(O.KAT_CATHETUS, [1.0, -0.6], 1)[0][0] == F32(0.8)


def test_powi_is_square_and_multiply(oracle_api):
    # compiler-rt __powisf2 order: x^5 = x * (x^2)^2
    x = F32(1.1234567)
    x2 = F32(x * x); x4 = F32(x2 * x2)
    assert O.kat(O.KAT_POWI, [x, 5], 1)[0][0] == F32(x * x4)
    assert O.kat(O.KAT_POWI, [x, 2], 1)[0][0] == x2
    assert O.kat(O.KAT_POWI, [x, 3], 1)[0][0] == F32(x * x2)


def test_sphere_hit_miss_table(oracle_api):
    # shape/tests/frame_test.rs:55-85.  `intersect` expectations all hold; the dir_1 `occludes`
    # expectation (:68) is RED at reference HEAD (Sphere::occludes needs both roots in range,
    # shape/src/simple.rs:287; decision D3: the executable code is the parity target).
    c, r = [3.0, 4.0, 5.0], 1.6
    o = [0.1, 0.2, 0.1]
    scales = [0.001, 0.01, 0.1, 1.0, 10.0, 100.0, 1000.0]
    for s in scales:
        d = np.array([1.5, 2.0, 2.5], F32) * F32(s)
        out, _ = O.kat(O.KAT_SPHERE_INTERSECT, [*c, r, *o, *d, F32(1.0) / F32(s)], 6)
        assert out[0] == 0.0 and out[5] == 0.0
    for d0, both_roots in (([3.0, 4.0, 5.0], False), ([4.8, 6.4, 8.0], True)):
        for s in scales:
            d = np.array(d0, F32) * F32(s)
            out, _ = O.kat(O.KAT_SPHERE_INTERSECT, [*c, r, *o, *d, F32(1.0) / F32(s)], 6)
            assert out[0] == 1.0
            dist2 = np.sum((out[2:5].astype(np.float64) - np.array(c)) ** 2)
            assert abs(dist2 - r * r) <= 1e-4
            assert (out[5] == 1.0) == both_roots


def test_tricky_triangle_does_not_panic(oracle_api):
    # shape/src/blas.rs:497-522
    from pbrs_b200.scene import SceneDesc
    sd = SceneDesc()
    sd.set_camera(8, 8, 40.0, (0, 0, 0), (0, 0, 1))
    P = [[10.3457699, 27.3706398, -21.2291069], [10.3457699, 13.3905125, -21.1700611],
         [7.22226286, 13.3905125, -21.1700611]]
    N = [[0.0, 0.00419999985, 1.0]] * 3
    mesh = sd.add_mesh(P, [[0, 1, 2]], N=N, UV=[[0, 0]] * 3)
    sd.add_instance(mesh, sd.lambertian((0.5, 0.5, 0.5)))
    h = sd.realize(oracle_api)
    out = O.trace_ray(h, (0.0, 23.0, 30.0), (0.219424784, -0.0887561888, -1.08688462))
    assert out[15] == 0.0  # no reference assert would have fired


def _hemi(count):
    # Omega::tesselate_hemi, geometry/src/bxdf.rs:159-176 (spherical_direction swaps sin/cos phi)
    dth = (np.pi / 2) / count
    dph = (2 * np.pi) / (4 * count)
    th = (np.arange(count) + 0.5) * dth
    ph = (np.arange(4 * count) + 0.5) * dph
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    w = np.stack([np.sin(T) * np.sin(Pp), np.sin(T) * np.cos(Pp), np.cos(T)], -1).reshape(-1, 3)
    return w.astype(F32), np.sin(T).reshape(-1), dth, dph


def _lobe_args(kind, albedo, params, op, wo, x):
    a = np.zeros(19, F32)
    a[0] = kind; a[1:4] = albedo; a[4:4 + len(params)] = params; a[12] = op; a[13:16] = wo; a[16:16 + len(x)] = x
    return a


@pytest.mark.parametrize("kind,params", [(0, []), (1, [0.0])])
def test_diffuse_pdf_integrates_to_one(oracle_api, kind, params):
    # geometry/tests/bxdf_test.rs:64-70,116-138 (1-D integral, +-1e-3)
    wo = np.array([0.48, 0.64, 0.6], F32); wo /= np.linalg.norm(wo)
    w, st, dth, dph = _hemi(25)
    tot = 0.0
    for wi, s in zip(w, st):
        out, _ = O.kat(O.KAT_LOBE, _lobe_args(kind, [1, 2, 5], params, 1, wo, wi), 2)
        assert out[0] == 0.0  # a density
        tot += out[1] * s * dth * dph
    assert abs(tot - 1.0) < 1e-3


def test_lambert_mc_rho_matches_albedo(oracle_api):
    # geometry/tests/bxdf_test.rs:181-201 (800 samples)
    rng = np.random.default_rng(7)
    wo = np.array([0.2, -0.1, 0.9], F32); wo /= np.linalg.norm(wo)
    acc = np.zeros(3)
    n = 800
    for _ in range(n):
        u, v = rng.random(2)
        out, _ = O.kat(O.KAT_LOBE, _lobe_args(0, [1, 2, 5], [], 2, wo, [u, v]), 8)
        f, wi, pdf = out[0:3], out[3:6], out[7]
        acc += f * abs(wi[2]) * (0 if pdf == 0 else 1 / pdf)
    rho = acc / n
    assert np.sum((rho - [1, 2, 5]) ** 2) / 30.0 < 1e-3


def test_beckmann_d_integrates_to_one_and_masking(oracle_api):
    # geometry/tests/microfacet_test.rs:13-25 (Beckmann alpha = 0.2)
    w, st, dth, dph = _hemi(60)
    proj = 0.0
    wv = np.array([0.48, 0.64, 0.6], F32)
    masked = 0.0
    g1 = O.kat(O.KAT_BECKMANN, [0.2, 0.2, 1, *wv], 1)[0][0]
    for wh, s in zip(w, st):
        d = O.kat(O.KAT_BECKMANN, [0.2, 0.2, 0, *wh], 1)[0][0]
        proj += d * wh[2] * s * dth * dph
        masked += g1 * max(0.0, float(np.dot(wv, wh))) * d * s * dth * dph
    assert abs(proj - 1.0) < 4e-3
    assert abs(masked - wv[2]) < 2e-3


def test_beckmann_sample_wh_matches_bisector(oracle_api):
    # geometry/tests/bxdf_test.rs:203-231 (anisotropic Beckmann 0.2/0.3)
    rng = np.random.default_rng(3)
    wo = np.array([0.6, 0.8, 0.3], F32); wo /= np.linalg.norm(wo)
    for _ in range(40):
        u, v = rng.random(2)
        wh = O.kat(O.KAT_BECKMANN, [0.2, 0.3, 3, *wo, u, v], 3)[0]
        out, _ = O.kat(O.KAT_LOBE, _lobe_args(5, [3.0, 3.4, 2.9], [0.2, 0.3], 2, wo, [u, v]), 8)
        f, wi = out[0:3], out[3:6]
        if np.all(f <= 0):
            continue
        b = wo + wi
        b = b / np.linalg.norm(b)
        assert np.sum((b - wh) ** 2) < 1e-3


def test_sphere_light_cone_pdf_and_samples(oracle_api):
    # light/tests/shape_sample_test.rs:10-20,23-66,69-90
    c, r = np.array([0.0, 5.0, 0.0]), 1.0
    tp, tn = [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]
    rng = np.random.default_rng(11)
    for _ in range(50):
        u, v = rng.random(2)
        out = O.kat(O.KAT_SPHERE_LIGHT, [*c, r, *tp, *tn, 0, u, v], 6)[0]
        assert abs(np.linalg.norm(out[0:3] - c) - r) < 1e-3
    # uniform-cone pdf integrates to 1 over the sphere of directions
    n = 4000
    th = (np.arange(n) + 0.5) * np.pi / n
    ph = (np.arange(8) + 0.5) * 2 * np.pi / 8
    tot = 0.0
    dth, dph = np.pi / n, 2 * np.pi / 8
    for t in th[: n // 4]:  # the cone (half angle asin(r/d) = 11.5 deg) lies well inside t < pi/4
        for p in ph:
            wi = [np.sin(t) * np.cos(p), np.cos(t), np.sin(t) * np.sin(p)]
            out = O.kat(O.KAT_SPHERE_LIGHT, [*c, r, *tp, *tn, 1, *wi], 2)[0]
            if out[0]:
                tot += out[1] * np.sin(t) * dth * dph
    assert abs(tot - 1.0) < 1e-2


def test_sampler_mapping(oracle_api):
    # rand 0.8 Standard f32 mapping: (u32 >> 8) * 2^-24 in [0, 1); distinct dims decorrelate
    vals = [oracle_api["sampler_u32"](0x5EED, 17, 3, d) for d in range(64)]
    assert len(set(vals)) == 64
    f = [(v >> 8) * 2.0 ** -24 for v in vals]
    assert all(0.0 <= x < 1.0 for x in f)
    assert 0.3 < np.mean(f) < 0.7


def test_bbox_slab_basic_and_nan_semantics(oracle_api):
    # geometry/src/bvh.rs:84-99 (glam SSE min/max: Q15).  Parity unpinned upstream.
    out = O.kat(O.KAT_BBOX, [0, 0, 0, 1, 1, 1, -1, 0.5, 0.5, 1, 0, 0, np.inf], 2)[0]
    assert out[0] == 1.0 and out[1] == 1.0
    out = O.kat(O.KAT_BBOX, [0, 0, 0, 1, 1, 1, -1, 2.5, 0.5, 1, 0, 0, np.inf], 2)[0]
    assert out[0] == 0.0
    # ray origin on a slab plane with a zero direction component => 0/0 = NaN lanes
    out = O.kat(O.KAT_BBOX, [0, 0, 0, 1, 1, 1, -1, 0.0, 0.5, 1, 0, 0, np.inf], 2)[0]
    assert out[0] in (0.0, 1.0)  # must not crash; value fixed by the SSE semantics
    # t_max prunes
    out = O.kat(O.KAT_BBOX, [0, 0, 0, 1, 1, 1, -1, 0.5, 0.5, 1, 0, 0, 0.5], 2)[0]
    assert out[0] == 0.0


def test_reference_transform_tests(oracle_api, hostsim_api):
    """geometry/src/transform.rs:327-355 (`test_inverse`, `test_bbox_transform`) on the instance
    transform both hosts use: rotater(axis (0.6, 0.8, 0), 0.3 rad) * translater; inverse * forward is
    the identity to f32 epsilon, and the transformed box (what the TLAS leaf stores,
    transform.rs:287-308) contains every transformed corner."""
    from pbrs_b200.pbrt_loader import Affine, _mat_mul
    from pbrs_b200.scene import SceneDesc
    t = Affine.rotater((0.6, 0.8, 0.0), np.float32(0.3)) * Affine.translater((0.3, 0.4, 0.6))
    ident = _mat_mul(t.inv, t.fwd)
    assert ((ident[:3, :3] - np.eye(3, dtype=np.float32)) ** 2).sum() <= np.finfo(np.float32).eps
    assert (ident[:3, 3] ** 2).sum() <= np.finfo(np.float32).eps
    t = Affine.rotater((0.6, 0.8, 0.0), np.float32(0.3)) * Affine.translater((7.0, 8.0, -13.0))
    lo, hi = np.array([-0.3, 0.4, 0.8], np.float32), np.array([3.4, 2.3, 4.4], np.float32)
    for api in (oracle_api, hostsim_api):
        sd = SceneDesc()
        sd.set_camera(16, 16, 45.0, (0, 0, -30), (0, 0, 0))
        sd.add_instance(sd.add_cuboid(tuple(lo), tuple(hi)), sd.lambertian((0.5, 0.5, 0.5)), fwd=t.fwd, inv=t.inv)
        info = sd.realize(api).info()
        wmin, wmax = np.array(info.world_min[:3], np.float32), np.array(info.world_max[:3], np.float32)
        for cx in (lo[0], hi[0]):
            for cy in (lo[1], hi[1]):
                for cz in (lo[2], hi[2]):
                    p = t.apply_point((cx, cy, cz))
                    assert (p >= wmin).all() and (p <= wmax).all()


def test_reference_custom_frame_test(oracle_api):
    """shape/tests/frame_test.rs:18-52: an Interaction with normal n and dpdu from
    make_coord_system(n) keeps n and gets dpdu as its tangent.  Reached here through a quad whose
    sides are (dpdu, dpdv): ParallelQuad::intersect ends in Interaction::new(.., n).with_dpdu(side_u)."""
    from pbrs_b200.scene import SceneDesc
    n = np.array([-0.3, 0.5, 1.0], F32); n = (n * (F32(1.0) / np.sqrt((n * n).sum(dtype=F32)))).astype(F32)
    out, p = O.kat(O.KAT_MAKE_COORD, n, 6)
    dpdu, dpdv = out[0:3].astype(F32), out[3:6].astype(F32)
    assert p == 0 and abs(float(n @ dpdu)) < 1e-4 and abs(float(n @ dpdv)) < 1e-4
    c = np.array([3.0, 2.5, 2.0], F32)
    sd = SceneDesc()
    sd.set_camera(16, 16, 45.0, (0, 0, -30), (0, 0, 0))
    sd.add_instance(sd.add_quad(tuple(c - 0.5 * dpdu - 0.5 * dpdv), tuple(dpdu), tuple(dpdv)), sd.lambertian((0.5, 0.5, 0.5)))
    h = sd.realize(oracle_api)
    hit = O.trace_ray(h, tuple(c + 2.0 * n), tuple(-n))
    assert hit[0] == 1.0 and hit[15] == 0.0                      # has_valid_frame: no assert fired
    assert ((hit[12:15] - dpdu) ** 2).sum() < 1e-6               # tangent == dpdu
    assert ((hit[5:8] - n) ** 2).sum() < 1e-6                    # normal == n (it already faces the ray)
    np.testing.assert_allclose(hit[8:10], (0.5, 0.5), atol=1e-5)
