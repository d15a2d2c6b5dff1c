"""Randomised scenes: triangle soups (incl. slivers, duplicates, axis-aligned and coplanar faces),
spheres, instances under random affine transforms (non-uniform scale, shear), random lights and
materials; a second family adds quads, cuboids, disks, a sphere BLAS and a quad or disk light.  The product's walks and shading must agree with the oracle on every one of them --
this is what exercises the exact-order TLAS/BLAS emulation (extent quirks, tie-breaks, oversized
leaves) far from the hand-made scenes.  CPU: host build of the stage functions; GPU: the kernels."""
import numpy as np
import pytest

from pbrs_b200 import _capi as K
from pbrs_b200.scene import SceneDesc
from pbrs_b200.scenes import GOLD, SILVER
from tests.util import assert_radiance_close, assert_stats_close, bits_equal


def _add_simple_shapes(sd, mats, seed):
    """Quads, cuboids, disks and a sphere BLAS under random transforms, and a quad or disk light
    (their own generator, so the base scenes of the plain seeds stay what they were)."""
    rng = np.random.default_rng(seed + 99991)
    shapes = []
    for _ in range(int(rng.integers(1, 4))):
        shapes.append(sd.add_quad(tuple(rng.uniform(-1.5, 1.5, 3)), tuple(rng.normal(size=3)), tuple(rng.normal(size=3))))
    for _ in range(int(rng.integers(1, 3))):
        shapes.append(sd.add_cuboid(tuple(rng.uniform(-1.5, 0.0, 3)), tuple(rng.uniform(0.0, 1.5, 3))))
    for _ in range(int(rng.integers(1, 3))):
        n = rng.normal(size=3)
        r = np.cross(n, rng.normal(size=3))
        r = r / np.linalg.norm(r) * rng.uniform(0.3, 1.2)
        n32 = (n / np.linalg.norm(n)).astype(np.float32)
        r32 = r.astype(np.float32)
        if abs(float(np.dot(r32, n32 / np.linalg.norm(n32)))) < 5e-7:   # Disk::new asserts |radial . n| < 1e-6
            shapes.append(sd.add_disk(tuple(rng.uniform(-1.5, 1.5, 3)), tuple(n32), tuple(r32)))
    shapes.append(sd.add_triangle(*(tuple(v) for v in rng.uniform(-1.5, 1.5, (3, 3)))))
    nb = int(rng.choice([1, 3, 4, 5, 17, 60]))
    shapes.append(sd.add_sphere_blas(np.concatenate([rng.uniform(-1.2, 1.2, (nb, 3)), rng.uniform(0.05, 0.5, (nb, 1))], axis=1)))
    for sh in shapes:
        for _ in range(int(rng.integers(1, 3))):
            if rng.random() < 0.3:
                fwd = None
            else:
                fwd = np.eye(4)
                fwd[:3, :3] = rng.normal(size=(3, 3)) * 0.3 + np.eye(3) * rng.uniform(0.6, 1.4)
                fwd[:3, 3] = rng.uniform(-2.5, 2.5, 3)
            sd.add_instance(sh, int(rng.choice(mats)), fwd=fwd)
    L = tuple(rng.uniform(5, 25, 3))
    if rng.random() < 0.5:
        o, u, v = tuple(rng.uniform(-1, 1, 3) + np.array([0, 3.0, 0])), (float(rng.uniform(0.5, 2)), 0.0, 0.2), (0.1, 0.0, float(rng.uniform(0.5, 2)))
        sd.add_instance(sd.add_quad(o, u, v), sd.diffuse_light(L))
        sd.add_area_light_quad(o, u, v, L)
    else:
        c, n, r = tuple(rng.uniform(-1, 1, 3) + np.array([0, 3.0, 0])), (0.0, -1.0, 0.0), (float(rng.uniform(0.3, 1.0)), 0.0, 0.0)
        sd.add_instance(sd.add_disk(c, n, r), sd.diffuse_light(L))
        sd.add_area_light_disk(c, n, r, L)


def random_scene(seed, w=72, h=56, ext=False):
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    eye = rng.uniform(-1.0, 1.0, 3) + np.array([0.0, 0.5, -7.0])
    sd.set_camera(w, h, float(rng.uniform(35, 70)), tuple(eye), tuple(rng.uniform(-0.5, 0.5, 3)), (0, 1, 0))
    mats = [sd.lambertian(tuple(rng.uniform(0.1, 0.9, 3))), sd.lambertian(tuple(rng.uniform(0.1, 0.9, 3))),
            sd.metal(GOLD[0], GOLD[1], float(rng.uniform(0.0, 0.4))), sd.metal(SILVER[0], SILVER[1], 0.0),
            sd.mirror((0.9, 0.9, 0.9)), sd.dielectric(float(rng.uniform(1.2, 1.7))),
            sd.plastic(tuple(rng.uniform(0.1, 0.8, 3)), (0.3, 0.3, 0.3), float(rng.uniform(0.05, 0.4))),
            sd.glossy(tuple(rng.uniform(0.3, 0.9, 3)), float(rng.uniform(0.02, 0.3)))]
    meshes = []
    for _ in range(int(rng.integers(1, 4))):
        nt = int(rng.integers(1, 60))
        kind = rng.integers(0, 3)
        if kind == 0:    # soup of random triangles, some tiny, some huge
            P = rng.uniform(-1.5, 1.5, (nt * 3, 3)) * rng.choice([0.05, 1.0, 1.0, 3.0], (nt * 3, 1))
            idx = np.arange(nt * 3, dtype=np.uint32).reshape(-1, 3)
        elif kind == 1:  # a grid in an axis-aligned plane (flat boxes, shared edges), with duplicated faces
            g = int(rng.integers(2, 6))
            u, v = np.meshgrid(np.linspace(-1, 1, g + 1), np.linspace(-1, 1, g + 1), indexing="ij")
            P = np.stack([u, np.zeros_like(u), v], -1).reshape(-1, 3)
            a = (np.arange(g)[:, None] * (g + 1) + np.arange(g)[None, :]).reshape(-1)
            idx = np.concatenate([np.stack([a, a + 1, a + g + 2], -1), np.stack([a, a + g + 2, a + g + 1], -1)], 0).astype(np.uint32)
            idx = np.concatenate([idx, idx[: max(1, len(idx) // 4)]], 0)  # coincident copies: tie-breaks
        else:            # many triangles sharing one centroid (zero-extent split -> an oversized leaf) plus slivers
            c = rng.uniform(-0.5, 0.5, 3)
            tris = []
            for _k in range(int(rng.integers(5, 12))):
                d1, d2 = rng.normal(size=3), rng.normal(size=3)
                tris += [c + d1, c + d2, c - d1 - d2]
            sl = rng.uniform(-1, 1, 3)
            tris += [sl, sl + np.array([2.0, 0, 0]), sl + np.array([1.0, 1e-4, 0])]
            P = np.array(tris)
            idx = np.arange(len(tris), dtype=np.uint32).reshape(-1, 3)
        N = rng.normal(size=P.shape) if rng.random() < 0.5 else None
        UV = rng.random((P.shape[0], 2)) if rng.random() < 0.5 else None
        meshes.append(sd.add_mesh(P.astype(np.float32), idx, N=N, UV=UV))
    spheres = [sd.add_sphere(tuple(rng.uniform(-0.5, 0.5, 3)), float(rng.uniform(0.2, 1.0))) for _ in range(2)]
    for _ in range(int(rng.integers(2, 14))):
        shape = int(rng.choice(meshes + spheres))
        if rng.random() < 0.25:
            fwd = None
        else:
            A = rng.normal(size=(3, 3)) * 0.4 + np.eye(3) * rng.uniform(0.5, 1.5)  # scale + shear + rotation-ish
            fwd = np.eye(4)
            fwd[:3, :3] = A
            fwd[:3, 3] = rng.uniform(-2.5, 2.5, 3)
        sd.add_instance(shape, int(rng.choice(mats)), fwd=fwd)
    if rng.random() < 0.7:
        c, r, L = tuple(rng.uniform(-2, 2, 3) + np.array([0, 3, 0])), float(rng.uniform(0.2, 0.8)), tuple(rng.uniform(5, 30, 3))
        sd.add_instance(sd.add_sphere(c, r), sd.diffuse_light(L))
        sd.add_area_light_sphere(c, r, L)
    if rng.random() < 0.5:
        q = rng.uniform(-2, 2, (3, 3)) + np.array([0, 3.5, 0])
        L = tuple(rng.uniform(5, 20, 3))
        sd.add_instance(sd.add_mesh(q.astype(np.float32), np.array([[0, 1, 2]], np.uint32)), sd.diffuse_light(L))
        sd.add_area_light_triangle(q[0], q[1], q[2], L)
    if rng.random() < 0.5:
        sd.add_point_light(tuple(rng.uniform(-4, 4, 3)), tuple(rng.uniform(5, 40, 3)))
    if rng.random() < 0.4:
        sd.add_distant_light(tuple(rng.normal(size=3)), tuple(rng.uniform(0.2, 1.5, 3)))
    if ext:
        _add_simple_shapes(sd, mats, seed)
    env = rng.integers(0, 4)
    if env == 0:
        sd.set_env_constant(tuple(rng.uniform(0.0, 0.4, 3)))
    elif env == 1:
        sd.set_env_fn(int(rng.integers(0, 3)))
    return sd


def _check(oracle_api, api, seed, n_bad_allowed=2e-3, ext=False):
    sd = random_scene(seed, ext=ext)
    ho, hp = sd.realize(oracle_api), sd.realize(api)
    a, b = ho.render_ids(0, msaa=1), hp.render_ids(0, msaa=1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all(), f"seed {seed}: primary hits differ"
    kw = dict(integrator="path", msaa=1, max_depth=4, flags=K.FLAG_COUNT_TRAVERSAL)
    fa, sa = ho.render_samples(**kw)
    fb, sb = hp.render_samples(**kw)
    assert_radiance_close(fb, fa, f"seed {seed}", outliers=n_bad_allowed)
    assert_stats_close(sb, sa, f"seed {seed}", rel=2e-3)


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_hostsim(oracle_api, hostsim_api, seed):
    _check(oracle_api, hostsim_api, 1000 + seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_random_scenes_gpu(oracle_api, gpu_api, seed):
    _check(oracle_api, gpu_api, 1000 + seed)


@pytest.mark.parametrize("seed", range(16))
def test_random_scenes_with_simple_shapes_hostsim(oracle_api, hostsim_api, seed):
    _check(oracle_api, hostsim_api, 2000 + seed, ext=True)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_random_scenes_with_simple_shapes_gpu(oracle_api, gpu_api, seed):
    _check(oracle_api, gpu_api, 2000 + seed, ext=True)
