"""CPU soak: the product's stage functions (host build, tests/hostsim) against the oracle on
randomised scenes far beyond the seeds of tests/test_fuzz_parity.py.
usage: python tools/soak_parity.py <first_seed> <last_seed> [zoo]
`zoo` switches to a second generator: every material kind (uber / substrate with texture slots),
image and Perlin textures, image / function / constant environments, sheared instances.
Every seed is run without and with the simple shapes, under path depth 4 (the suite's check), the
direct integrator, 4 spp, and path depth 6 without jitter.  Round 1: 1 400 + 1 440 renders, no
mismatch after the far-root flag."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_ffi  # noqa: E402  (test infrastructure: this is a test tool)
from pbrs_b200 import _capi as K  # noqa: E402
from tests.hostsim import load as hs_load  # noqa: E402
from tests.test_fuzz_parity import _check, random_scene  # noqa: E402
from tests.util import assert_radiance_close, assert_stats_close, bits_equal  # noqa: E402
import numpy as np  # noqa: E402
from pbrs_b200.scene import SceneDesc  # noqa: E402
from pbrs_b200.scenes import COPPER, checker_noise_image, icosphere, perlin_tables  # noqa: E402


def zoo_scene(seed):
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(64, 48, float(rng.uniform(35, 65)), tuple(rng.uniform(-1, 1, 3) + np.array([0, 1.0, -8.0])), (0, 0.3, 0))
    img = checker_noise_image(64, seed)
    timg = sd.add_texture_image(img)
    rv, px, py, pz = perlin_tables(seed)
    tper = sd.add_texture_perlin(float(rng.uniform(1, 8)), rv, px, py, pz)
    tsol = lambda: sd.add_texture_solid(tuple(rng.uniform(0.05, 0.9, 3)))
    mats = [sd.lambertian(tex=timg), sd.lambertian(tex=tper), sd.lambertian(tuple(rng.uniform(0.1, 0.9, 3))),
            sd.metal(COPPER[0], COPPER[1], float(rng.uniform(0, 0.5))), sd.glossy(tuple(rng.uniform(0.3, 0.9, 3)), float(rng.uniform(1e-4, 0.3))),
            sd.mirror(tuple(rng.uniform(0.5, 1, 3))), sd.dielectric(float(rng.uniform(1.1, 2.0)), tuple(rng.uniform(0.5, 1, 3)), tuple(rng.uniform(0.5, 1, 3))),
            sd.plastic(tuple(rng.uniform(0.1, 0.8, 3)), tuple(rng.uniform(0.1, 0.5, 3)), float(rng.uniform(0.01, 0.5)), remap_roughness=bool(rng.random() < 0.5)),
            sd.uber(tsol(), tsol(), tex_kr=tsol() if rng.random() < 0.5 else -1, tex_kt=tsol() if rng.random() < 0.5 else -1,
                    rough_u=float(rng.uniform(0.01, 0.4)), rough_v=float(rng.uniform(0.01, 0.4)), eta=float(rng.uniform(1.2, 1.8)), remap_roughness=bool(rng.random() < 0.5)),
            sd.uber(timg, tsol(), rough_u=0.1, rough_v=0.1), sd.substrate(tsol(), tsol()), sd.substrate(timg, tper)]
    P, N, UV, idx = icosphere(1, radius=0.8)
    ico = sd.add_mesh(P, idx, N=N, UV=UV)
    floor = sd.add_mesh(np.array([[-6, -1, -6], [6, -1, -6], [6, -1, 6], [-6, -1, 6]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32),
                        UV=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32))
    sd.add_instance(floor, int(rng.choice(mats)))
    for k in range(int(rng.integers(4, 12))):
        sh = ico if rng.random() < 0.5 else sd.add_sphere((0, 0, 0), float(rng.uniform(0.4, 0.9)))
        fwd = np.eye(4); fwd[:3, :3] = np.eye(3) * rng.uniform(0.6, 1.3) + rng.normal(size=(3, 3)) * 0.1; fwd[:3, 3] = rng.uniform(-3, 3, 3) * (1, 0.5, 1)
        sd.add_instance(sh, int(rng.choice(mats)), fwd=fwd)
    c, r, L = tuple(rng.uniform(-2, 2, 3) + np.array([0, 3.5, 0])), float(rng.uniform(0.2, 0.7)), tuple(rng.uniform(5, 30, 3))
    sd.add_instance(sd.add_sphere(c, r), sd.diffuse_light(L)); sd.add_area_light_sphere(c, r, L)
    if rng.random() < 0.5: sd.add_point_light(tuple(rng.uniform(-4, 4, 3) + np.array([0, 3, 0])), tuple(rng.uniform(5, 40, 3)))
    if rng.random() < 0.5: sd.add_distant_light(tuple(rng.normal(size=3) - np.array([0, 1.5, 0])), tuple(rng.uniform(0.2, 1.5, 3)))
    e = rng.integers(0, 4)
    if e == 0: sd.set_env_image(checker_noise_image(32, seed + 7), tuple(rng.uniform(0.2, 1.0, 3)))
    elif e == 1: sd.set_env_fn(int(rng.integers(0, 3)))
    elif e == 2: sd.set_env_constant(tuple(rng.uniform(0, 0.3, 3)))
    return sd


def check_zoo(o, h, seed):
    sd = zoo_scene(seed)
    ho, hp = sd.realize(o), sd.realize(h)
    a, b = ho.render_ids(0, msaa=1), hp.render_ids(0, msaa=1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and bits_equal(a[2], b[2]).all(), f"zoo {seed}: primary hits differ"
    for kw in (dict(integrator="path", msaa=1, max_depth=5), dict(integrator="direct", msaa=2, max_depth=5)):
        fa, sa = ho.render_samples(flags=1, **kw)
        fb, sb = hp.render_samples(flags=1, **kw)
        assert_radiance_close(fb, fa, f"zoo {seed} {kw}", outliers=3e-3)
        assert_stats_close(sb, sa, f"zoo {seed} {kw}", rel=3e-3)


o, h = oracle_ffi.load(), hs_load()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
zoo = len(sys.argv) > 3 and sys.argv[3] == "zoo"
for seed in range(lo, hi):
    for ext in ((None,) if zoo else (False, True)):
        try:
            if zoo:
                check_zoo(o, h, seed)
                continue
            _check(o, h, seed, ext=ext)
            sd = random_scene(seed, ext=ext)
            ho, hp = sd.realize(o), sd.realize(h)
            for kw in (dict(integrator="direct", msaa=1, max_depth=5), dict(integrator="path", msaa=2, max_depth=2),
                       dict(integrator="path", msaa=1, max_depth=6, extra=K.FLAG_NO_JITTER)):
                fl = K.FLAG_COUNT_TRAVERSAL | kw.pop("extra", 0)
                fa, sa = ho.render_samples(flags=fl, **kw)
                fb, sb = hp.render_samples(flags=fl, **kw)
                assert_radiance_close(fb, fa, f"seed {seed} {kw}", outliers=2e-3)
                assert_stats_close(sb, sa, f"seed {seed} {kw}", rel=2e-3)
        except AssertionError as e:
            bad += 1
            print("FAIL", seed, ext, str(e)[:200], flush=True)
print("done", lo, hi, "failures", bad)
sys.exit(1 if bad else 0)
