"""Shared helpers of the parity tests: the small scene families and the comparison metrics."""
import numpy as np

from pbrs_b200 import scenes

# name -> SceneDesc factory: small members of the five BASELINE.json config families plus the
# "zoo" that covers every material / texture / light / environment kind of the ABI
SMALL_SCENES = {
    "cornell": lambda: scenes.cornell_box(96, 96),
    "spheres": lambda: scenes.spheres500(128, 72, n_small=120),
    "terrain": lambda: scenes.mesh_terrain(128, 72, grid=48, ico_subdiv=2, tex_size=64),
    "field": lambda: scenes.instanced_field(128, 72, n_side=10, n_meshes=3, ico_subdiv=1, n_lights=3),
    "zoo_image": lambda: scenes.material_zoo(96, 72, env="image", delta_lights=True),
    "zoo_dusk": lambda: scenes.material_zoo(96, 72, env="dusk", delta_lights=False),
    "zoo_black": lambda: scenes.material_zoo(96, 72, env="black", delta_lights=True),
    # the remaining Shapes of shape/src/simple.rs + IsoBlas<Sphere>
    "shape_zoo": lambda: scenes.shape_zoo(96, 72),
    "preset_cornell": lambda: scenes.preset_cornell_box(96, 96),
    "preset_quad_light": lambda: scenes.preset_quad_light(96, 72),
    "preset_everything": lambda: scenes.preset_everything(96, 72, n_balls=150, n_boxes=6),
    "preset_plates": lambda: scenes.preset_plates(100, 80),
}

def _edge_single():
    """One instance only (the TLAS root is a leaf), a sheared / non-uniformly scaled sphere (Q13),
    delta lights only."""
    from pbrs_b200.scene import SceneDesc
    sd = SceneDesc()
    sd.set_camera(96, 72, 50.0, (0.0, 0.5, -6.0), (0.0, 0.0, 0.0))
    fwd = np.array([[1.6, 0.3, 0.0, 0.2], [0.0, 0.8, 0.1, -0.1], [0.2, 0.0, 1.2, 0.3], [0, 0, 0, 1]], np.float64)
    sd.add_instance(sd.add_sphere((0.1, 0.0, 0.2), 1.0), sd.plastic((0.6, 0.3, 0.2), (0.4, 0.4, 0.4), 0.2), fwd=fwd)
    sd.add_point_light((3.0, 4.0, -4.0), (40.0, 40.0, 35.0))
    sd.add_distant_light((-0.4, -1.0, 0.3), (1.5, 1.4, 1.2))
    return sd


def _edge_mesh():
    """Traversal corner cases in one mesh scene: degenerate and duplicated (coincident) triangles,
    a leaf of more than four identical triangles (zero centroid extent), axis-aligned walls hit by
    rays with exactly-zero direction components (render with PBRS_FLAG_NO_JITTER), a mirror, and the
    light-selection quirk Q1 (two delta lights + one area light + environment)."""
    from pbrs_b200.scene import SceneDesc
    sd = SceneDesc()
    sd.set_camera(96, 72, 60.0, (0.0, 0.0, -5.0), (0.0, 0.0, 0.0))
    P = np.array([[-2, -2, 2], [2, -2, 2], [2, 2, 2], [-2, 2, 2],          # back wall z = 2 (two triangles)
                  [-2, -2, -1], [2, -2, -1], [2, -2, 2], [-2, -2, 2],      # floor y = -2
                  [0, 0, 1], [0, 0, 1], [0, 0, 1],                         # a degenerate triangle
                  [-1, -1, 1.5], [1, -1, 1.5], [0, 1, 1.5]], np.float32)  # a triangle that is listed 7 times
    idx = [[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10]] + [[11, 12, 13]] * 7 + [[0, 2, 1]]  # + a coincident copy of the wall
    sd.add_instance(sd.add_mesh(P, np.array(idx, np.uint32)), sd.lambertian((0.7, 0.7, 0.6)))
    Pm = np.array([[-2, -2, -1], [-2, 2, -1], [-2, 2, 2], [-2, -2, 2]], np.float32)  # mirror wall x = -2
    sd.add_instance(sd.add_mesh(Pm, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)), sd.mirror((0.9, 0.9, 0.9)))
    sd.add_instance(sd.add_sphere((1.0, -1.2, 0.5), 0.8), sd.dielectric(1.5))
    c, r, L = (0.0, 1.6, 0.0), 0.3, (30.0, 30.0, 30.0)
    sd.add_instance(sd.add_sphere(c, r), sd.diffuse_light(L))
    sd.add_area_light_sphere(c, r, L)
    sd.add_point_light((1.5, 1.5, -2.0), (8.0, 6.0, 6.0))
    sd.add_distant_light((0.2, -1.0, 0.5), (0.6, 0.6, 0.7))
    sd.set_env_constant((0.05, 0.06, 0.08))
    return sd


SMALL_SCENES["edge_single"] = _edge_single
SMALL_SCENES["edge_mesh"] = _edge_mesh

# Documented tolerance (DESIGN.md "Parity"): integer outcomes are bit-exact; radiance agrees to
# 1e-4 relative except where a last-ulp difference in a transcendental (glibc on the CPU, FP64
# libdevice rounded to FP32 on the GPU) is amplified (exp of a large Beckmann exponent) or flips a
# discrete choice -- a tiny, bounded fraction of samples.
REL_TOL = 1e-4
OUTLIER_FRACTION = 2e-4


def bits_equal(a, b):
    return a.view(np.uint32) == b.view(np.uint32)


def rel_err(a, b, floor=1e-6):
    """max over channels of |a-b| / max(|b|, floor); non-finite pairs compare by bit pattern."""
    a = np.asarray(a); b = np.asarray(b)
    fin = np.isfinite(a) & np.isfinite(b)
    r = np.where(fin, np.abs(a - b) / np.maximum(np.abs(b), floor), np.where(bits_equal(a, b) | (np.isnan(a) & np.isnan(b)), 0.0, np.inf))
    return r.max(axis=-1)


def assert_radiance_close(got, want, what, tol=REL_TOL, outliers=OUTLIER_FRACTION):
    r = rel_err(got, want)
    n_bad = int((r > tol).sum())
    allowed = int(np.ceil(outliers * r.size))
    assert n_bad <= allowed, f"{what}: {n_bad} of {r.size} differ by more than {tol} (allowed {allowed}); max {r.max():.3e}"
    return n_bad


def assert_stats_close(got, want, what, rel=2e-4):
    for k in ("n_samples", "n_rays_extend", "n_rays_shadow", "n_nodes", "n_tris", "n_spheres", "n_instances"):
        g, w = got[k], want[k]
        assert abs(g - w) <= max(2, rel * w), f"{what}: stat {k}: {g} vs {w}"
    for k in ("trav_extend", "trav_shadow"):
        for g, w in zip(got[k], want[k]):
            assert abs(g - w) <= max(2, rel * w), f"{what}: stat {k}: {got[k]} vs {want[k]}"
    assert set(got["would_panic"]) == set(want["would_panic"]), f"{what}: would_panic kinds {got['would_panic']} vs {want['would_panic']}"


