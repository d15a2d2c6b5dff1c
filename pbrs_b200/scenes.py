"""Synthetic, seeded scene generators for the five BASELINE.json configs.

The reference ships no assets (`assets/` is git-ignored) and its presets draw from an OS-seeded
RNG (scene/src/preset.rs:18-20), so the workloads are generated here, deterministically, in the
vocabulary of scene/src/preset.rs and scene/src/loader.rs.  Every generator takes size knobs so
the parity tests can use small instances of the same family.
"""
import numpy as np

from . import _capi as K
from .scene import SceneDesc

SEED = 0x5EED

# metal IORs, scene/src/preset.rs:467-493
GOLD = ((0.143176, 0.373096, 1.443834), (3.982675, 2.387439, 1.602465))
SILVER = ((0.155184, 0.116681, 0.138360), (4.828131, 3.122411, 2.147082))
COPPER = ((0.195470, 0.925682, 1.102186), (3.910869, 2.451263, 2.142653))
ALUMINIUM = ((1.656937, 0.880173, 0.521201), (9.224230, 6.269670, 4.836996))


def translate(t):
    m = np.eye(4)
    m[:3, 3] = t
    return m


def rotate_y(deg):
    # AffineTransform::rotater(Y, angle), math/src/hcm.rs:508-520: X -> (cos, 0, sin), Z -> (-sin, 0, cos)
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[:3, 0] = (c, 0, s)
    m[:3, 2] = (-s, 0, c)
    return m


def rotate_axis(axis, rad):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    x, y, z = axis
    c, s = np.cos(rad), np.sin(rad)
    Kx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    R = np.eye(3) * c + s * Kx + (1 - c) * np.outer(axis, axis)
    m = np.eye(4)
    m[:3, :3] = R
    return m


def scale(s):
    m = np.eye(4)
    m[0, 0] = m[1, 1] = m[2, 2] = s
    return m


def _quad(p00, p10, p11, p01):
    """Two triangles over a quad; returns (P[4,3], idx[2,3])."""
    return np.array([p00, p10, p11, p01], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def _box(lo, hi):
    """Axis-aligned box as 12 triangles, outward winding not required (normals face the ray)."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    P = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0],
                  [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32)
    q = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
    idx = []
    for a, b, c, d in q:
        idx += [[a, b, c], [a, c, d]]
    return P, np.array(idx, np.uint32)


def cornell_box(width=512, height=512):
    """C1/C2: Cornell box, 34 triangles (5 walls x 2 + 2 boxes x 12) + a sphere area light.

    Colours from scene/src/preset.rs:196-199; camera from :249-254; the light is a sphere because
    the loader's area lights accept only sphere/plymesh (scene/src/loader.rs:396-434, Q16); it is
    built as the loader does (:176-195): Sphere at the origin under a Translate CTM, DiffuseLight
    instance + DiffuseAreaLight over the transformed shape.
    """
    sd = SceneDesc()
    sd.set_camera(width, height, 40.0, (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), (0, 1, 0))
    red = sd.lambertian((0.65, 0.05, 0.05))
    white = sd.lambertian((0.73, 0.73, 0.73))
    green = sd.lambertian((0.12, 0.45, 0.15))
    L = (15.0, 15.0, 15.0)
    light = sd.diffuse_light(L)
    S = 555.0
    walls = [
        (_quad((S, 0, 0), (S, S, 0), (S, S, S), (S, 0, S)), green),      # x = 555
        (_quad((0, 0, 0), (0, S, 0), (0, S, S), (0, 0, S)), red),        # x = 0
        (_quad((0, 0, 0), (S, 0, 0), (S, 0, S), (0, 0, S)), white),      # floor
        (_quad((0, S, 0), (S, S, 0), (S, S, S), (0, S, S)), white),      # ceiling
        (_quad((0, 0, S), (S, 0, S), (S, S, S), (0, S, S)), white),      # back
    ]
    for (P, idx), m in walls:
        sd.add_instance(sd.add_mesh(P, idx), m)
    P, idx = _box((0, 0, 0), (165, 165, 165))
    sd.add_instance(sd.add_mesh(P, idx), white, fwd=translate((265, 0, 105)) @ rotate_y(15.0))
    P, idx = _box((0, 0, 0), (165, 330, 165))
    sd.add_instance(sd.add_mesh(P, idx), white, fwd=translate((130, 0, 225)) @ rotate_y(-18.0))
    c, r = (278.0, 514.0, 279.5), 40.0
    sd.add_instance(sd.add_sphere((0, 0, 0), r), light, fwd=translate(c))
    sd.add_area_light_sphere(c, r, L)
    return sd


def _rotate_y_translate(deg, t):
    """identity().rotate_y(deg).translate(t) in the reference's own FP32 arithmetic
    (geometry/src/transform.rs:169-179): forward and inverse composed side by side."""
    from .pbrt_loader import Affine, _to_radians
    return Affine.translater(t) * Affine.rotater((0.0, 1.0, 0.0), _to_radians(deg))


def preset_cornell_box(width=600, height=600):
    """scene/src/preset.rs:194-257 as written: six ParallelQuads, two Cuboids under rotate_y +
    translate, one quad area light (Q11: under the path integrator the reference itself panics
    on this scene once a BSDF sample reaches the light's mirrored extension; those samples are
    counted in would_panic['quad'] here)."""
    sd = SceneDesc()
    sd.set_camera(width, height, 40.0, (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), (0, 1, 0))
    red = sd.lambertian((0.65, 0.05, 0.05))
    white = sd.lambertian((0.73, 0.73, 0.73))
    green = sd.lambertian((0.12, 0.45, 0.15))
    L = (15.0, 15.0, 15.0)
    light = sd.diffuse_light(L)
    shapes = [
        sd.add_quad_yz(555.0, (0.0, 555.0), (0.0, 555.0)),
        sd.add_quad_yz(0.0, (0.0, 555.0), (0.0, 555.0)),
        sd.add_quad_xz((213.0, 343.0), 554.0, (227.0, 332.0)),
        sd.add_quad_xz((0.0, 555.0), 0.0, (0.0, 555.0)),
        sd.add_quad_xz((0.0, 555.0), 555.0, (0.0, 555.0)),
        sd.add_quad_xy((0.0, 555.0), (0.0, 555.0), 555.0),
        sd.add_cuboid((0.0, 0.0, 0.0), (165.0, 165.0, 165.0)),
        sd.add_cuboid((0.0, 0.0, 0.0), (165.0, 330.0, 165.0)),
    ]
    mtls = [red, green, light, white, white, white, white, white]  # mtl_seq, preset.rs:230-232
    xf = {6: _rotate_y_translate(15.0, (265.0, 0.0, 105.0)), 7: _rotate_y_translate(-18.0, (130.0, 0.0, 225.0))}
    for i, (sh, m) in enumerate(zip(shapes, mtls)):
        t = xf.get(i)
        sd.add_instance(sh, m, fwd=t and t.fwd, inv=t and t.inv)
    sd.add_area_light_quad((213.0, 554.0, 227.0), (343.0 - 213.0, 0.0, 0.0), (0.0, 0.0, 332.0 - 227.0), L)
    return sd


def preset_quad(width=800, height=800):
    """scene/src/preset.rs:184-192: one quad under the blue sky, camera at its default pose."""
    sd = SceneDesc()
    # Camera::new without look_at (geometry/src/camera.rs:19-33): at the origin, identity
    # orientation = looking down +z, which look_at(origin -> +z, up = y) reproduces exactly
    sd.set_camera(width, height, 45.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0, 1, 0))
    sd.add_instance(sd.add_quad_xy((-0.5, 0.5), (-0.3, 0.6), 2.5), sd.lambertian((0.2, 0.3, 0.7)))
    sd.set_env_fn(K.ENV_BLUE_SKY)
    return sd


def preset_quad_light(width=800, height=800, seed=SEED):
    """scene/src/preset.rs:148-182: two perlin spheres lit by a quad light and a sphere light."""
    sd = SceneDesc()
    sd.set_camera(width, height, 20.0, (26.0, 3.0, -6.0), (0.0, 2.0, 0.0), (0, 1, 0))
    rv, px, py, pz = perlin_tables(seed)
    mtl = sd.lambertian(tex=sd.add_texture_perlin(4.0, rv, px, py, pz))
    L = (4.0, 4.0, 4.0)
    light = sd.diffuse_light(L)
    sd.add_instance(sd.add_sphere((0.0, -1000.0, 0.0), 1000.0), mtl)
    sd.add_instance(sd.add_sphere((0.0, 2.0, 0.0), 2.0), mtl)
    sd.add_instance(sd.add_quad_xy((3.0, 5.0), (1.0, 3.0), 2.1), light)
    sd.add_instance(sd.add_sphere((0.0, 7.0, 0.0), 2.0), light)
    sd.add_area_light_quad((3.0, 1.0, 2.1), (2.0, 0.0, 0.0), (0.0, 2.0, 0.0), L)
    sd.add_area_light_sphere((0.0, 7.0, 0.0), 2.0, L)
    sd.set_env_fn(K.ENV_DARK_ROOM)
    return sd


def preset_everything(width=800, height=800, n_balls=1000, n_boxes=20, seed=SEED):
    """scene/src/preset.rs:360-442 (seeded instead of OS-random, a generated image for
    assets/earthmap.png): a floor of random-height cuboids, a quad light, glass / metal / textured
    spheres and a rotated IsoBlas of `n_balls` small spheres."""
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(width, height, 40.0, (478.0, 278.0, -600.0), (278.0, 278.0, 0.0), (0, 1, 0))
    ground = sd.lambertian((0.48, 0.83, 0.53))
    w = 100.0 * 20 / n_boxes
    for i in range(n_boxes):
        for j in range(n_boxes):
            x0, z0 = -1000.0 + i * w, -1000.0 + j * w
            y1 = float(rng.uniform(1.0, 101.0))
            sd.add_instance(sd.add_cuboid((x0, 0.0, z0), (x0 + w, y1, z0 + w)), ground)
    L = (7.0, 7.0, 7.0)
    sd.add_instance(sd.add_quad_xz((123.0, 423.0), 554.0, (147.0, 412.0)), sd.diffuse_light(L))
    sd.add_area_light_quad((123.0, 554.0, 147.0), (300.0, 0.0, 0.0), (0.0, 0.0, 265.0), L)
    sd.add_instance(sd.add_sphere((400.0, 400.0, 200.0), 50.0), sd.lambertian((0.7, 0.3, 0.1)))
    sd.add_instance(sd.add_sphere((260.0, 150.0, 45.0), 50.0), sd.dielectric(1.5))
    sd.add_instance(sd.add_sphere((0.0, 150.0, 145.0), 50.0), sd.metal(SILVER[0], SILVER[1], 1.0))
    sd.add_instance(sd.add_sphere((360.0, 150.0, 145.0), 70.0), sd.dielectric(1.5))
    sd.add_instance(sd.add_sphere((400.0, 200.0, 400.0), 100.0),
                    sd.lambertian(tex=sd.add_texture_image(checker_noise_image(256, seed))))
    rv, px, py, pz = perlin_tables(seed)
    sd.add_instance(sd.add_sphere((220.0, 280.0, 300.0), 80.0), sd.lambertian(tex=sd.add_texture_perlin(10.0, rv, px, py, pz)))
    balls = np.concatenate([rng.uniform(0.0, 165.0, (n_balls, 3)), np.full((n_balls, 1), 10.0)], axis=1)
    t = _rotate_y_translate(15.0, (-100.0, 270.0, 395.0))
    sd.add_instance(sd.add_sphere_blas(balls), sd.lambertian((0.73, 0.73, 0.73)), fwd=t.fwd, inv=t.inv)
    sd.set_env_fn(K.ENV_DARK_ROOM)
    return sd


def preset_plates(width=1000, height=800):
    """scene/src/preset.rs:259-358: a wall and a floor quad, four glossy plates (two-triangle
    meshes tilted to mirror the lights into the camera, roughness 8e-5 .. 3e-3) and four sphere
    lights of decreasing size.  Vertex arithmetic in f32 like the reference's."""
    f = np.float32
    r = f(20.0)
    sd = SceneDesc()
    cam = np.array([0.0, f(0.4) * r, f(-2.8) * r], f)
    sd.set_camera(width, height, None, tuple(cam), tuple(cam + np.array([0, 0, 1], f)), (0, 1, 0),
                  fov_y_rad=f(np.pi) * f(0.19))  # Angle::pi() * 0.19
    matte = sd.lambertian((0.4, 0.4, 0.4))
    sd.add_instance(sd.add_quad_xy((-r, r), (0.0, r), 0.0), matte)
    sd.add_instance(sd.add_quad_xz((-r, r), 0.0, (-r, 0.0)), matte)
    lights_pos = np.array([0.0, r, f(-0.4) * r], f)
    left, right = -r * f(0.7), r * f(0.7)
    hat = lambda v: (v * (f(1.0) / np.sqrt(f(f(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])))).astype(f)
    plate_width = f(0.16) * r
    for (ky, kz), rough in zip([(0.6, -0.2), (0.45, -0.3), (0.3, -0.45), (0.2, -0.6)], [8e-5, 3e-4, 8e-4, 3e-3]):
        py, pz = f(ky) * r, f(kz) * r
        pl = np.array([0.0, lights_pos[1] - py, lights_pos[2] - pz], f)
        pc = np.array([0.0, cam[1] - py, cam[2] - pz], f)
        normal = hat(hat(pl) + hat(pc))
        tangent = hat(np.array([0.0, normal[2], -normal[1]], f)) * (plate_width * f(0.5))
        t00 = np.array([left, py, pz], f) + tangent
        t01 = t00 - tangent * f(2.0)
        t10 = np.array([right, py, pz], f) + tangent
        t11 = t10 - tangent * f(2.0)
        P = np.stack([t00, t01, t10, t11]).astype(f)
        mesh = sd.add_mesh(P, np.array([[0, 1, 2], [2, 1, 3]], np.uint32), N=np.tile(normal, (4, 1)),
                           UV=np.array([[0, 0], [0, 1], [1, 0], [1, 1]], f))
        sd.add_instance(mesh, sd.glossy((0.9, 0.9, 0.9), rough))
    a, b = left * f(0.9), right * f(0.9)
    spacing = (b - a) * f(1.0 / 4)                       # float::linspace, math/src/float.rs:140-155
    sizes = [f(0.1) * r, f(0.06) * r, f(0.03) * r, f(0.01) * r]
    colors = [(1.0, 0.8, 0.8), (1.0, 1.0, 0.8), (0.8, 1.0, 0.8), (0.8, 0.8, 1.0)]
    spheres = [((float(spacing * f(i + 0.5) + a), float(lights_pos[1]), float(lights_pos[2])), float(sizes[i])) for i in range(4)]
    for (c, rad), col in zip(spheres, colors):
        sd.add_area_light_sphere(c, rad, col)
    for (c, rad), col in zip(spheres, colors):
        sd.add_instance(sd.add_sphere(c, rad), sd.diffuse_light(col))
    return sd


def shape_zoo(width=128, height=96, seed=SEED):
    """Every Shape of shape/src/simple.rs and both BLAS kinds in one small scene, plus a disk and a
    quad area light: the parity fixture of the simple-shape code."""
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(width, height, 50.0, (0.0, 3.0, -9.0), (0.0, 1.0, 0.0), (0, 1, 0))
    grey = sd.lambertian((0.6, 0.6, 0.6))
    sd.add_instance(sd.add_quad_xz((-6.0, 6.0), 0.0, (-4.0, 6.0)), grey)                      # floor
    sd.add_instance(sd.add_quad((-6.0, 0.0, 6.0), (12.0, 0.0, 0.0), (1.5, 5.0, 0.5)), sd.lambertian((0.3, 0.4, 0.7)))  # slanted back
    sd.add_instance(sd.add_cuboid((-4.0, 0.0, 1.0), (-2.5, 1.5, 2.5)), sd.plastic((0.7, 0.2, 0.2), (0.3, 0.3, 0.3), 0.1),
                    fwd=translate((0.3, 0.0, 0.0)) @ rotate_y(20.0))
    sd.add_instance(sd.add_cuboid((2.0, 0.0, -1.0), (3.0, 2.5, 0.0)), sd.dielectric(1.5))
    sd.add_instance(sd.add_disk((0.0, 0.02, -1.0), (0.0, 2.0, 0.0), (1.2, 0.0, 0.0)), sd.metal(GOLD[0], GOLD[1], 0.3))
    sd.add_instance(sd.add_disk((0.5, 1.2, 3.0), (0.3, 0.2, -1.0), (0.0, 1.0, 0.2)),
                    sd.lambertian(tex=sd.add_texture_image(checker_noise_image(64, seed))),
                    fwd=translate((0.0, 0.3, 0.0)) @ scale(1.2))
    balls = np.concatenate([rng.uniform(-1.0, 1.0, (60, 3)) * (1.5, 0.8, 1.0) + (0.0, 1.6, 1.0), rng.uniform(0.08, 0.25, (60, 1))], axis=1)
    sd.add_instance(sd.add_sphere_blas(balls), sd.glossy((0.8, 0.7, 0.3), 0.2), fwd=translate((0.0, 0.0, 0.5)) @ rotate_y(-25.0))
    sd.add_instance(sd.add_sphere((-1.5, 0.6, -1.5), 0.6), sd.mirror((0.9, 0.9, 0.9)))
    Lq, Ld = (12.0, 11.0, 10.0), (20.0, 20.0, 25.0)
    qo, qu, qv = (-1.0, 4.5, 0.0), (2.0, 0.0, 0.0), (0.0, 0.0, 1.5)
    sd.add_instance(sd.add_quad(qo, qu, qv), sd.diffuse_light(Lq))
    sd.add_area_light_quad(qo, qu, qv, Lq)
    dc, dn, dr = (4.0, 3.0, 1.0), (-1.0, -0.5, 0.0), (0.0, 0.0, 0.6)
    sd.add_instance(sd.add_disk(dc, dn, dr), sd.diffuse_light(Ld))
    sd.add_area_light_disk(dc, dn, dr, Ld)
    sd.set_env_constant((0.02, 0.02, 0.03))
    return sd


def cornell_box_pbrt(width=512, height=512):
    """The same Cornell box as pbrt-v3-subset TEXT, the form BASELINE configs[0] names ("via
    scene_parser"): LookAt / Camera / Film, matte materials, `trianglemesh` walls and boxes under
    Translate / Rotate, and an `AreaLightSource "diffuse"` on a `sphere` (the loader's area lights
    accept only sphere / plymesh, scene/src/loader.rs:396-434)."""
    def mesh(P, idx):
        pts = " ".join(f"{v:g}" for v in np.asarray(P, np.float64).reshape(-1))
        ind = " ".join(str(int(v)) for v in np.asarray(idx).reshape(-1))
        return f'Shape "trianglemesh" "point P" [ {pts} ] "integer indices" [ {ind} ]'
    S = 555.0
    quads = [
        (_quad((S, 0, 0), (S, S, 0), (S, S, S), (S, 0, S)), "green"),
        (_quad((0, 0, 0), (0, S, 0), (0, S, S), (0, 0, S)), "red"),
        (_quad((0, 0, 0), (S, 0, 0), (S, 0, S), (0, 0, S)), "white"),
        (_quad((0, S, 0), (S, S, 0), (S, S, S), (0, S, S)), "white"),
        (_quad((0, 0, S), (S, 0, S), (S, S, S), (0, S, S)), "white"),
    ]
    out = [
        "# Cornell box, pbrt-v3 subset understood by pbrs (scene_parser + scene/src/loader.rs)",
        "LookAt 278 278 -800  278 278 0  0 1 0",
        'Camera "perspective" "float fov" [ 40 ]',
        f'Film "image" "integer xresolution" [ {width} ] "integer yresolution" [ {height} ]',
        "WorldBegin",
        'MakeNamedMaterial "red" "string type" "matte" "rgb Kd" [ 0.65 0.05 0.05 ]',
        'MakeNamedMaterial "white" "string type" "matte" "rgb Kd" [ 0.73 0.73 0.73 ]',
        'MakeNamedMaterial "green" "string type" "matte" "rgb Kd" [ 0.12 0.45 0.15 ]',
    ]
    for (P, idx), m in quads:
        out += ["AttributeBegin", f'  NamedMaterial "{m}"', "  " + mesh(P, idx), "AttributeEnd"]
    for (lo, hi, t, deg) in [((0, 0, 0), (165, 165, 165), (265, 0, 105), 15.0), ((0, 0, 0), (165, 330, 165), (130, 0, 225), -18.0)]:
        P, idx = _box(lo, hi)
        # pbrs negates Rotate angles (loader.rs:792-798), so the file carries the opposite sign
        out += ["AttributeBegin", '  NamedMaterial "white"', f"  Translate {t[0]} {t[1]} {t[2]}", f"  Rotate {-deg:g} 0 1 0", "  " + mesh(P, idx),
                "AttributeEnd"]
    out += ["AttributeBegin", '  AreaLightSource "diffuse" "rgb L" [ 15 15 15 ]', "  Translate 278 514 279.5", '  Shape "sphere" "float radius" [ 40 ]',
            "AttributeEnd", "WorldEnd", ""]
    return "\n".join(out)


def cornell_box_via_parser(width=512, height=512):
    from .pbrt_loader import load_pbrt_string
    return load_pbrt_string(cornell_box_pbrt(width, height))


def spheres500(width=1920, height=1080, n_small=496, seed=SEED):
    """C3: ground + 3 big + n_small small spheres, Lambertian/Metal/Dielectric, blue-sky env.

    In the style of preset::mixed_spheres (scene/src/preset.rs:55-113): each sphere is its own
    TLAS instance with the identity transform.
    """
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(width, height, 25.0, (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), (0, 1, 0))
    sd.add_instance(sd.add_sphere((0.0, -1000.0, 1.0), 1000.0), sd.lambertian((0.5, 0.5, 0.5)))
    sd.add_instance(sd.add_sphere((0.0, 1.0, 0.0), 1.0), sd.dielectric(1.5))
    sd.add_instance(sd.add_sphere((-4.0, 1.0, 0.0), 1.0), sd.lambertian((0.4, 0.2, 0.1)))
    sd.add_instance(sd.add_sphere((4.0, 1.0, 0.0), 1.0), sd.metal(GOLD[0], GOLD[1], 0.0))
    metals = [GOLD, SILVER, COPPER, ALUMINIUM]
    count = 0
    cells = [(a, b) for a in range(-12, 12) for b in range(-11, 10)]
    for a, b in cells:
        if count >= n_small:
            break
        choose = rng.random()
        center = np.array([a + 0.9 * rng.random(), 0.2 + 0.1 * rng.random() ** 3, b + 0.9 * rng.random()])
        if np.linalg.norm(center - np.array([4.0, 0.2, 0.0])) <= 0.9:
            continue
        if choose < 0.8:
            m = sd.lambertian(tuple(rng.random(3)))
        elif choose < 0.95:
            eta, k = metals[int(rng.integers(0, 4))]
            m = sd.metal(eta, k, float(rng.random() * 0.5))
        else:
            m = sd.dielectric(1.4)
        sd.add_instance(sd.add_sphere(tuple(center), 0.2), m)
        count += 1
    sd.set_env_fn(K.ENV_BLUE_SKY)
    return sd


def _value_noise(x, y, seed, octaves=4):
    """Seeded 2-D value noise, sum of octaves; x, y float arrays."""
    rng = np.random.default_rng(seed)
    tab = rng.random((256, 256))
    out = np.zeros_like(x, dtype=np.float64)
    amp, freq = 1.0, 1.0
    for _ in range(octaves):
        xf, yf = x * freq, y * freq
        x0, y0 = np.floor(xf).astype(int), np.floor(yf).astype(int)
        tx, ty = xf - x0, yf - y0
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        v00 = tab[x0 & 255, y0 & 255]; v10 = tab[(x0 + 1) & 255, y0 & 255]
        v01 = tab[x0 & 255, (y0 + 1) & 255]; v11 = tab[(x0 + 1) & 255, (y0 + 1) & 255]
        out += amp * ((v00 * (1 - tx) + v10 * tx) * (1 - ty) + (v01 * (1 - tx) + v11 * tx) * ty)
        amp *= 0.5
        freq *= 2.0
    return out


def compute_normals(P, idx):
    """geometry/src/lib.rs:16-32: area-weighted face normals summed per vertex, normalised."""
    P64 = P.astype(np.float64)
    n = np.cross(P64[idx[:, 1]] - P64[idx[:, 0]], P64[idx[:, 2]] - P64[idx[:, 0]])
    N = np.zeros_like(P64)
    for k in range(3):
        np.add.at(N, idx[:, k], n)
    N /= np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-30)
    return N.astype(np.float32)


def heightfield(grid, extent=40.0, height=3.0, seed=SEED):
    """(grid x grid quads) displaced height-field: P, N, UV, idx (2*grid^2 triangles)."""
    g = np.linspace(0.0, 1.0, grid + 1)
    U, V = np.meshgrid(g, g, indexing="ij")
    X = (U - 0.5) * extent
    Z = (V - 0.5) * extent
    Y = height * (_value_noise(U * 6.0, V * 6.0, seed) - 0.9)
    P = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    UV = np.stack([U, V], -1).reshape(-1, 2).astype(np.float32)
    i, j = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    v00 = (i * (grid + 1) + j).reshape(-1)
    v10 = v00 + (grid + 1)
    v01 = v00 + 1
    v11 = v10 + 1
    idx = np.concatenate([np.stack([v00, v01, v11], -1), np.stack([v00, v11, v10], -1)], 0).astype(np.uint32)
    # interleave the two triangles of each quad so that spatially close triangles are close in memory
    idx = idx.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    N = compute_normals(P, idx)
    # height-field normals must point up (+Y)
    N[N[:, 1] < 0] *= -1
    return P, N, UV, np.ascontiguousarray(idx)


def icosphere(subdiv, radius=1.0, bump=0.0, seed=0):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        edges = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], 0)
        edges.sort(axis=1)
        uniq, inv = np.unique(edges, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = V[uniq[:, 0]] + V[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = V.shape[0]
        V = np.concatenate([V, mid], 0)
        nf = F.shape[0]
        a, b, c = base + inv[:nf], base + inv[nf:2 * nf], base + inv[2 * nf:]
        F = np.concatenate([np.stack([F[:, 0], a, c], -1), np.stack([F[:, 1], b, a], -1),
                            np.stack([F[:, 2], c, b], -1), np.stack([a, b, c], -1)], 0)
    N = V.copy()
    if bump > 0.0:
        rng = np.random.default_rng(seed)
        k = rng.normal(size=(4, 3)) * 2.0
        ph = rng.random(4) * 6.28
        d = sum(np.sin(V @ k[i] + ph[i]) for i in range(4)) / 4.0
        V = V * (1.0 + bump * d)[:, None]
    P = (V * radius).astype(np.float32)
    theta = np.arccos(np.clip(N[:, 1], -1, 1)) / np.pi
    phi = (np.arctan2(N[:, 2], N[:, 0]) + np.pi) / (2 * np.pi)
    UV = np.stack([phi, theta], -1).astype(np.float32)
    idx = F.astype(np.uint32)
    Nn = compute_normals(P, idx) if bump > 0.0 else N.astype(np.float32)
    # make normals outward
    flip = np.sum(Nn * N, axis=1) < 0
    Nn[flip] *= -1
    return P, Nn, UV, idx


def checker_noise_image(size=1024, seed=SEED):
    g = np.arange(size)
    X, Y = np.meshgrid(g, g, indexing="ij")
    chk = ((X // (size // 16)) + (Y // (size // 16))) % 2
    n = _value_noise(X / 37.0, Y / 37.0, seed + 1)
    n = (n - n.min()) / (n.max() - n.min())
    base = np.where(chk[..., None] == 1, np.array([0.75, 0.62, 0.38]), np.array([0.28, 0.45, 0.22]))
    img = np.clip(base * (0.6 + 0.4 * n[..., None]), 0, 1)
    return (img * 255).astype(np.uint8)


def mesh_terrain(width=3840, height=2160, grid=708, ico_subdiv=5, tex_size=1024, seed=SEED):
    """C4: one TriangleMesh of 2*grid^2 triangles (708 -> 1,002,528) with an image-textured
    Lambertian, 8 floating icospheres (20*4^subdiv triangles each) in Plastic/Glossy, one sphere
    area light and a constant env of 0.1."""
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(width, height, 40.0, (0.0, 9.0, -30.0), (0.0, -1.0, 0.0), (0, 1, 0))
    P, N, UV, idx = heightfield(grid, seed=seed)
    tex = sd.add_texture_image(checker_noise_image(tex_size, seed))
    sd.add_instance(sd.add_mesh(P, idx, N=N, UV=UV), sd.lambertian(tex=tex))
    Pi, Ni, UVi, idxi = icosphere(ico_subdiv)
    ico = sd.add_mesh(Pi, idxi, N=Ni, UV=UVi)
    for k in range(8):
        pos = (float(rng.uniform(-14, 14)), float(rng.uniform(1.0, 5.0)), float(rng.uniform(-12, 12)))
        s = float(rng.uniform(0.8, 1.8))
        if k % 2 == 0:
            m = sd.plastic(tuple(rng.uniform(0.2, 0.8, 3)), (0.3, 0.3, 0.3), float(rng.uniform(0.05, 0.3)))
        else:
            m = sd.glossy(tuple(rng.uniform(0.5, 0.95, 3)), float(rng.uniform(0.01, 0.2)))
        sd.add_instance(ico, m, fwd=translate(pos) @ rotate_y(float(rng.uniform(0, 360))) @ scale(s))
    c, r, L = (6.0, 16.0, -6.0), 2.5, (40.0, 38.0, 34.0)
    sd.add_instance(sd.add_sphere((0, 0, 0), r), sd.diffuse_light(L), fwd=translate(c))
    sd.add_area_light_sphere(c, r, L)
    sd.set_env_constant((0.1, 0.1, 0.1))
    return sd


def instanced_field(width=3840, height=2160, n_side=100, n_meshes=10, ico_subdiv=3, n_lights=16, seed=SEED):
    """C5: n_side^2 instances (random rotation + uniform scale on a jittered grid) of n_meshes
    distinct bumpy icospheres (20*4^3 = 1280 triangles each; 10k instances ~ 12.8 M instanced
    triangles), a ground quad, n_lights sphere area lights, mixed materials."""
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    ext = float(n_side) * 2.5
    sd.set_camera(width, height, 40.0, (0.0, ext * 0.22, -ext * 0.62), (0.0, 0.0, -ext * 0.05), (0, 1, 0))
    Pg, ig = _quad((-ext, 0, -ext), (ext, 0, -ext), (ext, 0, ext), (-ext, 0, ext))
    sd.add_instance(sd.add_mesh(Pg, ig), sd.lambertian((0.45, 0.45, 0.42)))
    meshes = []
    for k in range(n_meshes):
        P, N, UV, idx = icosphere(ico_subdiv, bump=0.25, seed=seed + 100 + k)
        meshes.append(sd.add_mesh(P, idx, N=N, UV=UV))
    mats = []
    for k in range(24):
        c = rng.random()
        if c < 0.6:
            mats.append(sd.lambertian(tuple(rng.uniform(0.15, 0.9, 3))))
        elif c < 0.75:
            eta, kk = [GOLD, SILVER, COPPER, ALUMINIUM][int(rng.integers(0, 4))]
            mats.append(sd.metal(eta, kk, float(rng.uniform(0.02, 0.4))))
        elif c < 0.9:
            mats.append(sd.plastic(tuple(rng.uniform(0.2, 0.8, 3)), (0.25, 0.25, 0.25), float(rng.uniform(0.05, 0.3))))
        else:
            mats.append(sd.mirror((0.9, 0.9, 0.9)))
    cell = 2.0 * ext / n_side
    for i in range(n_side):
        for j in range(n_side):
            x = -ext + (i + 0.5 + rng.uniform(-0.3, 0.3)) * cell
            z = -ext + (j + 0.5 + rng.uniform(-0.3, 0.3)) * cell
            s = float(rng.uniform(0.6, 1.2)) * cell * 0.35
            axis = rng.normal(size=3)
            fwd = translate((x, s * 1.05, z)) @ rotate_axis(axis, float(rng.uniform(0, 6.28))) @ scale(s)
            sd.add_instance(meshes[int(rng.integers(0, n_meshes))], mats[int(rng.integers(0, len(mats)))], fwd=fwd)
    for k in range(n_lights):
        c = (float(rng.uniform(-ext * 0.8, ext * 0.8)), float(rng.uniform(ext * 0.12, ext * 0.3)),
             float(rng.uniform(-ext * 0.8, ext * 0.8)))
        r = float(rng.uniform(1.0, 2.5)) * max(1.0, n_side / 40.0)
        L = tuple(float(v) for v in rng.uniform(20.0, 60.0, 3))
        sd.add_instance(sd.add_sphere((0, 0, 0), r), sd.diffuse_light(L), fwd=translate(c))
        sd.add_area_light_sphere(c, r, L)
    return sd


# name -> (generator, integrator, msaa) : the BASELINE.json configs
def _wh(w, h, s):
    return max(2, int(round(w * s))), max(2, int(round(h * s)))


# `s` scales the frame (1.0 = the BASELINE.json size); used only to shorten profiling runs
CONFIGS = {
    "c1": (lambda s=1.0: cornell_box_via_parser(*_wh(512, 512, s)), "path", 4),
    "c2": (lambda s=1.0: cornell_box_via_parser(*_wh(1920, 1080, s)), "direct", 1),
    "c3": (lambda s=1.0: spheres500(*_wh(1920, 1080, s)), "path", 8),
    "c4": (lambda s=1.0: mesh_terrain(*_wh(3840, 2160, s)), "path", 16),
    "c5": (lambda s=1.0: instanced_field(*_wh(3840, 2160, s)), "path", 32),
}
WORKLOAD_NAMES = {
    "c1": "C1 cornell-box (pbrt text via the scene-file loader) 512x512 16spp path depth5",
    "c2": "C2 cornell-box (pbrt text via the scene-file loader) 1920x1080 1spp direct",
    "c3": "C3 500-spheres 1920x1080 64spp path depth5",
    "c4": "C4 1M-triangle terrain 3840x2160 256spp path depth5",
    "c5": "C5 10k-instance field (~12.8M tris) 3840x2160 1024spp path depth5",
}


def perlin_tables(seed=SEED):
    """Tables for Perlin::new (texture/src/lib.rs:66-96): 256 unit vectors, three permutations."""
    rng = np.random.default_rng(seed)
    v = rng.uniform(-1.0, 1.0, (256, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32), *(rng.permutation(256).astype(np.uint32) for _ in range(3))


def material_zoo(width=160, height=120, env="image", delta_lights=True, seed=SEED):
    """Every material, texture, light and environment kind of the ABI in one small scene: a ground
    quad (Perlin marble), a back wall (image texture), a row of spheres and icospheres in each
    material, a triangle area light + a sphere area light, a point and a distant light."""
    rng = np.random.default_rng(seed)
    sd = SceneDesc()
    sd.set_camera(width, height, 45.0, (0.0, 3.0, -9.0), (0.0, 1.0, 0.0), (0, 1, 0))
    rv, px, py, pz = perlin_tables(seed)
    marble = sd.add_texture_perlin(1.5, rv, px, py, pz)
    img = sd.add_texture_image(checker_noise_image(64, seed))
    Pg, ig = _quad((-8, 0, -8), (8, 0, -8), (8, 0, 8), (-8, 0, 8))
    UVg = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    sd.add_instance(sd.add_mesh(Pg, ig, UV=UVg), sd.lambertian(tex=marble))
    Pw, iw = _quad((-8, 0, 6), (8, 0, 6), (8, 7, 6), (-8, 7, 6))
    sd.add_instance(sd.add_mesh(Pw, iw, UV=UVg), sd.lambertian(tex=img))
    white = sd.add_texture_solid((0.8, 0.8, 0.8))
    spec = sd.add_texture_solid((0.4, 0.4, 0.4))
    dark = sd.add_texture_solid((0.0, 0.0, 0.0))
    mats = [
        sd.lambertian((0.7, 0.3, 0.2)),
        sd.metal(GOLD[0], GOLD[1], 0.15),
        sd.metal(SILVER[0], SILVER[1], 0.0),
        sd.glossy((0.6, 0.7, 0.9), 0.1),
        sd.mirror((0.9, 0.9, 0.9)),
        sd.dielectric(1.5),
        sd.plastic((0.2, 0.5, 0.3), (0.3, 0.3, 0.3), 0.1),
        sd.plastic((0.5, 0.2, 0.3), (0.3, 0.3, 0.3), 0.2, remap_roughness=False),
        sd.uber(img, spec, tex_kr=spec, tex_kt=-1, rough_u=0.1, rough_v=0.2, eta=1.4, opacity=1.0),
        sd.uber(white, spec, tex_kr=-1, tex_kt=spec, rough_u=0.05, rough_v=0.05, eta=1.3, opacity=0.6, remap_roughness=False),
        sd.substrate(img, spec),
        sd.uber(dark, dark, rough_u=0.1, rough_v=0.1),
    ]
    Pi, Ni, UVi, idxi = icosphere(2)
    ico = sd.add_mesh(Pi, idxi, N=Ni, UV=UVi)
    for k, m in enumerate(mats):
        x = -5.5 + 1.0 * k
        if k % 2 == 0:
            sd.add_instance(sd.add_sphere((x, 0.5, 0.0 + 0.3 * (k % 3)), 0.5), m)
        else:
            fwd = translate((x, 0.55, 0.5)) @ rotate_axis((0.3, 1.0, 0.2), 0.7 * k) @ np.diag([0.5, 0.55, 0.45, 1.0])
            sd.add_instance(ico, m, fwd=fwd)
    # a triangle area light (two triangles of an emissive quad) and a sphere area light
    L1 = (12.0, 11.0, 9.0)
    q = np.array([(-2, 5, -1), (2, 5, -1), (2, 5, 2), (-2, 5, 2)], np.float32)
    sd.add_instance(sd.add_mesh(q, np.array([[0, 2, 1], [0, 3, 2]], np.uint32)), sd.diffuse_light(L1))
    sd.add_area_light_triangle(q[0], q[2], q[1], L1)
    sd.add_area_light_triangle(q[0], q[3], q[2], L1)
    c, r, L2 = (4.5, 3.0, -2.0), 0.6, (20.0, 20.0, 25.0)
    sd.add_instance(sd.add_sphere((0, 0, 0), r), sd.diffuse_light(L2), fwd=translate(c))
    sd.add_area_light_sphere(c, r, L2)
    if delta_lights:
        sd.add_point_light((-5.0, 4.0, -3.0), (30.0, 25.0, 20.0))
        sd.add_distant_light((0.3, -1.0, 0.4), (0.8, 0.8, 0.9))
    if env == "image":
        g = np.linspace(0, 1, 32)
        sky = np.zeros((16, 32, 3), np.uint8)
        sky[..., 0] = (60 + 120 * g[None, :]).astype(np.uint8)
        sky[..., 1] = (90 + 100 * np.linspace(1, 0, 16)[:, None]).astype(np.uint8)
        sky[..., 2] = 200
        sd.set_env_image(sky, (0.6, 0.6, 0.7))
    elif env == "dusk":
        sd.set_env_fn(K.ENV_DUSK)
    elif env == "dark":
        sd.set_env_fn(K.ENV_DARK_ROOM)
    elif env == "black":
        sd.set_env_constant((0.0, 0.0, 0.0))
    else:
        sd.set_env_constant((0.3, 0.3, 0.35))
    return sd
