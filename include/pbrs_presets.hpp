// The reference's hard-coded scenes (scene/src/preset.rs) written against include/pbrs_gpu.hpp, the
// C++ mirror of the crates' constructors.  Random draws come from seeded generators where the
// reference uses thread_rng.  Used by pbrs_main (--scene_name) and the tests.
#pragma once
#include <cmath>

#include "pbrs_gpu.hpp"

namespace pbrs {
namespace preset {

inline ShapeRef quad(Point3 a, Point3 b, Point3 c, Point3 d) {
    return shape::TriangleMesh::from_soa({a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, d.x, d.y, d.z}, {}, {}, {0, 1, 2, 0, 2, 3});
}
inline ShapeRef box(Point3 lo, Point3 hi) {
    std::vector<float> P = {lo.x, lo.y, lo.z, hi.x, lo.y, lo.z, hi.x, hi.y, lo.z, lo.x, hi.y, lo.z, lo.x, lo.y, hi.z, hi.x, lo.y, hi.z, hi.x, hi.y, hi.z, lo.x, hi.y, hi.z};
    const uint32_t q[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};
    std::vector<uint32_t> idx;
    for (auto &f : q) { idx.insert(idx.end(), {f[0], f[1], f[2]}); idx.insert(idx.end(), {f[0], f[2], f[3]}); }
    return shape::TriangleMesh::from_soa(P, {}, {}, idx);
}

// preset::cornell_box, scene/src/preset.rs:194-257, as written: ParallelQuad walls, two Cuboids, a
// quad area light.  (Under the path integrator the reference panics here once a BSDF sample meets
// the light quad's mirrored extension; this back end counts would_panic[QUAD] and goes on.)
inline Scene cornell_box() {
    using Quad = shape::ParallelQuad;
    Camera camera({600, 600}, Angle::new_deg(40.0f));
    camera.look_at(point3(278, 278, -800), point3(278, 278, 0), Vec3::Y());
    MaterialRef red = mtl::Lambertian::solid({0.65f, 0.05f, 0.05f}), white = mtl::Lambertian::solid(Color::gray(0.73f)), green = mtl::Lambertian::solid({0.12f, 0.45f, 0.15f});
    const Color light_color = Color::gray(15.0f);
    MaterialRef light = mtl::DiffuseLight::create(light_color);
    const Quad light_quad = Quad::new_xz({213, 343}, 554, {227, 332});
    std::vector<Instance> inst = {
        Instance(Quad::new_yz(555, {0, 555}, {0, 555}), red), Instance(Quad::new_yz(0, {0, 555}, {0, 555}), green),  // mtl_seq order, preset.rs:230-232
        Instance(light_quad, light), Instance(Quad::new_xz({0, 555}, 0, {0, 555}), white), Instance(Quad::new_xz({0, 555}, 555, {0, 555}), white),
        Instance(Quad::new_xy({0, 555}, {0, 555}, 555), white),
        Instance(shape::Cuboid::from_points({0, 0, 0}, {165, 165, 165}), white).with_transform(AffineTransform::translater({265, 0, 105}) * AffineTransform::rotater(Vec3::Y(), Angle::new_deg(15))),
        Instance(shape::Cuboid::from_points({0, 0, 0}, {165, 330, 165}), white).with_transform(AffineTransform::translater({130, 0, 225}) * AffineTransform::rotater(Vec3::Y(), Angle::new_deg(-18))),
    };
    return Scene(std::move(inst), camera).with_lights({}, {light::DiffuseAreaLight(light_color, light::SamplableShape::Quad(light_quad))});
}

// preset::quad, scene/src/preset.rs:184-192 (Camera::new without look_at = identity orientation)
inline Scene quad_scene() {
    Camera camera({800, 800}, Angle::new_deg(45.0f));
    camera.look_at(point3(0, 0, 0), point3(0, 0, 1), Vec3::Y());
    std::vector<Instance> inst = {Instance(shape::ParallelQuad::new_xy({-0.5f, 0.5f}, {-0.3f, 0.6f}, 2.5f), mtl::Lambertian::solid({0.2f, 0.3f, 0.7f}))};
    return Scene(std::move(inst), camera).with_fn_env_light(light::EnvFn::BlueSky);
}

// preset::quad_light, scene/src/preset.rs:148-182 (seeded Perlin tables instead of thread_rng)
inline Scene quad_light() {
    Camera camera({800, 800}, Angle::new_deg(20.0f));
    camera.look_at(point3(26, 3, -6), point3(0, 2, 0), Vec3::Y());
    MaterialRef mtl = mtl::Lambertian::textured(tex::Perlin::with_freq(4.0f));
    const Color light_power = Color::gray(4.0f);
    MaterialRef light = mtl::DiffuseLight::create(light_power);
    const shape::ParallelQuad light_quad = shape::ParallelQuad::new_xy({3, 5}, {1, 3}, 2.1f);
    std::vector<Instance> inst = {
        Instance(shape::Sphere::from_raw(0, -1000, 0, 1000), mtl), Instance(shape::Sphere::from_raw(0, 2, 0, 2), mtl),
        Instance(light_quad, light), Instance(shape::Sphere::from_raw(0, 7, 0, 2), light),
    };
    return Scene(std::move(inst), camera).with_fn_env_light(light::EnvFn::DarkRoom)
        .with_lights({}, {light::DiffuseAreaLight(light_power, light::SamplableShape::Quad(light_quad)),
                          light::DiffuseAreaLight(light_power, light::SamplableShape::Sphere({0, 7, 0}, 2))});
}

// preset::two_perlin_spheres, scene/src/preset.rs:115-133
inline Scene two_perlin_spheres() {
    Camera camera({800, 800}, Angle::new_deg(20.0f));
    camera.look_at(point3(13, 2, -3), point3(0, 0, 0), Vec3::Y());
    MaterialRef mtl = mtl::Lambertian::textured(tex::Perlin::with_freq(4.0f));
    std::vector<Instance> inst = {Instance(shape::Sphere::from_raw(0, -1000, 0, 1000), mtl), Instance(shape::Sphere::from_raw(0, 2, 0, 2), mtl)};
    return Scene(std::move(inst), camera).with_fn_env_light(light::EnvFn::BlueSky);
}

// preset::plates, scene/src/preset.rs:259-358: four glossy plates tilted to mirror four sphere lights
// of decreasing size into the camera (the classic MIS test scene)
inline Scene plates() {
    const float r = 20.0f;
    auto hat = [](Vec3 v) { float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); return Vec3{v.x * inv, v.y * inv, v.z * inv}; };
    MaterialRef matte = mtl::Lambertian::solid(Color::gray(0.4f));
    std::vector<Instance> inst = {Instance(shape::ParallelQuad::new_xy({-r, r}, {0.0f, r}, 0.0f), matte),
                                  Instance(shape::ParallelQuad::new_xz({-r, r}, 0.0f, {-r, 0.0f}), matte)};
    const Point3 lights_pos = point3(0.0f, r, -0.4f * r), camera_pos = point3(0.0f, 0.4f * r, -2.8f * r);
    const float left = -r * 0.7f, right = r * 0.7f, plate_width = 0.16f * r;
    const float pos_yz[4][2] = {{0.6f * r, -0.2f * r}, {0.45f * r, -0.3f * r}, {0.3f * r, -0.45f * r}, {0.2f * r, -0.6f * r}};
    const float rough[4] = {8e-5f, 3e-4f, 8e-4f, 3e-3f};
    for (int k = 0; k < 4; ++k) {
        const float py = pos_yz[k][0], pz = pos_yz[k][1];
        const Vec3 pl = hat({0.0f, lights_pos.y - py, lights_pos.z - pz}), pc = hat({0.0f, camera_pos.y - py, camera_pos.z - pz});
        const Vec3 normal = hat({pl.x + pc.x, pl.y + pc.y, pl.z + pc.z});
        Vec3 tangent = hat({0.0f, normal.z, -normal.y});
        const float hw = plate_width * 0.5f;
        tangent = {tangent.x * hw, tangent.y * hw, tangent.z * hw};
        const Vec3 t00{left + tangent.x, py + tangent.y, pz + tangent.z}, t01{t00.x - tangent.x * 2.0f, t00.y - tangent.y * 2.0f, t00.z - tangent.z * 2.0f};
        const Vec3 t10{right + tangent.x, py + tangent.y, pz + tangent.z}, t11{t10.x - tangent.x * 2.0f, t10.y - tangent.y * 2.0f, t10.z - tangent.z * 2.0f};
        std::vector<float> P = {t00.x, t00.y, t00.z, t01.x, t01.y, t01.z, t10.x, t10.y, t10.z, t11.x, t11.y, t11.z}, N;
        for (int v = 0; v < 4; ++v) N.insert(N.end(), {normal.x, normal.y, normal.z});
        inst.emplace_back(shape::TriangleMesh::from_soa(P, N, {0, 0, 0, 1, 1, 0, 1, 1}, {0, 1, 2, 2, 1, 3}), mtl::Glossy::create(Color::gray(0.9f), rough[k]));
    }
    const float a = left * 0.9f, b = right * 0.9f, spacing = (b - a) * (1.0f / 4.0f);  // float::linspace, math/src/float.rs:140-155
    const float sizes[4] = {0.1f * r, 0.06f * r, 0.03f * r, 0.01f * r};
    const Color colors[4] = {{1.0f, 0.8f, 0.8f}, {1.0f, 1.0f, 0.8f}, {0.8f, 1.0f, 0.8f}, {0.8f, 0.8f, 1.0f}};
    std::vector<light::DiffuseAreaLight> area;
    for (int i = 0; i < 4; ++i) area.emplace_back(colors[i], light::SamplableShape::Sphere({spacing * (float(i) + 0.5f) + a, lights_pos.y, lights_pos.z}, sizes[i]));
    for (int i = 0; i < 4; ++i) inst.emplace_back(shape::Sphere::create({spacing * (float(i) + 0.5f) + a, lights_pos.y, lights_pos.z}, sizes[i]), mtl::DiffuseLight::create(colors[i]));
    Camera camera = Camera({1000, 800}, Angle::new_rad(3.14159265358979323846f * 0.19f))
                        .looking_at(camera_pos, point3(camera_pos.x, camera_pos.y, camera_pos.z + 1.0f), Vec3::Y());
    return Scene(std::move(inst), camera).with_lights({}, std::move(area));
}

// preset::everything, scene/src/preset.rs:360-442: a floor of 400 random-height Cuboids, a quad
// light, glass / metal / textured spheres and a rotated IsoBlas of 1000 small spheres.  Seeded
// draws in place of rand_f32(); a generated checker stands in for assets/earthmap.png.
inline Scene everything() {
    uint64_t state = 0x5EEDull;
    auto rnd = [&]() { state = state * 6364136223846793005ull + 1442695040888963407ull; return float((state >> 40) & 0xFFFFFF) / 16777216.0f; };
    MaterialRef ground = mtl::Lambertian::solid({0.48f, 0.83f, 0.53f});
    std::vector<Instance> inst;
    for (int i = 0; i < 20; ++i)
        for (int j = 0; j < 20; ++j) {
            const float x0 = -1000.0f + float(i) * 100.0f, z0 = -1000.0f + float(j) * 100.0f, y1 = rnd() * 100.0f + 1.0f;
            inst.emplace_back(shape::Cuboid::from_points(point3(x0, 0.0f, z0), point3(x0 + 100.0f, y1, z0 + 100.0f)), ground);
        }
    const Color L = Color::gray(7.0f);
    const shape::ParallelQuad light_quad = shape::ParallelQuad::new_xz({123, 423}, 554, {147, 412});
    inst.emplace_back(light_quad, mtl::DiffuseLight::create(L));
    inst.emplace_back(shape::Sphere::from_raw(260, 150, 45, 50), mtl::Dielectric::create(1.5f));
    const Color silver_r{0.155184f, 0.116681f, 0.138360f}, silver_i{4.828131f, 3.122411f, 2.147082f};  // preset.rs:467-472
    inst.emplace_back(shape::Sphere::from_raw(0, 150, 145, 50), mtl::Metal::from_ior(silver_r, silver_i, 1.0f));
    inst.emplace_back(shape::Sphere::from_raw(360, 150, 145, 70), mtl::Dielectric::create(1.5f));
    std::vector<uint8_t> img(256 * 256 * 3);
    for (int y = 0; y < 256; ++y)
        for (int x = 0; x < 256; ++x) {
            const bool c = ((x / 16) + (y / 16)) % 2 != 0;
            uint8_t *p = &img[3 * (y * 256 + x)];
            p[0] = c ? 190 : 70; p[1] = c ? 160 : 115; p[2] = c ? 95 : 55;
        }
    inst.emplace_back(shape::Sphere::from_raw(400, 200, 400, 100), mtl::Lambertian::textured(tex::Image::from_rgb8(256, 256, img)));
    inst.emplace_back(shape::Sphere::from_raw(220, 280, 300, 80), mtl::Lambertian::textured(tex::Perlin::with_freq(10.0f)));
    std::vector<shape::IsoBlas::Ball> balls(1000);
    for (auto &b : balls) b = {point3(rnd() * 165.0f, rnd() * 165.0f, rnd() * 165.0f), 10.0f};
    inst.push_back(Instance(shape::IsoBlas::build(balls), mtl::Lambertian::solid(Color::gray(0.73f)))
                       .with_transform(AffineTransform::translater({-100, 270, 395}) * AffineTransform::rotater(Vec3::Y(), Angle::new_deg(15))));
    Camera camera({800, 800}, Angle::new_deg(40.0f));
    camera.look_at(point3(478, 278, -600), point3(278, 278, 0), Vec3::Y());
    return Scene(std::move(inst), camera).with_fn_env_light(light::EnvFn::DarkRoom).with_lights({}, {light::DiffuseAreaLight(L, light::SamplableShape::Quad(light_quad))});
}

// The C1 workload of bench.py: the same box with triangle walls and a sphere light (what the
// loader can express, scene/src/loader.rs:396-434).
inline Scene cornell_box_mesh() {
    Camera camera({600, 600}, Angle::new_deg(40.0f));
    camera.look_at(point3(278, 278, -800), point3(278, 278, 0), Vec3::Y());
    MaterialRef red = mtl::Lambertian::solid({0.65f, 0.05f, 0.05f}), white = mtl::Lambertian::solid(Color::gray(0.73f)), green = mtl::Lambertian::solid({0.12f, 0.45f, 0.15f});
    const Color L{15, 15, 15};
    const float S = 555.0f;
    std::vector<Instance> inst = {
        Instance(quad({S, 0, 0}, {S, S, 0}, {S, S, S}, {S, 0, S}), green), Instance(quad({0, 0, 0}, {0, S, 0}, {0, S, S}, {0, 0, S}), red),
        Instance(quad({0, 0, 0}, {S, 0, 0}, {S, 0, S}, {0, 0, S}), white), Instance(quad({0, S, 0}, {S, S, 0}, {S, S, S}, {0, S, S}), white),
        Instance(quad({0, 0, S}, {S, 0, S}, {S, S, S}, {0, S, S}), white),
        Instance(box({0, 0, 0}, {165, 165, 165}), white).with_transform(AffineTransform::translater({265, 0, 105}) * AffineTransform::rotater(Vec3::Y(), Angle::new_deg(15))),
        Instance(box({0, 0, 0}, {165, 330, 165}), white).with_transform(AffineTransform::translater({130, 0, 225}) * AffineTransform::rotater(Vec3::Y(), Angle::new_deg(-18))),
        Instance(shape::Sphere::create({0, 0, 0}, 40), mtl::DiffuseLight::create(L)).with_transform(AffineTransform::translater({278, 514, 279.5f})),
    };
    return Scene(std::move(inst), camera).with_lights({}, {light::DiffuseAreaLight(L, light::SamplableShape::Sphere({278, 514, 279.5f}, 40))});
}

// preset::mixed_spheres (scene/src/preset.rs:55-113) with a seeded generator instead of thread_rng
inline Scene mixed_spheres() {
    Camera camera({1024, 768}, Angle::new_deg(25.0f));
    camera.look_at(point3(13, 2, 3), point3(0, 0, 0), Vec3::Y());
    uint64_t state = 0x5EEDull;
    auto rnd = [&]() { state = state * 6364136223846793005ull + 1442695040888963407ull; return float((state >> 40) & 0xFFFFFF) / 16777216.0f; };
    const Color gold_r{0.143176f, 0.373096f, 1.443834f}, gold_i{3.982675f, 2.387439f, 1.602465f};
    std::vector<Instance> inst = {
        Instance(shape::Sphere::from_raw(0, -1000, 1, 1000), mtl::Lambertian::solid(Color::gray(0.5f))), Instance(shape::Sphere::from_raw(0, 1, 0, 1), mtl::Dielectric::create(1.5f)),
        Instance(shape::Sphere::from_raw(-4, 1, 0, 1), mtl::Lambertian::solid({0.4f, 0.2f, 0.1f})), Instance(shape::Sphere::from_raw(4, 1, 0, 1), mtl::Metal::from_ior(gold_r, gold_i, 0.0f)),
    };
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            float choose = rnd(), h = rnd();
            Point3 c{a + 0.9f * rnd(), 0.2f + h * h * h * 0.1f, b + 0.9f * rnd()};
            float dx = c.x - 4, dy = c.y - 0.2f, dz = c.z;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) <= 0.9f) continue;
            MaterialRef m = choose < 0.8f ? mtl::Lambertian::solid({rnd(), rnd(), rnd()}) : choose < 0.95f ? mtl::Metal::from_ior(gold_r, gold_i, rnd() * 0.5f) : mtl::Dielectric::create(1.4f);
            inst.emplace_back(shape::Sphere::create(c, 0.2f), m);
        }
    return Scene(std::move(inst), camera).with_fn_env_light(light::EnvFn::BlueSky);
}

}  // namespace preset
}  // namespace pbrs
