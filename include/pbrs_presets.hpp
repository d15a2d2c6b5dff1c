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
