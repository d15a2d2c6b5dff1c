// pbrs_main -- the reference's driver (src/main.rs:56-246) with the B200 back end behind it.
//
//   pbrs_main --pbrt_file scene.pbrt [--integrator direct|path] [--msaa N]
//   pbrs_main --scene_name cornell_box|125_spheres ...
//
// Same command-line keys as src/cli_options.rs:52-59 (--use_multi_thread / --use_single_thread are
// accepted and ignored: there is one GPU path; --visualize_* are debug views outside the hot path).
// It parses the options, builds the Scene with the C++ mirror of the crates' constructors
// (include/pbrs_gpu.hpp, include/pbrs_scene_file.hpp), renders through the C ABI and writes
// "{scene}-{integrator}-{spp}spp.exr" like src/main.rs:238-245.  There is no CPU fallback: without
// a CUDA device the commit fails and the program exits non-zero with the library's message.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "../../../include/pbrs_gpu.hpp"
#include "../../../include/pbrs_scene_file.hpp"

using namespace pbrs;

namespace {

struct CliOptions {  // src/cli_options.rs:25-49
    bool use_multi_thread = true;
    std::string scene_name, pbrt_file;
    Integrator integrator = Integrator::Path;
    uint32_t msaa = 2;
};

bool parse_args(int argc, char **argv, CliOptions &o, std::string &err) {  // :63-115
    std::map<std::string, std::string> pairs;
    for (int i = 1; i < argc;) {
        std::string key = argv[i++];
        if (key.empty() || key[0] != '-') { err = "Unrecognized key " + key; return false; }
        if (i < argc && argv[i][0] != '-') pairs[key] = argv[i++]; else pairs[key] = "";
    }
    for (auto &kv : pairs) {
        const std::string &k = kv.first, &v = kv.second;
        if (k == "--use_multi_thread") o.use_multi_thread = true;
        else if (k == "--use_single_thread") o.use_multi_thread = false;
        else if (k == "--scene_name") o.scene_name = v;
        else if (k == "--pbrt_file") o.pbrt_file = v;
        else if (k == "--msaa") { if (v.empty()) { err = "'--msaa' should be followed by a number"; return false; } o.msaa = (uint32_t)std::strtoul(v.c_str(), nullptr, 10); }
        else if (k == "--integrator") {
            if (v == "direct") o.integrator = Integrator::Direct; else if (v == "path") o.integrator = Integrator::Path; else { err = "unsupported integrator"; return false; }
        } else if (k == "--help") {
            std::printf("usage:\n  --scene_name <scene_name>\n  --pbrt_file <file.pbrt>\n  --integrator <direct|path>\n  --msaa <N>\n");
            std::exit(0);
        } else { err = "Unrecognized key " + k; return false; }
    }
    return true;
}

ShapeRef quad(Point3 a, Point3 b, Point3 c, Point3 d) {
    return shape::TriangleMesh::from_soa({a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, d.x, d.y, d.z}, {}, {}, {0, 1, 2, 0, 2, 3});
}
ShapeRef box(Point3 lo, Point3 hi) {
    std::vector<float> P = {lo.x, lo.y, lo.z, hi.x, lo.y, lo.z, hi.x, hi.y, lo.z, lo.x, hi.y, lo.z, lo.x, lo.y, hi.z, hi.x, lo.y, hi.z, hi.x, hi.y, hi.z, lo.x, hi.y, hi.z};
    const uint32_t q[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};
    std::vector<uint32_t> idx;
    for (auto &f : q) { idx.insert(idx.end(), {f[0], f[1], f[2]}); idx.insert(idx.end(), {f[0], f[2], f[3]}); }
    return shape::TriangleMesh::from_soa(P, {}, {}, idx);
}

// preset::cornell_box (scene/src/preset.rs:194-257) with triangle walls and a sphere light: the
// reference's own uses ParallelQuad / Cuboid and a quad light, on which it panics.
Scene cornell_box() {
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
Scene mixed_spheres() {
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

}  // namespace

int main(int argc, char **argv) {
    CliOptions options;
    std::string err;
    if (!parse_args(argc, argv, options, err)) { std::fprintf(stderr, "Can't parse command-line options: %s\n", err.c_str()); return 2; }
    try {
        std::string scene_name;
        Scene scene = [&]() -> Scene {
            if (!options.pbrt_file.empty()) {
                std::string p = options.pbrt_file;
                size_t slash = p.find_last_of('/'), dot = p.find_last_of('.');
                scene_name = p.substr(slash == std::string::npos ? 0 : slash + 1, dot == std::string::npos ? std::string::npos : dot - (slash == std::string::npos ? 0 : slash + 1));
                return scene_file::build_scene(p);
            }
            scene_name = options.scene_name;
            if (options.scene_name == "cornell_box") return cornell_box();
            if (options.scene_name == "125_spheres") return mixed_spheres();
            std::fprintf(stderr, "No scene file or name specified. Abort.\nAvailable scenes: 125_spheres | cornell_box\n");
            std::exit(1);
        }();
        auto t0 = std::chrono::steady_clock::now();
        pbrs_stats st{};
        std::vector<Color> image_map = render(scene, options.integrator, options.msaa, &st);
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("whole render time = %.3fs (GPU %.1f ms, %llu samples, %llu rays)\n", secs, st.ms_total, (unsigned long long)st.n_samples,
                    (unsigned long long)(st.n_rays_extend + st.n_rays_shadow));
        std::string out = exr_file_name(scene_name, options.integrator, options.msaa);
        std::printf("Image written to %s\n", out.c_str());
        write_exr(out, image_map, scene.camera().resolution());
    } catch (const Error &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.code == PBRS_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
