// pbrs_main -- the reference's driver (src/main.rs:56-246) with the B200 back end behind it.
//
//   pbrs_main --pbrt_file scene.pbrt [--integrator direct|path] [--msaa N]
//   pbrs_main --scene_name cornell_box|cornell_box_mesh|125_spheres|quad|quad_light|plates|two_perlin_spheres|everything ...
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
#include "../../../include/pbrs_presets.hpp"
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
            if (options.scene_name == "cornell_box") return preset::cornell_box();
            if (options.scene_name == "cornell_box_mesh") return preset::cornell_box_mesh();
            if (options.scene_name == "125_spheres") return preset::mixed_spheres();
            if (options.scene_name == "quad") return preset::quad_scene();
            if (options.scene_name == "quad_light") return preset::quad_light();
            if (options.scene_name == "plates") return preset::plates();
            if (options.scene_name == "everything") return preset::everything();
            if (options.scene_name == "two_perlin_spheres") return preset::two_perlin_spheres();
            std::fprintf(stderr, "No scene file or name specified. Abort.\nAvailable scenes: 125_spheres | cornell_box | cornell_box_mesh | quad | quad_light | plates | two_perlin_spheres | everything\n");
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
