// TEST TOOLING: loads a pbrt-subset file with the C++ loader (include/pbrs_scene_file.hpp) or builds a
// C++ preset ("preset:<name>", include/pbrs_presets.hpp), commits it
// through the host build (libhostsim.so exports the product's pbrs_scene_* entry points) and dumps
// scene facts + the primary-hit ids, for comparison with the Python loader (tests/test_scene_io.py).
#include <cstdio>

#include <string>

#include "../../include/pbrs_presets.hpp"
#include "../../include/pbrs_scene_file.hpp"

extern "C" int hostsim_render_ids(const pbrs_scene *, const pbrs_render_opts *, uint32_t, uint32_t *, uint32_t *, float *);

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    try {
        const std::string what = argv[1];
        pbrs::Scene sc = what == "preset:cornell_box" ? pbrs::preset::cornell_box()
                         : what == "preset:quad" ? pbrs::preset::quad_scene()
                         : what == "preset:quad_light" ? pbrs::preset::quad_light()
                         : what == "preset:plates" ? pbrs::preset::plates()
                         : what == "preset:everything" ? pbrs::preset::everything()
                         : what == "preset:cornell_box_mesh" ? pbrs::preset::cornell_box_mesh()
                                                             : pbrs::scene_file::build_scene(what);
        pbrs_scene *s = sc.commit();
        pbrs_scene_info info;
        pbrs::check(pbrs_scene_get_info(s, &info), "get_info");
        size_t n = size_t(info.width) * info.height;
        std::vector<uint32_t> inst(n), prim(n);
        std::vector<float> t(n);
        pbrs_render_opts o{};
        o.integrator = PBRS_INTEGRATOR_PATH; o.msaa = 1; o.max_depth = 5; o.seed = 0x5EED; o.world_size = 1; o.flags = PBRS_FLAG_NO_JITTER;
        pbrs::check(hostsim_render_ids(s, &o, 0, inst.data(), prim.data(), t.data()), "render_ids");
        FILE *f = std::fopen(argv[2], "wb");
        uint32_t hdr[8] = {info.width, info.height, info.n_instances, info.n_meshes, info.n_spheres, info.n_triangles, info.n_lights, 0};
        std::fwrite(hdr, 4, 8, f);
        std::fwrite(inst.data(), 4, n, f); std::fwrite(prim.data(), 4, n, f); std::fwrite(t.data(), 4, n, f);
        std::fclose(f);
    } catch (const pbrs::Error &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.code == PBRS_ERR_UNSUPPORTED ? 5 : 1;
    }
    return 0;
}
