// TEST TOOLING ONLY -- not part of the product, never loaded by pbrs_b200/.
//
// Compiles the product's per-path stage functions (pbrs_b200/csrc/device_*.cuh, which are plain
// C++ behind the PB_DEV qualifier) and its host scene builder with g++ -ffp-contract=off, and
// runs the wavefront schedule sequentially, one path at a time.  Purpose: single-step the device
// logic against the oracle in the CPU-only container, where no GPU exists.  The scene-construction
// entry points are the product's own (api_scene.cpp); `device_upload` here just points a
// DeviceScene at host vectors.  Render entry points are named hostsim_* so that nothing can
// mistake this for the pbrs_render of include/pbrs_gpu.h.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../pbrs_b200/csrc/device_stages.cuh"
#include "../../pbrs_b200/csrc/scene_host.h"

using namespace pbrs;

struct pbrs_scene {
    SceneImpl impl;
};

namespace pbrs {
struct DeviceArrays {
    FlatScene flat;
};
void device_free(SceneImpl &s) {
    delete s.dev;
    s.dev = nullptr;
}
int device_upload(SceneImpl &s) {
    device_free(s);
    s.dev = new DeviceArrays();
    FlatScene &f = s.dev->flat;
    flatten_scene(s, f);
    DeviceScene &ds = s.dscene;
    std::memset(&ds, 0, sizeof ds);
    ds.tlas_nodes = s.tlas_nodes.data(); ds.blas_nodes = f.blas_nodes.data(); ds.tris = f.tris.data();
    ds.spheres = s.spheres.data(); ds.simples = s.simples.data(); ds.inst_trav = f.trav.data(); ds.inst_shade = f.shade.data(); ds.meshes = f.meshes.data();
    ds.tri_shade = f.tri_shade.data();
    ds.blas_node_parent = f.blas_node_parent.data(); ds.blas_leaf_parent = f.blas_leaf_parent.data();
    ds.tlas_node_parent = f.tlas_node_parent.data(); ds.tlas_leaf_parent = f.tlas_leaf_parent.data();
    ds.materials = s.materials.data(); ds.textures = f.textures.data(); ds.texels = f.texels.data();
    ds.perlin_vec = f.perlin_vec.data(); ds.perlin_perm = f.perlin_perm.data();
    ds.delta_lights = s.delta_lights.data(); ds.area_lights = s.area_lights.data();
    fill_scene_constants(s, f, ds);
    return 0;
}
}  // namespace pbrs

namespace {

struct Sim {
    std::vector<f4> a[16];
    std::vector<u4> hit;
    std::vector<float> sh_m;
    std::vector<uint32_t> q0, q1, sq, cq[PBRS_NUM_CLS];
    PathBuffers pb{};
    void resize(uint32_t n) {
        for (auto &v : a) v.assign(n, f4{0, 0, 0, 0});
        hit.assign(n, u4{0, 0, 0, 0}); sh_m.assign(n, 0.0f);
        q0.assign(n, 0); q1.assign(n, 0); sq.assign(n, 0);
        for (int c = 0; c < PBRS_NUM_CLS; ++c) { cq[c].assign(n, 0); pb.cls_queue[c] = cq[c].data(); }
        pb.ray_o = a[0].data(); pb.ray_d = a[1].data(); pb.hit = hit.data(); pb.beta = a[2].data(); pb.rad = a[3].data();
        pb.aux = a[4].data(); pb.sh_o1 = a[5].data(); pb.sh_d1 = a[6].data(); pb.sh_o2 = a[7].data(); pb.sh_d2 = a[8].data();
        pb.sh_c = a[9].data(); pb.sh_b = a[10].data(); pb.sh_m = sh_m.data();
        pb.sf_p = a[12].data(); pb.sf_n = a[13].data(); pb.sf_w = a[14].data(); pb.sf_t = a[15].data();
        pb.queue[0] = q0.data(); pb.queue[1] = q1.data(); pb.shadow_queue = sq.data();
        pb.capacity = n;
    }
};

struct Totals {
    uint64_t samples = 0, rays_extend = 0, rays_shadow = 0, nodes = 0, tris = 0, spheres = 0, insts = 0;
    uint64_t te[4] = {0, 0, 0, 0}, ts[4] = {0, 0, 0, 0};
    uint64_t panic[16] = {0};
    void add(const Totals &o) {
        samples += o.samples; rays_extend += o.rays_extend; rays_shadow += o.rays_shadow;
        nodes += o.nodes; tris += o.tris; spheres += o.spheres; insts += o.insts;
        for (int k = 0; k < 16; ++k) panic[k] += o.panic[k];
        for (int k = 0; k < 4; ++k) { te[k] += o.te[k]; ts[k] += o.ts[k]; }
    }
};
void note(Totals &t, Diag &dg) {
    for (int k = 0; k < 16; ++k) if (dg.panics & (1u << k)) t.panic[k]++;
    dg.panics = 0;
}

template <int CLS>
ShadeOut shade_cls(int integrator, const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j, int stage, Diag &dg) {
    return integrator == PBRS_INTEGRATOR_PATH ? stage_shade_path<CLS>(sc, pb, fp, bp, j, stage, dg) : stage_shade_direct<CLS>(sc, pb, fp, bp, j, stage, dg);
}
ShadeOut shade_dispatch(int cls, int integrator, const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j,
                        int stage, Diag &dg) {
    // like the kernels (PBRS_SHADE_SPLIT): the heavy classes of the path integrator go through the
    // surface record -- hit reconstruction first, then lobes / light sampling / BSDF sampling from it
    if (integrator == PBRS_INTEGRATOR_PATH && (cls == PBRS_CLS_LAMBERT || cls == PBRS_CLS_MICROFACET || cls == PBRS_CLS_MULTI)) {
        stage_shade_surface(sc, pb, j, stage, dg);
        if (cls == PBRS_CLS_LAMBERT) return stage_shade_scatter<PBRS_CLS_LAMBERT>(sc, pb, fp, bp, j, stage, dg);
        if (cls == PBRS_CLS_MICROFACET) return stage_shade_scatter<PBRS_CLS_MICROFACET>(sc, pb, fp, bp, j, stage, dg);
        return stage_shade_scatter<PBRS_CLS_MULTI>(sc, pb, fp, bp, j, stage, dg);
    }
    switch (cls) {
    case PBRS_CLS_MISS: return shade_cls<PBRS_CLS_MISS>(integrator, sc, pb, fp, bp, j, stage, dg);
    case PBRS_CLS_EMISSIVE: return shade_cls<PBRS_CLS_EMISSIVE>(integrator, sc, pb, fp, bp, j, stage, dg);
    case PBRS_CLS_LAMBERT: return shade_cls<PBRS_CLS_LAMBERT>(integrator, sc, pb, fp, bp, j, stage, dg);
    case PBRS_CLS_MICROFACET: return shade_cls<PBRS_CLS_MICROFACET>(integrator, sc, pb, fp, bp, j, stage, dg);
    case PBRS_CLS_SPECULAR: return shade_cls<PBRS_CLS_SPECULAR>(integrator, sc, pb, fp, bp, j, stage, dg);
    default: return shade_cls<PBRS_CLS_MULTI>(integrator, sc, pb, fp, bp, j, stage, dg);
    }
}

// mirrors render_frame (kernels.cu) for the tile range [tile_begin, tile_end) of the tile list
void run_tiles(const SceneImpl &s, FrameParams fp, const std::vector<uint32_t> &tiles, size_t t0, size_t t1, int n_stages, float *film,
               float *samples, uint32_t *ids_inst, uint32_t *ids_prim, float *ids_t, Totals &tot) {
    const DeviceScene &sc = s.dscene;
    Sim sim;
    std::vector<uint32_t> one(1);
    fp.tiles = one.data();
    fp.n_tiles = 1;
    sim.resize(4096u * std::max(fp.spp_r, 1u));
    for (size_t ti = t0; ti < t1; ++ti) {
        one[0] = tiles[ti];
        BatchParams bp; bp.first_pixel = 0; bp.n_pixels = 4096; bp.n_paths = 4096u * fp.spp_r;
        PathBuffers &pb = sim.pb;
        uint32_t n_in = 0;
        for (uint32_t j = 0; j < bp.n_paths; ++j) if (stage_generate(sc, pb, fp, bp, j)) pb.queue[0][n_in++] = j;
        tot.samples += n_in;
        for (int stage = 0; stage < n_stages; ++stage) {
            uint32_t *q_in = pb.queue[stage & 1], *q_out = pb.queue[(stage + 1) & 1];
            uint32_t n_out = 0, n_sh = 0;
            Diag dg; dg.panics = 0;
            TravCount tc{0, 0, 0, 0};
            tot.rays_extend += n_in;
            const bool counting = (fp.flags & PBRS_FLAG_COUNT_TRAVERSAL) != 0;  // like the kernels: two instantiations
            for (uint32_t i = 0; i < n_in; ++i) {
                if (counting) stage_extend<true>(sc, pb, q_in[i], dg, tc); else stage_extend<false>(sc, pb, q_in[i], dg, tc);
                note(tot, dg);
            }
            tot.te[0] += tc.nodes; tot.te[1] += tc.tris; tot.te[2] += tc.spheres; tot.te[3] += tc.insts;
            const TravCount te_snapshot = tc;
            if (fp.only_sample >= 0) { tot.nodes += tc.nodes; tot.tris += tc.tris; tot.spheres += tc.spheres; tot.insts += tc.insts; break; }
            // like the extend kernel: every finished walk joins the shade queue of its material class,
            // and each class runs its own specialisation of the shade stage
            uint32_t n_cls[PBRS_NUM_CLS] = {0};
            for (uint32_t i = 0; i < n_in; ++i) {
                uint32_t j = q_in[i];
                u4 hr = pb.hit[j];
                Hit h; h.t = u2f(hr.x); h.inst = hr.y; h.tri = hr.z;
                uint32_t c = hit_class(sc, h);
                pb.cls_queue[c][n_cls[c]++] = j;
            }
            for (int c = 0; c < PBRS_NUM_CLS; ++c)
                for (uint32_t i = 0; i < n_cls[c]; ++i) {
                    uint32_t j = pb.cls_queue[c][i];
                    ShadeOut so = shade_dispatch(c, fp.integrator, sc, pb, fp, bp, j, stage, dg);
                    note(tot, dg);
                    if (so.next) q_out[n_out++] = j;
                    if (so.shadow_rays > 0) pb.shadow_queue[n_sh++] = j;
                    tot.rays_shadow += (uint64_t)so.shadow_rays;
                }
            for (uint32_t i = 0; i < n_sh; ++i) {
                if (counting) stage_shadow<true>(sc, pb, pb.shadow_queue[i], dg, tc); else stage_shadow<false>(sc, pb, pb.shadow_queue[i], dg, tc);
                note(tot, dg);
            }
            tot.ts[0] += tc.nodes - te_snapshot.nodes; tot.ts[1] += tc.tris - te_snapshot.tris;
            tot.ts[2] += tc.spheres - te_snapshot.spheres; tot.ts[3] += tc.insts - te_snapshot.insts;
            tot.nodes += tc.nodes; tot.tris += tc.tris; tot.spheres += tc.spheres; tot.insts += tc.insts;
            n_in = n_out;
        }
        if (fp.only_sample >= 0) {
            for (uint32_t j = 0; j < bp.n_paths; ++j) {
                PathId id = decode_path(fp, bp, j);
                if (!id.valid) continue;
                u4 h = pb.hit[j];
                size_t k = (size_t)(id.y - fp.y0) * (fp.x1 - fp.x0) + (id.x - fp.x0);
                bool hit = h.y != 0xFFFFFFFFu;
                uint32_t prim = 0xFFFFFFFFu;
                if (hit) prim = sc.inst_trav[h.y].shape_kind == PBRS_SHAPE_MESH ? sc.tris[h.z].orig : 0u;
                if (ids_inst) ids_inst[k] = hit ? h.y : 0xFFFFFFFFu;
                if (ids_prim) ids_prim[k] = prim;
                if (ids_t) ids_t[k] = hit ? u2f(h.x) : PB_INF;
            }
        } else {
            if (film) for (uint32_t p = 0; p < bp.n_pixels; ++p) stage_accumulate(pb, fp, bp, p, film);
            if (samples)
                for (uint32_t j = 0; j < bp.n_paths; ++j) {
                    PathId id = decode_path(fp, bp, j);
                    if (!id.valid) continue;
                    f4 r = pb.rad[j];
                    size_t k = ((size_t)(id.y - fp.y0) * (fp.x1 - fp.x0) + (id.x - fp.x0)) * fp.spp + id.sample;
                    samples[3 * k] = r.x; samples[3 * k + 1] = r.y; samples[3 * k + 2] = r.z;
                }
        }
    }
}

int run(const pbrs_scene *scene, const pbrs_render_opts &o, float *film, float *samples, int32_t only_sample, uint32_t *ids_inst,
        uint32_t *ids_prim, float *ids_t, pbrs_stats *st) {
    const SceneImpl &s = scene->impl;
    if (!s.committed) return PBRS_ERR_STATE;
    const uint32_t W = s.cam.width, H = s.cam.height;
    FrameParams fp{};
    fp.seed = o.seed; fp.msaa = o.msaa ? o.msaa : 1; fp.spp = fp.msaa * fp.msaa;
    fp.rank = (uint32_t)o.rank; fp.world = (uint32_t)std::max(o.world_size, 1);
    fp.split_samples = (o.world_size > 1 && o.split == PBRS_SPLIT_SAMPLES) ? 1u : 0u;
    fp.only_sample = only_sample;
    fp.integrator = o.integrator; fp.max_depth = o.max_depth; fp.flags = o.flags;
    fp.width = W; fp.height = H;
    if (o.crop_w == 0 || o.crop_h == 0) { fp.x0 = 0; fp.y0 = 0; fp.x1 = W; fp.y1 = H; }
    else { fp.x0 = o.crop_x; fp.y0 = o.crop_y; fp.x1 = o.crop_x + o.crop_w; fp.y1 = o.crop_y + o.crop_h; }
    if (only_sample >= 0) fp.spp_r = 1;
    else if (fp.split_samples) fp.spp_r = fp.spp > fp.rank ? (fp.spp - fp.rank + fp.world - 1) / fp.world : 0;
    else fp.spp_r = fp.spp;
    const bool tile_split = o.world_size > 1 && o.split == PBRS_SPLIT_TILES;
    const uint32_t tiles_x = (W + 63) / 64;
    std::vector<uint32_t> tiles;
    for (uint32_t ty = fp.y0 / 64; ty <= (fp.y1 - 1) / 64; ++ty)
        for (uint32_t tx = fp.x0 / 64; tx <= (fp.x1 - 1) / 64; ++tx) {
            uint32_t t = ty * tiles_x + tx;
            if (tile_split && (t % fp.world) != fp.rank) continue;
            tiles.push_back(t);
        }
    int n_stages = o.integrator == PBRS_INTEGRATOR_PATH ? std::max(o.max_depth, 0) : (o.max_depth > 0 ? 2 : 0);
    if (only_sample >= 0) n_stages = 1;
    if (film) std::memset(film, 0, sizeof(float) * 3 * (size_t)W * H);
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::vector<Totals> tots(nt);
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            while (true) {
                size_t i = next.fetch_add(1);
                if (i >= tiles.size()) break;
                run_tiles(s, fp, tiles, i, i + 1, n_stages, film, samples, ids_inst, ids_prim, ids_t, tots[t]);
            }
        });
    for (auto &th : pool) th.join();
    Totals tot;
    for (auto &t : tots) tot.add(t);
    if (st) {
        std::memset(st, 0, sizeof *st);
        st->n_samples = tot.samples; st->n_rays_extend = tot.rays_extend; st->n_rays_shadow = tot.rays_shadow;
        st->n_nodes = tot.nodes; st->n_tris = tot.tris; st->n_spheres = tot.spheres; st->n_instances = tot.insts;
        for (int k = 0; k < 16; ++k) st->would_panic[k] = tot.panic[k];
        for (int k = 0; k < 4; ++k) { st->trav_extend[k] = tot.te[k]; st->trav_shadow[k] = tot.ts[k]; }
    }
    return 0;
}

}  // namespace

extern "C" {
int hostsim_render(const pbrs_scene *s, const pbrs_render_opts *o, float *out, pbrs_stats *st) {
    return run(s, *o, out, nullptr, -1, nullptr, nullptr, nullptr, st);
}
int hostsim_render_ids(const pbrs_scene *s, const pbrs_render_opts *o, uint32_t sample_index, uint32_t *oi, uint32_t *op, float *ot) {
    return run(s, *o, nullptr, nullptr, (int32_t)sample_index, oi, op, ot, nullptr);
}
int hostsim_render_samples(const pbrs_scene *s, const pbrs_render_opts *o, float *out, pbrs_stats *st) {
    return run(s, *o, nullptr, out, -1, nullptr, nullptr, nullptr, st);
}
}
