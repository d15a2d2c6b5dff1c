// The wavefront stages as per-path functions: generate, extend, shade, shadow, accumulate.
//
// One render call walks the frame in batches of paths resident in HBM.  A path is one
// (pixel, sample) integrator invocation of src/main.rs:197-205; its state lives in the SoA arrays
// of PathBuffers, indexed by the path's slot in the batch.  Per bounce the host enqueues
//     extend  (closest-hit walk for every queued path)              -> hit record
//     shade   (emission/env, lobes, light sampling, BSDF sampling)  -> shadow entry, next queue
//     shadow  (any-hit walks of the entry's <= 2 visibility rays)   -> radiance += direct term
// and after the last bounce `accumulate` sums each pixel's samples in sample order, exactly as the
// reference's sequential `color_sum` does (src/main.rs:195-208).
//
// The arithmetic order of every radiance update equals the reference's; the only thing the
// wavefront form changes is WHEN a visibility ray is traced (after shade, not inside it), which
// is why a shadow entry carries the candidate contributions instead of a boolean.
#pragma once
#include "device_shade.cuh"
#include "device_walk.cuh"

namespace pbrs {

struct PathBuffers {
    f4 *ray_o;  // origin.xyz, t_max
    f4 *ray_d;  // dir.xyz, -
    u4 *hit;    // t bits, instance, triangle record index, -
    f4 *beta;   // throughput rgb, w = flag bits (bit 0: last bounce was specular)
    f4 *rad;    // radiance rgb, -
    f4 *aux;    // direct integrator, specular stage: f.rgb, 1/mass
    // shadow entry of the path's current bounce
    f4 *sh_o1;  // ray 1 origin, t_max (< 0: absent)
    f4 *sh_d1;  // ray 1 dir, c1.r
    f4 *sh_o2;  // ray 2 origin, t_max (< 0: absent)
    f4 *sh_d2;  // ray 2 dir, c2.r
    f4 *sh_c;   // c1.g, c1.b, c2.g, c2.b
    f4 *sh_b;   // multiplier rgb (path: beta before the bounce; specular stage: f), 1/light_pdf
    float *sh_m;  // < 0: path mode (rad += beta * X); >= 0: specular stage (rad += (X * f) * m)
    // surface record of the split shade kernels (hit reconstruction -> scatter), 64 bytes per path:
    f4 *sf_p;   // pos.xyz, u
    f4 *sf_n;   // normal.xyz, v
    f4 *sf_w;   // wo.xyz, material id (bits)
    f4 *sf_t;   // tangent.xyz, t
    uint32_t *queue[2];            // extend queues (ping-pong between bounces)
    uint32_t *cls_queue[PBRS_NUM_CLS];  // shade queues, one per material class, filled by the extend kernel
    uint32_t *shadow_queue;
    uint32_t *counts;              // this batch's counter block (see PBRS_CNT_*)
    unsigned long long *stats;     // kStat* accumulators of the whole call
    uint32_t capacity;
};
// counter block layout: per batch, 16 words per stage
#define PBRS_MAX_STAGES 15
#define PBRS_CNT_STRIDE 16
#define PBRS_COUNTS_PER_BATCH (PBRS_CNT_STRIDE * (PBRS_MAX_STAGES + 1))
#define PBRS_CNT_EXTEND 0         // extend-queue length of the stage
#define PBRS_CNT_SHADOW 1         // shadow-queue length
#define PBRS_CNT_EXTEND_CURSOR 2  // work cursors the persistent extend / shadow kernels draw rays from
#define PBRS_CNT_SHADOW_CURSOR 3
#define PBRS_CNT_CLS 4            // [4 .. 4+PBRS_NUM_CLS): shade-queue lengths per material class
// [kStatTrav + 4*k + {0..3}] = nodes, tris, spheres, instances of the closest-hit (k=0) / any-hit (k=1) walks
enum { kStatShadowRays = 0, kStatTrav = 1, kStatPanic0 = 12, kStatCount = 28 };

struct FrameParams {
    unsigned long long seed;
    uint32_t msaa, spp, spp_r;    // spp_r: samples of each pixel owned by this rank
    uint32_t rank, world, split_samples;
    int32_t only_sample;          // >= 0: render just this sample index (render_ids)
    int32_t integrator, max_depth;
    uint32_t flags;
    uint32_t x0, y0, x1, y1;      // render region [x0,x1) x [y0,y1) (the crop)
    uint32_t width, height;
    const uint32_t *tiles;        // ids of the 64x64 tiles this rank renders
    uint32_t n_tiles;
};
struct BatchParams {
    uint32_t first_pixel;  // work-pixel index (tile-major) of the batch's first pixel
    uint32_t n_pixels;
    uint32_t n_paths;      // n_pixels * spp_r
};

// Morton decode of a 12-bit in-tile index: 8x4-pixel blocks stay together in a warp
PB_DEV uint32_t compact6(uint32_t v) {
    v &= 0x555u;
    v = (v | (v >> 1)) & 0x333u;
    v = (v | (v >> 2)) & 0x0F0Fu;
    v = (v | (v >> 4)) & 0x003Fu;
    return v;
}
struct PathId {
    uint32_t x, y, sample;
    bool valid;
};
PB_DEV PathId decode_pixel(const FrameParams &fp, uint32_t k) {
    uint32_t tile = ld_u32(fp.tiles + (k >> 12)), p = k & 4095u;
    uint32_t tiles_x = (fp.width + 63u) / 64u;
    PathId id;
    id.x = (tile % tiles_x) * 64u + compact6(p);
    id.y = (tile / tiles_x) * 64u + compact6(p >> 1);
    id.sample = 0;
    id.valid = id.x >= fp.x0 && id.x < fp.x1 && id.y >= fp.y0 && id.y < fp.y1;
    return id;
}
PB_DEV PathId decode_path(const FrameParams &fp, const BatchParams &bp, uint32_t j) {
    PathId id = decode_pixel(fp, bp.first_pixel + j / fp.spp_r);
    uint32_t si = j % fp.spp_r;
    id.sample = fp.only_sample >= 0 ? (uint32_t)fp.only_sample : (fp.split_samples ? si * fp.world + fp.rank : si);
    return id;
}
PB_DEV Sampler make_sampler(const FrameParams &fp, const PathId &id) {
    Sampler s;
    s.seed = fp.seed;
    s.pixel = id.y * fp.width + id.x;  // row * width + col
    s.sample = id.sample;
    return s;
}

PB_DEV void store_f4(f4 *p, float x, float y, float z, float w) {
    f4 v; v.x = x; v.y = y; v.z = z; v.w = w;
#ifdef __CUDA_ARCH__
    *reinterpret_cast<float4 *>(p) = make_float4(x, y, z, w);
#else
    *p = v;
#endif
}
PB_DEV f4 load_f4(const f4 *p) {
#ifdef __CUDA_ARCH__
    float4 v = *reinterpret_cast<const float4 *>(p);
    f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    return *p;
#endif
}

// ---------------------------------------------------------------------------------------------
// generate: src/main.rs:197-203 (stratified jitter) + Camera::shoot_ray (geometry/src/camera.rs:65-77,
// with orientation * {a, b, c} folded on the host).  Returns true if the path is live.
// ---------------------------------------------------------------------------------------------
PB_DEV bool stage_generate(const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j) {
    PathId id = decode_path(fp, bp, j);
    store_f4(pb.rad + j, 0.0f, 0.0f, 0.0f, 0.0f);
    if (!id.valid) return false;
    Sampler smp = make_sampler(fp, id);
    float dx = 0.0f, dy = 0.0f;
    if (!(fp.flags & PBRS_FLAG_NO_JITTER)) {
        float j0 = smp.f(0), j1 = smp.f(1);
        dx = ((float)(id.sample / fp.msaa) + j0) / (float)fp.msaa;
        dy = ((float)(id.sample % fp.msaa) + j1) / (float)fp.msaa;
    }
    float x = (float)id.x + fractf(dx);
    float y = (float)id.y + fractf(dy);
    vec3 a = mk(sc.cam.a[0], sc.cam.a[1], sc.cam.a[2]), b = mk(sc.cam.b[0], sc.cam.b[1], sc.cam.b[2]);
    vec3 c = mk(sc.cam.c[0], sc.cam.c[1], sc.cam.c[2]);
    vec3 dir = c + a * x + b * y;
    store_f4(pb.ray_o + j, sc.cam.center[0], sc.cam.center[1], sc.cam.center[2], PB_INF);
    store_f4(pb.ray_d + j, dir.x, dir.y, dir.z, 0.0f);
    store_f4(pb.beta + j, 1.0f, 1.0f, 1.0f, u2f(0u));
    return true;
}

PB_DEV Ray load_ray(const PathBuffers &pb, uint32_t j) {
    f4 o = load_f4(pb.ray_o + j), d = load_f4(pb.ray_d + j);
    Ray r;
    r.o = mk(o.x, o.y, o.z); r.d = mk(d.x, d.y, d.z); r.t_max = o.w;
    return r;
}

// ---------------------------------------------------------------------------------------------
// extend: scene.tlas.intersect(&mut ray) (tlas/src/bvh.rs:77).  The walk itself is device_walk.cuh;
// these are its two ends.
// ---------------------------------------------------------------------------------------------
PB_DEV void store_hit(const PathBuffers &pb, uint32_t j, const Hit &h) {
#ifdef __CUDA_ARCH__
    *reinterpret_cast<uint4 *>(pb.hit + j) = make_uint4(f2u(h.t), h.inst, h.tri, 0u);
#else
    u4 rec; rec.x = f2u(h.t); rec.y = h.inst; rec.z = h.tri; rec.w = 0u;
    pb.hit[j] = rec;
#endif
}
// which shade queue a finished closest-hit walk joins
PB_DEV uint32_t hit_class(const DeviceScene &sc, const Hit &h) {
    if (h.inst == PBRS_NONE) return PBRS_CLS_MISS;
    return ld_u32(&sc.inst_shade[h.inst].cls);
}
template <bool COUNT>
PB_DEV void stage_extend(const DeviceScene &sc, const PathBuffers &pb, uint32_t j, Diag &dg, TravCount &tc) {
    StackPair st_ent[PBRS_WALK_STACK];
    alignas(16) uint32_t st_park[PBRS_WALK_PARK];
    Walk<false, COUNT, true, ArrayStack<false>> w(ArrayStack<false>(st_ent, st_park));
    w.run(sc, load_ray(pb, j), dg, tc);
    store_hit(pb, j, w.best);
}

// Rebuilds the world-space Interaction of a recorded hit: Instance::intersect (tlas/src/instance.rs:50-72)
// = ray to object space, the shape's own intersect for the winning primitive, hit back to world
// (geometry/src/transform.rs:309-320).  Same inputs, same operations as during the walk, so the
// values equal what the reference computed when it found the hit.
PB_CALL_RECON void reconstruct_hit(const DeviceScene &sc, const Ray &world_ray, uint32_t inst, uint32_t tri, Isect &out, uint32_t &material,
                            Diag &dg) {
    Ray wr = world_ray;
    wr.t_max = PB_INF;
    uint32_t kind, index;
    Ray o = to_object(sc.inst_trav + inst, wr, kind, index);
    const char *sb = reinterpret_cast<const char *>(sc.inst_shade + inst);
    const f4 tail = ld16(sb + 48);
    Isect h;
    // one call site per primitive kind: this function is inlined into every shade kernel, which
    // are instruction-fetch bound (a second copy of sphere_intersect cost shade 3.5 % on C4)
    TriVerts tv;
    vec3 ball_c = mk(0.0f, 0.0f, 0.0f);
    float ball_r = 0.0f;
    bool is_ball = false;
    if (kind == PBRS_SHAPE_SPHERE) {
        f4 s = ld16(sc.spheres + index);
        ball_c = mk(s.x, s.y, s.z); ball_r = s.w; is_ball = true;
    } else if (kind == PBRS_SHAPE_MESH) {
        tv = load_tri(sc.tris + tri);
        if (tv.flags & PBRS_TRI_SPHERE) { ball_c = tv.p0; ball_r = tv.p1.x; is_ball = true; }  // a sphere of an IsoBlas<Sphere>
    }
    if (is_ball) {
        sphere_intersect(ball_c, ball_r, o, h, dg, f2u(tail.z) != 0u);
    } else if (kind == PBRS_SHAPE_MESH) {
        MeshHit mh;
        mh.pos = mk(0, 0, 0); mh.normal = mk(0, 0, 1); mh.dpdu = mk(1, 0, 0); mh.t = 0; mh.u = 0; mh.v = 0;
        if (!mesh_tri_shade(sc, tri, tv, o, mh, dg)) flag(dg, P_MISC);
        h = isect_new(mh.pos, mh.t, mh.u, mh.v, mh.normal, -o.d, dg);
        with_dpdu(h, mh.dpdu, dg);
    } else if (!simple_intersect(sc.simples + index, kind, o, h, dg)) {
        flag(dg, P_MISC);  // cannot happen: same inputs as during the walk
    }
    const char *tb = reinterpret_cast<const char *>(sc.inst_trav + inst);
    f4 i0 = ld16(tb), i1 = ld16(tb + 16), i2 = ld16(tb + 32);
    f4 f0 = ld16(sb), f1 = ld16(sb + 16), f2 = ld16(sb + 32);
    material = f2u(tail.x);
    vec3 new_pos = mk(row_pt(f0, h.pos), row_pt(f1, h.pos), row_pt(f2, h.pos));
    vec3 new_wo = mk(row_vec(f0, h.wo), row_vec(f1, h.wo), row_vec(f2, h.wo));
    // inverse.transpose() * normal: columns of the transpose are the rows of the inverse (three terms)
    vec3 new_normal = mk(i0.x, i0.y, i0.z) * h.normal.x + mk(i1.x, i1.y, i1.z) * h.normal.y + mk(i2.x, i2.y, i2.z) * h.normal.z;
    vec3 new_tan = mk(row_vec(f0, h.tangent), row_vec(f1, h.tangent), row_vec(f2, h.tangent));
    out = isect_new(new_pos, h.t, h.u, h.v, new_normal, new_wo, dg);
    with_dpdu(out, new_tan, dg);
}

// ---------------------------------------------------------------------------------------------
// uniform_sample_one_light, src/directlighting.rs:58-99, with the three estimators (:101-222).
// Fills the candidate contributions and their visibility rays; returns the number of rays.
// ---------------------------------------------------------------------------------------------
struct ShadowOut {
    Ray r1, r2;
    color c1, c2;
    float scale;
    bool has1, has2;
};
PB_DEV float power_heuristic2(float nf, float f_pdf, float ng, float g_pdf) {  // :224-232
    float f = nf * f_pdf, g = ng * g_pdf;
    return (f * f) / (f * f + g * g);
}
template <int K>
PB_DEV int sample_one_light_t(const DeviceScene &sc, const Isect &hit, const Lobes &L, const Frame &fr, const Sampler &smp, uint32_t base,
                              ShadowOut &so, Diag &dg) {
    so.has1 = so.has2 = false;
    uint32_t nd = sc.n_delta, na = sc.n_area;
    uint32_t n = nd + na + sc.has_env;
    if (n == 0u) return 0;
    float light_pdf = 1.0f / (float)n;
    uint32_t chosen = (uint32_t)(((unsigned long long)smp.u(base) * (unsigned long long)n) >> 32);
    float l0 = smp.f(base + 1), l1 = smp.f(base + 2);
    float s0 = smp.f(base + 3), s1 = smp.f(base + 4);
    so.scale = 1.0f / light_pdf;
    if (chosen < nd) {
        // estimate_direct_delta_light, :101-153
        if (L.n == 0) { flag(dg, P_EMPTY_BXDFS); return 0; }
        const DeltaLightRec &lt = sc.delta_lights[chosen];
        color lr;
        vec3 wi;
        Ray vis;
        if (lt.kind == PBRS_LIGHT_POINT) {  // light/src/lib.rs:70-75
            vec3 position = mk(lt.position[0], lt.position[1], lt.position[2]);
            lr = mkc(lt.color[0], lt.color[1], lt.color[2]) * weak_recip(len2(position - hit.pos));
            wi = hat(position - hit.pos, dg);
            vis = spawn_limited_ray_to(hit, position);
        } else {  // :76-89
            vec3 cd = mk(lt.casting_dir[0], lt.casting_dir[1], lt.casting_dir[2]);
            if (!(lt.world_radius > 0.0f)) flag(dg, P_MISC);
            vec3 outside_world = hit.pos - lt.world_radius * 2.0f * cd;
            vis = spawn_limited_ray_to(hit, outside_world);
            vec3 dummy = at(vis, vis.t_max);
            if (!(len(dummy - outside_world) < len(cd) * lt.world_radius * 0.01f)) flag(dg, P_MISC);
            lr = mkc(lt.color[0], lt.color[1], lt.color[2]);
            wi = -cd;
        }
        color bsdf_value = bsdf_eval<K>(fr, L, hit.wo, wi, dg) * fabsf(dot(hit.normal, wi));
        if (is_black(lr) || is_black(bsdf_value)) return 0;
        (void)bsdf_pdf<K>(fr, L, hit.wo, wi, dg);
        so.r1 = vis;
        so.c1 = bsdf_value * lr * 1.0f * weak_recip(1.0f);
        so.has1 = true;
        return 1;
    }
    if (chosen >= nd && chosen < na) {  // Q1: the bound is #area, not #delta + #area
        // estimate_direct_area_light, :155-222
        AreaLight lt = load_area_light(sc.area_lights + (chosen - nd));
        int rays = 0;
        {
            vec3 pol_pos, pol_n;
            area_shape_sample_towards(lt, hit, l0, l1, pol_pos, pol_n, dg);
            vec3 wi = hat(pol_pos - hit.pos, dg);
            color lr = !sign_neg(dot(pol_n, -wi)) ? lt.emit : blackc();  // radiance_from, light/src/lib.rs:127-133
            float pdf;
            if (!area_shape_pdf_at(lt, hit, wi, pdf, dg)) pdf = 0.0f;
            Ray vis = spawn_limited_ray_to(hit, pol_pos);
            if (pdf > 0.0f && !is_black(lr)) {
                color bsdf_value = bsdf_eval<K>(fr, L, hit.wo, wi, dg) * fabsf(dot(hit.normal, wi));
                float scatter_pdf = bsdf_pdf<K>(fr, L, hit.wo, wi, dg);
                if (!is_black(bsdf_value) && scatter_pdf > 0.0f) {
                    float weight = power_heuristic2(1.0f, pdf, 1.0f, scatter_pdf);
                    so.r1 = vis;
                    so.c1 = bsdf_value * lr * weight * weak_recip(pdf);
                    so.has1 = true;
                    rays++;
                }
            }
        }
        {
            color bv;
            vec3 wi2;
            Prob bp;
            bsdf_sample<K>(fr, L, hit.wo, s0, s1, bv, wi2, bp, dg);
            bv = bv * fabsf(dot(hit.normal, wi2));
            if (!(is_black(bv) || !(bp.v > 0.0f))) {
                // DiffuseAreaLight::radiance_to, light/src/lib.rs:141-146
                vec3 lpos, ln;
                float lpdf;
                if (area_shape_intersect(lt, spawn_ray(hit, wi2), lpos, ln, dg) && area_shape_pdf_at(lt, hit, wi2, lpdf, dg)) {
                    Ray vis2 = spawn_limited_ray_to(hit, lpos);
                    if (!(is_black(lt.emit) || lpdf <= 0.0f)) {
                        float weight = bp.is_mass ? 1.0f : power_heuristic2(1.0f, bp.v, 1.0f, lpdf);
                        so.r2 = vis2;
                        so.c2 = weight * (bv * lt.emit) * weak_recip(bp.v);
                        so.has2 = true;
                        rays++;
                    }
                }
            }
        }
        return rays;
    }
    // the environment, :80-96
    if (L.n == 0) { flag(dg, P_EMPTY_BXDFS); return 0; }
    color f;
    vec3 wi;
    Prob pr;
    bsdf_sample<K>(fr, L, hit.wo, s0, s1, f, wi, pr, dg);
    Ray incident = spawn_ray(hit, wi);
    so.r1 = incident;
    so.c1 = eval_env(sc, incident.d, dg) * f * fabsf(dot(wi, hit.normal)) * weak_recip(pr.v);
    so.has1 = true;
    return 1;
}

PB_CALL_DYN int sample_one_light_dyn(const DeviceScene &sc, const Isect &hit, const Lobes &L, const Frame &fr, const Sampler &smp, uint32_t base,
                                ShadowOut &so, Diag &dg) {
    return sample_one_light_t<-1>(sc, hit, L, fr, smp, base, so, dg);
}
template <int K>
PB_DEV int sample_one_light(const DeviceScene &sc, const Isect &hit, const Lobes &L, const Frame &fr, const Sampler &smp, uint32_t base,
                            ShadowOut &so, Diag &dg) {
    if constexpr (K < 0) return sample_one_light_dyn(sc, hit, L, fr, smp, base, so, dg);
    else return sample_one_light_t<K>(sc, hit, L, fr, smp, base, so, dg);
}

PB_DEV void store_shadow(const PathBuffers &pb, uint32_t j, const ShadowOut &so, color mult, float mode) {
    store_f4(pb.sh_o1 + j, so.r1.o.x, so.r1.o.y, so.r1.o.z, so.has1 ? so.r1.t_max : -1.0f);
    store_f4(pb.sh_d1 + j, so.r1.d.x, so.r1.d.y, so.r1.d.z, so.c1.r);
    store_f4(pb.sh_o2 + j, so.r2.o.x, so.r2.o.y, so.r2.o.z, so.has2 ? so.r2.t_max : -1.0f);
    store_f4(pb.sh_d2 + j, so.r2.d.x, so.r2.d.y, so.r2.d.z, so.c2.r);
    store_f4(pb.sh_c + j, so.c1.g, so.c1.b, so.c2.g, so.c2.b);
    store_f4(pb.sh_b + j, mult.r, mult.g, mult.b, so.scale);
    pb.sh_m[j] = mode;
}

struct ShadeOut {
    bool next;        // the path continues: its slot goes to the next extend queue
    int shadow_rays;  // > 0: its slot goes to the shadow queue
};

// ---------------------------------------------------------------------------------------------
// shade, path integrator: the body of the bounce loop, src/pathintegrator.rs:14-73
// ---------------------------------------------------------------------------------------------
// The two halves of the split shade kernels (kernels.cu PBRS_SHADE_SPLIT): `surface` = part 1 for
// every hit of a heavy material class, leaving the world-space Interaction in the path's 64-byte
// surface record; `scatter` = parts 2 and 3 from that record.  Same operations, same order as the
// one-piece body; the Interaction just crosses HBM as its own bit patterns.
PB_DEV void stage_shade_surface(const DeviceScene &sc, const PathBuffers &pb, uint32_t j, int bounce, Diag &dg) {
    Ray ray = load_ray(pb, j);
    u4 hr;
#ifdef __CUDA_ARCH__
    { uint4 v = *reinterpret_cast<const uint4 *>(pb.hit + j); hr.x = v.x; hr.y = v.y; hr.z = v.z; hr.w = v.w; }
#else
    hr = pb.hit[j];
#endif
    Isect h;
    uint32_t mtl_id = 0;
    reconstruct_hit(sc, ray, hr.y, hr.z, h, mtl_id, dg);
    f4 bt = load_f4(pb.beta + j);
    if (bounce == 0 || (f2u(bt.w) & 1u) != 0u) {  // :19-22
        f4 rd = load_f4(pb.rad + j);
        color beta = mkc(bt.x, bt.y, bt.z), radiance = mkc(rd.x, rd.y, rd.z);
        color env = eval_env(sc, ray.d, dg);
        (void)env;
        radiance = radiance + beta * mtl_emission(sc.materials[mtl_id]);
        store_f4(pb.rad + j, radiance.r, radiance.g, radiance.b, 0.0f);
    }
    store_f4(pb.sf_p + j, h.pos.x, h.pos.y, h.pos.z, h.u);
    store_f4(pb.sf_n + j, h.normal.x, h.normal.y, h.normal.z, h.v);
    store_f4(pb.sf_w + j, h.wo.x, h.wo.y, h.wo.z, u2f(mtl_id));
    store_f4(pb.sf_t + j, h.tangent.x, h.tangent.y, h.tangent.z, h.t);
}
template <int CLS>
PB_DEV ShadeOut stage_shade_scatter(const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j,
                                    int bounce, Diag &dg) {
    constexpr int K = cls_lobe_kind(CLS);
    ShadeOut out;
    out.next = false; out.shadow_rays = 0;
    const f4 sp = load_f4(pb.sf_p + j), sn = load_f4(pb.sf_n + j), sw = load_f4(pb.sf_w + j), st = load_f4(pb.sf_t + j);
    Isect h;
    h.pos = mk(sp.x, sp.y, sp.z); h.u = sp.w;
    h.normal = mk(sn.x, sn.y, sn.z); h.v = sn.w;
    h.wo = mk(sw.x, sw.y, sw.z);
    h.tangent = mk(st.x, st.y, st.z); h.t = st.w;
    const uint32_t mtl_id = f2u(sw.w);
    const f4 bt = load_f4(pb.beta + j), rdir = load_f4(pb.ray_d + j);
    color beta = mkc(bt.x, bt.y, bt.z);
    const vec3 ray_d = mk(rdir.x, rdir.y, rdir.z);
    Sampler smp = make_sampler(fp, decode_path(fp, bp, j));
    const uint32_t base = 2u + 8u * (uint32_t)bounce;
    const MaterialRec &m = sc.materials[mtl_id];
    Lobes L;
    bxdfs_at<CLS>(sc, m, h, L, dg);  // :31
    Frame fr = bsdf_frame(h, dg);
    ShadowOut so;
    out.shadow_rays = sample_one_light<K>(sc, h, L, fr, smp, base, so, dg);  // :35
    if (out.shadow_rays > 0) store_shadow(pb, j, so, beta, -1.0f);
    float r0 = smp.f(base + 5), r1 = smp.f(base + 6);  // :46
    color f;
    vec3 wi;
    Prob pr;
    bsdf_sample<K>(fr, L, -ray_d, r0, r1, f, wi, pr, dg);  // :47
    if (is_black(f) || pr.v == 0.0f) return out;  // :48
    bool specular_bounce = pr.is_mass;             // :55
    beta = beta * f * dot(wi, h.normal) * (1.0f / pr.v);  // :61 (Q7: signed cosine)
    Ray nr = spawn_ray(h, wi);                     // :62
    if (bounce > 3) {                              // :65-71
        float q = fmaxf(1.0f - luminance(beta), 0.05f);
        if (smp.f(base + 7) < q) return out;
        beta = beta * (1.0f / (1.0f - q));
    }
    if (bounce + 1 >= fp.max_depth) return out;
    store_f4(pb.ray_o + j, nr.o.x, nr.o.y, nr.o.z, nr.t_max);
    store_f4(pb.ray_d + j, nr.d.x, nr.d.y, nr.d.z, 0.0f);
    store_f4(pb.beta + j, beta.r, beta.g, beta.b, u2f(specular_bounce ? 1u : 0u));
    out.next = true;
    return out;
}

// The same body in one piece (the unsplit shade kernels and the host build).
template <int CLS>
PB_DEV ShadeOut stage_shade_path(const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j,
                                 int bounce, Diag &dg) {
    constexpr int K = cls_lobe_kind(CLS);
    ShadeOut out;
    out.next = false; out.shadow_rays = 0;
    Ray ray = load_ray(pb, j);
    u4 hr;
#ifdef __CUDA_ARCH__
    { uint4 v = *reinterpret_cast<const uint4 *>(pb.hit + j); hr.x = v.x; hr.y = v.y; hr.z = v.z; hr.w = v.w; }
#else
    hr = pb.hit[j];
#endif
    const bool hit = CLS == PBRS_CLS_ANY ? hr.y != 0xFFFFFFFFu : CLS != PBRS_CLS_MISS;
    f4 bt = load_f4(pb.beta + j), rd = load_f4(pb.rad + j);
    color beta = mkc(bt.x, bt.y, bt.z), radiance = mkc(rd.x, rd.y, rd.z);
    bool specular_bounce = (f2u(bt.w) & 1u) != 0u;
    Sampler smp = make_sampler(fp, decode_path(fp, bp, j));
    uint32_t base = 2u + 8u * (uint32_t)bounce;

    Isect h;
    uint32_t mtl_id = 0;
    if (hit) reconstruct_hit(sc, ray, hr.y, hr.z, h, mtl_id, dg);
    if (bounce == 0 || specular_bounce) {  // :19-22 (the environment is evaluated eagerly)
        color env = eval_env(sc, ray.d, dg);
        radiance = radiance + beta * (hit ? mtl_emission(sc.materials[mtl_id]) : env);
    }
    if (!hit) {
        store_f4(pb.rad + j, radiance.r, radiance.g, radiance.b, 0.0f);
        return out;
    }
    const MaterialRec &m = sc.materials[mtl_id];
    Lobes L;
    bxdfs_at<CLS>(sc, m, h, L, dg);  // :31
    Frame fr = bsdf_frame(h, dg);
    ShadowOut so;
    out.shadow_rays = sample_one_light<K>(sc, h, L, fr, smp, base, so, dg);  // :35
    if (out.shadow_rays > 0) store_shadow(pb, j, so, beta, -1.0f);
    float r0 = smp.f(base + 5), r1 = smp.f(base + 6);  // :46
    color f;
    vec3 wi;
    Prob pr;
    bsdf_sample<K>(fr, L, -ray.d, r0, r1, f, wi, pr, dg);  // :47
    store_f4(pb.rad + j, radiance.r, radiance.g, radiance.b, 0.0f);
    if (is_black(f) || pr.v == 0.0f) return out;  // :48
    specular_bounce = pr.is_mass;                  // :55
    beta = beta * f * dot(wi, h.normal) * (1.0f / pr.v);  // :61 (Q7: signed cosine)
    Ray nr = spawn_ray(h, wi);                     // :62
    if (bounce > 3) {                              // :65-71
        float q = fmaxf(1.0f - luminance(beta), 0.05f);
        if (smp.f(base + 7) < q) return out;
        beta = beta * (1.0f / (1.0f - q));
    }
    if (bounce + 1 >= fp.max_depth) return out;
    store_f4(pb.ray_o + j, nr.o.x, nr.o.y, nr.o.z, nr.t_max);
    store_f4(pb.ray_d + j, nr.d.x, nr.d.y, nr.d.z, 0.0f);
    store_f4(pb.beta + j, beta.r, beta.g, beta.b, u2f(specular_bounce ? 1u : 0u));
    out.next = true;
    return out;
}

// ---------------------------------------------------------------------------------------------
// shade, direct-lighting integrator: src/directlighting.rs:14-56.  Stage 0 is the primary hit,
// stage 1 the single specular bounce evaluated by direct_lighting_debug_integrator.
// ---------------------------------------------------------------------------------------------
template <int CLS>
PB_DEV ShadeOut stage_shade_direct(const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t j,
                                   int stage, Diag &dg) {
    constexpr int K = cls_lobe_kind(CLS);
    ShadeOut out;
    out.next = false; out.shadow_rays = 0;
    Ray ray = load_ray(pb, j);
    u4 hr;
#ifdef __CUDA_ARCH__
    { uint4 v = *reinterpret_cast<const uint4 *>(pb.hit + j); hr.x = v.x; hr.y = v.y; hr.z = v.z; hr.w = v.w; }
#else
    hr = pb.hit[j];
#endif
    const bool hit = CLS == PBRS_CLS_ANY ? hr.y != 0xFFFFFFFFu : CLS != PBRS_CLS_MISS;
    f4 rd = load_f4(pb.rad + j);
    color radiance = mkc(rd.x, rd.y, rd.z);
    Sampler smp = make_sampler(fp, decode_path(fp, bp, j));
    if (stage == 0) {
        if (!hit) {
            color env = eval_env(sc, ray.d, dg);
            store_f4(pb.rad + j, env.r, env.g, env.b, 0.0f);
            return out;
        }
        Isect h;
        uint32_t mtl_id;
        reconstruct_hit(sc, ray, hr.y, hr.z, h, mtl_id, dg);
        const MaterialRec &m = sc.materials[mtl_id];
        color e = mtl_emission(m);
        if (!is_black(e)) {
            store_f4(pb.rad + j, e.r, e.g, e.b, 0.0f);
            return out;
        }
        Lobes L;
        bxdfs_at<CLS>(sc, m, h, L, dg);
        Frame fr = bsdf_frame(h, dg);
        ShadowOut so;
        out.shadow_rays = sample_one_light<K>(sc, h, L, fr, smp, 2u, so, dg);
        if (out.shadow_rays > 0) store_shadow(pb, j, so, grayc(1.0f), -1.0f);
        color f;
        vec3 wi;
        Prob pr;
        if (bsdf_sample_specular<K>(fr, L, h.wo, f, wi, pr, dg)) {
            Ray refl = spawn_ray(h, wi);
            store_f4(pb.ray_o + j, refl.o.x, refl.o.y, refl.o.z, refl.t_max);
            store_f4(pb.ray_d + j, refl.d.x, refl.d.y, refl.d.z, 0.0f);
            store_f4(pb.aux + j, f.r, f.g, f.b, weak_recip(pr.is_mass ? pr.v : 0.0f));
            out.next = true;
        }
        return out;
    }
    f4 ax = load_f4(pb.aux + j);
    color f = mkc(ax.x, ax.y, ax.z);
    if (!hit) {
        color sr = eval_env(sc, ray.d, dg);
        radiance = radiance + sr * f * ax.w;
        store_f4(pb.rad + j, radiance.r, radiance.g, radiance.b, 0.0f);
        return out;
    }
    Isect h;
    uint32_t mtl_id;
    reconstruct_hit(sc, ray, hr.y, hr.z, h, mtl_id, dg);
    const MaterialRec &m = sc.materials[mtl_id];
    Lobes L;
    bxdfs_at<CLS>(sc, m, h, L, dg);
    Frame fr = bsdf_frame(h, dg);
    ShadowOut so;
    out.shadow_rays = sample_one_light<K>(sc, h, L, fr, smp, 10u, so, dg);
    if (out.shadow_rays > 0) store_shadow(pb, j, so, f, ax.w);
    return out;
}

// ---------------------------------------------------------------------------------------------
// shadow: scene.tlas.occludes(&vis_ray) for the entry's rays, then the radiance update of
// src/pathintegrator.rs:35 / src/directlighting.rs:36,44.
// ---------------------------------------------------------------------------------------------
// ray `which` (0/1) of path j's shadow entry; false if absent
PB_DEV bool shadow_ray(const PathBuffers &pb, uint32_t j, int which, Ray &r) {
    f4 o = load_f4((which ? pb.sh_o2 : pb.sh_o1) + j);
    if (o.w < 0.0f) return false;
    f4 d = load_f4((which ? pb.sh_d2 : pb.sh_d1) + j);
    r.o = mk(o.x, o.y, o.z); r.d = mk(d.x, d.y, d.z); r.t_max = o.w;
    return true;
}
// vis: bit 0 / 1 = ray 1 / 2 was traced and found unoccluded
PB_DEV void shadow_finish(const PathBuffers &pb, uint32_t j, uint32_t vis) {
    f4 d1 = load_f4(pb.sh_d1 + j), d2 = load_f4(pb.sh_d2 + j), cc = load_f4(pb.sh_c + j), mb = load_f4(pb.sh_b + j);
    float mode = pb.sh_m[j];
    color ld = blackc();
    if (vis & 1u) ld = ld + mkc(d1.w, cc.x, cc.y);
    if (vis & 2u) ld = ld + mkc(d2.w, cc.z, cc.w);
    color x = ld * mb.w;  // one_light_incident_radiance * (1.0 / light_pdf)
    f4 rd = load_f4(pb.rad + j);
    color radiance = mkc(rd.x, rd.y, rd.z), mult = mkc(mb.x, mb.y, mb.z);
    if (mode < 0.0f) radiance = radiance + mult * x;
    else radiance = radiance + x * mult * mode;
    store_f4(pb.rad + j, radiance.r, radiance.g, radiance.b, 0.0f);
}
template <bool COUNT>
PB_DEV void stage_shadow(const DeviceScene &sc, const PathBuffers &pb, uint32_t j, Diag &dg, TravCount &tc) {
    uint32_t vis = 0u;
    uint32_t st_ref[PBRS_WALK_STACK];
    alignas(16) uint32_t st_park[PBRS_WALK_PARK];
    for (int which = 0; which < 2; ++which) {
        Ray r;
        if (!shadow_ray(pb, j, which, r)) continue;
        Walk<true, COUNT, true, ArrayStack<true>> w(ArrayStack<true>(st_ref, st_park));
        w.run(sc, r, dg, tc);
        if (!w.occluded) vis |= 1u << which;
    }
    shadow_finish(pb, j, vis);
}

// ---------------------------------------------------------------------------------------------
// accumulate: src/main.rs:195-209.  `p` is the pixel's index in the batch.
// ---------------------------------------------------------------------------------------------
PB_DEV void stage_accumulate(const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp, uint32_t p, float *film) {
    PathId id = decode_pixel(fp, bp.first_pixel + p);
    if (!id.valid) return;
    color sum = blackc();
    for (uint32_t s = 0; s < fp.spp_r; ++s) {
        f4 r = load_f4(pb.rad + (size_t)p * fp.spp_r + s);
        sum = sum + mkc(r.x, r.y, r.z);
    }
    color px = (fp.flags & PBRS_FLAG_RAW_SUM) ? sum : sum * (1.0f / (float)fp.spp);
    float *o = film + 3u * ((size_t)id.y * fp.width + id.x);
    o[0] = px.r; o[1] = px.g; o[2] = px.b;
}

}  // namespace pbrs
