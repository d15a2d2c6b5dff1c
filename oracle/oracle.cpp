// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.h header).
// TLAS build/traversal, material lobes, BSDF, light estimators, the two integrators, the
// per-pixel sample loop, and the `oracle_*` C API (same shape as include/pbrs_gpu.h so that a
// scene description can be replayed into either side by tests/ and bench.py).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>

#include "../include/pbrs_gpu.h"
#include "oracle_scene.h"

namespace orc {

thread_local Diag *g_diag = nullptr;

struct Hit {
    Interaction isect;
    int inst;
    uint32_t prim;
};

// ---------------- instance: tlas/src/instance.rs:47-72 ----------------
static BBox shape_bbox(const Scene &sc, int shape_id) {
    const ShapeRef &s = sc.shapes[shape_id];
    switch (s.kind) {
    case SHAPE_SPHERE: return sphere_bbox(sc.spheres[s.index]);
    case SHAPE_QUAD: return quad_bbox(sc.quads[s.index]);
    case SHAPE_CUBOID: return cuboid_bbox(sc.cuboids[s.index]);
    case SHAPE_DISK: return disk_bbox(sc.disks[s.index]);
    case SHAPE_TRIANGLE: return isotri_bbox(sc.triangles[s.index]);
    default: return mesh_bbox(*sc.meshes[s.index]);
    }
}
static BBox instance_bbox(const Scene &sc, const Instance &in) {
    return affine_bbox(in.xf.fwd, shape_bbox(sc, in.shape_id));
}
static bool instance_intersect(const Scene &sc, const Instance &in, const Ray &ray, Hit *out) {
    if (g_diag) g_diag->n_instances++;
    Ray inv_ray = affine_ray(in.xf.inv, ray);
    if (!(norm_squared(inv_ray.dir) > 1e-3f)) panic_flag(P_MISC);
    const ShapeRef &s = sc.shapes[in.shape_id];
    Interaction hit;
    uint32_t prim = 0;
    if (s.kind == SHAPE_SPHERE) {
        if (g_diag) g_diag->n_spheres++;
        if (!sphere_intersect(sc.spheres[s.index], inv_ray, &hit)) return false;
    } else if (s.kind == SHAPE_QUAD) {
        if (!quad_intersect(sc.quads[s.index], inv_ray, &hit)) return false;
    } else if (s.kind == SHAPE_CUBOID) {
        if (!cuboid_intersect(sc.cuboids[s.index], inv_ray, &hit)) return false;
    } else if (s.kind == SHAPE_DISK) {
        if (!disk_intersect(sc.disks[s.index], inv_ray, &hit)) return false;
    } else if (s.kind == SHAPE_TRIANGLE) {
        if (g_diag) g_diag->n_tris++;
        if (!isotri_intersect(sc.triangles[s.index], inv_ray, &hit)) return false;
    } else {
        if (!mesh_intersect(*sc.meshes[s.index], inv_ray, &hit, &prim)) return false;
    }
    out->isect = affine_isect(in.xf, hit);
    out->inst = in.id;
    out->prim = prim;
    return true;
}
static bool instance_occludes(const Scene &sc, const Instance &in, const Ray &ray) {
    if (g_diag) g_diag->n_instances++;
    Ray inv_ray = affine_ray(in.xf.inv, ray);
    if (!(norm_squared(inv_ray.dir) > 1e-6f)) panic_flag(P_MISC);
    const ShapeRef &s = sc.shapes[in.shape_id];
    if (s.kind == SHAPE_SPHERE) {
        if (g_diag) g_diag->n_spheres++;
        return sphere_occludes(sc.spheres[s.index], inv_ray);
    }
    if (s.kind == SHAPE_QUAD) return quad_occludes(sc.quads[s.index], inv_ray);
    if (s.kind == SHAPE_CUBOID) return cuboid_occludes(sc.cuboids[s.index], inv_ray);
    if (s.kind == SHAPE_DISK) return disk_occludes(sc.disks[s.index], inv_ray);
    if (s.kind == SHAPE_TRIANGLE) {
        if (g_diag) g_diag->n_tris++;
        return isotri_occludes(sc.triangles[s.index], inv_ray);
    }
    return mesh_occludes(*sc.meshes[s.index], inv_ray);
}

// ---------------- TLAS: tlas/src/bvh.rs ----------------
// bvh.rs:116-152
static std::unique_ptr<TlasNode> tlas_build(Scene &sc, std::vector<int> instances) {
    auto node = std::make_unique<TlasNode>();
    if (instances.size() == 1) {
        node->is_leaf = true;
        node->inst = instances[0];
        node->bbox = instance_bbox(sc, sc.instances[instances[0]]);
        return node;
    }
    size_t num_all = instances.size();
    BBox all = bbox_empty();
    for (int i : instances) all = bbox_union(all, instance_bbox(sc, sc.instances[i]));
    int axis = max_dimension(bbox_diag(all));
    float split_plane = bbox_midpoint(all)[axis];
    std::vector<int> left, right;
    for (int i : instances) {
        if (bbox_midpoint(instance_bbox(sc, sc.instances[i]))[axis] < split_plane) left.push_back(i);
        else right.push_back(i);
    }
    if (left.empty()) {
        for (size_t k = 0; k < num_all / 2; ++k) { left.push_back(right.back()); right.pop_back(); }
    } else if (right.empty()) {
        for (size_t k = 0; k < num_all / 2; ++k) { right.push_back(left.back()); left.pop_back(); }
    }
    node->is_leaf = false;
    node->child[0] = tlas_build(sc, left);
    node->child[1] = tlas_build(sc, right);
    node->bbox = bbox_union(node->child[0]->bbox, node->child[1]->bbox);
    sc.n_tlas_inner++;
    return node;
}
// bvh.rs:77-103 (mutates ray.t_max: Q14)
static bool tlas_intersect(const Scene &sc, const TlasNode *n, Ray &ray, Hit *out) {
    if (!bbox_intersect(n->bbox, ray)) return false;
    if (n->is_leaf) return instance_intersect(sc, sc.instances[n->inst], ray, out);
    if (g_diag) g_diag->n_nodes++;
    Hit l, r;
    bool hl = tlas_intersect(sc, n->child[0].get(), ray, &l);
    if (hl) ray.t_max = l.isect.ray_t;
    bool hr = tlas_intersect(sc, n->child[1].get(), ray, &r);
    if (!hl && !hr) return false;
    if (hl && !hr) { *out = l; return true; }
    if (!hl && hr) { *out = r; return true; }
    *out = (l.isect.ray_t < r.isect.ray_t) ? l : r;
    return true;
}
// bvh.rs:105-113
static bool tlas_occludes(const Scene &sc, const TlasNode *n, const Ray &ray) {
    if (!bbox_intersect(n->bbox, ray)) return false;
    if (n->is_leaf) return instance_occludes(sc, sc.instances[n->inst], ray);
    if (g_diag) g_diag->n_nodes++;
    return tlas_occludes(sc, n->child[0].get(), ray) || tlas_occludes(sc, n->child[1].get(), ray);
}
struct TravSnapshot {
    uint64_t v[4];
    explicit TravSnapshot(const Diag &d) : v{d.n_nodes, d.n_tris, d.n_spheres, d.n_instances} {}
    void charge(const Diag &d, uint64_t *dst) const {
        dst[0] += d.n_nodes - v[0]; dst[1] += d.n_tris - v[1]; dst[2] += d.n_spheres - v[2]; dst[3] += d.n_instances - v[3];
    }
};
static bool scene_intersect(const Scene &sc, Ray &ray, Hit *out) {
    if (!g_diag) return tlas_intersect(sc, sc.tlas.get(), ray, out);
    g_diag->n_rays_extend++;
    TravSnapshot snap(*g_diag);
    bool r = tlas_intersect(sc, sc.tlas.get(), ray, out);
    snap.charge(*g_diag, g_diag->trav_extend);
    return r;
}
static bool scene_occludes(const Scene &sc, const Ray &ray) {
    if (!g_diag) return tlas_occludes(sc, sc.tlas.get(), ray);
    g_diag->n_rays_shadow++;
    TravSnapshot snap(*g_diag);
    bool r = tlas_occludes(sc, sc.tlas.get(), ray);
    snap.charge(*g_diag, g_diag->trav_shadow);
    return r;
}

// ---------------- environment: scene/src/lib.rs:96-117; scene/src/preset.rs:25-51 ----------------
static bool has_env_light(const Scene &sc) {
    if (sc.env_kind == ENV_CONSTANT) return !is_black(sc.env_color);
    return true;
}
static Color eval_env_light(const Scene &sc, const Ray &ray) {
    switch (sc.env_kind) {
    case ENV_CONSTANT: return sc.env_color;
    case ENV_IMAGE: {
        float phi = std::atan2(ray.dir.z, ray.dir.x);
        float u = f_fract(phi * kFrac1Pi * 0.5f + 1.0f);
        float cos_t = ray.dir.y / norm(ray.dir);
        float v = std::acos(cos_t) / kPi;
        return texture_value(sc.env_image, u, v, V3{0, 0, 0}) * sc.env_scale;
    }
    default:
        switch (sc.env_fn) {
        case PBRS_ENV_BLUE_SKY: {
            float y = (hat(ray.dir).y + 1.0f) * 0.5f;
            return rgb(0.5f, 0.7f, 1.0f) * y + gray(1.0f) * (1.0f - y);
        }
        case PBRS_ENV_DARK_ROOM: {
            float y = (hat(ray.dir).y + 1.0f) * 0.5f;
            return gray(0.1f) * y + gray(0.1f) * (1.0f - y);
        }
        default: {
            Color horizon = rgb8(245, 174, 82), dome = rgb8(109, 150, 204);
            float tilt = std::acos(hat(ray.dir).y);
            if (tilt > kPi * 0.25f) return dome;
            if (tilt > 0.0f) {
                float t = tilt / (kPi * 0.25f);
                return dome * t + horizon * (1.0f - t);
            }
            return gray(0.2f);
        }
        }
    }
}

// ---------------- materials -> lobes: material/src/lib.rs ----------------
static Color mtl_emission(const Material &m) {
    return m.kind == PBRS_MTL_DIFFUSE_LIGHT ? m.a : black();
}
static Color tex_at(const Scene &sc, int id, const Interaction &h) {
    return texture_value(sc.textures[id], h.u, h.v, h.pos);
}
static Lobes bxdfs_at(const Scene &sc, const Material &m, const Interaction &h) {
    Lobes L;
    L.n = 0;
    switch (m.kind) {
    case PBRS_MTL_LAMBERTIAN:  // :180-184
        L.l[L.n++] = lobe_lambert(tex_at(sc, m.tex_kd, h));
        break;
    case PBRS_MTL_METAL: {  // :200-206
        float alpha = roughness_to_alpha(m.f[0]);
        L.l[L.n++] = lobe_microfacet(gray(1.0f), Beckmann{alpha, alpha}, fresnel_conductor(m.a, m.b));
        break;
    }
    case PBRS_MTL_GLOSSY: {  // :72-78, :216-218
        float alpha = roughness_to_alpha(m.f[0]);
        L.l[L.n++] = lobe_microfacet(m.a, Beckmann{alpha, alpha}, fresnel_nop());
        break;
    }
    case PBRS_MTL_MIRROR:  // :229-232
        L.l[L.n++] = lobe_mirror(m.a);
        break;
    case PBRS_MTL_DIELECTRIC:  // :265-268
        L.l[L.n++] = lobe_dielectric(m.a, 1.0f, m.f[0]);
        break;
    case PBRS_MTL_DIFFUSE_LIGHT:  // :291-293
        break;
    case PBRS_MTL_PLASTIC: {  // :433-445
        float alpha = m.remap ? roughness_to_alpha(m.f[0]) : m.f[0];
        L.l[L.n++] = lobe_microfacet(m.b, Beckmann{alpha, alpha}, fresnel_nop());
        L.l[L.n++] = lobe_lambert(m.a);
        break;
    }
    case PBRS_MTL_UBER: {  // :317-365
        Color transmission = gray(f_clamp(1.0f - m.f[3], 0.0f, 1.0f));
        if (!is_black(transmission)) L.l[L.n++] = lobe_transmit(transmission, 1.0f, m.f[2]);
        Color kd = tex_at(sc, m.tex_kd, h);
        if (!is_black(kd)) L.l[L.n++] = lobe_lambert(kd);
        Color ks = tex_at(sc, m.tex_ks, h);
        if (!is_black(ks)) {
            float au = m.remap ? roughness_to_alpha(m.f[0]) : m.f[0];
            float av = m.remap ? roughness_to_alpha(m.f[1]) : m.f[1];
            L.l[L.n++] = lobe_microfacet(ks, Beckmann{au, av}, fresnel_dielectric(1.0f, m.f[2]));
        }
        if (m.tex_kr >= 0) {
            Color kr = tex_at(sc, m.tex_kr, h);
            if (!is_black(kr)) L.l[L.n++] = lobe_dielectric(kr, 1.0f, m.f[2]);
        }
        if (m.tex_kt >= 0) {
            Color kt = tex_at(sc, m.tex_kt, h);
            if (!is_black(kt)) L.l[L.n++] = lobe_transmit(kt, 1.0f, m.f[2]);
        }
        break;
    }
    case PBRS_MTL_SUBSTRATE: {  // :393-420
        Color d = tex_at(sc, m.tex_kd, h), s = tex_at(sc, m.tex_ks, h);
        if (!(is_black(d) && is_black(s))) L.l[L.n++] = lobe_lambert(d);
        break;
    }
    }
    return L;
}

// ---------------- BSDF: src/bsdf.rs ----------------
struct BSDF {
    M3 frame;
    const Lobes *lobes;
};
static bool frame_valid(const M3 &f) {  // bsdf.rs:125-137
    float det = dot(cross(f.c[0], f.c[1]), f.c[2]);
    return std::fabs(det - 1.0f) < 1e-4f;
}
static BSDF bsdf_new(const Interaction &isect, const Lobes *lobes) {  // bsdf.rs:18-31
    V3 normal = hat(isect.normal);
    V3 bitangent = hat(cross(isect.normal, isect.tbn.c[0]));
    V3 tangent = cross(bitangent, normal);
    if (!(std::fabs(dot(normal, bitangent)) < 1e-4f) || !(std::fabs(dot(normal, tangent)) < 1e-4f) ||
        !(std::fabs(dot(tangent, bitangent)) < 1e-4f))
        panic_flag(P_BSDF_FRAME);
    BSDF b{M3{{tangent, bitangent, normal}}, lobes};
    if (!frame_valid(b.frame)) panic_flag(P_BSDF_FRAME);
    return b;
}
static V3 world_to_local(const BSDF &b, V3 w) {  // bsdf.rs:113-117
    return hat(V3{dot(b.frame.c[0], w), dot(b.frame.c[1], w), dot(b.frame.c[2], w)});
}
static V3 local_to_world(const BSDF &b, V3 l) {  // bsdf.rs:119-123
    return l.x * b.frame.c[0] + l.y * b.frame.c[1] + l.z * b.frame.c[2];
}
static Color bsdf_eval(const BSDF &b, V3 wo_w, V3 wi_w) {  // bsdf.rs:43-51
    V3 wi = world_to_local(b, wi_w);
    V3 wo = world_to_local(b, wo_w);
    if (wo.z == 0.0f) return black();
    Color s = black();
    for (int i = 0; i < b.lobes->n; ++i) s = s + lobe_eval(b.lobes->l[i], wo, wi);
    return s;
}
static float bsdf_pdf(const BSDF &b, V3 wo_w, V3 wi_w) {  // bsdf.rs:53-57 (Q4: a sum)
    V3 wi = world_to_local(b, wi_w);
    V3 wo = world_to_local(b, wo_w);
    float s = 0.0f;
    for (int i = 0; i < b.lobes->n; ++i) s += density(lobe_prob(b.lobes->l[i], wo, wi));
    return s;
}
// bsdf.rs:59-103
static void bsdf_sample(const BSDF &b, V3 wo_world, float u, float v, Color *f, V3 *wi_out, Prob *pr) {
    if (!(u < 1.0f)) panic_flag(P_MISC);
    V3 wo = world_to_local(b, wo_world);
    int n = b.lobes->n;
    if (n == 0) { *f = black(); *wi_out = V3{0, 0, 0}; *pr = Mass(0.0f); return; }
    const Lobe *list[5];
    for (int i = 0; i < n; ++i) list[i] = &b.lobes->l[i];
    float un = u * (float)n;
    int chosen = (int)un;
    if (chosen >= n) chosen = n - 1;  // unreachable for u < 1; keeps the oracle memory-safe
    float remapped_u = f_fract(un);
    const Lobe *chosen_lobe = list[chosen];
    list[chosen] = list[n - 1];  // swap_remove
    int rest = n - 1;
    Color value;
    V3 wi;
    Prob prob;
    lobe_sample(*chosen_lobe, wo, v, remapped_u, &value, &wi, &prob);  // Q2: (v, remapped_u)
    if (prob.is_mass) { *f = value; *wi_out = local_to_world(b, wi); *pr = prob; return; }
    int count = 0;
    float other_sum = 0.0f;
    for (int i = 0; i < rest; ++i) {
        Prob p = lobe_prob(*list[i], wo, wi);
        if (!p.is_mass) { count++; other_sum += density(p); }
    }
    float overall = (density(prob) + other_sum) / (float)(1 + count);
    Color others = black();
    for (int i = 0; i < rest; ++i) others = others + lobe_eval(*list[i], wo, wi);
    *f = value + others;
    *wi_out = local_to_world(b, wi);
    *pr = Density(overall);
}
// bsdf.rs:104-112
static bool bsdf_sample_specular(const BSDF &b, V3 wo_world, Color *f, V3 *wi_out, Prob *pr) {
    V3 wo = world_to_local(b, wo_world);
    for (int i = 0; i < b.lobes->n; ++i) {
        if (b.lobes->l[i].kind == LOBE_SPECULAR) {
            V3 wi;
            lobe_sample(b.lobes->l[i], wo, 0.0f, 0.0f, f, &wi, pr);
            *wi_out = local_to_world(b, wi);
            return true;
        }
    }
    return false;
}

// ---------------- lights: light/src/lib.rs:66-92,141-172 ----------------
static void delta_sample(const DeltaLight &l, const Interaction &target, Color *rad, V3 *wi, Prob *pr, Ray *vis) {
    if (l.kind == DELTA_POINT) {
        *rad = l.intensity * weak_recip(squared_distance_to(l.position, target.pos));
        *wi = hat(l.position - target.pos);
        *vis = spawn_limited_ray_to(target, l.position);
        *pr = Mass(1.0f);
    } else {
        if (!(l.world_radius > 0.0f)) panic_flag(P_MISC);
        V3 outside_world = target.pos - l.world_radius * 2.0f * l.casting_dir;
        *vis = spawn_limited_ray_to(target, outside_world);
        V3 dummy = position_at(*vis, vis->t_max);
        if (!(distance_to(dummy, outside_world) < norm(l.casting_dir) * l.world_radius * 0.01f)) panic_flag(P_MISC);
        *rad = l.intensity;
        *wi = -l.casting_dir;
        *pr = Mass(1.0f);
    }
}
static void area_sample(const AreaLight &l, const Interaction &target, float u, float v, Color *rad, V3 *wi, Prob *pr, Ray *vis) {
    Interaction pol = area_shape_sample_towards(l, target, u, v);
    *wi = hat(pol.pos - target.pos);
    *rad = !std::signbit(dot(pol.normal, -*wi)) ? l.emit : black();  // radiance_from :127-133
    float pdf;
    if (!area_shape_pdf_at(l, target, *wi, &pdf)) pdf = 0.0f;
    *pr = Density(pdf);
    *vis = spawn_limited_ray_to(target, pol.pos);
}
static bool area_radiance_to(const AreaLight &l, const Interaction &target, V3 wi, Color *rad, float *pdf, Ray *vis) {
    Interaction light_hit;
    if (!area_shape_intersect(l, spawn_ray(target, wi), &light_hit)) return false;
    if (!area_shape_pdf_at(l, target, wi, pdf)) return false;
    *vis = spawn_limited_ray_to(target, light_hit.pos);
    *rad = l.emit;
    return true;
}

// ---------------- direct lighting: src/directlighting.rs ----------------
static float power_heuristic2(float nf, float f_pdf, float ng, float g_pdf) {  // :224-232
    float f = nf * f_pdf, g = ng * g_pdf;
    return f_powi(f, 2) / (f_powi(f, 2) + f_powi(g, 2));
}
static Color estimate_direct_delta(const Scene &sc, const Interaction &hit, const Material &m, const DeltaLight &light) {  // :101-153
    Lobes lobes = bxdfs_at(sc, m, hit);
    if (lobes.n == 0) { panic_flag(P_EMPTY_BXDFS); return black(); }
    BSDF bsdf = bsdf_new(hit, &lobes);
    Color lr; V3 wi; Prob lp; Ray vis;
    delta_sample(light, hit, &lr, &wi, &lp, &vis);
    Color bsdf_value = bsdf_eval(bsdf, hit.wo, wi) * std::fabs(dot(hit.normal, wi));
    if (!is_positive(lp) || is_black(lr) || is_black(bsdf_value)) return black();
    float scatter_pdf = bsdf_pdf(bsdf, hit.wo, wi);
    if (scene_occludes(sc, vis)) return black();
    float weight, pr;
    if (lp.is_mass) { weight = 1.0f; pr = lp.v; }
    else { weight = power_heuristic2(1.0f, lp.v, 1.0f, scatter_pdf); pr = lp.v; }
    return bsdf_value * lr * weight * weak_recip(pr);
}
static Color estimate_direct_area(const Scene &sc, const Interaction &hit, const Material &m, float s0, float s1,
                                  const AreaLight &light, float l0, float l1) {  // :155-222
    Color radiance_d = black();
    Lobes lobes = bxdfs_at(sc, m, hit);
    BSDF bsdf = bsdf_new(hit, &lobes);
    Color lr; V3 wi; Prob lp; Ray vis;
    area_sample(light, hit, l0, l1, &lr, &wi, &lp, &vis);
    if (is_positive(lp) && !is_black(lr)) {
        float light_pdf = density(lp);
        Color bsdf_value = bsdf_eval(bsdf, hit.wo, wi) * std::fabs(dot(hit.normal, wi));
        float scatter_pdf = bsdf_pdf(bsdf, hit.wo, wi);
        if (!is_black(bsdf_value) && scatter_pdf > 0.0f && !scene_occludes(sc, vis)) {
            float weight = power_heuristic2(1.0f, light_pdf, 1.0f, scatter_pdf);
            radiance_d = radiance_d + bsdf_value * lr * weight * weak_recip(light_pdf);
        }
    }
    {
        Color bv; V3 wi2; Prob bp;
        bsdf_sample(bsdf, hit.wo, s0, s1, &bv, &wi2, &bp);
        bv = bv * std::fabs(dot(hit.normal, wi2));
        if (!(is_black(bv) || !is_positive(bp))) {
            Color ir; float lpdf; Ray vis2;
            if (area_radiance_to(light, hit, wi2, &ir, &lpdf, &vis2)) {
                if (!(is_black(ir) || lpdf <= 0.0f || scene_occludes(sc, vis2))) {
                    float weight, pr;
                    if (bp.is_mass) { weight = 1.0f; pr = bp.v; }
                    else { weight = power_heuristic2(1.0f, bp.v, 1.0f, lpdf); pr = bp.v; }
                    radiance_d = radiance_d + weight * (bv * ir) * weak_recip(pr);
                }
            }
        }
    }
    return radiance_d;
}

struct SampleCtx {
    uint64_t seed;
    uint32_t pixel, sample;
    float draw(uint32_t dim) const { return u32_to_f32(sampler_u32(seed, pixel, sample, dim)); }
    uint32_t draw_u32(uint32_t dim) const { return sampler_u32(seed, pixel, sample, dim); }
};

// :58-99; `base` = first sampler dimension of this bounce (DESIGN.md "Sampler")
static Color uniform_sample_one_light(const Scene &sc, const Interaction &hit, const Material &m,
                                      const SampleCtx &ctx, uint32_t base) {
    size_t nd = sc.delta_lights.size(), na = sc.area_lights.size();
    size_t num_lights = nd + na + (has_env_light(sc) ? 1 : 0);
    if (num_lights == 0) return black();
    float light_pdf = 1.0f / (float)num_lights;
    size_t chosen = (size_t)(((uint64_t)ctx.draw_u32(base + 0) * (uint64_t)num_lights) >> 32);
    float l0 = ctx.draw(base + 1), l1 = ctx.draw(base + 2);
    float s0 = ctx.draw(base + 3), s1 = ctx.draw(base + 4);
    Color one;
    if (chosen < nd) {
        one = estimate_direct_delta(sc, hit, m, sc.delta_lights[chosen]);
    } else if (chosen >= nd && chosen < na) {  // Q1: bound is #area, not #delta + #area
        one = estimate_direct_area(sc, hit, m, s0, s1, sc.area_lights[chosen - nd], l0, l1);
    } else {
        Lobes lobes = bxdfs_at(sc, m, hit);
        if (lobes.n == 0) { panic_flag(P_EMPTY_BXDFS); return black() * (1.0f / light_pdf); }
        BSDF bsdf = bsdf_new(hit, &lobes);
        Color f; V3 wi; Prob pr;
        bsdf_sample(bsdf, hit.wo, s0, s1, &f, &wi, &pr);
        Ray incident = spawn_ray(hit, wi);
        Color ir = scene_occludes(sc, incident) ? black() : eval_env_light(sc, incident);
        one = ir * f * std::fabs(dot(wi, hit.normal)) * weak_recip(pr.v);
    }
    return one * (1.0f / light_pdf);
}

// ---------------- integrators ----------------
// src/pathintegrator.rs:9-74
static Color path_integrator(const Scene &sc, Ray ray, int depth, const SampleCtx &ctx) {
    Color radiance = black();
    bool specular_bounce = false;
    Color beta = gray(1.0f);
    for (int bounces = 0; bounces < depth; ++bounces) {
        uint32_t base = 2 + 8 * (uint32_t)bounces;
        Hit h;
        bool hit = scene_intersect(sc, ray, &h);
        if (bounces == 0 || specular_bounce) {
            Color env = eval_env_light(sc, ray);
            radiance = radiance + beta * (hit ? mtl_emission(sc.materials[sc.instances[h.inst].mtl_id]) : env);
        }
        if (!hit) break;
        const Material &m = sc.materials[sc.instances[h.inst].mtl_id];
        Lobes lobes = bxdfs_at(sc, m, h.isect);
        radiance = radiance + beta * uniform_sample_one_light(sc, h.isect, m, ctx, base);
        BSDF sp = bsdf_new(h.isect, &lobes);
        float r0 = ctx.draw(base + 5), r1 = ctx.draw(base + 6);
        Color f; V3 wi; Prob pr;
        bsdf_sample(sp, -ray.dir, r0, r1, &f, &wi, &pr);
        if (is_black(f) || is_zero(pr)) break;
        specular_bounce = pr.is_mass;
        beta = beta * f * dot(wi, h.isect.normal) * f_recip(pr.v);  // Q7: signed cosine
        ray = spawn_ray(h.isect, wi);
        if (bounces > 3) {
            float q = f_max(1.0f - luminance(beta), 0.05f);
            if (ctx.draw(base + 7) < q) break;
            beta = beta * f_recip(1.0f - q);
        }
    }
    return radiance;
}
// src/directlighting.rs:14-56
static Color direct_lighting_integrator(const Scene &sc, Ray ray, int depth, const SampleCtx &ctx) {
    if (depth <= 0) return black();
    Hit h;
    if (!scene_intersect(sc, ray, &h)) return eval_env_light(sc, ray);
    const Material &m = sc.materials[sc.instances[h.inst].mtl_id];
    if (!is_black(mtl_emission(m))) return mtl_emission(m);
    Color direct = uniform_sample_one_light(sc, h.isect, m, ctx, 2);
    Lobes lobes = bxdfs_at(sc, m, h.isect);
    BSDF bsdf = bsdf_new(h.isect, &lobes);
    Color spec_refl = black();
    Color f; V3 wi; Prob pr;
    if (bsdf_sample_specular(bsdf, h.isect.wo, &f, &wi, &pr)) {
        Ray refl = spawn_ray(h.isect, wi);
        Hit h2;
        Color sr;
        if (scene_intersect(sc, refl, &h2)) {
            const Material &m2 = sc.materials[sc.instances[h2.inst].mtl_id];
            sr = uniform_sample_one_light(sc, h2.isect, m2, ctx, 10);
        } else {
            sr = eval_env_light(sc, refl);
        }
        spec_refl = sr * f * weak_recip(mass(pr));
    }
    return direct + spec_refl;
}

static Ray primary_ray(const Scene &sc, uint32_t row, uint32_t col, uint32_t i, uint32_t msaa, const SampleCtx &ctx, bool no_jitter) {
    // src/main.rs:197-203
    float j0 = no_jitter ? 0.0f : ctx.draw(0), j1 = no_jitter ? 0.0f : ctx.draw(1);
    float dx = no_jitter ? 0.0f : ((float)(i / msaa) + j0) / (float)msaa;
    float dy = no_jitter ? 0.0f : ((float)(i % msaa) + j1) / (float)msaa;
    return shoot_ray(sc.camera, row, col, dx, dy);
}

static Color render_sample(const Scene &sc, const pbrs_render_opts &o, uint32_t row, uint32_t col, uint32_t i) {
    SampleCtx ctx{o.seed, row * sc.camera.width + col, i};
    Ray ray = primary_ray(sc, row, col, i, o.msaa, ctx, (o.flags & PBRS_FLAG_NO_JITTER) != 0);
    return o.integrator == PBRS_INTEGRATOR_PATH ? path_integrator(sc, ray, o.max_depth, ctx)
                                                 : direct_lighting_integrator(sc, ray, o.max_depth, ctx);
}

static bool owns_pixel(const pbrs_render_opts &o, uint32_t width, uint32_t row, uint32_t col) {
    if (o.world_size <= 1 || o.split != PBRS_SPLIT_TILES) return true;
    uint32_t tiles_x = (width + 63) / 64;
    uint32_t tile = (row / 64) * tiles_x + (col / 64);
    return (int)(tile % (uint32_t)o.world_size) == o.rank;
}
static bool owns_sample(const pbrs_render_opts &o, uint32_t i) {
    if (o.world_size <= 1 || o.split != PBRS_SPLIT_SAMPLES) return true;
    return (int)(i % (uint32_t)o.world_size) == o.rank;
}

}  // namespace orc

// =====================================================================================
//                                       C API
// =====================================================================================
using namespace orc;

static thread_local std::string g_err;
static int fail(int code, const char *msg) { g_err = msg; return code; }

struct oracle_scene {
    Scene sc;
};

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }
oracle_scene *oracle_scene_create(void) { return new oracle_scene(); }
void oracle_scene_destroy(oracle_scene *s) { delete s; }

int oracle_scene_set_camera(oracle_scene *s, uint32_t w, uint32_t h, float fov, const float eye[3], const float target[3], const float up[3]) {
    if (!s || !eye || !target || !up || w == 0 || h == 0) return fail(PBRS_ERR_INVALID_ARG, "camera: bad args");
    s->sc.camera = camera_new(w, h, fov);
    camera_look_at(s->sc.camera, V3{eye[0], eye[1], eye[2]}, V3{target[0], target[1], target[2]}, V3{up[0], up[1], up[2]});
    s->sc.has_camera = true;
    return 0;
}
int oracle_scene_add_texture_solid(oracle_scene *s, const float c[3]) {
    Texture t{};
    t.kind = TEX_SOLID; t.value = rgb(c[0], c[1], c[2]);
    s->sc.textures.push_back(t);
    return (int)s->sc.textures.size() - 1;
}
static Texture make_image(uint32_t w, uint32_t h, const uint8_t *data) {
    Texture t{};
    t.kind = TEX_IMAGE; t.width = w; t.height = h;
    t.data.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) t.data[i] = rgb8(data[3 * i], data[3 * i + 1], data[3 * i + 2]);
    return t;
}
int oracle_scene_add_texture_image_rgb8(oracle_scene *s, uint32_t w, uint32_t h, const uint8_t *data) {
    if (!data || w == 0 || h == 0) return fail(PBRS_ERR_INVALID_ARG, "image: bad args");
    s->sc.textures.push_back(make_image(w, h, data));
    return (int)s->sc.textures.size() - 1;
}
int oracle_scene_add_texture_perlin(oracle_scene *s, float freq, const float *rv, const uint32_t *px, const uint32_t *py, const uint32_t *pz) {
    Texture t{};
    t.kind = TEX_PERLIN; t.freq = freq;
    t.rand_vec.resize(256); t.perm_x.assign(px, px + 256); t.perm_y.assign(py, py + 256); t.perm_z.assign(pz, pz + 256);
    for (int i = 0; i < 256; ++i) {
        t.rand_vec[i] = V3{rv[3 * i], rv[3 * i + 1], rv[3 * i + 2]};
        if (px[i] > 255 || py[i] > 255 || pz[i] > 255) return fail(PBRS_ERR_INVALID_ARG, "perlin: perm out of range");
    }
    s->sc.textures.push_back(t);
    return (int)s->sc.textures.size() - 1;
}
int oracle_scene_add_material(oracle_scene *s, const pbrs_material_desc *d) {
    if (!d || d->kind < 0 || d->kind > PBRS_MTL_SUBSTRATE) return fail(PBRS_ERR_INVALID_ARG, "material: bad kind");
    Material m;
    m.kind = d->kind; m.tex_kd = d->tex_kd; m.tex_ks = d->tex_ks; m.tex_kr = d->tex_kr; m.tex_kt = d->tex_kt;
    m.a = rgb(d->color_a[0], d->color_a[1], d->color_a[2]);
    m.b = rgb(d->color_b[0], d->color_b[1], d->color_b[2]);
    for (int i = 0; i < 4; ++i) m.f[i] = d->f[i];
    m.remap = d->remap_roughness != 0;
    s->sc.materials.push_back(m);
    return (int)s->sc.materials.size() - 1;
}
int oracle_scene_add_sphere(oracle_scene *s, const float c[3], float r) {
    s->sc.spheres.push_back(Sphere{V3{c[0], c[1], c[2]}, r});
    s->sc.shapes.push_back(ShapeRef{SHAPE_SPHERE, (int)s->sc.spheres.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_mesh(oracle_scene *s, const float *P, const float *N, const float *UV, uint32_t nverts, const uint32_t *idx, uint32_t ntris) {
    if (!P || !idx || ntris == 0) return fail(PBRS_ERR_INVALID_ARG, "mesh: bad args");
    auto m = std::make_unique<Mesh>();
    mesh_build(*m, P, N, UV, nverts, idx, ntris);
    s->sc.meshes.push_back(std::move(m));
    s->sc.shapes.push_back(ShapeRef{SHAPE_MESH, (int)s->sc.meshes.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_quad(oracle_scene *s, const float o[3], const float su[3], const float sv[3]) {
    s->sc.quads.push_back(Quad{V3{o[0], o[1], o[2]}, V3{su[0], su[1], su[2]}, V3{sv[0], sv[1], sv[2]}});
    s->sc.shapes.push_back(ShapeRef{SHAPE_QUAD, (int)s->sc.quads.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_cuboid(oracle_scene *s, const float p0[3], const float p1[3]) {
    s->sc.cuboids.push_back(cuboid_from_points(V3{p0[0], p0[1], p0[2]}, V3{p1[0], p1[1], p1[2]}));
    s->sc.shapes.push_back(ShapeRef{SHAPE_CUBOID, (int)s->sc.cuboids.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
// Disk::new, simple.rs:42-52: the normal is normalised; the two asserts are argument errors here
static int make_disk(const float c[3], const float n[3], const float rad[3], Disk *out) {
    V3 nv{n[0], n[1], n[2]};
    float n2 = norm_squared(nv);
    if (!(n2 != 0.0f && std::isfinite(n2))) return fail(PBRS_ERR_INVALID_ARG, "disk: normal cannot be normalised");
    V3 normal = hat(nv), radial{rad[0], rad[1], rad[2]};
    if (!std::isfinite(norm_squared(radial))) return fail(PBRS_ERR_INVALID_ARG, "disk: radial is not finite");
    if (!(std::fabs(dot(radial, normal)) < 1e-6f)) return fail(PBRS_ERR_INVALID_ARG, "disk: radial is not perpendicular to the normal");
    *out = Disk{V3{c[0], c[1], c[2]}, normal, radial};
    return 0;
}
int oracle_scene_add_disk(oracle_scene *s, const float c[3], const float n[3], const float rad[3]) {
    Disk d;
    if (int rc = make_disk(c, n, rad, &d)) return rc;
    s->sc.disks.push_back(d);
    s->sc.shapes.push_back(ShapeRef{SHAPE_DISK, (int)s->sc.disks.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_triangle(oracle_scene *s, const float p0[3], const float p1[3], const float p2[3]) {
    s->sc.triangles.push_back(IsoTriangle{V3{p0[0], p0[1], p0[2]}, V3{p1[0], p1[1], p1[2]}, V3{p2[0], p2[1], p2[2]}});
    s->sc.shapes.push_back(ShapeRef{SHAPE_TRIANGLE, (int)s->sc.triangles.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_sphere_blas(oracle_scene *s, const float *centers_radii, uint32_t n) {
    if (!centers_radii || n == 0) return fail(PBRS_ERR_INVALID_ARG, "sphere_blas: bad args");
    for (uint32_t i = 0; i < 4 * n; ++i)
        if (std::isnan(centers_radii[i])) return fail(PBRS_ERR_INVALID_ARG, "sphere_blas: NaN (Sphere::from_raw asserts)");
    auto m = std::make_unique<Mesh>();
    sphere_blas_build(*m, centers_radii, n);
    s->sc.meshes.push_back(std::move(m));
    s->sc.shapes.push_back(ShapeRef{SHAPE_MESH, (int)s->sc.meshes.size() - 1});
    return (int)s->sc.shapes.size() - 1;
}
int oracle_scene_add_instance(oracle_scene *s, int shape, int mtl, const float *fwd, const float *inv) {
    if (shape < 0 || shape >= (int)s->sc.shapes.size() || mtl < 0 || mtl >= (int)s->sc.materials.size())
        return fail(PBRS_ERR_INVALID_ARG, "instance: bad ids");
    Instance in;
    in.shape_id = shape; in.mtl_id = mtl; in.id = (int)s->sc.instances.size();
    in.xf.fwd = m4_identity(); in.xf.inv = m4_identity();
    in.xf.is_identity = !(fwd && inv);
    if (fwd && inv)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) { in.xf.fwd.c[c][r] = fwd[4 * c + r]; in.xf.inv.c[c][r] = inv[4 * c + r]; }
    s->sc.instances.push_back(in);
    return in.id;
}
int oracle_scene_add_point_light(oracle_scene *s, const float p[3], const float i[3]) {
    DeltaLight l{};
    l.kind = DELTA_POINT; l.position = V3{p[0], p[1], p[2]}; l.intensity = rgb(i[0], i[1], i[2]);
    s->sc.delta_lights.push_back(l);
    return 0;
}
int oracle_scene_add_distant_light(oracle_scene *s, const float d[3], const float rad[3], float world_radius) {
    DeltaLight l{};
    l.kind = DELTA_DISTANT; l.casting_dir = V3{d[0], d[1], d[2]}; l.intensity = rgb(rad[0], rad[1], rad[2]);
    l.world_radius = world_radius;
    s->sc.delta_lights.push_back(l);
    return 0;
}
int oracle_scene_add_area_light_sphere(oracle_scene *s, const float c[3], float r, const float e[3]) {
    AreaLight l{};
    l.shape_kind = AREA_SPHERE; l.sphere = Sphere{V3{c[0], c[1], c[2]}, r}; l.emit = rgb(e[0], e[1], e[2]);
    l.area = sphere_area(l.sphere);
    s->sc.area_lights.push_back(l);
    return 0;
}
int oracle_scene_add_area_light_triangle(oracle_scene *s, const float p0[3], const float p1[3], const float p2[3], const float e[3]) {
    AreaLight l{};
    l.shape_kind = AREA_TRIANGLE;
    l.tri = IsoTriangle{V3{p0[0], p0[1], p0[2]}, V3{p1[0], p1[1], p1[2]}, V3{p2[0], p2[1], p2[2]}};
    l.emit = rgb(e[0], e[1], e[2]);
    l.area = isotri_area(l.tri);
    s->sc.area_lights.push_back(l);
    return 0;
}
int oracle_scene_add_area_light_quad(oracle_scene *s, const float o[3], const float su[3], const float sv[3], const float e[3]) {
    AreaLight l{};
    l.shape_kind = AREA_QUAD;
    l.quad = Quad{V3{o[0], o[1], o[2]}, V3{su[0], su[1], su[2]}, V3{sv[0], sv[1], sv[2]}};
    l.emit = rgb(e[0], e[1], e[2]);
    l.area = quad_area(l.quad);
    s->sc.area_lights.push_back(l);
    return 0;
}
int oracle_scene_add_area_light_disk(oracle_scene *s, const float c[3], const float n[3], const float rad[3], const float e[3]) {
    AreaLight l{};
    l.shape_kind = AREA_DISK;
    if (int rc = make_disk(c, n, rad, &l.disk)) return rc;
    l.emit = rgb(e[0], e[1], e[2]);
    l.area = disk_area(l.disk);
    s->sc.area_lights.push_back(l);
    return 0;
}
int oracle_scene_set_env_constant(oracle_scene *s, const float c[3]) {
    s->sc.env_kind = ENV_CONSTANT; s->sc.env_color = rgb(c[0], c[1], c[2]);
    return 0;
}
int oracle_scene_set_env_fn(oracle_scene *s, int kind) {
    if (kind < 0 || kind > 2) return fail(PBRS_ERR_INVALID_ARG, "env fn kind");
    s->sc.env_kind = ENV_FN; s->sc.env_fn = kind;
    return 0;
}
int oracle_scene_set_env_image(oracle_scene *s, uint32_t w, uint32_t h, const uint8_t *data, const float scale[3]) {
    s->sc.env_kind = ENV_IMAGE; s->sc.env_image = make_image(w, h, data);
    s->sc.env_scale = rgb(scale[0], scale[1], scale[2]);
    return 0;
}
int oracle_scene_commit(oracle_scene *s) {
    Scene &sc = s->sc;
    if (!sc.has_camera) return fail(PBRS_ERR_STATE, "commit: no camera");
    if (sc.instances.empty()) return fail(PBRS_ERR_STATE, "commit: empty instances");  // tlas/src/bvh.rs:117
    std::vector<int> all(sc.instances.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
    sc.n_tlas_inner = 0;
    sc.tlas = tlas_build(sc, all);
    // scene/src/lib.rs:54-58
    for (auto &l : sc.delta_lights)
        if (l.kind == DELTA_DISTANT && !(l.world_radius > 0.0f && std::isfinite(l.world_radius)))
            l.world_radius = norm(bbox_diag(sc.tlas->bbox)) * 0.5f;
    sc.committed = true;
    return 0;
}

}  // extern "C" (helpers below are C++)

static void fill_stats(pbrs_stats *st, const Diag &d, uint64_t n_samples, double ms) {
    if (!st) return;
    std::memset(st, 0, sizeof *st);
    st->n_samples = n_samples;
    st->n_rays_extend = d.n_rays_extend; st->n_rays_shadow = d.n_rays_shadow;
    st->n_nodes = d.n_nodes; st->n_tris = d.n_tris; st->n_spheres = d.n_spheres; st->n_instances = d.n_instances;
    for (int i = 0; i < 16; ++i) st->would_panic[i] = d.would_panic[i];
    for (int i = 0; i < 4; ++i) { st->trav_extend[i] = d.trav_extend[i]; st->trav_shadow[i] = d.trav_shadow[i]; }
    st->ms_total = ms;
}

struct Crop { uint32_t x, y, w, h; };
static Crop get_crop(const Scene &sc, const pbrs_render_opts &o) {
    if (o.crop_w == 0 || o.crop_h == 0) return Crop{0, 0, sc.camera.width, sc.camera.height};
    return Crop{o.crop_x, o.crop_y, o.crop_w, o.crop_h};
}

// Row-parallel driver (= the rayon par_iter over rows, src/main.rs:219-224).  `fn(row, diag)`.
template <class F>
static void parallel_rows(uint32_t y0, uint32_t y1, uint32_t row_step, int threads, Diag &total, F fn) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::atomic<uint32_t> next(y0);
    std::vector<Diag> diags(threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            g_diag = &diags[t];
            while (true) {
                uint32_t row = next.fetch_add(row_step);
                if (row >= y1) break;
                fn(row);
            }
            g_diag = nullptr;
        });
    }
    for (auto &th : pool) th.join();
    for (auto &d : diags) total.add(d);
}

static int g_threads = 0;

extern "C" {

void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
    int t = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    return t > 0 ? t : 1;
}

// Film of this rank's share; out is W*H*3 (full frame; pixels outside the crop/rank stay 0).
// row_step > 1 renders only rows y0, y0+row_step, ... (bench.py's bounded CPU sample).
int oracle_render_rows(const oracle_scene *s, const pbrs_render_opts *o, float *out, pbrs_stats *st, uint32_t row_step) {
    if (!s || !o || !out) return fail(PBRS_ERR_INVALID_ARG, "render: null");
    const Scene &sc = s->sc;
    if (!sc.committed) return fail(PBRS_ERR_STATE, "render before commit");
    if (o->msaa == 0) return fail(PBRS_ERR_INVALID_ARG, "msaa == 0");
    if (row_step == 0) row_step = 1;
    uint32_t W = sc.camera.width, H = sc.camera.height, spp = o->msaa * o->msaa;
    std::memset(out, 0, sizeof(float) * 3 * (size_t)W * H);
    Crop c = get_crop(sc, *o);
    Diag total;
    std::atomic<uint64_t> n_samples(0);
    auto t0 = std::chrono::steady_clock::now();
    parallel_rows(c.y, c.y + c.h, row_step, g_threads, total, [&](uint32_t row) {
        uint64_t local = 0;
        for (uint32_t col = c.x; col < c.x + c.w; ++col) {
            if (!owns_pixel(*o, W, row, col)) continue;
            Color sum = black();
            for (uint32_t i = 0; i < spp; ++i) {
                if (!owns_sample(*o, i)) continue;
                sum = sum + render_sample(sc, *o, row, col, i);  // src/main.rs:205
                local++;
            }
            Color px = (o->flags & PBRS_FLAG_RAW_SUM) ? sum : sum * (1.0f / (float)spp);  // :208, color.rs:90-95
            float *p = out + 3 * ((size_t)row * W + col);
            p[0] = px.r; p[1] = px.g; p[2] = px.b;
        }
        n_samples += local;
    });
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fill_stats(st, total, n_samples.load(), ms);
    return 0;
}
int oracle_render(const oracle_scene *s, const pbrs_render_opts *o, float *out, pbrs_stats *st) {
    return oracle_render_rows(s, o, out, st, 1);
}

int oracle_render_ids(const oracle_scene *s, const pbrs_render_opts *o, uint32_t sample_index, uint32_t *out_inst, uint32_t *out_prim, float *out_t) {
    if (!s || !o) return fail(PBRS_ERR_INVALID_ARG, "render_ids: null");
    const Scene &sc = s->sc;
    if (!sc.committed) return fail(PBRS_ERR_STATE, "render before commit");
    Crop c = get_crop(sc, *o);
    Diag total;
    parallel_rows(c.y, c.y + c.h, 1, g_threads, total, [&](uint32_t row) {
        for (uint32_t col = c.x; col < c.x + c.w; ++col) {
            SampleCtx ctx{o->seed, row * sc.camera.width + col, sample_index};
            Ray ray = primary_ray(sc, row, col, sample_index, o->msaa ? o->msaa : 1, ctx, (o->flags & PBRS_FLAG_NO_JITTER) != 0);
            Hit h;
            bool hit = scene_intersect(sc, ray, &h);
            size_t k = (size_t)(row - c.y) * c.w + (col - c.x);
            if (out_inst) out_inst[k] = hit ? (uint32_t)h.inst : 0xFFFFFFFFu;
            if (out_prim) out_prim[k] = hit ? h.prim : 0xFFFFFFFFu;
            if (out_t) out_t[k] = hit ? h.isect.ray_t : kInf;
        }
    });
    return 0;
}

int oracle_render_samples(const oracle_scene *s, const pbrs_render_opts *o, float *out, pbrs_stats *st) {
    if (!s || !o || !out) return fail(PBRS_ERR_INVALID_ARG, "render_samples: null");
    const Scene &sc = s->sc;
    if (!sc.committed) return fail(PBRS_ERR_STATE, "render before commit");
    Crop c = get_crop(sc, *o);
    uint32_t spp = o->msaa * o->msaa;
    Diag total;
    std::atomic<uint64_t> n_samples(0);
    auto t0 = std::chrono::steady_clock::now();
    parallel_rows(c.y, c.y + c.h, 1, g_threads, total, [&](uint32_t row) {
        for (uint32_t col = c.x; col < c.x + c.w; ++col)
            for (uint32_t i = 0; i < spp; ++i) {
                Color r = render_sample(sc, *o, row, col, i);
                float *p = out + 3 * (((size_t)(row - c.y) * c.w + (col - c.x)) * spp + i);
                p[0] = r.r; p[1] = r.g; p[2] = r.b;
            }
        n_samples += (uint64_t)c.w * spp;
    });
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fill_stats(st, total, n_samples.load(), ms);
    return 0;
}

int oracle_scene_get_info(const oracle_scene *s, pbrs_scene_info *info) {
    const Scene &sc = s->sc;
    std::memset(info, 0, sizeof *info);
    info->width = sc.camera.width; info->height = sc.camera.height;
    info->n_instances = (uint32_t)sc.instances.size();
    info->n_meshes = (uint32_t)sc.meshes.size();
    uint32_t nt = 0, nb = 0;
    for (auto &m : sc.meshes) {
        if (m->balls.empty()) nt += (uint32_t)m->tris.size(); else nb += (uint32_t)m->balls.size();
    }
    info->n_spheres = (uint32_t)sc.spheres.size() + nb;  // incl. the spheres of sphere BLASes
    info->n_triangles = nt;
    info->n_tlas_nodes = sc.n_tlas_inner;
    info->n_lights = (uint32_t)(sc.delta_lights.size() + sc.area_lights.size() + (has_env_light(sc) ? 1 : 0));
    if (sc.tlas)
        for (int i = 0; i < 3; ++i) { info->world_min[i] = sc.tlas->bbox.mn[i]; info->world_max[i] = sc.tlas->bbox.mx[i]; }
    return 0;
}

uint32_t oracle_sampler_u32(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    return sampler_u32(seed, pixel, sample, dim);
}

// Trace one world-space ray through the committed scene (closest hit), for KATs.
// out: [0]=hit(0/1) [1]=t [2..4]=pos [5..7]=normal [8..9]=uv [10]=inst [11]=prim [12..14]=tangent
int oracle_trace_ray(const oracle_scene *s, const float o[3], const float d[3], float t_max, float *out) {
    const Scene &sc = s->sc;
    if (!sc.committed) return fail(PBRS_ERR_STATE, "trace before commit");
    Ray r{V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, t_max};
    Hit h;
    Diag dg;
    g_diag = &dg;
    bool hit = scene_intersect(sc, r, &h);
    g_diag = nullptr;
    for (int i = 0; i < 16; ++i) out[i] = 0.0f;
    out[0] = hit ? 1.0f : 0.0f;
    if (hit) {
        out[1] = h.isect.ray_t;
        out[2] = h.isect.pos.x; out[3] = h.isect.pos.y; out[4] = h.isect.pos.z;
        out[5] = h.isect.normal.x; out[6] = h.isect.normal.y; out[7] = h.isect.normal.z;
        out[8] = h.isect.u; out[9] = h.isect.v;
        out[10] = (float)h.inst; out[11] = (float)h.prim;
        out[12] = h.isect.tbn.c[0].x; out[13] = h.isect.tbn.c[0].y; out[14] = h.isect.tbn.c[0].z;
    }
    uint64_t panics = 0;
    for (int i = 0; i < 16; ++i) panics += dg.would_panic[i];
    out[15] = (float)panics;
    return 0;
}
int oracle_occludes_ray(const oracle_scene *s, const float o[3], const float d[3], float t_max) {
    const Scene &sc = s->sc;
    if (!sc.committed) return fail(PBRS_ERR_STATE, "trace before commit");
    Ray r{V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, t_max};
    return scene_occludes(sc, r) ? 1 : 0;
}

// ---- known-answer hooks for the function-level KATs of SURVEY.md section 4 ----
enum {
    KAT_FRESNEL_DIELECTRIC = 1,   // in: eta_front, eta_back, cos            out: refl_coeff
    KAT_OMEGA_TRIG = 2,           // in: x,y,z   out: cos,cos2,sin2,sin,cos_phi,sin_phi,cos2_phi,sin2_phi
    KAT_SPECULAR_DIELECTRIC = 3,  // in: albedo(3), eta_o, eta_i, wo(3), u, v  out: f(3), wi(3), is_mass, p
    KAT_REFLECT = 4,              // in: n(3), wi(3)   out: v(3)
    KAT_REFRACT = 5,              // in: n(3), wi(3), ratio  out: transmit?, v(3)
    KAT_SPHERE_INTERSECT = 6,     // in: c(3), r, o(3), d(3), tmax   out: hit, t, pos(3), occludes
    KAT_MAKE_COORD = 7,           // in: v(3)  out: v1(3), v2(3)
    KAT_LOBE = 8,                 // in: kind, params...; see tests/oracle_ffi.py
    KAT_BECKMANN = 9,             // in: ax, ay, op, w(3), w2(3)/uv   out: value(s)
    KAT_SPHERE_LIGHT = 10,        // in: c(3), r, target pos(3), target normal(3), op, u,v / wi(3)
    KAT_ROUGHNESS_TO_ALPHA = 11,
    KAT_CONCENTRIC = 12,          // in: u, v  out: x, y
    KAT_FRESNEL_CONDUCTOR = 13,   // in: eta(3), k(3), cos  out: rgb
    KAT_TRIANGLE = 14,            // in: p0,p1,p2 (9), o(3), d(3), tmax  out: hit,t,pos(3),uv(2),pred
    KAT_BBOX = 15,                // in: mn(3), mx(3), o(3), d(3), tmax  out: hit, t_low
    KAT_POWI = 16,                // in: x, n  out: x^n
    KAT_LUMINANCE = 17,
    KAT_CATHETUS = 18,            // in: h, o  out: cathetus
};

static Lobe kat_make_lobe(const float *in) {
    int kind = (int)in[0];
    Color albedo = rgb(in[1], in[2], in[3]);
    switch (kind) {
    case 0: return lobe_lambert(albedo);
    case 1: return lobe_oren_nayar(albedo, in[4]);
    case 2: return lobe_mirror(albedo);
    case 3: return lobe_dielectric(albedo, in[4], in[5]);
    case 4: return lobe_transmit(albedo, in[4], in[5]);
    case 5: return lobe_microfacet(albedo, Beckmann{in[4], in[5]}, fresnel_nop());
    case 6: return lobe_microfacet(albedo, Beckmann{in[4], in[5]}, fresnel_dielectric(in[6], in[7]));
    default: return lobe_microfacet(albedo, Beckmann{in[4], in[5]}, fresnel_conductor(rgb(in[6], in[7], in[8]), rgb(in[9], in[10], in[11])));
    }
}

int oracle_kat(int op, const float *in, int n_in, float *out, int n_out) {
    (void)n_in; (void)n_out;
    Diag dg;
    g_diag = &dg;
    int rc = 0;
    switch (op) {
    case KAT_FRESNEL_DIELECTRIC:
        out[0] = fresnel_refl_coeff(fresnel_dielectric(in[0], in[1]), in[2]);
        break;
    case KAT_OMEGA_TRIG: {
        V3 w{in[0], in[1], in[2]};
        out[0] = cos_theta(w); out[1] = cos2_theta(w); out[2] = sin2_theta(w); out[3] = sin_theta(w);
        out[4] = cos_phi(w); out[5] = sin_phi(w); out[6] = cos2_phi(w); out[7] = sin2_phi(w);
        break;
    }
    case KAT_SPECULAR_DIELECTRIC: {
        Lobe l = lobe_dielectric(rgb(in[0], in[1], in[2]), in[3], in[4]);
        Color f; V3 wi; Prob p;
        lobe_sample(l, V3{in[5], in[6], in[7]}, in[8], in[9], &f, &wi, &p);
        out[0] = f.r; out[1] = f.g; out[2] = f.b; out[3] = wi.x; out[4] = wi.y; out[5] = wi.z;
        out[6] = p.is_mass ? 1.0f : 0.0f; out[7] = p.v;
        break;
    }
    case KAT_REFLECT: {
        V3 v = reflect(V3{in[0], in[1], in[2]}, V3{in[3], in[4], in[5]});
        out[0] = v.x; out[1] = v.y; out[2] = v.z;
        break;
    }
    case KAT_REFRACT: {
        V3 v;
        bool tr = refract(V3{in[0], in[1], in[2]}, V3{in[3], in[4], in[5]}, in[6], &v);
        out[0] = tr ? 1.0f : 0.0f; out[1] = v.x; out[2] = v.y; out[3] = v.z;
        break;
    }
    case KAT_SPHERE_INTERSECT: {
        Sphere s{V3{in[0], in[1], in[2]}, in[3]};
        Ray r{V3{in[4], in[5], in[6]}, V3{in[7], in[8], in[9]}, in[10]};
        Interaction h;
        bool hit = sphere_intersect(s, r, &h);
        out[0] = hit ? 1.0f : 0.0f;
        if (hit) { out[1] = h.ray_t; out[2] = h.pos.x; out[3] = h.pos.y; out[4] = h.pos.z; }
        out[5] = sphere_occludes(s, r) ? 1.0f : 0.0f;
        break;
    }
    case KAT_MAKE_COORD: {
        V3 a, b;
        make_coord_system(V3{in[0], in[1], in[2]}, &a, &b);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = b.z;
        break;
    }
    case KAT_LOBE: {
        // in[0..11] lobe, in[12] op (0 eval, 1 prob, 2 sample), in[13..15] wo, in[16..18] wi or (u,v)
        Lobe l = kat_make_lobe(in);
        int sub = (int)in[12];
        V3 wo{in[13], in[14], in[15]};
        if (sub == 0) { Color c = lobe_eval(l, wo, V3{in[16], in[17], in[18]}); out[0] = c.r; out[1] = c.g; out[2] = c.b; }
        else if (sub == 1) { Prob p = lobe_prob(l, wo, V3{in[16], in[17], in[18]}); out[0] = p.is_mass ? 1.0f : 0.0f; out[1] = p.v; }
        else {
            Color f; V3 wi; Prob p;
            lobe_sample(l, wo, in[16], in[17], &f, &wi, &p);
            out[0] = f.r; out[1] = f.g; out[2] = f.b; out[3] = wi.x; out[4] = wi.y; out[5] = wi.z;
            out[6] = p.is_mass ? 1.0f : 0.0f; out[7] = p.v;
        }
        break;
    }
    case KAT_BECKMANN: {
        Beckmann m{in[0], in[1]};
        int sub = (int)in[2];
        V3 w{in[3], in[4], in[5]};
        if (sub == 0) out[0] = mf_d(m, w);
        else if (sub == 1) out[0] = mf_g1(m, w);
        else if (sub == 2) out[0] = mf_pdf(m, w, V3{in[6], in[7], in[8]});
        else if (sub == 3) { V3 h = mf_sample_wh(m, w, in[6], in[7]); out[0] = h.x; out[1] = h.y; out[2] = h.z; }
        else out[0] = mf_g(m, w, V3{in[6], in[7], in[8]});
        break;
    }
    case KAT_SPHERE_LIGHT: {
        Sphere s{V3{in[0], in[1], in[2]}, in[3]};
        Interaction t = isect_rayless(V3{in[4], in[5], in[6]}, 0, 0, V3{in[7], in[8], in[9]});
        int sub = (int)in[10];
        if (sub == 0) {
            Interaction p = sphere_sample_towards(s, t, in[11], in[12]);
            out[0] = p.pos.x; out[1] = p.pos.y; out[2] = p.pos.z; out[3] = p.normal.x; out[4] = p.normal.y; out[5] = p.normal.z;
        } else {
            float pdf = 0.0f;
            bool ok = sphere_pdf_at(s, t, V3{in[11], in[12], in[13]}, &pdf);
            out[0] = ok ? 1.0f : 0.0f; out[1] = pdf;
        }
        break;
    }
    case KAT_ROUGHNESS_TO_ALPHA: out[0] = roughness_to_alpha(in[0]); break;
    case KAT_CONCENTRIC: concentric_sample_disk(in[0], in[1], &out[0], &out[1]); break;
    case KAT_FRESNEL_CONDUCTOR: {
        Color c = fresnel_eval(fresnel_conductor(rgb(in[0], in[1], in[2]), rgb(in[3], in[4], in[5])), in[6]);
        out[0] = c.r; out[1] = c.g; out[2] = c.b;
        break;
    }
    case KAT_TRIANGLE: {
        V3 p0{in[0], in[1], in[2]}, p1{in[3], in[4], in[5]}, p2{in[6], in[7], in[8]};
        Ray r{V3{in[9], in[10], in[11]}, V3{in[12], in[13], in[14]}, in[15]};
        Interaction h;
        bool hit = intersect_triangle(p0, p1, p2, r, &h);
        out[0] = hit ? 1.0f : 0.0f;
        if (hit) { out[1] = h.ray_t; out[2] = h.pos.x; out[3] = h.pos.y; out[4] = h.pos.z; out[5] = h.u; out[6] = h.v; }
        out[7] = intersect_triangle_pred(p0, p1, p2, r) ? 1.0f : 0.0f;
        break;
    }
    case KAT_BBOX: {
        BBox b{{in[0], in[1], in[2]}, {in[3], in[4], in[5]}};
        Ray r{V3{in[6], in[7], in[8]}, V3{in[9], in[10], in[11]}, in[12]};
        float tl = 0.0f;
        out[0] = bbox_intersect(b, r, &tl) ? 1.0f : 0.0f;
        out[1] = tl;
        break;
    }
    case KAT_POWI: out[0] = f_powi(in[0], (int)in[1]); break;
    case KAT_LUMINANCE: out[0] = luminance(rgb(in[0], in[1], in[2])); break;
    case KAT_CATHETUS: out[0] = cathetus(in[0], in[1]); break;
    default: rc = fail(PBRS_ERR_INVALID_ARG, "unknown KAT op");
    }
    g_diag = nullptr;
    uint64_t panics = 0;
    for (int i = 0; i < 16; ++i) panics += dg.would_panic[i];
    return rc < 0 ? rc : (int)panics;
}

}  // extern "C"
