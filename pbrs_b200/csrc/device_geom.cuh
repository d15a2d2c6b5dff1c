// Ray / box / triangle / sphere arithmetic over the flattened records of records.h (the BVH walks
// that use it are in device_walk.cuh).
//
// Parity contract (DESIGN.md "Traversal"): the walks visit the same boxes, in the same order and
// with the same ray extent as the reference's recursive/stack walks, so that the integer outcome
// (which primitive wins, including ties and the extent quirks Q14/Q17) is bit-identical:
//   TLAS closest  tlas/src/bvh.rs:77-103     (left then right, `ray.t_max` mutated in place)
//   TLAS any      tlas/src/bvh.rs:105-113
//   BLAS closest  shape/src/blas.rs:422-476  (explicit stack, near child first by split axis)
//   BLAS any      shape/src/blas.rs:478-495
// What differs is the data layout: one 64-byte record holds BOTH children's boxes, so a node is
// fetched once per expansion with two 32-byte loads, and the far child waits on the stack with
// its entry distance instead of being re-fetched.
#pragma once
#include "../../include/pbrs_gpu.h"
#include "device_math.cuh"
#include "records.h"

namespace pbrs {

struct f4 {
    float x, y, z, w;
};
struct u4 {
    uint32_t x, y, z, w;
};

// 16-byte read-only load (LDG.E.128.CONSTANT on the device)
PB_DEV f4 ld16(const void *p) {
#ifdef __CUDA_ARCH__
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    f4 r; memcpy(&r, p, 16); return r;
#endif
}
// 32-byte read-only load of a 32-byte aligned half record (LDG.E.256.CONSTANT, new with sm_100): a
// 64-byte node / triangle / instance record is two requests to the L1 instead of four -- the
// traversal kernels are bound by the L1's load pipe (ncu: l1tex data-pipe wavefronts 74 % of peak)
#ifndef PBRS_LD256
#define PBRS_LD256 1
#endif

PB_DEV void ld32(const void *p, f4 &a, f4 &b) {
#if defined(__CUDA_ARCH__) && PBRS_LD256
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
#else
    a = ld16(p);
    b = ld16(reinterpret_cast<const char *>(p) + 16);
#endif
}
PB_DEV uint32_t ld_u32(const uint32_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
PB_DEV float ld_f32(const float *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Traversal work counters (SURVEY.md 8d definitions); only live in the COUNT instantiations.
struct TravCount {
    uint32_t nodes, tris, spheres, insts;
};

// ---------------------------------------------------------------------------------------------
// BBox::intersect, geometry/src/bvh.rs:84-99.  True divisions (the reference keeps no reciprocal),
// glam Vec3A lane semantics for min/max, f32::max/min for the clamps.  Split in two so the far
// child can be re-tested against a shrunken extent without recomputing the slabs:
//   t_low  = max(max_element(min(t0,t1)), 0)      -- independent of the extent
//   min_el = min_element(max(t0,t1))
//   hit    = t_low <= min(min_el, t_max)
// ---------------------------------------------------------------------------------------------
PB_DEV void slab(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const Ray &r, float &t_low,
                 float &min_el) {
    float t0x = (mnx - r.o.x) / r.d.x, t1x = (mxx - r.o.x) / r.d.x;
    float t0y = (mny - r.o.y) / r.d.y, t1y = (mxy - r.o.y) / r.d.y;
    float t0z = (mnz - r.o.z) / r.d.z, t1z = (mxz - r.o.z) / r.d.z;
    float lox = lane_min(t0x, t1x), hix = lane_max(t0x, t1x);
    float loy = lane_min(t0y, t1y), hiy = lane_max(t0y, t1y);
    float loz = lane_min(t0z, t1z), hiz = lane_max(t0z, t1z);
    float max_el = lane_max(lane_max(lox, loz), lane_max(loy, loz));
    min_el = lane_min(lane_min(hix, hiz), lane_min(hiy, hiz));
    t_low = fmaxf(max_el, 0.0f);
}
PB_DEV bool box_pass(float t_low, float min_el, float t_max) { return t_low <= fminf(min_el, t_max); }

// ---------------------------------------------------------------------------------------------
// Triangles.  shape/src/simple.rs:435-475 (closest) and :477-495 (predicate).
// ---------------------------------------------------------------------------------------------
struct TriHit {
    float t, b0, b1, b2;  // normalised barycentrics (uv = (b1, b2))
    vec3 normal;          // unit geometric normal facing the ray
    vec3 pos;
};
// intersect_triangle from the unit normal n = hat((p0 - p1) x (p2 - p1)) on (simple.rs:441-475)
PB_DEV bool tri_intersect_n(vec3 p0, vec3 p1, vec3 p2, vec3 n, const Ray &r, TriHit &h, Diag &dg) {
    vec3 normal = facing(n, r.d);
    if (!(dot(normal, r.d) <= 0.0f)) flag(dg, P_MISC);
    float t = dot(normal, p0 - r.o) / dot(normal, r.d);
    if (!in_extent(t, r.t_max)) return false;
    vec3 p = at(r, t);
    float b2 = dot(cross(p - p0, p - p1), normal);
    float b0 = dot(cross(p - p1, p - p2), normal);
    float b1 = dot(cross(p - p2, p - p0), normal);
    if (is_nan(b0) || is_nan(b1) || is_nan(b2)) return false;
    bool g0 = b0 > 0.0f, g1 = b1 > 0.0f, g2 = b2 > 0.0f;
    if (!((g0 && g1 && g2) || (!g0 && !g1 && !g2))) return false;
    float total = b0 + b1 + b2;
    b0 = b0 / total; b1 = b1 / total; b2 = b2 / total;
    vec3 hp = bary_lerp(p0, p1, p2, b0, b1);
    if (any_nan(hp)) return false;
    h.t = t; h.b0 = b0; h.b1 = b1; h.b2 = b2; h.normal = normal; h.pos = hp;
    return true;
}
PB_DEV bool tri_intersect(vec3 p0, vec3 p1, vec3 p2, const Ray &r, TriHit &h, Diag &dg) {
    vec3 n;
    if (!try_hat(cross(p0 - p1, p2 - p1), n)) return false;
    return tri_intersect_n(p0, p1, p2, n, r, h, dg);
}
// intersect_triangle_pred from the unit normal on (simple.rs:481-495)
PB_DEV bool tri_occludes_n(vec3 p0, vec3 p1, vec3 p2, vec3 normal, const Ray &r, Diag &dg) {
    float t = dot(normal, p0 - r.o) / dot(normal, r.d);
    if (!in_extent(t, r.t_max)) return false;
    vec3 p = at(r, t);
    float b0 = dot(cross(p - p0, p - p1), normal);
    float b1 = dot(cross(p - p1, p - p2), normal);
    float b2 = dot(cross(p - p2, p - p0), normal);
    if (is_nan(b0) || is_nan(b1) || is_nan(b2)) flag(dg, P_MISC);
    bool g0 = b0 > 0.0f, g1 = b1 > 0.0f, g2 = b2 > 0.0f;
    return (g0 && g1 && g2) || (!g0 && !g1 && !g2);
}
PB_DEV bool tri_occludes(vec3 p0, vec3 p1, vec3 p2, const Ray &r, Diag &dg) {
    vec3 normal;
    if (!try_hat(cross(p0 - p1, p2 - p1), normal)) return false;
    return tri_occludes_n(p0, p1, p2, normal, r, dg);
}

// A mesh triangle as it sits in HBM: the three positions and the precomputed unit normal
// (records.h TriRec: bit-identical to what simple.rs:441 computes per test).
struct TriVerts {
    vec3 p0, p1, p2, n;
    uint32_t orig, flags;
};
// (WIDE = two 32-byte loads: the traversal kernels only -- ptxas 12.9 crashes on the 256-bit load
// inside the out-of-line functions of the shade kernels)
template <bool WIDE = false>
PB_DEV TriVerts load_tri(const TriRec *t) {
    const char *b = reinterpret_cast<const char *>(t);
    f4 a, c, d, e;
    if constexpr (WIDE) { ld32(b, a, c); ld32(b + 32, d, e); }
    else { a = ld16(b); c = ld16(b + 16); d = ld16(b + 32); e = ld16(b + 48); }
    TriVerts v;
    v.p0 = mk(a.x, a.y, a.z); v.orig = f2u(a.w);
    v.p1 = mk(c.x, c.y, c.z); v.flags = f2u(c.w);
    v.p2 = mk(d.x, d.y, d.z);
    v.n = mk(e.x, e.y, e.z);
    return v;
}
PB_DEV bool mesh_tri_intersect(const TriVerts &tv, const Ray &r, TriHit &h, Diag &dg) {
    if (tv.flags & PBRS_TRI_DEGENERATE) return false;
    return tri_intersect_n(tv.p0, tv.p1, tv.p2, tv.n, r, h, dg);
}
PB_DEV bool mesh_tri_occludes(const TriVerts &tv, const Ray &r, Diag &dg) {
    if (tv.flags & PBRS_TRI_DEGENERATE) return false;
    return tri_occludes_n(tv.p0, tv.p1, tv.p2, tv.n, r, dg);
}
// The rare tail of mesh_tri_hit_t: the hit position interpolated from the normalised barycentrics
// must not be NaN (simple.rs:462-469).
PB_CALL bool tri_hit_tail(vec3 p0, vec3 p1, vec3 p2, float b0, float b1, float total) {
    b0 = b0 / total; b1 = b1 / total;
    return !any_nan(bary_lerp(p0, p1, p2, b0, b1));
}
// What the closest-hit WALK needs of intersect_triangle: hit or not, and t.  Same decisions as
// tri_intersect_n; the normalisation of the barycentrics and the interpolated position only feed
// the final NaN rejection, which cannot fire when the barycentric sum is an ordinary number and
// the vertices are moderate (all three terms have one sign, so |b_i / total| <= 1 + 3 ulp and
// bary_lerp stays finite): then it is skipped, else evaluated out of line.
PB_DEV bool mesh_tri_hit_t(const TriVerts &tv, const Ray &r, float &t_out, Diag &dg) {
    if (tv.flags & PBRS_TRI_DEGENERATE) return false;
    const float s = dot(tv.n, r.d);
    const bool keep = sign_neg(s);            // facing(): n if n.d < 0, else -n
    const vec3 normal = keep ? tv.n : -tv.n;
    const float nd = dot(normal, r.d);
    if (!(nd <= 0.0f)) flag(dg, P_MISC);
    const float t = dot(normal, tv.p0 - r.o) / nd;
    if (!in_extent(t, r.t_max)) return false;
    const vec3 p = at(r, t);
    const vec3 a = p - tv.p0, b = p - tv.p1, c = p - tv.p2;
    const float b2 = dot(cross(a, b), normal);
    const float b0 = dot(cross(b, c), normal);
    const float b1 = dot(cross(c, a), normal);
    if (is_nan(b0) || is_nan(b1) || is_nan(b2)) return false;
    const bool g0 = b0 > 0.0f, g1 = b1 > 0.0f, g2 = b2 > 0.0f;
    if (!((g0 && g1 && g2) || (!g0 && !g1 && !g2))) return false;
    const float total = b0 + b1 + b2, mag = fabsf(total);
    if (!(mag > 1e-30f && mag < 1e30f) || (tv.flags & PBRS_TRI_HUGE)) {
        if (!tri_hit_tail(tv.p0, tv.p1, tv.p2, b0, b1, total)) return false;
    }
    t_out = t;
    return true;
}

struct MeshHead {
    float bmin[3], bmax[3];
    uint32_t node_base, tri_base, n_tris, root_is_leaf;
};
PB_DEV MeshHead load_mesh_head(const MeshRec *m) {
    const char *b = reinterpret_cast<const char *>(m);
    f4 a = ld16(b), c = ld16(b + 16), d = ld16(b + 32);
    MeshHead h;
    h.bmin[0] = a.x; h.bmin[1] = a.y; h.bmin[2] = a.z; h.bmax[0] = a.w; h.bmax[1] = c.x; h.bmax[2] = c.y;
    h.node_base = f2u(c.z); h.tri_base = f2u(c.w);
    h.n_tris = f2u(d.x); h.root_is_leaf = f2u(d.y);
    return h;
}

// TriangleMesh::intersect_triangle, shape/src/blas.rs:161-207: the geometric hit plus the shading
// interpolation (normal, uv, tangent).  Returns false if the geometric test misses OR the tangent
// check (:193-201) rejects.  `mesh_tri_accepts` below runs it inside the walk only for the
// triangles the host flagged as possibly rejecting (scene_host.cpp tri_may_reject).
struct MeshHit {
    vec3 pos, normal, dpdu;
    float t, u, v;
};
PB_DEV bool mesh_tri_shade(const DeviceScene &sc, uint32_t tri_index, const TriVerts &tv, const Ray &r, MeshHit &out,
                           Diag &dg) {
    TriHit h;
    if (!mesh_tri_intersect(tv, r, h, dg)) return false;
    float b0 = 1.0f - h.b1 - h.b2, b1 = h.b1, b2 = h.b2;
    vec3 hit_by_uv = tv.p0 + (tv.p1 - tv.p0) * b1 + (tv.p2 - tv.p0) * b2;
    if (!(len2(hit_by_uv - h.pos) < 1e-6f)) flag(dg, P_MESH_UV);
    // the triangle's shading record: normals and uvs of its three vertices in TriRec order
    // ((i, k, j) = index_triple: vertex 1 is idx.2, vertex 2 is idx.1, blas.rs:162)
    const char *sr = reinterpret_cast<const char *>(sc.tri_shade + tri_index);
    const f4 s0 = ld16(sr), s1 = ld16(sr + 16), s2 = ld16(sr + 32), s3 = ld16(sr + 48);
    vec3 n0 = mk(s0.x, s0.y, s0.z), n1 = mk(s0.w, s1.x, s1.y), n2 = mk(s1.z, s1.w, s2.x);
    const float ui = s2.y, vi = s2.z, uj = s2.w, vj = s3.x, uk = s3.y, vk = s3.z;
    vec3 bn;
    if (!try_hat(bary_lerp(n0, n1, n2, b0, b1), bn)) bn = h.normal;
    bn = facing(bn, r.d);
    float uu = bary_lerp(ui, uj, uk, b0, b1);
    float vv = bary_lerp(vi, vj, vk, b0, b1);
    float u1 = uj - ui, v1 = vj - vi;
    float u2 = uk - ui, v2 = vk - vi;
    vec3 dpdu = ((tv.p2 - tv.p0) * v2 - (tv.p1 - tv.p0) * v1) / (u1 * v2 - u2 * v1);
    if (!is_fin(len2(dpdu))) dpdu = tv.p1 - tv.p0;
    dpdu = hat(dpdu - projected_onto(dpdu, bn), dg);
    if (fabsf(dot(dpdu, bn)) >= 1e-3f) return false;
    out.pos = h.pos; out.normal = bn; out.dpdu = dpdu; out.t = h.t; out.u = uu; out.v = vv;
    return true;
}

// The same as a real call: the closest-hit walk meets flagged triangles rarely, and its loop is
// sensitive to every inlined instruction (registers, instruction cache).
PB_CALL bool mesh_tri_shade_t(const DeviceScene &sc, uint32_t tri_index, const TriVerts &tv, const Ray &r, float &t, Diag &dg) {
    MeshHit mh;
    const bool hit = mesh_tri_shade(sc, tri_index, tv, r, mh, dg);
    t = mh.t;
    return hit;
}

// ---------------------------------------------------------------------------------------------
// Spheres.  shape/src/simple.rs:207-288.
// ---------------------------------------------------------------------------------------------
PB_DEV bool sphere_roots(vec3 c, float radius, const Ray &r, float &t0, float &t1) {
    vec3 f = r.o - c;
    float a = len2(r.d);
    float b_prime = -dot(f, r.d);
    float delta = radius * radius - len2(f + b_prime / a * r.d);
    if (delta < 0.0f) return false;
    float cc = len2(f) - radius * radius;
    float q = b_prime + signumf(b_prime) * sqrtf(delta * a);
    t0 = cc / q;
    t1 = q / a;
    return true;
}
// `far_root`: the hit is the far root, i.e. the ray leaves a sphere it started in.  The reference
// builds the Interaction of every candidate and its normal.wo >= 0 assert fires there (D1) even when
// the candidate then loses to a nearer hit; the walk records that with this flag instead of
// normalising a normal per candidate (the winner gets the exact test in reconstruct_hit).
PB_DEV bool sphere_hit_t(vec3 c, float radius, const Ray &r, float &t, bool &far_root) {
    float t0, t1;
    far_root = false;
    if (!sphere_roots(c, radius, r, t0, t1)) return false;
    float t_low, t_high;
    if (t0 < t1) { t_low = t0; t_high = t1; } else { t_low = t1; t_high = t0; }
    if (in_extent(t_low, r.t_max)) { t = t_low; return true; }
    if (in_extent(t_high, r.t_max)) { t = t_high; far_root = true; return true; }
    return false;
}
PB_DEV bool sphere_hit_t(vec3 c, float radius, const Ray &r, float &t) {
    bool far_root;
    return sphere_hit_t(c, radius, r, t, far_root);
}
// Q10: both roots must lie inside the extent (simple.rs:287)
PB_DEV bool sphere_occludes(vec3 c, float radius, const Ray &r) {
    float t0, t1;
    if (!sphere_roots(c, radius, r, t0, t1)) return false;
    return in_extent(t0, r.t_max) && in_extent(t1, r.t_max);
}
// One sphere of a sphere BLAS, out of line so the triangle-run loop only carries a call site.
PB_CALL bool ball_test(vec3 c, float radius, const Ray &r, bool any, float &t, Diag &dg) {
    if (any) return sphere_occludes(c, radius, r);
    bool far_root;
    const bool hit = sphere_hit_t(c, radius, r, t, far_root);
    if (far_root) flag(dg, P_SPHERE_INSIDE);
    return hit;
}
// The full Interaction of Sphere::intersect in the sphere's own space.  D1: a hit from
// inside trips Interaction::new's assert upstream; flag it and face the normal to the ray.
// need_uv = false skips the (u, v) of the hit (two FP64 transcendentals): only image textures read them.
PB_DEV bool sphere_intersect(vec3 c, float radius, const Ray &r, Isect &out, Diag &dg, bool need_uv = true) {
    float ray_t;
    if (!sphere_hit_t(c, radius, r, ray_t)) return false;
    vec3 pos = at(r, ray_t);
    vec3 normal = hat(pos - c, dg);
    pos = c + normal * radius * 1.00001f;
    float u = 0.0f, v = 0.0f;
    if (need_uv) {
        float theta = t_acos(normal.y);
        float phi = t_atan2(normal.z, normal.x) + kPi;
        u = phi / (2.0f * kPi);
        v = theta / kPi;
    }
    vec3 dpdu;
    if (!try_hat(mk(-normal.y, normal.x, 0.0f), dpdu)) dpdu = mk(1.0f, 0.0f, 0.0f);
    if (!(len(pos - c) >= radius)) flag(dg, P_MISC);
    vec3 wo = -r.d;
    if (!(dot(normal, wo) >= 0.0f)) {
        flag(dg, P_SPHERE_INSIDE);
        normal = -normal;
    }
    out = isect_new(pos, ray_t, u, v, normal, wo, dg);
    with_dpdu(out, dpdu, dg);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Instance transforms.  geometry/src/transform.rs:267-286 over math/src/hcm.rs:539-544:
// Mat4 * Vec4 = ((c0*v0 + c1*v1) + c2*v2) + c3*v3 per lane, with v3 = 1 for points, 0 for vectors.
// Rows are stored so one 16-byte load brings (c0[r], c1[r], c2[r], c3[r]).
// ---------------------------------------------------------------------------------------------
PB_DEV float row_pt(f4 m, vec3 p) { return m.x * p.x + m.y * p.y + m.z * p.z + m.w * 1.0f; }
PB_DEV float row_vec(f4 m, vec3 v) { return m.x * v.x + m.y * v.y + m.z * v.z + m.w * 0.0f; }
template <bool WIDE = false>
PB_DEV Ray to_object(const InstTravRec *it, const Ray &r, uint32_t &shape_kind, uint32_t &shape_index) {
    const char *b = reinterpret_cast<const char *>(it);
    f4 r0, r1, r2, tail;
    if constexpr (WIDE) { ld32(b, r0, r1); ld32(b + 32, r2, tail); }
    else { r0 = ld16(b); r1 = ld16(b + 16); r2 = ld16(b + 32); tail = ld16(b + 48); }
    shape_kind = f2u(tail.x);
    shape_index = f2u(tail.y);
    Ray o;
    o.o = mk(row_pt(r0, r.o), row_pt(r1, r.o), row_pt(r2, r.o));
    o.d = mk(row_vec(r0, r.d), row_vec(r1, r.d), row_vec(r2, r.d));
    o.t_max = r.t_max;
    return o;
}

#define PBRS_LEAF_BIT 0x80000000u

// The winner of a closest-hit walk: t, instance id, triangle record index (0 for spheres).
struct Hit {
    float t;
    uint32_t inst, tri;
};

}  // namespace pbrs
