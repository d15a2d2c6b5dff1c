// ParallelQuad, Cuboid, Disk and IsolatedTriangle (shape/src/simple.rs:33-195,291-431) over SimpleRec.
//
// These shapes sit directly under an instance (one per TLAS leaf), so they are met once per walk at
// most a few times; the three dispatchers at the bottom are deliberately out of line (PB_CALL) so
// the traversal loop of device_walk.cuh pays one call site for them, not their code.
#pragma once
#include "device_geom.cuh"

namespace pbrs {

struct Simple {
    vec3 a, b, c;
};
PB_DEV Simple load_simple(const SimpleRec *r) {
    const char *p = reinterpret_cast<const char *>(r);
    f4 x = ld16(p), y = ld16(p + 16), z = ld16(p + 32);
    Simple s;
    s.a = mk(x.x, x.y, x.z); s.b = mk(y.x, y.y, y.z); s.c = mk(z.x, z.y, z.z);
    return s;
}
PB_DEV bool inside01(float x) { return 0.0f <= x && x <= 1.0f; }  // math/src/float.rs:210-213

// ---------------------------------------------------------------------------------------------
// ParallelQuad.  Q11: u and v come from cross-product NORMS (simple.rs:136-137), so the mirrored
// extensions of the quad pass `inside` as well and then trip the accurate-vs-coarse assert
// (:140-147), which is counted as P_QUAD; the hit is kept, as the code after the assert would.
// ---------------------------------------------------------------------------------------------
PB_DEV void quad_uv(vec3 origin, vec3 su, vec3 sv, vec3 coarse_hit, float &u, float &v) {
    vec3 d = coarse_hit - origin;
    v = len(cross(su, d)) / len(cross(su, sv));
    u = len(cross(sv, d)) / len(cross(sv, su));
}
struct QuadHit {
    float t, u, v;
    vec3 pos, normal;  // normal: facing, not normalised
};
PB_DEV bool quad_hit(vec3 origin, vec3 su, vec3 sv, const Ray &r, QuadHit &h, Diag &dg) {  // :120-148
    h.normal = facing(cross(su, sv), r.d);
    h.t = dot(origin - r.o, h.normal) / dot(r.d, h.normal);
    if (!in_extent(h.t, r.t_max)) return false;
    vec3 coarse_hit = at(r, h.t);
    quad_uv(origin, su, sv, coarse_hit, h.u, h.v);
    if (!(inside01(h.v) && inside01(h.u))) return false;
    h.pos = origin + h.u * su + sv * h.v;
    if (!(len(h.pos - coarse_hit) < 1e-3f)) flag(dg, P_QUAD);
    return true;
}
PB_DEV bool quad_intersect(vec3 origin, vec3 su, vec3 sv, const Ray &r, Isect &out, Diag &dg) {
    QuadHit h;
    if (!quad_hit(origin, su, sv, r, h, dg)) return false;
    out = isect_new(h.pos, h.t, h.u, h.v, hat(h.normal, dg), -r.d, dg);
    with_dpdu(out, su, dg);
    return true;
}
// :151-163 -- `t` is the reciprocal of the plane distance (Q11), transcribed
PB_DEV bool quad_occludes(vec3 origin, vec3 su, vec3 sv, const Ray &r) {
    vec3 normal = cross(su, sv);
    float t = dot(r.d, normal) / dot(origin - r.o, normal);
    if (!in_extent(t, r.t_max)) return false;
    float u, v;
    quad_uv(origin, su, sv, at(r, t), u, v);
    return inside01(v) && inside01(u);
}

// ---------------------------------------------------------------------------------------------
// Cuboid::intersect, simple.rs:343-411: per-axis slabs with a reciprocal (unlike BBox::intersect),
// the exit face when the interval holds 0 (ray starts inside), else the entry face.
// ---------------------------------------------------------------------------------------------
struct CuboidHit {
    float t, bound;
    int axis;
};
PB_DEV bool cuboid_hit(vec3 mn, vec3 mx, const Ray &r, CuboidHit &out, Diag &dg) {
    CuboidHit hit_min, hit_max;
    hit_min.t = 0.0f; hit_min.bound = PB_INF; hit_min.axis = 0;
    hit_max.t = r.t_max; hit_max.bound = -PB_INF; hit_max.axis = 0;
    for (int axis = 0; axis < 3; ++axis) {
        float inv_dir = 1.0f / comp(r.d, axis);
        float lo = comp(mn, axis), hi = comp(mx, axis);
        float t0 = (lo - comp(r.o, axis)) * inv_dir;
        float t1 = (hi - comp(r.o, axis)) * inv_dir;
        if (t0 > t1) { float s = t0; t0 = t1; t1 = s; s = lo; lo = hi; hi = s; }
        if (t0 > hit_min.t) { hit_min.t = t0; hit_min.bound = lo; hit_min.axis = axis; }
        if (t1 < hit_max.t) { hit_max.t = t1; hit_max.bound = hi; hit_max.axis = axis; }
        if (hit_max.t < hit_min.t) return false;
    }
    if (is_nan(hit_min.t) || is_nan(hit_max.t)) flag(dg, P_MISC);  // Interval::new, math/src/float.rs:162-164
    float lo = hit_min.t < hit_max.t ? hit_min.t : hit_max.t, hi = hit_min.t < hit_max.t ? hit_max.t : hit_min.t;
    out = (0.0f >= lo && 0.0f <= hi) ? hit_max : hit_min;  // Interval::contains(0.0), float.rs:174-176
    return !is_inf(out.bound);
}
PB_DEV bool cuboid_intersect(vec3 mn, vec3 mx, const Ray &r, Isect &out, Diag &dg) {
    CuboidHit h;
    if (!cuboid_hit(mn, mx, r, h, dg)) return false;
    vec3 hit_pos = at(r, h.t);
    set_comp(hit_pos, h.axis, h.bound);
    vec3 normal = mk(0.0f, 0.0f, 0.0f), tangent = mk(0.0f, 0.0f, 0.0f);
    set_comp(normal, h.axis, signumf(comp(r.d, h.axis)) * -1.0f);
    set_comp(tangent, (h.axis + 1) % 3, 1.0f);
    out = isect_new(hit_pos, h.t, 0.5f, 0.5f, normal, -r.d, dg);
    with_dpdu(out, tangent, dg);
    return true;
}
// :412-415: BBox::intersect of the cuboid's own box (true divisions, the ray's extent)
PB_DEV bool cuboid_occludes(vec3 mn, vec3 mx, const Ray &r) {
    float t_low, min_el;
    slab(lane_min(mn.x, mx.x), lane_min(mn.y, mx.y), lane_min(mn.z, mx.z), lane_max(mn.x, mx.x), lane_max(mn.y, mx.y),
         lane_max(mn.z, mx.z), r, t_low, min_el);
    return box_pass(t_low, min_el, r.t_max);
}

// ---------------------------------------------------------------------------------------------
// Disk, simple.rs:306-333.  `occludes` never consults the ray extent (Q10).
// ---------------------------------------------------------------------------------------------
PB_DEV bool disk_hit(vec3 center, vec3 normal, vec3 radial, const Ray &r, float &t, vec3 &isect_point) {
    t = dot(center - r.o, normal) / dot(r.d, normal);
    if (!in_extent(t, r.t_max)) return false;
    isect_point = at(r, t);
    return len2(isect_point - center) <= len2(radial);
}
PB_DEV bool disk_intersect(vec3 center, vec3 dn, vec3 radial, const Ray &r, Isect &out, Diag &dg) {
    float t;
    vec3 isect_point;
    if (!disk_hit(center, dn, radial, r, t, isect_point)) return false;
    vec3 cp = isect_point - center;
    cp = cp - dot(cp, dn) * dn;
    if (!(fabsf(dot(cp, dn)) < 1e-6f)) flag(dg, P_MISC);
    vec3 normal = dn * signumf(dot(dn, -r.d));
    vec3 tangent = hat(cross(normal, cp), dg);
    float u = t_atan2(dot(cross(radial, cp), normal), dot(radial, cp));
    u = fractf(u * kFrac1Pi + 1.0f);
    float v = len(cp) / len(radial);
    out = isect_new(center + cp, t, u, v, normal, -r.d, dg);
    with_dpdu(out, tangent, dg);
    return true;
}
PB_DEV bool disk_occludes(vec3 center, vec3 normal, vec3 radial, const Ray &r) {
    float t = dot(center - r.o, normal) / dot(r.d, normal);
    return len2(at(r, t) - center) <= len2(radial);
}

// ---------------------------------------------------------------------------------------------
// IsolatedTriangle, simple.rs:418-431: intersect_triangle with (u, v) = (b1, b2) and dpdu = p1 - p0.
// ---------------------------------------------------------------------------------------------
PB_DEV bool isotri_shape_intersect(vec3 p0, vec3 p1, vec3 p2, const Ray &r, Isect &out, Diag &dg) {
    TriHit h;
    if (!tri_intersect(p0, p1, p2, r, h, dg)) return false;
    out = isect_new(h.pos, h.t, h.b1, h.b2, h.normal, -r.d, dg);
    with_dpdu(out, p1 - p0, dg);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Dispatch by PBRS_SHAPE_* (out of line on purpose, see the header comment).
// ---------------------------------------------------------------------------------------------
PB_CALL bool simple_hit_t(const SimpleRec *rec, uint32_t kind, const Ray &r, float &t, Diag &dg) {
    const Simple s = load_simple(rec);
    if (kind == PBRS_SHAPE_QUAD) {
        QuadHit h;
        if (!quad_hit(s.a, s.b, s.c, r, h, dg)) return false;
        t = h.t;
        return true;
    }
    if (kind == PBRS_SHAPE_CUBOID) {
        CuboidHit h;
        if (!cuboid_hit(s.a, s.b, r, h, dg)) return false;
        t = h.t;
        return true;
    }
    if (kind == PBRS_SHAPE_TRIANGLE) {
        TriHit h;
        if (!tri_intersect(s.a, s.b, s.c, r, h, dg)) return false;
        t = h.t;
        return true;
    }
    vec3 p;
    return disk_hit(s.a, s.b, s.c, r, t, p);
}
PB_CALL bool simple_occludes(const SimpleRec *rec, uint32_t kind, const Ray &r, Diag &dg) {
    const Simple s = load_simple(rec);
    if (kind == PBRS_SHAPE_TRIANGLE) return tri_occludes(s.a, s.b, s.c, r, dg);
    if (kind == PBRS_SHAPE_QUAD) return quad_occludes(s.a, s.b, s.c, r);
    if (kind == PBRS_SHAPE_CUBOID) return cuboid_occludes(s.a, s.b, r);
    return disk_occludes(s.a, s.b, s.c, r);
}
PB_CALL bool simple_intersect(const SimpleRec *rec, uint32_t kind, const Ray &r, Isect &out, Diag &dg) {
    const Simple s = load_simple(rec);
    out.pos = r.o; out.t = 0.0f; out.u = out.v = 0.0f; out.normal = out.wo = -r.d; out.tangent = mk(1.0f, 0.0f, 0.0f);  // defined on a miss too
    if (kind == PBRS_SHAPE_QUAD) return quad_intersect(s.a, s.b, s.c, r, out, dg);
    if (kind == PBRS_SHAPE_CUBOID) return cuboid_intersect(s.a, s.b, r, out, dg);
    if (kind == PBRS_SHAPE_TRIANGLE) return isotri_shape_intersect(s.a, s.b, s.c, r, out, dg);
    return disk_intersect(s.a, s.b, s.c, r, out, dg);
}

}  // namespace pbrs
