// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.h header).
// Restates shapes (shape/src/simple.rs, shape/src/blas.rs), the TLAS (tlas/src), textures,
// materials, lights and the Scene container, as pointer-chasing CPU data structures that
// follow the reference's own (recursive, boxed) layout.
#pragma once
#include <algorithm>
#include <memory>
#include <vector>

#include "oracle_geom.h"

namespace orc {

// ---------------- Sphere: shape/src/simple.rs:10-31,199-289 ----------------
struct Sphere {
    V3 center;
    float radius;
};
inline BBox sphere_bbox(const Sphere &s) {  // :203-206
    V3 hd = V3{1.0f, 1.0f, 1.0f} * s.radius;
    return bbox_new(s.center - hd, s.center + hd);
}
// :207-267.  D1: a hit from inside would trip Interaction::new's assert; we flag
// it and flip the normal to face the ray (what intersect_triangle does via `facing`).
inline bool sphere_intersect(const Sphere &s, const Ray &r, Interaction *out) {
    V3 f = r.origin - s.center;
    float a = norm_squared(r.dir);
    float b_prime = -dot(f, r.dir);
    float delta = s.radius * s.radius - norm_squared(f + b_prime / a * r.dir);
    if (delta < 0.0f) return false;
    float c = norm_squared(f) - s.radius * s.radius;
    float q = b_prime + f_signum(b_prime) * std::sqrt(delta * a);
    float t0 = c / q, t1 = q / a;
    float t_low, t_high;
    if (t0 < t1) { t_low = t0; t_high = t1; } else { t_low = t1; t_high = t0; }
    bool lo_ok = truncated_t(r, t_low), hi_ok = truncated_t(r, t_high);
    float ray_t;
    if (lo_ok) ray_t = t_low;
    else if (hi_ok) ray_t = t_high;
    else return false;
    V3 pos = position_at(r, ray_t);
    V3 normal = hat(pos - s.center);
    pos = s.center + normal * s.radius * 1.00001f;
    float theta = std::acos(normal.y);
    float phi = std::atan2(normal.z, normal.x) + kPi;
    float u = phi / (2.0f * kPi), v = theta / kPi;
    V3 dpdu;
    if (!try_hat(V3{-normal.y, normal.x, 0.0f}, &dpdu)) dpdu = V3{1, 0, 0};
    if (!(distance_to(pos, s.center) >= s.radius)) panic_flag(P_MISC);
    V3 wo = -r.dir;
    if (!(dot(normal, wo) >= 0.0f)) {  // D1
        panic_flag(P_SPHERE_INSIDE);
        normal = -normal;
    }
    *out = with_dpdu(isect_new(pos, ray_t, u, v, normal, wo), dpdu);
    return true;
}
// :268-288 (Q10: both roots must be inside the extent)
inline bool sphere_occludes(const Sphere &s, const Ray &r) {
    V3 f = r.origin - s.center;
    float a = norm_squared(r.dir);
    float b_prime = -dot(f, r.dir);
    float delta = s.radius * s.radius - norm_squared(f + b_prime / a * r.dir);
    if (delta < 0.0f) return false;
    float c = norm_squared(f) - s.radius * s.radius;
    float q = b_prime + f_signum(b_prime) * std::sqrt(delta * a);
    float t0 = c / q, t1 = q / a;
    return truncated_t(r, t0) && truncated_t(r, t1);
}

// ---------------- triangles: shape/src/simple.rs:435-495 ----------------
inline bool intersect_triangle(V3 p0, V3 p1, V3 p2, const Ray &r, Interaction *out) {
    V3 n;
    if (!try_hat(cross(p0 - p1, p2 - p1), &n)) return false;
    V3 normal = facing(n, r.dir);
    if (!(dot(normal, r.dir) <= 0.0f)) panic_flag(P_MISC);
    float t = dot(normal, p0 - r.origin) / dot(normal, r.dir);
    if (!truncated_t(r, t)) return false;
    V3 p = position_at(r, t);
    float b2 = dot(cross(p - p0, p - p1), normal);
    float b0 = dot(cross(p - p1, p - p2), normal);
    float b1 = dot(cross(p - p2, p - p0), normal);
    if (std::isnan(b0) || std::isnan(b1) || std::isnan(b2)) return false;
    bool g0 = b0 > 0.0f, g1 = b1 > 0.0f, g2 = b2 > 0.0f;
    if (!((g0 && g1 && g2) || (!g0 && !g1 && !g2))) return false;
    float total = b0 + b1 + b2;
    b0 = b0 / total; b1 = b1 / total; b2 = b2 / total;
    V3 hit_pos = barycentric_lerp(p0, p1, p2, b0, b1);
    if (has_nan(hit_pos)) return false;
    *out = isect_new(hit_pos, t, b1, b2, normal, -r.dir);
    return true;
}
inline bool intersect_triangle_pred(V3 p0, V3 p1, V3 p2, const Ray &r) {
    V3 normal;
    if (!try_hat(cross(p0 - p1, p2 - p1), &normal)) return false;
    float t = dot(normal, p0 - r.origin) / dot(normal, r.dir);
    if (!truncated_t(r, t)) return false;
    V3 p = position_at(r, t);
    float b0 = dot(cross(p - p0, p - p1), normal);
    float b1 = dot(cross(p - p1, p - p2), normal);
    float b2 = dot(cross(p - p2, p - p0), normal);
    if (std::isnan(b0) || std::isnan(b1) || std::isnan(b2)) panic_flag(P_MISC);
    bool g0 = b0 > 0.0f, g1 = b1 > 0.0f, g2 = b2 > 0.0f;
    return (g0 && g1 && g2) || (!g0 && !g1 && !g2);
}
// IsolatedTriangle (area-light shape): simple.rs:184-195,418-431
struct IsoTriangle {
    V3 p0, p1, p2;
};
inline BBox isotri_bbox(const IsoTriangle &t) { return bbox_union_pt(bbox_new(t.p0, t.p1), t.p2); }  // simple.rs:422-424
inline bool isotri_occludes(const IsoTriangle &t, const Ray &r) { return intersect_triangle_pred(t.p0, t.p1, t.p2, r); }  // :428-430
inline bool isotri_intersect(const IsoTriangle &t, const Ray &r, Interaction *out) {
    Interaction i;
    if (!intersect_triangle(t.p0, t.p1, t.p2, r, &i)) return false;
    *out = with_dpdu(i, t.p1 - t.p0);
    return true;
}

// ---------------- ParallelQuad: shape/src/simple.rs:69-164 ----------------
struct Quad {
    V3 origin, side_u, side_v;
};
inline BBox quad_bbox(const Quad &q) {  // :105-113
    BBox bu = bbox_new(q.origin, q.origin + q.side_u);
    BBox bv = bbox_new(q.origin + q.side_v, q.origin + q.side_u + q.side_v);
    return bbox_union(bu, bv);
}
// :136-137 (Q11: u and v come from cross-product NORMS, so the mirrored extensions pass too)
inline void quad_uv(const Quad &q, V3 coarse_hit, float *u, float *v) {
    V3 a = q.side_u, b = q.side_v, d = coarse_hit - q.origin;
    *v = norm(cross(a, d)) / norm(cross(a, b));
    *u = norm(cross(b, d)) / norm(cross(b, a));
}
inline bool inside01(float x) { return 0.0f <= x && x <= 1.0f; }  // math/src/float.rs:210-213
// :120-150.  The accurate-vs-coarse assert (:140-147) is counted as P_QUAD and the hit kept.
inline bool quad_intersect(const Quad &q, const Ray &r, Interaction *out) {
    V3 normal = facing(cross(q.side_u, q.side_v), r.dir);
    float t = dot(q.origin - r.origin, normal) / dot(r.dir, normal);
    if (!truncated_t(r, t)) return false;
    V3 coarse_hit = position_at(r, t);
    float u, v;
    quad_uv(q, coarse_hit, &u, &v);
    if (!(inside01(v) && inside01(u))) return false;
    V3 accurate_hit = q.origin + u * q.side_u + q.side_v * v;
    if (!(distance_to(accurate_hit, coarse_hit) < 1e-3f)) panic_flag(P_QUAD);
    *out = with_dpdu(isect_new(accurate_hit, t, u, v, hat(normal), -r.dir), q.side_u);
    return true;
}
// :151-163 (Q11: `t` is the reciprocal of the plane distance)
inline bool quad_occludes(const Quad &q, const Ray &r) {
    V3 normal = cross(q.side_u, q.side_v);
    float t = dot(r.dir, normal) / dot(q.origin - r.origin, normal);
    if (!truncated_t(r, t)) return false;
    float u, v;
    quad_uv(q, position_at(r, t), &u, &v);
    return inside01(v) && inside01(u);
}

// ---------------- Cuboid: shape/src/simple.rs:166-182,335-416 ----------------
struct Cuboid {
    V3 mn, mx;
};
inline Cuboid cuboid_from_points(V3 p0, V3 p1) {  // :173-181 with float::min_max (float.rs:197-203)
    Cuboid c;
    for (int k = 0; k < 3; ++k) {
        if (p0[k] < p1[k]) { c.mn[k] = p0[k]; c.mx[k] = p1[k]; } else { c.mn[k] = p1[k]; c.mx[k] = p0[k]; }
    }
    return c;
}
inline BBox cuboid_bbox(const Cuboid &c) { return bbox_new(c.mn, c.mx); }
// :343-411
inline bool cuboid_intersect(const Cuboid &c, const Ray &r, Interaction *out) {
    struct HitInfo { float t, bound; int axis; };
    HitInfo hit_min{0.0f, kInf, 0}, hit_max{r.t_max, -kInf, 0};
    for (int axis = 0; axis < 3; ++axis) {
        float inv_dir = 1.0f / r.dir[axis];
        float t0 = (c.mn[axis] - r.origin[axis]) * inv_dir;
        float t1 = (c.mx[axis] - r.origin[axis]) * inv_dir;
        HitInfo hit_0{t0, c.mn[axis], axis}, hit_1{t1, c.mx[axis], axis};
        if (t0 > t1) { std::swap(hit_0, hit_1); std::swap(t0, t1); }
        if (t0 > hit_min.t) hit_min = hit_0;
        if (t1 < hit_max.t) hit_max = hit_1;
        if (hit_max.t < hit_min.t) return false;
    }
    if (std::isnan(hit_min.t) || std::isnan(hit_max.t)) panic_flag(P_MISC);  // Interval::new, float.rs:162-164
    float lo = hit_min.t < hit_max.t ? hit_min.t : hit_max.t, hi = hit_min.t < hit_max.t ? hit_max.t : hit_min.t;
    HitInfo h = (0.0f >= lo && 0.0f <= hi) ? hit_max : hit_min;  // Interval::contains(0.0)
    if (std::isinf(h.bound)) return false;
    V3 hit_pos = position_at(r, h.t);
    hit_pos[h.axis] = h.bound;
    V3 normal{0, 0, 0}, tangent{0, 0, 0};
    normal[h.axis] = f_signum(r.dir[h.axis]) * -1.0f;
    tangent[(h.axis + 1) % 3] = 1.0f;
    *out = with_dpdu(isect_new(hit_pos, h.t, 0.5f, 0.5f, normal, -r.dir), tangent);
    return true;
}
inline bool cuboid_occludes(const Cuboid &c, const Ray &r) { return bbox_intersect(cuboid_bbox(c), r); }  // :412-415

// ---------------- Disk: shape/src/simple.rs:33-66,291-333 ----------------
struct Disk {
    V3 center, normal, radial;  // normal is unit length (Disk::new, :42-51)
};
inline BBox disk_bbox(const Disk &d) {  // :298-305
    V3 v1, v2;
    make_coord_system(d.normal, &v1, &v2);
    v1 = v1 * norm(d.radial);
    v2 = v2 * norm(d.radial);
    return bbox_union(bbox_new(d.center + v1 + v2, d.center + v1 - v2), bbox_new(d.center - v1 - v2, d.center - v1 + v2));
}
// :306-326
inline bool disk_intersect(const Disk &d, const Ray &r, Interaction *out) {
    float t = dot(d.center - r.origin, d.normal) / dot(r.dir, d.normal);
    if (!truncated_t(r, t)) return false;
    V3 isect_point = position_at(r, t);
    if (!(squared_distance_to(isect_point, d.center) <= norm_squared(d.radial))) return false;
    V3 cp = isect_point - d.center;
    cp = cp - dot(cp, d.normal) * d.normal;
    if (!(std::fabs(dot(cp, d.normal)) < 1e-6f)) panic_flag(P_MISC);
    V3 normal = d.normal * f_signum(dot(d.normal, -r.dir));
    V3 tangent = hat(cross(normal, cp));
    float u = std::atan2(dot(cross(d.radial, cp), normal), dot(d.radial, cp));
    u = f_fract(u * kFrac1Pi + 1.0f);
    float v = norm(cp) / norm(d.radial);
    *out = with_dpdu(isect_new(d.center + cp, t, u, v, normal, -r.dir), tangent);
    return true;
}
// :328-332 (Q10: the ray extent is not consulted at all)
inline bool disk_occludes(const Disk &d, const Ray &r) {
    float t = dot(d.center - r.origin, d.normal) / dot(r.dir, d.normal);
    return squared_distance_to(position_at(r, t), d.center) <= norm_squared(d.radial);
}

// ---------------- BLAS: shape/src/blas.rs ----------------
struct BlasNode {
    BBox bbox;
    bool is_leaf;
    int axis;
    std::unique_ptr<BlasNode> child[2];
    uint32_t begin, end;  // leaf range
};
struct MeshTri {
    uint32_t i0, i1, i2;  // index_triple as given by the caller
    BBox bbox;
    uint32_t orig;        // the triangle's index in the caller's array (prim id)
};
struct Mesh {
    std::vector<V3> positions, normals;
    std::vector<float> us, vs;
    std::vector<MeshTri> tris;  // permuted in place by the build (blas.rs:388)
    std::vector<Sphere> balls;  // IsoBlas<Sphere> (blas.rs:36-70): primitives are these, `tris` only carries box + index
    std::unique_ptr<BlasNode> root;
};

// crate `partition` 0.1.2 (absent from /root/reference; restated from its published source):
// Hoare-style in-place unstable partition; returns the size of the `true` part.
template <class T, class P>
size_t partition_crate(T *data, size_t len, P pred) {
    if (len == 0) return 0;
    size_t l = 0, r = len - 1;
    while (true) {
        while (l < len && pred(data[l])) ++l;
        while (r > 0 && !pred(data[r])) --r;
        if (l >= r) return l;
        std::swap(data[l], data[r]);
    }
}

// blas.rs:333-420
inline std::unique_ptr<BlasNode> blas_build(std::vector<MeshTri> &shapes, size_t start, size_t end) {
    auto node = std::make_unique<BlasNode>();
    size_t len = end - start;
    if (len <= 4) {
        BBox b = bbox_empty();
        for (size_t i = start; i < end; ++i) b = bbox_union(b, shapes[i].bbox);
        node->bbox = b; node->is_leaf = true; node->begin = (uint32_t)start; node->end = (uint32_t)end;
        return node;
    }
    std::vector<BBox> bboxes;
    bboxes.reserve(len);
    for (size_t i = start; i < end; ++i) bboxes.push_back(shapes[i].bbox);
    BBox centroid = bbox_empty();
    for (auto &b : bboxes) centroid = bbox_union_pt(centroid, bbox_midpoint(b));
    int axis = max_dimension(bbox_diag(centroid));
    if (bbox_diag(centroid)[axis] < 1e-8f) {
        BBox b = bbox_empty();
        for (auto &bb : bboxes) b = bbox_union(b, bb);
        node->bbox = b; node->is_leaf = true; node->begin = (uint32_t)start; node->end = (uint32_t)end;
        return node;
    }
    // sort_by is a stable merge sort
    std::stable_sort(bboxes.begin(), bboxes.end(), [axis](const BBox &a, const BBox &b) {
        return bbox_midpoint(a)[axis] < bbox_midpoint(b)[axis];
    });
    float area_sum = 0.0f;
    for (auto &b : bboxes) area_sum += bbox_area(b);
    float pivot_area = area_sum * 0.5f;
    float partial = 0.0f;
    size_t split_index = 0;
    for (size_t i = 0; i < bboxes.size(); ++i) {
        partial += bbox_area(bboxes[i]);
        if (partial >= pivot_area) { split_index = i; break; }
    }
    float pivot_value = bbox_midpoint(bboxes[split_index])[axis];
    size_t left_len = partition_crate(shapes.data() + start, len, [axis, pivot_value](const MeshTri &t) {
        return bbox_midpoint(t.bbox)[axis] <= pivot_value;
    });
    size_t mid = start + left_len;
    if (left_len == 0 || left_len == len) {
        // blas.rs:403-410 uses select_nth_unstable_by (Rust std's pdqselect), whose order
        // inside the two halves is an implementation detail.  Deviation D4: a stable sort by
        // the same key, split at len/2 (satisfies the same postcondition).
        std::stable_sort(shapes.begin() + start, shapes.begin() + end, [axis](const MeshTri &a, const MeshTri &b) {
            return bbox_midpoint(a.bbox)[axis] < bbox_midpoint(b.bbox)[axis];
        });
        mid = start + len / 2;
    }
    node->is_leaf = false;
    node->axis = axis;
    node->child[0] = blas_build(shapes, start, mid);
    node->child[1] = blas_build(shapes, mid, end);
    node->bbox = bbox_union(node->child[0]->bbox, node->child[1]->bbox);
    return node;
}

// blas.rs:134-159 (TriangleMesh::from_soa)
inline void mesh_build(Mesh &m, const float *P, const float *N, const float *UV, uint32_t nverts,
                       const uint32_t *idx, uint32_t ntris) {
    m.positions.resize(nverts); m.normals.resize(nverts); m.us.resize(nverts); m.vs.resize(nverts);
    for (uint32_t i = 0; i < nverts; ++i) {
        m.positions[i] = V3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        m.normals[i] = N ? V3{N[3 * i], N[3 * i + 1], N[3 * i + 2]} : V3{0, 0, 0};
        m.us[i] = UV ? UV[2 * i] : 0.0f;
        m.vs[i] = UV ? UV[2 * i + 1] : 0.0f;
    }
    m.tris.resize(ntris);
    for (uint32_t t = 0; t < ntris; ++t) {
        uint32_t i = idx[3 * t], j = idx[3 * t + 1], k = idx[3 * t + 2];
        BBox b = bbox_union_pt(bbox_new(m.positions[i], m.positions[j]), m.positions[k]);
        m.tris[t] = MeshTri{i, j, k, b, t};
    }
    m.root = blas_build(m.tris, 0, ntris);
}

// IsoBlas::build, blas.rs:60-69
inline void sphere_blas_build(Mesh &m, const float *centers_radii, uint32_t n) {
    m.balls.resize(n);
    m.tris.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        m.balls[i] = Sphere{V3{centers_radii[4 * i], centers_radii[4 * i + 1], centers_radii[4 * i + 2]}, centers_radii[4 * i + 3]};
        m.tris[i] = MeshTri{0, 0, 0, sphere_bbox(m.balls[i]), i};
    }
    m.root = blas_build(m.tris, 0, n);
}

// blas.rs:161-207 (note the (i, k, j) destructuring: p1 = pos[idx.2], p2 = pos[idx.1])
inline bool mesh_intersect_triangle(const Mesh &m, const MeshTri &tri, const Ray &r, Interaction *out) {
    uint32_t i = tri.i0, k = tri.i1, j = tri.i2;
    V3 p0 = m.positions[i], p1 = m.positions[j], p2 = m.positions[k];
    Interaction hit;
    if (!intersect_triangle(p0, p1, p2, r, &hit)) return false;
    float b0 = 1.0f - hit.u - hit.v, b1 = hit.u, b2 = hit.v;
    V3 hit_by_uv = p0 + (p1 - p0) * b1 + (p2 - p0) * b2;
    if (!(squared_distance_to(hit_by_uv, hit.pos) < 1e-6f)) panic_flag(P_MESH_UV);
    V3 n0 = m.normals[i], n1 = m.normals[j], n2 = m.normals[k];
    V3 bn;
    if (!try_hat(barycentric_lerp(n0, n1, n2, b0, b1), &bn)) bn = hit.normal;
    bn = facing(bn, r.dir);
    float uu = barycentric_lerp(m.us[i], m.us[j], m.us[k], b0, b1);
    float vv = barycentric_lerp(m.vs[i], m.vs[j], m.vs[k], b0, b1);
    float u0 = m.us[i], v0 = m.vs[i];
    float u1 = m.us[j] - u0, v1 = m.vs[j] - v0;
    float u2 = m.us[k] - u0, v2 = m.vs[k] - v0;
    V3 dpdu = ((p2 - p0) * v2 - (p1 - p0) * v1) / (u1 * v2 - u2 * v1);
    if (!std::isfinite(norm_squared(dpdu))) dpdu = p1 - p0;
    dpdu = hat(dpdu - projected_onto(dpdu, bn));
    if (std::fabs(dot(dpdu, bn)) >= 1e-3f) return false;
    *out = with_dpdu(isect_new(hit.pos, hit.ray_t, uu, vv, bn, hit.wo), dpdu);
    return true;
}
inline bool mesh_intersect_triangle_pred(const Mesh &m, const MeshTri &tri, const Ray &r) {
    uint32_t i = tri.i0, k = tri.i1, j = tri.i2;
    return intersect_triangle_pred(m.positions[i], m.positions[j], m.positions[k], r);
}

// blas.rs:422-476.  Note the quirk: after the first popped node that passes
// its box test, ray.t_max becomes outer_hit.ray_t, i.e. +inf until this BLAS finds its own hit;
// the t_max the TLAS handed in only prunes the root.
inline bool mesh_intersect(const Mesh &m, const Ray &r, Interaction *out, uint32_t *prim) {
    if (!m.root) return false;
    if (!bbox_intersect(m.root->bbox, r)) return false;
    std::vector<const BlasNode *> stack;
    stack.reserve(60);
    stack.push_back(m.root.get());
    Interaction outer;
    outer.ray_t = kInf;
    uint32_t best_prim = 0xFFFFFFFFu;
    Ray ray = r;
    while (!stack.empty()) {
        const BlasNode *node = stack.back();
        stack.pop_back();
        if (!bbox_intersect(node->bbox, ray)) continue;
        if (node->is_leaf) {
            for (uint32_t s = node->begin; s < node->end; ++s) {
                Interaction h;
                bool got;
                if (!m.balls.empty()) {  // blas.rs:267-270: the closure is the shape's own intersect
                    if (g_diag) g_diag->n_spheres++;
                    got = sphere_intersect(m.balls[m.tris[s].orig], ray, &h);
                } else {
                    if (g_diag) g_diag->n_tris++;
                    got = mesh_intersect_triangle(m, m.tris[s], ray, &h);
                }
                if (got) {
                    if (h.ray_t < outer.ray_t) { outer = h; best_prim = m.tris[s].orig; }
                }
            }
        } else {
            if (g_diag) g_diag->n_nodes++;
            if (ray.dir[node->axis] > 0.0f) {
                stack.push_back(node->child[1].get());
                stack.push_back(node->child[0].get());
            } else {
                stack.push_back(node->child[0].get());
                stack.push_back(node->child[1].get());
            }
        }
        ray.t_max = outer.ray_t;
    }
    if (outer.ray_t < kInf) { *out = outer; *prim = best_prim; return true; }
    return false;
}
// blas.rs:478-495
inline bool blas_pred(const Mesh &m, const BlasNode *n, const Ray &r) {
    if (!bbox_intersect(n->bbox, r)) return false;
    if (n->is_leaf) {
        for (uint32_t s = n->begin; s < n->end; ++s) {
            if (!m.balls.empty()) {
                if (g_diag) g_diag->n_spheres++;
                if (sphere_occludes(m.balls[m.tris[s].orig], r)) return true;
                continue;
            }
            if (g_diag) g_diag->n_tris++;
            if (mesh_intersect_triangle_pred(m, m.tris[s], r)) return true;
        }
        return false;
    }
    if (g_diag) g_diag->n_nodes++;
    return blas_pred(m, n->child[0].get(), r) || blas_pred(m, n->child[1].get(), r);
}
inline bool mesh_occludes(const Mesh &m, const Ray &r) { return blas_pred(m, m.root.get(), r); }
inline BBox mesh_bbox(const Mesh &m) { return m.root->bbox; }  // blas.rs:313-320

// ---------------- textures: texture/src/lib.rs ----------------
enum TexKind { TEX_SOLID = 0, TEX_IMAGE = 1, TEX_PERLIN = 2 };
struct Texture {
    int kind;
    Color value;
    uint32_t width, height;
    std::vector<Color> data;
    std::vector<V3> rand_vec;
    std::vector<uint32_t> perm_x, perm_y, perm_z;
    float freq;
};
// lib.rs:98-138
inline float perlin_noise(const Texture &t, V3 p) {
    auto split = [](float f, int *i, float *fr) { float fl = std::floor(f); *i = (int)fl; *fr = f - fl; };
    int i, j, k;
    float u, v, w;
    split(p.x * t.freq, &i, &u);
    split(p.y * t.freq, &j, &v);
    split(p.z * t.freq, &k, &w);
    u = u * u * (3.0f - 2.0f * u);
    v = v * v * (3.0f - 2.0f * v);
    w = w * w * (3.0f - 2.0f * w);
    V3 c[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                uint32_t ii = (uint32_t)((i + di) & 255), jj = (uint32_t)((j + dj) & 255), kk = (uint32_t)((k + dk) & 255);
                uint32_t index = t.perm_x[ii] ^ t.perm_y[jj] ^ t.perm_z[kk];
                c[di][dj][dk] = t.rand_vec[index];
            }
    float accum = 0.0f;
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                V3 wv{u - (float)di, v - (float)dj, w - (float)dk};
                float dp = dot(c[di][dj][dk], wv);
                accum += ((float)di * u + (float)(1 - di) * (1.0f - u)) *
                         ((float)dj * v + (float)(1 - dj) * (1.0f - v)) *
                         ((float)dk * w + (float)(1 - dk) * (1.0f - w)) * dp;
            }
    if (!(accum >= -1.0f) || !(accum <= 1.0f)) panic_flag(P_PERLIN);
    return accum;
}
// lib.rs:140-147
inline float perlin_turbulence(const Texture &t, V3 p) {
    float acc = 0.0f;
    for (int i = 0; i < 7; ++i) {
        float s = f_powi(2.0f, i);
        acc = acc + f_powi(0.5f, i) * perlin_noise(t, V3{p.x * s, p.y * s, p.z * s});
    }
    return std::fabs(acc);
}
inline Color texture_value(const Texture &t, float u, float v, V3 p) {
    switch (t.kind) {
    case TEX_SOLID: return t.value;  // :29-33
    case TEX_IMAGE: {                // :211-223 (`as usize` saturates: NaN/negative -> 0)
        u = f_clamp(u, 0.0f, 1.0f);
        v = f_clamp(v, 0.0f, 1.0f);
        float fu = u * (float)t.width, fv = v * (float)t.height;
        size_t col = (fu > 0.0f ? (size_t)fu : 0) % (size_t)t.width;
        size_t row = (fv > 0.0f ? (size_t)fv : 0) % (size_t)t.height;
        return t.data[row * (size_t)t.width + col];
    }
    default: {                       // :150-160 marble
        float s = std::fma(std::sin(t.freq * p.z + 10.0f * perlin_turbulence(t, p)), 0.5f, 0.5f);
        return s * gray(1.0f);
    }
    }
}

// ---------------- materials: material/src/lib.rs ----------------
struct Material {
    int kind;  // pbrs_material_kind
    int tex_kd, tex_ks, tex_kr, tex_kt;
    Color a, b;
    float f[4];
    bool remap;
};
struct Lobes {
    int n;
    Lobe l[5];
};

// ---------------- lights: light/src/lib.rs, light/src/sample_shape.rs ----------------
enum DeltaKind { DELTA_POINT = 0, DELTA_DISTANT = 1 };
struct DeltaLight {
    int kind;
    V3 position;       // point
    Color intensity;   // point: intensity; distant: radiance
    float world_radius;
    V3 casting_dir;
};
enum AreaShapeKind { AREA_SPHERE = 0, AREA_TRIANGLE = 1, AREA_QUAD = 2, AREA_DISK = 3 };
struct AreaLight {
    int shape_kind;
    Sphere sphere;
    IsoTriangle tri;
    Quad quad;
    Disk disk;
    Color emit;
    float area;
};

// sample_shape.rs:184-195
inline Interaction sphere_sample(const Sphere &s, float u, float v) {
    float theta = 2.0f * kPi * u;
    float phi = std::acos(2.0f * v - 1.0f);
    V3 dir{std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta), 2.0f * v - 1.0f};
    return isect_rayless(s.center + s.radius * dir, u, v, dir);
}
// sample_shape.rs:197-236
inline Interaction sphere_sample_towards(const Sphere &s, const Interaction &target, float u, float v) {
    V3 wc = s.center - target.pos;
    if (norm_squared(wc) < f_powi(s.radius, 2)) return sphere_sample(s, u, v);
    float sin_theta_max_2 = f_powi(s.radius, 2) / norm_squared(wc);
    float cos_theta_max = std::sqrt(f_max(1.0f - sin_theta_max_2, 0.0f));
    float cos_t = (1.0f - u) + u * cos_theta_max;
    float sin_t2 = f_max(1.0f - f_powi(cos_t, 2), 0.0f);
    float phi = v * 2.0f * kPi;
    float dc = norm(wc);
    float ds = dc * cos_t - std::sqrt(f_max(f_powi(s.radius, 2) - norm_squared(wc) * sin_t2, 0.0f));
    float cos_alpha = (norm_squared(wc) + f_powi(s.radius, 2) - f_powi(ds, 2)) / (2.0f * dc * s.radius);
    float sin_alpha = std::sqrt(f_max(1.0f - f_powi(cos_alpha, 2), 0.0f));
    V3 n_obj = spherical_direction(sin_alpha, cos_alpha, phi);
    V3 wcx, wcy;
    make_coord_system(-hat(wc), &wcx, &wcy);
    M3 frame{{wcx, wcy, -hat(wc)}};
    V3 n_world = frame * n_obj;
    V3 point = n_world * s.radius + s.center;
    return isect_rayless(point, u, v, n_world);
}
inline float sphere_area(const Sphere &s) { return f_powi(s.radius, 2) * 4.0f * kPi; }  // :253-255
// sample_shape.rs:238-251
inline bool sphere_pdf_at(const Sphere &s, const Interaction &ref, V3 wi, float *pdf) {
    V3 rc = s.center - ref.pos;
    if (norm_squared(rc) < f_powi(s.radius, 2)) { *pdf = 1.0f / sphere_area(s); return true; }
    float sin_theta_max_2 = f_powi(s.radius, 2) / norm_squared(rc);
    float cos_theta_max = std::sqrt(f_max(1.0f - sin_theta_max_2, 0.0f));
    float cos_t = dot(rc, wi) / (norm(rc) * norm(wi));
    if (cos_t > cos_theta_max) { *pdf = 1.0f / (2.0f * kPi * (1.0f - cos_theta_max)); return true; }
    return false;
}
// sample_shape.rs:275-293
inline Interaction isotri_sample(const IsoTriangle &t, float u, float v) {
    if (u + v > 1.0f) { float nu = 1.0f - v, nv = 1.0f - u; u = nu; v = nv; }
    V3 position = t.p0 + (t.p1 - t.p0) * u + (t.p2 - t.p0) * v;
    V3 normal = hat(cross(t.p0 - t.p1, t.p2 - t.p1));
    return isect_rayless(position, u, v, normal);
}
inline float isotri_area(const IsoTriangle &t) { return norm(cross(t.p0 - t.p1, t.p2 - t.p1)) * 0.5f; }

// sample_shape.rs:296-309 (the normal is the unnormalised cross product)
inline Interaction quad_sample(const Quad &q, float u, float v) {
    V3 position = q.origin + u * q.side_u + v * q.side_v;
    return isect_rayless(position, u, v, cross(q.side_u, q.side_v));
}
inline float quad_area(const Quad &q) { return norm(cross(q.side_u, q.side_v)); }
// sample_shape.rs:257-274
inline Interaction disk_sample(const Disk &d, float u, float v) {
    float cos_t, sin_t;
    concentric_sample_disk(u, v, &cos_t, &sin_t);
    V3 radial2 = cross(d.normal, d.radial);
    V3 cp = d.radial * cos_t + radial2 * sin_t;
    return isect_rayless(d.center + cp, u, v, d.normal);
}
inline Interaction disk_sample_towards(const Disk &d, const Interaction &target, float u, float v) {
    Interaction res = disk_sample(d, u, v);
    res.normal = facing(res.normal, target.normal);
    return res;
}
inline float disk_area(const Disk &d) { return norm_squared(d.radial) * kPi; }

inline bool area_shape_intersect(const AreaLight &l, const Ray &r, Interaction *out) {
    switch (l.shape_kind) {
    case AREA_SPHERE: return sphere_intersect(l.sphere, r, out);
    case AREA_TRIANGLE: return isotri_intersect(l.tri, r, out);
    case AREA_QUAD: return quad_intersect(l.quad, r, out);
    default: return disk_intersect(l.disk, r, out);
    }
}
// sample_shape.rs:28-33 default pdf_at (Q12: distance, not distance squared)
inline bool area_shape_pdf_at(const AreaLight &l, const Interaction &ref, V3 wi, float *pdf) {
    if (l.shape_kind == AREA_SPHERE) return sphere_pdf_at(l.sphere, ref, wi, pdf);
    Ray ray = spawn_ray(ref, wi);
    Interaction hit;
    if (!area_shape_intersect(l, ray, &hit)) return false;
    *pdf = distance_to(ref.pos, hit.pos) / (std::fabs(dot(hit.normal, -wi)) * l.area);
    return true;
}
inline Interaction area_shape_sample_towards(const AreaLight &l, const Interaction &t, float u, float v) {
    switch (l.shape_kind) {
    case AREA_SPHERE: return sphere_sample_towards(l.sphere, t, u, v);
    case AREA_TRIANGLE: return isotri_sample(l.tri, u, v);
    case AREA_QUAD: return quad_sample(l.quad, u, v);
    default: return disk_sample_towards(l.disk, t, u, v);
    }
}

// ---------------- instances + TLAS: tlas/src/instance.rs, tlas/src/bvh.rs ----------------
enum ShapeKind { SHAPE_SPHERE = 0, SHAPE_MESH = 1, SHAPE_QUAD = 2, SHAPE_CUBOID = 3, SHAPE_DISK = 4, SHAPE_TRIANGLE = 5 };  // a sphere BLAS is a SHAPE_MESH
struct ShapeRef {
    int kind;
    int index;  // into spheres / meshes
};
struct Instance {
    int shape_id, mtl_id;
    Affine xf;
    int id;
};
struct TlasNode {
    BBox bbox;
    bool is_leaf;
    int inst;
    std::unique_ptr<TlasNode> child[2];
};

enum EnvKind { ENV_CONSTANT = 0, ENV_FN = 1, ENV_IMAGE = 2 };

struct Scene {
    bool has_camera = false, committed = false;
    Camera camera;
    std::vector<Texture> textures;
    std::vector<Material> materials;
    std::vector<ShapeRef> shapes;
    std::vector<Sphere> spheres;
    std::vector<Quad> quads;
    std::vector<Cuboid> cuboids;
    std::vector<Disk> disks;
    std::vector<IsoTriangle> triangles;  // IsolatedTriangle as a Shape (the loader's PLY area lights, loader.rs:408-433)
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<Instance> instances;
    std::vector<DeltaLight> delta_lights;
    std::vector<AreaLight> area_lights;
    int env_kind = ENV_CONSTANT;
    Color env_color = black();
    int env_fn = 0;
    Texture env_image;
    Color env_scale = gray(1.0f);
    std::unique_ptr<TlasNode> tlas;
    uint32_t n_tlas_inner = 0;
};

}  // namespace orc
