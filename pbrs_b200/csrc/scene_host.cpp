// Host-side scene building for the B200 back end.  Emits the flattened records of records.h.
//
// The BVH TOPOLOGY must equal the reference's, because primary-hit parity (bit-exact primitive
// ids including traversal-order tie-breaks) depends on visiting the same boxes in the same order:
//   BLAS: shape/src/blas.rs:333-420 `recursive_build` (leaf <= 4, widest centroid axis,
//         area-median pivot over midpoint-sorted boxes, in-place partition, median fallback)
//   TLAS: tlas/src/bvh.rs:116-152 `build_bvh` (midpoint split on the widest axis of the union
//         box, one instance per leaf)
// Unlike the reference (boxed recursive nodes, one box per node) the builders here work on
// SoA key arrays and a permutation, and write 64-byte two-child records in preorder.
// Compiled with -ffp-contract=off: all box arithmetic is plain IEEE FP32.
#include "scene_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>

namespace pbrs {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

namespace {

constexpr float INF = std::numeric_limits<float>::infinity();

// glam::Vec3A lanes are SSE registers: min/max return the second operand on NaN
// (geometry/src/bvh.rs:28-35,138-143).
inline float lane_min(float a, float b) { return a < b ? a : b; }
inline float lane_max(float a, float b) { return a > b ? a : b; }

inline HostBox box_empty() { return HostBox{{INF, INF, INF}, {-INF, -INF, -INF}}; }
inline HostBox box_of_points(const float *p, const float *q) {  // BBox::new
    HostBox b;
    for (int k = 0; k < 3; ++k) { b.mn[k] = lane_min(p[k], q[k]); b.mx[k] = lane_max(p[k], q[k]); }
    return b;
}
inline void box_grow_point(HostBox &b, const float *p) {  // BBox::union(Point3): scalar f32::min/max
    for (int k = 0; k < 3; ++k) { b.mn[k] = std::fmin(b.mn[k], p[k]); b.mx[k] = std::fmax(b.mx[k], p[k]); }
}
inline HostBox box_merge(const HostBox &a, const HostBox &b) {  // bvh::union(b0, b1)
    HostBox r;
    for (int k = 0; k < 3; ++k) { r.mn[k] = lane_min(a.mn[k], b.mn[k]); r.mx[k] = lane_max(a.mx[k], b.mx[k]); }
    return r;
}
inline float box_mid(const HostBox &b, int k) { return (b.mx[k] - b.mn[k]) * 0.5f + b.mn[k]; }  // bvh.rs:46-49
inline float box_area(const HostBox &b) {  // bvh.rs:75-82
    float x = b.mx[0] - b.mn[0], y = b.mx[1] - b.mn[1], z = b.mx[2] - b.mn[2];
    if (!std::signbit(x) && !std::signbit(y) && !std::signbit(z)) return (x * y + y * z + z * x) * 2.0f;
    return 0.0f;
}
inline int widest_axis(float x, float y, float z) {  // Vec3::max_dimension, math/src/hcm.rs:156-163
    int r = x > y ? 0 : 1;
    float rv = r == 0 ? x : y;
    return z > rv ? 2 : r;
}

constexpr uint32_t PBRS_LEAF_LINK = 0x80000000u;  // == PBRS_LEAF_BIT of the device headers

struct ChildLink {
    bool leaf;
    uint32_t ref;    // node index or first primitive
    uint32_t count;  // leaf primitive count
    HostBox box;
};

void write_child(NodeRec &n, int side, const ChildLink &c) {
    float *mn = side == 0 ? n.lmin : n.rmin, *mx = side == 0 ? n.lmax : n.rmax;
    for (int k = 0; k < 3; ++k) { mn[k] = c.box.mn[k]; mx[k] = c.box.mx[k]; }
    n.child[side] = c.leaf ? (c.ref | PBRS_LEAF_LINK) : c.ref;
    if (c.leaf) {
        n.meta |= side == 0 ? PBRS_NODE_LEFT_LEAF : PBRS_NODE_RIGHT_LEAF;
        n.meta |= (c.count & 0x3FFFu) << (side == 0 ? 4 : 18);
    }
}

// ------------------------------------------------------------------------------------------
// BLAS
// ------------------------------------------------------------------------------------------
struct BlasBuilder {
    const std::vector<HostBox> &tbox;       // per triangle (caller order)
    std::vector<float> mid[3], area;        // keys, per triangle
    std::vector<uint32_t> perm;             // triangles in current (build) order
    std::vector<NodeRec> nodes;
    std::vector<uint32_t> leaf_last;        // position (in build order) of each leaf's last triangle
    uint32_t max_depth = 0;

    explicit BlasBuilder(const std::vector<HostBox> &boxes) : tbox(boxes) {
        size_t n = boxes.size();
        for (int k = 0; k < 3; ++k) mid[k].resize(n);
        area.resize(n);
        perm.resize(n);
        for (size_t i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) mid[k][i] = box_mid(boxes[i], k);
            area[i] = box_area(boxes[i]);
            perm[i] = (uint32_t)i;
        }
    }

    ChildLink make_leaf(size_t s, size_t e) {
        HostBox b = box_empty();
        for (size_t i = s; i < e; ++i) b = box_merge(b, tbox[perm[i]]);  // .sum::<BBox>()
        if (e > s) leaf_last.push_back((uint32_t)(e - 1));
        const uint32_t run = (e - s) <= PBRS_LEAF_COUNT_MAX ? (uint32_t)(e - s) : 0u;  // short runs carry their length
        return ChildLink{true, (uint32_t)s | (run << PBRS_LEAF_COUNT_SHIFT), (uint32_t)std::min<size_t>(e - s, PBRS_MAX_LEAF_PRIMS), b};
    }

    ChildLink build(size_t s, size_t e, uint32_t depth = 1) {
        size_t len = e - s;
        if (len <= 4) return make_leaf(s, e);  // blas.rs:338-344
        max_depth = std::max(max_depth, depth);

        // centroid box and split axis (blas.rs:350-360)
        float cmn[3] = {INF, INF, INF}, cmx[3] = {-INF, -INF, -INF};
        for (size_t i = s; i < e; ++i)
            for (int k = 0; k < 3; ++k) {
                float m = mid[k][perm[i]];
                cmn[k] = std::fmin(cmn[k], m);
                cmx[k] = std::fmax(cmx[k], m);
            }
        float dx = cmx[0] - cmn[0], dy = cmx[1] - cmn[1], dz = cmx[2] - cmn[2];
        int axis = widest_axis(dx, dy, dz);
        float extent = axis == 0 ? dx : (axis == 1 ? dy : dz);
        if (extent < 1e-8f) return make_leaf(s, e);

        // stable sort of the boxes by midpoint on the axis, then the area-median pivot
        // (blas.rs:366-385).  Sorting (key, triangle) pairs is equivalent to sorting the boxes.
        const std::vector<float> &key = mid[axis];
        std::vector<uint32_t> sorted(perm.begin() + s, perm.begin() + e);
        std::stable_sort(sorted.begin(), sorted.end(), [&key](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        float total = 0.0f;
        for (uint32_t t : sorted) total += area[t];
        float half = total * 0.5f;
        float run = 0.0f;
        size_t split_index = 0;
        for (size_t i = 0; i < sorted.size(); ++i) {
            run += area[sorted[i]];
            if (run >= half) { split_index = i; break; }
        }
        float pivot = key[sorted[split_index]];

        // crate `partition` 0.1.2 (pinned by shape/Cargo.toml, not vendored): Hoare-style
        // in-place unstable partition, true-part first (blas.rs:388-390).
        uint32_t *d = perm.data() + s;
        size_t l = 0, r = len - 1;
        while (true) {
            while (l < len && key[d[l]] <= pivot) ++l;
            while (r > 0 && !(key[d[r]] <= pivot)) --r;
            if (l >= r) break;
            std::swap(d[l], d[r]);
        }
        size_t left_len = l;
        size_t m = s + left_len;
        if (left_len == 0 || left_len == len) {
            // blas.rs:403-410 `select_nth_unstable_by`: order inside the halves is an
            // implementation detail of Rust's std; deviation D4: stable sort, split
            // at len/2.
            std::stable_sort(perm.begin() + s, perm.begin() + e, [&key](uint32_t a, uint32_t b) { return key[a] < key[b]; });
            m = s + len / 2;
        }

        uint32_t self = (uint32_t)nodes.size();
        nodes.emplace_back();
        ChildLink lc = build(s, m, depth + 1);
        ChildLink rc = build(m, e, depth + 1);
        NodeRec rec;
        std::memset(&rec, 0, sizeof rec);
        rec.meta = (uint32_t)axis;
        write_child(rec, 0, lc);
        write_child(rec, 1, rc);
        nodes[self] = rec;
        return ChildLink{false, self, 0, box_merge(lc.box, rc.box)};
    }
};

inline void apply_point(const float m[4][4], const float p[3], float out[3]) {
    // Mat4 * Vec4(p, 1): ((c0*x + c1*y) + c2*z) + c3*1, math/src/hcm.rs:539-544
    for (int r = 0; r < 3; ++r) out[r] = m[0][r] * p[0] + m[1][r] * p[1] + m[2][r] * p[2] + m[3][r] * 1.0f;
}

// geometry/src/transform.rs:287-308
HostBox transform_box(const float m[4][4], const HostBox &b) {
    HostBox res = box_empty();
    float diag[3] = {b.mx[0] - b.mn[0], b.mx[1] - b.mn[1], b.mx[2] - b.mn[2]};
    for (int i = 0; i < 8; ++i) {
        float corner[3];
        apply_point(m, b.mn, corner);
        for (int a = 0; a < 3; ++a)
            if (i & (1 << a))
                for (int r = 0; r < 3; ++r) corner[r] = corner[r] + m[a][r] * diag[a];
        box_grow_point(res, corner);
    }
    return res;
}

// ------------------------------------------------------------------------------------------
// TLAS
// ------------------------------------------------------------------------------------------
struct TlasBuilder {
    const std::vector<HostInstance> &inst;
    std::vector<NodeRec> nodes;
    uint32_t max_depth = 0;

    ChildLink build(std::vector<uint32_t> ids, uint32_t depth = 1) {
        if (ids.size() == 1) return ChildLink{true, ids[0], 1, inst[ids[0]].box};
        max_depth = std::max(max_depth, depth);
        size_t n = ids.size();
        HostBox all = box_empty();
        for (uint32_t i : ids) all = box_merge(all, inst[i].box);
        int axis = widest_axis(all.mx[0] - all.mn[0], all.mx[1] - all.mn[1], all.mx[2] - all.mn[2]);
        float plane = box_mid(all, axis);
        std::vector<uint32_t> left, right;
        for (uint32_t i : ids) (box_mid(inst[i].box, axis) < plane ? left : right).push_back(i);
        if (left.empty()) {
            for (size_t k = 0; k < n / 2; ++k) { left.push_back(right.back()); right.pop_back(); }
        } else if (right.empty()) {
            for (size_t k = 0; k < n / 2; ++k) { right.push_back(left.back()); left.pop_back(); }
        }
        uint32_t self = (uint32_t)nodes.size();
        nodes.emplace_back();
        ChildLink lc = build(std::move(left), depth + 1);
        ChildLink rc = build(std::move(right), depth + 1);
        NodeRec rec;
        std::memset(&rec, 0, sizeof rec);
        write_child(rec, 0, lc);
        write_child(rec, 1, rc);
        nodes[self] = rec;
        return ChildLink{false, self, 0, box_merge(lc.box, rc.box)};
    }
};

inline void v_sub(const float *a, const float *b, float *o) { for (int k = 0; k < 3; ++k) o[k] = a[k] - b[k]; }
inline void v_cross(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float v_dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void v_hat(const float *a, float *o) {  // Vec3::hat, math/src/hcm.rs:112-117
    float inv = 1.0f / std::sqrt(v_dot(a, a));
    for (int k = 0; k < 3; ++k) o[k] = a[k] * inv;
}

// Can TriangleMesh::intersect_triangle's tangent check (blas.rs:193-201) ever reject a hit on
// this triangle?  Conservative (double precision, wide margins): returns true = "maybe".
inline HostBox sphere_box(const SphereRec &sp) {  // Sphere::bbox, shape/src/simple.rs:203-206
    float hd[3] = {1.0f * sp.r, 1.0f * sp.r, 1.0f * sp.r};
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = sp.c[k] - hd[k]; hi[k] = sp.c[k] + hd[k]; }
    return box_of_points(lo, hi);
}
// make_coord_system, math/src/hcm.rs:595-605 (abs_min_dimension :165-176)
inline void host_make_coord_system(const float v[3], float o1[3], float o2[3]) {
    float ax = std::fabs(v[0]), ay = std::fabs(v[1]), az = std::fabs(v[2]);
    int i0 = ax < ay ? 0 : 1;
    i0 = (i0 == 0 ? ax : ay) < az ? i0 : 2;
    int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    float v1[3] = {0.0f, 0.0f, 0.0f}, v2[3];
    v1[i1] = v[i2];
    v1[i2] = -v[i1];
    v_cross(v, v1, v2);
    v_hat(v1, o1);
    v_hat(v2, o2);
}

bool tri_may_reject(const HostMesh &m, uint32_t t) {
    uint32_t i = m.idx[3 * t], k = m.idx[3 * t + 1], j = m.idx[3 * t + 2];
    auto P = [&](uint32_t v, int c) { return (double)m.P[3 * v + c]; };
    auto N = [&](uint32_t v, int c) { return (double)m.N[3 * v + c]; };
    double e1[3], e2[3], g[3];
    for (int c = 0; c < 3; ++c) { e1[c] = P(j, c) - P(i, c); e2[c] = P(k, c) - P(i, c); }
    g[0] = e1[1] * e2[2] - e1[2] * e2[1]; g[1] = e1[2] * e2[0] - e1[0] * e2[2]; g[2] = e1[0] * e2[1] - e1[1] * e2[0];
    double gl = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (!(gl > 0.0) || !std::isfinite(gl)) return true;
    for (int c = 0; c < 3; ++c) g[c] /= gl;
    // candidate tangents: uv-derived and the p1 - p0 fallback
    double u1 = (double)m.UV[2 * j] - m.UV[2 * i], v1 = (double)m.UV[2 * j + 1] - m.UV[2 * i + 1];
    double u2 = (double)m.UV[2 * k] - m.UV[2 * i], v2 = (double)m.UV[2 * k + 1] - m.UV[2 * i + 1];
    double cand[2][3];
    int nc = 0;
    double det = u1 * v2 - u2 * v1;
    {
        double d[3];
        for (int c = 0; c < 3; ++c) d[c] = (e2[c] * v2 - e1[c] * v1) / det;
        double dl = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (std::isfinite(dl)) {
            if (!(dl > 1e-30)) return true;  // zero tangent: NaN frame, keep the exact path
            // FP32 cancellation in either difference makes the device-side direction unreliable
            double e1l = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
            double e2l = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
            if (std::fabs(det) < 1e-3 * (std::fabs(u1 * v2) + std::fabs(u2 * v1))) return true;
            if (dl * std::fabs(det) < 1e-3 * (e2l * std::fabs(v2) + e1l * std::fabs(v1))) return true;
            for (int c = 0; c < 3; ++c) cand[nc][c] = d[c] / dl;
            nc++;
        }
    }
    {
        double dl = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        if (!(dl > 1e-30) || !std::isfinite(dl)) return true;
        for (int c = 0; c < 3; ++c) cand[nc][c] = e1[c] / dl;
        nc++;
    }
    uint32_t vs[3] = {i, j, k};
    double nh[3][3];
    int zeros = 0;
    for (int a = 0; a < 3; ++a) {
        double l = std::sqrt(N(vs[a], 0) * N(vs[a], 0) + N(vs[a], 1) * N(vs[a], 1) + N(vs[a], 2) * N(vs[a], 2));
        if (l == 0.0) { zeros++; continue; }
        if (!(l > 1e-12) || !(l < 1e12)) return true;
        for (int c = 0; c < 3; ++c) nh[a][c] = N(vs[a], c) / l;
    }
    if (zeros == 3) {  // shading normal = geometric normal
        for (int q = 0; q < nc; ++q)
            if (std::fabs(cand[q][0] * g[0] + cand[q][1] * g[1] + cand[q][2] * g[2]) > 0.5) return true;
        return false;
    }
    if (zeros != 0) return true;
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (nh[a][0] * nh[b][0] + nh[a][1] * nh[b][1] + nh[a][2] * nh[b][2] < 0.5) return true;
    for (int q = 0; q < nc; ++q) {
        for (int a = 0; a < 3; ++a)
            if (std::fabs(cand[q][0] * nh[a][0] + cand[q][1] * nh[a][1] + cand[q][2] * nh[a][2]) > 0.5) return true;
        // the lerp may also degenerate to the geometric normal
        if (std::fabs(cand[q][0] * g[0] + cand[q][1] * g[1] + cand[q][2] * g[2]) > 0.5) return true;
    }
    return false;
}

}  // namespace

// geometry/src/camera.rs:19-44,65-77
int host_set_camera(SceneImpl &s, uint32_t w, uint32_t h, float fov, const float *eye, const float *target, const float *up) {
    float aspect = (float)w / (float)h;
    float half_v = std::tan(fov * 0.5f);
    float half_h = half_v * aspect;
    float a[3] = {half_h / (float)(w / 2), 0.0f, 0.0f};
    float b[3] = {0.0f, -half_v / (float)(h / 2), 0.0f};
    float c[3] = {-half_h, half_v, 1.0f};
    float fwd[3], right[3], up2[3], tmp[3];
    v_sub(target, eye, tmp);
    v_hat(tmp, fwd);
    v_cross(up, fwd, tmp);
    v_hat(tmp, right);
    v_cross(fwd, right, up2);
    const float *cols[3] = {right, up2, fwd};
    auto rot = [&](const float *v, float *o) {  // Mat3 * Vec3, math/src/hcm.rs:448-453
        for (int r = 0; r < 3; ++r) o[r] = cols[0][r] * v[0] + cols[1][r] * v[1] + cols[2][r] * v[2];
    };
    rot(a, s.cam.a);
    rot(b, s.cam.b);
    rot(c, s.cam.c);
    for (int k = 0; k < 3; ++k) s.cam.center[k] = eye[k];
    s.cam.width = w;
    s.cam.height = h;
    for (int k = 0; k < 3; ++k)
        if (!std::isfinite(s.cam.a[k]) || !std::isfinite(s.cam.b[k]) || !std::isfinite(s.cam.c[k])) {
            set_error("set_camera: degenerate look-at (the reference panics in Vec3::hat)");
            return PBRS_ERR_INVALID_ARG;
        }
    s.has_camera = true;
    return 0;
}

int host_add_mesh(SceneImpl &s, const float *P, const float *N, const float *UV, uint32_t nverts, const uint32_t *idx, uint32_t ntris) {
    HostMesh m;
    m.P.assign(P, P + 3 * (size_t)nverts);
    if (N) m.N.assign(N, N + 3 * (size_t)nverts); else m.N.assign(3 * (size_t)nverts, 0.0f);
    if (UV) m.UV.assign(UV, UV + 2 * (size_t)nverts); else m.UV.assign(2 * (size_t)nverts, 0.0f);
    m.idx.assign(idx, idx + 3 * (size_t)ntris);
    for (float v : m.P)
        if (std::isnan(v)) { set_error("add_mesh: NaN position (the reference's BVH build panics on partial_cmp)"); return PBRS_ERR_INVALID_ARG; }
    for (uint32_t v : m.idx)
        if (v >= nverts) { set_error("add_mesh: vertex index out of range"); return PBRS_ERR_INVALID_ARG; }
    s.meshes.push_back(std::move(m));
    s.shapes.push_back(HostShape{PBRS_SHAPE_MESH, (uint32_t)s.meshes.size() - 1});
    return (int)s.shapes.size() - 1;
}

// IsoBlas::<Sphere>::build, shape/src/blas.rs:60-69
int host_add_sphere_blas(SceneImpl &s, const float *centers_radii, uint32_t n) {
    HostMesh m;
    m.balls.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 4; ++k)
            if (std::isnan(centers_radii[4 * i + k])) { set_error("add_sphere_blas: NaN (Sphere::from_raw asserts, shape/src/simple.rs:21-22)"); return PBRS_ERR_INVALID_ARG; }
        for (int k = 0; k < 3; ++k) m.balls[i].c[k] = centers_radii[4 * i + k];
        m.balls[i].r = centers_radii[4 * i + 3];
    }
    s.meshes.push_back(std::move(m));
    s.shapes.push_back(HostShape{PBRS_SHAPE_MESH, (uint32_t)s.meshes.size() - 1});
    return (int)s.shapes.size() - 1;
}

int host_add_simple(SceneImpl &s, uint32_t kind, const float a[3], const float b[3], const float c[3]) {
    SimpleRec r;
    std::memset(&r, 0, sizeof r);
    for (int k = 0; k < 3; ++k) { r.a[k] = a[k]; r.b[k] = b[k]; r.c[k] = c ? c[k] : 0.0f; }
    s.simples.push_back(r);
    s.shapes.push_back(HostShape{kind, (uint32_t)s.simples.size() - 1});
    return (int)s.shapes.size() - 1;
}

// Disk::new, shape/src/simple.rs:42-52: normalises the normal; its two asserts are argument errors.
int host_make_disk(const float center[3], const float normal[3], const float radial[3], float out_normal[3]) {
    (void)center;
    float n2 = v_dot(normal, normal);
    if (!(n2 != 0.0f && std::isfinite(n2))) { set_error("disk: the normal cannot be normalised"); return PBRS_ERR_INVALID_ARG; }
    v_hat(normal, out_normal);
    float r2 = v_dot(radial, radial);
    if (!std::isfinite(r2)) { set_error("disk: radial is not finite (simple.rs:44)"); return PBRS_ERR_INVALID_ARG; }
    if (!(std::fabs(v_dot(radial, out_normal)) < 1e-6f)) { set_error("disk: radial is not perpendicular to the normal (simple.rs:45)"); return PBRS_ERR_INVALID_ARG; }
    return 0;
}

// Shape::bbox of the simple shapes in object space.
static HostBox simple_box(uint32_t kind, const SimpleRec &r) {
    if (kind == PBRS_SHAPE_QUAD) {  // simple.rs:105-113
        float ou[3], ov[3], ouv[3];
        for (int k = 0; k < 3; ++k) { ou[k] = r.a[k] + r.b[k]; ov[k] = r.a[k] + r.c[k]; ouv[k] = r.a[k] + r.b[k] + r.c[k]; }
        return box_merge(box_of_points(r.a, ou), box_of_points(ov, ouv));
    }
    if (kind == PBRS_SHAPE_CUBOID) return box_of_points(r.a, r.b);  // :339-341
    if (kind == PBRS_SHAPE_TRIANGLE) {  // :422-424: BBox::new(p0, p1).union(p2)
        HostBox b = box_of_points(r.a, r.b);
        box_grow_point(b, r.c);
        return b;
    }
    // Disk, :298-305: make_coord_system(normal) scaled by |radial|
    float v1[3], v2[3];
    host_make_coord_system(r.b, v1, v2);
    float rn = std::sqrt(v_dot(r.c, r.c));
    float p[4][3];
    for (int k = 0; k < 3; ++k) {
        float a = v1[k] * rn, b = v2[k] * rn;
        p[0][k] = r.a[k] + a + b; p[1][k] = r.a[k] + a - b; p[2][k] = r.a[k] - a - b; p[3][k] = r.a[k] - a + b;
    }
    return box_merge(box_of_points(p[0], p[1]), box_of_points(p[2], p[3]));
}

int host_add_instance(SceneImpl &s, int shape, int mtl, const float *fwd, const float *inv) {
    HostInstance in;
    in.shape = shape;
    in.material = mtl;
    in.identity = !(fwd && inv);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            in.fwd[c][r] = in.identity ? (c == r ? 1.0f : 0.0f) : fwd[4 * c + r];
            in.inv[c][r] = in.identity ? (c == r ? 1.0f : 0.0f) : inv[4 * c + r];
        }
    for (int c = 0; c < 4; ++c) {
        float want = c == 3 ? 1.0f : 0.0f;
        if (in.fwd[c][3] != want || in.inv[c][3] != want) {
            set_error("add_instance: bottom matrix row must be (0,0,0,1) (geometry/src/transform.rs:277)");
            return PBRS_ERR_INVALID_ARG;
        }
    }
    s.instances.push_back(in);
    return (int)s.instances.size() - 1;
}

int host_build(SceneImpl &s) {
    // ---- BLAS per mesh ----
    uint32_t total_nodes = 0, total_tris = 0, total_balls = 0;
    for (HostMesh &m : s.meshes) {
        const bool is_balls = !m.balls.empty();
        size_t nt = is_balls ? m.balls.size() : m.idx.size() / 3;
        std::vector<HostBox> boxes(nt);
        for (size_t t = 0; is_balls && t < nt; ++t) boxes[t] = sphere_box(m.balls[t]);  // blas.rs:66: |s| s.bbox()
        for (size_t t = 0; !is_balls && t < nt; ++t) {  // blas.rs:141: BBox::new(p[i], p[j]).union(p[k])
            const float *pi = &m.P[3 * m.idx[3 * t]], *pj = &m.P[3 * m.idx[3 * t + 1]], *pk = &m.P[3 * m.idx[3 * t + 2]];
            HostBox b = box_of_points(pi, pj);
            box_grow_point(b, pk);
            boxes[t] = b;
        }
        if (nt > PBRS_LEAF_FIRST_MASK) { set_error("commit: a mesh has more primitives than a leaf link can address (2^28)"); return PBRS_ERR_UNSUPPORTED; }
        BlasBuilder bb(boxes);
        ChildLink root = bb.build(0, nt);
        if (bb.max_depth > 60) { set_error("commit: a BLAS is deeper than 60 levels (traversal stack)"); return PBRS_ERR_UNSUPPORTED; }
        m.depth = bb.max_depth;
        m.nodes = std::move(bb.nodes);
        m.order = std::move(bb.perm);
        m.leaf_last = std::move(bb.leaf_last);
        m.root_box = root.box;
        m.root_is_leaf = root.leaf;
        total_nodes += (uint32_t)m.nodes.size();
        if (is_balls) total_balls += (uint32_t)nt; else total_tris += (uint32_t)nt;
    }
    // ---- instance boxes (tlas/src/instance.rs:47-49) ----
    for (HostInstance &in : s.instances) {
        const HostShape &sh = s.shapes[in.shape];
        HostBox sb;
        if (sh.kind == PBRS_SHAPE_SPHERE) {
            sb = sphere_box(s.spheres[sh.index]);
        } else if (sh.kind != PBRS_SHAPE_MESH) {
            sb = simple_box(sh.kind, s.simples[sh.index]);
        } else {
            sb = s.meshes[sh.index].root_box;
        }
        in.box = transform_box(in.fwd, sb);
    }
    // ---- TLAS ----
    if (s.instances.size() > PBRS_LEAF_FIRST_MASK) { set_error("commit: more instances than a leaf link can address (2^28)"); return PBRS_ERR_UNSUPPORTED; }
    TlasBuilder tb{s.instances, {}, 0};
    std::vector<uint32_t> all(s.instances.size());
    std::iota(all.begin(), all.end(), 0u);
    ChildLink root = tb.build(std::move(all));
    s.tlas_nodes = std::move(tb.nodes);
    s.tlas_box = root.box;
    s.tlas_root_is_leaf = root.leaf;
    s.tlas_depth = tb.max_depth;
    if (tb.max_depth > 60) { set_error("commit: the TLAS is deeper than 60 levels (traversal stack)"); return PBRS_ERR_UNSUPPORTED; }
    {   // the walk keeps at most one entry per TLAS level, the EXIT tag and one far child per BLAS level
        // (device_walk.cuh): with this bound the stack-overflow flag of the kernels is provably dead
        uint32_t deepest_blas = 0;
        for (const HostMesh &m : s.meshes) deepest_blas = std::max(deepest_blas, m.depth);
        if (tb.max_depth + deepest_blas + 2 > PBRS_WALK_STACK_ENTRIES) { set_error("commit: TLAS + BLAS depth exceeds the traversal stack"); return PBRS_ERR_UNSUPPORTED; }
    }
    // scene/src/lib.rs:54-58: distant lights get half the TLAS diagonal as world radius
    float d[3] = {root.box.mx[0] - root.box.mn[0], root.box.mx[1] - root.box.mn[1], root.box.mx[2] - root.box.mn[2]};
    float half_diag = std::sqrt(v_dot(d, d)) * 0.5f;
    for (DeltaLightRec &l : s.delta_lights)
        if (l.kind == PBRS_LIGHT_DISTANT && !(l.world_radius > 0.0f && std::isfinite(l.world_radius))) l.world_radius = half_diag;

    std::memset(&s.info, 0, sizeof s.info);
    s.info.width = s.cam.width;
    s.info.height = s.cam.height;
    s.info.n_instances = (uint32_t)s.instances.size();
    s.info.n_meshes = (uint32_t)s.meshes.size();
    s.info.n_spheres = (uint32_t)s.spheres.size() + total_balls;  // incl. the spheres of sphere BLASes
    s.info.n_triangles = total_tris;
    s.info.n_tlas_nodes = (uint32_t)s.tlas_nodes.size();
    s.info.n_blas_nodes = total_nodes;
    bool has_env = s.env_kind != PBRS_ENV_KIND_CONSTANT || !(s.env_color[0] <= 0.0f && s.env_color[1] <= 0.0f && s.env_color[2] <= 0.0f);
    s.info.n_lights = (uint32_t)(s.delta_lights.size() + s.area_lights.size() + (has_env ? 1 : 0));
    for (int k = 0; k < 3; ++k) { s.info.world_min[k] = root.box.mn[k]; s.info.world_max[k] = root.box.mx[k]; }
    return 0;
}

bool host_tri_may_reject(const HostMesh &m, uint32_t t) { return tri_may_reject(m, t); }

// Flattens the host scene into the record arrays of records.h (DESIGN.md "Data layout"): a mesh's
// inner nodes in preorder, its triangles in leaf order with the (i, k, j) vertex swap applied.
void flatten_scene(const SceneImpl &s, FlatScene &f) {
    auto &blas_nodes = f.blas_nodes; auto &tris = f.tris; auto &meshes = f.meshes; auto &tri_shade = f.tri_shade;
    auto &trav = f.trav; auto &shade = f.shade; auto &textures = f.textures; auto &texels = f.texels;
    auto &perlin_vec = f.perlin_vec; auto &perlin_perm = f.perlin_perm;
    // ---- meshes: nodes, triangles, attributes ----
    for (const HostMesh &m : s.meshes) {
        MeshRec r;
        std::memset(&r, 0, sizeof r);
        for (int k = 0; k < 3; ++k) { r.bmin[k] = m.root_box.mn[k]; r.bmax[k] = m.root_box.mx[k]; }
        r.node_base = (uint32_t)blas_nodes.size();
        r.tri_base = (uint32_t)tris.size();
        r.n_tris = (uint32_t)m.order.size();
        r.root_is_leaf = m.root_is_leaf ? 1u : 0u;
        meshes.push_back(r);
        blas_nodes.insert(blas_nodes.end(), m.nodes.begin(), m.nodes.end());
        size_t t0 = tris.size();
        tris.resize(t0 + m.order.size());
        tri_shade.resize(t0 + m.order.size());
        for (size_t q = 0; q < m.order.size() && !m.balls.empty(); ++q) {
            uint32_t t = m.order[q];
            TriRec &tr = tris[t0 + q];
            std::memset(&tr, 0, sizeof tr);
            for (int c = 0; c < 3; ++c) tr.p0[c] = m.balls[t].c[c];
            tr.p1[0] = m.balls[t].r;
            tr.orig = t;
            tr.flags = PBRS_TRI_SPHERE;
            std::memset(&tri_shade[t0 + q], 0, sizeof(TriShadeRec));
        }
        for (size_t q = 0; q < m.order.size() && m.balls.empty(); ++q) {
            uint32_t t = m.order[q];
            // (i, k, j) = index_triple, shape/src/blas.rs:162-163
            uint32_t i = m.idx[3 * t], k = m.idx[3 * t + 1], j = m.idx[3 * t + 2];
            TriRec &tr = tris[t0 + q];
            for (int c = 0; c < 3; ++c) { tr.p0[c] = m.P[3 * i + c]; tr.p1[c] = m.P[3 * j + c]; tr.p2[c] = m.P[3 * k + c]; }
            tr.orig = t;
            tr.flags = host_tri_may_reject(m, t) ? PBRS_TRI_CHECK_SHADING : 0u;
            tr.pad = 0u; tr.pad2 = 0u;
            {   // n = try_hat((p0 - p1) x (p2 - p1)), shape/src/simple.rs:441,481 over math/src/hcm.rs:118-121
                float e0[3], e1[3], cr[3];
                v_sub(tr.p0, tr.p1, e0);
                v_sub(tr.p2, tr.p1, e1);
                v_cross(e0, e1, cr);
                const float inv = 1.0f / std::sqrt(v_dot(cr, cr));
                if (std::isfinite(inv) && inv != 0.0f) {
                    for (int c = 0; c < 3; ++c) tr.n[c] = inv * cr[c];
                } else {
                    for (int c = 0; c < 3; ++c) tr.n[c] = 0.0f;
                    tr.flags |= PBRS_TRI_DEGENERATE;
                }
                for (int c = 0; c < 3; ++c)
                    if (!(std::fabs(tr.p0[c]) < 1e18f && std::fabs(tr.p1[c]) < 1e18f && std::fabs(tr.p2[c]) < 1e18f)) tr.flags |= PBRS_TRI_HUGE;
            }
            TriShadeRec &sr = tri_shade[t0 + q];
            const uint32_t v3[3] = {i, j, k};
            float *nd[3] = {sr.n0, sr.n1, sr.n2}, *ud[3] = {sr.uv0, sr.uv1, sr.uv2};
            for (int a = 0; a < 3; ++a) {
                for (int c = 0; c < 3; ++c) nd[a][c] = m.N[3 * v3[a] + c];
                for (int c = 0; c < 2; ++c) ud[a][c] = m.UV[2 * v3[a] + c];
            }
            sr.pad = 0.0f;
        }
        for (uint32_t last : m.leaf_last) tris[t0 + last].flags |= PBRS_TRI_LAST_IN_LEAF;
        // parent links of this mesh's nodes and leaves (for the exact re-test of a stacked child)
        f.blas_node_parent.resize(blas_nodes.size(), 0u);
        f.blas_leaf_parent.resize(tris.size(), 0u);
        for (size_t p = 0; p < m.nodes.size(); ++p)
            for (uint32_t side = 0; side < 2; ++side) {
                const uint32_t link = m.nodes[p].child[side], tag = (uint32_t)p | (side << 31);
                if (link & PBRS_LEAF_LINK) f.blas_leaf_parent[t0 + (link & PBRS_LEAF_FIRST_MASK)] = tag;
                else f.blas_node_parent[r.node_base + link] = tag;
            }
    }
    f.tlas_node_parent.assign(s.tlas_nodes.size(), 0u);
    f.tlas_leaf_parent.assign(s.instances.size(), 0u);
    for (size_t p = 0; p < s.tlas_nodes.size(); ++p)
        for (uint32_t side = 0; side < 2; ++side) {
            const uint32_t link = s.tlas_nodes[p].child[side], tag = (uint32_t)p | (side << 31);
            if (link & PBRS_LEAF_LINK) f.tlas_leaf_parent[link & PBRS_LEAF_FIRST_MASK] = tag;
            else f.tlas_node_parent[link] = tag;
        }

    // ---- instances ----
    trav.assign(s.instances.size(), InstTravRec{});
    shade.assign(s.instances.size(), InstShadeRec{});
    for (size_t i = 0; i < s.instances.size(); ++i) {
        const HostInstance &in = s.instances[i];
        std::memset(&trav[i], 0, sizeof trav[i]);
        std::memset(&shade[i], 0, sizeof shade[i]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) { trav[i].inv[r][c] = in.inv[c][r]; shade[i].fwd[r][c] = in.fwd[c][r]; }
        trav[i].shape_kind = s.shapes[in.shape].kind;
        trav[i].shape_index = s.shapes[in.shape].index;
        trav[i].identity = in.identity ? 1u : 0u;
        shade[i].material = (uint32_t)in.material;
        {
            const MaterialRec &mr = s.materials[in.material];
            const int ids[4] = {mr.tex_kd, mr.tex_ks, mr.tex_kr, mr.tex_kt};
            shade[i].uses_uv = 0u;
            for (int id : ids)
                if (id >= 0 && id < (int)s.textures.size() && s.textures[id].rec.kind == PBRS_TEX_IMAGE) shade[i].uses_uv = 1u;
        }
        switch (s.materials[in.material].kind) {
        case PBRS_MTL_LAMBERTIAN: case PBRS_MTL_SUBSTRATE: shade[i].cls = PBRS_CLS_LAMBERT; break;
        case PBRS_MTL_METAL: case PBRS_MTL_GLOSSY: shade[i].cls = PBRS_CLS_MICROFACET; break;
        case PBRS_MTL_MIRROR: case PBRS_MTL_DIELECTRIC: shade[i].cls = PBRS_CLS_SPECULAR; break;
        case PBRS_MTL_DIFFUSE_LIGHT: shade[i].cls = PBRS_CLS_EMISSIVE; break;
        default: shade[i].cls = PBRS_CLS_MULTI; break;
        }
    }

    // ---- textures ----
    auto add_texture = [&](const HostTexture &t) {
        TextureRec r = t.rec;
        if (r.kind == PBRS_TEX_IMAGE) { r.texel_base = (uint32_t)texels.size(); texels.insert(texels.end(), t.texels.begin(), t.texels.end()); }
        if (r.kind == PBRS_TEX_PERLIN) {
            r.perlin_base = (uint32_t)(perlin_vec.size() / 768);
            perlin_vec.insert(perlin_vec.end(), t.perlin_vec.begin(), t.perlin_vec.end());
            perlin_perm.insert(perlin_perm.end(), t.perlin_perm.begin(), t.perlin_perm.end());
        }
        return r;
    };
    for (const HostTexture &t : s.textures) textures.push_back(add_texture(t));
    std::memset(&f.env_image, 0, sizeof f.env_image);
    if (s.env_kind == PBRS_ENV_KIND_IMAGE) f.env_image = add_texture(s.env_image);

}

// Fills every non-pointer field of the DeviceScene.
void fill_scene_constants(const SceneImpl &s, const FlatScene &f, DeviceScene &ds) {
    ds.n_delta = (uint32_t)s.delta_lights.size();
    ds.n_area = (uint32_t)s.area_lights.size();
    // Scene::has_env_light, scene/src/lib.rs:96-103
    ds.has_env = (s.env_kind != PBRS_ENV_KIND_CONSTANT || !(s.env_color[0] <= 0.0f && s.env_color[1] <= 0.0f && s.env_color[2] <= 0.0f)) ? 1u : 0u;
    ds.env_kind = s.env_kind;
    ds.env_fn = s.env_fn;
    for (int k = 0; k < 3; ++k) { ds.env_color[k] = s.env_color[k]; ds.env_scale[k] = s.env_scale[k]; }
    ds.env_image = f.env_image;
    ds.cam = s.cam;
    for (int k = 0; k < 3; ++k) { ds.tlas_min[k] = s.tlas_box.mn[k]; ds.tlas_max[k] = s.tlas_box.mx[k]; }
    ds.tlas_root_is_leaf = s.tlas_root_is_leaf ? 1u : 0u;
    ds.n_instances = (uint32_t)s.instances.size();
    // 8 is the optimum when rays spend their time inside a few big meshes (C4); with thousands of
    // mesh instances the leaf phase is mostly instance entries and a larger vote pays (C5 +3 %)
    ds.leaf_vote = (!s.meshes.empty() && s.instances.size() >= PBRS_MANY_INSTANCES) ? 12u : 8u;
    if (const char *e = std::getenv("PBRS_LEAF_VOTE_CLOSEST")) ds.leaf_vote = (uint32_t)std::atoi(e);  // development knob
    ds.has_mesh = s.meshes.empty() ? 0u : 1u;
    ds.coop_closest = f.tris.size() >= 1024 ? 1u : 0u;  // a Cornell box (34 triangles) loses 8 % of its extend time to the bookkeeping
    if (const char *e = std::getenv("PBRS_COOP_CLOSEST")) ds.coop_closest = (uint32_t)std::atoi(e);  // development knob
    // The split shade kernels (k_surface + k_scatter) were a gain on big meshes (-11 % on the 1 M-triangle
    // terrain) only while the one-piece kernels kept their Interaction, lobes and scene in local memory;
    // since those live in registers / the constant bank the one-piece kernels win everywhere (C4 shade
    // -13 %, C5 -23 %, C3 -20 % against the split: profiles/r2_exp_shade_inline_gridconstant.log), so the
    // split is off unless asked for.
    ds.shade_split = 0u;
    if (const char *e = std::getenv("PBRS_SHADE_SPLIT")) ds.shade_split = (uint32_t)std::atoi(e);
    ds.cls_mask = 1u << PBRS_CLS_MISS;
    for (const InstShadeRec &r : f.shade) ds.cls_mask |= 1u << r.cls;
    ds.has_ext = s.simples.empty() ? 0u : 1u;
    for (const HostMesh &m : s.meshes)
        if (!m.balls.empty()) ds.has_ext = 1u;
    for (const TriRec &t : f.tris)
        if (t.flags & PBRS_TRI_CHECK_SHADING) { ds.has_ext = 1u; break; }
}

}  // namespace pbrs
