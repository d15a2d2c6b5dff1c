// The traversal kernels' walker: closest-hit and any-hit walks of the two-level BVH as ONE
// resumable state machine per lane.
//
// Why a state machine: in SIMT the naive nesting (TLAS loop -> instance -> BLAS loop -> leaf)
// lets a lane that is deep inside a mesh run alone while its 31 neighbours wait at the TLAS
// level (ncu on the first version: 7-11 of 32 lanes active, profiles/r1_v0_*).  Here every
// lane, whatever level it is on, meets the others in the same two phases:
//     phase 1  until the lane stands at a leaf: expand inner nodes (fetch the 64-byte record, test
//              both child boxes, choose / push) and unwind the stack   (the hot loop, TLAS and BLAS alike)
//     phase 2  one leaf: a triangle run, a sphere, or an instance entry
// and the kernels (kernels.cu) refill finished lanes from the ray queue between steps.
//
// The whole state of a lane is the register `next`:
//     >= 0           an inner node to expand          (link as loaded from the parent record)
//     PBRS_NONE  -1  dead end: unwind the stack
//     PBRS_DONE  -2  the walk is over (also: the lane is idle)
//     <  -2          a leaf link (PBRS_LEAF_BIT | run length << 28 | first primitive)
// so "advancing" and "at a leaf" are one signed compare each, and the warp's vote is two ballots.
//
// Exactness (DESIGN.md "Traversal"): the visit order, the extent each box/primitive is tested
// against and every tie-break equal the reference's walks (tlas/src/bvh.rs:77-113,
// shape/src/blas.rs:422-495, incl. quirks Q14/Q17).  The box test is the reference's slab test
// (geometry/src/bvh.rs:84-99, true divisions); a reciprocal-multiply pre-test decides the clear
// cases and hands everything within a few ulps of the boundary to the exact divisions, so the
// pass/fail outcome of every box is bit-identical to the reference's.
#pragma once
#include <type_traits>

#include "device_geom.cuh"
#include "device_simple.cuh"

namespace pbrs {

#define PBRS_NONE 0xFFFFFFFFu
#define PBRS_DONE 0xFFFFFFFEu
#define PBRS_TAG_COMBINE 0x40000000u  // TLAS closest: [lv] left value waiting for the right subtree's
#define PBRS_TAG_EXIT 0x20000000u     // boundary between the TLAS entries and a mesh walk's
// Stack entries alive at once: <= one per TLAS level (a pending right child or a COMBINE), the
// EXIT tag, one far child per BLAS level; commit rejects scenes whose depths do not fit.
#define PBRS_WALK_STACK PBRS_WALK_STACK_ENTRIES
#define PBRS_WALK_PARK 16  // words of world-ray state parked beside the stack during a mesh walk

PB_DEV bool ref_advancing(uint32_t n) { return (int32_t)n >= -1; }
PB_DEV bool ref_at_leaf(uint32_t n) { return (int32_t)n < -2; }

// Relative margin of the pre-test.  With t = (mn - o) * fl(1/d) the product is within 2.5 ulp of the
// reference's fl(fl(mn - o) / d); 16 * 2^-24 leaves a 3x safety factor.  PBRS_BOX_FMA evaluates
// t = fma(mn, fl(1/d), fl(-o * fl(1/d))) instead: one instruction per slab instead of two, at the
// price of an ABSOLUTE error of one ulp of |o / d|, which the per-ray constant `c0` covers.
#define PBRS_BOX_MARGIN 9.5367431640625e-7f
#define PBRS_BOX_TINY 1e-30f
#ifndef PBRS_BOX_FMA
#define PBRS_BOX_FMA 0
#endif
#ifndef PBRS_PREFETCH_CHILDREN
#define PBRS_PREFETCH_CHILDREN 0
#endif

struct BoxTest {
    bool pass;     // t_low <= min(min_el, t_max): exactly the reference's outcome
    bool overlap;  // !(t_low > min_el): could pass under a larger extent
    float tl;      // t_low, approximate (within the margin) or exact
};

// The reference's slab test with its true divisions; out of line, it is the rare path.
PB_CALL BoxTest box_exact(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, vec3 o, vec3 d, float t_max) {
    Ray ray; ray.o = o; ray.d = d; ray.t_max = t_max;
    float tl, me;
    slab(mnx, mny, mnz, mxx, mxy, mxz, ray, tl, me);
    BoxTest r;
    r.pass = box_pass(tl, me, t_max);
    r.overlap = !(tl > me);
    r.tl = tl;
    return r;
}

// One box against the ray (root boxes: once per walk / instance entry).  `fast` = the ray's
// direction has no zero / non-finite reciprocal, so the products below are finite or overflow to
// inf (never NaN).
PB_DEV BoxTest test_box(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, vec3 o, vec3 d, vec3 rd, bool fast, float t_max) {
    BoxTest r;
    if (fast) {
        float ax = (mnx - o.x) * rd.x, bx = (mxx - o.x) * rd.x;
        float ay = (mny - o.y) * rd.y, by = (mxy - o.y) * rd.y;
        float az = (mnz - o.z) * rd.z, bz = (mxz - o.z) * rd.z;
        float tl = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), 0.0f);
        float me = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        float m_tl = PBRS_BOX_MARGIN * tl + PBRS_BOX_TINY;            // tl >= 0
        float m_me = PBRS_BOX_MARGIN * fabsf(me) + PBRS_BOX_TINY;
        bool ov_yes = tl + m_tl <= me - m_me, ov_no = tl - m_tl > me + m_me;
        bool t_yes = tl + m_tl <= t_max, t_no = tl - m_tl > t_max;
        if ((ov_yes || ov_no) && (t_yes || t_no)) {  // false whenever tl or me is inf / NaN
            r.pass = ov_yes && t_yes;
            r.overlap = ov_yes;
            r.tl = tl;
            return r;
        }
    }
    return box_exact(mnx, mny, mnz, mxx, mxy, mxz, o, d, t_max);
}

// Both child boxes of a node at once: one pre-test, one shared branch to the exact path.
// A child's decision "t_low <= min(min_el, t_max)" is taken from the approximate values when the
// difference of the two sides exceeds the error bound M * (|hi| + t_low) + c0 of that difference
// (M = PBRS_BOX_MARGIN, c0 = the ray's absolute term), else from the exact divisions.
// need_overlap: the caller also wants `overlap` of the right child (TLAS closest-hit only).
struct PairTest {
    bool lp, rp, rov;  // left / right pass, right overlap
    float ltl, rtl;    // entry distances (approximate within the margin, or exact)
};
PB_DEV PairTest test_pair(f4 q0, f4 q1, f4 q2, vec3 o, vec3 d, vec3 rd, vec3 nord, float c0, bool fast, float t_max, bool need_overlap) {
    PairTest p;
    if (fast) {
#if PBRS_BOX_FMA
        float ax = fmaf(q0.x, rd.x, nord.x), bx = fmaf(q0.w, rd.x, nord.x), cx = fmaf(q1.z, rd.x, nord.x), dx = fmaf(q2.y, rd.x, nord.x);
        float ay = fmaf(q0.y, rd.y, nord.y), by = fmaf(q1.x, rd.y, nord.y), cy = fmaf(q1.w, rd.y, nord.y), dy = fmaf(q2.z, rd.y, nord.y);
        float az = fmaf(q0.z, rd.z, nord.z), bz = fmaf(q1.y, rd.z, nord.z), cz = fmaf(q2.x, rd.z, nord.z), dz = fmaf(q2.w, rd.z, nord.z);
#else
        float ax = (q0.x - o.x) * rd.x, bx = (q0.w - o.x) * rd.x, cx = (q1.z - o.x) * rd.x, dx = (q2.y - o.x) * rd.x;
        float ay = (q0.y - o.y) * rd.y, by = (q1.x - o.y) * rd.y, cy = (q1.w - o.y) * rd.y, dy = (q2.z - o.y) * rd.y;
        float az = (q0.z - o.z) * rd.z, bz = (q1.y - o.z) * rd.z, cz = (q2.x - o.z) * rd.z, dz = (q2.w - o.z) * rd.z;
#endif
        float ltl = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), 0.0f);
        float lme = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        float rtl = fmaxf(fmaxf(fmaxf(fminf(cx, dx), fminf(cy, dy)), fminf(cz, dz)), 0.0f);
        float rme = fminf(fminf(fmaxf(cx, dx), fmaxf(cy, dy)), fmaxf(cz, dz));
        // (explicit fma: this is the approximate side, contraction is harmless here)
        const float lhi = fminf(lme, t_max), rhi = fminf(rme, t_max);
        const float ldiff = lhi - ltl, rdiff = rhi - rtl;
        const float lthr = fmaf(fabsf(lhi) + ltl, PBRS_BOX_MARGIN, c0), rthr = fmaf(fabsf(rhi) + rtl, PBRS_BOX_MARGIN, c0);
        bool sure = fabsf(ldiff) > lthr && fabsf(rdiff) > rthr;  // false whenever a value is inf / NaN
        p.lp = ldiff >= 0.0f; p.rp = rdiff >= 0.0f; p.rov = p.rp;
        if (need_overlap) {
            const float odiff = rme - rtl;
            sure = sure && fabsf(odiff) > fmaf(fabsf(rme) + rtl, PBRS_BOX_MARGIN, c0);
            p.rov = odiff >= 0.0f;
        }
        if (sure) { p.ltl = ltl; p.rtl = rtl; return p; }
    }
    const BoxTest l = box_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, d, t_max);
    const BoxTest r = box_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, d, t_max);
    p.lp = l.pass; p.rp = r.pass; p.rov = r.overlap; p.ltl = l.tl; p.rtl = r.tl;
    return p;
}
// Re-test of a stacked child against the extent of the moment (it overlapped when pushed):
// pass iff t_low <= t_max.  Returns 1 pass, 0 fail, -1 too close to call with an approximate tl.
PB_DEV int retest(float tl, float t_max, float c0) {
    float m = fmaf(PBRS_BOX_MARGIN, tl, c0);
    if (tl + m <= t_max) return 1;
    if (tl - m > t_max) return 0;
    return -1;
}

// The walk's stack in a plain array (local memory in the traversal kernels; the shared-memory ring
// of kernels.cu has the same interface).  A closest-hit entry is one 8-byte (link, t_low) pair --
// one local load / store per pop / push instead of two: local-memory requests outnumber the
// global ones 2.7 : 1 in the closest-hit kernel (ncu r2) -- an any-hit entry a bare link; `park` is
// PBRS_WALK_PARK words beside the stack.
struct alignas(16) ParkVec {
    uint32_t x, y, z, w;
};
struct alignas(8) StackPair {
    uint32_t ref;
    float tl;
};
template <bool ANY, bool CHECK = true>
struct ArrayStack {
    using Entry = typename std::conditional<ANY, uint32_t, StackPair>::type;
    Entry *ent;
    uint32_t *park;
    int sp;
    PB_DEV ArrayStack(Entry *e, uint32_t *p) : ent(e), park(p), sp(0) {}
    PB_DEV void reset() { sp = 0; }
    PB_DEV bool empty() const { return sp == 0; }
    // CHECK = false: no overflow test.  pbrs_scene_commit bounds TLAS depth + BLAS depth + 2 by the stack
    // size (scene_host.cpp) and at most one entry per level plus the EXIT tag is alive at once, so the
    // test can never fire on a committed scene; it costs 1-3 % of the traversal kernels
    // (profiles/r2_exp_stack_check_votes.log).  The counting kernels, which the parity tests run on every
    // scene, and the host build keep it.
    PB_DEV void push(uint32_t r, float t, Diag &dg) {
        if (!CHECK || sp < PBRS_WALK_STACK) {
            if constexpr (ANY) ent[sp] = r;
            else { StackPair e; e.ref = r; e.tl = t; ent[sp] = e; }
            ++sp;
        } else {
            flag(dg, P_STACK);
        }
    }
    PB_DEV void pop(uint32_t &r, float &t) {
        --sp;
        if constexpr (ANY) { r = ent[sp]; t = 0.0f; }
        else { const StackPair e = ent[sp]; r = e.ref; t = e.tl; }
    }
    // the park words, four at a time (16-byte aligned: one local-memory request instead of four)
    PB_DEV void park_set4(int q, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        ParkVec v; v.x = a; v.y = b; v.z = c; v.w = d;
        reinterpret_cast<ParkVec *>(park)[q] = v;
    }
    PB_DEV u4 park_get4(int q) const {
        const ParkVec v = reinterpret_cast<const ParkVec *>(park)[q];
        u4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        return r;
    }
};

// EXT = the scene holds quads / cuboids / disks / isolated triangles, a sphere BLAS, or a triangle
// whose hit the shading interpolation can reject (PBRS_TRI_CHECK_SHADING): DeviceScene::has_ext.
// Without them the leaf code is the sphere + plain triangle code alone (3.5 % faster on C4).
template <bool ANY, bool COUNT, bool EXT, class STK>
struct Walk {
    // ray in the current space (world on the TLAS level, object inside a mesh instance)
    vec3 o, d, rd;
#if PBRS_BOX_FMA
    vec3 nord;  // fl(-o * rd)
    float c0;   // absolute error term of the pre-test: 4 * 2^-24 * max |o * rd| + tiny
#endif
    float t_max;
    // bits 0..2: d.x / d.y / d.z > 0 (near-child choice, blas.rs:456); bit 3: `fast`; bit 4: inside a mesh
    uint32_t bits;
    // (while a lane is inside a mesh instance its world ray, reciprocal direction and extent wait
    // in the park slots of the stack -- see enter / leave -- not in registers)
    uint32_t next;
    // closest: running winner ("smallest t, right-most on ties"); any: occluded
    Hit best;
    bool occluded;
    // the mesh instance being walked
    float l_best_t;
    uint32_t l_best_tri, cur_inst;
    const NodeRec *nodes;  // node array of the current level (TLAS, or the mesh's first node)
    uint32_t tri_base;     // of that mesh
    float ret;             // TLAS closest: value of the subtree that just completed
    STK st;

    PB_DEV explicit Walk(const STK &s) : st(s) { next = PBRS_DONE; }

    PB_DEV bool fast() const { return (bits & 8u) != 0u; }
    PB_DEV bool in_mesh() const { return (bits & 16u) != 0u; }
    PB_DEV bool done() const { return next == PBRS_DONE; }
    PB_DEV bool at_leaf() const { return ref_at_leaf(next); }
    PB_DEV bool advancing() const { return ref_advancing(next); }
    PB_DEV float abs_term() const {
#if PBRS_BOX_FMA
        return c0;
#else
        return PBRS_BOX_TINY;
#endif
    }

    PB_DEV void set_space(vec3 no, vec3 nd, float nt, uint32_t mesh_bit) {
        o = no; d = nd; t_max = nt;
        rd = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        const bool f = is_fin(rd.x) && is_fin(rd.y) && is_fin(rd.z) && !any_nan(o);
        bits = (d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u) | (f ? 8u : 0u) | mesh_bit;
#if PBRS_BOX_FMA
        set_fma_terms();
#endif
    }
#if PBRS_BOX_FMA
    PB_DEV void set_fma_terms() {
        nord = mk(-(o.x * rd.x), -(o.y * rd.y), -(o.z * rd.z));
        c0 = fmaf(fmaxf(fmaxf(fabsf(nord.x), fabsf(nord.y)), fabsf(nord.z)), 2.384185791015625e-7f, PBRS_BOX_TINY);
    }
#endif
    // World-ray state parked for the duration of a mesh walk (restored bit for bit; three IEEE
    // divisions for rd at 3 of 32 lanes cost more than three loads, ncu r2).
    PB_DEV void save_world() {
        st.park_set4(0, f2u(o.x), f2u(o.y), f2u(o.z), f2u(t_max));
        st.park_set4(1, f2u(d.x), f2u(d.y), f2u(d.z), bits);
        if (ANY) st.park_set4(2, f2u(rd.x), f2u(rd.y), f2u(rd.z), 0u);
        else {
            st.park_set4(2, f2u(rd.x), f2u(rd.y), f2u(rd.z), cur_inst);
            st.park_set4(3, f2u(best.t), best.inst, best.tri, 0u);
        }
    }
    PB_DEV void restore_world() {
        const u4 a = st.park_get4(0), b = st.park_get4(1), c = st.park_get4(2);
        o = mk(u2f(a.x), u2f(a.y), u2f(a.z)); t_max = u2f(a.w);
        d = mk(u2f(b.x), u2f(b.y), u2f(b.z)); bits = b.w;
        rd = mk(u2f(c.x), u2f(c.y), u2f(c.z));
#if PBRS_BOX_FMA
        set_fma_terms();
#endif
        if (!ANY) {
            const u4 e = st.park_get4(3);
            cur_inst = c.w;
            best.t = u2f(e.x); best.inst = e.y; best.tri = e.z;
        }
    }
    // exact pass of the stacked child `ref` (the rare re-test): its box is in its parent's record
    PB_DEV bool exact_child(const DeviceScene &sc, uint32_t ref, float extent) const {
        uint32_t par;
        const bool leaf = (ref & PBRS_LEAF_BIT) != 0u;
        if (in_mesh()) {
            const uint32_t node_base = (uint32_t)(nodes - sc.blas_nodes);
            par = leaf ? ld_u32(sc.blas_leaf_parent + tri_base + (ref & PBRS_LEAF_FIRST_MASK)) : ld_u32(sc.blas_node_parent + node_base + ref);
        } else {
            par = leaf ? ld_u32(sc.tlas_leaf_parent + (ref & PBRS_LEAF_FIRST_MASK)) : ld_u32(sc.tlas_node_parent + ref);
        }
        const char *b = reinterpret_cast<const char *>(nodes + (par & 0x7FFFFFFFu));
        f4 q0 = ld16(b), q1 = ld16(b + 16), q2 = ld16(b + 32);
        if (par & 0x80000000u) return box_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, d, extent).pass;
        return box_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, d, extent).pass;
    }

    // ---- start: the root box with the ray's own extent (tlas/src/bvh.rs:78,106) ----
    PB_DEV void begin(const DeviceScene &sc, const Ray &ray) {
        set_space(ray.o, ray.d, ray.t_max, 0u);
        st.reset();
        occluded = false; ret = PB_INF;
        best.t = PB_INF; best.inst = PBRS_NONE; best.tri = PBRS_NONE;
        nodes = sc.tlas_nodes;
        BoxTest b = test_box(sc.tlas_min[0], sc.tlas_min[1], sc.tlas_min[2], sc.tlas_max[0], sc.tlas_max[1], sc.tlas_max[2], o, d, rd, fast(), t_max);
        if (!b.pass) { next = PBRS_DONE; return; }
        next = sc.tlas_root_is_leaf ? PBRS_LEAF_BIT : 0u;
    }

    // phase 1: expansions and stack unwinds until the lane stands at a leaf (or is done)
    PB_DEV void advance(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        if (next != PBRS_NONE) expand(sc, dg, tc);
        if (next == PBRS_NONE) unwind(sc, dg);  // dead end: pop right away, same iteration
    }

    // ---- phase 1: expand the inner node `next` ----
    PB_DEV void expand(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        if (COUNT) tc.nodes++;
        const char *b = reinterpret_cast<const char *>(nodes + next);
        f4 q0, q1, q2, q3;
        ld32(b, q0, q1); ld32(b + 32, q2, q3);
#if PBRS_BOX_FMA
        const PairTest pt = test_pair(q0, q1, q2, o, d, rd, nord, c0, fast(), t_max, !ANY && !in_mesh());
#else
        const PairTest pt = test_pair(q0, q1, q2, o, d, rd, o, PBRS_BOX_TINY, fast(), t_max, !ANY && !in_mesh());
#endif
        const uint32_t lref = f2u(q3.x), rref = f2u(q3.y);  // leaf bit and run length are part of the link
#if PBRS_PREFETCH_CHILDREN && defined(__CUDA_ARCH__)
        // the right child's record while the boxes are being tested (the left one is the next record in preorder)
        if ((int32_t)rref >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + rref));
#endif
        if (ANY) {
            // `left || right`, depth first (tlas/src/bvh.rs:105-113, blas.rs:478-495); the extent
            // never changes, so a pass is final.  (Visiting the nearer child first was measured:
            // 4 % MORE nodes on the C4 scene, so the reference's order stays.)
            if (pt.lp) {
                if (pt.rp) st.push(rref, 0.0f, dg);
                next = lref;
            } else {
                next = pt.rp ? rref : PBRS_NONE;
            }
            return;
        }
        if (in_mesh()) {
            // shape/src/blas.rs:456-466: near = left iff dir[axis] > 0; the far child is pushed
            // first and re-tested against the extent of the moment when it is popped
            const bool left_near = ((bits >> (f2u(q3.z) & 3u)) & 1u) != 0u;
            const bool np = left_near ? pt.lp : pt.rp, fp = left_near ? pt.rp : pt.lp;
            // (near fails, far passes: the far child would be pushed and popped right back against
            // the unchanged extent -- the same outcome without the round trip through the stack)
            const uint32_t nref = left_near ? lref : rref, fref = left_near ? rref : lref;
            if (np) {
                if (fp) st.push(fref, left_near ? pt.rtl : pt.ltl, dg);
                next = nref;
            } else {
                next = fp ? fref : PBRS_NONE;
            }
            return;
        }
        // tlas/src/bvh.rs:83-100: left, then right against whatever extent the left subtree leaves
        if (pt.rov) st.push(rref, pt.rtl, dg);
        if (pt.lp) next = lref;
        else { next = PBRS_NONE; ret = PB_INF; }
    }

    // One leaf run of a mesh, from record `s` to the first one flagged LAST_IN_LEAF, against `ray`
    // (shape/src/blas.rs:447-454).  Closest-hit: the running (bt, btri) of the mesh walk is improved
    // by strictly smaller hits; any-hit: returns true at the first occluder.
    PB_DEV bool run_tris(const DeviceScene &sc, uint32_t s, const Ray &ray, float &bt, uint32_t &btri, Diag &dg, TravCount &tc) const {
        while (true) {
            TriVerts tv = load_tri<true>(sc.tris + s);
            if (EXT && (tv.flags & PBRS_TRI_SPHERE)) {
                // IsoBlas<Sphere>: the leaf closure is the sphere's own test (blas.rs:267-274)
                if (COUNT) tc.spheres++;
                float t;
                if (ball_test(tv.p0, tv.p1.x, ray, ANY, t, dg)) {
                    if (ANY) return true;
                    if (t < bt) { bt = t; btri = s; }
                }
            } else if (ANY) {
                if (COUNT) tc.tris++;
                if (mesh_tri_occludes(tv, ray, dg)) return true;
            } else {
                if (COUNT) tc.tris++;
                float t;
                bool hit;
                if (EXT && (tv.flags & PBRS_TRI_CHECK_SHADING)) hit = mesh_tri_shade_t(sc, s, tv, ray, t, dg);
                else hit = mesh_tri_hit_t(tv, ray, t, dg);
                if (hit && t < bt) { bt = t; btri = s; }
            }
            if (tv.flags & PBRS_TRI_LAST_IN_LEAF) return false;
            ++s;
        }
    }

    // ---- phase 2a: a leaf ----
    PB_DEV void leaf(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        const uint32_t first = next & PBRS_LEAF_FIRST_MASK;
        next = PBRS_NONE;
        if (in_mesh()) {
            // a run of triangles (shape/src/blas.rs:447-454): all see the extent of the pop
            Ray ray; ray.o = o; ray.d = d; ray.t_max = t_max;
            if (run_tris(sc, tri_base + first, ray, l_best_t, l_best_tri, dg, tc)) { occluded = true; next = PBRS_DONE; return; }
            if (!ANY) t_max = l_best_t;  // blas.rs:468
            return;
        }
        // an instance (tlas/src/instance.rs:50-72): the ray goes to object space
        if (COUNT) tc.insts++;
        Ray wr; wr.o = o; wr.d = d; wr.t_max = t_max;
        uint32_t kind, index;
        Ray obj = to_object<true>(sc.inst_trav + first, wr, kind, index);
        if (!(len2(obj.d) > (ANY ? 1e-6f : 1e-3f))) flag(dg, P_MISC);
        if (kind != PBRS_SHAPE_MESH) {
            float t = PB_INF;
            bool hit;
            if (kind == PBRS_SHAPE_SPHERE) {
                if (COUNT) tc.spheres++;
                f4 s = ld16(sc.spheres + index);
                bool far_root = false;
                hit = ANY ? sphere_occludes(mk(s.x, s.y, s.z), s.w, obj) : sphere_hit_t(mk(s.x, s.y, s.z), s.w, obj, t, far_root);
                if (!ANY && far_root) flag(dg, P_SPHERE_INSIDE);
            } else if (!EXT) {
                hit = false;  // unreachable: has_ext selects the EXT kernels
            } else {  // quad, cuboid, disk: shape/src/simple.rs
                if (COUNT && kind == PBRS_SHAPE_TRIANGLE) tc.tris++;
                hit = ANY ? simple_occludes(sc.simples + index, kind, obj, dg) : simple_hit_t(sc.simples + index, kind, obj, t, dg);
            }
            if (ANY) {
                if (hit) { occluded = true; next = PBRS_DONE; }
                return;
            }
            if (hit) {
                ret = t;
                if (t <= best.t) { best.t = t; best.inst = first; best.tri = 0u; }
            } else {
                ret = PB_INF;
            }
            return;
        }
        // a mesh: its root box sees the incoming extent (blas.rs:428 and the root's own pop, :441)
        const MeshHead mesh = load_mesh_head(sc.meshes + index);
        // (Testing a one-leaf mesh -- a Cornell wall: two triangles -- in place, without the mesh protocol,
        // was measured: the exact root-box test it needs and the EXT kernels it selects cost more than the
        // protocol saves; C1 extend +6 %, C5 +9 %: profiles/r2_exp_inplace_leaf_mesh_class_mask.log)
        cur_inst = first;
        save_world();
        set_space(obj.o, obj.d, obj.t_max, 16u);
        BoxTest rb = test_box(mesh.bmin[0], mesh.bmin[1], mesh.bmin[2], mesh.bmax[0], mesh.bmax[1], mesh.bmax[2], o, d, rd, fast(), t_max);
        if (!rb.pass) {
            restore_world();
            if (!ANY) ret = PB_INF;
            return;
        }
        nodes = sc.blas_nodes + mesh.node_base; tri_base = mesh.tri_base;
        l_best_t = PB_INF; l_best_tri = PBRS_NONE;
        st.push(PBRS_TAG_EXIT, 0.0f, dg);
        if (mesh.root_is_leaf) {
            next = PBRS_LEAF_BIT;  // its triangles see the incoming extent (the clone of `r`)
        } else {
            next = 0u;
            if (!ANY) t_max = PB_INF;  // Q17: after the root's pop the extent is the walk's own best
        }
    }

    // ---- phase 2b: unwind the stack until there is something to visit ----
    PB_DEV void unwind(const DeviceScene &sc, Diag &dg) {
        while (true) {
            if (st.empty()) { next = PBRS_DONE; return; }
            uint32_t ref;
            float e_tl;
            st.pop(ref, e_tl);
            if (in_mesh()) {
                if (ref == PBRS_TAG_EXIT) {
                    // the mesh walk is over: back to the world ray
                    restore_world();
                    nodes = sc.tlas_nodes;
                    if (ANY) continue;
                    if (l_best_t < PB_INF) {
                        ret = l_best_t;
                        if (l_best_t <= best.t) { best.t = l_best_t; best.inst = cur_inst; best.tri = l_best_tri; }
                    } else {
                        ret = PB_INF;
                    }
                    continue;
                }
                if (ANY) { next = ref; return; }
                int rt = retest(e_tl, t_max, abs_term());
                if (rt < 0) rt = exact_child(sc, ref, t_max) ? 1 : 0;
                if (rt) { next = ref; return; }
                continue;
            }
            if (ANY) { next = ref; return; }
            if (ref == PBRS_TAG_COMBINE) {
                ret = (e_tl < ret) ? e_tl : ret;  // pick(l, r).t
                continue;
            }
            // a right child whose left sibling subtree just completed with value `ret`
            if (ret < PB_INF) {
                t_max = ret;  // ray.set_extent(isect.ray_t), tlas/src/bvh.rs:85-87
                st.push(PBRS_TAG_COMBINE, ret, dg);
            }
            int rt = retest(e_tl, t_max, abs_term());
            if (rt < 0) rt = exact_child(sc, ref, t_max) ? 1 : 0;
            if (rt) { next = ref; return; }
            ret = PB_INF;
        }
    }

    // the whole walk, sequentially (host-sim and single-ray callers)
    PB_DEV void run(const DeviceScene &sc, const Ray &ray, Diag &dg, TravCount &tc) {
        begin(sc, ray);
        while (!done()) {
            while (advancing()) advance(sc, dg, tc);
            if (at_leaf()) leaf(sc, dg, tc);
        }
    }
};

}  // namespace pbrs
