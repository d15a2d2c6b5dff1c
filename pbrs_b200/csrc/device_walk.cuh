// The traversal kernels' walker: closest-hit and any-hit walks of the two-level BVH as ONE
// resumable state machine per lane.
//
// Why a state machine: in SIMT the naive nesting (TLAS loop -> instance -> BLAS loop -> leaf)
// lets a lane that is deep inside a mesh run alone while its 31 neighbours wait at the TLAS
// level (ncu on the first version: 7-11 of 32 lanes active, profiles/r1_v0_*).  Here every
// lane, whatever level it is on, meets the others in the same two phases of `step()`:
//     phase 1  until the lane stands at a leaf: expand inner nodes (fetch the 64-byte record, test
//              both child boxes, choose / push) and unwind the stack   (the hot loop, TLAS and BLAS alike)
//     phase 2  one leaf: a triangle run, a sphere, or an instance entry
// and the kernels (kernels.cu) refill finished lanes from the ray queue between steps.
//
// Exactness (DESIGN.md "Traversal"): the visit order, the extent each box/primitive is tested
// against and every tie-break equal the reference's walks (tlas/src/bvh.rs:77-113,
// shape/src/blas.rs:422-495, incl. quirks Q14/Q17).  The box test is the reference's slab test
// (geometry/src/bvh.rs:84-99, true divisions); a reciprocal-multiply pre-test decides the clear
// cases and hands everything within a few ulps of the boundary to the exact divisions, so the
// pass/fail outcome of every box is bit-identical to the reference's.
#pragma once
#include "device_geom.cuh"
#include "device_simple.cuh"

namespace pbrs {

#define PBRS_NONE 0xFFFFFFFFu
#define PBRS_TAG_COMBINE 0x40000000u  // TLAS closest: [lv] left value waiting for the right subtree's
#define PBRS_TAG_EXIT 0x20000000u     // closest: boundary between the TLAS entries and a mesh walk's
#define PBRS_WALK_STACK 160

// relative margin of the pre-test: the product (mn - o) * fl(1/d) is within 1.5 * 2^-23 of the
// quotient fl((mn - o) / d); 8 * 2^-23 leaves a 5x safety factor
#define PBRS_BOX_MARGIN 9.5367431640625e-7f
#define PBRS_BOX_TINY 1e-30f

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

// One child box against the ray.  `fast` = the ray's direction has no zero / non-finite
// reciprocal, so the products below are finite or overflow to inf (never NaN).
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
// need_overlap: the caller also wants `overlap` of the right child (TLAS closest-hit only).
struct PairTest {
    BoxTest l, r;
};
PB_DEV PairTest test_pair(f4 q0, f4 q1, f4 q2, vec3 o, vec3 d, vec3 rd, bool fast, float t_max, bool need_overlap) {
    PairTest p;
    if (fast) {
        float ax = (q0.x - o.x) * rd.x, bx = (q0.w - o.x) * rd.x, cx = (q1.z - o.x) * rd.x, dx = (q2.y - o.x) * rd.x;
        float ay = (q0.y - o.y) * rd.y, by = (q1.x - o.y) * rd.y, cy = (q1.w - o.y) * rd.y, dy = (q2.z - o.y) * rd.y;
        float az = (q0.z - o.z) * rd.z, bz = (q1.y - o.z) * rd.z, cz = (q2.x - o.z) * rd.z, dz = (q2.w - o.z) * rd.z;
        float ltl = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), 0.0f);
        float lme = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        float rtl = fmaxf(fmaxf(fmaxf(fminf(cx, dx), fminf(cy, dy)), fminf(cz, dz)), 0.0f);
        float rme = fminf(fminf(fmaxf(cx, dx), fmaxf(cy, dy)), fmaxf(cz, dz));
        // margins (explicit fma: this is the approximate side, contraction is harmless here)
        float lm = fmaf(PBRS_BOX_MARGIN, ltl, PBRS_BOX_TINY), ln = fmaf(PBRS_BOX_MARGIN, fabsf(lme), PBRS_BOX_TINY);
        float rm = fmaf(PBRS_BOX_MARGIN, rtl, PBRS_BOX_TINY), rn = fmaf(PBRS_BOX_MARGIN, fabsf(rme), PBRS_BOX_TINY);
        // pass = tl <= min(me, t_max): surely yes / surely no
        bool l_yes = ltl + lm <= fminf(lme - ln, t_max), l_no = ltl - lm > fminf(lme + ln, t_max);
        bool r_yes = rtl + rm <= fminf(rme - rn, t_max), r_no = rtl - rm > fminf(rme + rn, t_max);
        bool sure = (l_yes || l_no) && (r_yes || r_no);  // false whenever a value is inf / NaN
        bool r_ov_yes = rtl + rm <= rme - rn;
        if (need_overlap) sure = sure && (r_ov_yes || rtl - rm > rme + rn);
        if (sure) {
            p.l.pass = l_yes; p.l.overlap = l_yes; p.l.tl = ltl;
            p.r.pass = r_yes; p.r.overlap = r_ov_yes; p.r.tl = rtl;
            return p;
        }
    }
    p.l = box_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, d, t_max);
    p.r = box_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, d, t_max);
    return p;
}
// Re-test of a stacked child against the extent of the moment (it overlapped when pushed):
// pass iff t_low <= t_max.  Returns 1 pass, 0 fail, -1 too close to call with an approximate tl.
PB_DEV int retest(float tl, float t_max) {
    float m = PBRS_BOX_MARGIN * tl + PBRS_BOX_TINY;
    if (tl + m <= t_max) return 1;
    if (tl - m > t_max) return 0;
    return -1;
}

// EXT = the scene holds quads / cuboids / disks / isolated triangles, a sphere BLAS, or a triangle
// whose hit the shading interpolation can reject (PBRS_TRI_CHECK_SHADING): DeviceScene::has_ext.
// Without them the leaf code is the sphere + plain triangle code alone (3.5 % faster on C4).
template <bool ANY, bool COUNT, bool EXT = true>
struct Walk {
    // ray in the current space (world on the TLAS level, object inside a mesh instance)
    vec3 o, d, rd;
    float t_max;
    bool fast;
    // (while a lane is inside a mesh instance its world ray, reciprocal direction and extent wait
    // on the stack under the walk's entries -- see enter_mesh / leave_mesh -- not in registers)
    uint32_t next;  // ref to visit (PBRS_LEAF_BIT = leaf), PBRS_NONE = unwind
    uint32_t lvl;   // 0 = TLAS, 1 = inside a mesh instance
    int sp;
    bool done;
    // closest: running winner ("smallest t, right-most on ties"); any: occluded
    Hit best;
    bool occluded;
    // the mesh instance being walked
    float l_best_t;
    uint32_t l_best_tri, cur_inst;
    uint32_t node_base, tri_base, mesh_index;  // of that mesh
    float ret;  // TLAS closest: value of the subtree that just completed
    // the stack lives in arrays owned by the caller (so that the scalars above stay in registers)
    uint32_t *st_ref;
    float *st_tl;
    uint32_t *st_par;  // parent node (| side in bit 31) for the rare exact re-test

    PB_DEV Walk(uint32_t *ref, float *tl, uint32_t *par) : st_ref(ref), st_tl(tl), st_par(par) {}

    PB_DEV void set_space(vec3 no, vec3 nd, float nt) {
        o = no; d = nd; t_max = nt;
        rd = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        fast = is_fin(rd.x) && is_fin(rd.y) && is_fin(rd.z) && !any_nan(o);
    }
    // Eleven words of world-ray state parked on the stack for the duration of a mesh walk.
    PB_DEV void park(uint32_t a, uint32_t b, uint32_t c, Diag &dg) {
        if (ANY) { push(a, 0.0f, 0u, dg); push(b, 0.0f, 0u, dg); push(c, 0.0f, 0u, dg); }
        else push(a, u2f(b), c, dg);
    }
    PB_DEV void unpark(uint32_t &a, uint32_t &b, uint32_t &c) {
        if (ANY) { c = st_ref[sp - 1]; b = st_ref[sp - 2]; a = st_ref[sp - 3]; sp -= 3; }
        else { --sp; a = st_ref[sp]; b = f2u(st_tl[sp]); c = st_par[sp]; }
    }
    PB_DEV void save_world(Diag &dg) {
        park(f2u(o.x), f2u(o.y), f2u(o.z), dg);
        park(f2u(d.x), f2u(d.y), f2u(d.z), dg);
        park(f2u(rd.x), f2u(rd.y), f2u(rd.z), dg);
        park(f2u(t_max), fast ? 1u : 0u, cur_inst, dg);
        if (!ANY) park(f2u(best.t), best.inst, best.tri, dg);
    }
    PB_DEV void restore_world() {
        uint32_t a, b, c;
        if (!ANY) { unpark(a, b, c); best.t = u2f(a); best.inst = b; best.tri = c; }
        unpark(a, b, c); t_max = u2f(a); fast = b != 0u; cur_inst = c;
        unpark(a, b, c); rd = mk(u2f(a), u2f(b), u2f(c));
        unpark(a, b, c); d = mk(u2f(a), u2f(b), u2f(c));
        unpark(a, b, c); o = mk(u2f(a), u2f(b), u2f(c));
    }
    PB_DEV void push(uint32_t ref, float tl, uint32_t par, Diag &dg) {
        if (sp < PBRS_WALK_STACK) {
            st_ref[sp] = ref;
            if (!ANY) { st_tl[sp] = tl; st_par[sp] = par; }  // any-hit entries are bare refs
            ++sp;
        } else {
            flag(dg, P_STACK);
        }
    }
    PB_DEV const NodeRec *node_ptr(const DeviceScene &sc, uint32_t idx) const {
        return lvl ? sc.blas_nodes + node_base + idx : sc.tlas_nodes + idx;
    }
    // exact pass of child `side` of node `par` (the rare re-test)
    PB_DEV bool exact_child(const DeviceScene &sc, uint32_t par, float extent) const {
        const char *b = reinterpret_cast<const char *>(node_ptr(sc, par & 0x7FFFFFFFu));
        f4 q0 = ld16(b), q1 = ld16(b + 16), q2 = ld16(b + 32);
        if (par & 0x80000000u) return box_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, d, extent).pass;
        return box_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, d, extent).pass;
    }

    // ---- start: the root box with the ray's own extent (tlas/src/bvh.rs:78,106) ----
    PB_DEV void begin(const DeviceScene &sc, const Ray &ray) {
        set_space(ray.o, ray.d, ray.t_max);
        lvl = 0u; sp = 0; done = false; occluded = false; ret = PB_INF;
        best.t = PB_INF; best.inst = PBRS_NONE; best.tri = PBRS_NONE;
        BoxTest b = test_box(sc.tlas_min[0], sc.tlas_min[1], sc.tlas_min[2], sc.tlas_max[0], sc.tlas_max[1], sc.tlas_max[2], o, d, rd, fast, t_max);
        if (!b.pass) { done = true; next = PBRS_NONE; return; }
        next = sc.tlas_root_is_leaf ? PBRS_LEAF_BIT : 0u;
    }

    PB_DEV bool at_leaf() const { return !done && next != PBRS_NONE && (next & PBRS_LEAF_BIT); }
    // phase 1: expansions and stack unwinds until the lane stands at a leaf (or is done)
    PB_DEV bool advancing() const { return !done && !(next != PBRS_NONE && (next & PBRS_LEAF_BIT)); }
    PB_DEV void advance(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        if (next != PBRS_NONE) expand(sc, dg, tc);
        if (!done && next == PBRS_NONE) unwind(sc, dg);  // dead end: pop right away, same iteration
    }

    // ---- phase 1: expand the inner node `next` ----
    PB_DEV void expand(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        if (COUNT) tc.nodes++;
        const uint32_t self = next;
        const char *b = reinterpret_cast<const char *>(node_ptr(sc, self));
        f4 q0 = ld16(b), q1 = ld16(b + 16), q2 = ld16(b + 32), q3 = ld16(b + 48);
        const PairTest pt = test_pair(q0, q1, q2, o, d, rd, fast, t_max, !ANY && lvl == 0u);
        const BoxTest &L = pt.l, &R = pt.r;
        uint32_t meta = f2u(q3.z);
        uint32_t lref = f2u(q3.x) | ((meta & PBRS_NODE_LEFT_LEAF) ? PBRS_LEAF_BIT : 0u);
        uint32_t rref = f2u(q3.y) | ((meta & PBRS_NODE_RIGHT_LEAF) ? PBRS_LEAF_BIT : 0u);
        if (ANY) {
            // `left || right`, depth first (tlas/src/bvh.rs:105-113, blas.rs:478-495); the extent
            // never changes, so a pass is final.  (Visiting the nearer child first was measured:
            // 4 % MORE nodes on the C4 scene, so the reference's order stays.)
            if (L.pass) {
                if (R.pass) push(rref, 0.0f, 0u, dg);
                next = lref;
            } else {
                next = R.pass ? rref : PBRS_NONE;
            }
            return;
        }
        if (lvl) {
            // shape/src/blas.rs:456-466: near = left iff dir[axis] > 0; the far child is pushed
            // first and re-tested against the extent of the moment when it is popped
            bool left_near = comp(d, (int)(meta & 3u)) > 0.0f;
            const BoxTest &N = left_near ? L : R, &F = left_near ? R : L;
            if (F.pass) push(left_near ? rref : lref, F.tl, self | (left_near ? 0x80000000u : 0u), dg);
            next = N.pass ? (left_near ? lref : rref) : PBRS_NONE;
            return;
        }
        // tlas/src/bvh.rs:83-100: left, then right against whatever extent the left subtree leaves
        if (R.overlap) push(rref, R.tl, self | 0x80000000u, dg);
        if (L.pass) next = lref;
        else { next = PBRS_NONE; ret = PB_INF; }
    }

    // ---- phase 2a: a leaf ----
    PB_DEV void leaf(const DeviceScene &sc, Diag &dg, TravCount &tc) {
        const uint32_t first = next & PBRS_LEAF_FIRST_MASK;
        next = PBRS_NONE;
        if (lvl) {
            // a run of triangles (shape/src/blas.rs:447-454): all see the extent of the pop
            Ray ray; ray.o = o; ray.d = d; ray.t_max = t_max;
            uint32_t s = tri_base + first;
            while (true) {
                TriVerts tv = load_tri(sc.tris + s);
                if (EXT && (tv.flags & PBRS_TRI_SPHERE)) {
                    // IsoBlas<Sphere>: the leaf closure is the sphere's own test (blas.rs:267-274)
                    if (COUNT) tc.spheres++;
                    float t;
                    if (ball_test(tv.p0, tv.p1.x, ray, ANY, t, dg)) {
                        if (ANY) { occluded = true; done = true; return; }
                        if (t < l_best_t) { l_best_t = t; l_best_tri = s; }
                    }
                } else if (ANY) {
                    if (COUNT) tc.tris++;
                    if (tri_occludes(tv.p0, tv.p1, tv.p2, ray, dg)) { occluded = true; done = true; return; }
                } else {
                    if (COUNT) tc.tris++;
                    float t;
                    bool hit;
                    if (EXT && (tv.flags & PBRS_TRI_CHECK_SHADING)) {
#if PBRS_TRISHADE_CALL
                        hit = mesh_tri_shade_t(sc, s, tv, ray, t, dg);
#else
                        MeshHit mh;
                        hit = mesh_tri_shade(sc, s, tv, ray, mh, dg);
                        t = mh.t;
#endif
                    } else {
                        TriHit h;
                        hit = tri_intersect(tv.p0, tv.p1, tv.p2, ray, h, dg);
                        t = h.t;
                    }
                    if (hit && t < l_best_t) { l_best_t = t; l_best_tri = s; }
                }
                if (tv.flags & PBRS_TRI_LAST_IN_LEAF) break;
                ++s;
            }
            if (!ANY) t_max = l_best_t;  // blas.rs:468
            return;
        }
        // an instance (tlas/src/instance.rs:50-72): the ray goes to object space
        if (COUNT) tc.insts++;
        Ray wr; wr.o = o; wr.d = d; wr.t_max = t_max;
        uint32_t kind, index;
        Ray obj = to_object(sc.inst_trav + first, wr, kind, index);
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
                if (hit) { occluded = true; done = true; }
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
        cur_inst = first;
        save_world(dg);
        set_space(obj.o, obj.d, obj.t_max);
        BoxTest rb = test_box(mesh.bmin[0], mesh.bmin[1], mesh.bmin[2], mesh.bmax[0], mesh.bmax[1], mesh.bmax[2], o, d, rd, fast, t_max);
        if (!rb.pass) {
            restore_world();
            if (!ANY) ret = PB_INF;
            return;
        }
        node_base = mesh.node_base; tri_base = mesh.tri_base; mesh_index = index;
        l_best_t = PB_INF; l_best_tri = PBRS_NONE;
        lvl = 1u;
        push(PBRS_TAG_EXIT, 0.0f, 0u, dg);
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
            if (sp == 0) { done = true; return; }
            --sp;
            const uint32_t ref = st_ref[sp];
            if (lvl) {
                if (ref == PBRS_TAG_EXIT) {
                    // the mesh walk is over: back to the world ray
                    lvl = 0u;
                    restore_world();
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
                int rt = retest(st_tl[sp], t_max);
                if (rt < 0) rt = exact_child(sc, st_par[sp], t_max) ? 1 : 0;
                if (rt) { next = ref; return; }
                continue;
            }
            if (ANY) { next = ref; return; }
            if (ref == PBRS_TAG_COMBINE) {
                float lv = st_tl[sp];
                ret = (lv < ret) ? lv : ret;  // pick(l, r).t
                continue;
            }
            // a right child whose left sibling subtree just completed with value `ret`
            const float r_tl = st_tl[sp];
            const uint32_t par = st_par[sp];
            if (ret < PB_INF) {
                t_max = ret;  // ray.set_extent(isect.ray_t), tlas/src/bvh.rs:85-87
                st_ref[sp] = PBRS_TAG_COMBINE; st_tl[sp] = ret; ++sp;
            }
            int rt = retest(r_tl, t_max);
            if (rt < 0) rt = exact_child(sc, par, t_max) ? 1 : 0;
            if (rt) { next = ref; return; }
            ret = PB_INF;
        }
    }

    // the whole walk, sequentially (host-sim and single-ray callers)
    PB_DEV void run(const DeviceScene &sc, const Ray &ray, Diag &dg, TravCount &tc) {
        begin(sc, ray);
        while (!done) {
            while (advancing()) advance(sc, dg, tc);
            if (at_leaf()) leaf(sc, dg, tc);
        }
    }
};

}  // namespace pbrs
