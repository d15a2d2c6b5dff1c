// sm_100a kernels of the wavefront and the host loop that enqueues them.
//
// Every kernel is persistent: the grid is sized once to fill the 148 SMs (occupancy x SM count)
// and each warp strides over its queue in 32-path groups, reading the queue length from HBM, so
// the host never synchronises between stages -- a whole frame is one stream of launches.
// Queue appends are compacted per warp: one ballot, one atomicAdd by the leader, a popc prefix.
// Tensor cores are unused on purpose: there is no dense contraction anywhere on this path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "device_stages.cuh"
#include "render.h"

namespace pbrs {

namespace {

constexpr int kThreads = 128;
// Kernel parameters that are handed on by reference to out-of-line device functions (the scene, the
// path buffers): __grid_constant__ lets their address be taken in place, in the constant bank; without
// it every such kernel starts by copying the struct into its local-memory stack frame and the callees
// read the scene through local loads.
#ifndef PBRS_GRID_CONSTANT
#define PBRS_GRID_CONSTANT 1
#endif
#if PBRS_GRID_CONSTANT
#define PBRS_GC const __grid_constant__
#else
#define PBRS_GC
#endif
#ifndef PBRS_LANES
#define PBRS_LANES 3  // batches in flight on their own streams: 3 beats 2 by 1.0 % on the full C4 frame, 4 equals 3 (profiles/r2_exp_lanes.log)
#endif
constexpr int kMaxLanes = 4;
static_assert(PBRS_LANES >= 1 && PBRS_LANES <= kMaxLanes, "PBRS_LANES");

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// one atomic per warp; returns this lane's slot (valid only where pred)
__device__ __forceinline__ uint32_t warp_push(uint32_t *count, bool pred) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0u) return 0u;
    int leader = __ffs(m) - 1;
    uint32_t base = 0u;
    if ((int)lane_id() == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane_id()) - 1u));
}
// two pushes whose atomics are in flight together (one L2 round trip instead of two)
__device__ __forceinline__ void warp_push2(uint32_t *c1, bool p1, uint32_t *c2, bool p2, uint32_t &s1, uint32_t &s2) {
    const unsigned m1 = __ballot_sync(0xFFFFFFFFu, p1), m2 = __ballot_sync(0xFFFFFFFFu, p2);
    uint32_t b1 = 0u, b2 = 0u;
    if (lane_id() == 0u) {
        if (m1) b1 = atomicAdd(c1, (uint32_t)__popc(m1));
        if (m2) b2 = atomicAdd(c2, (uint32_t)__popc(m2));
    }
    b1 = __shfl_sync(0xFFFFFFFFu, b1, 0);
    b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
    const unsigned below = (1u << lane_id()) - 1u;
    s1 = b1 + (uint32_t)__popc(m1 & below);
    s2 = b2 + (uint32_t)__popc(m2 & below);
}
__device__ __forceinline__ void warp_add_stat(unsigned long long *dst, uint32_t v) {
    uint32_t s = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0u && s != 0u) atomicAdd(dst, (unsigned long long)s);
}
__device__ __forceinline__ void flush_diag(unsigned long long *stats, const Diag &dg) {
    uint32_t any = __reduce_or_sync(0xFFFFFFFFu, dg.panics);
    if (any == 0u) return;
    for (int k = 0; k < 16; ++k)
        if (any & (1u << k)) {
            uint32_t c = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, (dg.panics >> k) & 1u));
            if (lane_id() == 0u) atomicAdd(stats + kStatPanic0 + k, (unsigned long long)c);
        }
}
__device__ __forceinline__ void flush_count(unsigned long long *stats, const TravCount &tc, int which) {
    unsigned long long *d = stats + kStatTrav + 4 * which;
    warp_add_stat(d + 0, tc.nodes);
    warp_add_stat(d + 1, tc.tris);
    warp_add_stat(d + 2, tc.spheres);
    warp_add_stat(d + 3, tc.insts);
}

// warp-uniform strided loop over [0, n): `body(idx, active)` runs with all 32 lanes converged
#define PBRS_WARP_LOOP(n_expr, idx, active)                                                           \
    const uint32_t _n = (n_expr);                                                                     \
    const uint32_t _warps = (gridDim.x * blockDim.x) >> 5;                                            \
    for (uint32_t _b = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; _b < _n; _b += _warps * 32u) \
        if (uint32_t idx = _b + lane_id(); true)                                                      \
            if (bool active = idx < _n; true)

// A warp takes 128 consecutive paths per round and appends their live ones with ONE atomicAdd: the
// queue counter is a single address, on which the L2 serialises every atomic (about half of this
// kernel's time with one atomic per 32 paths; the queue's order is irrelevant to the film).
__global__ void __launch_bounds__(kThreads) k_generate(DeviceScene sc, PathBuffers pb, FrameParams fp, BatchParams bp, uint32_t *out_count) {
    const uint32_t n = bp.n_paths, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 128u; b < n; b += warps * 128u) {
        unsigned m[4];
        uint32_t total = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = b + 32u * k + lane_id();
            m[k] = __ballot_sync(0xFFFFFFFFu, j < n && stage_generate(sc, pb, fp, bp, j));
            total += (uint32_t)__popc(m[k]);
        }
        if (total == 0u) continue;
        uint32_t base = 0u;
        if (lane_id() == 0u) base = atomicAdd(out_count, total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((m[k] >> lane_id()) & 1u) pb.queue[0][base + (uint32_t)__popc(m[k] & ((1u << lane_id()) - 1u))] = b + 32u * k + lane_id();
            base += (uint32_t)__popc(m[k]);
        }
    }
}

// Persistent traversal kernel (closest hit: ANY = false, shadow entries: ANY = true).
// Every lane owns one Walk; the warp alternates phase 1 (inner-node expansions, all levels) and
// phase 2 (one leaf or unwind) and, whenever enough lanes have run dry, hands them the next rays
// of the warp's chunk of the queue (one atomicAdd on the launch's work cursor per chunk).
#ifndef PBRS_REFILL_IDLE_LANES
#define PBRS_REFILL_IDLE_LANES 8  // a refill costs no atomic (the warp owns a chunk of the queue): 8 beats 12 and 20 (profiles/r2_exp_coop_closest_refill_vote.log)
#endif
#ifndef PBRS_REFILL_IDLE_LANES_ANY
#define PBRS_REFILL_IDLE_LANES_ANY PBRS_REFILL_IDLE_LANES
#endif
// refill as soon as this many lanes are idle (closest-hit / any-hit walks tuned separately)
template <bool ANY> constexpr int kRefillIdleLanes = ANY ? PBRS_REFILL_IDLE_LANES_ANY : PBRS_REFILL_IDLE_LANES;
#ifndef PBRS_REFILL_CHUNK
#define PBRS_REFILL_CHUNK 64  // rays a warp draws from the queue per atomicAdd
#endif
#ifndef PBRS_LEAF_VOTE_ANY
#define PBRS_LEAF_VOTE_ANY 12  // any-hit leaves are cheap to wait for: C5 shadow -7 %, C3/C4 unchanged (profiles/r1_exp_anyhit_vote.log)
#endif
#ifndef PBRS_NODE_STEPS
#define PBRS_NODE_STEPS 3  // node steps between two votes of the warp (3 beats 2 by 1-4 % of the extend time, 1 loses: profiles/r2_exp_*.log)
#endif

// `cnt` = this stage's counter block (PBRS_CNT_*).
#ifndef PBRS_TRACE_BLOCKS_PER_SM
#define PBRS_TRACE_BLOCKS_PER_SM 8
#endif
#ifndef PBRS_COOP_ANY
#define PBRS_COOP_ANY 1
#endif
#ifndef PBRS_COOP_FILL_SERIAL
#define PBRS_COOP_FILL_SERIAL 0  // 1: every owner writes all its (lane, triangle) slots one by one (the first version: 5 % of the kernel's instructions at 4 of 32 lanes)
#endif
// Deferred unwinds: a lane whose walk ran dry waits for the end of the round and pops together with
// the others instead of alone, right away (a pop runs at 3-5 of 32 lanes and is two dependent
// local-memory accesses).  Any-hit walks gain (shadow -5 % on C4, -6 % on C5 with 2 node steps per
// round); closest-hit walks do not (C4 equal, C3 / C1 +6-10 %): profiles/r2_exp_ld256_packed_stack_defer.log
#ifndef PBRS_DEFER_UNWIND_ANY
#define PBRS_DEFER_UNWIND_ANY 1
#endif
#ifndef PBRS_DEFER_UNWIND_CLOSEST
#define PBRS_DEFER_UNWIND_CLOSEST 0
#endif
#ifndef PBRS_NODE_STEPS_ANY
#define PBRS_NODE_STEPS_ANY 2
#endif
template <bool ANY> constexpr bool kDeferUnwind = ANY ? (PBRS_DEFER_UNWIND_ANY != 0) : (PBRS_DEFER_UNWIND_CLOSEST != 0);
template <bool ANY> constexpr int kNodeSteps = ANY ? PBRS_NODE_STEPS_ANY : PBRS_NODE_STEPS;
#ifndef PBRS_COOP_CLOSEST
#define PBRS_COOP_CLOSEST 1  // with the stack in local memory: extend -4.5 % on C4 and C5; scenes of a few triangles lose 8 % and switch it off (DeviceScene::coop_closest)
#endif

// The walk's stack on the device: a ring of the TOP entries per lane in shared memory (entry s of
// lane l at ring[s * kThreads + l]: one bank per lane, so a push or pop is one conflict-free
// wavefront whatever the lanes' stack depths are -- in local memory, lanes at different depths
// touch different 128-byte lines), spilling its oldest entry to local memory when it is full and
// reading back from there only after the ring has run empty.  Entries [lo, sp) are in the ring,
// [0, lo) in local memory.  The parked world-ray state of a mesh walk sits in local memory at
// fixed slots (all lanes the same address offset: coalesced).
// Measured (profiles/r2_exp_walker_v2_variants.log, r2_exp_coop_closest_refill_vote.log,
// r2_exp_stack_prefetch_steps.log): the ring costs more than it saves -- extend +3.4 % on C4, +4.7 % on
// C5, +6 % on C3 against the same walk with its stack in local memory, and a ring of 4 entries is worse
// than one of 8.  Local-memory stack traffic is 2.6 % of the instructions and hits L1 88 % of the time
// (ncu before the change, profiles/r2_ncu_trace_c4.md); the ring adds its full / empty bookkeeping to every push and
// pop.  It therefore stays a build option (-DPBRS_SMEM_STACK=1) and the default stack is local memory.
#ifndef PBRS_SMEM_STACK
#define PBRS_SMEM_STACK 0
#endif
#ifndef PBRS_SMEM_STACK_CLOSEST
#define PBRS_SMEM_STACK_CLOSEST 8   // (link, t_low) pairs: 8 KB per block of 128 threads
#endif
#ifndef PBRS_SMEM_STACK_ANY
#define PBRS_SMEM_STACK_ANY 16      // bare links: 8 KB per block
#endif
template <bool ANY>
struct SmemStack {
    static constexpr int S = ANY ? PBRS_SMEM_STACK_ANY : PBRS_SMEM_STACK_CLOSEST;
    static_assert((S & (S - 1)) == 0, "ring size must be a power of two");
    using Entry = typename std::conditional<ANY, uint32_t, uint2>::type;
    static constexpr uint32_t kSlotBytes = (uint32_t)sizeof(Entry) * kThreads;  // distance between two slots of a lane
    uint32_t ring;        // shared-space byte address of this lane's slot 0
    uint32_t *spill_ref;  // local memory, PBRS_WALK_STACK + PBRS_WALK_PARK words
    float *spill_tl;
    int sp, lo;
    __device__ __forceinline__ SmemStack(Entry *r, uint32_t *sr, float *stl)
        : ring((uint32_t)__cvta_generic_to_shared(r)), spill_ref(sr), spill_tl(stl), sp(0), lo(0) {}
    __device__ __forceinline__ uint32_t slot(int i) const { return ring + ((uint32_t)i & (uint32_t)(S - 1)) * kSlotBytes; }
    __device__ __forceinline__ void store(int i, uint32_t r, float t) {
        if constexpr (ANY) asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot(i)), "r"(r) : "memory");
        else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot(i)), "r"(r), "r"(__float_as_uint(t)) : "memory");
    }
    __device__ __forceinline__ void load(int i, uint32_t &r, float &t) const {
        if constexpr (ANY) {
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(slot(i)) : "memory");
            t = 0.0f;
        } else {
            uint32_t tb;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r), "=r"(tb) : "r"(slot(i)) : "memory");
            t = __uint_as_float(tb);
        }
    }
    __device__ __forceinline__ void reset() { sp = 0; lo = 0; }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
    __device__ __forceinline__ void push(uint32_t r, float t, Diag &dg) {
        if (sp - lo == S) {  // full: the oldest entry goes to local memory
            if (lo < PBRS_WALK_STACK) {
                uint32_t er; float et;
                load(lo, er, et);
                spill_ref[lo] = er;
                if constexpr (!ANY) spill_tl[lo] = et;
            } else {
                flag(dg, P_STACK);
            }
            ++lo;
        }
        store(sp, r, t);
        ++sp;
    }
    __device__ __forceinline__ void pop(uint32_t &r, float &t) {
        --sp;
        if (sp >= lo) {
            load(sp, r, t);
        } else {  // the ring ran empty: the entry is in local memory
            const int k = sp < PBRS_WALK_STACK ? sp : PBRS_WALK_STACK - 1;
            r = spill_ref[k];
            t = ANY ? 0.0f : spill_tl[k];
            lo = sp;
        }
    }
    __device__ __forceinline__ void park_set4(int q, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        uint32_t *p = spill_ref + PBRS_WALK_STACK + 4 * q;
        p[0] = a; p[1] = b; p[2] = c; p[3] = d;
    }
    __device__ __forceinline__ u4 park_get4(int q) const {
        const uint32_t *p = spill_ref + PBRS_WALK_STACK + 4 * q;
        u4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3];
        return v;
    }
};

template <bool ANY, bool COUNT, bool EXT>
__global__ void __launch_bounds__(kThreads, PBRS_TRACE_BLOCKS_PER_SM) k_trace(PBRS_GC DeviceScene sc, PBRS_GC PathBuffers pb, const uint32_t *queue, uint32_t *cnt) {
    const uint32_t *count = cnt + (ANY ? PBRS_CNT_SHADOW : PBRS_CNT_EXTEND);
    uint32_t *cursor = cnt + (ANY ? PBRS_CNT_SHADOW_CURSOR : PBRS_CNT_EXTEND_CURSOR);
    Diag dg; dg.panics = 0u;
    TravCount tc; tc.nodes = tc.tris = tc.spheres = tc.insts = 0u;
    const uint32_t n = *count;
    // local memory: the spill area of the stack + the park slots (one array, dynamically indexed, so
    // that the statically indexed park slots are not promoted to registers)
#if PBRS_SMEM_STACK
    uint32_t st_ref[PBRS_WALK_STACK + PBRS_WALK_PARK];
    float st_tl[ANY ? 1 : PBRS_WALK_STACK];
    using Stk = SmemStack<ANY>;
    __shared__ typename Stk::Entry ring_mem[Stk::S * kThreads];
    Walk<ANY, COUNT, EXT, Stk> w(Stk(ring_mem + threadIdx.x, st_ref, st_tl));
#else
    using Stk = ArrayStack<ANY, COUNT>;
    alignas(16) typename Stk::Entry st_mem[PBRS_WALK_STACK + PBRS_WALK_PARK];  // (closest-hit: the park words take half of their 16 entries)
    Walk<ANY, COUNT, EXT, Stk> w(Stk(st_mem, reinterpret_cast<uint32_t *>(st_mem + PBRS_WALK_STACK)));
#endif
    // the warp's chunk of the queue: [chunk[0], chunk[1]) is still to be handed out
    __shared__ uint32_t chunk_mem[kThreads / 32][2];
    uint32_t *chunk = chunk_mem[threadIdx.x >> 5];
    if (lane_id() == 0u) { chunk[0] = 0u; chunk[1] = 0u; }
    __syncwarp();
    bool busy = false, exhausted = false;
    uint32_t j = 0u, vis = 0u;
    int which = 0;
    const int vote = ANY ? PBRS_LEAF_VOTE_ANY : (int)sc.leaf_vote;
    while (true) {
        // ---- refill idle lanes ----
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (!exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= kRefillIdleLanes<ANY>)) {
            uint32_t c_next = chunk[0], c_end = chunk[1];
            if (c_next == c_end) {
                uint32_t base = 0u;
                if (lane_id() == 0u) base = atomicAdd(cursor, (uint32_t)PBRS_REFILL_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base >= n) { exhausted = true; c_next = c_end = 0u; }
                else { c_next = base; c_end = min(base + (uint32_t)PBRS_REFILL_CHUNK, n); }
            }
            const uint32_t take = min((uint32_t)__popc(idle), c_end - c_next);
            if (!busy) {
                const uint32_t r = (uint32_t)__popc(idle & ((1u << lane_id()) - 1u));
                if (r < take) {
                    j = queue[c_next + r];
                    busy = true;
                    if (ANY) {
                        vis = 0u;
                        Ray ray;
                        which = 0;
                        if (!shadow_ray(pb, j, 0, ray)) { which = 1; shadow_ray(pb, j, 1, ray); }
                        w.begin(sc, ray);
                    } else {
                        w.begin(sc, load_ray(pb, j));
                    }
                }
            }
            __syncwarp();
            if (lane_id() == 0u) { chunk[0] = c_next + take; chunk[1] = c_end; }
            __syncwarp();
        } else if (idle == 0xFFFFFFFFu) {
            break;
        }
        // ---- phase 1: expansions / unwinds until enough lanes stand at a leaf (or all are done) ----
#pragma unroll 1
        while (true) {
            const unsigned m_adv = __ballot_sync(0xFFFFFFFFu, w.advancing());
            const unsigned m_leaf = __ballot_sync(0xFFFFFFFFu, w.at_leaf());
            if (m_adv == 0u || __popc(m_leaf) >= vote) break;
            // (a real loop: three unrolled copies of the step push the kernel past 50 KB of SASS and the
            // extend time up 10 %, profiles/r2_exp_shade_phased_and_code_size.log)
            if constexpr (kDeferUnwind<ANY>) {
                // dead ends are not unwound one lane at a time: every lane that ran dry (left a leaf, or
                // both children failed) pops at the top of the next round, together with the others
                if (w.next == PBRS_NONE) w.unwind(sc, dg);
#pragma unroll 1
                for (int k = 0; k < kNodeSteps<ANY>; ++k)
                    if ((int32_t)w.next >= 0) w.expand(sc, dg, tc);
            } else {
#pragma unroll 1
                for (int k = 0; k < kNodeSteps<ANY>; ++k)
                    if (w.advancing()) w.advance(sc, dg, tc);
            }
        }
        __syncwarp();
#if PBRS_COOP_ANY
        // ---- phase 2, any-hit: the triangle runs of all lanes at a BLAS leaf, spread over the warp ----
        // A lane at a leaf has 1..6 triangles to test against its ray and typically 5-6 of the 32
        // lanes are there, so testing them lane by lane, triangle after triangle, issues the test
        // code ~3 times for ~6 lanes.  Here every (lane, triangle) pair gets a lane of its own:
        // the owner's ray comes over by shuffle, the result goes back as a ballot.  Any-hit only:
        // the outcome is an OR, no order to preserve.  (COUNT kernels keep the sequential walk,
        // whose counters stop at the first occluder like the reference's.)
        if (ANY && !COUNT && sc.has_mesh) {
            __shared__ uint8_t coop_slots[kThreads];
            uint8_t *slot = coop_slots + (threadIdx.x & ~31u);
            bool mine = w.at_leaf() && w.in_mesh() && ((w.next >> PBRS_LEAF_COUNT_SHIFT) & 7u) != 0u;
            unsigned owners = __ballot_sync(0xFFFFFFFFu, mine);
            while (owners) {
                const uint32_t c = mine ? ((w.next >> PBRS_LEAF_COUNT_SHIFT) & 7u) : 0u;
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if ((int)lane_id() >= d) incl += v;
                }
                const uint32_t excl = incl - c;
                const bool in_pass = mine && incl <= 32u;  // the first owner always fits (c <= 6)
#if PBRS_COOP_FILL_SERIAL
                if (in_pass)
                    for (uint32_t k = 0; k < c; ++k) slot[excl + k] = (uint8_t)(lane_id() | (k << 5));
                const uint32_t total = __reduce_max_sync(0xFFFFFFFFu, in_pass ? incl : 0u);
                __syncwarp();
                const bool work = lane_id() < total;
                const uint32_t e = work ? slot[lane_id()] : 0u;
#else
                // slot p = first pair of a run: its owner leaves its lane there (one store); a working
                // lane finds the start of its run as the highest start bit at or below its own index
                const unsigned starts = __reduce_or_sync(0xFFFFFFFFu, in_pass ? (1u << excl) : 0u);
                if (in_pass) slot[excl] = (uint8_t)lane_id();
                const uint32_t total = __reduce_max_sync(0xFFFFFFFFu, in_pass ? incl : 0u);
                __syncwarp();
                const bool work = lane_id() < total;
                const uint32_t p0 = 31u - (uint32_t)__clz((int)(starts & (0xFFFFFFFFu >> (31u - lane_id()))));
                const uint32_t e = work ? ((uint32_t)slot[p0 & 31u] | ((lane_id() - p0) << 5)) : 0u;
#endif
                const int src = (int)(e & 31u);
                Ray r;
                r.o.x = __shfl_sync(0xFFFFFFFFu, w.o.x, src); r.o.y = __shfl_sync(0xFFFFFFFFu, w.o.y, src); r.o.z = __shfl_sync(0xFFFFFFFFu, w.o.z, src);
                r.d.x = __shfl_sync(0xFFFFFFFFu, w.d.x, src); r.d.y = __shfl_sync(0xFFFFFFFFu, w.d.y, src); r.d.z = __shfl_sync(0xFFFFFFFFu, w.d.z, src);
                r.t_max = __shfl_sync(0xFFFFFFFFu, w.t_max, src);
                const uint32_t first = __shfl_sync(0xFFFFFFFFu, w.tri_base + (w.next & PBRS_LEAF_FIRST_MASK), src);
                bool hit = false;
                if (work) {
                    const TriVerts tv = load_tri<true>(sc.tris + first + (e >> 5));
                    if (EXT && (tv.flags & PBRS_TRI_SPHERE)) { float t; hit = ball_test(tv.p0, tv.p1.x, r, true, t, dg); }
                    else hit = mesh_tri_occludes(tv, r, dg);
                }
                const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
                if (in_pass) {
                    // the leaf is consumed: occluded, or phase 1 unwinds from here
                    if (hits & (((1u << c) - 1u) << excl)) { w.occluded = true; w.next = PBRS_DONE; }
                    else w.next = PBRS_NONE;
                    mine = false;
                }
                __syncwarp();
                owners = __ballot_sync(0xFFFFFFFFu, mine);
            }
        }
#endif
#if PBRS_COOP_CLOSEST
        // ---- phase 2, closest-hit: the same spreading of (lane, triangle) pairs over the warp ----
        // The leaf's triangles all see the extent of the pop and the winner is the smallest t, the
        // FIRST of the run on ties (`t < best` in run order, shape/src/blas.rs:447-454): a segmented
        // min over the run's lanes that only takes a later candidate when it is strictly smaller,
        // then one compare against the walk's best.  Incoherent rays stand at their leaves 3-7
        // lanes at a time (ncu, bounce 1 of C4): lane by lane the triangle code ran at 4-5 of 32 lanes.
        if (!ANY && sc.coop_closest) {
            __shared__ uint16_t coop_slots2[kThreads];
            uint16_t *slot = coop_slots2 + (threadIdx.x & ~31u);
            bool mine = w.at_leaf() && w.in_mesh() && ((w.next >> PBRS_LEAF_COUNT_SHIFT) & 7u) != 0u;
            unsigned owners = __ballot_sync(0xFFFFFFFFu, mine);
            while (owners) {
                const uint32_t c = mine ? ((w.next >> PBRS_LEAF_COUNT_SHIFT) & 7u) : 0u;
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if ((int)lane_id() >= d) incl += v;
                }
                const uint32_t excl = incl - c;
                const bool in_pass = mine && incl <= 32u;  // the first owner always fits (c <= 6)
#if PBRS_COOP_FILL_SERIAL
                if (in_pass)
                    for (uint32_t k = 0; k < c; ++k) slot[excl + k] = (uint16_t)(lane_id() | (k << 5) | (c << 8));
                const uint32_t total = __reduce_max_sync(0xFFFFFFFFu, in_pass ? incl : 0u);
                __syncwarp();
                const bool work = lane_id() < total;
                const uint32_t e = work ? slot[lane_id()] : 0u;
#else
                const unsigned starts = __reduce_or_sync(0xFFFFFFFFu, in_pass ? (1u << excl) : 0u);
                if (in_pass) slot[excl] = (uint16_t)(lane_id() | (c << 8));
                const uint32_t total = __reduce_max_sync(0xFFFFFFFFu, in_pass ? incl : 0u);
                __syncwarp();
                const bool work = lane_id() < total;
                const uint32_t p0 = 31u - (uint32_t)__clz((int)(starts & (0xFFFFFFFFu >> (31u - lane_id()))));
                const uint32_t e0 = work ? (uint32_t)slot[p0 & 31u] : 0u;
                const uint32_t e = work ? ((e0 & 31u) | ((lane_id() - p0) << 5) | (e0 & 0xFF00u)) : 0u;
#endif
                const int src = (int)(e & 31u);
                const uint32_t k = (e >> 5) & 7u, run = e >> 8;
                Ray r;
                r.o.x = __shfl_sync(0xFFFFFFFFu, w.o.x, src); r.o.y = __shfl_sync(0xFFFFFFFFu, w.o.y, src); r.o.z = __shfl_sync(0xFFFFFFFFu, w.o.z, src);
                r.d.x = __shfl_sync(0xFFFFFFFFu, w.d.x, src); r.d.y = __shfl_sync(0xFFFFFFFFu, w.d.y, src); r.d.z = __shfl_sync(0xFFFFFFFFu, w.d.z, src);
                r.t_max = __shfl_sync(0xFFFFFFFFu, w.t_max, src);
                const uint32_t first = __shfl_sync(0xFFFFFFFFu, w.tri_base + (w.next & PBRS_LEAF_FIRST_MASK), src);
                float val = PB_INF;
                if (work) {
                    const uint32_t s = first + k;
                    const TriVerts tv = load_tri<true>(sc.tris + s);
                    float t;
                    bool hit;
                    if (EXT && (tv.flags & PBRS_TRI_SPHERE)) { if (COUNT) tc.spheres++; hit = ball_test(tv.p0, tv.p1.x, r, false, t, dg); }
                    else {
                        if (COUNT) tc.tris++;
                        if (EXT && (tv.flags & PBRS_TRI_CHECK_SHADING)) hit = mesh_tri_shade_t(sc, s, tv, r, t, dg);
                        else hit = mesh_tri_hit_t(tv, r, t, dg);
                    }
                    if (hit) val = t;
                }
                uint32_t arg = k;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {  // runs are <= 6 long
                    const float ov = __shfl_down_sync(0xFFFFFFFFu, val, d);
                    const uint32_t oa = __shfl_down_sync(0xFFFFFFFFu, arg, d);
                    if (work && k + (uint32_t)d < run && ov < val) { val = ov; arg = oa; }  // strictly smaller only: the earlier triangle keeps a tie
                }
                const float bv = __shfl_sync(0xFFFFFFFFu, val, (int)(excl & 31u));
                const uint32_t ba = __shfl_sync(0xFFFFFFFFu, arg, (int)(excl & 31u));
                if (in_pass) {
                    if (bv < w.l_best_t) { w.l_best_t = bv; w.l_best_tri = w.tri_base + (w.next & PBRS_LEAF_FIRST_MASK) + ba; }
                    w.t_max = w.l_best_t;  // blas.rs:468
                    w.next = PBRS_NONE;
                    mine = false;
                }
                __syncwarp();
                owners = __ballot_sync(0xFFFFFFFFu, mine);
            }
        }
#endif
        // ---- phase 2: one leaf ----
        if (w.at_leaf()) w.leaf(sc, dg, tc);
        if (ANY) {
            if (busy && w.done()) {
                if (!w.occluded) vis |= 1u << which;
                Ray r;
                if (which == 0 && shadow_ray(pb, j, 1, r)) {
                    which = 1;
                    w.begin(sc, r);
                } else {
                    shadow_finish(pb, j, vis);
                    busy = false;
                }
            }
        } else {
            // a finished walk: hit record out, path into the shade queue of its material class
            // (lanes of one class share one atomicAdd)
            const bool fin = busy && w.done();
            const unsigned fmask = __ballot_sync(0xFFFFFFFFu, fin);
            if (fin) {
                store_hit(pb, j, w.best);
                const uint32_t cls = hit_class(sc, w.best);
                const unsigned peers = __match_any_sync(fmask, cls);
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0u;
                if ((int)lane_id() == leader) base = atomicAdd(cnt + PBRS_CNT_CLS + cls, (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                pb.cls_queue[cls][base + (uint32_t)__popc(peers & ((1u << lane_id()) - 1u))] = j;
                busy = false;
            }
        }
    }
    __syncwarp();
    flush_diag(pb.stats, dg);
    if (COUNT) flush_count(pb.stats, tc, ANY ? 1 : 0);
}

// One shade kernel per material class CLS (its queue was filled by the extend kernel) and
// integrator: only the code of that class's lobes is compiled in, and the lanes of a warp run
// the same material code.  `cnt` / `next_cnt` = counter blocks of this stage / the next one.
#ifndef PBRS_SHADE_BLOCKS_PER_SM
#define PBRS_SHADE_BLOCKS_PER_SM 4
#endif
// PBRS_SHADE_SPLIT (path integrator): the classes with a large body -- Lambert, microfacet, multi-
// lobe -- are shaded by two kernels instead of one: k_surface rebuilds the hit (one kernel for all
// three classes: the code does not depend on the material) and leaves it in the path's 64-byte
// surface record, k_scatter<class> does lobes, light sampling and BSDF sampling from the record.
// Each half is well under the instruction-cache footprint and the register count of the one-piece
// kernel (186 KB of SASS, 128 registers, 4.4 stall cycles per issue waiting for instructions:
// profiles/r1_final3_ncu_shade_c4.md); the price is 128 bytes of state traffic per path and bounce.
// Measured (profiles/r2_exp_shade_split.log): shade -11 % on C4, where rebuilding a hit means the
// triangle's shading record, the normal / uv interpolation and two tangent frames; +12 % on C5 and
// +18 % on C3, where the hit is a sphere or a small instanced mesh and the record is pure overhead.
// Commit therefore switches it on per scene (DeviceScene::shade_split: >= 100 000 triangles).
#ifndef PBRS_SHADE_SPLIT
#define PBRS_SHADE_SPLIT 1
#endif
#ifndef PBRS_SURFACE_BLOCKS_PER_SM
#define PBRS_SURFACE_BLOCKS_PER_SM 6
#endif
#ifndef PBRS_SCATTER_BLOCKS_PER_SM
#define PBRS_SCATTER_BLOCKS_PER_SM 4
#endif
template <int CLS, int INTEGRATOR>
__global__ void __launch_bounds__(kThreads, PBRS_SHADE_BLOCKS_PER_SM) k_shade(PBRS_GC DeviceScene sc, PBRS_GC PathBuffers pb, PBRS_GC FrameParams fp, PBRS_GC BatchParams bp, uint32_t *cnt,
                                                    uint32_t *next_queue, uint32_t *next_cnt, int bounce) {
    Diag dg; dg.panics = 0u;
    uint32_t rays = 0u;
    const uint32_t *queue = pb.cls_queue[CLS];
    uint32_t *next_count = next_cnt + PBRS_CNT_EXTEND, *shadow_count = cnt + PBRS_CNT_SHADOW;
    PBRS_WARP_LOOP(cnt[PBRS_CNT_CLS + CLS], i, active) {
        ShadeOut so; so.next = false; so.shadow_rays = 0;
        uint32_t j = 0u;
        if (active) {
            j = queue[i];
            so = INTEGRATOR == PBRS_INTEGRATOR_PATH ? stage_shade_path<CLS>(sc, pb, fp, bp, j, bounce, dg)
                                                    : stage_shade_direct<CLS>(sc, pb, fp, bp, j, bounce, dg);
        }
        uint32_t s1, s2;
        warp_push2(next_count, so.next, shadow_count, so.shadow_rays > 0, s1, s2);
        if (so.next) next_queue[s1] = j;
        if (so.shadow_rays > 0) pb.shadow_queue[s2] = j;
        rays += (uint32_t)so.shadow_rays;
    }
    __syncwarp();
    warp_add_stat(pb.stats + kStatShadowRays, rays);
    flush_diag(pb.stats, dg);
}
// the three heavy classes' queues, one after the other
__global__ void __launch_bounds__(kThreads, PBRS_SURFACE_BLOCKS_PER_SM) k_surface(PBRS_GC DeviceScene sc, PBRS_GC PathBuffers pb, const uint32_t *cnt, int bounce) {
    Diag dg; dg.panics = 0u;
    const int classes[3] = {PBRS_CLS_LAMBERT, PBRS_CLS_MICROFACET, PBRS_CLS_MULTI};
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        const uint32_t *queue = pb.cls_queue[classes[c]];
        PBRS_WARP_LOOP(cnt[PBRS_CNT_CLS + classes[c]], i, active) {
            if (active) stage_shade_surface(sc, pb, queue[i], bounce, dg);
        }
    }
    __syncwarp();
    flush_diag(pb.stats, dg);
}
template <int CLS>
__global__ void __launch_bounds__(kThreads, PBRS_SCATTER_BLOCKS_PER_SM) k_scatter(PBRS_GC DeviceScene sc, PBRS_GC PathBuffers pb, PBRS_GC FrameParams fp, PBRS_GC BatchParams bp, uint32_t *cnt,
                                                                                  uint32_t *next_queue, uint32_t *next_cnt, int bounce) {
    Diag dg; dg.panics = 0u;
    uint32_t rays = 0u;
    const uint32_t *queue = pb.cls_queue[CLS];
    uint32_t *next_count = next_cnt + PBRS_CNT_EXTEND, *shadow_count = cnt + PBRS_CNT_SHADOW;
    PBRS_WARP_LOOP(cnt[PBRS_CNT_CLS + CLS], i, active) {
        ShadeOut so; so.next = false; so.shadow_rays = 0;
        uint32_t j = 0u;
        if (active) {
            j = queue[i];
            so = stage_shade_scatter<CLS>(sc, pb, fp, bp, j, bounce, dg);
        }
        uint32_t s1, s2;
        warp_push2(next_count, so.next, shadow_count, so.shadow_rays > 0, s1, s2);
        if (so.next) next_queue[s1] = j;
        if (so.shadow_rays > 0) pb.shadow_queue[s2] = j;
        rays += (uint32_t)so.shadow_rays;
    }
    __syncwarp();
    warp_add_stat(pb.stats + kStatShadowRays, rays);
    flush_diag(pb.stats, dg);
}

__global__ void __launch_bounds__(kThreads) k_accumulate(PathBuffers pb, FrameParams fp, BatchParams bp, float *film) {
    PBRS_WARP_LOOP(bp.n_pixels, p, active) {
        if (active) stage_accumulate(pb, fp, bp, p, film);
    }
}

// parity side channels ------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_write_ids(DeviceScene sc, PathBuffers pb, FrameParams fp, BatchParams bp, uint32_t *out_inst,
                                                        uint32_t *out_prim, float *out_t) {
    PBRS_WARP_LOOP(bp.n_paths, j, active) {
        if (!active) continue;
        PathId id = decode_path(fp, bp, j);
        if (!id.valid) continue;
        uint4 h = *reinterpret_cast<const uint4 *>(pb.hit + j);
        size_t k = (size_t)(id.y - fp.y0) * (fp.x1 - fp.x0) + (id.x - fp.x0);
        bool hit = h.y != 0xFFFFFFFFu;
        uint32_t prim = 0xFFFFFFFFu;
        if (hit) {
            uint32_t kind = sc.inst_trav[h.y].shape_kind;
            prim = kind == PBRS_SHAPE_MESH ? sc.tris[h.z].orig : 0u;
        }
        if (out_inst) out_inst[k] = hit ? h.y : 0xFFFFFFFFu;
        if (out_prim) out_prim[k] = prim;
        if (out_t) out_t[k] = hit ? __uint_as_float(h.x) : PB_INF;
    }
}
__global__ void __launch_bounds__(kThreads) k_write_samples(PathBuffers pb, FrameParams fp, BatchParams bp, float *out) {
    PBRS_WARP_LOOP(bp.n_paths, j, active) {
        if (!active) continue;
        PathId id = decode_path(fp, bp, j);
        if (!id.valid) continue;
        float4 r = *reinterpret_cast<const float4 *>(pb.rad + j);
        size_t k = ((size_t)(id.y - fp.y0) * (fp.x1 - fp.x0) + (id.x - fp.x0)) * fp.spp + id.sample;
        out[3 * k] = r.x; out[3 * k + 1] = r.y; out[3 * k + 2] = r.z;
    }
}


struct Grid {
    int extend, extend_ext, extend_count, shadow, shadow_ext, shadow_count, small;
    int shade[2][PBRS_NUM_CLS];
    int surface, scatter[3];
};

template <int INTEGRATOR>
void launch_shade(const Grid &g, cudaStream_t stream, const DeviceScene &sc, const PathBuffers &pb, const FrameParams &fp, const BatchParams &bp,
                  uint32_t *cnt, uint32_t *next_queue, uint32_t *next_cnt, int stage) {
    const int *gs = g.shade[INTEGRATOR];
    // (a class no instance's material belongs to has an empty queue in every stage: its kernels are not launched at all)
    const auto has = [&sc](int cls) { return (sc.cls_mask >> cls) & 1u; };
    k_shade<PBRS_CLS_MISS, INTEGRATOR><<<gs[PBRS_CLS_MISS], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
    if (has(PBRS_CLS_EMISSIVE)) k_shade<PBRS_CLS_EMISSIVE, INTEGRATOR><<<gs[PBRS_CLS_EMISSIVE], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
    if (PBRS_SHADE_SPLIT && INTEGRATOR == PBRS_INTEGRATOR_PATH && sc.shade_split) {
        if (has(PBRS_CLS_LAMBERT) || has(PBRS_CLS_MICROFACET) || has(PBRS_CLS_MULTI)) k_surface<<<g.surface, kThreads, 0, stream>>>(sc, pb, cnt, stage);
        if (has(PBRS_CLS_LAMBERT)) k_scatter<PBRS_CLS_LAMBERT><<<g.scatter[0], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
        if (has(PBRS_CLS_MICROFACET)) k_scatter<PBRS_CLS_MICROFACET><<<g.scatter[1], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
        if (has(PBRS_CLS_MULTI)) k_scatter<PBRS_CLS_MULTI><<<g.scatter[2], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
    } else {
        if (has(PBRS_CLS_LAMBERT)) k_shade<PBRS_CLS_LAMBERT, INTEGRATOR><<<gs[PBRS_CLS_LAMBERT], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
        if (has(PBRS_CLS_MICROFACET)) k_shade<PBRS_CLS_MICROFACET, INTEGRATOR><<<gs[PBRS_CLS_MICROFACET], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
        if (has(PBRS_CLS_MULTI)) k_shade<PBRS_CLS_MULTI, INTEGRATOR><<<gs[PBRS_CLS_MULTI], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
    }
    if (has(PBRS_CLS_SPECULAR)) k_shade<PBRS_CLS_SPECULAR, INTEGRATOR><<<gs[PBRS_CLS_SPECULAR], kThreads, 0, stream>>>(sc, pb, fp, bp, cnt, next_queue, next_cnt, stage);
}
// kernels launched by launch_shade
template <int INTEGRATOR>
int shade_launches(const DeviceScene &sc) {
    const auto has = [&sc](int cls) { return (int)((sc.cls_mask >> cls) & 1u); };
    const int heavy = has(PBRS_CLS_LAMBERT) + has(PBRS_CLS_MICROFACET) + has(PBRS_CLS_MULTI);
    const bool split = PBRS_SHADE_SPLIT && INTEGRATOR == PBRS_INTEGRATOR_PATH && sc.shade_split;
    return 1 + has(PBRS_CLS_EMISSIVE) + has(PBRS_CLS_SPECULAR) + heavy + (split && heavy ? 1 : 0);
}
template <int INTEGRATOR>
void size_shade(Grid &g, int sms);

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                              \
            return PBRS_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)

template <class K>
int blocks_for(K kernel, int sms) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return sms * per_sm;
}
template <int INTEGRATOR>
void size_shade(Grid &g, int sms) {
    int *gs = g.shade[INTEGRATOR];
    gs[PBRS_CLS_MISS] = blocks_for(k_shade<PBRS_CLS_MISS, INTEGRATOR>, sms);
    gs[PBRS_CLS_EMISSIVE] = blocks_for(k_shade<PBRS_CLS_EMISSIVE, INTEGRATOR>, sms);
    gs[PBRS_CLS_LAMBERT] = blocks_for(k_shade<PBRS_CLS_LAMBERT, INTEGRATOR>, sms);
    gs[PBRS_CLS_MICROFACET] = blocks_for(k_shade<PBRS_CLS_MICROFACET, INTEGRATOR>, sms);
    gs[PBRS_CLS_SPECULAR] = blocks_for(k_shade<PBRS_CLS_SPECULAR, INTEGRATOR>, sms);
    gs[PBRS_CLS_MULTI] = blocks_for(k_shade<PBRS_CLS_MULTI, INTEGRATOR>, sms);
}

}  // namespace

struct Workspace {
    int device = -1;
    // Lanes: consecutive batches rotate over PBRS_LANES sets of path buffers, each on its own stream, so
    // that the tail of one batch's persistent kernels (a few straggling warps) and its small late-bounce
    // kernels overlap the other batches' work instead of leaving SMs idle.
    uint32_t capacity = 0;
    int n_lanes = 0;
    char *slab[kMaxLanes] = {};
    PathBuffers pb[kMaxLanes]{};
    cudaStream_t lane_stream[kMaxLanes] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[kMaxLanes] = {};
    uint32_t *counts = nullptr;
    uint32_t counts_cap = 0;   // in batches
    uint32_t *tiles = nullptr;
    uint32_t tiles_cap = 0;
    std::vector<uint32_t> tiles_host;  // what `tiles` holds
    unsigned long long *stats = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> stage_ev;  // PBRS_FLAG_TIME_STAGES: boundaries between launches
    int sms = 0;
    Grid grid{};
    bool grid_ready = false;
    // Small frames are launch-bound (C2: ~20 dependent launches against 0.8 ms of work), so their
    // whole enqueue sequence -- memsets, the fork onto the lanes, every kernel, the join -- is
    // captured once per distinct parameter set and replayed as one CUDA graph.
    cudaStream_t graph_stream = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_key;
};

void workspace_free(Workspace *w) {
    if (!w) return;
    for (auto &p : w->slab) if (p) cudaFree(p);
    for (auto &q : w->lane_stream) if (q) cudaStreamDestroy(q);
    if (w->fork_ev) cudaEventDestroy(w->fork_ev);
    for (auto &e : w->join_ev) if (e) cudaEventDestroy(e);
    if (w->counts) cudaFree(w->counts);
    if (w->tiles) cudaFree(w->tiles);
    if (w->stats) cudaFree(w->stats);
    for (auto &e : w->ev) if (e) cudaEventDestroy(e);
    for (auto &e : w->stage_ev) cudaEventDestroy(e);
    if (w->graph_exec) cudaGraphExecDestroy(w->graph_exec);
    if (w->graph_stream) cudaStreamDestroy(w->graph_stream);
    delete w;
}

static int workspace_prepare(Workspace *&wp, int device, uint32_t capacity, int n_lanes, uint32_t n_batches, uint32_t n_tiles) {
    if (!wp) wp = new Workspace();
    Workspace &w = *wp;
    if (w.device != device) {
        w.device = device;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        w.sms = prop.multiProcessorCount;
    }
    if (!w.grid_ready) {
        w.grid.extend = blocks_for(k_trace<false, false, false>, w.sms);
        w.grid.extend_ext = blocks_for(k_trace<false, false, true>, w.sms);
        w.grid.extend_count = blocks_for(k_trace<false, true, true>, w.sms);
        size_shade<PBRS_INTEGRATOR_DIRECT>(w.grid, w.sms);
        size_shade<PBRS_INTEGRATOR_PATH>(w.grid, w.sms);
        w.grid.shadow = blocks_for(k_trace<true, false, false>, w.sms);
        w.grid.shadow_ext = blocks_for(k_trace<true, false, true>, w.sms);
        w.grid.shadow_count = blocks_for(k_trace<true, true, true>, w.sms);
        w.grid.small = blocks_for(k_generate, w.sms);
        w.grid.surface = blocks_for(k_surface, w.sms);
        w.grid.scatter[0] = blocks_for(k_scatter<PBRS_CLS_LAMBERT>, w.sms);
        w.grid.scatter[1] = blocks_for(k_scatter<PBRS_CLS_MICROFACET>, w.sms);
        w.grid.scatter[2] = blocks_for(k_scatter<PBRS_CLS_MULTI>, w.sms);
        w.grid_ready = true;
    }
    if (!w.ev[0]) { CK(cudaEventCreate(&w.ev[0])); CK(cudaEventCreate(&w.ev[1])); }
    if (!w.stats) CK(cudaMalloc(&w.stats, sizeof(unsigned long long) * kStatCount));
    if (!w.lane_stream[0]) {
        for (int l = 0; l < kMaxLanes; ++l) { CK(cudaStreamCreateWithFlags(&w.lane_stream[l], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&w.join_ev[l], cudaEventDisableTiming)); }
        CK(cudaEventCreateWithFlags(&w.fork_ev, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&w.graph_stream, cudaStreamNonBlocking));
    }
    if (w.capacity < capacity || w.n_lanes < n_lanes) {
        for (auto &p : w.slab) if (p) { cudaFree(p); p = nullptr; }
        w.capacity = 0; w.n_lanes = 0;
        // 16 x 16-byte arrays + 1 float + (3 + PBRS_NUM_CLS) queues per path slot
        size_t per = 16 * sizeof(f4) + sizeof(float) + (3 + PBRS_NUM_CLS) * sizeof(uint32_t);
        size_t bytes = per * (size_t)capacity + 4096;
        for (int l = 0; l < n_lanes; ++l) {
            cudaError_t e = cudaMalloc(&w.slab[l], bytes);
            if (e != cudaSuccess) { cudaGetLastError(); set_error("path workspace: out of device memory"); return PBRS_ERR_OOM; }
            char *p = w.slab[l];
            auto take = [&](size_t elem) { char *q = p; p += elem * (size_t)capacity; return q; };
            PathBuffers &pb = w.pb[l];
            pb.ray_o = (f4 *)take(16); pb.ray_d = (f4 *)take(16); pb.hit = (u4 *)take(16); pb.beta = (f4 *)take(16);
            pb.rad = (f4 *)take(16); pb.aux = (f4 *)take(16);
            pb.sh_o1 = (f4 *)take(16); pb.sh_d1 = (f4 *)take(16); pb.sh_o2 = (f4 *)take(16); pb.sh_d2 = (f4 *)take(16);
            pb.sh_c = (f4 *)take(16); pb.sh_b = (f4 *)take(16);
            pb.sf_p = (f4 *)take(16); pb.sf_n = (f4 *)take(16); pb.sf_w = (f4 *)take(16); pb.sf_t = (f4 *)take(16);
            pb.sh_m = (float *)take(4);
            pb.queue[0] = (uint32_t *)take(4); pb.queue[1] = (uint32_t *)take(4); pb.shadow_queue = (uint32_t *)take(4);
            for (int c = 0; c < PBRS_NUM_CLS; ++c) pb.cls_queue[c] = (uint32_t *)take(4);
            pb.capacity = capacity;
        }
        w.capacity = capacity;
        w.n_lanes = n_lanes;
    }
    if (w.counts_cap < n_batches) {
        if (w.counts) cudaFree(w.counts);
        CK(cudaMalloc(&w.counts, sizeof(uint32_t) * PBRS_COUNTS_PER_BATCH * (size_t)n_batches));
        w.counts_cap = n_batches;
    }
    if (w.tiles_cap < n_tiles) {
        if (w.tiles) cudaFree(w.tiles);
        CK(cudaMalloc(&w.tiles, sizeof(uint32_t) * (size_t)n_tiles));
        w.tiles_cap = n_tiles;
        w.tiles_host.clear();
    }
    for (auto &pb : w.pb) pb.stats = w.stats;
    return 0;
}

int check_last_frame(Replica &r) {
    if (!r.workspace || !r.workspace->stats) return 0;
    unsigned long long overflow = 0;
    CK(cudaMemcpy(&overflow, r.workspace->stats + kStatPanic0 + P_STACK, sizeof overflow, cudaMemcpyDeviceToHost));
    if (overflow) { set_error("a traversal stack overflowed (scene deeper than the commit-time bound allows?)"); return PBRS_ERR_UNSUPPORTED; }
    return 0;
}

namespace {
// the caller's current device, put back when the call returns
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

// Tile t of the frame's 64x64 grid belongs to rank t % world_size (tile split), so every rank's
// share is a regular lattice over the whole image; without a tile split a rank renders them all.
void owned_tiles(const SceneImpl &s, const pbrs_render_opts &o, std::vector<uint32_t> &tiles) {
    const uint32_t W = s.cam.width, H = s.cam.height;
    uint32_t x0 = 0, y0 = 0, x1 = W, y1 = H;
    if (o.crop_w != 0 && o.crop_h != 0) { x0 = o.crop_x; y0 = o.crop_y; x1 = o.crop_x + o.crop_w; y1 = o.crop_y + o.crop_h; }
    const bool tile_split = o.world_size > 1 && o.split == PBRS_SPLIT_TILES;
    const uint32_t tiles_x = (W + 63) / 64;
    tiles.clear();
    if (x1 <= x0 || y1 <= y0) return;
    for (uint32_t ty = y0 / 64; ty <= (y1 - 1) / 64; ++ty)
        for (uint32_t tx = x0 / 64; tx <= (x1 - 1) / 64; ++tx) {
            uint32_t t = ty * tiles_x + tx;
            if (tile_split && (t % (uint32_t)o.world_size) != (uint32_t)o.rank) continue;
            tiles.push_back(t);
        }
}

int render_frame(const SceneImpl &s, Replica &r, const pbrs_render_opts &o, const RenderTargets &tg, cudaStream_t stream, pbrs_stats *st) {
    if (!s.committed) { set_error("render before pbrs_scene_commit"); return PBRS_ERR_STATE; }
    DeviceGuard device_guard;
    if (o.msaa == 0) { set_error("msaa must be >= 1"); return PBRS_ERR_INVALID_ARG; }
    if (o.world_size < 1 || o.rank < 0 || o.rank >= o.world_size) { set_error("bad rank/world_size"); return PBRS_ERR_INVALID_ARG; }
    if (o.integrator != PBRS_INTEGRATOR_PATH && o.integrator != PBRS_INTEGRATOR_DIRECT) { set_error("unknown integrator"); return PBRS_ERR_INVALID_ARG; }
    CK(cudaSetDevice(r.device));
    const uint32_t W = s.cam.width, H = s.cam.height;
    FrameParams fp{};
    fp.seed = o.seed; fp.msaa = o.msaa; fp.spp = o.msaa * o.msaa;
    fp.rank = (uint32_t)o.rank; fp.world = (uint32_t)o.world_size;
    fp.split_samples = (o.world_size > 1 && o.split == PBRS_SPLIT_SAMPLES) ? 1u : 0u;
    fp.only_sample = tg.only_sample;
    fp.integrator = o.integrator; fp.max_depth = o.max_depth; fp.flags = o.flags;
    fp.width = W; fp.height = H;
    if (o.crop_w == 0 || o.crop_h == 0) { fp.x0 = 0; fp.y0 = 0; fp.x1 = W; fp.y1 = H; }
    else {
        if (o.crop_x + o.crop_w > W || o.crop_y + o.crop_h > H) { set_error("crop outside the frame"); return PBRS_ERR_INVALID_ARG; }
        fp.x0 = o.crop_x; fp.y0 = o.crop_y; fp.x1 = o.crop_x + o.crop_w; fp.y1 = o.crop_y + o.crop_h;
    }
    if (fp.only_sample >= 0) fp.spp_r = 1;
    else if (fp.split_samples) fp.spp_r = fp.spp > fp.rank ? (fp.spp - fp.rank + fp.world - 1) / fp.world : 0;
    else fp.spp_r = fp.spp;
    std::vector<uint32_t> tiles;
    owned_tiles(s, o, tiles);
    fp.n_tiles = (uint32_t)tiles.size();

    // stages per path: the path integrator's bounce loop, or the direct integrator's two stages
    int n_stages = o.integrator == PBRS_INTEGRATOR_PATH ? std::max(o.max_depth, 0) : (o.max_depth > 0 ? 2 : 0);
    if (tg.only_sample >= 0) n_stages = 1;
    if (n_stages > PBRS_MAX_STAGES) { set_error("max_depth too large (at most 15 bounces)"); return PBRS_ERR_INVALID_ARG; }

    // default 16 Mi paths (~4 GB of path state): measured +15 % over 4 Mi (fewer, longer launches; tails amortised)
    // (a caller-supplied value is clamped from below: a tiny batch would mean millions of launches)
    uint32_t capacity = o.paths_in_flight ? std::max(o.paths_in_flight, 1u << 16) : (1u << 24);
    const uint64_t total_pixels = (uint64_t)fp.n_tiles * 4096u;
    if (total_pixels >> 32) { set_error("frame too large: more than 2^32 work pixels"); return PBRS_ERR_UNSUPPORTED; }
    const uint64_t total_paths = total_pixels * fp.spp_r;
    if (total_paths < capacity) capacity = (uint32_t)std::max<uint64_t>(total_paths, 32);
    if (capacity < fp.spp_r) capacity = fp.spp_r;

    uint32_t ppb = std::max<uint32_t>(1u, capacity / std::max(fp.spp_r, 1u));  // pixels per batch
    // End-to-end path (pbrs_render into a page-locked film, whole frame, no tile split): batches of
    // whole tile rows are contiguous row bands of the film, so each batch's band goes home on its own
    // lane stream as soon as it is accumulated -- the copy overlaps the other lanes' kernels.  A frame
    // that would fit one batch is cut in up to four, else the copy could not overlap anything.
    tg.host_copied = false;
    bool band_copies = false;
    {
        const bool whole = o.crop_w == 0 || o.crop_h == 0;
        const bool tile_split_now = o.world_size > 1 && o.split == PBRS_SPLIT_TILES;
        cudaPointerAttributes pa;
        const bool pinned = tg.host_film && cudaPointerGetAttributes(&pa, tg.host_film) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        static const bool env_no_bands = std::getenv("PBRS_NO_BAND_COPIES") != nullptr;  // development knob for A/B runs
        if (pinned && whole && !tile_split_now && tg.film && fp.spp_r > 0 && !env_no_bands) {
            const uint32_t tiles_x = (W + 63) / 64, tiles_y = (H + 63) / 64, row_px = tiles_x * 4096u;
            uint32_t rows_per_batch = ppb / row_px;
            if (rows_per_batch >= tiles_y && tiles_y >= 2) rows_per_batch = (tiles_y + std::min(4u, tiles_y) - 1) / std::min(4u, tiles_y);
            if (rows_per_batch >= 1) { ppb = rows_per_batch * row_px; band_copies = true; }
        }
    }
    const uint32_t n_batches = fp.spp_r == 0 ? 0 : (uint32_t)((total_pixels + ppb - 1) / ppb);

    const bool count_trav = (o.flags & PBRS_FLAG_COUNT_TRAVERSAL) != 0;
    const bool want_stats = st != nullptr;
    // per-stage timing needs one in-order stream; everything else pipelines batches over PBRS_LANES lanes
    const int n_lanes = (n_batches >= 2 && !(want_stats && (o.flags & PBRS_FLAG_TIME_STAGES))) ? PBRS_LANES : 1;
    Workspace *&wp = r.workspace;
    int rc = workspace_prepare(wp, r.device, capacity, n_lanes, std::max(n_batches, 1u), std::max(fp.n_tiles, 1u));
    if (rc < 0) return rc;
    Workspace &w = *wp;
    cudaStream_t const caller_stream = stream;

    // the tile list is uploaded only when it differs from the one already on the device (a pageable
    // source makes the copy synchronise with the host)
    if (fp.n_tiles && tiles != w.tiles_host) {
        CK(cudaMemcpyAsync(w.tiles, tiles.data(), sizeof(uint32_t) * tiles.size(), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        w.tiles_host = tiles;
    }
    fp.tiles = w.tiles;
    if (want_stats) CK(cudaEventRecord(w.ev[0], stream));

    uint64_t launches = 0, launches_extend = 0, launches_shadow = 0;
    const DeviceScene &sc = r.dscene;
    const bool time_stages = want_stats && (o.flags & PBRS_FLAG_TIME_STAGES) != 0;
    std::vector<int> ev_kind;
    enum { T_GEN = 0, T_EXT = 1, T_SHADE = 2, T_SHADOW = 3, T_ACC = 4 };
    size_t ev_used = 0;
    // Everything the frame puts on the GPU, on `origin` (and, forked from it, the lane streams).
    auto enqueue = [&](cudaStream_t origin) -> int {
        cudaStream_t stream = origin;
        cudaStream_t const caller_stream = origin;
        launches = launches_extend = launches_shadow = 0;
        CK(cudaMemsetAsync(w.stats, 0, sizeof(unsigned long long) * kStatCount, stream));
        if (n_batches) CK(cudaMemsetAsync(w.counts, 0, sizeof(uint32_t) * PBRS_COUNTS_PER_BATCH * (size_t)n_batches, stream));
        if (tg.film) CK(cudaMemsetAsync(tg.film, 0, sizeof(float) * 3 * (size_t)W * H, stream));
        // PBRS_FLAG_TIME_STAGES: one event after every launch; stage time = sum of the gaps that end
        // with a launch of that stage (the stream is in order, so a gap is that kernel's duration).

        auto mark = [&](int kind) -> int {
            if (!time_stages) return 0;
            if (ev_used == w.stage_ev.size()) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return -1;
                w.stage_ev.push_back(e);
            }
            cudaEventRecord(w.stage_ev[ev_used++], stream);
            ev_kind.push_back(kind);
            return 0;
        };
        mark(-1);
        if (n_lanes > 1) {
            CK(cudaEventRecord(w.fork_ev, caller_stream));
            for (int l = 0; l < n_lanes; ++l) CK(cudaStreamWaitEvent(w.lane_stream[l], w.fork_ev, 0));
        }
        for (uint32_t b = 0; b < n_batches; ++b) {
            PathBuffers pb = w.pb[b % (uint32_t)n_lanes];
            stream = n_lanes > 1 ? w.lane_stream[b % (uint32_t)n_lanes] : caller_stream;
            BatchParams bp;
            bp.first_pixel = (uint32_t)((uint64_t)b * ppb);
            bp.n_pixels = (uint32_t)std::min<uint64_t>(ppb, total_pixels - (uint64_t)b * ppb);
            bp.n_paths = bp.n_pixels * fp.spp_r;
            pb.counts = w.counts + (size_t)b * PBRS_COUNTS_PER_BATCH;
            k_generate<<<w.grid.small, kThreads, 0, stream>>>(sc, pb, fp, bp, pb.counts + PBRS_CNT_EXTEND);
            ++launches;
            mark(T_GEN);
            for (int stage = 0; stage < n_stages; ++stage) {
                uint32_t *q_in = pb.queue[stage & 1], *q_out = pb.queue[(stage + 1) & 1];
                uint32_t *cnt = pb.counts + PBRS_CNT_STRIDE * stage, *next_cnt = cnt + PBRS_CNT_STRIDE;
                const uint32_t *q_ext = q_in;
                if (count_trav) k_trace<false, true, true><<<w.grid.extend_count, kThreads, 0, stream>>>(sc, pb, q_ext, cnt);
                else if (sc.has_ext) k_trace<false, false, true><<<w.grid.extend_ext, kThreads, 0, stream>>>(sc, pb, q_ext, cnt);
                else k_trace<false, false, false><<<w.grid.extend, kThreads, 0, stream>>>(sc, pb, q_ext, cnt);
                ++launches; ++launches_extend;
                mark(T_EXT);
                if (tg.only_sample >= 0) break;
                if (o.integrator == PBRS_INTEGRATOR_PATH) launch_shade<PBRS_INTEGRATOR_PATH>(w.grid, stream, sc, pb, fp, bp, cnt, q_out, next_cnt, stage);
                else launch_shade<PBRS_INTEGRATOR_DIRECT>(w.grid, stream, sc, pb, fp, bp, cnt, q_out, next_cnt, stage);
                mark(T_SHADE);
                const uint32_t *q_sh = pb.shadow_queue;
                if (count_trav) k_trace<true, true, true><<<w.grid.shadow_count, kThreads, 0, stream>>>(sc, pb, q_sh, cnt);
                else if (sc.has_ext) k_trace<true, false, true><<<w.grid.shadow_ext, kThreads, 0, stream>>>(sc, pb, q_sh, cnt);
                else k_trace<true, false, false><<<w.grid.shadow, kThreads, 0, stream>>>(sc, pb, q_sh, cnt);
                mark(T_SHADOW);
                launches += (o.integrator == PBRS_INTEGRATOR_PATH ? shade_launches<PBRS_INTEGRATOR_PATH>(sc) : shade_launches<PBRS_INTEGRATOR_DIRECT>(sc)) + 1; ++launches_shadow;
            }
            if (tg.only_sample >= 0) {
                k_write_ids<<<w.grid.small, kThreads, 0, stream>>>(sc, pb, fp, bp, tg.ids_inst, tg.ids_prim, tg.ids_t);
                ++launches;
            } else {
                if (tg.film) { k_accumulate<<<w.grid.small, kThreads, 0, stream>>>(pb, fp, bp, tg.film); ++launches; mark(T_ACC); }
                if (band_copies) {  // this batch's tile rows are rows [y0, y1) of the film, full width: one contiguous copy
                    const uint32_t row_px = ((W + 63) / 64) * 4096u;
                    const uint32_t y0 = (bp.first_pixel / row_px) * 64u, y1 = std::min(H, ((bp.first_pixel + bp.n_pixels) / row_px) * 64u);
                    const size_t off = 3 * (size_t)W * y0;
                    if (y1 > y0) CK(cudaMemcpyAsync(tg.host_film + off, tg.film + off, sizeof(float) * 3 * (size_t)W * (y1 - y0), cudaMemcpyDeviceToHost, stream));
                }
                if (tg.samples) { k_write_samples<<<w.grid.small, kThreads, 0, stream>>>(pb, fp, bp, tg.samples); ++launches; mark(T_ACC); }
            }
        }
        if (n_lanes > 1)
            for (int l = 0; l < n_lanes; ++l) { CK(cudaEventRecord(w.join_ev[l], w.lane_stream[l])); CK(cudaStreamWaitEvent(caller_stream, w.join_ev[l], 0)); }
        return 0;
    };  // enqueue

    // a frame of a few batches is replayed from its graph; large frames (hundreds of batches of
    // millisecond kernels) gain nothing from it and are enqueued directly
    static const bool env_no_graph = std::getenv("PBRS_NO_GRAPH") != nullptr;  // development knob for A/B runs
    const bool use_graph = !time_stages && !env_no_graph && !(o.flags & PBRS_FLAG_NO_GRAPH) && n_batches >= 1 && n_batches <= 8;
    if (use_graph) {
        std::vector<unsigned char> key;
        auto put = [&key](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; key.insert(key.end(), b, b + n); };
        const uint32_t shape[8] = {n_batches, ppb, capacity, (uint32_t)n_stages, (uint32_t)n_lanes, count_trav ? 1u : 0u, (uint32_t)o.integrator, 0u};
        FrameParams fpk;
        std::memset(&fpk, 0, sizeof fpk);  // field by field: padding bytes must not enter the key
        fpk.seed = fp.seed; fpk.msaa = fp.msaa; fpk.spp = fp.spp; fpk.spp_r = fp.spp_r; fpk.rank = fp.rank; fpk.world = fp.world;
        fpk.split_samples = fp.split_samples; fpk.only_sample = fp.only_sample; fpk.integrator = fp.integrator; fpk.max_depth = fp.max_depth;
        fpk.flags = fp.flags; fpk.x0 = fp.x0; fpk.y0 = fp.y0; fpk.x1 = fp.x1; fpk.y1 = fp.y1; fpk.width = fp.width; fpk.height = fp.height;
        fpk.tiles = fp.tiles; fpk.n_tiles = fp.n_tiles;
        put(shape, sizeof shape); put(&fpk, sizeof fpk); put(&total_pixels, sizeof total_pixels);
        const void *ptrs[15] = {tg.film, tg.samples, tg.ids_inst, tg.ids_prim, tg.ids_t, w.slab[0], w.slab[1], w.slab[2], w.slab[3], w.counts, w.stats, w.tiles,
                                r.dscene.tlas_nodes, r.dscene.inst_trav, band_copies ? tg.host_film : nullptr};
        put(ptrs, sizeof ptrs);
        if (!w.graph_exec || key != w.graph_key) {
            if (w.graph_exec) { cudaGraphExecDestroy(w.graph_exec); w.graph_exec = nullptr; }
            CK(cudaStreamBeginCapture(w.graph_stream, cudaStreamCaptureModeThreadLocal));
            const int erc = enqueue(w.graph_stream);
            cudaGraph_t g = nullptr;
            const cudaError_t ee = cudaStreamEndCapture(w.graph_stream, &g);
            if (erc < 0 || ee != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                if (erc < 0) return erc;
                set_error(std::string("graph capture: ") + cudaGetErrorString(ee));
                return PBRS_ERR_CUDA;
            }
            const cudaError_t ie = cudaGraphInstantiate(&w.graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { w.graph_exec = nullptr; set_error(std::string("graph instantiate: ") + cudaGetErrorString(ie)); return PBRS_ERR_CUDA; }
            w.graph_key = key;
        } else {
            // the counters the capture pass would have produced
            for (uint32_t b = 0; b < n_batches; ++b) {
                ++launches;
                for (int stage = 0; stage < n_stages; ++stage) {
                    ++launches; ++launches_extend;
                    if (tg.only_sample >= 0) break;
                    launches += (o.integrator == PBRS_INTEGRATOR_PATH ? shade_launches<PBRS_INTEGRATOR_PATH>(sc) : shade_launches<PBRS_INTEGRATOR_DIRECT>(sc)) + 1; ++launches_shadow;
                }
                if (tg.only_sample >= 0) ++launches;
                else launches += (tg.film ? 1 : 0) + (tg.samples ? 1 : 0);
            }
        }
        CK(cudaGraphLaunch(w.graph_exec, caller_stream));
    } else {
        const int erc = enqueue(caller_stream);
        if (erc < 0) return erc;
    }
    stream = caller_stream;
    CK(cudaGetLastError());
    tg.host_copied = band_copies;
    if (want_stats) {
        CK(cudaEventRecord(w.ev[1], stream));
        CK(cudaEventSynchronize(w.ev[1]));
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]));
        std::vector<uint32_t> counts((size_t)PBRS_COUNTS_PER_BATCH * std::max(n_batches, 1u), 0u);
        unsigned long long stats[kStatCount];
        if (n_batches) CK(cudaMemcpy(counts.data(), w.counts, sizeof(uint32_t) * counts.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(stats, w.stats, sizeof stats, cudaMemcpyDeviceToHost));
        std::memset(st, 0, sizeof *st);
        uint64_t rays_extend = 0;
        for (uint32_t b = 0; b < n_batches; ++b)
            for (int stage = 0; stage < n_stages; ++stage) rays_extend += counts[(size_t)b * PBRS_COUNTS_PER_BATCH + PBRS_CNT_STRIDE * stage + PBRS_CNT_EXTEND];
        st->n_samples = counts.empty() ? 0 : 0;
        for (uint32_t b = 0; b < n_batches; ++b) st->n_samples += counts[(size_t)b * PBRS_COUNTS_PER_BATCH + PBRS_CNT_EXTEND];
        st->n_rays_extend = rays_extend;
        st->n_rays_shadow = stats[kStatShadowRays];
        for (int k = 0; k < 4; ++k) { st->trav_extend[k] = stats[kStatTrav + k]; st->trav_shadow[k] = stats[kStatTrav + 4 + k]; }
        st->n_nodes = st->trav_extend[0] + st->trav_shadow[0]; st->n_tris = st->trav_extend[1] + st->trav_shadow[1];
        st->n_spheres = st->trav_extend[2] + st->trav_shadow[2]; st->n_instances = st->trav_extend[3] + st->trav_shadow[3];
        st->launches_shadow = launches_shadow;
        if (time_stages) {
            double acc[5] = {0, 0, 0, 0, 0};
            for (size_t i = 1; i < ev_used; ++i) {
                float g = 0.0f;
                CK(cudaEventElapsedTime(&g, w.stage_ev[i - 1], w.stage_ev[i]));
                if (ev_kind[i] >= 0) acc[ev_kind[i]] += g;
            }
            st->ms_generate = acc[T_GEN]; st->ms_extend = acc[T_EXT]; st->ms_shade = acc[T_SHADE];
            st->ms_shadow = acc[T_SHADOW]; st->ms_accumulate = acc[T_ACC];
        }
        for (int k = 0; k < 16; ++k) st->would_panic[k] = stats[kStatPanic0 + k];
        st->ms_total = ms;
        st->launches = launches;
        st->launches_extend = launches_extend;
        if (stats[kStatPanic0 + P_STACK]) { set_error("a traversal stack overflowed"); return PBRS_ERR_UNSUPPORTED; }
    }
    return 0;
}

}  // namespace pbrs
