// Vote-scheduled persistent BVH8/Tri4 traversal for sm_100a: one ray per lane, but the warp
// decides together which kind of step runs next.
//
// The reference's single-ray loop (cpu_traverse_single_helper, src/traversal/mapping_cpu.impala:
// 138-256) alternates between two bodies: an inner-node visit (8 slab tests, pushes, sort) and a
// Tri4 test.  When 32 rays run it as 32 independent per-lane loops ("while-while", as in
// tools/bench_aila/kepler_dynamic_fetch.cu:70-371), the hardware serialises the two bodies and
// every nested loop exit; on the incoherent Sponza set that left 6.5 of 32 lanes active in the
// node body and 1.9 in the leaf body (profiles/r01_traverse_thread_blocks.txt).  Here each ray
// is an explicit state machine whose transitions are exactly the reference's, and the warp runs
// ONE straight-line step per iteration, chosen by majority vote (__ballot_sync + __popc) between
// the lanes that need a node step and those that need a Tri4 step; lanes of the other kind wait
// with their state in registers.  Finished lanes are refilled from a global counter together
// (one atomicAdd per warp, ranks from the ballot), as before.
//
// Per-ray order of node visits, pushes, sorts, Tri4 tests and tie-breaks is untouched, so hit
// records stay bit-identical to the oracle; only the interleaving BETWEEN rays changes.
#pragma once

#include "traverse.cuh"

namespace rb200 {

// The memory part of the traversal stack (stack.impala:53-54: 64 entries): the first SMEM_DEPTH
// levels in shared memory, [level][thread] so a warp's accesses to one level are conflict-free;
// deeper levels (< 0.1 % of the accesses on Sponza at depth 24) go to a thread-local array that
// is a separate object, so that the walker's scalar state stays in registers.
template <int SMEM_DEPTH, int BLOCK>
struct SplitStack {
    StackEntry* smem;                                   // this thread's column
    StackEntry* overflow;                               // kStackSize - SMEM_DEPTH thread-local entries
    __device__ __forceinline__ StackEntry load(int i) const {
        if (i < SMEM_DEPTH) return smem[i * BLOCK];
        return overflow[i - SMEM_DEPTH];
    }
    __device__ __forceinline__ void store(int i, StackEntry e) {
        if (i < SMEM_DEPTH) smem[i * BLOCK] = e;
        else overflow[i - SMEM_DEPTH] = e;
    }
};

// ARITY: 8 for Node8 (BVH8), 4 for Node4 (BVH4); the leaves are Tri4 in both.
// WIDE: fetch records with 256-bit loads (7 per Node8 instead of 14 x 128 bit, 7 per Tri4 instead of 13);
// needs 32-byte aligned arrays.
// FMA: contract the slab and triangle arithmetic (renderer only, see common.cuh).
template <bool ANY, int SMEM_DEPTH, int BLOCK, int ARITY = 8, bool WIDE = false, bool FMA = false>
struct RayWalker {
    RaySetup ray;
    float tmax;
    int top_node; float top_t; int ptr;
    int leaf;                       // next Tri4 of the leaf being tested, -1 when not inside a leaf
    HitRecord hit;
    SplitStack<SMEM_DEPTH, BLOCK> st;

    __device__ __forceinline__ void push(int n, float t) { ++ptr; st.store(ptr, StackEntry{top_node, top_t}); top_node = n; top_t = t; }
    __device__ __forceinline__ void push_after(int n, float t) { ++ptr; st.store(ptr, StackEntry{n, t}); }
    __device__ __forceinline__ void pop() { const StackEntry e = st.load(ptr); top_node = e.node; top_t = e.tmin; --ptr; }

    // The head of the reference's outer loop (:168-174): drop entries that start behind the current hit.
    __device__ __forceinline__ void cull() {
        if constexpr (!ANY) {
            while (top_node != 0 && top_t > tmax) pop();
        }
    }

    // `root`: 1 but for the lanes the hybrid packet kernel hands over at an inner entry of its stack (:309-319)
    __device__ __forceinline__ void begin(float4 r0, float4 r1, int root = 1) {
        ray.template init<ARITY / 4>(r0, r1);
        tmax = r1.w;
        hit.prim = -1; hit.geom = -1; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;   // empty_hit, intersection.impala:134-136
        ptr = -1; top_node = 0; top_t = kFltMax; leaf = -1;
        push(root, ray.tmin);                                                       // :153
        cull();
    }

    __device__ __forceinline__ bool finished() const { return leaf < 0 && top_node == 0; }
    __device__ __forceinline__ bool wants_node() const { return leaf < 0 && top_node > 0; }
    __device__ __forceinline__ bool wants_leaf() const { return leaf >= 0 || top_node < 0; }

    // One iteration of the inner `while (top_node > 0)` loop (:177-219).
    template <bool X86_NAN>
    __device__ __forceinline__ void node_step(const void* __restrict__ nodes) {
        constexpr int ROW = ARITY / 4;                       // float4 per bounds row
        // a node is 6 rows of ARITY floats, ARITY child ids and ARITY ints of padding: 8 * ARITY float4 / 4
        const float4* nb = reinterpret_cast<const float4*>(nodes) + size_t(top_node - 1) * (2 * ARITY);
        pop();
        float nx[ARITY], ny[ARITY], nz[ARITY], fx[ARITY], fy[ARITY], fz[ARITY];
        int child[ARITY];
        if constexpr (WIDE && ARITY == 8) {                  // a bounds row of a Node8 is one 256-bit load
            const F8 a = ldg8(nb + ray.near_x), b = ldg8(nb + ray.near_y), c = ldg8(nb + ray.near_z);
            const F8 d = ldg8(nb + ray.far_x), e = ldg8(nb + ray.far_y), f = ldg8(nb + ray.far_z);
            const F8 g = ldg8(nb + 6 * ROW);
#define RB_ROW(dst, src)                                                                             \
            dst[0] = src.lo.x; dst[1] = src.lo.y; dst[2] = src.lo.z; dst[3] = src.lo.w;              \
            dst[4] = src.hi.x; dst[5] = src.hi.y; dst[6] = src.hi.z; dst[7] = src.hi.w;
            RB_ROW(nx, a) RB_ROW(ny, b) RB_ROW(nz, c) RB_ROW(fx, d) RB_ROW(fy, e) RB_ROW(fz, f)
#undef RB_ROW
            child[0] = __float_as_int(g.lo.x); child[1] = __float_as_int(g.lo.y); child[2] = __float_as_int(g.lo.z); child[3] = __float_as_int(g.lo.w);
            child[4] = __float_as_int(g.hi.x); child[5] = __float_as_int(g.hi.y); child[6] = __float_as_int(g.hi.z); child[7] = __float_as_int(g.hi.w);
        } else
#pragma unroll
        for (int k = 0; k < ROW; k++) {                      // all loads of the node are issued before the first use
            const float4 a = ldg4(nb + ray.near_x + k), b = ldg4(nb + ray.near_y + k), c = ldg4(nb + ray.near_z + k);
            const float4 d = ldg4(nb + ray.far_x + k), e = ldg4(nb + ray.far_y + k), f = ldg4(nb + ray.far_z + k);
            const int4 g = ldg4(reinterpret_cast<const int4*>(nb) + 6 * ROW + k);
            nx[4 * k] = a.x; nx[4 * k + 1] = a.y; nx[4 * k + 2] = a.z; nx[4 * k + 3] = a.w;
            ny[4 * k] = b.x; ny[4 * k + 1] = b.y; ny[4 * k + 2] = b.z; ny[4 * k + 3] = b.w;
            nz[4 * k] = c.x; nz[4 * k + 1] = c.y; nz[4 * k + 2] = c.z; nz[4 * k + 3] = c.w;
            fx[4 * k] = d.x; fx[4 * k + 1] = d.y; fx[4 * k + 2] = d.z; fx[4 * k + 3] = d.w;
            fy[4 * k] = e.x; fy[4 * k + 1] = e.y; fy[4 * k + 2] = e.z; fy[4 * k + 3] = e.w;
            fz[4 * k] = f.x; fz[4 * k + 1] = f.y; fz[4 * k + 2] = f.z; fz[4 * k + 3] = f.w;
            child[4 * k] = g.x; child[4 * k + 1] = g.y; child[4 * k + 2] = g.z; child[4 * k + 3] = g.w;
        }

        // ordered slab test, intersection.impala:194-208 with integer min/max (:123-133, :184)
        float tentry[ARITY];
        unsigned mask = 0;
#pragma unroll
        for (int i = 0; i < ARITY; i++) {
            const float t0x = slab<X86_NAN, FMA>(ray.idx, nx[i], ray.iox);
            const float t0y = slab<X86_NAN, FMA>(ray.idy, ny[i], ray.ioy);
            const float t0z = slab<X86_NAN, FMA>(ray.idz, nz[i], ray.ioz);
            const float t1x = slab<X86_NAN, FMA>(ray.idx, fx[i], ray.iox);
            const float t1y = slab<X86_NAN, FMA>(ray.idy, fy[i], ray.ioy);
            const float t1z = slab<X86_NAN, FMA>(ray.idz, fz[i], ray.ioz);
            const float te = imax2(imax3(t0x, t0y, t0z), ray.tmin);
            const float tx = imin2(imin3(t1x, t1y, t1z), tmax);
            tentry[i] = te;
            if (!(__float_as_int(tx) < __float_as_int(te))) mask |= 1u << i;
        }
        if (mask == 0) {
            // :189-191.  Closest hit: back to the head of the outer loop (cull, then whatever is on top).
            // Any hit: `continue` of the inner loop, no culling in that mode anyway.
            cull();
            return;
        }
        // pushes in lane order, :195-208
#pragma unroll
        for (int i = 0; i < ARITY; i++) {
            if (mask & (1u << i)) {
                if (ANY || tentry[i] < top_t) push(child[i], tentry[i]);
                else push_after(child[i], tentry[i]);
            }
        }
        if (!ANY) {                                                                  // :210-218
            const int n = __popc(mask);
            if (n >= 3) {
                if (ARITY == 8) sort_entries(st, ptr - n + 1, n);
                else            sort_entries_bvh4(st, ptr - n + 1, n);
            }
        }
        // The reference goes on with the new top without a cull check (inner loop / :221-224).
    }

    // One Tri4 packet of the leaf loop (:224-249).  Returns true when an any-hit ray terminated (:252).
    template <bool WANT_GEOM>
    __device__ __forceinline__ bool leaf_step(const Tri4* __restrict__ tris) {
        if (leaf < 0) { leaf = ~top_node; pop(); }                                   // :224-225
        const float4* tp = reinterpret_cast<const float4*>(tris + leaf);
        int4 pid, gid_wide;
        float4 v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, nnx, nny, nnz;
        if constexpr (WIDE) {                                                         // a Tri4 is seven 256-bit loads
            const F8 p6 = ldg8(tp + 12);
            const F8 p0 = ldg8(tp + 0), p1 = ldg8(tp + 2), p2 = ldg8(tp + 4), p3 = ldg8(tp + 6), p4 = ldg8(tp + 8), p5 = ldg8(tp + 10);
            pid = make_int4(__float_as_int(p6.lo.x), __float_as_int(p6.lo.y), __float_as_int(p6.lo.z), __float_as_int(p6.lo.w));
            gid_wide = make_int4(__float_as_int(p6.hi.x), __float_as_int(p6.hi.y), __float_as_int(p6.hi.z), __float_as_int(p6.hi.w));
            v0x = p0.lo; v0y = p0.hi; v0z = p1.lo; e1x = p1.hi; e1y = p2.lo; e1z = p2.hi;
            e2x = p3.lo; e2y = p3.hi; e2z = p4.lo; nnx = p4.hi; nny = p5.lo; nnz = p5.hi;
        } else {
            pid = ldg4(reinterpret_cast<const int4*>(tp) + 12);
            v0x = ldg4(tp + 0); v0y = ldg4(tp + 1); v0z = ldg4(tp + 2);
            e1x = ldg4(tp + 3); e1y = ldg4(tp + 4); e1z = ldg4(tp + 5);
            e2x = ldg4(tp + 6); e2y = ldg4(tp + 7); e2z = ldg4(tp + 8);
            nnx = ldg4(tp + 9); nny = ldg4(tp + 10); nnz = ldg4(tp + 11);
            gid_wide = make_int4(0, 0, 0, 0);
        }

        float lt[4], lu[4], lv[4];
        unsigned hm = 0;
#define RB_LANE(j, c)                                                                               \
        lt[j] = kFltMax; lu[j] = 0.0f; lv[j] = 0.0f;                                                \
        if (pid.c != -1 &&                                                                          \
            intersect_tri_lane<FMA>(ray, tmax, v0x.c, v0y.c, v0z.c, e1x.c, e1y.c, e1z.c,                 \
                               e2x.c, e2y.c, e2z.c, nnx.c, nny.c, nnz.c, lt[j], lu[j], lv[j]))      \
            hm |= 1u << j;
        RB_LANE(0, x) RB_LANE(1, y) RB_LANE(2, z) RB_LANE(3, w)
#undef RB_LANE
        if (hm) {
            int lane;
            if (ANY) {
                lane = __ffs(hm) - 1;                                                // :234-237 (hit.prim >= 0 is `terminated`)
            } else {
                // cpu_reduce with integer min, then first lane equal to it (:239-242)
                const float mn = imin2(imin2(lt[0], lt[2]), imin2(lt[1], lt[3]));
                lane = lt[0] == mn ? 0 : lt[1] == mn ? 1 : lt[2] == mn ? 2 : 3;
            }
            const int p = lane == 0 ? pid.x : lane == 1 ? pid.y : lane == 2 ? pid.z : pid.w;
            hit.prim = p & 0x7FFFFFFF;                                               // mapping_cpu.impala:34
            hit.t = lane == 0 ? lt[0] : lane == 1 ? lt[1] : lane == 2 ? lt[2] : lt[3];
            hit.u = lane == 0 ? lu[0] : lane == 1 ? lu[1] : lane == 2 ? lu[2] : lu[3];
            hit.v = lane == 0 ? lv[0] : lane == 1 ? lv[1] : lane == 2 ? lv[2] : lv[3];
            if (WANT_GEOM) {
                const int4 gid = WIDE ? gid_wide : ldg4(reinterpret_cast<const int4*>(tp) + 13);
                hit.geom = lane == 0 ? gid.x : lane == 1 ? gid.y : lane == 2 ? gid.z : gid.w;
            }
            if (!ANY) tmax = hit.t;                                                  // :243
        }
        if (pid.w < 0) {                                                             // is_last, mapping_cpu.impala:40
            leaf = -1;
            if (ANY && hit.prim >= 0) { top_node = 0; return true; }                 // :252: the leaf is finished first
            cull();                                                                  // back to the head of the outer loop
        } else {
            leaf++;
        }
        return false;
    }
};

// What happens to a finished ray's record.
//   StoreAtOnce  the sink is called when the ray finishes (records in device memory).
//   PushHome     the host-pointer entry points with pinned caller memory (traverse.cu: run_host_direct): the records go
//                to the caller's array over PCIe, written by the kernel itself -- but never one by one: a write of
//                part of a cache line makes the host read the line first, and incoherent rays finish one at a time.
//                A finished ray's record stays in its lane until the warp refills (or drains); the sink then stores it
//                in device memory, the warp counts the records of every group of 16 rays (256 bytes), and the warp
//                whose count completes a group copies it home as whole lines.  Small groups, because a group waits
//                for its slowest ray.  (`counts` == nullptr: no groups, the sink writes wherever it likes.)
//                Any hit: what goes home is the triangle id alone (the caller's t, u, v stay as they are), 4 bytes per
//                ray in a dense array the host scatters into the records -- groups of 64 rays, 256 bytes again.
struct StoreAtOnce { static constexpr bool kActive = false; };
constexpr int kPushShift = 4;                      // records per group: 16 (closest hit)
constexpr int kPushShiftIds = 6;                   // ids per group: 64 (any hit); a group is 16 float4 either way
//                The rays may still be on their way when the kernel starts (`arriving` != nullptr): a copy engine is
//                writing them into the device array, whose slots were armed with all-ones words where tmin and tmax go.
//                A warp reserves rays only when the last one it would get has arrived (it goes on with the rays it
//                has meanwhile), every lane then checks its own slot and re-arms it for the next call.  A ray whose
//                tmin or tmax really is the all-ones NaN looks as if it had not arrived: `*copied` == epoch (written by
//                the same stream behind the copy) ends every wait.
constexpr unsigned kArmed = 0xFFFFFFFFu;
constexpr int kArrivalPatience = 1 << 26;          // polls of >= 128 ns: tens of seconds
struct PushHome {
    static constexpr bool kActive = true;
    unsigned* counts;                              // finished records per group, zeroed by the launcher
    const float4* staged; float4* home;            // the records in device memory; the caller's array
    float4* arriving; const unsigned* copied; unsigned epoch;
    static __device__ __forceinline__ unsigned peek(const void* p) {
        unsigned v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ bool arrived(int i) const {
        const unsigned* slot = reinterpret_cast<const unsigned*>(arriving + 2 * i);
        return (peek(slot + 3) != kArmed && peek(slot + 7) != kArmed) || peek(copied) == epoch;
    }
    // the fetch functor of a kernel whose rays are arriving: wait for the slot, take the ray, re-arm the slot
    __device__ __forceinline__ void take(int i, float4& r0, float4& r1) const {
        float4* slot = arriving + 2 * i;
        for (int polls = 0;; polls++) {
            r0 = __ldcv(slot); r1 = __ldcv(slot + 1);
            if (__float_as_uint(r0.w) != kArmed && __float_as_uint(r1.w) != kArmed) break;
            if (peek(copied) == epoch) { r0 = __ldcv(slot); r1 = __ldcv(slot + 1); break; }
            if (polls > kArrivalPatience) __trap();              // seconds: the copies have stopped coming; fail loudly rather than hang
            __nanosleep(128);
        }
        reinterpret_cast<unsigned*>(slot)[3] = kArmed;
        reinterpret_cast<unsigned*>(slot)[7] = kArmed;
    }
};

// The persistent loop shared by the bench_traversal kernels and the renderer's stream kernels.
//   fetch(i, r0, r1)  loads ray i (origin+tmin, direction+tmax)
//   sink(i, hit)      consumes the finished ray's record
// `refill_min`: idle lanes are refilled once at least this many wait (or none is busy).
// `node_streak_min`: see the node branch at the end of the loop (33: one step per vote).
template <bool ANY, bool WANT_GEOM, int SMEM_DEPTH, int BLOCK, int ARITY = 8, bool WIDE = false, bool FMA = false, typename Fetch, typename Sink, typename Records = StoreAtOnce>
__device__ __forceinline__ void traverse_vote_scheduled(const void* __restrict__ nodes, const Tri4* __restrict__ tris,
                                                        StackEntry* smem_column, int num_rays, int* __restrict__ work_counter,
                                                        int refill_min, Fetch fetch, Sink sink, int node_streak_min = 8, Records records = Records()) {
    const unsigned lane = lane_id();
    StackEntry overflow[kStackSize - SMEM_DEPTH];
    RayWalker<ANY, SMEM_DEPTH, BLOCK, ARITY, WIDE, FMA> w;
    w.st.smem = smem_column;
    w.st.overflow = overflow;
    w.leaf = -1; w.top_node = 0;
    int ray_idx = -1;
    bool drained = false;
    [[maybe_unused]] int starved = 0;          // PushHome, arriving rays: polls in a row with nothing to trace and nothing to fetch
    for (;;) {
        // a finished ray leaves its lane (PushHome: its record waits there, ray_idx = -2 - index, for the warp's next refill)
        if (ray_idx >= 0 && w.finished()) {
            if constexpr (Records::kActive) ray_idx = -2 - ray_idx;
            else { sink(ray_idx, w.hit); ray_idx = -1; }
        }
        // ---- refill idle lanes: one atomicAdd per warp, ranks from ballot/popc ----
        const unsigned idle = __ballot_sync(0xffffffffu, ray_idx < 0);
        if constexpr (Records::kActive) {
            const unsigned um = __ballot_sync(0xffffffffu, ray_idx < -1);
            if (um != 0 && (drained || __popc(idle) >= refill_min || idle == 0xffffffffu)) {
                const bool mine = ray_idx < -1;
                const int done = -2 - ray_idx;
                if (mine) { sink(done, w.hit); ray_idx = -1; }
                if (records.counts != nullptr) {
                    __threadfence();                                   // the records before their count
                    unsigned complete = 0;                             // this lane's count filled a group
                    constexpr int kShift = ANY ? kPushShiftIds : kPushShift;
                    const int units = ANY ? (num_rays + 3) >> 2 : num_rays;            // float4 to send home in all
                    const int group = done >> kShift;
                    if (mine) {
                        const unsigned same = __match_any_sync(um, group);
                        if (int(lane) == __ffs(same) - 1) {
                            const unsigned n = unsigned(__popc(same));
                            const unsigned size = unsigned(min(1 << kShift, num_rays - (group << kShift)));
                            complete = atomicAdd(records.counts + group, n) + n == size;
                        }
                    }
                    // two complete groups per round, 16 lanes each: 256 contiguous bytes = four whole lines per group
                    for (unsigned cm = __ballot_sync(0xffffffffu, complete != 0); cm != 0;) {
                        const int a = __ffs(cm) - 1; cm &= cm - 1;
                        const int b = cm != 0 ? __ffs(cm) - 1 : a; cm &= cm - 1;
                        const int ga = __shfl_sync(0xffffffffu, group, a), gb = __shfl_sync(0xffffffffu, group, b);
                        const bool second = lane >= 16u;
                        const int j = ((second ? gb : ga) << 4) + int(lane & 15u);
                        __threadfence();                               // the other warps' records after their counts
                        if (j < units && !(second && b == a)) records.home[j] = __ldcg(records.staged + j);
                    }
                }
            }
        }
        bool refill = !drained && (__popc(idle) >= refill_min || idle == 0xffffffffu);
        if constexpr (Records::kActive) {
            if (refill && records.arriving != nullptr) {       // only rays that have arrived are reserved
                const int leader = __ffs(idle) - 1;
                int ok = 1;
                if (int(lane) == leader) {
                    const int next = *reinterpret_cast<volatile int*>(work_counter);
                    ok = next >= num_rays || records.arrived(min(next + __popc(idle), num_rays) - 1);
                }
                refill = __shfl_sync(0xffffffffu, ok, leader) != 0;
                if (!refill && idle == 0xffffffffu) {
                    __nanosleep(256);
                    if (++starved > kArrivalPatience) __trap();
                } else {
                    starved = 0;
                }
            }
        }
        if (refill) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (int(lane) == leader) base = atomicAdd(work_counter, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ray_idx < 0) {
                const int i = base + __popc(idle & lanemask_lt());
                if (i < num_rays) {
                    ray_idx = i;
                    float4 r0, r1;
                    fetch(i, r0, r1);
                    w.begin(r0, r1);
                }
            }
            if (base + __popc(idle) >= num_rays) drained = true;
        }
        const bool has = ray_idx >= 0;
        const bool want_n = has && w.wants_node();
        const bool want_l = has && w.wants_leaf();
        const unsigned bn = __ballot_sync(0xffffffffu, want_n), bl = __ballot_sync(0xffffffffu, want_l);
        if ((bn | bl) == 0) {
            if (__ballot_sync(0xffffffffu, has) == 0 && drained) break;
            continue;                       // only lanes whose fresh ray was culled at once; they leave next round
        }
        if (__popc(bn) >= __popc(bl)) {
            // rays with a clamped axis need the x86 NaN pattern (RaySetup::degenerate); that variant is a
            // superset of the plain one, so the whole warp takes it when any of its lanes does
            if (__ballot_sync(0xffffffffu, want_n && w.ray.degenerate) != 0) {
                if (want_n) w.template node_step<true>(nodes);
            } else {
                // Node steps follow each other (most of a ray's steps are node steps): as long as at least
                // `node_streak_min` lanes want another one right away, take it without going through the refill /
                // finish / vote bookkeeping of the outer loop.  Lanes that reached a leaf or ran dry just wait.
                bool go = want_n;
                do {
                    if (go) w.template node_step<false>(nodes);
                    go = has && w.wants_node() && !w.ray.degenerate;
                } while (__popc(__ballot_sync(0xffffffffu, go)) >= node_streak_min);
            }
        } else {
            // (streaks of Tri4 steps were measured too and do not pay: most leaves are one packet)
            if (want_l) w.template leaf_step<WANT_GEOM>(tris);
        }
    }
}

}  // namespace rb200
