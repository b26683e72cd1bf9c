// bench_traversal entry points: persistent BVH8/Tri4 (and BVH4/Tri4) traversal kernels for sm_100a and the C ABI
// around them (include/rodent_b200.h).
//
// They replace gpu_traverse_single (src/traversal/mapping_gpu.impala:182-203) and compete with the Aila-Laine kernel
// (tools/bench_aila/kepler_dynamic_fetch.cu:70-371).  All variants are persistent grids (SMs x resident CTAs) that pull
// rays from a global counter with one atomicAdd per warp (ranks from ballot + popc) and follow the reference's own
// per-ray order; they differ in how a warp's rays share its lanes:
//   traverse_bvh8_vote / traverse_bvh4_vote   one ray per lane, the warp votes on the kind of step  (traverse_sched.cuh, default)
//   traverse_bvh2_vote                        the reference's GPU semantics on its BVH2 / Tri1 layout (traverse_bvh2.cuh)
//   traverse_bvh8_pool                        64 rays per warp in shared memory, compacted per step  (traverse_pool.cuh)
//   traverse_bvh8_persistent / _grid          one ray per lane, per-lane while-while loops          (traverse.cuh)
//   traverse_bvh8_quad                        four lanes per ray                                    (traverse_quad.cuh)
// Measurements and what each experiment showed: profiles/r01_experiments.md.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <vector>
#include <type_traits>

#include <emmintrin.h>

#include "traverse.cuh"
#include "traverse_quad.cuh"
#include "traverse_sched.cuh"
#include "traverse_pool.cuh"
#include "traverse_bvh2.cuh"

namespace rb200 {

constexpr int kBlock = 128;          // 4 warps per CTA
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr int kSmemStackDepth = 16;  // stack levels kept in shared memory by the persistent thread-per-ray kernel

struct Tuning {
    int mapping = 2;         // 3 = ray pool per warp (traverse_pool.cuh), 2 = vote-scheduled thread per ray (traverse_sched.cuh), 1 = while-while thread per ray (traverse.cuh),
                             // 4 = four lanes per ray (traverse_quad.cuh)
    int refill_min = 24;     // (mapping 2) refill idle lanes once this many wait
    int node_streak_min = 8;  // (mapping 2) consecutive node steps without re-voting while this many lanes want one (33: off)
    int vote_smem_depth = 24; // (mapping 2, 5 blocks, wide) stack levels in shared memory: 24, 16 or 12
    int wide_loads = 1;      // (mapping 2, 5 blocks) 256-bit record loads when the arrays are 32-byte aligned
    int bvh2_streak_min = 4; // BVH2 / Tri1 kernel: consecutive steps of one kind while this many lanes want one (measured 4 > 8 > 16)
    int bvh2_min_blocks = 8; // BVH2 / Tri1 kernel: __launch_bounds__ min blocks per SM of the variant launched: 8, 10 or 12
    int vote_min_blocks = 5; // (mapping 2) __launch_bounds__ min blocks per SM of the variant launched: 4, 5 or 6
    int persistent = 1;      // (mapping 1) 0: one thread per ray, plain grid
    int quad_refill_below = 6;   // (mapping 4) refill a warp when fewer than this many quads are busy
    int refill_below = 8;    // refill a warp when fewer than this many lanes are busy
    int pool_prefetch = 0;   // (mapping 3) L1 prefetch of a ray's next node / leaf while it waits in the pool
    int pool_refill_min = 24; // (mapping 3) refill a pool once this many of its 64 slots are empty
    int blocks_per_sm = 0;   // 0: occupancy API
    int packet_order = 0;    // packet / hybrid entry points: 1 = the reference's packet order to the bit (traverse_packets_ordered: one thread per packet), 0 = every ray gets the single-ray kernel's record
    int host_staging = 1;    // host-pointer entry points: pageable caller buffers go through pinned staging memory (0: straight to cudaMemcpyAsync)
    int host_direct = 1;     // host-pointer entry points, closest hit: one launch per call that follows its rays as they arrive and sends its records home itself (run_host_direct; 0: copy-engine pieces)
    int host_direct_rays = 1;    // ... 1 = a copy engine brings the rays in while the kernel runs (armed slots, traverse_sched.cuh), 0 = the warps read them from the caller's memory as they refill
    int host_staged_direct = 1;  // ... 1 = pageable records and packets go through the staging arrays with the same single launch, 2 = pageable rays as well (staged before the launch), 0 = neither (copy-engine pieces)
    int host_direct_push = 1;    // ... 1 = records staged in device memory and sent home group by group as whole lines, 0 = every record stored in the caller's memory by its lane, 2 = one copy after the kernel
    int host_trace = 0;          // ... print the device time of every such call (developer probe)
    int host_stream_stores = 1;  // ... staging copies with non-temporal stores
    int host_copy_parts = 8; // host-pointer entry points, pageable buffers: threads that share one staging memcpy (the caller's included; 8 measured a little better than 4: profiles/r02_probe_pageable.txt)
    int host_ramp = 0;       // host-pointer entry points: 1 = a small first piece as well (1, k-1, k-2, ..., 1 parts): traversal starts sooner
    int host_chunks = 5;     // host-pointer entry points: pieces the ray array is cut into for copy/compute overlap (two calls in flight: 5..6 measured best, one: 3..4)
};
static Tuning g_tuning;

template <bool ANY>
__device__ __forceinline__ void store_hit(Hit1* __restrict__ hits, int i, const HitRecord& h) {
    // make_cpu_hit1, tools/bench_traversal/bench_traversal.impala:121-131
    if (ANY) hits[i].tri_id = h.prim;
    else *reinterpret_cast<float4*>(hits + i) = make_float4(__int_as_float(h.prim), h.t, h.u, h.v);
}

template <bool ANY>
__global__ void __launch_bounds__(kBlock)
traverse_bvh8_grid(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= num_rays) return;
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    Traversal<ANY> tr;
    tr.begin(ldg4(rp), ldg4(rp + 1));
    tr.template run<false>(nodes, tris, [] { return false; });
    store_hit<ANY>(hits, i, tr.hit);
}

template <bool ANY>
__global__ void __launch_bounds__(kBlock, 6)
traverse_bvh8_persistent(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                         const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                         int* __restrict__ work_counter, int refill_below) {
    __shared__ int busy_lanes[kWarpsPerBlock];
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    volatile int* busy = &busy_lanes[warp];
    if (lane == 0) *busy = 0;
    __syncwarp();

    __shared__ StackEntry smem_stack[kSmemStackDepth][kBlock];
    Traversal<ANY, kSmemStackDepth, kBlock> tr;
    tr.st.smem = &smem_stack[0][threadIdx.x];
    int ray_idx = -1;            // -1: this lane holds no ray
    bool drained = false;        // the global queue is empty

    for (;;) {
        // ---- refill idle lanes: one atomicAdd per warp, ranks from ballot/popc ----
        const unsigned idle = __ballot_sync(0xffffffffu, ray_idx < 0);
        if (idle != 0 && !drained) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (int(lane) == leader) base = atomicAdd(work_counter, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ray_idx < 0) {
                const int i = base + __popc(idle & lanemask_lt());
                if (i < num_rays) {
                    ray_idx = i;
                    const float4* rp = reinterpret_cast<const float4*>(rays + i);
                    tr.begin(ldg4(rp), ldg4(rp + 1));
                }
            }
            if (base + __popc(idle) >= num_rays) drained = true;
        }
        const unsigned active = __ballot_sync(0xffffffffu, ray_idx >= 0);
        if (active == 0) break;
        if (lane == 0) *busy = __popc(active);
        __syncwarp();

        if (ray_idx >= 0) {
            const bool done = tr.template run<false>(nodes, tris, [&] { return !drained && *busy < refill_below; });
            if (done) {
                store_hit<ANY>(hits, ray_idx, tr.hit);
                ray_idx = -1;
                atomicSub(const_cast<int*>(busy), 1);
            }
        }
        __syncwarp();
    }
}


// Vote-scheduled persistent kernel (traverse_sched.cuh): the default.
constexpr int kVoteSmemDepth = 24;
// WIDE: 256-bit record loads (needs 32-byte aligned nodes / tris: the launcher checks).
template <bool ANY, int MIN_BLOCKS, bool WIDE = false, int DEPTH = kVoteSmemDepth>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
traverse_bvh8_vote(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                   int* __restrict__ work_counter, int refill_min, int node_streak_min) {
    __shared__ StackEntry smem_stack[DEPTH][kBlock];
    traverse_vote_scheduled<ANY, false, DEPTH, kBlock, 8, WIDE>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min,
        [rays](int i, float4& r0, float4& r1) {
            const float4* rp = reinterpret_cast<const float4*>(rays + i);
            r0 = ldg4(rp); r1 = ldg4(rp + 1);
        },
        [hits](int i, const HitRecord& h) { store_hit<ANY>(hits, i, h); }, node_streak_min);
}

// The host-pointer entry points with pinned caller memory: the default kernel reads its rays straight from the caller's
// array over PCIe as its warps refill, and sends the records home itself (PushHome, traverse_sched.cuh).  No copy engine,
// no second buffer for the rays, one launch per call.
template <bool ANY, int ARITY>
__global__ void __launch_bounds__(kBlock, ARITY == 8 ? 5 : 6)
traverse_direct(const void* __restrict__ nodes, const Tri4* __restrict__ tris,
                const Ray1* __restrict__ caller_rays, Hit1* __restrict__ hits, int num_rays,
                int* __restrict__ work_counter, int refill_min, int node_streak_min, PushHome records) {
    __shared__ StackEntry smem_stack[kVoteSmemDepth][kBlock];
    traverse_vote_scheduled<ANY, false, kVoteSmemDepth, kBlock, ARITY, ARITY == 8>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min,
        [caller_rays, records](int i, float4& r0, float4& r1) {
            if (records.arriving != nullptr) { records.take(i, r0, r1); return; }
            const float4* rp = reinterpret_cast<const float4*>(caller_rays + i);
            r0 = ldg4(rp); r1 = ldg4(rp + 1);        // (ld.global.cv of 16 bytes per lane over PCIe is ten times slower)
        },
        [hits, records](int i, const HitRecord& h) {
            if (ANY && records.counts != nullptr) reinterpret_cast<int32_t*>(hits)[i] = h.prim;      // the dense id array (see PushHome)
            else store_hit<ANY>(hits, i, h);
        }, node_streak_min, records);
}

// The same loop over a BVH4 (Node4: 6 rows of 4 floats, 4 children; leaves are Tri4 as well): the CPU single-ray
// path of the reference at its default --bvh-width (cpu_{intersect,occluded}_single_ray1_bvh4_tri4,
// tools/bench_traversal/bench_traversal.impala:279-305).
template <bool ANY>
__global__ void __launch_bounds__(kBlock, 6)
traverse_bvh4_vote(const Node4* __restrict__ nodes, const Tri4* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                   int* __restrict__ work_counter, int refill_min, int node_streak_min) {
    __shared__ StackEntry smem_stack[kVoteSmemDepth][kBlock];
    traverse_vote_scheduled<ANY, false, kVoteSmemDepth, kBlock, 4>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min,
        [rays](int i, float4& r0, float4& r1) {
            const float4* rp = reinterpret_cast<const float4*>(rays + i);
            r0 = ldg4(rp); r1 = ldg4(rp + 1);
        },
        [hits](int i, const HitRecord& h) { store_hit<ANY>(hits, i, h); }, node_streak_min);
}

// The reference's GPU path on its own layout (nvvm_{intersect,occluded}_single_ray1_bvh2_tri1,
// tools/bench_traversal/bench_traversal.impala:459-493): Node2 / Tri1 in, all four hit fields out in both modes
// (make_gpu_hit1, :78-83).
// MIN_BLOCKS: resident CTAs per SM asked of ptxas (8: 53 registers, 10: 48, 12: 40 with 24..36 bytes of spills); the
// incoherent set is bound by L2 latency and wants warps, the coherent one by issue slots (profiles/r01_traverse_bvh2_ncu.txt).
template <bool ANY, int MIN_BLOCKS, int SMEM_DEPTH>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
traverse_bvh2_vote(const Node2* __restrict__ nodes, const Tri1* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                   int* __restrict__ work_counter, int refill_min, int node_streak_min) {
    __shared__ int smem_stack[SMEM_DEPTH][kBlock];
    traverse_bvh2_scheduled<ANY, SMEM_DEPTH, kBlock>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min, node_streak_min,
        [rays](int i, float4& r0, float4& r1) {
            const float4* rp = reinterpret_cast<const float4*>(rays + i);
            r0 = ldg4(rp); r1 = ldg4(rp + 1);
        },
        [hits](int i, const HitRecord& h) {
            *reinterpret_cast<float4*>(hits + i) = make_float4(__int_as_float(h.prim), h.t, h.u, h.v);
        });
}

// Packet input (Ray4 / Ray8 in, Hit4 / Hit8 out, bench_traversal.impala:32-65): the same loop, fed from and draining
// into the structure-of-arrays packets; ray i is lane i % W of packet i / W.
template <bool ANY, int ARITY, int W>
__global__ void __launch_bounds__(kBlock, 5)
traverse_packets_vote(const void* __restrict__ nodes, const Tri4* __restrict__ tris,
                      const float* __restrict__ rays, float* __restrict__ hits, int num_rays,
                      int* __restrict__ work_counter, int refill_min, int node_streak_min) {
    __shared__ StackEntry smem_stack[kVoteSmemDepth][kBlock];
    traverse_vote_scheduled<ANY, false, kVoteSmemDepth, kBlock, ARITY>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min,
        [rays](int i, float4& r0, float4& r1) {
            const float* p = rays + size_t(i / W) * (8 * W) + (i % W);          // org[3][W] dir[3][W] tmin[W] tmax[W]
            r0 = make_float4(__ldg(p), __ldg(p + W), __ldg(p + 2 * W), __ldg(p + 6 * W));
            r1 = make_float4(__ldg(p + 3 * W), __ldg(p + 4 * W), __ldg(p + 5 * W), __ldg(p + 7 * W));
        },
        [hits](int i, const HitRecord& h) {
            float* p = hits + size_t(i / W) * (4 * W) + (i % W);                // tri_id[W] t[W] u[W] v[W]
            p[0] = __int_as_float(h.prim);
            if (!ANY) { p[W] = h.t; p[2 * W] = h.u; p[3 * W] = h.v; }           // make_cpu_hit4/8, bench_traversal.impala:133-157
        }, node_streak_min);
}

// The reference's packet / hybrid kernel in ITS order (cpu_traverse_hybrid_helper, src/traversal/mapping_cpu.impala:259-384):
// one stack of (node, tmin[W]) per packet, a node visited when any ray of the packet enters it, children pushed in child
// order without a sort, the triangles of a leaf tested one after the other against every interested ray (`t <= tmax`: of
// two hits at the same distance the later one stays), and -- hybrid -- the last few interested rays handed to the
// single-ray walk of the subtree.  One THREAD per packet, its W rays in turn: this is the mode for callers that need the
// packet kernels' records to the bit (rodent_b200_tune("packet_order", 1); tests/test_packet_oracle.py), not a fast one --
// the default gives every ray of a packet the single-ray kernel's record, which differs only where hits tie.
template <bool ANY, int ARITY, int W>
__global__ void __launch_bounds__(64)
traverse_packets_ordered(const void* __restrict__ nodes, const Tri4* __restrict__ tris,
                         const float* __restrict__ rays, float* __restrict__ hits, int num_packets, int hybrid) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num_packets) return;
    const float* rp = rays + size_t(p) * (8 * W);
    constexpr int kSwitch = W == 4 ? 3 : (ARITY == 4 ? 4 : 6);                       // :268-273
    RaySetup ray[W];
    float tmax[W], hit_t[W], hit_u[W], hit_v[W];
    int hit_prim[W];
    bool terminated[W];
    for (int l = 0; l < W; l++) {
        ray[l].template init<ARITY / 4>(make_float4(rp[l], rp[W + l], rp[2 * W + l], rp[6 * W + l]),
                                         make_float4(rp[3 * W + l], rp[4 * W + l], rp[5 * W + l], rp[7 * W + l]));
        tmax[l] = rp[7 * W + l];
        hit_prim[l] = -1; hit_t[l] = tmax[l]; hit_u[l] = 0.0f; hit_v[l] = 0.0f; terminated[l] = false;
    }
    int st_node[kStackSize + 8];
    float st_t[kStackSize + 8][W];
    int ptr = -1, top_node = 0;
    float top_t[W];
    for (int l = 0; l < W; l++) top_t[l] = kFltMax;
    auto push = [&](int n, const float* t) { ++ptr; st_node[ptr] = top_node; for (int l = 0; l < W; l++) st_t[ptr][l] = top_t[l]; top_node = n; for (int l = 0; l < W; l++) top_t[l] = t[l]; };
    auto push_after = [&](int n, const float* t) { ++ptr; st_node[ptr] = n; for (int l = 0; l < W; l++) st_t[ptr][l] = t[l]; };
    auto pop = [&] { top_node = st_node[ptr]; for (int l = 0; l < W; l++) top_t[l] = st_t[ptr][l]; --ptr; };
    { float t0[W]; for (int l = 0; l < W; l++) t0[l] = ray[l].tmin; push(1, t0); }   // :278

    StackEntry walk_stack[kStackSize];                                               // of the single-ray walks of the hybrid form
    for (;;) {
        // cull; hand the last interested rays to the single-ray walk (:303-326)
        bool done = false;
        for (;;) {
            if (top_node == 0) { done = true; break; }
            unsigned mask = 0;
            for (int l = 0; l < W; l++) if (top_t[l] <= tmax[l] && !terminated[l]) mask |= 1u << l;
            if (mask != 0) {
                if (hybrid && __popc(mask) <= kSwitch) {
                    for (unsigned m = mask; m != 0; m &= m - 1) {
                        const int l = __ffs(m) - 1;
                        RayWalker<ANY, 0, 64, ARITY, false, false> w;
                        w.st.smem = nullptr; w.st.overflow = walk_stack;
                        w.begin(make_float4(ray[l].ox, ray[l].oy, ray[l].oz, ray[l].tmin), make_float4(ray[l].dx, ray[l].dy, ray[l].dz, tmax[l]), top_node);
                        while (!w.finished()) {
                            if (w.wants_node()) w.template node_step<true>(nodes);
                            else if (w.template leaf_step<false>(tris)) break;
                        }
                        if (w.hit.prim >= 0) {
                            hit_prim[l] = w.hit.prim;
                            if (!ANY) { hit_t[l] = w.hit.t; hit_u[l] = w.hit.u; hit_v[l] = w.hit.v; tmax[l] = w.hit.t; }
                        }
                    }
                    if (ANY) for (int l = 0; l < W; l++) terminated[l] = hit_prim[l] >= 0;
                } else {
                    break;
                }
            }
            pop();
        }
        if (done) break;

        // inner nodes (:329-355)
        bool culled = false;
        while (top_node > 0) {
            const float* nb = reinterpret_cast<const float*>(nodes) + size_t(top_node - 1) * (8 * ARITY);
            const int* child = reinterpret_cast<const int*>(nb + 6 * ARITY);
            pop();
            bool pushed = false;
            for (int i = 0; i < ARITY; i++) {
                const int child_id = __ldg(child + i);
                if (child_id == 0) break;
                const float bx0 = __ldg(nb + i), bx1 = __ldg(nb + ARITY + i), by0 = __ldg(nb + 2 * ARITY + i), by1 = __ldg(nb + 3 * ARITY + i);
                const float bz0 = __ldg(nb + 4 * ARITY + i), bz1 = __ldg(nb + 5 * ARITY + i);
                float thit[W];
                bool any = false, nearer = false;
                for (int l = 0; l < W; l++) {                                        // intersect_ray_box, unordered, integer min / max
                    const float t0x = slab<true>(ray[l].idx, bx0, ray[l].iox), t1x = slab<true>(ray[l].idx, bx1, ray[l].iox);
                    const float t0y = slab<true>(ray[l].idy, by0, ray[l].ioy), t1y = slab<true>(ray[l].idy, by1, ray[l].ioy);
                    const float t0z = slab<true>(ray[l].idz, bz0, ray[l].ioz), t1z = slab<true>(ray[l].idz, bz1, ray[l].ioz);
                    const float tentry = imax2(imax2(imin2(t0x, t1x), imin2(t0y, t1y)), imax2(imin2(t0z, t1z), ray[l].tmin));
                    const float texit = imin2(imin2(imax2(t0x, t1x), imax2(t0y, t1y)), imin2(imax2(t0z, t1z), tmax[l]));
                    const bool miss = __float_as_int(texit) < __float_as_int(tentry);
                    thit[l] = miss ? kFltMax : tentry;
                    any |= !miss;
                }
                if (any) {
                    for (int l = 0; l < W; l++) nearer |= top_t[l] > thit[l];
                    if (ANY || nearer) push(child_id, thit);
                    else push_after(child_id, thit);
                    pushed = true;
                }
            }
            if (!pushed) { culled = true; break; }
        }
        if (culled) continue;

        // leaf (:357-381)
        if (top_node < 0) {
            bool active[W];
            for (int l = 0; l < W; l++) active[l] = top_t[l] <= tmax[l] && !terminated[l];
            int prim_id = ~top_node;
            pop();
            bool all_out = false;
            for (;;) {
                const float* tp = reinterpret_cast<const float*>(tris + prim_id);
                const int* ids = reinterpret_cast<const int*>(tp + 48);
                prim_id++;
                for (int i = 0; i < 4; i++) {
                    const int id = __ldg(ids + i);
                    if (id == -1) break;                                             // is_valid
                    const float v0x = __ldg(tp + i), v0y = __ldg(tp + 4 + i), v0z = __ldg(tp + 8 + i);
                    const float e1x = __ldg(tp + 12 + i), e1y = __ldg(tp + 16 + i), e1z = __ldg(tp + 20 + i);
                    const float e2x = __ldg(tp + 24 + i), e2y = __ldg(tp + 28 + i), e2z = __ldg(tp + 32 + i);
                    const float nx = __ldg(tp + 36 + i), ny = __ldg(tp + 40 + i), nz = __ldg(tp + 44 + i);
                    for (int l = 0; l < W; l++) {
                        if (!active[l]) continue;
                        float t, u, v;
                        if (intersect_tri_lane(ray[l], tmax[l], v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz, t, u, v)) {
                            hit_prim[l] = id & 0x7FFFFFFF; hit_t[l] = t; hit_u[l] = u; hit_v[l] = v;
                            tmax[l] = t;
                            if (ANY) { terminated[l] = true; active[l] = false; }
                        }
                    }
                    if (ANY) { bool all = true; for (int l = 0; l < W; l++) all &= terminated[l]; if (all) { all_out = true; break; } }
                }
                if (all_out) break;
                if (__ldg(ids + 3) < 0) break;                                       // is_last
            }
            if (all_out) break;
        }
    }
    float* hp = hits + size_t(p) * (4 * W);
    for (int l = 0; l < W; l++) {                                                    // make_cpu_hit4/8
        hp[l] = __int_as_float(hit_prim[l]);
        if (!ANY) { hp[W + l] = hit_t[l]; hp[2 * W + l] = hit_u[l]; hp[3 * W + l] = hit_v[l]; }
    }
}

// Ray-pool kernel (traverse_pool.cuh): 64 rays per warp in shared memory, compacted onto the lanes per step.
constexpr int kPoolBlock = 64;       // two warps (two pools, 29 KB) per CTA, seven CTAs per SM
template <bool ANY>
__global__ void __launch_bounds__(kPoolBlock, 7)
traverse_bvh8_pool(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                   int* __restrict__ work_counter, int refill_min, StackEntry* __restrict__ overflow, int prefetch) {
    __shared__ RayPool pools[kPoolBlock / 32];
    const int warp = threadIdx.x >> 5;
    const size_t global_warp = size_t(blockIdx.x) * (kPoolBlock / 32) + warp;
    traverse_pooled<ANY, false>(
        nodes, tris, pools[warp], overflow + global_warp * kPoolSlots * kPoolOverflow, num_rays, work_counter, refill_min, prefetch != 0,
        [rays](int i, float4& r0, float4& r1) {
            const float4* rp = reinterpret_cast<const float4*>(rays + i);
            r0 = ldg4(rp); r1 = ldg4(rp + 1);
        },
        [hits](int i, const HitRecord& h) { store_hit<ANY>(hits, i, h); });
}

// Quad-per-ray persistent kernel: eight rays per warp (traverse_quad.cuh).  Idle quads are
// refilled together: one atomicAdd per warp, ray index = base + rank of the quad among the
// idle ones (ballot + popc).
template <bool ANY>
__global__ void __launch_bounds__(kQuadBlock)
traverse_bvh8_quad(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                   const Ray1* __restrict__ rays, Hit1* __restrict__ hits, int num_rays,
                   int* __restrict__ work_counter, int refill_below) {
    __shared__ Entry stacks[kQuadBlock / 32][kStackSize][kQuadsPerWarp];
    __shared__ int busy_quads[kQuadBlock / 32];
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int q = lane >> 2;
    volatile int* busy = &busy_quads[warp];

    QuadTraversal<ANY> tr;
    tr.s = lane & 3;
    tr.qmask = 0xFu << (q * 4);
    tr.sp = &stacks[warp][0][q];
    int ray_idx = -1;
    bool drained = false;

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, ray_idx < 0) & 0x11111111u;   // one bit per idle quad
        if (idle != 0 && !drained) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (int(lane) == leader) base = atomicAdd(work_counter, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ray_idx < 0) {
                const int i = base + __popc(idle & ((1u << (q * 4)) - 1u));
                if (i < num_rays) {
                    ray_idx = i;
                    const float4* rp = reinterpret_cast<const float4*>(rays + i);
                    tr.begin(ldg4(rp), ldg4(rp + 1));
                }
            }
            if (base + __popc(idle) >= num_rays) drained = true;
        }
        const unsigned active = __ballot_sync(0xffffffffu, ray_idx >= 0) & 0x11111111u;
        if (active == 0) break;
        if (lane == 0) *busy = __popc(active);
        __syncwarp();

        if (ray_idx >= 0) {
            const bool done = tr.template run<false>(nodes, tris, [&] { return !drained && *busy < refill_below; });
            if (done) {
                if (tr.s == 0) {
                    store_hit<ANY>(hits, ray_idx, tr.hit);
                    atomicSub(const_cast<int*>(busy), 1);
                }
                ray_idx = -1;
            }
        }
        __syncwarp();
    }
}

// ---- per-device state ---------------------------------------------------------
// A BVH uploaded on behalf of a host-pointer call (cached_bvh below).
struct BvhCopy {
    void* d_nodes = nullptr; Tri4* d_tris = nullptr;
    size_t num_nodes = 0, num_tri4 = 0, node_size = 0;
    uint64_t fingerprint = 0, last_use = 0;
    unsigned char root[sizeof(Node8)] = {};       // the root node as uploaded
};
struct DeviceState {
    std::atomic<bool> init{false};
    int sm_count = 0;
    int* counter = nullptr;    // 64 ints: slot 0 for the synchronous entry points, 8/16/24 for the host-path streams
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    const char* last_kernel = "";   // the variant the last traversal launch on this device picked (rodent_b200_last_kernel_name)
    int occ[2] = {0, 0};       // resident CTAs per SM of the persistent kernels (closest, any)
    int occ_quad[2] = {0, 0};
    int occ_pool[2] = {0, 0};
    int occ_bvh4[2] = {0, 0};
    int occ_bvh2[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // [min blocks 8, 10, 12][closest, any]
    StackEntry* pool_overflow = nullptr; size_t pool_overflow_warps = 0;   // global backing of the pools' deep stack levels
    // host-pointer path: staging contexts (one per call in flight, reused) and the uploaded BVHs
    std::vector<struct HostContext*> idle_contexts;
    std::map<std::pair<const void*, const void*>, struct BvhCopy> bvh_cache;
    uint64_t bvh_clock = 0; int64_t bvh_uploads = 0, bvh_reuploads = 0;
    std::map<const void*, int> occupancy;      // kernel -> resident CTAs per SM
    // the ray copies of all direct host-pointer calls go through ONE stream, first come first served at the full rate of
    // the link: the first call's kernel has its rays early, the second's arrive while the first is traced
    cudaStream_t copy_in = nullptr; std::mutex copy_in_mutex;
};
// What one host-pointer call needs on the device.  The reference's cpu_* functions are reentrant, so their drop-ins are
// too: every call takes a context of its own, and concurrent calls (from several host threads) overlap on the device.
struct HostContext {
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    Ray1* d_rays = nullptr; Hit1* d_hits = nullptr; size_t ray_capacity = 0;
    int* counters = nullptr;   // 3 x 8 ints: one work counter per stream
    // pinned staging for callers whose buffers are pageable (grown on demand, kept)
    Ray1* h_rays = nullptr; Hit1* h_hits = nullptr; size_t stage_capacity = 0;
    cudaEvent_t piece_done[16] = {};
    unsigned* group_counts = nullptr; size_t group_capacity = 0;   // run_host_direct: finished records per group of 16 rays
    int32_t* h_ids = nullptr; size_t ids_capacity = 0;   // any-hit calls: the triangle ids as they come home (pinned, armed with kRecordArmed)
    bool hits_armed = false;               // every record of h_hits carries kRecordArmed in tri_id (the copy-out re-arms what it takes)
    bool rays_armed = false;               // every slot of d_rays carries the all-ones words (the kernel re-arms what it takes)
    unsigned* copied = nullptr; unsigned* epochs = nullptr; unsigned epoch = 0;   // device word / pinned values: "this call's rays are all in"
};
static DeviceState g_dev[64];
static std::mutex g_mutex;
static std::atomic<int64_t> g_launches{0};
static int g_host_dev = 0;
static std::vector<int> g_host_devs;      // rodent_b200_set_devices: the host-pointer entry points then cut every call over these

static DeviceState& device_state(int dev) {
    if (dev < 0 || dev >= 64) { std::fprintf(stderr, "rodent_b200: bad device %d\n", dev); std::abort(); }
    DeviceState& s = g_dev[dev];
    RB_CUDA_CHECK(cudaSetDevice(dev));
    if (!s.init.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!s.init.load(std::memory_order_relaxed)) {
            cudaDeviceProp prop;
            RB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
            if (prop.major != 10) {
                std::fprintf(stderr, "rodent_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", dev, prop.major, prop.minor);
                std::abort();
            }
            s.sm_count = prop.multiProcessorCount;
            RB_CUDA_CHECK(cudaMalloc(&s.counter, 256));
            RB_CUDA_CHECK(cudaEventCreate(&s.ev0));
            RB_CUDA_CHECK(cudaEventCreate(&s.ev1));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ[0], traverse_bvh8_persistent<false>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ[1], traverse_bvh8_persistent<true>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_quad[0], traverse_bvh8_quad<false>, kQuadBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_quad[1], traverse_bvh8_quad<true>, kQuadBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh4[0], traverse_bvh4_vote<false>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh4[1], traverse_bvh4_vote<true>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[0][0], traverse_bvh2_vote<false, 8, 32>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[0][1], traverse_bvh2_vote<true, 8, 32>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[1][0], traverse_bvh2_vote<false, 10, 24>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[1][1], traverse_bvh2_vote<true, 10, 24>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[2][0], traverse_bvh2_vote<false, 12, 24>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_bvh2[2][1], traverse_bvh2_vote<true, 12, 24>, kBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_pool[0], traverse_bvh8_pool<false>, kPoolBlock, 0));
            RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_pool[1], traverse_bvh8_pool<true>, kPoolBlock, 0));
            s.init.store(true, std::memory_order_release);
        }
    }
    return s;
}

// Resident CTAs per SM of `kernel` at `block` threads (static shared memory only), looked up once per device and kernel.
static int occupancy(DeviceState& s, const void* kernel, int block) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = s.occupancy.find(kernel);
    if (it != s.occupancy.end()) return it->second;
    int n = 0;
    RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, 0));
    s.occupancy[kernel] = std::max(n, 1);
    return std::max(n, 1);
}

template <bool ANY>
static void launch(DeviceState& s, const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits,
                   int num_rays, cudaStream_t stream, int* counter) {
    if (num_rays <= 0) return;
    if (g_tuning.mapping == 3) {
        if (!counter) counter = s.counter;
        RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
        const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : s.occ_pool[ANY ? 1 : 0];
        const int rays_per_block = kPoolBlock / 32 * kPoolSlots;
        const int needed = (num_rays + rays_per_block - 1) / rays_per_block;
        const int grid = std::min(needed, s.sm_count * per_sm);
        const size_t warps = size_t(s.sm_count) * 16 * (kPoolBlock / 32);       // upper bound of any grid launched here
        if (s.pool_overflow_warps < warps) {
            std::lock_guard<std::mutex> lock(g_mutex);
            if (s.pool_overflow_warps < warps) {
                if (s.pool_overflow) RB_CUDA_CHECK(cudaFree(s.pool_overflow));
                RB_CUDA_CHECK(cudaMalloc(&s.pool_overflow, warps * kPoolSlots * kPoolOverflow * sizeof(StackEntry)));
                s.pool_overflow_warps = warps;
            }
        }
        s.last_kernel = "traverse_bvh8_pool";
        traverse_bvh8_pool<ANY><<<std::min(grid, s.sm_count * 16), kPoolBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.pool_refill_min, s.pool_overflow, g_tuning.pool_prefetch);
    } else if (g_tuning.mapping == 2) {
        if (!counter) counter = s.counter;
        RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
        const int v = std::min(std::max(g_tuning.vote_min_blocks, 4), 6);
        const bool wide = g_tuning.wide_loads && ((reinterpret_cast<uintptr_t>(nodes) | reinterpret_cast<uintptr_t>(tris)) & 31) == 0;
        using Kernel = void (*)(const Node8*, const Tri4*, const Ray1*, Hit1*, int, int*, int, int);
        Kernel kernel = v == 4 ? traverse_bvh8_vote<ANY, 4> : v == 6 ? traverse_bvh8_vote<ANY, 6> :
                        !wide ? traverse_bvh8_vote<ANY, 5> :
                        g_tuning.vote_smem_depth >= 24 ? traverse_bvh8_vote<ANY, 5, true> :
                        g_tuning.vote_smem_depth >= 16 ? traverse_bvh8_vote<ANY, 5, true, 16> : traverse_bvh8_vote<ANY, 5, true, 12>;
        s.last_kernel = v == 4 ? (ANY ? "traverse_bvh8_vote<true, 4>" : "traverse_bvh8_vote<false, 4>") :
                        v == 6 ? (ANY ? "traverse_bvh8_vote<true, 6>" : "traverse_bvh8_vote<false, 6>") :
                        !wide ? (ANY ? "traverse_bvh8_vote<true, 5>" : "traverse_bvh8_vote<false, 5>") :
                        g_tuning.vote_smem_depth >= 24 ? (ANY ? "traverse_bvh8_vote<true, 5, true, 24>" : "traverse_bvh8_vote<false, 5, true, 24>") :
                        g_tuning.vote_smem_depth >= 16 ? (ANY ? "traverse_bvh8_vote<true, 5, true, 16>" : "traverse_bvh8_vote<false, 5, true, 16>") :
                                                         (ANY ? "traverse_bvh8_vote<true, 5, true, 12>" : "traverse_bvh8_vote<false, 5, true, 12>");
        // the persistent grid is sized from the occupancy of the very instantiation that is launched
        const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : occupancy(s, reinterpret_cast<const void*>(kernel), kBlock);
        const int needed = (num_rays + kBlock - 1) / kBlock;
        const int grid = std::min(needed, s.sm_count * per_sm);
        kernel<<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_min, g_tuning.node_streak_min);
    } else if (g_tuning.mapping == 4) {
        if (!counter) counter = s.counter;
        RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
        const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : s.occ_quad[ANY ? 1 : 0];
        const int rays_per_block = kQuadBlock / 4;
        const int needed = (num_rays + rays_per_block - 1) / rays_per_block;
        const int grid = std::min(needed, s.sm_count * per_sm);
        s.last_kernel = "traverse_bvh8_quad";
        traverse_bvh8_quad<ANY><<<grid, kQuadBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.quad_refill_below);
    } else if (g_tuning.persistent) {
        if (!counter) counter = s.counter;
        RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
        const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : s.occ[ANY ? 1 : 0];
        const int needed = (num_rays + kBlock - 1) / kBlock;
        const int grid = std::min(needed, s.sm_count * per_sm);     // a multiple of the SM count when saturated
        s.last_kernel = "traverse_bvh8_persistent";
        traverse_bvh8_persistent<ANY><<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_below);
    } else {
        traverse_bvh8_grid<ANY><<<(num_rays + kBlock - 1) / kBlock, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays);
    }
    RB_CUDA_CHECK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// BVH4 input: the vote-scheduled kernel only.
template <bool ANY>
static void launch(DeviceState& s, const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits,
                   int num_rays, cudaStream_t stream, int* counter) {
    if (num_rays <= 0) return;
    if (!counter) counter = s.counter;
    RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : s.occ_bvh4[ANY ? 1 : 0];
    const int grid = std::min((num_rays + kBlock - 1) / kBlock, s.sm_count * per_sm);
    traverse_bvh4_vote<ANY><<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_min, g_tuning.node_streak_min);
    RB_CUDA_CHECK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// BVH2 / Tri1 input.
template <bool ANY>
static void launch(DeviceState& s, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits,
                   int num_rays, cudaStream_t stream, int* counter) {
    if (num_rays <= 0) return;
    if (reinterpret_cast<uintptr_t>(nodes) & 31) {
        std::fprintf(stderr, "rodent_b200: the Node2 array must be 32-byte aligned (cudaMalloc / rodent_b200_alloc_device give 256)\n");
        std::abort();
    }
    if (!counter) counter = s.counter;
    RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    const int v = g_tuning.bvh2_min_blocks >= 12 ? 2 : g_tuning.bvh2_min_blocks >= 10 ? 1 : 0;
    const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : s.occ_bvh2[v][ANY ? 1 : 0];
    const int grid = std::min((num_rays + kBlock - 1) / kBlock, s.sm_count * per_sm);
    if (v == 0) traverse_bvh2_vote<ANY, 8, 32><<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_min, g_tuning.bvh2_streak_min);
    if (v == 1) traverse_bvh2_vote<ANY, 10, 24><<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_min, g_tuning.bvh2_streak_min);
    if (v == 2) traverse_bvh2_vote<ANY, 12, 24><<<grid, kBlock, 0, stream>>>(nodes, tris, rays, hits, num_rays, counter, g_tuning.refill_min, g_tuning.bvh2_streak_min);
    RB_CUDA_CHECK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// The synchronous device-pointer calls share the device's default work counter, its timing events and last_ms, so they
// are serialised per device (the reference's nvvm_* exports are called from one thread); callers that want
// concurrent launches use the _async forms with their own counters.
static std::mutex g_sync_mutex[64];
template <bool ANY, typename NodeT, typename TriT>
static void run_sync(int dev, const NodeT* nodes, const TriT* tris, const Ray1* rays, Hit1* hits, int num_rays) {
    DeviceState& s = device_state(dev);
    std::lock_guard<std::mutex> lock(g_sync_mutex[dev]);
    RB_CUDA_CHECK(cudaEventRecord(s.ev0, 0));
    launch<ANY>(s, nodes, tris, rays, hits, num_rays, 0, nullptr);
    RB_CUDA_CHECK(cudaEventRecord(s.ev1, 0));
    RB_CUDA_CHECK(cudaEventSynchronize(s.ev1));
    float ms = 0.0f;
    RB_CUDA_CHECK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    s.last_ms = ms;
}

// ---- host-pointer path ----------------------------------------------------------
// Extent of a BVH given only its arrays: walk from the root (node 1).  The arrays come from the caller and are not
// trusted: ids are bounded, every node is visited once, a leaf's packet run is bounded.
constexpr size_t kMaxBvhNodes = size_t(1) << 26, kMaxBvhTri4 = size_t(1) << 28;
template <typename NodeT>
static void bvh_extent(const NodeT* nodes, const Tri4* tris, size_t& num_nodes, size_t& num_tri4) {
    std::vector<int> todo{1};
    std::vector<bool> seen;
    num_nodes = 0; num_tri4 = 0;
    size_t visited = 0;
    auto bad = [](const char* why) { std::fprintf(stderr, "rodent_b200: malformed BVH passed to a host-pointer entry point (%s)\n", why); std::abort(); };
    while (!todo.empty()) {
        const int id = todo.back(); todo.pop_back();
        if (size_t(id) > kMaxBvhNodes) bad("node id out of range");
        if (seen.size() < size_t(id)) seen.resize(std::max(size_t(id), seen.size() * 2), false);
        if (seen[id - 1]) bad("a node is reachable twice");
        seen[id - 1] = true;
        if (++visited > kMaxBvhNodes) bad("too many nodes");
        num_nodes = std::max(num_nodes, size_t(id));
        const NodeT& n = nodes[id - 1];
        for (int i = 0; i < int(sizeof(n.child) / sizeof(n.child[0])); i++) {
            const int c = n.child[i];
            if (c > 0) todo.push_back(c);
            else if (c < 0) {
                size_t k = size_t(~c);
                if (k >= kMaxBvhTri4) bad("leaf index out of range");
                while (tris[k].prim_id[3] >= 0) {        // is_last sentinel, mapping_cpu.impala:40
                    if (++k >= kMaxBvhTri4) bad("leaf without an end marker");
                }
                num_tri4 = std::max(num_tri4, k + 1);
            }
        }
    }
}

// The cpu_* functions these entry points replace are pure: they read whatever the arrays hold at the time of the call.
// The uploaded copy is therefore keyed on CONTENT, not on the addresses alone: every call hashes the first and last KB of
// both arrays and 64 cache lines spread over each (about 12 KB, ~1 us) and compares with what was uploaded; a caller
// that rebuilt a BVH in place or reused an allocation gets a fresh upload.  (A change confined to bytes the samples miss
// is not seen: callers that patch a BVH in place call rodent_b200_forget_bvh.)  At most kBvhCacheEntries copies are kept
// per device, least recently used first out.
constexpr size_t kBvhCacheEntries = 8;
static uint64_t sample_hash(const void* base, size_t bytes, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(base);
    auto mix = [&h](const unsigned char* q, size_t n) {
        for (size_t i = 0; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, q + i, 8); h = (h ^ w) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; }
    };
    const size_t edge = std::min<size_t>(bytes, 1024);
    mix(p, edge);
    mix(p + bytes - edge, edge);
    if (bytes > 2048)
        for (int k = 1; k <= 64; k++) mix(p + ((bytes - 64) * size_t(k) / 65 & ~size_t(7)), 64);
    return h ^ bytes;
}
static uint64_t bvh_fingerprint(const void* nodes, size_t node_bytes, const void* tris, size_t tri_bytes) {
    return sample_hash(tris, tri_bytes, sample_hash(nodes, node_bytes, 0xCBF29CE484222325ull));
}

template <typename NodeT>
static std::pair<NodeT*, Tri4*> cached_bvh(DeviceState& s, const NodeT* nodes, const Tri4* tris) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto key = std::make_pair((const void*)nodes, (const void*)tris);
    auto it = s.bvh_cache.find(key);
    if (it != s.bvh_cache.end()) {
        BvhCopy& c = it->second;
        // The root node first: it exists in every BVH, so reading it is safe whatever now lives at this address; only
        // when it is the one that was uploaded are the sampled lines of the (then almost surely same-sized) arrays read.
        if (c.node_size == sizeof(NodeT) && std::memcmp(nodes, c.root, sizeof(NodeT)) == 0 &&
            c.fingerprint == bvh_fingerprint(nodes, c.num_nodes * sizeof(NodeT), tris, c.num_tri4 * sizeof(Tri4))) {
            c.last_use = ++s.bvh_clock;
            return std::make_pair(static_cast<NodeT*>(c.d_nodes), c.d_tris);
        }
        RB_CUDA_CHECK(cudaDeviceSynchronize());          // another call may still trace the stale copy
        RB_CUDA_CHECK(cudaFree(c.d_nodes)); RB_CUDA_CHECK(cudaFree(c.d_tris));
        s.bvh_cache.erase(it);
        s.bvh_reuploads++;
    }
    if (s.bvh_cache.size() >= kBvhCacheEntries) {
        auto lru = s.bvh_cache.begin();
        for (auto i = s.bvh_cache.begin(); i != s.bvh_cache.end(); ++i) if (i->second.last_use < lru->second.last_use) lru = i;
        RB_CUDA_CHECK(cudaDeviceSynchronize());
        RB_CUDA_CHECK(cudaFree(lru->second.d_nodes)); RB_CUDA_CHECK(cudaFree(lru->second.d_tris));
        s.bvh_cache.erase(lru);
    }
    BvhCopy c;
    bvh_extent(nodes, tris, c.num_nodes, c.num_tri4);
    c.node_size = sizeof(NodeT);
    c.fingerprint = bvh_fingerprint(nodes, c.num_nodes * sizeof(NodeT), tris, c.num_tri4 * sizeof(Tri4));
    NodeT* dn; Tri4* dt;
    RB_CUDA_CHECK(cudaMalloc(&dn, c.num_nodes * sizeof(NodeT)));
    RB_CUDA_CHECK(cudaMalloc(&dt, std::max<size_t>(c.num_tri4, 1) * sizeof(Tri4)));
    RB_CUDA_CHECK(cudaMemcpy(dn, nodes, c.num_nodes * sizeof(NodeT), cudaMemcpyHostToDevice));
    RB_CUDA_CHECK(cudaMemcpy(dt, tris, c.num_tri4 * sizeof(Tri4), cudaMemcpyHostToDevice));
    std::memcpy(c.root, nodes, sizeof(NodeT));
    c.d_nodes = dn; c.d_tris = dt; c.last_use = ++s.bvh_clock;
    s.bvh_cache[key] = c;
    s.bvh_uploads++;
    return std::make_pair(dn, dt);
}

static HostContext* acquire_host_context(DeviceState& s, size_t num_rays) {
    HostContext* c = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        // prefer an idle context that is already large enough
        for (size_t i = 0; i < s.idle_contexts.size(); i++)
            if (s.idle_contexts[i]->ray_capacity >= num_rays) { c = s.idle_contexts[i]; s.idle_contexts.erase(s.idle_contexts.begin() + i); break; }
        if (!c && !s.idle_contexts.empty()) { c = s.idle_contexts.back(); s.idle_contexts.pop_back(); }
    }
    if (!c) {
        c = new HostContext();
        for (auto& st : c->streams) RB_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        RB_CUDA_CHECK(cudaMalloc(&c->counters, 3 * 8 * sizeof(int)));
        RB_CUDA_CHECK(cudaMalloc(&c->copied, sizeof(unsigned)));
        RB_CUDA_CHECK(cudaMemset(c->copied, 0, sizeof(unsigned)));
        RB_CUDA_CHECK(cudaMallocHost(&c->epochs, 8 * sizeof(unsigned)));
    }
    if (c->ray_capacity < num_rays) {
        if (c->d_rays) { RB_CUDA_CHECK(cudaFree(c->d_rays)); RB_CUDA_CHECK(cudaFree(c->d_hits)); }
        RB_CUDA_CHECK(cudaMalloc(&c->d_rays, num_rays * sizeof(Ray1)));
        RB_CUDA_CHECK(cudaMalloc(&c->d_hits, num_rays * sizeof(Hit1)));
        c->ray_capacity = num_rays;
        c->rays_armed = false;
    }
    const size_t groups = (num_rays >> kPushShift) + 1;
    if (c->group_capacity < groups) {
        if (c->group_counts) RB_CUDA_CHECK(cudaFree(c->group_counts));
        RB_CUDA_CHECK(cudaMalloc(&c->group_counts, groups * sizeof(unsigned)));
        c->group_capacity = groups;
    }
    return c;
}
static void release_host_context(DeviceState& s, HostContext* c) {
    std::lock_guard<std::mutex> lock(g_mutex);
    s.idle_contexts.push_back(c);
}
// The ray-pool variant (mapping 3) indexes one per-device overflow buffer by warp: its launches must not overlap.
static std::mutex g_pool_serial;

// A few helper threads for the host side of calls with pageable buffers: copies between the caller's memory and pinned
// staging memory (one memcpy runs at ~10 GB/s on one core; the PCIe link takes five times that).  A task returns true when
// it is finished and false to be run again later (a copy-out task whose records have not all arrived yet); run_all
// returns when all its tasks are finished, the calling thread works along.
// memcpy with non-temporal stores: the destination of a staging copy is not read by this core again (the copy engine or
// the caller takes it from memory), so the lines need neither be fetched for ownership nor displace anything cached.
static void stream_copy(void* dst, const void* src, size_t bytes) {
    char* d = static_cast<char*>(dst); const char* s = static_cast<const char*>(src);
    const size_t head = std::min(bytes, size_t(-reinterpret_cast<uintptr_t>(d)) & 15);
    std::memcpy(d, s, head); d += head; s += head; bytes -= head;
    size_t blocks = bytes / 64;
    for (; blocks > 0; blocks--, d += 64, s += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
    }
    _mm_sfence();
    std::memcpy(d, s, bytes % 64);
}

class CopyPool {
public:
    using Task = std::function<bool()>;
    // never destroyed: its threads wait on the condition variable for the life of the process, and destroying a
    // condition variable that has waiters blocks (glibc) -- a static instance would hang every process at exit
    static CopyPool& get() { static CopyPool* p = new CopyPool; return *p; }
    void run_all(std::vector<Task>& tasks) {
        if (tasks.empty()) return;
        std::atomic<int> pending{int(tasks.size())};
        {
            std::lock_guard<std::mutex> lock(m_);
            for (size_t k = tasks.size(); k-- > 0;) jobs_.push_back(Job{&tasks[k], &pending});     // popped from the back: in order
        }
        cv_.notify_all();
        while (pending.load(std::memory_order_acquire) > 0) {
            Job j;
            if (take(j, &pending)) run(j);
            else for (int k = 0; k < 32; k++) __builtin_ia32_pause();
        }
    }
    void parallel_copy(void* dst, const void* src, size_t bytes) {
        const size_t kMin = size_t(1) << 20;
        const int parts = int(std::min<size_t>(size_t(std::max(1, std::min(g_tuning.host_copy_parts, threads_ + 1))), std::max<size_t>(1, bytes / kMin)));
        const bool stream = g_tuning.host_stream_stores != 0;
        if (parts <= 1) { if (stream) stream_copy(dst, src, bytes); else std::memcpy(dst, src, bytes); return; }
        const size_t each = ((bytes / parts) + 4095) & ~size_t(4095);
        std::vector<Task> tasks;
        for (int k = 0; k < parts; k++) {
            const size_t b = std::min(bytes, each * k), e = std::min(bytes, each * (k + 1));
            tasks.push_back([=] {
                if (stream) stream_copy(static_cast<char*>(dst) + b, static_cast<const char*>(src) + b, e - b);
                else std::memcpy(static_cast<char*>(dst) + b, static_cast<const char*>(src) + b, e - b);
                return true;
            });
        }
        run_all(tasks);
    }
private:
    struct Job { Task* task; std::atomic<int>* pending; };
    // a copy is cut into at most g_tuning.host_copy_parts parts (the caller's thread takes one); the pool is large enough
    // for the calls of several devices' threads at once (rodent_b200_set_devices) without oversubscribing a small host
    int threads_ = 0;
    CopyPool() {
        threads_ = int(std::max(3u, std::min(16u, std::thread::hardware_concurrency() / 2)));
        for (int i = 0; i < threads_; i++) std::thread([this] { work(); }).detach();
    }
    // a job of the caller's own batch if there is one (`mine`), else nothing: the caller never takes on another call's wait
    bool take(Job& j, std::atomic<int>* mine) {
        std::lock_guard<std::mutex> lock(m_);
        for (size_t k = jobs_.size(); k-- > 0;)
            if (jobs_[k].pending == mine) { j = jobs_[k]; jobs_.erase(jobs_.begin() + k); return true; }
        return false;
    }
    void run(const Job& j) {
        if ((*j.task)()) { j.pending->fetch_sub(1, std::memory_order_release); return; }
        {
            std::lock_guard<std::mutex> lock(m_);
            jobs_.insert(jobs_.begin(), j);                      // not finished: behind everything that is queued
        }
        for (int k = 0; k < 64; k++) __builtin_ia32_pause();
    }
    void work() {
        // A call hands out its copies piece after piece, a few hundred microseconds apart: a helper that went to sleep on
        // the condition variable after every job would need ~50 us to wake up for the next one.  So it keeps looking
        // at the queue for a while after its last job and only then blocks.
        auto last_job = std::chrono::steady_clock::now() - std::chrono::seconds(1);
        for (;;) {
            Job j;
            bool have = false;
            {
                std::unique_lock<std::mutex> lock(m_);
                if (jobs_.empty() && std::chrono::steady_clock::now() - last_job > std::chrono::microseconds(kLingerUs))
                    cv_.wait(lock, [this] { return !jobs_.empty(); });
                if (!jobs_.empty()) { j = jobs_.back(); jobs_.pop_back(); have = true; }
            }
            if (!have) {
                for (int k = 0; k < 64; k++) __builtin_ia32_pause();
                continue;
            }
            run(j);
            last_job = std::chrono::steady_clock::now();
        }
    }
    static constexpr int kLingerUs = 300;
    std::mutex m_; std::condition_variable cv_; std::vector<Job> jobs_;
};

static bool is_pinned(const void* p, bool host_only = false) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || (!host_only && a.type == cudaMemoryTypeManaged);
}

// Copy-in / trace / copy-out, pipelined in chunks over three streams so the PCIe
// transfers of one chunk overlap the traversal of another.
template <bool ANY, typename NodeT>
static void run_host_on(int dev, const NodeT* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int num_rays);

// With several devices (rodent_b200_set_devices) a call is cut into contiguous ray ranges, one per device, BVH replicated
// (uploaded to each on first use); every range is traced and copied straight back into its slice of the caller's array,
// so the path needs no collective.
template <bool ANY, typename NodeT>
static void run_host(const NodeT* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int num_rays) {
    if (num_rays <= 0) return;
    const std::vector<int> devs = g_host_devs;
    if (devs.size() <= 1 || num_rays < int(devs.size()) * 4096) return run_host_on<ANY>(devs.empty() ? g_host_dev : devs[0], nodes, tris, rays, hits, num_rays);
    std::vector<std::thread> threads;
    const int64_t n = int64_t(devs.size());
    for (int64_t k = 0; k < n; k++) {
        const int b = int(num_rays * k / n), e = int(num_rays * (k + 1) / n);
        threads.emplace_back([=] { run_host_on<ANY>(devs[size_t(k)], nodes, tris, rays + b, hits + b, e - b); });
    }
    for (auto& t : threads) t.join();
}

// Pinned caller buffers, default kernel (BVH8 or BVH4): one launch per call and one copy (traverse_direct).  Against
// the copy-engine pieces below: the traversal starts at once instead of after a fifth of the rays, pays the tail of a
// launch once per call, keeps all its CTAs for the whole call, and the records need no pass of their own.
static void ensure_staging(HostContext* c, size_t num_rays) {
    if (c->stage_capacity >= num_rays) return;
    if (c->h_rays) { RB_CUDA_CHECK(cudaFreeHost(c->h_rays)); RB_CUDA_CHECK(cudaFreeHost(c->h_hits)); }
    RB_CUDA_CHECK(cudaMallocHost(&c->h_rays, num_rays * sizeof(Ray1)));
    RB_CUDA_CHECK(cudaMallocHost(&c->h_hits, num_rays * sizeof(Hit1)));
    c->stage_capacity = num_rays;
    c->hits_armed = false;
}

// Ray packets (org[3][W] dir[3][W] tmin[W] tmax[W]) -> rays [first, first + n) of the staging array, by the helper threads;
// first and n are multiples of 4.  SSE 4 x 4 transposes, non-temporal stores.
template <int W>
static void stage_packets(HostContext* c, const float* ray_packets, int first, int n) {
    std::vector<CopyPool::Task> tasks;
    const int parts = std::max(1, std::min(g_tuning.host_copy_parts, 16)), each = std::max(((n / parts + W * 4 - 1) / (W * 4)) * (W * 4), W * 4);
    for (int b0 = first; b0 < first + n; b0 += each) {
        const int e0 = std::min(first + n, b0 + each);
        tasks.push_back([=] {
            for (int i = b0; i < e0; i += 4) {                       // four rays of one packet at a time
                const float* q = ray_packets + size_t(i / W) * (8 * W) + (i % W);
                __m128 ox = _mm_loadu_ps(q), oy = _mm_loadu_ps(q + W), oz = _mm_loadu_ps(q + 2 * W), t0 = _mm_loadu_ps(q + 6 * W);
                __m128 dx = _mm_loadu_ps(q + 3 * W), dy = _mm_loadu_ps(q + 4 * W), dz = _mm_loadu_ps(q + 5 * W), t1 = _mm_loadu_ps(q + 7 * W);
                _MM_TRANSPOSE4_PS(ox, oy, oz, t0);
                _MM_TRANSPOSE4_PS(dx, dy, dz, t1);
                float* r = reinterpret_cast<float*>(c->h_rays + i);
                _mm_stream_ps(r, ox); _mm_stream_ps(r + 4, dx); _mm_stream_ps(r + 8, oy); _mm_stream_ps(r + 12, dy);
                _mm_stream_ps(r + 16, oz); _mm_stream_ps(r + 20, dz); _mm_stream_ps(r + 24, t0); _mm_stream_ps(r + 28, t1);
            }
            _mm_sfence();
            return true;
        });
    }
    CopyPool::get().run_all(tasks);
}
// ... and records [first, first + n) of the staging array -> hit packets (tri_id[W] t[W] u[W] v[W]).
template <int W>
static void unstage_packets(HostContext* c, float* hit_packets, int first, int n) {
    std::vector<CopyPool::Task> tasks;
    const int parts = std::max(1, std::min(g_tuning.host_copy_parts, 16)), each = std::max(((n / parts + W * 4 - 1) / (W * 4)) * (W * 4), W * 4);
    for (int b0 = first; b0 < first + n; b0 += each) {
        const int e0 = std::min(first + n, b0 + each);
        tasks.push_back([=] {
            for (int i = b0; i < e0; i += 4) {
                const float* r = reinterpret_cast<const float*>(c->h_hits + i);
                __m128 a = _mm_load_ps(r), b = _mm_load_ps(r + 4), d = _mm_load_ps(r + 8), e = _mm_load_ps(r + 12);
                _MM_TRANSPOSE4_PS(a, b, d, e);                       // rows: tri_id, t, u, v of the four rays
                float* q = hit_packets + size_t(i / W) * (4 * W) + (i % W);
                _mm_storeu_ps(q, a); _mm_storeu_ps(q + W, b); _mm_storeu_ps(q + 2 * W, d); _mm_storeu_ps(q + 3 * W, e);
            }
            return true;
        });
    }
    CopyPool::get().run_all(tasks);
}

constexpr int32_t kRecordArmed = int32_t(0x80000000u);   // tri_id of a staging record that has not arrived (a real one is -1 or >= 0)

// `stage_in` / `stage_out`: the caller's rays / hits are pageable and go through the context's pinned staging arrays --
// still ONE launch.  The calling thread and the helper threads copy the rays into the staging array piece by piece, each
// piece followed by its copy to the device; the kernel is launched when the last copy is queued and follows what is
// still on its way (the armed slots tell it what is in); the records arrive in the (armed) staging array by groups
// while the kernel runs, and the same threads move them on to the caller's array chunk by chunk as they turn up,
// re-arming the staging array on the way.
//
// W = 4 or 8: the caller's arrays are the packet layouts of the packet / hybrid entry points (RayW: org[3][W] dir[3][W]
// tmin[W] tmax[W]; HitW: tri_id[W] t[W] u[W] v[W]).  They always go through the staging arrays: the helper threads
// transpose packets into Ray1 on the way in and records back into packets on the way out, the device sees rays.
template <bool ANY, typename NodeT, int W = 1>
static bool run_host_direct(DeviceState& s, HostContext* c, const NodeT* d_nodes, const Tri4* d_tris, const void* rays_v, void* hits_v, int num_rays,
                            bool stage_in = false, bool stage_out = false) {
    constexpr int ARITY = int(sizeof(NodeT::child) / sizeof(int32_t));
    const Ray1* rays = static_cast<const Ray1*>(rays_v); Hit1* hits = static_cast<Hit1*>(hits_v);     // W == 1
    const float* ray_packets = static_cast<const float*>(rays_v); float* hit_packets = static_cast<float*>(hits_v);   // W > 1
    if (W > 1) { stage_in = true; stage_out = true; }
    // Any hit: only tri_id changes.  The ids come home as a dense array (pinned, armed) and the helper threads write them
    // into the caller's records as they arrive -- whatever memory those are in, so nothing is staged on the way out.
    if (ANY) {
        stage_out = false;
        if (c->ids_capacity < size_t(num_rays) + 64) {
            if (c->h_ids) RB_CUDA_CHECK(cudaFreeHost(c->h_ids));
            c->ids_capacity = size_t(num_rays) + 64;
            RB_CUDA_CHECK(cudaMallocHost(&c->h_ids, c->ids_capacity * sizeof(int32_t)));
            std::fill(c->h_ids, c->h_ids + c->ids_capacity, kRecordArmed);
        }
    }
    const Ray1* src = stage_in ? c->h_rays : rays;
    Hit1* home = ANY ? reinterpret_cast<Hit1*>(c->h_ids) : stage_out ? c->h_hits : hits;
    const Ray1* caller_rays = nullptr; Hit1* caller_hits = nullptr;
    if (cudaHostGetDevicePointer(const_cast<void**>(reinterpret_cast<const void**>(&caller_rays)), const_cast<Ray1*>(src), 0) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&caller_hits), home, 0) != cudaSuccess) {
        cudaGetLastError();                      // page-locked but not mapped for this device: the copy-engine path takes it
        return false;
    }
    cudaStream_t run = c->streams[0];
    const bool staged = stage_in || stage_out;
    const bool push = ANY || staged || g_tuning.host_direct_push == 1, copy_after = !push && g_tuning.host_direct_push == 2;
    const bool by_copy_engine = staged || g_tuning.host_direct_rays;
    PushHome records{push ? c->group_counts : nullptr, reinterpret_cast<const float4*>(c->d_hits), reinterpret_cast<float4*>(caller_hits), nullptr, c->copied, ++c->epoch};
    unsigned* epoch_value = c->epochs + (records.epoch & 7);
    *epoch_value = records.epoch;
    auto copy_in = [&](int first, int n, bool last) {
        std::lock_guard<std::mutex> lock(s.copy_in_mutex);
        if (!s.copy_in) RB_CUDA_CHECK(cudaStreamCreateWithFlags(&s.copy_in, cudaStreamNonBlocking));
        RB_CUDA_CHECK(cudaMemcpyAsync(c->d_rays + first, src + first, size_t(n) * sizeof(Ray1), cudaMemcpyHostToDevice, s.copy_in));
        if (last) RB_CUDA_CHECK(cudaMemcpyAsync(c->copied, epoch_value, sizeof(unsigned), cudaMemcpyHostToDevice, s.copy_in));
    };
    if (by_copy_engine) {
        // arm the slots if somebody else wrote to this context's ray array since (the kernel re-arms every slot it takes)
        if (!c->rays_armed) { RB_CUDA_CHECK(cudaMemsetAsync(c->d_rays, 0xFF, c->ray_capacity * sizeof(Ray1), run)); RB_CUDA_CHECK(cudaStreamSynchronize(run)); c->rays_armed = true; }
        if (!stage_in) copy_in(0, num_rays, true);           // everything the kernel waits for is queued before it
        records.arriving = reinterpret_cast<float4*>(c->d_rays);
    }
    if (stage_out && !c->hits_armed) {
        std::vector<CopyPool::Task> arm;
        const size_t step = size_t(1) << 16;
        for (size_t b0 = 0; b0 < c->stage_capacity; b0 += step)
            arm.push_back([=] { for (size_t i = b0; i < std::min(c->stage_capacity, b0 + step); i++) c->h_hits[i] = Hit1{kRecordArmed, 0.0f, 0.0f, 0.0f}; return true; });
        CopyPool::get().run_all(arm);
        c->hits_armed = true;
    }
    // Pieces of ~4 MB: into the staging array, then queued for the copy engine.  The kernel is launched only when every
    // copy it will wait for is queued: a running kernel must never depend on a CUDA call this process has yet to make (a
    // cudaFree in another thread waits for the device to go idle and may keep other threads' calls from going in
    // meanwhile; a profiler may run a kernel to its end inside the launch call).  The copies overlap the staging of the
    // later pieces, and the kernel finds most of its rays in place and follows the rest.
    if (stage_in) {
        const int piece = 1 << 17;
        for (int first = 0; first < num_rays; first += piece) {
            const int n = std::min(piece, num_rays - first);
            if (W == 1) {
                CopyPool::get().parallel_copy(c->h_rays + first, rays + first, size_t(n) * sizeof(Ray1));
            } else {
                stage_packets<W>(c, ray_packets, first, n);
            }
            copy_in(first, n, first + n >= num_rays);
        }
    }
    cudaEvent_t ev[2] = {};
    if (g_tuning.host_trace) { for (auto& e : ev) RB_CUDA_CHECK(cudaEventCreate(&e)); RB_CUDA_CHECK(cudaEventRecord(ev[0], run)); }
    if (push) RB_CUDA_CHECK(cudaMemsetAsync(c->group_counts, 0, (size_t((num_rays - 1) >> (ANY ? kPushShiftIds : kPushShift)) + 1) * sizeof(unsigned), run));
    RB_CUDA_CHECK(cudaMemsetAsync(c->counters, 0, sizeof(int), run));
    const int per_sm = g_tuning.blocks_per_sm > 0 ? g_tuning.blocks_per_sm : occupancy(s, reinterpret_cast<const void*>(traverse_direct<ANY, ARITY>), kBlock);
    const int grid = std::min((num_rays + kBlock - 1) / kBlock, s.sm_count * per_sm);
    s.last_kernel = ANY ? (ARITY == 8 ? "traverse_direct<true, 8>" : "traverse_direct<true, 4>") : (ARITY == 8 ? "traverse_direct<false, 8>" : "traverse_direct<false, 4>");
    traverse_direct<ANY, ARITY><<<grid, kBlock, 0, run>>>(d_nodes, d_tris, caller_rays, push || copy_after ? c->d_hits : caller_hits, num_rays, c->counters,
                                                         g_tuning.refill_min, g_tuning.node_streak_min, records);
    RB_CUDA_CHECK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (ev[1]) RB_CUDA_CHECK(cudaEventRecord(ev[1], run));
    if (copy_after) RB_CUDA_CHECK(cudaMemcpyAsync(hits, c->d_hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyDeviceToHost, run));
    std::atomic<int> kernel_done{0};
    if (stage_out || ANY) RB_CUDA_CHECK(cudaLaunchHostFunc(run, [](void* p) { static_cast<std::atomic<int>*>(p)->store(1, std::memory_order_release); }, &kernel_done));
    if (stage_out) {                                         // chunks of 512 KB of records, each taken home as far as it has arrived
        const int chunk = 1 << 15;
        const bool aligned_out = (reinterpret_cast<uintptr_t>(hits) & 15) == 0 && g_tuning.host_stream_stores != 0;
        std::vector<int> pos;
        for (int first = 0; first < num_rays; first += chunk) pos.push_back(first);
        std::vector<CopyPool::Task> tasks;
        for (size_t k = 0; k < pos.size(); k++)
            tasks.push_back([&, k] {
                const int end = std::min(num_rays, int(k + 1) * chunk);
                std::atomic_thread_fence(std::memory_order_acquire);
                for (int i = pos[k]; i < end; i++) {
                    __m128i v = _mm_load_si128(reinterpret_cast<const __m128i*>(c->h_hits + i));
                    if (_mm_cvtsi128_si32(v) == kRecordArmed) {
                        if (!kernel_done.load(std::memory_order_acquire)) { pos[k] = i; return false; }
                        v = _mm_load_si128(reinterpret_cast<const __m128i*>(c->h_hits + i));      // the kernel is gone: everything it wrote is here
                        if (_mm_cvtsi128_si32(v) == kRecordArmed) { std::fprintf(stderr, "rodent_b200: record %d never arrived in the staging array\n", i); std::abort(); }
                    }
                    if (W > 1) {                             // record -> its lane of the hit packet
                        alignas(16) float rec[4];
                        _mm_store_si128(reinterpret_cast<__m128i*>(rec), v);
                        float* q = hit_packets + size_t(i / W) * (4 * W) + (i % W);
                        q[0] = rec[0]; q[W] = rec[1]; q[2 * W] = rec[2]; q[3 * W] = rec[3];
                    } else if (aligned_out) _mm_stream_si128(reinterpret_cast<__m128i*>(hits + i), v);
                    else _mm_storeu_si128(reinterpret_cast<__m128i*>(hits + i), v);
                    c->h_hits[i].tri_id = kRecordArmed;
                }
                pos[k] = end;
                _mm_sfence();
                return true;
            });
        CopyPool::get().run_all(tasks);
    }
    if (ANY) {                                               // chunks of 128 KB of ids, each written into the records as far as it has arrived
        const int chunk = 1 << 15;
        std::vector<int> pos;
        for (int first = 0; first < num_rays; first += chunk) pos.push_back(first);
        std::vector<CopyPool::Task> tasks;
        for (size_t k = 0; k < pos.size(); k++)
            tasks.push_back([&, k] {
                const int end = std::min(num_rays, int(k + 1) * chunk);
                std::atomic_thread_fence(std::memory_order_acquire);
                for (int i = pos[k]; i < end; i++) {
                    int32_t v = *const_cast<volatile int32_t*>(c->h_ids + i);
                    if (v == kRecordArmed) {
                        if (!kernel_done.load(std::memory_order_acquire)) { pos[k] = i; return false; }
                        v = *const_cast<volatile int32_t*>(c->h_ids + i);                          // the kernel is gone: everything it wrote is here
                        if (v == kRecordArmed) { std::fprintf(stderr, "rodent_b200: the id of ray %d never arrived\n", i); std::abort(); }
                    }
                    if (W > 1) std::memcpy(hit_packets + size_t(i / W) * (4 * W) + (i % W), &v, sizeof v);
                    else hits[i].tri_id = v;
                    c->h_ids[i] = kRecordArmed;
                }
                if (end == num_rays) for (int i = end; i < ((end + 3) & ~3); i++) c->h_ids[i] = kRecordArmed;   // the rest of the last float4
                pos[k] = end;
                return true;
            });
        CopyPool::get().run_all(tasks);
    }
    RB_CUDA_CHECK(cudaStreamSynchronize(run));      // (the kernel has seen every ray arrive, or the word behind the copy)
    if (ev[1]) {
        float ms = 0;
        RB_CUDA_CHECK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        std::fprintf(stderr, "direct call, %d rays: %.3f ms on the device\n", num_rays, ms);
        for (auto& e : ev) RB_CUDA_CHECK(cudaEventDestroy(e));
    }
    return true;
}

template <bool ANY, typename NodeT>
static void run_host_on(int dev, const NodeT* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int num_rays) {
    if (num_rays <= 0) return;
    DeviceState& s = device_state(dev);
    auto bvh = cached_bvh(s, nodes, tris);
    std::unique_lock<std::mutex> serial(g_pool_serial, std::defer_lock);
    if (g_tuning.mapping == 3) serial.lock();
    HostContext* c = acquire_host_context(s, size_t(num_rays));
    // pageable buffers are staged through pinned memory of this context (cudaMemcpyAsync on pageable memory is
    // synchronous and slow); g_tuning.host_staging = 0 hands them to the driver as they are
    const bool stage_in = g_tuning.host_staging && !is_pinned(rays), stage_out = g_tuning.host_staging && !is_pinned(hits);
    if (stage_in || stage_out) ensure_staging(c, size_t(num_rays));
    const Ray1* src = stage_in ? c->h_rays : rays;
    Hit1* dst = stage_out ? c->h_hits : hits;
    {
        const bool aligned = ((reinterpret_cast<uintptr_t>(bvh.first) | reinterpret_cast<uintptr_t>(bvh.second)) & 31) == 0;
        // The single launch needs its rays where a copy engine can take them: page-locked.  Pageable rays would have
        // to be staged before the launch (a running kernel must not wait for copies this process has yet to queue, see
        // run_host_direct), which gives away the overlap the pieces below have: those keep the pageable rays
        // (host_staged_direct = 2 forces the single launch for them as well).  Pageable RECORDS are no obstacle: they
        // are taken home from the staging array while the kernel runs, and an any-hit call touches the caller's
        // records from the host only (tri_id, as the ids arrive).
        const bool in_ok = stage_in ? g_tuning.host_staged_direct == 2 : is_pinned(rays, true);
        const bool out_ok = ANY ? true : stage_out ? g_tuning.host_staged_direct != 0 : is_pinned(hits, true);
        if (g_tuning.host_direct && g_tuning.mapping == 2 && g_tuning.wide_loads && aligned && in_ok && out_ok) {
            if (run_host_direct<ANY>(s, c, bvh.first, bvh.second, rays, hits, num_rays, stage_in, stage_out)) {
                release_host_context(s, c);
                return;
            }
        }
    }
    if (ANY) {  // occluded leaves t/u/v untouched: round-trip the caller's records
        RB_CUDA_CHECK(cudaMemcpyAsync(c->d_hits, hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyHostToDevice, c->streams[0]));
        RB_CUDA_CHECK(cudaStreamSynchronize(c->streams[0]));
    }
    // Pieces of decreasing size (k, k-1, ..., 1 parts of k(k+1)/2): the copy engine is the critical resource, and what
    // follows the last byte of input is one piece's traversal -- including its stragglers -- and its copy out, so that
    // last piece is kept small.
    c->rays_armed = false;                            // the copies below overwrite the slots run_host_direct keeps armed
    if (stage_out) c->hits_armed = false;
    const int pieces = std::max(1, std::min(g_tuning.host_chunks, 16));
    const bool ramp = g_tuning.host_ramp && pieces >= 3;      // weights 1, k-1, k-2, ..., 1: nothing runs before the first piece is in
    const int64_t parts = ramp ? int64_t(pieces - 1) * pieces / 2 + 1 : int64_t(pieces) * (pieces + 1) / 2;
    int k = 0;
    int piece_first[17], piece_n[17];
    for (int first = 0; first < num_rays; k++) {
        const int weight = ramp ? (k == 0 ? 1 : std::max(1, pieces - k)) : std::max(1, pieces - k);
        int n = int((int64_t(num_rays) * weight / parts + 3) & ~int64_t(3));
        n = std::max(n, 1 << 14);
        if (k >= pieces - 1 || n > num_rays - first) n = num_rays - first;
        cudaStream_t st = c->streams[k % 3];
        if (stage_in) CopyPool::get().parallel_copy(c->h_rays + first, rays + first, size_t(n) * sizeof(Ray1));
        RB_CUDA_CHECK(cudaMemcpyAsync(c->d_rays + first, src + first, size_t(n) * sizeof(Ray1), cudaMemcpyHostToDevice, st));
        launch<ANY>(s, bvh.first, bvh.second, c->d_rays + first, c->d_hits + first, n, st, c->counters + 8 * (k % 3));
        RB_CUDA_CHECK(cudaMemcpyAsync(dst + first, c->d_hits + first, size_t(n) * sizeof(Hit1), cudaMemcpyDeviceToHost, st));
        if (stage_out) {
            if (!c->piece_done[k]) RB_CUDA_CHECK(cudaEventCreateWithFlags(&c->piece_done[k], cudaEventDisableTiming));
            RB_CUDA_CHECK(cudaEventRecord(c->piece_done[k], st));
        }
        piece_first[k] = first; piece_n[k] = n;
        first += n;
    }
    if (stage_out) {
        for (int j = 0; j < k; j++) {            // drain piece by piece while the later ones are still in flight
            RB_CUDA_CHECK(cudaEventSynchronize(c->piece_done[j]));
            CopyPool::get().parallel_copy(hits + piece_first[j], c->h_hits + piece_first[j], size_t(piece_n[j]) * sizeof(Hit1));
        }
    }
    for (auto& st : c->streams) RB_CUDA_CHECK(cudaStreamSynchronize(st));
    release_host_context(s, c);
}

// Packet entry points: closest hit in copy-engine pieces, any hit on the staged single launch (run_host_direct, W = 4 or 8);
// with host_staged_direct = 0, copy in, one launch of the packet kernel, copy out.
template <bool ANY, typename NodeT, int W>
static void run_host_packets(const NodeT* nodes, const Tri4* tris, const void* rays, void* hits, int num_packets, bool hybrid) {
    if (num_packets <= 0) return;
    constexpr int ARITY = int(sizeof(NodeT::child) / sizeof(int32_t));
    DeviceState& s = device_state(g_host_dev);
    auto bvh = cached_bvh(s, nodes, tris);
    const int num_rays = num_packets * W;
    HostContext* c = acquire_host_context(s, size_t(num_rays));
    if (g_tuning.packet_order) {                 // the reference's packet order to the bit: copy in, one launch, copy out
        cudaStream_t st = c->streams[0];
        c->rays_armed = false;
        RB_CUDA_CHECK(cudaMemcpyAsync(c->d_rays, rays, size_t(num_rays) * sizeof(Ray1), cudaMemcpyHostToDevice, st));
        if (ANY) RB_CUDA_CHECK(cudaMemcpyAsync(c->d_hits, hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyHostToDevice, st));   // t, u, v stay the caller's
        traverse_packets_ordered<ANY, ARITY, W><<<(num_packets + 63) / 64, 64, 0, st>>>(bvh.first, bvh.second, reinterpret_cast<const float*>(c->d_rays),
                                                                                        reinterpret_cast<float*>(c->d_hits), num_packets, hybrid ? 1 : 0);
        RB_CUDA_CHECK(cudaGetLastError());
        g_launches.fetch_add(1, std::memory_order_relaxed);
        s.last_kernel = "traverse_packets_ordered";
        RB_CUDA_CHECK(cudaMemcpyAsync(hits, c->d_hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyDeviceToHost, st));
        RB_CUDA_CHECK(cudaStreamSynchronize(st));
        release_host_context(s, c);
        return;
    }
    if (g_tuning.host_direct && g_tuning.mapping == 2 && g_tuning.host_staged_direct != 0) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(bvh.first) | reinterpret_cast<uintptr_t>(bvh.second)) & 31) == 0;
        ensure_staging(c, size_t(num_rays));
        if constexpr (ANY) {
            // the single launch of the single-ray calls behind the staging of the rays (transposed by the helper
            // threads); the triangle ids come home as a dense array and are written into the hit packets
            if (g_tuning.wide_loads && aligned && run_host_direct<ANY, NodeT, W>(s, c, bvh.first, bvh.second, rays, hits, num_rays)) {
                release_host_context(s, c);
                return;
            }
        } else {
            // The copy-engine pieces of the single-ray calls (run_host_on), with the helper threads transposing packets
            // into rays on the way in and records into hit packets on the way out: the traversal of one piece runs
            // under the staging of the next.
            c->rays_armed = false; c->hits_armed = false;
            const int pieces = std::max(1, std::min(g_tuning.host_chunks, 16));
            const int64_t parts = int64_t(pieces) * (pieces + 1) / 2;
            int k = 0, piece_first[17], piece_n[17];
            for (int first = 0; first < num_rays; k++) {
                int n = int((int64_t(num_rays) * std::max(1, pieces - k) / parts + 31) & ~int64_t(31));
                n = std::max(n, 1 << 14);
                if (k >= pieces - 1 || n > num_rays - first) n = num_rays - first;
                cudaStream_t st = c->streams[k % 3];
                stage_packets<W>(c, static_cast<const float*>(rays), first, n);
                RB_CUDA_CHECK(cudaMemcpyAsync(c->d_rays + first, c->h_rays + first, size_t(n) * sizeof(Ray1), cudaMemcpyHostToDevice, st));
                launch<ANY>(s, bvh.first, bvh.second, c->d_rays + first, c->d_hits + first, n, st, c->counters + 8 * (k % 3));
                RB_CUDA_CHECK(cudaMemcpyAsync(c->h_hits + first, c->d_hits + first, size_t(n) * sizeof(Hit1), cudaMemcpyDeviceToHost, st));
                if (!c->piece_done[k]) RB_CUDA_CHECK(cudaEventCreateWithFlags(&c->piece_done[k], cudaEventDisableTiming));
                RB_CUDA_CHECK(cudaEventRecord(c->piece_done[k], st));
                piece_first[k] = first; piece_n[k] = n;
                first += n;
            }
            for (int j = 0; j < k; j++) {
                RB_CUDA_CHECK(cudaEventSynchronize(c->piece_done[j]));
                unstage_packets<W>(c, static_cast<float*>(hits), piece_first[j], piece_n[j]);
            }
            for (auto& st : c->streams) RB_CUDA_CHECK(cudaStreamSynchronize(st));
            release_host_context(s, c);
            return;
        }
    }
    cudaStream_t st = c->streams[0];
    int* counter = c->counters;
    c->rays_armed = false;
    RB_CUDA_CHECK(cudaMemcpyAsync(c->d_rays, rays, size_t(num_rays) * sizeof(Ray1), cudaMemcpyHostToDevice, st));     // a packet is W * 32 bytes
    if (ANY) RB_CUDA_CHECK(cudaMemcpyAsync(c->d_hits, hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyHostToDevice, st)); // t/u/v stay the caller's
    RB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
    const int grid = std::min((num_rays + kBlock - 1) / kBlock,
                              s.sm_count * occupancy(s, reinterpret_cast<const void*>(traverse_packets_vote<ANY, ARITY, W>), kBlock));
    traverse_packets_vote<ANY, ARITY, W><<<grid, kBlock, 0, st>>>(bvh.first, bvh.second, reinterpret_cast<const float*>(c->d_rays),
                                                                  reinterpret_cast<float*>(c->d_hits), num_rays, counter,
                                                                  g_tuning.refill_min, g_tuning.node_streak_min);
    RB_CUDA_CHECK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    RB_CUDA_CHECK(cudaMemcpyAsync(hits, c->d_hits, size_t(num_rays) * sizeof(Hit1), cudaMemcpyDeviceToHost, st));
    RB_CUDA_CHECK(cudaStreamSynchronize(st));
    release_host_context(s, c);
}

}  // namespace rb200

using namespace rb200;

extern "C" {

void cuda_intersect_single_ray1_bvh8_tri4(int32_t dev, const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<false>(dev, nodes, tris, rays, hits, num_rays);
}
void cuda_occluded_single_ray1_bvh8_tri4(int32_t dev, const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<true>(dev, nodes, tris, rays, hits, num_rays);
}
void cuda_intersect_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<false>(dev, nodes, tris, rays, hits, num_rays);
}
void cuda_occluded_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<true>(dev, nodes, tris, rays, hits, num_rays);
}
void cuda_intersect_single_ray1_bvh8_tri4_async(int32_t dev, const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits,
                                                int32_t num_rays, void* stream, int32_t* work_counter) {
    launch<false>(device_state(dev), nodes, tris, rays, hits, num_rays, (cudaStream_t)stream, work_counter);
}
void cuda_occluded_single_ray1_bvh8_tri4_async(int32_t dev, const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits,
                                               int32_t num_rays, void* stream, int32_t* work_counter) {
    launch<true>(device_state(dev), nodes, tris, rays, hits, num_rays, (cudaStream_t)stream, work_counter);
}

void b200_intersect_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    run_host<false>(nodes, tris, rays, hits, num_packets);
}
void b200_occluded_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    run_host<true>(nodes, tris, rays, hits, num_packets);
}
void cuda_intersect_single_ray1_bvh4_tri4(int32_t dev, const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<false>(dev, nodes, tris, rays, hits, num_rays);
}
void cuda_occluded_single_ray1_bvh4_tri4(int32_t dev, const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    run_sync<true>(dev, nodes, tris, rays, hits, num_rays);
}
void b200_intersect_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    run_host<false>(nodes, tris, rays, hits, num_packets);
}
void b200_occluded_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    run_host<true>(nodes, tris, rays, hits, num_packets);
}
#define RB_PACKET_API(kind, W, B, HYBRID)                                                                                       \
    void b200_intersect_##kind##_ray##W##_bvh##B##_tri4(const Node##B* n, const Tri4* t, const Ray##W* r, Hit##W* h, int32_t k) { \
        run_host_packets<false, Node##B, W>(n, t, r, h, k, HYBRID);                                                             \
    }                                                                                                                           \
    void b200_occluded_##kind##_ray##W##_bvh##B##_tri4(const Node##B* n, const Tri4* t, const Ray##W* r, Hit##W* h, int32_t k) {  \
        run_host_packets<true, Node##B, W>(n, t, r, h, k, HYBRID);                                                              \
    }
RB_PACKET_API(packet, 4, 4, false) RB_PACKET_API(packet, 8, 4, false) RB_PACKET_API(packet, 4, 8, false) RB_PACKET_API(packet, 8, 8, false)
RB_PACKET_API(hybrid, 4, 4, true) RB_PACKET_API(hybrid, 8, 4, true) RB_PACKET_API(hybrid, 4, 8, true) RB_PACKET_API(hybrid, 8, 8, true)
#undef RB_PACKET_API

void rodent_b200_set_packet_order(int32_t on) { g_tuning.packet_order = on != 0; }

void rodent_b200_forget_bvh(const void* nodes, const Tri4* tris) {
    DeviceState& s = device_state(g_host_dev);
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = s.bvh_cache.find(std::make_pair((const void*)nodes, (const void*)tris));
    if (it == s.bvh_cache.end()) return;
    RB_CUDA_CHECK(cudaDeviceSynchronize());
    RB_CUDA_CHECK(cudaFree(it->second.d_nodes));
    RB_CUDA_CHECK(cudaFree(it->second.d_tris));
    s.bvh_cache.erase(it);
}
// uploads / re-uploads (content changed under a known address) / live copies of the host-pointer BVH cache
void rodent_b200_bvh_cache_stats(int64_t out[3]) {
    DeviceState& s = device_state(g_host_dev);
    std::lock_guard<std::mutex> lock(g_mutex);
    out[0] = s.bvh_uploads; out[1] = s.bvh_reuploads; out[2] = int64_t(s.bvh_cache.size());
}

int32_t rodent_b200_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
void rodent_b200_set_device(int32_t dev) { device_state(dev); g_host_dev = dev; g_host_devs.clear(); }
void rodent_b200_set_devices(const int32_t* devs, int32_t num_devs) {
    g_host_devs.assign(devs, devs + std::max(num_devs, 0));
    for (int d : g_host_devs) device_state(d);
    if (!g_host_devs.empty()) { g_host_dev = g_host_devs[0]; device_state(g_host_dev); }
}
void* rodent_b200_alloc_device(int32_t dev, size_t bytes) {
    device_state(dev);
    void* p = nullptr;
    RB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return p;
}
void rodent_b200_free_device(int32_t dev, void* ptr) { device_state(dev); RB_CUDA_CHECK(cudaFree(ptr)); }
void* rodent_b200_alloc_host(size_t bytes) {
    void* p = nullptr;
    RB_CUDA_CHECK(cudaMallocHost(&p, std::max<size_t>(bytes, 16)));
    return p;
}
void rodent_b200_free_host(void* ptr) { RB_CUDA_CHECK(cudaFreeHost(ptr)); }
int32_t rodent_b200_pin_host(void* ptr, size_t bytes) {
    if (ptr == nullptr || bytes == 0) return -1;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}
// The host side of the pageable-buffer path without a device: the helper-thread pool and the staging copies (plain and
// non-temporal) over odd sizes and alignments, tasks that ask to be run again.  0 when everything checks out.
int32_t rodent_b200_selftest_host_copies(void) {
    const size_t n = (size_t(5) << 20) + 777;
    std::vector<unsigned char> src(n + 64), dst(n + 64), want(n + 64);
    for (size_t i = 0; i < src.size(); i++) src[i] = static_cast<unsigned char>(i * 2654435761u >> 24);
    int bad = 0;
    const int saved_parts = g_tuning.host_copy_parts, saved_stream = g_tuning.host_stream_stores;
    for (int stream = 0; stream < 2; stream++)
        for (int parts : {1, 3, 4, 9})
            for (size_t off : {size_t(0), size_t(1), size_t(13), size_t(16)})
                for (size_t bytes : {size_t(0), size_t(1), size_t(63), size_t(64), size_t(4097), n - 16}) {
                    g_tuning.host_stream_stores = stream; g_tuning.host_copy_parts = parts;
                    std::fill(dst.begin(), dst.end(), static_cast<unsigned char>(0xA5));
                    want = dst;
                    std::memcpy(want.data() + off, src.data() + 3, bytes);
                    CopyPool::get().parallel_copy(dst.data() + off, src.data() + 3, bytes);
                    bad += dst != want;
                    std::fill(dst.begin(), dst.end(), static_cast<unsigned char>(0xA5));
                    if (stream) stream_copy(dst.data() + off, src.data() + 3, bytes); else std::memcpy(dst.data() + off, src.data() + 3, bytes);
                    bad += dst != want;
                }
    g_tuning.host_copy_parts = saved_parts; g_tuning.host_stream_stores = saved_stream;
    // tasks that are not finished the first few times they run, from two callers at once
    auto batch = [&bad] {
        std::vector<int> runs(40, 0);
        std::vector<CopyPool::Task> tasks;
        for (size_t k = 0; k < runs.size(); k++) tasks.push_back([&runs, k] { return ++runs[k] > int(k % 4); });
        CopyPool::get().run_all(tasks);
        for (size_t k = 0; k < runs.size(); k++) if (runs[k] != int(k % 4) + 1) __atomic_fetch_add(&bad, 1, __ATOMIC_RELAXED);
    };
    std::thread other(batch);
    batch();
    other.join();
    return bad;
}
int32_t rodent_b200_unpin_host(void* ptr) {
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}
void rodent_b200_copy_to_device(int32_t dev, void* dst, const void* src, size_t bytes) {
    device_state(dev);
    RB_CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
}
void rodent_b200_copy_to_host(int32_t dev, void* dst, const void* src, size_t bytes) {
    device_state(dev);
    RB_CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
}
// ---- peer memory (one process per GPU) -------------------------------------------------------------------------
// A device allocation of another process mapped into this one (CUDA IPC; the devices are NVLink / NVSwitch peers): what
// lets a rank's traversal kernel write its hit records straight into the gathering rank's HBM instead of into local memory
// and through a collective afterwards.  Failures are reported, not fatal: the caller falls back to a gather.
int32_t rodent_b200_ipc_export(int32_t dev, const void* device_ptr, void* handle_out) {
    device_state(dev);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t h;
    const cudaError_t err = cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr));
    if (err != cudaSuccess) { std::fprintf(stderr, "rodent_b200: cudaIpcGetMemHandle: %s\n", cudaGetErrorString(err)); cudaGetLastError(); return 0; }
    std::memcpy(handle_out, &h, sizeof h);
    return 1;
}
void* rodent_b200_ipc_open(int32_t dev, const void* handle) {
    device_state(dev);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) { std::fprintf(stderr, "rodent_b200: cudaIpcOpenMemHandle: %s\n", cudaGetErrorString(err)); cudaGetLastError(); return nullptr; }
    return p;
}
void rodent_b200_ipc_close(int32_t dev, void* ptr) { device_state(dev); if (ptr) RB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr)); }

void rodent_b200_sync(int32_t dev) { device_state(dev); RB_CUDA_CHECK(cudaDeviceSynchronize()); }
double rodent_b200_last_kernel_ms(int32_t dev) { return device_state(dev).last_ms; }
const char* rodent_b200_last_kernel_name(int32_t dev) { return device_state(dev).last_kernel; }
int64_t rodent_b200_launch_count(void) { return g_launches.load(); }
void rodent_b200_count_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }   // for the other translation units
const char* rodent_b200_version(void) { return "rodent_b200 0.1 sm_100a"; }

void rodent_b200_render_tune(const char* key, int32_t value);   // render.cu

// Tuning knobs for experiments (not part of the drop-in surface).  Values are clamped to what the kernels can take: a
// streak or refill threshold below 1 would never leave its loop (or elect lane -1), so thresholds live in [1, 33]
// (33 = never), block counts and variants in their implemented ranges.
void rodent_b200_tune(const char* key, int32_t value) {
    auto clamp = [](int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); };
    if (!std::strcmp(key, "persistent")) g_tuning.persistent = value != 0;
    else if (!std::strcmp(key, "mapping")) g_tuning.mapping = clamp(value, 0, 4);
    else if (!std::strcmp(key, "quad_refill_below")) g_tuning.quad_refill_below = clamp(value, 1, 9);
    else if (!std::strcmp(key, "refill_below")) g_tuning.refill_below = clamp(value, 1, 33);
    else if (!std::strcmp(key, "refill_min")) g_tuning.refill_min = clamp(value, 1, 32);
    else if (!std::strcmp(key, "vote_min_blocks")) g_tuning.vote_min_blocks = clamp(value, 4, 6);
    else if (!std::strcmp(key, "node_streak_min")) g_tuning.node_streak_min = clamp(value, 1, 33);
    else if (!std::strcmp(key, "blocks_per_sm")) g_tuning.blocks_per_sm = clamp(value, 0, 32);
    else if (!std::strcmp(key, "pool_refill_min")) g_tuning.pool_refill_min = clamp(value, 1, 64);
    else if (!std::strcmp(key, "pool_prefetch")) g_tuning.pool_prefetch = value != 0;
    else if (!std::strcmp(key, "host_chunks")) g_tuning.host_chunks = clamp(value, 1, 16);
    else if (!std::strcmp(key, "packet_order")) g_tuning.packet_order = value != 0;
    else if (!std::strcmp(key, "host_staging")) g_tuning.host_staging = value != 0;
    else if (!std::strcmp(key, "host_stream_stores")) g_tuning.host_stream_stores = value != 0;
    else if (!std::strcmp(key, "host_copy_parts")) g_tuning.host_copy_parts = clamp(value, 1, 17);
    else if (!std::strcmp(key, "host_ramp")) g_tuning.host_ramp = value != 0;
    else if (!std::strcmp(key, "host_direct")) g_tuning.host_direct = value != 0;
    else if (!std::strcmp(key, "host_direct_rays")) g_tuning.host_direct_rays = value != 0;
    else if (!std::strcmp(key, "host_staged_direct")) g_tuning.host_staged_direct = clamp(value, 0, 2);
    else if (!std::strcmp(key, "host_direct_push")) g_tuning.host_direct_push = clamp(value, 0, 2);
    else if (!std::strcmp(key, "host_trace")) g_tuning.host_trace = value != 0;
    else if (!std::strcmp(key, "bvh2_min_blocks")) g_tuning.bvh2_min_blocks = clamp(value, 8, 12);
    else if (!std::strcmp(key, "wide_loads")) g_tuning.wide_loads = value != 0;
    else if (!std::strcmp(key, "vote_smem_depth")) g_tuning.vote_smem_depth = clamp(value, 12, 24);
    else if (!std::strcmp(key, "bvh2_streak_min")) g_tuning.bvh2_streak_min = clamp(value, 1, 33);
    else if (!std::strncmp(key, "render_", 7)) rodent_b200_render_tune(key, value);
    else { std::fprintf(stderr, "rodent_b200_tune: unknown key '%s'\n", key); std::abort(); }
}

}  // extern "C"
