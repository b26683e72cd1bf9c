// Ray-pool BVH8/Tri4 traversal for sm_100a: every warp owns a pool of 64 rays in shared memory and, for each
// step, COMPACTS up to 32 rays that need the same kind of step onto its lanes.
//
// traverse_sched.cuh binds one ray to one lane: when the warp votes for a node step, the lanes whose ray needs
// a Tri4 step (or has finished) idle -- 17.8 of 32 lanes were busy in the node step and 11.4 in the leaf step on
// the incoherent Sponza set (profiles/r01_traverse_vote_blocks.txt).  Here a ray is not bound to a lane.  Its
// state (ray constants, register-top of the stack, hit, 16 stack levels) lives in a slot of the warp's pool;
// each iteration the warp
//   1. reads the phase of all 64 slots (two per lane) and ballots them,
//   2. picks the phase with more candidates (node step / Tri4 step),
//   3. ranks the candidates (ballot + popc) and hands candidate #i to lane i through a 32-entry scratch array
//      -- the ballot/shuffle ray compaction of the Aila-Laine scheme, done per step instead of per refill,
//   4. every lane loads the state its step needs, runs the SAME RayWalker step as the vote-scheduled kernel
//      (node_step / leaf_step of traverse_sched.cuh: the reference's transitions, untouched), stores what
//      changed and the ray's next phase.
// Empty slots are refilled from the global ray counter 24..32 at a time (one atomicAdd per refill).
// All of it is warp-private: no block barrier, no lock, only __syncwarp().
#pragma once

#include "traverse_sched.cuh"

namespace rb200 {

constexpr int kPoolSlots = 64;                 // rays per warp
constexpr int kPoolDepth = 16;                 // stack levels per ray in shared memory; deeper ones go to global memory
constexpr int kPoolOverflow = kStackSize - kPoolDepth;

enum PoolPhase : int { kPhaseEmpty = 0, kPhaseNode = 1, kPhaseLeaf = 2 };

// One warp's pool.  Structure of arrays over the slot index: a lane reads the fields of whichever slot it was handed.
struct alignas(16) RayPool {
    float ox[kPoolSlots], oy[kPoolSlots], oz[kPoolSlots], dx[kPoolSlots], dy[kPoolSlots], dz[kPoolSlots];
    float idx[kPoolSlots], idy[kPoolSlots], idz[kPoolSlots], iox[kPoolSlots], ioy[kPoolSlots], ioz[kPoolSlots];
    float tmin[kPoolSlots], tmax[kPoolSlots];
    int flags[kPoolSlots];                     // bit 0..2: direction component > 0 (x, y, z); bit 3: RaySetup::degenerate
    int top_node[kPoolSlots]; float top_t[kPoolSlots]; int ptr[kPoolSlots]; int leaf[kPoolSlots];
    int prim[kPoolSlots], geom[kPoolSlots]; float hu[kPoolSlots], hv[kPoolSlots];
    int ray_idx[kPoolSlots];
    int phase[kPoolSlots];
    int scratch[32];                           // candidate rank -> slot
    StackEntry stack[kPoolDepth][kPoolSlots];
};

template <bool ANY>
using PoolWalker = RayWalker<ANY, kPoolDepth, kPoolSlots>;

template <bool ANY>
__device__ __forceinline__ void pool_set_octant(PoolWalker<ANY>& w, int flags) {
    const int px = flags & 1, py = (flags >> 1) & 1, pz = (flags >> 2) & 1;
    w.ray.near_x = 2 * (1 - px); w.ray.far_x = 2 * px;
    w.ray.near_y = 4 + 2 * (1 - py); w.ray.far_y = 4 + 2 * py;
    w.ray.near_z = 8 + 2 * (1 - pz); w.ray.far_z = 8 + 2 * pz;
    w.ray.degenerate = (flags & 8) != 0;
}

// A ray waits in the pool between two of its steps: ask L1 meanwhile for what the next step will read (the two
// 128-byte lines of the Node8, or of the leaf's first Tri4).
template <bool ANY>
__device__ __forceinline__ void pool_prefetch(const PoolWalker<ANY>& w, const Node8* __restrict__ nodes, const Tri4* __restrict__ tris, bool enable) {
    if (!enable || w.top_node == 0) return;
    const char* p = w.leaf >= 0 ? reinterpret_cast<const char*>(tris + w.leaf)
                  : w.top_node > 0 ? reinterpret_cast<const char*>(nodes + (w.top_node - 1))
                                   : reinterpret_cast<const char*>(tris + ~w.top_node);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 128));
}

// Phase of a ray after a step (or after begin): the slot is free again when the ray has finished.
template <bool ANY>
__device__ __forceinline__ int pool_phase_of(const PoolWalker<ANY>& w) {
    return w.finished() ? kPhaseEmpty : (w.wants_node() ? kPhaseNode : kPhaseLeaf);
}

// Hands candidate #i (slots 0..31 first, then 32..63) to lane i.  c0 / c1: ballots over the lanes' two home slots.
__device__ __forceinline__ int pool_assign(RayPool& pool, unsigned c0, unsigned c1, unsigned lane) {
    const unsigned lt = lanemask_lt();
    const int n0 = __popc(c0);
    if (c0 & (1u << lane)) pool.scratch[__popc(c0 & lt)] = int(lane);
    const int r1 = n0 + __popc(c1 & lt);
    if ((c1 & (1u << lane)) && r1 < 32) pool.scratch[r1] = int(lane) + 32;
    __syncwarp();
    const int count = min(32, n0 + __popc(c1));
    const int slot = int(lane) < count ? pool.scratch[lane] : -1;
    __syncwarp();
    return slot;
}

//   fetch(i, r0, r1)  loads ray i (origin+tmin, direction+tmax)
//   sink(i, hit)      consumes the finished ray's record
// `overflow`: this warp's kPoolSlots * kPoolOverflow stack entries in global memory.
template <bool ANY, bool WANT_GEOM, typename Fetch, typename Sink>
__device__ __forceinline__ void traverse_pooled(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris, RayPool& pool,
                                                StackEntry* overflow, int num_rays, int* __restrict__ work_counter, int refill_min,
                                                bool prefetch, Fetch fetch, Sink sink) {
    const unsigned lane = lane_id();
    pool.phase[lane] = kPhaseEmpty;
    pool.phase[lane + 32] = kPhaseEmpty;
    bool drained = false;
    for (;;) {
        __syncwarp();
        const int ph0 = pool.phase[lane], ph1 = pool.phase[lane + 32];
        const unsigned e0 = __ballot_sync(0xffffffffu, ph0 == kPhaseEmpty), e1 = __ballot_sync(0xffffffffu, ph1 == kPhaseEmpty);
        const unsigned n0 = __ballot_sync(0xffffffffu, ph0 == kPhaseNode), n1 = __ballot_sync(0xffffffffu, ph1 == kPhaseNode);
        const unsigned l0 = ~(e0 | n0), l1 = ~(e1 | n1);
        const int num_empty = __popc(e0) + __popc(e1), num_node = __popc(n0) + __popc(n1), num_leaf = 64 - num_empty - num_node;

        // ---- refill: up to 32 fresh rays into empty slots, one atomicAdd ----
        if (!drained && (num_empty >= refill_min || num_node + num_leaf == 0)) {
            const int slot = pool_assign(pool, e0, e1, lane);
            const int take = min(32, num_empty);
            int base = 0;
            if (lane == 0) base = atomicAdd(work_counter, take);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int i = base + int(lane);
            if (slot >= 0 && i < num_rays) {
                float4 r0, r1;
                fetch(i, r0, r1);
                PoolWalker<ANY> w;
                w.st.smem = &pool.stack[0][slot];
                w.st.overflow = overflow + slot * kPoolOverflow;
                w.begin(r0, r1);
                if (w.finished()) {
                    sink(i, w.hit);                                  // culled at once (tmin > tmax): the slot stays empty
                } else {
                    pool.ox[slot] = w.ray.ox; pool.oy[slot] = w.ray.oy; pool.oz[slot] = w.ray.oz;
                    pool.dx[slot] = w.ray.dx; pool.dy[slot] = w.ray.dy; pool.dz[slot] = w.ray.dz;
                    pool.idx[slot] = w.ray.idx; pool.idy[slot] = w.ray.idy; pool.idz[slot] = w.ray.idz;
                    pool.iox[slot] = w.ray.iox; pool.ioy[slot] = w.ray.ioy; pool.ioz[slot] = w.ray.ioz;
                    pool.tmin[slot] = w.ray.tmin; pool.tmax[slot] = w.tmax;
                    pool.flags[slot] = (w.ray.dx > 0.0f ? 1 : 0) | (w.ray.dy > 0.0f ? 2 : 0) | (w.ray.dz > 0.0f ? 4 : 0) | (w.ray.degenerate ? 8 : 0);
                    pool.top_node[slot] = w.top_node; pool.top_t[slot] = w.top_t; pool.ptr[slot] = w.ptr; pool.leaf[slot] = -1;
                    pool.prim[slot] = -1; pool.geom[slot] = -1; pool.hu[slot] = 0.0f; pool.hv[slot] = 0.0f;
                    pool.ray_idx[slot] = i;
                    pool.phase[slot] = pool_phase_of(w);
                }
            }
            if (base + take >= num_rays) drained = true;
            continue;
        }
        if (num_node + num_leaf == 0) {
            if (drained) break;
            continue;
        }

        if (num_node >= num_leaf) {
            // ---- node step for up to 32 rays ----
            const int slot = pool_assign(pool, n0, n1, lane);
            const bool any_degenerate = __ballot_sync(0xffffffffu, slot >= 0 && (pool.flags[max(slot, 0)] & 8)) != 0;
            if (slot >= 0) {
                PoolWalker<ANY> w;
                w.st.smem = &pool.stack[0][slot];
                w.st.overflow = overflow + slot * kPoolOverflow;
                w.ray.idx = pool.idx[slot]; w.ray.idy = pool.idy[slot]; w.ray.idz = pool.idz[slot];
                w.ray.iox = pool.iox[slot]; w.ray.ioy = pool.ioy[slot]; w.ray.ioz = pool.ioz[slot];
                w.ray.tmin = pool.tmin[slot]; w.tmax = pool.tmax[slot];
                pool_set_octant(w, pool.flags[slot]);
                w.top_node = pool.top_node[slot]; w.top_t = pool.top_t[slot]; w.ptr = pool.ptr[slot]; w.leaf = -1;
                if (any_degenerate) w.template node_step<true>(nodes);
                else                w.template node_step<false>(nodes);
                pool.top_node[slot] = w.top_node; pool.top_t[slot] = w.top_t; pool.ptr[slot] = w.ptr;
                const int next = pool_phase_of(w);
                pool_prefetch(w, nodes, tris, prefetch);
                if (next == kPhaseEmpty) {
                    HitRecord h;
                    h.prim = pool.prim[slot]; h.geom = pool.geom[slot]; h.t = w.tmax; h.u = pool.hu[slot]; h.v = pool.hv[slot];
                    sink(pool.ray_idx[slot], h);
                }
                pool.phase[slot] = next;
            }
        } else {
            // ---- Tri4 step for up to 32 rays ----
            const int slot = pool_assign(pool, l0, l1, lane);
            if (slot >= 0) {
                PoolWalker<ANY> w;
                w.st.smem = &pool.stack[0][slot];
                w.st.overflow = overflow + slot * kPoolOverflow;
                w.ray.ox = pool.ox[slot]; w.ray.oy = pool.oy[slot]; w.ray.oz = pool.oz[slot];
                w.ray.dx = pool.dx[slot]; w.ray.dy = pool.dy[slot]; w.ray.dz = pool.dz[slot];
                w.ray.tmin = pool.tmin[slot]; w.tmax = pool.tmax[slot];
                w.top_node = pool.top_node[slot]; w.top_t = pool.top_t[slot]; w.ptr = pool.ptr[slot]; w.leaf = pool.leaf[slot];
                w.hit.prim = pool.prim[slot]; w.hit.geom = pool.geom[slot]; w.hit.t = w.tmax; w.hit.u = pool.hu[slot]; w.hit.v = pool.hv[slot];
                w.template leaf_step<WANT_GEOM>(tris);
                pool.tmax[slot] = ANY ? pool.tmax[slot] : w.tmax;
                pool.top_node[slot] = w.top_node; pool.top_t[slot] = w.top_t; pool.ptr[slot] = w.ptr; pool.leaf[slot] = w.leaf;
                pool.prim[slot] = w.hit.prim; pool.geom[slot] = w.hit.geom; pool.hu[slot] = w.hit.u; pool.hv[slot] = w.hit.v;
                const int next = pool_phase_of(w);
                pool_prefetch(w, nodes, tris, prefetch);
                if (next == kPhaseEmpty) sink(pool.ray_idx[slot], w.hit);
                pool.phase[slot] = next;
            }
        }
    }
}

}  // namespace rb200
