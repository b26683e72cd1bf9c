// The reference's GPU single-ray traversal on its own BVH2 / Tri1 layout, for sm_100a.
//
// Semantics: gpu_traverse_single_helper (src/traversal/mapping_gpu.impala:94-178) with the arity-2 branch, the NVVM
// min/max set (make_nvvm_min_max, :74-85: fminf / fmaxf plus vmin / vmax on the float bits), the unordered slab test
// (src/traversal/intersection.impala:194-208), the single-triangle leaves of make_gpu_bvh2_tri1 (:19-69, n = cross(e1, e2)
// per test) and the register-top stack of src/traversal/stack.impala:52-123 (whose keys this path never reads).
// The reference runs it as one thread per ray, block 64, one launch over the whole stream (:182-203).  Here the same
// per-ray state machine runs under the vote scheduler of traverse_sched.cuh: the warp takes ONE straight-line step per
// iteration -- a node step (2 x LDG.256, two slab tests, at most one push) or a triangle step (3 x LDG.128, one
// Moeller-Trumbore test) -- chosen by majority, idle lanes refilled together from a global counter.
// NaN handling needs no special path: the reference is NVIDIA code, and these are the same instructions.
#pragma once

#include "traverse.cuh"

namespace rb200 {

// A Node2 in 32 bytes for the renderer's stream kernels: the twelve bounds as 16-bit steps of the scene's extent (lower
// bounds rounded down, upper bounds up, one step of margin each way), the two child ids as they are.  One 256-bit load
// per node instead of two, twice as many nodes per cache line -- the kernels wait on L2 hits for half of their stall
// samples (profiles/r02_render_kernels_ncu.txt).  Boxes only grow, so a ray finds the same triangles; it may enter a few
// more nodes.  The bench_traversal entry points keep the reference's 64-byte Node2 (their records are compared bit for bit).
struct Node2q { uint16_t b[12]; int32_t child[2]; };
static_assert(sizeof(Node2q) == 32, "Node2q is one 256-bit load");
struct QuantGrid { float origin[3], step[3]; };       // value of code q on axis a: origin[a] + q * step[a]

// Node ids only (the GPU path pushes undef keys): first SMEM_DEPTH levels in shared memory, [level][thread].
template <int SMEM_DEPTH, int BLOCK>
struct IdStack {
    int* smem;
    int* overflow;
    __device__ __forceinline__ int load(int i) const { return i < SMEM_DEPTH ? smem[i * BLOCK] : overflow[i - SMEM_DEPTH]; }
    __device__ __forceinline__ void store(int i, int v) {
        if (i < SMEM_DEPTH) smem[i * BLOCK] = v;
        else overflow[i - SMEM_DEPTH] = v;
    }
};

// QUANT: the node array holds Node2q; `qa`, `qb` are the ray's slab coefficients on the quantisation grid.
template <bool ANY, int SMEM_DEPTH, int BLOCK, bool FMA = false, bool QUANT = false>
struct Bvh2Walker {
    RaySetup ray;
    float qax, qay, qaz, qbx, qby, qbz;          // t = qa * code + qb  (= inv_dir * (origin + code * step) + inv_org)
    float tmax;
    int top, ptr;
    int leaf;                       // next Tri1 of the leaf being tested, -1 when not inside a leaf
    HitRecord hit;
    IdStack<SMEM_DEPTH, BLOCK> st;

    __device__ __forceinline__ void pop() { top = st.load(ptr); --ptr; }

    __device__ __forceinline__ void begin(float4 r0, float4 r1) {
        ray.template init<1>(r0, r1);                                      // make_gpu_ray1 + make_ray
        tmax = r1.w;
        hit.prim = -1; hit.geom = -1; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;   // empty_hit
        st.store(0, 0); ptr = 0; top = 1; leaf = -1;                       // stack.push(1, undef) on the empty stack, :103
    }
    // A 16-bit code is turned into a float without a conversion instruction: 0x4B000000 | q is the float 2^23 + q, and
    // the 2^23 is folded into the ray's offset, qb = inv_dir * origin + inv_org - qa * 2^23 (its rounding is half a grid
    // step at most, inside the two steps of margin the quantisation leaves).  One PRMT and one FMA per plane.
    __device__ __forceinline__ void set_grid(const QuantGrid& g) {
        qax = ray.idx * g.step[0]; qay = ray.idy * g.step[1]; qaz = ray.idz * g.step[2];
        qbx = __fmaf_rn(-qax, 8388608.0f, __fmaf_rn(ray.idx, g.origin[0], ray.iox));
        qby = __fmaf_rn(-qay, 8388608.0f, __fmaf_rn(ray.idy, g.origin[1], ray.ioy));
        qbz = __fmaf_rn(-qaz, 8388608.0f, __fmaf_rn(ray.idz, g.origin[2], ray.ioz));
    }
    // the same test on a quantised box: lo / hi codes of the three axes packed two to a word
    __device__ __forceinline__ bool hit_box_q(unsigned wx, unsigned wy, unsigned wz, float& tentry) const {
        auto lo = [](unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); };     // 2^23 + low half
        auto hi = [](unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); };     // 2^23 + high half
        const float t0x = __fmaf_rn(qax, lo(wx), qbx), t1x = __fmaf_rn(qax, hi(wx), qbx);
        const float t0y = __fmaf_rn(qay, lo(wy), qby), t1y = __fmaf_rn(qay, hi(wy), qby);
        const float t0z = __fmaf_rn(qaz, lo(wz), qbz), t1z = __fmaf_rn(qaz, hi(wz), qbz);
        const int zmin = max(min(__float_as_int(t0z), __float_as_int(t1z)), __float_as_int(ray.tmin));
        const int zmax = min(max(__float_as_int(t0z), __float_as_int(t1z)), __float_as_int(tmax));
        tentry = __int_as_float(__vimax3_s32(__float_as_int(fminf(t0x, t1x)), __float_as_int(fminf(t0y, t1y)), zmin));
        const float texit = __int_as_float(__vimin3_s32(__float_as_int(fmaxf(t0x, t1x)), __float_as_int(fmaxf(t0y, t1y)), zmax));
        return tentry <= texit;
    }
    __device__ __forceinline__ bool finished() const { return leaf < 0 && top == 0; }
    __device__ __forceinline__ bool wants_node() const { return leaf < 0 && top > 0; }
    __device__ __forceinline__ bool wants_leaf() const { return leaf >= 0 || top < 0; }

    // intersect_ray_box(min_max, ordered = false, ...), intersection.impala:194-208
    __device__ __forceinline__ bool hit_box(float lx, float hx, float ly, float hy, float lz, float hz, float& tentry) const {
        const float t0x = madd<FMA>(ray.idx, lx, ray.iox), t1x = madd<FMA>(ray.idx, hx, ray.iox);
        const float t0y = madd<FMA>(ray.idy, ly, ray.ioy), t1y = madd<FMA>(ray.idy, hy, ray.ioy);
        const float t0z = madd<FMA>(ray.idz, lz, ray.ioz), t1z = madd<FMA>(ray.idz, hz, ray.ioz);
        const int zmin = max(min(__float_as_int(t0z), __float_as_int(t1z)), __float_as_int(ray.tmin));      // fminmaxf
        const int zmax = min(max(__float_as_int(t0z), __float_as_int(t1z)), __float_as_int(tmax));          // fmaxminf
        tentry = __int_as_float(__vimax3_s32(__float_as_int(fminf(t0x, t1x)), __float_as_int(fminf(t0y, t1y)), zmin));
        const float texit = __int_as_float(__vimin3_s32(__float_as_int(fmaxf(t0x, t1x)), __float_as_int(fmaxf(t0y, t1y)), zmax));
        return tentry <= texit;                                            // mapping_gpu.impala:114
    }

    // One iteration of the outer loop up to the leaf loop (:106-135).
    __device__ __forceinline__ void node_step(const void* __restrict__ nodes_v) {
        int2 ch;
        float t0, t1;
        bool h0, h1;
        if constexpr (QUANT) {
            const F8 n = ldg8(reinterpret_cast<const float4*>(static_cast<const Node2q*>(nodes_v) + (top - 1)));   // the whole node
            ch = make_int2(__float_as_int(n.hi.z), __float_as_int(n.hi.w));
            h0 = hit_box_q(__float_as_uint(n.lo.x), __float_as_uint(n.lo.y), __float_as_uint(n.lo.z), t0);
            h1 = hit_box_q(__float_as_uint(n.lo.w), __float_as_uint(n.hi.x), __float_as_uint(n.hi.y), t1);
        } else {
            const float4* p = reinterpret_cast<const float4*>(static_cast<const Node2*>(nodes_v) + (top - 1));
            const F8 lo = ldg8(p), hi = ldg8(p + 2);                       // a Node2 is two 256-bit loads (32-byte aligned array)
            const float4 b0 = lo.lo, b1 = lo.hi, b2 = hi.lo;
            ch = make_int2(__float_as_int(hi.hi.x), __float_as_int(hi.hi.y));
            h0 = hit_box(b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, t0);          // box 0: lo/hi x, y, z (:33-36)
            h1 = hit_box(b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, t1);          // box 1 (:37-40)
        }
        if (h0 && h1) {
            const bool first0 = t0 < t1;                                   // :127-131
            top = first0 ? ch.x : ch.y;
            st.store(++ptr, first0 ? ch.y : ch.x);
        } else if (h0 || h1) {
            top = h0 ? ch.x : ch.y;                                        // :133
        } else {
            pop();                                                         // :122-123
        }
    }

    // One triangle of the leaf loop (:155-174).
    __device__ __forceinline__ void leaf_step(const Tri1* __restrict__ tris) {
        if (leaf < 0) { leaf = ~top; pop(); }
        const float4* p = reinterpret_cast<const float4*>(tris + leaf);
        const float4 a = ldg4(p), b = ldg4(p + 1), c = ldg4(p + 2);       // v0 | e1, geom | e2, prim
        const int prim = __float_as_int(c.w);
        const float nx = msub<FMA>(b.y, c.z, b.z, c.y);                    // cross(e1, e2), vector.impala:62-66
        const float ny = msub<FMA>(b.z, c.x, b.x, c.z);
        const float nz = msub<FMA>(b.x, c.y, b.y, c.x);
        float t, u, v;
        if (intersect_tri_lane<FMA>(ray, tmax, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, nx, ny, nz, t, u, v)) {
            hit.prim = prim & 0x7FFFFFFF; hit.geom = __float_as_int(b.w); hit.t = t; hit.u = u; hit.v = v;
            tmax = t;
            if (ANY) { top = 0; leaf = -1; return; }                       // early_exit, :168
        }
        leaf = prim < 0 ? -1 : leaf + 1;                                   // is_last, :62
    }
};

// The scheduler of traverse_vote_scheduled (traverse_sched.cuh) for this walker.
template <bool ANY, int SMEM_DEPTH, int BLOCK, bool FMA = false, bool QUANT = false, typename Fetch, typename Sink>
__device__ __forceinline__ void traverse_bvh2_scheduled(const void* __restrict__ nodes, const Tri1* __restrict__ tris, int* smem_column,
                                                        int num_rays, int* __restrict__ work_counter, int refill_min, int node_streak_min,
                                                        Fetch fetch, Sink sink, int leaf_streak_min = 0, QuantGrid grid = QuantGrid{}) {
    if (leaf_streak_min <= 0) leaf_streak_min = node_streak_min;
    const unsigned lane = lane_id();
    int overflow[kStackSize - SMEM_DEPTH];
    Bvh2Walker<ANY, SMEM_DEPTH, BLOCK, FMA, QUANT> w;
    w.st.smem = smem_column;
    w.st.overflow = overflow;
    w.leaf = -1; w.top = 0;
    int ray_idx = -1;
    bool drained = false;
    for (;;) {
        if (ray_idx >= 0 && w.finished()) { sink(ray_idx, w.hit); ray_idx = -1; }
        const unsigned idle = __ballot_sync(0xffffffffu, ray_idx < 0);
        if (!drained && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
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
                    if constexpr (QUANT) w.set_grid(grid);
                }
            }
            if (base + __popc(idle) >= num_rays) drained = true;
        }
        const bool has = ray_idx >= 0;
        const bool want_n = has && w.wants_node();
        const bool want_l = has && w.wants_leaf();
        const unsigned bn = __ballot_sync(0xffffffffu, want_n), bl = __ballot_sync(0xffffffffu, want_l);
        if ((bn | bl) == 0) break;                  // a fresh ray always wants a node step: nothing left and nothing to fetch
        if (__popc(bn) >= __popc(bl)) {
            bool go = want_n;
            do {
                if (go) w.node_step(nodes);
                go = has && w.wants_node();
            } while (__popc(__ballot_sync(0xffffffffu, go)) >= node_streak_min);
        } else {
            bool go = want_l;
            do {
                if (go) w.leaf_step(tris);
                go = has && w.wants_leaf();
            } while (__popc(__ballot_sync(0xffffffffu, go)) >= leaf_streak_min);
        }
    }
}

}  // namespace rb200
