// BVH8/Tri4 single-ray traversal for sm_100a: the device-side core shared by the
// bench_traversal entry points (traverse.cu) and the wavefront renderer.
//
// Semantics are those of the reference's CPU single-ray kernel
//   cpu_traverse_single_helper, src/traversal/mapping_cpu.impala:138-256
// with its stack discipline (src/traversal/stack.impala:52-123), Batcher network
// (src/core/sort.impala:34-66), ordered slab test with integer min/max
// (src/traversal/intersection.impala:194-208, mapping_cpu.impala:123-133) and
// triangle test (intersection.impala:164-192), under the arithmetic contract in
// common.cuh.  One thread owns one ray; the visit order, and with it every tie, is
// the reference's, so hit records are bit-identical to the oracle by construction.
#pragma once

#include "common.cuh"

namespace rb200 {

constexpr int kStackSize = 64;   // src/traversal/stack.impala:53-54

struct StackEntry { int node; float tmin; };

// Per-ray constants, make_ray + ray_octant (intersection.impala:88-99,128-132).
struct RaySetup {
    float ox, oy, oz, dx, dy, dz;
    float idx, idy, idz, iox, ioy, ioz;
    float tmin;
    // float4 index (within a node) of the near/far plane rows per axis:
    // ordered_bbox, mapping_cpu.impala:88-106.  Row r of bounds[6][8] = float4 2r, 2r+1.
    int near_x, far_x, near_y, far_y, near_z, far_z;
    // A slab term inv_dir * plane + inv_org can only become NaN (inf - inf, 0 * inf) when
    // inv_org overflowed or inv_dir is 0, i.e. for clamped (|d| < 1e-8) axes.  The integer
    // min/max then orders the NaN by its bits, and the reference's x86 NaN (0xFFC00000,
    // the SSE "indefinite") is not the GPU's (0x7FFFFFFF): such rays take a slow path that
    // rewrites generated NaNs to the x86 pattern.
    bool degenerate;

    // ROW: float4 per bounds row, 2 for a Node8 (bounds[6][8]), 1 for a Node4 (bounds[6][4]).
    template <int ROW = 2>
    __device__ __forceinline__ void init(float4 r0, float4 r1) {
        ox = r0.x; oy = r0.y; oz = r0.z; tmin = r0.w;
        dx = r1.x; dy = r1.y; dz = r1.z;
        idx = safe_rcp(dx); idy = safe_rcp(dy); idz = safe_rcp(dz);
        iox = -mul(ox, idx); ioy = -mul(oy, idy); ioz = -mul(oz, idz);
        const int px = dx > 0.0f, py = dy > 0.0f, pz = dz > 0.0f;
        near_x = ROW * (1 - px); far_x = ROW * px;
        near_y = 2 * ROW + ROW * (1 - py); far_y = 2 * ROW + ROW * py;
        near_z = 4 * ROW + ROW * (1 - pz); far_z = 4 * ROW + ROW * pz;
        const float big = kFltMax;
        degenerate = !(fabsf(iox) <= big && fabsf(ioy) <= big && fabsf(ioz) <= big) ||
                     idx == 0.0f || idy == 0.0f || idz == 0.0f;
    }
};

struct HitRecord { int prim; int geom; float t, u, v; };

// inv_dir * plane + inv_org (intersection.impala:195-196); X86_NAN: see RaySetup::degenerate.
template <bool X86_NAN, bool FMA = false>
__device__ __forceinline__ float slab(float inv_dir, float plane, float inv_org) {
    const float r = madd<FMA>(inv_dir, plane, inv_org);
    if (X86_NAN) return r != r ? __int_as_float(int(0xFFC00000u)) : r;
    return r;
}

// One lane of a Tri4 (mapping_cpu.impala:24-42 + intersection.impala:164-192).
template <bool FMA = false>
__device__ __forceinline__ bool intersect_tri_lane(const RaySetup& r, float tmax,
                                                   float v0x, float v0y, float v0z,
                                                   float e1x, float e1y, float e1z,
                                                   float e2x, float e2y, float e2z,
                                                   float nx, float ny, float nz,
                                                   float& t_out, float& u_out, float& v_out) {
    const float cx = sub(v0x, r.ox), cy = sub(v0y, r.oy), cz = sub(v0z, r.oz);
    const float rx = msub<FMA>(r.dy, cz, r.dz, cy);
    const float ry = msub<FMA>(r.dz, cx, r.dx, cz);
    const float rz = msub<FMA>(r.dx, cy, r.dy, cx);
    const float det = dot3f<FMA>(nx, ny, nz, r.dx, r.dy, r.dz);
    const float abs_det = fabsf(det);
    const float u = prodsign(dot3f<FMA>(rx, ry, rz, e2x, e2y, e2z), det);
    const float v = prodsign(dot3f<FMA>(rx, ry, rz, e1x, e1y, e1z), det);
    if (!(u >= 0.0f && v >= 0.0f && add(u, v) <= abs_det)) return false;
    const float t = prodsign(dot3f<FMA>(cx, cy, cz, nx, ny, nz), det);
    if (!(abs_det != 0.0f && t >= mul(abs_det, r.tmin) && t <= mul(abs_det, tmax))) return false;
    const float inv_det = __fdiv_rn(1.0f, abs_det);
    t_out = mul(t, inv_det); u_out = mul(u, inv_det); v_out = mul(v, inv_det);
    return true;
}

#define RB_CSWAP(i, j)                                                     \
    {                                                                      \
        const bool s__ = e[i].tmin < e[j].tmin;                            \
        const StackEntry a__ = e[i], b__ = e[j];                           \
        e[i].node = s__ ? b__.node : a__.node; e[i].tmin = s__ ? b__.tmin : a__.tmin; \
        e[j].node = s__ ? a__.node : b__.node; e[j].tmin = s__ ? a__.tmin : b__.tmin; \
    }

// sort_n of stack.impala:79-111 on the n entries st[first .. first+n-1], 3 <= n <= 8,
// with batcher_sort(n) of sort.impala:34-66.  Slots >= n are padded with -inf keys:
// a comparator (i, j >= n) can then never fire, which is the reference's "remove
// comparators for non-existing elements"; for n <= 4 the network is the 4-input one.
template <typename Stack>
__device__ __forceinline__ void sort_entries(Stack& st, int first, int n) {
    StackEntry e[8];
    if (n <= 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < n) e[k] = st.load(first + k);
            else { e[k].node = 0; e[k].tmin = -INFINITY; }
        }
        RB_CSWAP(0, 1) RB_CSWAP(2, 3) RB_CSWAP(0, 2) RB_CSWAP(1, 3) RB_CSWAP(1, 2)
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < n) st.store(first + k, e[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < n) e[k] = st.load(first + k);
            else { e[k].node = 0; e[k].tmin = -INFINITY; }
        }
        RB_CSWAP(0, 1) RB_CSWAP(2, 3) RB_CSWAP(0, 2) RB_CSWAP(1, 3) RB_CSWAP(1, 2)
        RB_CSWAP(4, 5) RB_CSWAP(6, 7) RB_CSWAP(4, 6) RB_CSWAP(5, 7) RB_CSWAP(5, 6)
        RB_CSWAP(0, 4) RB_CSWAP(2, 6) RB_CSWAP(2, 4) RB_CSWAP(1, 5) RB_CSWAP(3, 7) RB_CSWAP(3, 5)
        RB_CSWAP(1, 2) RB_CSWAP(3, 4) RB_CSWAP(5, 6)
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < n) st.store(first + k, e[k]);
    }
}
// sort_n for a BVH4 (arity 4): bose_nelson_sort(n) of sort.impala:3-32, n = 3 or 4.  Its three-input network
// is (1,2)(0,2)(0,1) -- not the (0,1)(0,2)(1,2) that Batcher's gives for an arity-8 node.
template <typename Stack>
__device__ __forceinline__ void sort_entries_bvh4(Stack& st, int first, int n) {
    StackEntry e[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < n) e[k] = st.load(first + k);
        else { e[k].node = 0; e[k].tmin = -INFINITY; }
    }
    if (n == 3) { RB_CSWAP(1, 2) RB_CSWAP(0, 2) RB_CSWAP(0, 1) }
    else        { RB_CSWAP(0, 1) RB_CSWAP(2, 3) RB_CSWAP(0, 2) RB_CSWAP(1, 3) RB_CSWAP(1, 2) }
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n) st.store(first + k, e[k]);
}
#undef RB_CSWAP

// The memory part of the traversal stack (stack.impala:53-54: 64 entries).  The first
// SMEM_DEPTH levels live in shared memory, laid out [level][thread] so that a warp's
// accesses to one level are conflict-free; deeper levels (0.02 % of the accesses on Sponza)
// spill to thread-local memory.  SMEM_DEPTH = 0 keeps everything thread-local.
template <int SMEM_DEPTH, int BLOCK>
struct HybridStack {
    StackEntry* smem;                                   // this thread's column: smem[level * BLOCK]
    StackEntry local[kStackSize - SMEM_DEPTH];
    __device__ __forceinline__ StackEntry load(int i) const {
        if (SMEM_DEPTH > 0 && i < SMEM_DEPTH) return smem[i * BLOCK];
        return local[i - SMEM_DEPTH];
    }
    __device__ __forceinline__ void store(int i, StackEntry e) {
        if (SMEM_DEPTH > 0 && i < SMEM_DEPTH) smem[i * BLOCK] = e;
        else local[i - SMEM_DEPTH] = e;
    }
};

// Traversal state of one ray.  `step()` runs the reference's outer loop until the
// ray is finished or `should_yield()` asks to return to the scheduler (persistent
// kernels use that to refill idle lanes); state survives across calls.
template <bool ANY, int SMEM_DEPTH = 0, int BLOCK = 128>
struct Traversal {
    RaySetup ray;
    float tmax;
    int top_node; float top_t; int ptr;
    HitRecord hit;
    HybridStack<SMEM_DEPTH, BLOCK> st;

    __device__ __forceinline__ void push(int n, float t) { ++ptr; st.store(ptr, StackEntry{top_node, top_t}); top_node = n; top_t = t; }
    __device__ __forceinline__ void push_after(int n, float t) { ++ptr; st.store(ptr, StackEntry{n, t}); }
    __device__ __forceinline__ void pop() { const StackEntry e = st.load(ptr); top_node = e.node; top_t = e.tmin; --ptr; }

    __device__ __forceinline__ void begin(float4 r0, float4 r1) {
        ray.init(r0, r1);
        tmax = r1.w;
        hit.prim = -1; hit.geom = -1; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;   // empty_hit, intersection.impala:134-136
        ptr = -1; top_node = 0; top_t = kFltMax;
        push(1, ray.tmin);                                                          // mapping_cpu.impala:153
    }

    // Returns true when the ray is finished.
    template <bool WANT_GEOM, typename Yield>
    __device__ __forceinline__ bool run(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris, Yield should_yield) {
        for (;;) {
            if (top_node == 0) return true;                                         // :168
            if (!ANY && top_t > tmax) { pop(); continue; }                          // :170-174
            if (should_yield()) return false;

            bool restart = false;
            while (top_node > 0) {                                                  // :177
                const float4* nb = reinterpret_cast<const float4*>(nodes + (top_node - 1));
                pop();
                // all 14 loads of the node are issued before the first use
                const float4 nxa = ldg4(nb + ray.near_x), nxb = ldg4(nb + ray.near_x + 1);
                const float4 nya = ldg4(nb + ray.near_y), nyb = ldg4(nb + ray.near_y + 1);
                const float4 nza = ldg4(nb + ray.near_z), nzb = ldg4(nb + ray.near_z + 1);
                const float4 fxa = ldg4(nb + ray.far_x), fxb = ldg4(nb + ray.far_x + 1);
                const float4 fya = ldg4(nb + ray.far_y), fyb = ldg4(nb + ray.far_y + 1);
                const float4 fza = ldg4(nb + ray.far_z), fzb = ldg4(nb + ray.far_z + 1);
                const int4 ca = ldg4(reinterpret_cast<const int4*>(nb) + 12), cb = ldg4(reinterpret_cast<const int4*>(nb) + 13);

                const float nx[8] = {nxa.x, nxa.y, nxa.z, nxa.w, nxb.x, nxb.y, nxb.z, nxb.w};
                const float ny[8] = {nya.x, nya.y, nya.z, nya.w, nyb.x, nyb.y, nyb.z, nyb.w};
                const float nz[8] = {nza.x, nza.y, nza.z, nza.w, nzb.x, nzb.y, nzb.z, nzb.w};
                const float fx[8] = {fxa.x, fxa.y, fxa.z, fxa.w, fxb.x, fxb.y, fxb.z, fxb.w};
                const float fy[8] = {fya.x, fya.y, fya.z, fya.w, fyb.x, fyb.y, fyb.z, fyb.w};
                const float fz[8] = {fza.x, fza.y, fza.z, fza.w, fzb.x, fzb.y, fzb.z, fzb.w};
                const int child[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};

                // ordered slab test, intersection.impala:194-208 with integer min/max
                float tentry[8];
                unsigned mask = 0;
#define RB_SLABS(X86_NAN)                                                                        \
                _Pragma("unroll")                                                                \
                for (int i = 0; i < 8; i++) {                                                    \
                    const float t0x = slab<X86_NAN>(ray.idx, nx[i], ray.iox);                    \
                    const float t0y = slab<X86_NAN>(ray.idy, ny[i], ray.ioy);                    \
                    const float t0z = slab<X86_NAN>(ray.idz, nz[i], ray.ioz);                    \
                    const float t1x = slab<X86_NAN>(ray.idx, fx[i], ray.iox);                    \
                    const float t1y = slab<X86_NAN>(ray.idy, fy[i], ray.ioy);                    \
                    const float t1z = slab<X86_NAN>(ray.idz, fz[i], ray.ioz);                    \
                    const float te = imax2(imax3(t0x, t0y, t0z), ray.tmin);                      \
                    const float tx = imin2(imin3(t1x, t1y, t1z), tmax);                          \
                    tentry[i] = te;                                                              \
                    if (!(__float_as_int(tx) < __float_as_int(te))) mask |= 1u << i; /* :184 */  \
                }
                if (ray.degenerate) { RB_SLABS(true) } else { RB_SLABS(false) }
#undef RB_SLABS
                if (mask == 0) {                                                     // :189-191
                    if (ANY) continue;
                    restart = true;
                    break;
                }
                // pushes in lane order, :195-208
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (mask & (1u << i)) {
                        if (ANY || tentry[i] < top_t) push(child[i], tentry[i]);
                        else push_after(child[i], tentry[i]);
                    }
                }
                if (!ANY) {                                                          // :210-218
                    const int n = __popc(mask);
                    if (n >= 3) sort_entries(st, ptr - n + 1, n);
                }
            }
            if (restart) continue;
            if (top_node == 0) return true;   // :221 (closest-hit: reference would read prim -1; unreachable with finite keys)

            int prim_id = ~top_node;                                                 // :224
            pop();
            bool terminated = false;
            for (;;) {
                const float4* tp = reinterpret_cast<const float4*>(tris + prim_id);
                prim_id++;
                const int4 pid = ldg4(reinterpret_cast<const int4*>(tp) + 12);
                const float4 v0x = ldg4(tp + 0), v0y = ldg4(tp + 1), v0z = ldg4(tp + 2);
                const float4 e1x = ldg4(tp + 3), e1y = ldg4(tp + 4), e1z = ldg4(tp + 5);
                const float4 e2x = ldg4(tp + 6), e2y = ldg4(tp + 7), e2z = ldg4(tp + 8);
                const float4 nnx = ldg4(tp + 9), nny = ldg4(tp + 10), nnz = ldg4(tp + 11);

                float lt[4], lu[4], lv[4];
                unsigned hm = 0;
#define RB_LANE(j, c)                                                                               \
                lt[j] = kFltMax; lu[j] = 0.0f; lv[j] = 0.0f;                                       \
                if (pid.c != -1 &&                                                                 \
                    intersect_tri_lane(ray, tmax, v0x.c, v0y.c, v0z.c, e1x.c, e1y.c, e1z.c,        \
                                       e2x.c, e2y.c, e2z.c, nnx.c, nny.c, nnz.c, lt[j], lu[j], lv[j])) \
                    hm |= 1u << j;
                RB_LANE(0, x) RB_LANE(1, y) RB_LANE(2, z) RB_LANE(3, w)
#undef RB_LANE
                if (hm) {
                    int lane;
                    if (ANY) {
                        lane = __ffs(hm) - 1;                                        // :234-237
                        terminated = true;
                    } else {
                        // cpu_reduce with integer min, then first lane equal to it (:239-242)
                        const float mn = imin2(imin2(lt[0], lt[2]), imin2(lt[1], lt[3]));
                        lane = lt[0] == mn ? 0 : lt[1] == mn ? 1 : lt[2] == mn ? 2 : 3;
                    }
                    const int p = lane == 0 ? pid.x : lane == 1 ? pid.y : lane == 2 ? pid.z : pid.w;
                    hit.prim = p & 0x7FFFFFFF;                                       // mapping_cpu.impala:34
                    hit.t = lane == 0 ? lt[0] : lane == 1 ? lt[1] : lane == 2 ? lt[2] : lt[3];
                    hit.u = lane == 0 ? lu[0] : lane == 1 ? lu[1] : lane == 2 ? lu[2] : lu[3];
                    hit.v = lane == 0 ? lv[0] : lane == 1 ? lv[1] : lane == 2 ? lv[2] : lv[3];
                    if (WANT_GEOM) {
                        const int4 gid = ldg4(reinterpret_cast<const int4*>(tp) + 13);
                        hit.geom = lane == 0 ? gid.x : lane == 1 ? gid.y : lane == 2 ? gid.z : gid.w;
                    }
                    if (!ANY) tmax = hit.t;                                          // :243
                }
                if (pid.w < 0) break;                                                // is_last, mapping_cpu.impala:40
            }
            if (ANY && terminated) return true;                                      // :252
        }
    }
};

}  // namespace rb200
