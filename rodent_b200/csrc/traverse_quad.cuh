// Quad-per-ray BVH8/Tri4 traversal for sm_100a: four lanes cooperate on one ray, a
// warp carries eight rays.
//
// Why four lanes: the reference's layouts are SIMD layouts -- a Node8 is six rows of
// eight floats, a Tri4 twelve rows of four -- and its CPU kernel runs the slab test
// 8-wide and the triangle test 4-wide (src/traversal/mapping_cpu.impala:177-187,
// 224-249).  Here lane s of a quad tests children 2s and 2s+1 of the node (7 coalesced
// 8-byte loads) and triangle s of a Tri4 packet (13 4-byte loads); the reference's
// rv_ballot / cpu_reduce / rv_extract become __ballot_sync / __shfl_sync on the quad's
// lane mask.  All control flow is uniform within a quad, so divergence only exists
// between the eight quads of a warp, and it costs a fraction of what thread-per-ray
// divergence costs on incoherent rays.
//
// The visit order is exactly the reference's (stack.impala:52-123, sort.impala:34-66):
//  * the sequential "push / push_after" loop over hit children (mapping_cpu.impala:
//    195-208) is evaluated as a first-minimum prefix scan over the quad, which yields
//    the same register top and the same memory entries in the same positions;
//  * sort_n runs Batcher's network with one (n <= 4) or two (n <= 8) stack entries per
//    lane, comparators in dependency order.
// Arithmetic contract: common.cuh.  Hit records are bit-identical to the oracle.
#pragma once

#include "traverse.cuh"

namespace rb200 {

constexpr int kQuadBlock = 128;                       // 4 warps = 32 rays in flight per CTA
constexpr int kQuadsPerWarp = 8;
constexpr int kQuadStackBytesPerWarp = kStackSize * kQuadsPerWarp * 8;   // 4 KB

struct Entry { float t; int node; };                  // key first: compared far more often

__device__ __forceinline__ Entry shfl_entry(unsigned m, Entry e, int src) {
    Entry r;
    r.t = __shfl_sync(m, e.t, src, 4);
    r.node = __shfl_sync(m, e.node, src, 4);
    return r;
}
__device__ __forceinline__ Entry shfl_up_entry(unsigned m, Entry e, int d) {
    Entry r;
    r.t = __shfl_up_sync(m, e.t, d, 4);
    r.node = __shfl_up_sync(m, e.node, d, 4);
    return r;
}
__device__ __forceinline__ Entry shfl_xor_entry(unsigned m, Entry e, int x) {
    Entry r;
    r.t = __shfl_xor_sync(m, e.t, x, 4);
    r.node = __shfl_xor_sync(m, e.node, x, 4);
    return r;
}
// `later` displaces `earlier` only when strictly nearer: "t < stack.top().tmin" (mapping_cpu.impala:202)
__device__ __forceinline__ Entry first_min(Entry earlier, Entry later) { return later.t < earlier.t ? later : earlier; }

// One comparator of the network between my entry and a partner's: position i < j,
// swap iff key[i] < key[j] (stack.impala:91-96 with cmp = a < b).
__device__ __forceinline__ Entry cmp_exchange(Entry mine, Entry other, bool i_am_lower) {
    const bool swap = i_am_lower ? (mine.t < other.t) : (other.t < mine.t);
    return swap ? other : mine;
}

template <bool ANY>
struct QuadTraversal {
    // ray (replicated in the four lanes)
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, iox, ioy, ioz, tmin, tmax;
    unsigned o_nx, o_fx, o_ny, o_fy, o_nz, o_fz;     // byte offsets of this lane's float2 in the near/far rows
    bool degenerate;
    // stack: register top + shared-memory part (this quad's column)
    Entry top;
    int ptr;
    Entry* sp;
    // result
    HitRecord hit;
    // lane geometry
    unsigned qmask; int s;

    __device__ __forceinline__ Entry& slot(int i) { return sp[i * kQuadsPerWarp]; }
    // (the quad's lanes all read the slot; one of them may overwrite it with a push right after: order the two)
    __device__ __forceinline__ void pop() { top = slot(ptr); __syncwarp(qmask); --ptr; }

    __device__ __forceinline__ void begin(float4 r0, float4 r1) {
        ox = r0.x; oy = r0.y; oz = r0.z; tmin = r0.w;
        dx = r1.x; dy = r1.y; dz = r1.z; tmax = r1.w;
        idx = safe_rcp(dx); idy = safe_rcp(dy); idz = safe_rcp(dz);
        iox = -mul(ox, idx); ioy = -mul(oy, idy); ioz = -mul(oz, idz);
        const unsigned px = dx > 0.0f, py = dy > 0.0f, pz = dz > 0.0f;      // ray_octant
        const unsigned lane_off = unsigned(s) * 8u;
        o_nx = (1u - px) * 32u + lane_off;       o_fx = px * 32u + lane_off;           // rows 0/1
        o_ny = 64u + (1u - py) * 32u + lane_off; o_fy = 64u + py * 32u + lane_off;     // rows 2/3
        o_nz = 128u + (1u - pz) * 32u + lane_off; o_fz = 128u + pz * 32u + lane_off;   // rows 4/5
        degenerate = !(fabsf(iox) <= kFltMax && fabsf(ioy) <= kFltMax && fabsf(ioz) <= kFltMax) ||
                     idx == 0.0f || idy == 0.0f || idz == 0.0f;
        hit.prim = -1; hit.geom = -1; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;
        // stack.push(root, ray.tmin) on the empty stack (node 0, FLT_MAX)
        ptr = 0;
        if (s == 0) { slot(0).t = kFltMax; slot(0).node = 0; }
        top.node = 1; top.t = tmin;
        __syncwarp(qmask);
    }

    template <bool X86_NAN>
    __device__ __forceinline__ void slabs(float2 nx, float2 ny, float2 nz, float2 fx, float2 fy, float2 fz,
                                          float& ta, float& tb, bool& ha, bool& hb) const {
        {
            const float t0x = slab<X86_NAN>(idx, nx.x, iox), t0y = slab<X86_NAN>(idy, ny.x, ioy), t0z = slab<X86_NAN>(idz, nz.x, ioz);
            const float t1x = slab<X86_NAN>(idx, fx.x, iox), t1y = slab<X86_NAN>(idy, fy.x, ioy), t1z = slab<X86_NAN>(idz, fz.x, ioz);
            ta = imax2(imax3(t0x, t0y, t0z), tmin);
            const float tx = imin2(imin3(t1x, t1y, t1z), tmax);
            ha = !(__float_as_int(tx) < __float_as_int(ta));
        }
        {
            const float t0x = slab<X86_NAN>(idx, nx.y, iox), t0y = slab<X86_NAN>(idy, ny.y, ioy), t0z = slab<X86_NAN>(idz, nz.y, ioz);
            const float t1x = slab<X86_NAN>(idx, fx.y, iox), t1y = slab<X86_NAN>(idy, fy.y, ioy), t1z = slab<X86_NAN>(idz, fz.y, ioz);
            tb = imax2(imax3(t0x, t0y, t0z), tmin);
            const float tx = imin2(imin3(t1x, t1y, t1z), tmax);
            hb = !(__float_as_int(tx) < __float_as_int(tb));
        }
    }

    // sort_n (stack.impala:79-111) over slots first .. first+n-1, 3 <= n <= 8.
    __device__ __forceinline__ void sort_n(int first, int n) {
        const Entry pad = {-INFINITY, 0};            // never moves: comparators beyond n are no-ops
        if (n <= 4) {
            // one entry per lane; batcher_sort(3|4) = (0,1)(2,3) | (0,2)(1,3) | (1,2)
            Entry e = s < n ? slot(first + s) : pad;
            e = cmp_exchange(e, shfl_xor_entry(qmask, e, 1), (s & 1) == 0);
            e = cmp_exchange(e, shfl_xor_entry(qmask, e, 2), (s & 2) == 0);
            const Entry o = shfl_xor_entry(qmask, e, 3);
            if (s == 1 || s == 2) e = cmp_exchange(e, o, s == 1);
            if (s < n) slot(first + s) = e;
        } else {
            // two entries per lane, positions 2s and 2s+1; batcher_sort(5..8) in dependency order:
            // (0,1)(2,3)(4,5)(6,7) | (0,2)(1,3)(4,6)(5,7) | (1,2)(5,6) | (0,4)(2,6)(1,5)(3,7) | (2,4)(3,5) | (1,2)(3,4)(5,6)
            Entry a = 2 * s < n ? slot(first + 2 * s) : pad;
            Entry b = 2 * s + 1 < n ? slot(first + 2 * s + 1) : pad;
            { const bool sw = a.t < b.t; const Entry t = a; a = sw ? b : a; b = sw ? t : b; }
            {   // lanes 0<->1, 2<->3, same slots
                const Entry oa = shfl_xor_entry(qmask, a, 1), ob = shfl_xor_entry(qmask, b, 1);
                a = cmp_exchange(a, oa, (s & 1) == 0); b = cmp_exchange(b, ob, (s & 1) == 0);
            }
            {   // (1,2): lane0.b <-> lane1.a ; (5,6): lane2.b <-> lane3.a
                const Entry from_hi = shfl_xor_entry(qmask, a, 1), from_lo = shfl_xor_entry(qmask, b, 1);
                if ((s & 1) == 0) b = cmp_exchange(b, from_hi, true); else a = cmp_exchange(a, from_lo, false);
            }
            {   // lanes 0<->2, 1<->3, same slots
                const Entry oa = shfl_xor_entry(qmask, a, 2), ob = shfl_xor_entry(qmask, b, 2);
                a = cmp_exchange(a, oa, (s & 2) == 0); b = cmp_exchange(b, ob, (s & 2) == 0);
            }
            {   // (2,4)(3,5): lane1 <-> lane2, same slots
                const Entry oa = shfl_xor_entry(qmask, a, 3), ob = shfl_xor_entry(qmask, b, 3);
                if (s == 1 || s == 2) { a = cmp_exchange(a, oa, s == 1); b = cmp_exchange(b, ob, s == 1); }
            }
            {   // (1,2)(3,4)(5,6): lane s.b <-> lane s+1.a
                const Entry up = shfl_up_entry(qmask, b, 1);                      // from lane s-1: its b
                Entry dn; dn.t = __shfl_down_sync(qmask, a.t, 1, 4); dn.node = __shfl_down_sync(qmask, a.node, 1, 4);   // from lane s+1: its a
                if (s < 3) b = cmp_exchange(b, dn, true);
                if (s > 0) a = cmp_exchange(a, up, false);
            }
            if (2 * s < n) slot(first + 2 * s) = a;
            if (2 * s + 1 < n) slot(first + 2 * s + 1) = b;
        }
        __syncwarp(qmask);
    }

    // Returns true when the ray is finished.
    template <bool WANT_GEOM, typename Yield>
    __device__ __forceinline__ bool run(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris, Yield should_yield) {
        const char* nbase = reinterpret_cast<const char*>(nodes);
        const unsigned lt_s = (1u << s) - 1u;
        for (;;) {
            if (top.node == 0) return true;                                         // mapping_cpu.impala:168
            if (!ANY && top.t > tmax) { pop(); continue; }                          // :170-174
            if (should_yield()) return false;

            bool restart = false;
            while (top.node > 0) {                                                  // :177
                const unsigned noff = unsigned(top.node - 1) << 8;
                pop();
                const float2 nx = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_nx)));
                const float2 ny = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_ny)));
                const float2 nz = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_nz)));
                const float2 fx = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_fx)));
                const float2 fy = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_fy)));
                const float2 fz = __ldg(reinterpret_cast<const float2*>(nbase + (noff + o_fz)));
                const int2 ch = __ldg(reinterpret_cast<const int2*>(nbase + (noff + 192u + unsigned(s) * 8u)));

                float ta, tb; bool ha, hb;
                if (degenerate) slabs<true>(nx, ny, nz, fx, fy, fz, ta, tb, ha, hb);
                else            slabs<false>(nx, ny, nz, fx, fy, fz, ta, tb, ha, hb);
                const int shift = (threadIdx.x & 28);                                // first lane of the quad
                const unsigned ba = (__ballot_sync(qmask, ha) >> shift) & 0xFu;      // children 0,2,4,6
                const unsigned bb = (__ballot_sync(qmask, hb) >> shift) & 0xFu;      // children 1,3,5,7
                if ((ba | bb) == 0) {                                                // :189-191
                    if (ANY) continue;
                    restart = true;
                    break;
                }
                const int n = __popc(ba) + __popc(bb);
                const int ka = __popc(ba & lt_s) + __popc(bb & lt_s);               // hit children before child 2s
                const int kb = ka + int(ha);
                const Entry ea = {ta, ch.x}, eb = {tb, ch.y};
                if (ANY) {
                    // every hit child goes on top (:202): memory gets [old top, c_1 .. c_{n-1}], top = c_n
                    if (s == 0) slot(ptr + 1) = top;
                    if (ha && ka + 1 < n) slot(ptr + 2 + ka) = ea;
                    if (hb && kb + 1 < n) slot(ptr + 2 + kb) = eb;
                    const int last_lane = 31 - __clz(int(ba | bb));
                    const Entry last = (bb >> last_lane) & 1u ? eb : ea;
                    top = shfl_entry(qmask, last, last_lane);
                    ptr += n;
                    __syncwarp(qmask);
                } else {
                    // The push loop of :195-208 as a scan.  Sequence: old top, then hit children in lane
                    // order; running top = first minimum; each child leaves either itself or the top it
                    // displaced in memory, at the position of its rank among the hit children.
                    const Entry none = {INFINITY, 0};
                    const Entry la = ha ? ea : none, lb = hb ? eb : none;
                    Entry incl = first_min(la, lb);
                    { const Entry up = shfl_up_entry(qmask, incl, 1); if (s >= 1) incl = first_min(up, incl); }
                    { const Entry up = shfl_up_entry(qmask, incl, 2); if (s >= 2) incl = first_min(up, incl); }
                    Entry before = shfl_up_entry(qmask, incl, 1);                    // children of lanes < s
                    if (s == 0) before = none;
                    Entry cur = first_min(top, before);                              // running top before child 2s
                    if (ha) {
                        const bool nearer = ta < cur.t;
                        slot(ptr + 1 + ka) = nearer ? cur : ea;
                        cur = nearer ? ea : cur;
                    }
                    if (hb) {
                        const bool nearer = tb < cur.t;
                        slot(ptr + 1 + kb) = nearer ? cur : eb;
                        cur = nearer ? eb : cur;
                    }
                    top = shfl_entry(qmask, cur, 3);
                    ptr += n;
                    __syncwarp(qmask);
                    if (n >= 3) sort_n(ptr - n + 1, n);                              // :210-218
                }
            }
            if (restart) continue;
            if (top.node == 0) return true;   // :221; closest-hit: unreachable with finite keys (see traverse.cuh)

            int prim_id = ~top.node;                                                 // :224
            pop();
            bool terminated = false;
            for (;;) {
                const char* tp = reinterpret_cast<const char*>(tris + prim_id) + s * 4;
                prim_id++;
                const int pid = __ldg(reinterpret_cast<const int*>(tp + 192));
#define RB_ROW(r) __ldg(reinterpret_cast<const float*>(tp + (r) * 16))
                const float v0x = RB_ROW(0), v0y = RB_ROW(1), v0z = RB_ROW(2);
                const float e1x = RB_ROW(3), e1y = RB_ROW(4), e1z = RB_ROW(5);
                const float e2x = RB_ROW(6), e2y = RB_ROW(7), e2z = RB_ROW(8);
                const float nnx = RB_ROW(9), nny = RB_ROW(10), nnz = RB_ROW(11);
#undef RB_ROW
                float lt = kFltMax, lu = 0.0f, lv = 0.0f;
                bool h = false;
                if (pid != -1) {                                                     // is_valid
                    // intersect_ray_tri, intersection.impala:164-192 (one lane of the Tri4)
                    const float cx = sub(v0x, ox), cy = sub(v0y, oy), cz = sub(v0z, oz);
                    const float rx = sub(mul(dy, cz), mul(dz, cy));
                    const float ry = sub(mul(dz, cx), mul(dx, cz));
                    const float rz = sub(mul(dx, cy), mul(dy, cx));
                    const float det = dot3(nnx, nny, nnz, dx, dy, dz);
                    const float abs_det = fabsf(det);
                    const float u = prodsign(dot3(rx, ry, rz, e2x, e2y, e2z), det);
                    const float v = prodsign(dot3(rx, ry, rz, e1x, e1y, e1z), det);
                    if (u >= 0.0f && v >= 0.0f && add(u, v) <= abs_det) {
                        const float t = prodsign(dot3(cx, cy, cz, nnx, nny, nnz), det);
                        if (abs_det != 0.0f && t >= mul(abs_det, tmin) && t <= mul(abs_det, tmax)) {
                            const float inv_det = __fdiv_rn(1.0f, abs_det);
                            lt = mul(t, inv_det); lu = mul(u, inv_det); lv = mul(v, inv_det);
                            h = true;
                        }
                    }
                }
                const unsigned hm = __ballot_sync(qmask, h) & qmask;
                if (hm) {
                    int lane;                                                        // lane within the quad
                    if (ANY) {
                        lane = (__ffs(hm) - 1) & 3;                                  // :234-237
                        terminated = true;
                    } else {
                        // cpu_reduce with the integer min, then the first lane holding it (:239-242)
                        float mn = imin2(lt, __shfl_xor_sync(qmask, lt, 2, 4));
                        mn = imin2(mn, __shfl_xor_sync(qmask, mn, 1, 4));
                        lane = (__ffs(__ballot_sync(qmask, lt == mn) & qmask) - 1) & 3;
                    }
                    hit.prim = __shfl_sync(qmask, pid, lane, 4) & 0x7FFFFFFF;        // mapping_cpu.impala:34
                    hit.t = __shfl_sync(qmask, lt, lane, 4);
                    hit.u = __shfl_sync(qmask, lu, lane, 4);
                    hit.v = __shfl_sync(qmask, lv, lane, 4);
                    if (WANT_GEOM) {
                        const int gid = __ldg(reinterpret_cast<const int*>(tp + 208));
                        hit.geom = __shfl_sync(qmask, gid, lane, 4);
                    }
                    if (!ANY) tmax = hit.t;                                          // :243
                }
                if (__shfl_sync(qmask, pid, 3, 4) < 0) break;                        // is_last
            }
            if (ANY && terminated) return true;                                      // :252
        }
    }
};

}  // namespace rb200
