/*
 * traversal_oracle.c -- CPU restatement of the reference's single-ray BVH
 * traversal.  TEST INFRASTRUCTURE ONLY: nothing under rodent_b200/ may link or
 * call this; it exists to check the CUDA path (tests/, __graft_entry__.smoke())
 * and to be timed as the CPU baseline (bench.py cpu_baseline / --impl reference).
 *
 * Parity status: PINNED for hit distance t by the reference's golden images
 * testing/ref-primary.png and testing/ref-random.png (tests/test_oracle_golden.py);
 * tri_id/u/v have no golden vector in the reference, for those this file *is* the
 * definition (see DESIGN.md "Parity contract").
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest, no FMA contraction
 * (build with -ffp-contract=off, no -ffast-math; the AVX2 forms of the node and triangle tests use separate
 * multiply and add intrinsics and give the same bits as the scalar code, which stays as the fallback), operations in the source order
 * of the reference, integer compares of float bits where the reference uses them
 * (enable_cpu_int_min_max = true, tools/bench_traversal/bench_traversal.impala:11).
 * The reference binary itself is built -ffast-math (CMakeLists.txt:12), so its own
 * last bits are not reproducible; this contract is what "bit-exact" means here.
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference tree).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#if defined(__AVX2__)
#include <immintrin.h>    /* the 8-lane node test and the 4-lane triangle test below, as the reference's vectorised build has them */
#endif

#include "../include/rodent_b200.h"

#define FLT_MAX_ 3.4028234664e+38f  /* src/core/common.impala:4 */
#define STACK_SIZE 64               /* src/traversal/stack.impala:53-54 */

typedef struct { int32_t node; float tmin; } NodeRef;

/* Optional step trace for scheduling studies (scripts/sim_sched.c defines it): one
 * call per inner-node visit ('N') and per Tri4 packet ('L').  Empty in the oracle build. */
#ifndef ORACLE_TRACE
#define ORACLE_TRACE(kind) ((void)0)
#endif
/* ... and per inner-node visit the hit mask and which of the hit children became the new top (bit set)
 * rather than going below it: the push blocks a SIMT lane would execute (scripts/sim_sched2.c). */
#ifndef ORACLE_TRACE_PUSHES
#define ORACLE_TRACE_PUSHES(mask, tops) ((void)0)
#endif

/* Work counters, for the roofline's algorithmic bytes. */
typedef struct OracleStats {
    uint64_t nodes;   /* inner nodes popped and box-tested  */
    uint64_t tri4;    /* Tri4 packets fetched in leaves     */
    uint64_t max_stack;
} OracleStats;

static inline int32_t f2i(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float   i2f(int32_t i) { float x; memcpy(&x, &i, 4); return x; }

/* src/core/common.impala:78-80 */
static inline float prodsign(float x, float y) {
    return i2f(f2i(x) ^ (f2i(y) & (int32_t)0x80000000u));
}
/* src/core/common.impala:82-85 */
static inline float safe_rcp(float x) {
    const float min_rcp = 1e-8f;
    float ax = x > 0.0f ? x : -x;
    return ax < min_rcp ? prodsign(FLT_MAX_, x) : 1.0f / x;
}
/* src/traversal/mapping_cpu.impala:123-133 : min/max on the float bits as signed ints */
static inline float imin(float a, float b) { int32_t x = f2i(a), y = f2i(b); return i2f(x < y ? x : y); }
static inline float imax(float a, float b) { int32_t x = f2i(a), y = f2i(b); return i2f(x > y ? x : y); }

/* ---- sorting networks, src/core/sort.impala:3-66 ------------------------- */
typedef struct { int n; int8_t a[32], b[32]; } Network;
static Network g_batcher[9];      /* arity 8: batcher_sort(n), n = 3..8      */
static Network g_bose_nelson[5];  /* arity <= 4: bose_nelson_sort(n), n = 3,4 */
static pthread_once_t g_net_once = PTHREAD_ONCE_INIT;

static void net_add(Network* w, int i, int j) { w->a[w->n] = (int8_t)i; w->b[w->n] = (int8_t)j; w->n++; }

/* src/core/common.impala:107-116 : smallest p with i <= 2^p */
static int ilog2_(int i) { int p = 0; while (i > (1 << p)) p++; return p; }

/* src/core/sort.impala:35-52 */
static void batcher_merge(Network* w, int n, int i, int len, int r) {
    int step = r * 2;
    if (step < len) {
        batcher_merge(w, n, i, len, step);
        batcher_merge(w, n, i + r, len, step);
        for (int j = i + r; j < i + len - r; j += step)
            if (j < n && j + r < n) net_add(w, j, j + r);
    } else {
        if (i < n && i + r < n) net_add(w, i, i + r);
    }
}
/* src/core/sort.impala:54-61 */
static void batcher_sort_rec(Network* w, int n, int i, int len) {
    if (len > 1) {
        int m = len / 2;
        batcher_sort_rec(w, n, i, m);
        batcher_sort_rec(w, n, i + m, m);
        batcher_merge(w, n, i, len, 1);
    }
}
/* src/core/sort.impala:13-29 */
static void bn_bracket(Network* w, int i1, int len1, int i2, int len2) {
    if (len1 == 1 && len2 == 1) {
        net_add(w, i1, i2);
    } else if (len1 == 1 && len2 == 2) {
        net_add(w, i1, i2 + 1);
        net_add(w, i1, i2);
    } else if (len1 == 2 && len2 == 1) {
        net_add(w, i1, i2);
        net_add(w, i1 + 1, i2);
    } else {
        int a = len1 / 2;
        int b = (len1 % 2 != 0) ? len2 / 2 : (len2 + 1) / 2;
        bn_bracket(w, i1, a, i2, b);
        bn_bracket(w, i1 + a, len1 - a, i2 + b, len2 - b);
        bn_bracket(w, i1 + a, len1 - a, i2, b);
    }
}
/* src/core/sort.impala:4-11 */
static void bn_star(Network* w, int i, int len) {
    if (len > 1) {
        int m = len / 2;
        bn_star(w, i, m);
        bn_star(w, i + m, len - m);
        bn_bracket(w, i, m, i + m, len - m);
    }
}
static void init_networks(void) {
    for (int n = 3; n <= 8; n++) {
        g_batcher[n].n = 0;
        batcher_sort_rec(&g_batcher[n], n, 0, 1 << ilog2_(n));   /* sort.impala:63-65 */
    }
    for (int n = 3; n <= 4; n++) {
        g_bose_nelson[n].n = 0;
        bn_star(&g_bose_nelson[n], 0, n);
    }
}

/* Exposed so the tests can pin the comparator sequences. */
int oracle_network(int arity, int n, int8_t* a, int8_t* b) {
    pthread_once(&g_net_once, init_networks);
    const Network* w = arity == 8 ? &g_batcher[n] : &g_bose_nelson[n];
    memcpy(a, w->a, (size_t)w->n); memcpy(b, w->b, (size_t)w->n);
    return w->n;
}

/* ---- ray/triangle, src/traversal/intersection.impala:164-192 -------------
 * One lane of a Tri4 (src/traversal/mapping_cpu.impala:24-42).  Returns 1 on hit. */
static inline int intersect_ray_tri_lane(const Tri4* tp, int i,
                                         const float org[3], const float dir[3],
                                         float tmin, float tmax,
                                         float* out_t, float* out_u, float* out_v) {
    const float v0x = tp->v0[0][i], v0y = tp->v0[1][i], v0z = tp->v0[2][i];
    const float e1x = tp->e1[0][i], e1y = tp->e1[1][i], e1z = tp->e1[2][i];
    const float e2x = tp->e2[0][i], e2y = tp->e2[1][i], e2z = tp->e2[2][i];
    const float nx  = tp->n[0][i],  ny  = tp->n[1][i],  nz  = tp->n[2][i];
    /* c = v0 - org ; r = dir x c ; det = n . dir   (vector.impala:60-67) */
    const float cx = v0x - org[0], cy = v0y - org[1], cz = v0z - org[2];
    const float rx = dir[1] * cz - dir[2] * cy;
    const float ry = dir[2] * cx - dir[0] * cz;
    const float rz = dir[0] * cy - dir[1] * cx;
    const float det = nx * dir[0] + ny * dir[1] + nz * dir[2];
    const float abs_det = i2f(f2i(det) & 0x7FFFFFFF);

    const float u = prodsign(rx * e2x + ry * e2y + rz * e2z, det);
    int mask = u >= 0.0f;
    const float v = prodsign(rx * e1x + ry * e1y + rz * e1z, det);
    mask &= v >= 0.0f;
    mask &= (u + v) <= abs_det;
    if (!mask) return 0;

    const float t = prodsign(cx * nx + cy * ny + cz * nz, det);
    mask &= abs_det != 0.0f;                 /* no backface culling, mapping_cpu.impala:33 */
    mask &= t >= abs_det * tmin;
    mask &= t <= abs_det * tmax;
    if (!mask) return 0;

    const float inv_det = 1.0f / abs_det;
    *out_t = t * inv_det; *out_u = u * inv_det; *out_v = v * inv_det;
    return 1;
}

#if defined(__AVX2__)
/* The four lanes of a Tri4 at once: the same operations in the same order as intersect_ray_tri_lane (so the same bits),
 * lanes that miss or are invalid report (FLT_MAX, 0, 0).  Returns the hit mask. */
static inline int intersect_ray_tri4(const Tri4* tp, const float org[3], const float dir[3], float tmin, float tmax,
                                     float lt[4], float lu[4], float lv[4]) {
    const __m128 ox = _mm_set1_ps(org[0]), oy = _mm_set1_ps(org[1]), oz = _mm_set1_ps(org[2]);
    const __m128 dx = _mm_set1_ps(dir[0]), dy = _mm_set1_ps(dir[1]), dz = _mm_set1_ps(dir[2]);
    const __m128 e1x = _mm_loadu_ps(tp->e1[0]), e1y = _mm_loadu_ps(tp->e1[1]), e1z = _mm_loadu_ps(tp->e1[2]);
    const __m128 e2x = _mm_loadu_ps(tp->e2[0]), e2y = _mm_loadu_ps(tp->e2[1]), e2z = _mm_loadu_ps(tp->e2[2]);
    const __m128 nx = _mm_loadu_ps(tp->n[0]), ny = _mm_loadu_ps(tp->n[1]), nz = _mm_loadu_ps(tp->n[2]);
    const __m128 cx = _mm_sub_ps(_mm_loadu_ps(tp->v0[0]), ox), cy = _mm_sub_ps(_mm_loadu_ps(tp->v0[1]), oy), cz = _mm_sub_ps(_mm_loadu_ps(tp->v0[2]), oz);
    const __m128 rx = _mm_sub_ps(_mm_mul_ps(dy, cz), _mm_mul_ps(dz, cy));
    const __m128 ry = _mm_sub_ps(_mm_mul_ps(dz, cx), _mm_mul_ps(dx, cz));
    const __m128 rz = _mm_sub_ps(_mm_mul_ps(dx, cy), _mm_mul_ps(dy, cx));
#define DOT3(ax, ay, az, bx, by, bz) _mm_add_ps(_mm_add_ps(_mm_mul_ps(ax, bx), _mm_mul_ps(ay, by)), _mm_mul_ps(az, bz))
    const __m128 det = DOT3(nx, ny, nz, dx, dy, dz);
    const __m128 sign = _mm_and_ps(det, _mm_castsi128_ps(_mm_set1_epi32((int32_t)0x80000000u)));
    const __m128 abs_det = _mm_and_ps(det, _mm_castsi128_ps(_mm_set1_epi32(0x7FFFFFFF)));
    const __m128 u = _mm_xor_ps(DOT3(rx, ry, rz, e2x, e2y, e2z), sign);
    const __m128 v = _mm_xor_ps(DOT3(rx, ry, rz, e1x, e1y, e1z), sign);
    const __m128 zero = _mm_setzero_ps();
    __m128 m = _mm_and_ps(_mm_and_ps(_mm_cmpge_ps(u, zero), _mm_cmpge_ps(v, zero)), _mm_cmple_ps(_mm_add_ps(u, v), abs_det));
    m = _mm_and_ps(m, _mm_castsi128_ps(_mm_xor_si128(_mm_cmpeq_epi32(_mm_loadu_si128((const __m128i*)tp->prim_id), _mm_set1_epi32(-1)), _mm_set1_epi32(-1))));
    const __m128 t = _mm_xor_ps(DOT3(cx, cy, cz, nx, ny, nz), sign);
#undef DOT3
    m = _mm_and_ps(m, _mm_cmpneq_ps(abs_det, zero));
    m = _mm_and_ps(m, _mm_cmpge_ps(t, _mm_mul_ps(abs_det, _mm_set1_ps(tmin))));
    m = _mm_and_ps(m, _mm_cmple_ps(t, _mm_mul_ps(abs_det, _mm_set1_ps(tmax))));
    const int hm = _mm_movemask_ps(m);
    if (!hm) { for (int j = 0; j < 4; j++) { lt[j] = FLT_MAX_; lu[j] = 0.0f; lv[j] = 0.0f; } return 0; }
    const __m128 inv_det = _mm_div_ps(_mm_set1_ps(1.0f), abs_det);
    _mm_storeu_ps(lt, _mm_blendv_ps(_mm_set1_ps(FLT_MAX_), _mm_mul_ps(t, inv_det), m));
    _mm_storeu_ps(lu, _mm_and_ps(_mm_mul_ps(u, inv_det), m));
    _mm_storeu_ps(lv, _mm_and_ps(_mm_mul_ps(v, inv_det), m));
    return hm;
}
#endif

/* ---- the traversal kernel, src/traversal/mapping_cpu.impala:138-256 ------
 * N = arity (4 or 8); `nodes` is Node4* or Node8* (same field order, bounds[6][N],
 * child[N], pad[N]).  Stack discipline of src/traversal/stack.impala:52-123: the
 * top entry lives in (top_node, top_t); `st` is the memory part. */
static inline __attribute__((always_inline))
void traverse_single_from(const int N, const int any_hit,
                          const void* nodes_v, const Tri4* tris, const Ray1* rp, Hit1* hp,
                          OracleStats* stats, int32_t* geom_out, const int32_t root) {
    const size_t node_stride = (size_t)N * 32;       /* 6N floats + N + N ints */
    const char* nodes = (const char*)nodes_v;
    const Network* nets = N == 8 ? g_batcher : g_bose_nelson;

    /* make_cpu_ray1 + make_ray, bench_traversal.impala:85-95, intersection.impala:88-99 */
    const float org[3] = { rp->org[0], rp->org[1], rp->org[2] };
    const float dir[3] = { rp->dir[0], rp->dir[1], rp->dir[2] };
    const float tmin = rp->tmin;
    float tmax = rp->tmax;
    const float idx = safe_rcp(dir[0]), idy = safe_rcp(dir[1]), idz = safe_rcp(dir[2]);
    const float iox = -(org[0] * idx), ioy = -(org[1] * idy), ioz = -(org[2] * idz);
    /* ray_octant, intersection.impala:128-132 ; ordered_bbox, mapping_cpu.impala:88-106 */
    const int ox = dir[0] > 0.0f ? N : 0, oy = dir[1] > 0.0f ? N : 0, oz = dir[2] > 0.0f ? N : 0;
    const int near_x = N - ox, far_x = ox;
    const int near_y = 3 * N - oy, far_y = 2 * N + oy;
    const int near_z = 5 * N - oz, far_z = 4 * N + oz;

    /* empty_hit, intersection.impala:134-136 (u, v undefined there; 0 here) */
    int32_t hit_prim = -1, hit_geom = -1; float hit_t = tmax, hit_u = 0.0f, hit_v = 0.0f;

    NodeRef st[STACK_SIZE + 8];
    int ptr = -1;
    int32_t top_node = 0; float top_t = FLT_MAX_;
#define PUSH(n_, t_)       do { ++ptr; st[ptr].node = top_node; st[ptr].tmin = top_t; top_node = (n_); top_t = (t_); } while (0)
#define PUSH_AFTER(n_, t_) do { ++ptr; st[ptr].node = (n_); st[ptr].tmin = (t_); } while (0)
#define POP()              do { top_node = st[ptr].node; top_t = st[ptr].tmin; --ptr; } while (0)
    PUSH(root, tmin);                                                    /* :153 (root = 1 but for the hybrid kernel's lanes) */

    uint64_t n_nodes = 0, n_tri4 = 0; int max_ptr = 0;

    for (;;) {
        if (top_node == 0) break;                                        /* :168 */
        if (!any_hit && top_t > tmax) { POP(); ORACLE_TRACE('c'); continue; }   /* :170-174 */

        int restart = 0;
        while (top_node > 0) {                                           /* :177 */
            const float* nb = (const float*)(nodes + (size_t)(top_node - 1) * node_stride);
            const int32_t* child = (const int32_t*)(nb + 6 * N);
            POP();
            n_nodes++;
            ORACLE_TRACE('N');

            /* intersect_ray_box ordered, intersection.impala:194-208, integer min/max */
            float tentry[8]; int mask = 0;
#if defined(__AVX2__)
            if (N == 8) {
                /* The same operations on eight children at once (what the reference's RV-vectorised CPU build does): one
                 * multiply and one add per plane, never fused (-ffp-contract=off), integer min / max in the scalar nesting.
                 * Bit-identical to the loop below, NaNs included (same SSE default NaN). */
                const __m256 vix = _mm256_set1_ps(idx), viy = _mm256_set1_ps(idy), viz = _mm256_set1_ps(idz);
                const __m256 vox = _mm256_set1_ps(iox), voy = _mm256_set1_ps(ioy), voz = _mm256_set1_ps(ioz);
                const __m256 t0x = _mm256_add_ps(_mm256_mul_ps(vix, _mm256_loadu_ps(nb + near_x)), vox);
                const __m256 t0y = _mm256_add_ps(_mm256_mul_ps(viy, _mm256_loadu_ps(nb + near_y)), voy);
                const __m256 t0z = _mm256_add_ps(_mm256_mul_ps(viz, _mm256_loadu_ps(nb + near_z)), voz);
                const __m256 t1x = _mm256_add_ps(_mm256_mul_ps(vix, _mm256_loadu_ps(nb + far_x)), vox);
                const __m256 t1y = _mm256_add_ps(_mm256_mul_ps(viy, _mm256_loadu_ps(nb + far_y)), voy);
                const __m256 t1z = _mm256_add_ps(_mm256_mul_ps(viz, _mm256_loadu_ps(nb + far_z)), voz);
#define CI(x) _mm256_castps_si256(x)
                const __m256i te = _mm256_max_epi32(_mm256_max_epi32(CI(t0x), CI(t0y)), _mm256_max_epi32(CI(t0z), CI(_mm256_set1_ps(tmin))));
                const __m256i tx = _mm256_min_epi32(_mm256_min_epi32(CI(t1x), CI(t1y)), _mm256_min_epi32(CI(t1z), CI(_mm256_set1_ps(tmax))));
#undef CI
                _mm256_storeu_ps(tentry, _mm256_castsi256_ps(te));
                mask = ~_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(te, tx))) & 0xFF;      /* :184 */
            } else
#endif
            for (int i = 0; i < N; i++) {
                const float t0x = idx * nb[near_x + i] + iox;
                const float t0y = idy * nb[near_y + i] + ioy;
                const float t0z = idz * nb[near_z + i] + ioz;
                const float t1x = idx * nb[far_x + i] + iox;
                const float t1y = idy * nb[far_y + i] + ioy;
                const float t1z = idz * nb[far_z + i] + ioz;
                const float te = imax(imax(t0x, t0y), imax(t0z, tmin));
                const float tx = imin(imin(t1x, t1y), imin(t1z, tmax));
                tentry[i] = te;
                mask |= (f2i(tx) < f2i(te) ? 0 : 1) << i;                /* :184 */
            }
            if (mask == 0) {                                             /* :189-191 */
                ORACLE_TRACE('0');
                if (any_hit) continue;
                restart = 1; break;
            }

            int num_intrs = 0, tops = 0;                                 /* :195-208 */
            for (int m = mask; m != 0; m &= m - 1) {
                const int lane = __builtin_ctz((unsigned)m);
                const int32_t child_id = child[lane];
                const float t = tentry[lane];
                num_intrs++;
                if (any_hit || t < top_t) { PUSH(child_id, t); tops |= 1 << lane; }
                else                      PUSH_AFTER(child_id, t);
            }
            ORACLE_TRACE_PUSHES(mask, tops); (void)tops;
            if (ptr > max_ptr) max_ptr = ptr;
            ORACLE_TRACE('0' + num_intrs);

            if (!any_hit && num_intrs >= 3) {                            /* :210-218, stack.impala:79-111 */
                NodeRef* e = &st[ptr - num_intrs + 1];
                const Network* w = &nets[num_intrs];
                for (int c = 0; c < w->n; c++) {
                    const int i = w->a[c], j = w->b[c];
                    if (e[i].tmin < e[j].tmin) { NodeRef tmp = e[i]; e[i] = e[j]; e[j] = tmp; }
                }
            }
        }
        if (restart) continue;

        if (top_node == 0) break;   /* :221 for any_hit; for closest-hit the reference would read
                                       prim -1 here (only reachable with a non-finite/FLT_MAX entry key) */

        int32_t prim_id = ~top_node;                                     /* :224 */
        POP();
        int terminated = 0;
        for (;;) {
            const Tri4* tp = &tris[prim_id++];
            n_tri4++;
            ORACLE_TRACE('L');
            float lt[4], lu[4], lv[4]; int hm = 0;
#if defined(__AVX2__)
            hm = intersect_ray_tri4(tp, org, dir, tmin, tmax, lt, lu, lv);
#else
            for (int j = 0; j < 4; j++) {
                lt[j] = FLT_MAX_; lu[j] = 0.0f; lv[j] = 0.0f;
                if (tp->prim_id[j] == -1) continue;                      /* is_valid, mapping_cpu.impala:39 */
                if (intersect_ray_tri_lane(tp, j, org, dir, tmin, tmax, &lt[j], &lu[j], &lv[j]))
                    hm |= 1 << j;
                else
                    lt[j] = FLT_MAX_;
            }
#endif
            if (hm) {
                int lane;
                if (any_hit) {
                    lane = __builtin_ctz((unsigned)hm);                  /* :234-237 */
                    terminated = 1;
                } else {
                    /* cpu_reduce with the integer min + cpu_index_of, cpu_common.impala:37-56, :239-243 */
                    const float mn = imin(imin(lt[0], lt[2]), imin(lt[1], lt[3]));
                    lane = 0;
                    while (!(lt[lane] == mn)) lane++;
                }
                hit_prim = tp->prim_id[lane] & 0x7FFFFFFF;               /* mapping_cpu.impala:34 */
                hit_geom = tp->geom_id[lane];
                hit_t = lt[lane]; hit_u = lu[lane]; hit_v = lv[lane];
                if (!any_hit) tmax = hit_t;                              /* :243 */
            }
            if (tp->prim_id[3] < 0) break;                               /* is_last, mapping_cpu.impala:40 */
        }
        if (any_hit && terminated) break;                                /* :252 */
    }
#undef PUSH
#undef PUSH_AFTER
#undef POP

    /* make_cpu_hit1, bench_traversal.impala:121-131 */
    hp->tri_id = hit_prim;
    if (geom_out) *geom_out = hit_geom;
    if (!any_hit) { hp->t = hit_t; hp->u = hit_u; hp->v = hit_v; }
    if (stats) {
        stats->nodes += n_nodes; stats->tri4 += n_tri4;
        if ((uint64_t)max_ptr > stats->max_stack) stats->max_stack = (uint64_t)max_ptr;
    }
}

static inline __attribute__((always_inline))
void traverse_single(const int N, const int any_hit,
                     const void* nodes_v, const Tri4* tris, const Ray1* rp, Hit1* hp,
                     OracleStats* stats, int32_t* geom_out) {
    traverse_single_from(N, any_hit, nodes_v, tris, rp, hp, stats, geom_out, 1 /*root*/);
}

/* ---- the packet / hybrid kernel, src/traversal/mapping_cpu.impala:259-384 ---------------------------------------------
 * One packet of W rays (W = 4 or 8 lanes of the reference's vectorised region) with ONE stack of (node, tmin[W]):
 * a node is visited when any lane enters it, the children are pushed in child order (no sort), the triangles of a leaf
 * are tested one after the other (Tri4 lane i against every ray lane; `t <= tmax` accepts an equal distance, so of two
 * triangles hit at the same t the LATER one stays -- the single-ray kernel keeps the FIRST of a Tri4).  `hybrid`: when
 * at most `switch_threshold` lanes are still interested in the entry on top, those lanes walk its subtree with the
 * single-ray kernel (:309-319).  Records therefore differ from the single-ray kernel's where hits tie; how often they do
 * on the reference's ray sets is measured in tests/test_packet_oracle.py.
 * rays: org[3][W] dir[3][W] tmin[W] tmax[W]; hits: tri_id[W] t[W] u[W] v[W] (make_cpu_ray4/8, make_cpu_hit4/8,
 * tools/bench_traversal/bench_traversal.impala:96-157). */
static void traverse_packet(const int N, const int W, const int hybrid, const int any_hit,
                            const void* nodes_v, const Tri4* tris, const float* rp, float* hp) {
    const size_t node_stride = (size_t)N * 32;
    const char* nodes = (const char*)nodes_v;
    const int switch_threshold = W == 4 ? 3 : W == 8 ? (N == 4 ? 4 : 6) : W / 2;           /* :268-273 */
    float org[8][3], dir[8][3], inv_dir[8][3], inv_org[8][3], tmin[8], tmax[8];
    int32_t hit_prim[8]; float hit_t[8], hit_u[8], hit_v[8]; int terminated[8];
    for (int l = 0; l < W; l++) {                                        /* make_ray, intersection.impala:88-99 */
        for (int a = 0; a < 3; a++) {
            org[l][a] = rp[a * W + l]; dir[l][a] = rp[(3 + a) * W + l];
            inv_dir[l][a] = safe_rcp(dir[l][a]);
            inv_org[l][a] = -(org[l][a] * inv_dir[l][a]);
        }
        tmin[l] = rp[6 * W + l]; tmax[l] = rp[7 * W + l];
        hit_prim[l] = -1; hit_t[l] = tmax[l]; hit_u[l] = 0.0f; hit_v[l] = 0.0f; terminated[l] = 0;      /* empty_hit */
    }
    int32_t st_node[STACK_SIZE + 8]; float st_t[STACK_SIZE + 8][8];
    int ptr = -1; int32_t top_node = 0; float top_t[8];
    for (int l = 0; l < W; l++) top_t[l] = FLT_MAX_;
#define PPUSH(n_, t_)       do { ++ptr; st_node[ptr] = top_node; memcpy(st_t[ptr], top_t, sizeof top_t); top_node = (n_); memcpy(top_t, (t_), sizeof top_t); } while (0)
#define PPUSH_AFTER(n_, t_) do { ++ptr; st_node[ptr] = (n_); memcpy(st_t[ptr], (t_), sizeof top_t); } while (0)
#define PPOP()              do { top_node = st_node[ptr]; memcpy(top_t, st_t[ptr], sizeof top_t); --ptr; } while (0)
    { float t0[8]; for (int l = 0; l < 8; l++) t0[l] = l < W ? tmin[l] : FLT_MAX_; PPUSH(1 /*root*/, t0); }    /* :278 */

    for (;;) {
        /* cull, and hand the last interested lanes to the single-ray kernel (:303-326) */
        int done = 0;
        for (;;) {
            if (top_node == 0) { done = 1; break; }
            unsigned mask = 0;
            for (int l = 0; l < W; l++) if (top_t[l] <= tmax[l] && !terminated[l]) mask |= 1u << l;
            if (mask != 0) {
                if (hybrid && __builtin_popcount(mask) <= switch_threshold) {
                    for (unsigned m = mask; m != 0; m &= m - 1) {
                        const int l = __builtin_ctz(m);
                        const Ray1 lane_ray = { { org[l][0], org[l][1], org[l][2] }, tmin[l], { dir[l][0], dir[l][1], dir[l][2] }, tmax[l] };
                        Hit1 lane_hit = { -1, tmax[l], 0.0f, 0.0f };
                        traverse_single_from(N, any_hit, nodes_v, tris, &lane_ray, &lane_hit, NULL, NULL, top_node);
                        if (lane_hit.tri_id >= 0) {
                            hit_prim[l] = lane_hit.tri_id;
                            if (!any_hit) { hit_t[l] = lane_hit.t; hit_u[l] = lane_hit.u; hit_v[l] = lane_hit.v; tmax[l] = lane_hit.t; }
                        }
                    }
                    if (any_hit) for (int l = 0; l < W; l++) terminated[l] = hit_prim[l] >= 0;
                } else {
                    break;
                }
            }
            PPOP();
        }
        if (done) break;

        /* inner nodes (:329-355) */
        int culled = 0;
        while (top_node > 0) {
            const float* nb = (const float*)(nodes + (size_t)(top_node - 1) * node_stride);
            const int32_t* child = (const int32_t*)(nb + 6 * N);
            PPOP();
            int pushed = 0;
            for (int i = 0; i < N; i++) {
                const int32_t child_id = child[i];
                if (child_id == 0) break;
                float thit[8]; int any = 0, nearer = 0;
                for (int l = 0; l < W; l++) {                            /* intersect_ray_box, unordered, integer min / max */
                    const float t0x = inv_dir[l][0] * nb[i] + inv_org[l][0],         t1x = inv_dir[l][0] * nb[N + i] + inv_org[l][0];
                    const float t0y = inv_dir[l][1] * nb[2 * N + i] + inv_org[l][1], t1y = inv_dir[l][1] * nb[3 * N + i] + inv_org[l][1];
                    const float t0z = inv_dir[l][2] * nb[4 * N + i] + inv_org[l][2], t1z = inv_dir[l][2] * nb[5 * N + i] + inv_org[l][2];
                    const float tentry = imax(imax(imin(t0x, t1x), imin(t0y, t1y)), imax(imin(t0z, t1z), tmin[l]));
                    const float texit  = imin(imin(imax(t0x, t1x), imax(t0y, t1y)), imin(imax(t0z, t1z), tmax[l]));
                    const int miss = f2i(texit) < f2i(tentry);
                    thit[l] = miss ? FLT_MAX_ : tentry;
                    any |= !miss;
                }
                for (int l = W; l < 8; l++) thit[l] = FLT_MAX_;
                if (any) {
                    for (int l = 0; l < W; l++) nearer |= top_t[l] > thit[l];
                    if (any_hit || nearer) PPUSH(child_id, thit);
                    else                   PPUSH_AFTER(child_id, thit);
                    pushed = 1;
                }
            }
            if (!pushed) { culled = 1; break; }
        }
        if (culled) continue;

        /* leaf (:357-381) */
        if (top_node < 0) {
            int active[8];
            for (int l = 0; l < W; l++) active[l] = top_t[l] <= tmax[l] && !terminated[l];
            int32_t prim_id = ~top_node;
            PPOP();
            int all_out = 0;
            for (;;) {
                const Tri4* tp = &tris[prim_id++];
                for (int i = 0; i < 4; i++) {
                    if (tp->prim_id[i] == -1) break;                     /* is_valid */
                    for (int l = 0; l < W; l++) {
                        if (!active[l]) continue;
                        float t, u, v;
                        if (intersect_ray_tri_lane(tp, i, org[l], dir[l], tmin[l], tmax[l], &t, &u, &v)) {
                            hit_prim[l] = tp->prim_id[i] & 0x7FFFFFFF; hit_t[l] = t; hit_u[l] = u; hit_v[l] = v;
                            tmax[l] = t;
                            if (any_hit) { terminated[l] = 1; active[l] = 0; }
                        }
                    }
                    if (any_hit) { int all = 1; for (int l = 0; l < W; l++) all &= terminated[l]; if (all) { all_out = 1; break; } }
                }
                if (all_out) break;
                if (tp->prim_id[3] < 0) break;                           /* is_last */
            }
            if (all_out) break;
        }
    }
#undef PPUSH
#undef PPUSH_AFTER
#undef PPOP
    for (int l = 0; l < W; l++) {                                        /* make_cpu_hit4/8 */
        memcpy(&hp[l], &hit_prim[l], 4);
        if (!any_hit) { hp[W + l] = hit_t[l]; hp[2 * W + l] = hit_u[l]; hp[3 * W + l] = hit_v[l]; }
    }
}

typedef struct { int arity, width, hybrid, any_hit; const void* nodes; const Tri4* tris; const float* rays; float* hits; int32_t begin, end; } PacketJob;
static void* run_packets_thread(void* p) {
    PacketJob* j = (PacketJob*)p;
    for (int32_t i = j->begin; i < j->end; i++)
        traverse_packet(j->arity, j->width, j->hybrid, j->any_hit, j->nodes, j->tris, j->rays + (size_t)i * 8 * j->width, j->hits + (size_t)i * 4 * j->width);
    return NULL;
}
/* cpu_{intersect,occluded}_{packet,hybrid}_ray{4,8}_bvh{4,8}_tri4 of the reference, by name of their parameters */
void oracle_traverse_packets(int arity, int width, int hybrid, int any_hit, const void* nodes, const Tri4* tris,
                             const float* rays, float* hits, int32_t num_packets, int threads) {
    pthread_once(&g_net_once, init_networks);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    PacketJob jobs[256]; pthread_t th[256];
    for (int k = 0; k < threads; k++) {
        jobs[k] = (PacketJob){ arity, width, hybrid, any_hit, nodes, tris, rays, hits,
                               (int32_t)((int64_t)num_packets * k / threads), (int32_t)((int64_t)num_packets * (k + 1) / threads) };
        pthread_create(&th[k], NULL, run_packets_thread, &jobs[k]);
    }
    for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
}

/* ---- cpu_traverse_single, src/traversal/mapping_cpu.impala:404-418 -------- */
typedef struct {
    int arity, any_hit;
    const void* nodes; const Tri4* tris; const Ray1* rays; Hit1* hits;
    int32_t begin, end;
    OracleStats stats; int want_stats;
} Job;

static void run_range(Job* j) {
    OracleStats* s = j->want_stats ? &j->stats : NULL;
    if (j->arity == 8) {
        if (j->any_hit) for (int32_t i = j->begin; i < j->end; i++) traverse_single(8, 1, j->nodes, j->tris, &j->rays[i], &j->hits[i], s, NULL);
        else            for (int32_t i = j->begin; i < j->end; i++) traverse_single(8, 0, j->nodes, j->tris, &j->rays[i], &j->hits[i], s, NULL);
    } else {
        if (j->any_hit) for (int32_t i = j->begin; i < j->end; i++) traverse_single(4, 1, j->nodes, j->tris, &j->rays[i], &j->hits[i], s, NULL);
        else            for (int32_t i = j->begin; i < j->end; i++) traverse_single(4, 0, j->nodes, j->tris, &j->rays[i], &j->hits[i], s, NULL);
    }
}
static void* run_range_thread(void* p) { run_range((Job*)p); return NULL; }

/* Contiguous ray ranges over `threads` host threads (the reference loop is
 * serial; threads > 1 is this repo's "all host cores" baseline). */
void oracle_traverse(int arity, int any_hit, const void* nodes, const Tri4* tris,
                     const Ray1* rays, Hit1* hits, int32_t num_rays, int threads,
                     OracleStats* stats) {
    pthread_once(&g_net_once, init_networks);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    Job jobs[256]; pthread_t th[256];
    for (int k = 0; k < threads; k++) {
        Job* j = &jobs[k];
        j->arity = arity; j->any_hit = any_hit; j->nodes = nodes; j->tris = tris; j->rays = rays; j->hits = hits;
        j->begin = (int32_t)((int64_t)num_rays * k / threads);
        j->end   = (int32_t)((int64_t)num_rays * (k + 1) / threads);
        memset(&j->stats, 0, sizeof j->stats); j->want_stats = stats != NULL;
    }
    if (threads == 1) run_range(&jobs[0]);
    else {
        for (int k = 0; k < threads; k++) pthread_create(&th[k], NULL, run_range_thread, &jobs[k]);
        for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int k = 0; k < threads; k++) {
            stats->nodes += jobs[k].stats.nodes; stats->tri4 += jobs[k].stats.tri4;
            if (jobs[k].stats.max_stack > stats->max_stack) stats->max_stack = jobs[k].stats.max_stack;
        }
    }
}

/* The reference's exported names, tools/bench_traversal/bench_traversal.impala:279-305,429-455 */
void cpu_intersect_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    oracle_traverse(8, 0, nodes, tris, rays, hits, num_packets, 1, NULL);
}
void cpu_occluded_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    oracle_traverse(8, 1, nodes, tris, rays, hits, num_packets, 1, NULL);
}
void cpu_intersect_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    oracle_traverse(4, 0, nodes, tris, rays, hits, num_packets, 1, NULL);
}
void cpu_occluded_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    oracle_traverse(4, 1, nodes, tris, rays, hits, num_packets, 1, NULL);
}

/* Brute force over every Tri4 lane in array order: closest hit with the same
 * acceptance rule; a self-check for the traversal above. */
void oracle_brute_force(const Tri4* tris, int32_t num_tri4, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    for (int32_t r = 0; r < num_rays; r++) {
        const Ray1* rp = &rays[r];
        const float org[3] = { rp->org[0], rp->org[1], rp->org[2] };
        const float dir[3] = { rp->dir[0], rp->dir[1], rp->dir[2] };
        float tmax = rp->tmax; Hit1 h = { -1, tmax, 0.0f, 0.0f };
        for (int32_t k = 0; k < num_tri4; k++)
            for (int j = 0; j < 4; j++) {
                float t, u, v;
                if (tris[k].prim_id[j] == -1) continue;
                if (intersect_ray_tri_lane(&tris[k], j, org, dir, rp->tmin, tmax, &t, &u, &v) && t < h.t) {
                    h.tri_id = tris[k].prim_id[j] & 0x7FFFFFFF; h.t = t; h.u = u; h.v = v; tmax = t;
                }
            }
        hits[r] = h;
    }
}
