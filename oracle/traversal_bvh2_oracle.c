/*
 * traversal_bvh2_oracle.c -- CPU restatement of the reference's GPU single-ray traversal over its BVH2 / Tri1 layout
 * (what nvvm_{intersect,occluded}_single_ray1_bvh2_tri1 run).  TEST INFRASTRUCTURE ONLY: nothing under rodent_b200/ may
 * link or call this; it checks cuda_{intersect,occluded}_single_ray1_bvh2_tri1 (tests/).
 *
 * Parity status: PINNED for the hit distance t by the reference's golden images testing/ref-primary.png and
 * testing/ref-random.png, on the reference's own BVH2 block of testing/sponza.bvh (tests/test_bvh2.py); tri_id / u / v
 * have no golden vector, for those this file is the definition.
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest, no FMA contraction, source order.  The reference runs this
 * code on NVIDIA hardware, whose arithmetic instructions return the canonical NaN 0x7FFFFFFF (x86 would give
 * 0xFFC00000 for inf - inf); the bit pattern matters because the slab test compares float bits as integers
 * (make_nvvm_min_max, src/traversal/mapping_gpu.impala:74-85), so every arithmetic result is canonicalised the NVIDIA way
 * here.  NaNs arise for rays parallel to an axis: inv_dir = +-FLT_MAX, inv_dir * bound = +-inf, inv_org = -+inf.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>

#include "../include/rodent_b200.h"

#define FLT_MAX_ 3.4028234664e+38f  /* src/core/common.impala:4 */
#define STACK_SIZE 64               /* src/traversal/stack.impala:53-54 */

static inline int32_t f2i(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float   i2f(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
static inline float nv(float x) { return x != x ? i2f(0x7FFFFFFF) : x; }       /* NVIDIA's canonical NaN */

/* src/core/common.impala:78-85 */
static inline float prodsign(float x, float y) { return i2f(f2i(x) ^ (f2i(y) & (int32_t)0x80000000u)); }
static inline float safe_rcp(float x) {
    const float ax = x > 0.0f ? x : -x;
    return ax < 1e-8f ? prodsign(FLT_MAX_, x) : 1.0f / x;
}
/* make_nvvm_min_max, mapping_gpu.impala:74-85: fminf / fmaxf are the float instructions (a NaN operand loses), the
 * three-operand forms are vmin / vmax on the bits as signed integers */
static inline float fmin_nv(float a, float b) { return nv(fminf(a, b)); }
static inline float fmax_nv(float a, float b) { return nv(fmaxf(a, b)); }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline float minmin(float a, float b, float c) { return i2f(imin(imin(f2i(a), f2i(b)), f2i(c))); }
static inline float maxmax(float a, float b, float c) { return i2f(imax(imax(f2i(a), f2i(b)), f2i(c))); }
static inline float minmax(float a, float b, float c) { return i2f(imax(imin(f2i(a), f2i(b)), f2i(c))); }
static inline float maxmin(float a, float b, float c) { return i2f(imin(imax(f2i(a), f2i(b)), f2i(c))); }

typedef struct { float org[3], dir[3], inv_dir[3], inv_org[3], tmin, tmax; } Ray;

/* intersect_ray_box, unordered: src/traversal/intersection.impala:194-208; box k of a Node2: mapping_gpu.impala:33-43 */
static inline int hit_box(const Ray* r, const float* b, float* tentry) {
    float t0[3], t1[3];
    for (int a = 0; a < 3; a++) {
        t0[a] = nv(nv(r->inv_dir[a] * b[2 * a]) + r->inv_org[a]);
        t1[a] = nv(nv(r->inv_dir[a] * b[2 * a + 1]) + r->inv_org[a]);
    }
    const float te = maxmax(fmin_nv(t0[0], t1[0]), fmin_nv(t0[1], t1[1]), minmax(t0[2], t1[2], r->tmin));
    const float tx = minmin(fmax_nv(t0[0], t1[0]), fmax_nv(t0[1], t1[1]), maxmin(t0[2], t1[2], r->tmax));
    *tentry = te;
    return te <= tx;                                                     /* mapping_gpu.impala:114 */
}

/* intersect_ray_tri, intersection.impala:164-192, with the triangle of make_gpu_bvh2_tri1 (mapping_gpu.impala:52-60):
 * n = cross(e1, e2) computed per test (vector.impala:60-67).  Returns 1 on hit. */
static inline int hit_tri(const Ray* r, const Tri1* tp, float* out_t, float* out_u, float* out_v) {
    const float* e1 = tp->e1; const float* e2 = tp->e2;
    const float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    const float cx = tp->v0[0] - r->org[0], cy = tp->v0[1] - r->org[1], cz = tp->v0[2] - r->org[2];
    const float rx = r->dir[1] * cz - r->dir[2] * cy, ry = r->dir[2] * cx - r->dir[0] * cz, rz = r->dir[0] * cy - r->dir[1] * cx;
    const float det = nx * r->dir[0] + ny * r->dir[1] + nz * r->dir[2];
    const float abs_det = i2f(f2i(det) & 0x7FFFFFFF);
    const float u = prodsign(rx * e2[0] + ry * e2[1] + rz * e2[2], det);
    const float v = prodsign(rx * e1[0] + ry * e1[1] + rz * e1[2], det);
    if (!(u >= 0.0f && v >= 0.0f && u + v <= abs_det)) return 0;
    const float t = prodsign(cx * nx + cy * ny + cz * nz, det);
    if (!(abs_det != 0.0f && t >= abs_det * r->tmin && t <= abs_det * r->tmax)) return 0;
    const float inv_det = 1.0f / abs_det;
    *out_t = t * inv_det; *out_u = u * inv_det; *out_v = v * inv_det;
    return 1;
}

/* gpu_traverse_single_helper, mapping_gpu.impala:94-178, arity 2.  The stack keeps its top in a register and never
 * looks at the entry distances on this path (stack.impala:52-123 with undef keys). */
static void traverse_bvh2(int any_hit, const Node2* nodes, const Tri1* tris, const Ray1* rp, Hit1* hp, uint64_t* counters, int32_t* geom_out) {
    Ray ray;
    for (int a = 0; a < 3; a++) {                                        /* make_gpu_ray1 + make_ray, intersection.impala:88-99 */
        ray.org[a] = rp->org[a]; ray.dir[a] = rp->dir[a];
        ray.inv_dir[a] = safe_rcp(rp->dir[a]);
        ray.inv_org[a] = -nv(rp->org[a] * ray.inv_dir[a]);
    }
    ray.tmin = rp->tmin; ray.tmax = rp->tmax;
    Hit1 hit = { -1, ray.tmax, 0.0f, 0.0f };                             /* empty_hit, :134-136 (u, v undefined there) */
    int32_t geom = -1;

    int32_t mem[STACK_SIZE + 8];
    int ptr = 0; int32_t top = 1; mem[0] = 0;                            /* push(root) onto the empty stack, :103 */
    while (top != 0) {                                                   /* :105 */
        const Node2* node = &nodes[top - 1];
        float t0, t1;
        const int h0 = hit_box(&ray, node->bounds, &t0), h1 = hit_box(&ray, node->bounds + 6, &t1);
        if (counters) counters[0]++;
        if (!h0 && !h1) top = mem[ptr--];                                /* :122-123 */
        else if (h0 && h1) {                                             /* :127-131 */
            const int first0 = t0 < t1;
            top = first0 ? node->child[0] : node->child[1];
            mem[++ptr] = first0 ? node->child[1] : node->child[0];
        } else top = h0 ? node->child[0] : node->child[1];               /* :133 */

        while (top < 0) {                                                /* :155-174 */
            int32_t k = ~top;
            top = mem[ptr--];
            for (;;) {
                const Tri1* tp = &tris[k++];
                if (counters) counters[1]++;
                float t, u, v;
                if (hit_tri(&ray, tp, &t, &u, &v)) {
                    hit.tri_id = tp->prim_id & 0x7FFFFFFF; hit.t = t; hit.u = u; hit.v = v; geom = tp->geom_id;
                    ray.tmax = t;
                    if (any_hit) { *hp = hit; if (geom_out) *geom_out = geom; return; }   /* :168 */
                }
                if (tp->prim_id < 0) break;                              /* is_last, :62 */
            }
        }
    }
    *hp = hit;                                                           /* make_gpu_hit1, bench_traversal.impala:78-83 */
    if (geom_out) *geom_out = geom;
}

/* One ray, with the geometry id of the hit (make_hit(geom_id, ...), mapping_gpu.impala:60): for the path-tracing oracle. */
void oracle_bvh2_trace_one(int any_hit, const Node2* nodes, const Tri1* tris, const Ray1* ray, Hit1* hit, int32_t* geom, uint64_t* counters) {
    traverse_bvh2(any_hit, nodes, tris, ray, hit, counters, geom);      /* counters (optional): [0] += Node2 visited, [1] += Tri1 tested */
}

typedef struct { int any_hit; const Node2* nodes; const Tri1* tris; const Ray1* rays; Hit1* hits; int32_t begin, end; uint64_t counters[2]; } Job;
static void* run_range(void* p) {
    Job* j = (Job*)p;
    for (int32_t i = j->begin; i < j->end; i++) traverse_bvh2(j->any_hit, j->nodes, j->tris, &j->rays[i], &j->hits[i], j->counters, NULL);
    return NULL;
}

/* counters (optional): [0] inner nodes visited, [1] triangles tested */
void oracle_traverse_bvh2(int any_hit, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits, int32_t num_rays,
                          int threads, uint64_t* counters) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    Job jobs[256]; pthread_t th[256];
    for (int k = 0; k < threads; k++) {
        Job* j = &jobs[k];
        j->any_hit = any_hit; j->nodes = nodes; j->tris = tris; j->rays = rays; j->hits = hits;
        j->begin = (int32_t)((int64_t)num_rays * k / threads); j->end = (int32_t)((int64_t)num_rays * (k + 1) / threads);
        j->counters[0] = j->counters[1] = 0;
    }
    for (int k = 0; k < threads; k++) pthread_create(&th[k], NULL, run_range, &jobs[k]);
    for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
    if (counters) { counters[0] = counters[1] = 0; for (int k = 0; k < threads; k++) { counters[0] += jobs[k].counters[0]; counters[1] += jobs[k].counters[1]; } }
}
