/*
 * rodent_b200.h -- C ABI of the B200-native traversal / path-tracing core.
 *
 * This is the drop-in boundary for the one hot path of AnyDSL/rodent: BVH
 * traversal + ray/triangle test (bench_traversal) and the wavefront render loop
 * (rodent).  Every entry point is `extern "C"`, takes plain pointers and sizes,
 * and names the reference interface it stands in for (paths are relative to the
 * reference tree).
 *
 * Data layouts are the reference's, byte for byte, so buffers written by the
 * reference's loaders (tools/common/load_bvh.h, load_rays.h) can be passed
 * through unchanged.
 */
#ifndef RODENT_B200_H
#define RODENT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Layouts ------------------------------------------------------------ */

/* src/traversal/mapping_cpu.impala:18-22 (Impala [[f32*8]*6] == C float[6][8]).
 * bounds rows: lo_x, hi_x, lo_y, hi_y, lo_z, hi_z; child > 0: inner node id
 * (1-based), child < 0: leaf, ~child = first Tri4; child == 0: empty lane whose
 * box is (+inf, -inf) (src/driver/converter.cpp:185-195). */
typedef struct Node8 {
    float   bounds[6][8];
    int32_t child[8];
    int32_t pad[8];
} Node8;

/* src/traversal/mapping_cpu.impala:12-16 */
typedef struct Node4 {
    float   bounds[6][4];
    int32_t child[4];
    int32_t pad[4];
} Node4;

/* src/traversal/mapping_cpu.impala:3-10.  e1 = v0 - v1, e2 = v2 - v0,
 * n = cross(e1, e2).  prim_id == -1: invalid lane; prim_id[3] < 0: last packet
 * of its leaf; the reported id is prim_id & 0x7FFFFFFF. */
typedef struct Tri4 {
    float   v0[3][4];
    float   e1[3][4];
    float   e2[3][4];
    float   n[3][4];
    int32_t prim_id[4];
    int32_t geom_id[4];
} Tri4;

/* tools/bench_traversal/bench_traversal.impala:25-30 */
typedef struct Ray1 {
    float org[3];
    float tmin;
    float dir[3];
    float tmax;
} Ray1;

/* tools/bench_traversal/bench_traversal.impala:46-51.  Miss: tri_id = -1,
 * t = tmax, u/v undefined in the reference; this library writes u = v = 0. */
typedef struct Hit1 {
    int32_t tri_id;
    float   t;
    float   u;
    float   v;
} Hit1;

/* ---- Traversal, device pointers ----------------------------------------- *
 * Same contract as the reference's GPU exports
 *   nvvm_{intersect,occluded}_single_ray1_bvh2_tri1(dev, nodes, tris, rays, hits, num_rays)
 *   (tools/bench_traversal/bench_traversal.impala:459-493, called from
 *   tools/bench_traversal/bench_traversal.cpp:124-135)
 * but on the BVH8/Tri4 layout of the CPU single-ray path they must agree with
 * (cpu_{intersect,occluded}_single_ray1_bvh8_tri4, bench_traversal.impala:429-455).
 * All four arrays are DEVICE pointers on CUDA device `dev`, owned by the caller;
 * the call returns after the device finished (reference: device.sync(), :473).
 * `occluded` writes tri_id only semantics of make_cpu_hit1(any_hit=true)
 * (:121-131): tri_id >= 0 iff something was hit; t/u/v are left untouched.
 * Errors: CUDA failures print file:line and abort(), as the reference does
 * (bench_traversal.impala:15-21, src/driver/common.h:43-46). */
void cuda_intersect_single_ray1_bvh8_tri4(int32_t dev, const Node8* nodes, const Tri4* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_rays);
void cuda_occluded_single_ray1_bvh8_tri4(int32_t dev, const Node8* nodes, const Tri4* tris,
                                         const Ray1* rays, Hit1* hits, int32_t num_rays);

/* Asynchronous forms: enqueue on `stream` (a cudaStream_t passed as void*; NULL
 * = the legacy default stream) and return without synchronising.  `work_counter`
 * is a caller-owned device int32 used by the persistent kernel's dynamic ray
 * fetch; the call resets it on the stream before the launch.  Pass NULL to use
 * the library's per-device counter (then calls on different streams of one
 * device must not overlap). */
void cuda_intersect_single_ray1_bvh8_tri4_async(int32_t dev, const Node8* nodes, const Tri4* tris,
                                                const Ray1* rays, Hit1* hits, int32_t num_rays,
                                                void* stream, int32_t* work_counter);
void cuda_occluded_single_ray1_bvh8_tri4_async(int32_t dev, const Node8* nodes, const Tri4* tris,
                                               const Ray1* rays, Hit1* hits, int32_t num_rays,
                                               void* stream, int32_t* work_counter);

/* ---- Traversal, host pointers ------------------------------------------- *
 * Drop-ins for the call sites of
 *   cpu_{intersect,occluded}_single_ray1_bvh8_tri4(nodes, tris, rays, hits, num_packets)
 *   (tools/bench_traversal/bench_traversal.cpp:76-82): identical signature, HOST
 * buffers.  The BVH is uploaded to the current device on first use and cached by
 * (nodes, tris) address; its extent is found by walking it from the root (node 1),
 * because the reference signature carries no sizes.  Rays are copied in and hits
 * copied out on every call.  rodent_b200_forget_bvh drops a cached copy (call it
 * before freeing or rewriting a BVH that was passed here). */
void b200_intersect_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_packets);
void b200_occluded_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris,
                                         const Ray1* rays, Hit1* hits, int32_t num_packets);
void rodent_b200_forget_bvh(const Node8* nodes, const Tri4* tris);

/* ---- Memory helpers (stand-ins for anydsl::Array / anydsl_alloc / anydsl_copy,
 * tools/common/load_bvh.h:58-68, load_rays.h:85-88) ------------------------ */
int32_t rodent_b200_device_count(void);
void    rodent_b200_set_device(int32_t dev);          /* device used by the host-pointer entry points */
void*   rodent_b200_alloc_device(int32_t dev, size_t bytes);
void    rodent_b200_free_device(int32_t dev, void* ptr);
void*   rodent_b200_alloc_host(size_t bytes);         /* page-locked */
void    rodent_b200_free_host(void* ptr);
void    rodent_b200_copy_to_device(int32_t dev, void* dst, const void* src, size_t bytes);
void    rodent_b200_copy_to_host(int32_t dev, void* dst, const void* src, size_t bytes);
void    rodent_b200_sync(int32_t dev);

/* Kernel time of the most recent synchronous traversal call on `dev`, in ms,
 * measured with CUDA events around the launch (the role anydsl_get_kernel_time
 * plays at tools/bench_traversal/bench_traversal.cpp:125-133). */
double  rodent_b200_last_kernel_ms(int32_t dev);

/* Number of kernels this library launched so far in this process. */
int64_t rodent_b200_launch_count(void);

/* Library/version string, e.g. "rodent_b200 0.1 sm_100a". */
const char* rodent_b200_version(void);

#ifdef __cplusplus
}
#endif

#endif /* RODENT_B200_H */
