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

/* ---- Layouts ------------------------------------------------------------ *
 * Define RODENT_B200_NO_LAYOUTS before including this header when the reference's
 * generated tools/common/traversal.h (same structs, same names) is included too. */
#ifndef RODENT_B200_NO_LAYOUTS

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

/* The layout of the reference's own GPU path, src/traversal/mapping_gpu.impala:3-16.  Node2: bounds = lo_x, hi_x,
 * lo_y, hi_y, lo_z, hi_z of child 0, then of child 1; child as in Node8.  Tri1: one triangle, n is recomputed as
 * cross(e1, e2); prim_id < 0: last triangle of its leaf, the reported id is prim_id & 0x7FFFFFFF. */
typedef struct Node2 {
    float   bounds[12];
    int32_t child[2];
    int32_t pad[2];
} Node2;
typedef struct Tri1 {
    float   v0[3]; int32_t pad;
    float   e1[3]; int32_t geom_id;
    float   e2[3]; int32_t prim_id;
} Tri1;

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
#endif /* RODENT_B200_NO_LAYOUTS */

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

/* The reference's GPU exports themselves, on their own BVH2 / Tri1 layout: drop-ins for
 *   nvvm_{intersect,occluded}_single_ray1_bvh2_tri1(dev, nodes, tris, rays, hits, num_rays)
 *   (tools/bench_traversal/bench_traversal.impala:459-493), i.e. `bench_traversal -gpu nvvm` unchanged up to the call.
 * Semantics of gpu_traverse_single_helper (src/traversal/mapping_gpu.impala:94-178): unordered slab test with the
 * video min/max instructions on the float bits (:74-85), both children hit -> nearer entry first, no culling of
 * stacked entries, leaves of single triangles.  The hit writer stores all four fields in both modes
 * (make_gpu_hit1, bench_traversal.impala:78-83): miss -> tri_id = -1, t = tmax, u = v = 0; `occluded` returns the first
 * accepted triangle's record.  `nodes` must be 32-byte aligned (device allocations are). */
void cuda_intersect_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_rays);
void cuda_occluded_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris,
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
 * buffers.  The BVH is uploaded to the current device on first use; its extent is found by walking it from the root
 * (node 1), because the reference signature carries no sizes (a malformed BVH -- ids out of range, cycles, a leaf
 * without its end marker -- aborts with a message).  The cpu_* functions are pure, so the cached copy is keyed on
 * CONTENT: every call re-hashes samples of both arrays (first and last KB plus 64 lines spread over each) and uploads
 * again when they changed under a known address; at most 8 copies per device are kept (least recently used goes first).
 * A caller that patches a BVH in place in bytes the samples may miss calls rodent_b200_forget_bvh.
 * Rays come in and hits go out on every call.  With pinned (cudaMallocHost / rodent_b200_alloc_host) buffers a
 * closest-hit call (BVH8 or BVH4) is ONE launch: a copy engine brings the rays in while the kernel already traces the first of
 * them, and the kernel writes its records into the caller's array itself (traverse.cu: run_host_direct).  Pageable
 * rays (malloc, std::vector, anydsl::Array host memory) go through the library's own pinned staging buffers in copy /
 * launch / copy pieces, filled and drained by a few helper threads; rodent_b200_pin_host page-locks such arrays in
 * place.  Any-hit calls bring home the triangle ids only; the helper threads write them into the caller's records,
 * whose t, u, v are never touched.
 * Like the cpu_* functions they replace, the calls are reentrant: each takes its own device
 * staging buffers and streams, so calls from several host threads overlap on the device (one set's transfers under
 * another set's traversal). */
void b200_intersect_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_packets);
void b200_occluded_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris,
                                         const Ray1* rays, Hit1* hits, int32_t num_packets);
void rodent_b200_forget_bvh(const void* nodes, const Tri4* tris);
/* uploads, re-uploads after a content change, live copies -- of the cache above (current device) */
void rodent_b200_bvh_cache_stats(int64_t out[3]);

/* ---- The same four over a BVH4 -------------------------------------------- *
 * cpu_{intersect,occluded}_single_ray1_bvh4_tri4 (tools/bench_traversal/bench_traversal.impala:279-305) is what
 * `bench_traversal -s` runs at its default --bvh-width 4 (bench_traversal.cpp:147).  Node4 arrays, Tri4 leaves,
 * the arity-4 sorting networks (bose_nelson_sort, src/core/sort.impala:3-32); contracts as above. */
void cuda_intersect_single_ray1_bvh4_tri4(int32_t dev, const Node4* nodes, const Tri4* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_rays);
void cuda_occluded_single_ray1_bvh4_tri4(int32_t dev, const Node4* nodes, const Tri4* tris,
                                         const Ray1* rays, Hit1* hits, int32_t num_rays);
void b200_intersect_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris,
                                          const Ray1* rays, Hit1* hits, int32_t num_packets);
void b200_occluded_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris,
                                         const Ray1* rays, Hit1* hits, int32_t num_packets);

/* ---- Packet interfaces ----------------------------------------------------- *
 * tools/bench_traversal/bench_traversal.impala:32-65: structure-of-arrays packets of 4 or 8 rays / hits. */
typedef struct Ray4 { float org[3][4]; float dir[3][4]; float tmin[4]; float tmax[4]; } Ray4;
typedef struct Ray8 { float org[3][8]; float dir[3][8]; float tmin[8]; float tmax[8]; } Ray8;
typedef struct Hit4 { int32_t tri_id[4]; float t[4]; float u[4]; float v[4]; } Hit4;
typedef struct Hit8 { int32_t tri_id[8]; float t[8]; float u[8]; float v[8]; } Hit8;

/* Drop-ins for the call sites of the reference's packet and hybrid variants,
 *   cpu_{intersect,occluded}_{packet,hybrid}_ray{4,8}_bvh{4,8}_tri4(nodes, tris, rays, hits, num_packets)
 *   (bench_traversal.impala:159-427, called at bench_traversal.cpp:44-74,84-122): identical signatures, HOST buffers.
 * A GPU has no use for the CPU's packet traversal order, so every ray of a packet is traced by the single-ray kernel:
 * each ray gets the closest hit (or an any-hit flag) exactly as cpu_*_single_ray1_* defines it.  The reference's packet
 * kernels visit nodes in a per-packet order, which can pick another of several triangles hit at the same distance; `t`
 * agrees (the reference's own test accepts every variant against one golden image, tools/CMakeLists.txt:26-31).
 * Measured against a restatement of that packet kernel (oracle/traversal_oracle.c: traverse_packet,
 * tests/test_packet_oracle.py) on the two Sponza sets: the same rays hit, t within 1 ulp everywhere, another triangle of
 * a tie on 6 .. 694 of 1 Mi primary rays and on 0.03 .. 2.1 % of the incoherent ones.  A caller that needs the packet
 * kernels' records to the bit switches these entry points to the packet walk itself with
 * rodent_b200_set_packet_order(1) (one thread per packet; several times slower; process-wide, default 0). */
void rodent_b200_set_packet_order(int32_t on);
void b200_intersect_packet_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_occluded_packet_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_intersect_packet_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_occluded_packet_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_intersect_packet_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_occluded_packet_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_intersect_packet_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_occluded_packet_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_intersect_hybrid_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_occluded_hybrid_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_intersect_hybrid_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_occluded_hybrid_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_intersect_hybrid_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_occluded_hybrid_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets);
void b200_intersect_hybrid_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);
void b200_occluded_hybrid_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets);

/* ---- Memory helpers (stand-ins for anydsl::Array / anydsl_alloc / anydsl_copy,
 * tools/common/load_bvh.h:58-68, load_rays.h:85-88) ------------------------ */
int32_t rodent_b200_device_count(void);
void    rodent_b200_set_device(int32_t dev);          /* device used by the host-pointer entry points */
/* ... or several: every host-pointer call is then cut into contiguous ray ranges, one per device, BVH replicated, each
 * range copied back into its slice of the caller's array (no collective on this path). */
void    rodent_b200_set_devices(const int32_t* devs, int32_t num_devs);
/* One process per GPU: a device allocation (from rodent_b200_alloc_device) exported as a 64-byte CUDA IPC handle, and
 * such a handle opened in another process on its own device -- the pointer it returns is peer memory over NVLink /
 * NVSwitch and can be passed as the `hits` array of the cuda_* entry points, so that a rank's kernel writes its records
 * straight into the gathering rank's HBM (no collective afterwards; bench.py --gpus N).  export returns 0 and open NULL
 * (with a message on stderr) where IPC or peer access is not available. */
int32_t rodent_b200_ipc_export(int32_t dev, const void* device_ptr, void* handle_out_64_bytes);
void*   rodent_b200_ipc_open(int32_t dev, const void* handle_64_bytes);
void    rodent_b200_ipc_close(int32_t dev, void* ptr);
void*   rodent_b200_alloc_device(int32_t dev, size_t bytes);
void    rodent_b200_free_device(int32_t dev, void* ptr);
void*   rodent_b200_alloc_host(size_t bytes);         /* page-locked */
void    rodent_b200_free_host(void* ptr);
/* Page-lock an array the host allocated itself (malloc, std::vector, anydsl::Array host memory) in place, and release it
 * before the host frees it: the b200_* calls then treat it like rodent_b200_alloc_host memory (one launch per call, no
 * staging).  0 on success, -1 when the driver refuses (not page-lockable, already registered, unknown pointer). */
int32_t rodent_b200_pin_host(void* ptr, size_t bytes);
int32_t rodent_b200_unpin_host(void* ptr);
/* Checks the host side of the pageable-buffer path (helper-thread pool, staging copies) without touching a device:
 * 0 when everything checks out.  tests/test_abi.py runs it where there is no GPU. */
int32_t rodent_b200_selftest_host_copies(void);
void    rodent_b200_copy_to_device(int32_t dev, void* dst, const void* src, size_t bytes);
void    rodent_b200_copy_to_host(int32_t dev, void* dst, const void* src, size_t bytes);
void    rodent_b200_sync(int32_t dev);

/* Kernel time of the most recent synchronous traversal call on `dev`, in ms,
 * measured with CUDA events around the launch (the role anydsl_get_kernel_time
 * plays at tools/bench_traversal/bench_traversal.cpp:125-133). */
double  rodent_b200_last_kernel_ms(int32_t dev);
const char* rodent_b200_last_kernel_name(int32_t dev);   /* the BVH8 kernel variant the last launch on `dev` picked */

/* Number of kernels this library launched so far in this process. */
int64_t rodent_b200_launch_count(void);

/* Library/version string, e.g. "rodent_b200 0.1 sm_100a". */
const char* rodent_b200_version(void);


/* ======================================================================== *
 * Scene description and wavefront path tracer (the `rodent` driver surface)
 * ======================================================================== */

/* src/dummy_main.impala:3-13, src/driver/converter.cpp:617-626 */
typedef struct Vec3 { float x, y, z; } Vec3;
typedef struct Settings {
    Vec3  eye, dir, up, right;
    float width, height;       /* half extents of the image plane: w = tan(fov/2), h = w / ratio (driver.cpp:36-37) */
} Settings;

/* The reference compiles one shader per material into the renderer
 * (src/driver/converter.cpp:857-919).  Here the same rules produce a table that the
 * shade kernel interprets; `bsdf` selects the constructor of src/render/material.impala. */
enum RodentBsdf {
    RODENT_BSDF_BLACK   = 0,   /* make_black_bsdf,   material.impala:65-72   */
    RODENT_BSDF_DIFFUSE = 1,   /* make_diffuse_bsdf, material.impala:75-91   */
    RODENT_BSDF_PHONG   = 2,   /* make_phong_bsdf,   material.impala:94-116  */
    RODENT_BSDF_MIX     = 3,   /* make_mix_bsdf(diffuse, phong, k), material.impala:167-192, k from converter.cpp:897-902 */
    RODENT_BSDF_MIRROR  = 4,   /* make_mirror_bsdf,  material.impala:119-128 */
    RODENT_BSDF_GLASS   = 5    /* make_glass_bsdf(1.0, ni, ks, tf), material.impala:131-164 */
};
typedef struct RodentMaterial {
    int32_t bsdf;              /* enum RodentBsdf */
    int32_t is_emissive;       /* make_emissive_material with lights(light_ids[prim]) (converter.cpp:915) */
    float   ns, ni;
    float   kd[3], mix_k;
    float   ks[3]; int32_t map_kd;   /* 1 + index of the diffuse texture (MTL map_Kd), 0 = the constant kd  (converter.cpp:879-885) */
    float   tf[3]; int32_t map_ks;   /* 1 + index of the specular texture (MTL map_Ks), 0 = the constant ks (converter.cpp:887-893) */
    float   ke[3]; int32_t map_ke;   /* emitted radiance of an emissive material (MTL Ke); 1 + index of the emission texture
                                      * (MTL map_Ke, converter.cpp:794-801), 0 = the constant ke */
} RodentMaterial;

/* One image of the scene as load_png leaves it (src/driver/image.cpp:10-93): RGBA8 packed little-endian in a uint32_t
 * (r = bits 0-7: make_image_rgba32, src/render/image.impala:24-38), bottom row first, gamma 2.2 applied to r, g, b.
 * Materials sample it with the repeat border and the bilinear filter (image.impala:48-92). */
typedef struct RodentTexture {
    int32_t width, height;
    int64_t offset;            /* first pixel in RodentSceneView::texture_pixels */
} RodentTexture;

/* make_precomputed_triangle_light inputs (src/render/light.impala:147-154,
 * data/light_{verts,norms,areas,colors}.bin of converter.cpp:807-818) */
typedef struct RodentLight {
    float v0[3], inv_area;
    float v1[3]; int32_t prim;       /* the triangle this light is (its texture coordinates are looked up through it) */
    float v2[3]; int32_t map_ke;     /* the material's map_ke: with it, `color` is replaced by the texture at the sampled point */
    float n[3],  pad2;
    float color[3], pad3;
} RodentLight;

/* Host-side view of a loaded scene: the buffers the reference's converter writes to
 * its data directory (vertices.bin, normals.bin, ...: converter.cpp:403-409, 832) with the 16-byte padding it uses for GPU
 * targets, plus the BVH8/Tri4 arrays and the material / light tables. */
typedef struct RodentSceneView {
    int32_t num_tris, num_vertices, num_materials, num_lights, num_nodes, num_tri4;
    const float*          vertices;      /* num_vertices x float4 (xyz, 0)            */
    const float*          normals;       /* num_vertices x float4                     */
    const float*          face_normals;  /* num_tris x float4                         */
    const float*          texcoords;     /* num_vertices x float4 (uv, 0, 0)          */
    const int32_t*        indices;       /* num_tris x int4: i0, i1, i2, material id  */
    const int32_t*        light_ids;     /* num_tris: light of an emissive triangle   */
    const RodentMaterial* materials;
    const RodentLight*    lights;
    const Node8*          nodes;
    const Tri4*           tris;
    const RodentTexture*  textures;       /* num_textures */
    const uint32_t*       texture_pixels; /* num_texture_pixels, all images back to back */
    int64_t               num_texture_pixels;
    int32_t               num_textures, pad;
    const Node2*          nodes2;         /* the same triangles under a BVH2 / Tri1 (NULL until built or set), */
    const Tri1*           tris1;          /* the layout of the reference's GPU renderer (mapping_gpu.impala:19-69) */
    int32_t               num_nodes2, num_tri1;
} RodentSceneView;

typedef struct RodentScene RodentScene;

/* Runtime replacement of the reference's scene compiler (src/driver/converter.cpp:
 * convert_obj): OBJ/MTL parsing, material clean-up and de-duplication (:440-557),
 * triangle mesh (src/driver/obj.cpp:412-509), light extraction (:778-818), material
 * rules (:857-913) and a BVH8/Tri4 build (plus the BVH2/Tri1 of the reference's GPU device for scenes of 4096 triangles and
 * more, see rodent_b200_scene_build_bvh2).  Returns NULL and prints the reason on error. */
RodentScene* rodent_b200_scene_load_obj(const char* obj_file);
/* A scene whose geometry is an existing BVH8/Tri4 (e.g. the Sponza block of
 * testing/sponza.bvh, which carries no materials): triangles are recovered as
 * v1 = v0 - e1, v2 = v0 + e2 (src/traversal/intersection.impala:110-119), flat normals;
 * `material_of_prim[p]` indexes `materials`; emissive materials make their triangles lights. */
RodentScene* rodent_b200_scene_from_bvh8(const Node8* nodes, int32_t num_nodes, const Tri4* tris, int32_t num_tri4,
                                         const RodentMaterial* materials, int32_t num_materials,
                                         const int32_t* material_of_prim, int32_t num_prims);
/* Appends an image (RGBA8 as in RodentTexture, already gamma-corrected, bottom row first) to the scene and returns the value
 * to store in RodentMaterial::map_kd / map_ks (1 + its index).  What device.load_png hands the generated shaders
 * (src/render/mapping_gpu.impala:518-525), for scenes that are not built from an OBJ file.  Call before creating renderers. */
int32_t rodent_b200_scene_add_texture(RodentScene* scene, const uint32_t* rgba, int32_t width, int32_t height);
/* Decodes a PNG file the way load_png does (image.cpp:25-93: 8-bit RGBA, palette / grey expanded, 16 bit stripped, rows
 * flipped, gamma 2.2) and appends it; returns the map_kd / map_ks value or 0 on error (reason printed). */
int32_t rodent_b200_scene_add_png(RodentScene* scene, const char* png_file);
/* The same for a JPEG file, as load_jpg leaves it (image.cpp:186-238: libjpeg defaults, (r, g, b, 0) or (grey, 0, 0, 0),
 * rows flipped, gamma 2.2).  Sequential and progressive Huffman files; arithmetic-coded ones are reported as unsupported. */
int32_t rodent_b200_scene_add_jpg(RodentScene* scene, const char* jpg_file);
/* The same for a TGA file (types 1, 2, 3 and their run-length forms, 8 / 15 / 16 / 24 / 32 bit).  The reference's converter
 * emits device.load_tga for such images (converter.cpp:759-762) but no device provides it; the layout is that of the PNG
 * loader (bottom row first, gamma 2.2, alpha kept). */
int32_t rodent_b200_scene_add_tga(RodentScene* scene, const char* tga_file);
void rodent_b200_scene_view(const RodentScene* scene, RodentSceneView* out);
void rodent_b200_scene_free(RodentScene* scene);
/* The scene's triangles under a BVH4 as well (built on first use, owned by the scene), for writing .bvh files with both
 * blocks as tools/bvh_extractor/extract_bvh4_8.cpp:9-42 does. */
void rodent_b200_scene_bvh4(RodentScene* scene, const Node4** nodes, int32_t* num_nodes, const Tri4** tris, int32_t* num_tri4);
/* BVH2 / Tri1 over the scene's triangles -- what the reference's GPU device renders from (device.load_bvh2_tri1,
 * src/render/mapping_gpu.impala:505-509).  `build` makes one from the scene's own builder; `set` adopts a caller's arrays
 * (e.g. the BVH2 block of a .bvh file over the same triangles; copied, Tri1::geom_id rewritten to the scene's material
 * ids; returns 0 if a prim_id is out of range).  Renderers created afterwards trace their rays through it with the
 * reference GPU path's traversal (Sponza: 1.7x the samples/s of the BVH8 walk). */
void    rodent_b200_scene_build_bvh2(RodentScene* scene);
/* Replaces the scene's BVH8 / Tri4 by one from this library's split-BVH builder over the scene's own triangles (a scene
 * made with rodent_b200_scene_from_bvh8 carries the caller's tree until then). */
void    rodent_b200_scene_rebuild_bvh8(RodentScene* scene);
int32_t rodent_b200_scene_set_bvh2(RodentScene* scene, const Node2* nodes, int32_t num_nodes, const Tri1* tris, int32_t num_tri1);

/* The converter's data/ directory (convert_obj, src/driver/converter.cpp:403-438, 682-745): the mesh as LZ4-framed buffers
 * (vertices / normals / face_normals / texcoords .bin, 16-byte elements for the GPU targets, and indices.bin), the BVH of
 * the target's layout in bvh.bin, and bvh.stamp ("<target> <obj file>").  rodent_b200_scene_load_data renders from such a
 * directory: the arrays and whatever BVH it holds are adopted (a BVH8 as the scene's, a BVH2 / Tri1 as its second tree; what
 * is missing is built), --fusion's simple_kd / ks / ns buffers become table entries again, and since the reference keeps
 * materials and lights as generated code, those are read from the OBJ / MTL that `obj_file` -- or, if NULL, the stamp --
 * names.  rodent_b200_scene_write_data writes the directory from a loaded scene (`bvh_arity` 2, 4 or 8; `padded` as for the
 * GPU targets).  Both return NULL / 0 with a message on stderr when something does not fit. */
RodentScene* rodent_b200_scene_load_data(const char* data_dir, const char* obj_file);
int32_t rodent_b200_scene_write_data(RodentScene* scene, const char* data_dir, int32_t bvh_arity, int32_t padded, const char* obj_file);

typedef struct RodentRenderer RodentRenderer;

/* A wavefront path tracer bound to one scene and one device (the role of the generated
 * `render` + make_nvvm_device(dev, streaming) of src/render/mapping_gpu.impala:537-593).
 * `spp` and `max_path_len` are the reference's CMake cache variables SPP / MAX_PATH_LEN
 * (src/CMakeLists.txt:6-8).  Multi-GPU: the renderer owns the rows y with
 * (y / band) % num_parts == part; pass part 0 of 1 for the whole film. */
RodentRenderer* rodent_b200_renderer_create(const RodentScene* scene, int32_t dev, int32_t width, int32_t height,
                                            int32_t spp, int32_t max_path_len, int32_t part, int32_t num_parts, int32_t band);
/* The same renderer spread over several devices of this process: device devs[k] owns the row bands with
 * (y / band) % num_devs == k, the scene is replicated, and every render call ends with one ncclReduce that sums the
 * films onto devs[0] (NCCL is bound at run time: csrc/nccl_dyn.h).  The handle is used like a single-device one --
 * render, present, film, clear, stats, free.  The reference has no counterpart (one device per process). */
RodentRenderer* rodent_b200_renderer_create_multi(const RodentScene* scene, const int32_t* devs, int32_t num_devs,
                                                  int32_t width, int32_t height, int32_t spp, int32_t max_path_len, int32_t band);
void  rodent_b200_renderer_free(RodentRenderer* r);
/* render(settings, iter) of the generated main (converter.cpp:628-967): adds
 * (sum over spp samples) / spp of iteration `iter` to the device film and copies the
 * film to the host (device.present -> rodent_present, interface.cpp:494-496). */
void  rodent_b200_render(RodentRenderer* r, const Settings* settings, int32_t iter);
/* Same without the device->host film copy (for timing the device work alone). */
void  rodent_b200_render_device(RodentRenderer* r, const Settings* settings, int32_t iter);
void  rodent_b200_present(RodentRenderer* r);
float* rodent_b200_film(RodentRenderer* r);            /* host film, width*height*3 floats (get_pixels, interface.cpp:520-522) */
void*  rodent_b200_film_device(RodentRenderer* r);     /* device film, same layout                                          */
/* Accumulate into a caller-owned device buffer (width*height*3 floats on the renderer's device) instead of
 * the renderer's own film, e.g. a tensor that a collective reduces in place when the image rows are dealt out
 * across GPUs; NULL switches back.  The reference has one film per process (interface.cpp:348). */
void   rodent_b200_renderer_bind_film(RodentRenderer* r, float* device_film);
void  rodent_b200_clear(RodentRenderer* r);            /* clear_pixels, interface.cpp:524-526                               */
/* Work counters of the last render call: [0] camera samples, [1] closest-hit rays,
 * [2] shadow rays, [3] wavefronts, [4] kernels launched. */
void  rodent_b200_render_stats(const RodentRenderer* r, int64_t out[5]);
double rodent_b200_render_last_ms(const RodentRenderer* r);   /* CUDA-event time of the last render call */

/* The reference driver's own entry points (src/driver/driver.cpp:55-58,266-296 calls
 * them; interface.cpp:512-526 and the generated main define them).  They act on the
 * renderer made current by rodent_b200_bind(), which stands for what the reference
 * bakes in at build time (SCENE_FILE, SPP, MAX_PATH_LEN, TARGET_DEVICE). */
void   rodent_b200_bind(const RodentScene* scene, int32_t dev, int32_t spp, int32_t max_path_len);
/* ... spread over several devices (setup_interface then builds a rodent_b200_renderer_create_multi renderer). */
void   rodent_b200_bind_multi(const RodentScene* scene, const int32_t* devs, int32_t num_devs, int32_t spp, int32_t max_path_len);
void   setup_interface(size_t width, size_t height);
void   cleanup_interface(void);
float* get_pixels(void);
void   clear_pixels(void);
int32_t get_spp(void);
void   render(const Settings* settings, int32_t iter);



/* ======================================================================== *
 * The reference's data containers (src/driver/buffer.h, data/bvh.bin)
 * ======================================================================== */

/* The LZ4 block format, both directions (the reference calls LZ4_compress_default / LZ4_decompress_safe of liblz4,
 * src/driver/buffer.h:17-20,39-44).  decompress: bytes written, or -1 on malformed input or too small a destination;
 * compress: bytes written, or -1 unless dst_capacity >= rodent_b200_lz4_compress_bound(n). */
int64_t rodent_b200_lz4_decompress(const void* src, int64_t src_size, void* dst, int64_t dst_capacity);
int64_t rodent_b200_lz4_compress_bound(int64_t n);
int64_t rodent_b200_lz4_compress(const void* src, int64_t n, void* dst, int64_t dst_capacity);
/* One buffer file of the converter's data directory, [u32 raw size][u32 compressed size][LZ4 block] (read_buffer /
 * write_buffer, buffer.h:22-61; the role of rodent_load_buffer, interface.cpp:598-601, without the per-device cache).
 * load: a malloc'ed copy of the raw bytes (free it with rodent_b200_free_buffer) or NULL; write: 1 on success. */
void*   rodent_b200_load_buffer(const char* file, int64_t* size);
void    rodent_b200_free_buffer(void* p);
int32_t rodent_b200_write_buffer(const char* file, const void* data, int64_t bytes);
/* data/bvh.bin: repeated { u32 sizeof(Node), u32 sizeof(Tri), buffer(nodes), buffer(tris) } (write_bvh, converter.cpp:428-438;
 * load_bvh<Node, Tri>, interface.cpp:432-454).  load returns the first entry with these record sizes (256 / 224: BVH8 / Tri4,
 * 128 / 224: BVH4 / Tri4, 64 / 48: BVH2 / Tri1) as malloc'ed arrays, 0 if there is none; append adds an entry. */
int32_t rodent_b200_load_bvh_bin(const char* file, int32_t node_size, int32_t tri_size,
                                 void** nodes, int64_t* num_nodes, void** tris, int64_t* num_tris);
int32_t rodent_b200_append_bvh_bin(const char* file, int32_t node_size, int32_t tri_size,
                                   const void* nodes, int64_t num_nodes, const void* tris, int64_t num_tris);

/* ======================================================================== *
 * Shading-interface micro-benchmark (tools/bench_interface)
 * ======================================================================== */

/* tools/bench_interface/bench_interface.impala:8-42 (the structs of the generated tools/common/interface.h). */
typedef struct Vec2  { float x, y; } Vec2;
typedef struct Color { float r, g, b; } Color;
typedef struct Tex {
    const Color* pixels;        /* device pointer, width * height */
    Color    border_color;
    uint32_t border;            /* 0 clamp, 1 repeat, 2 constant colour (:1-3) */
    uint32_t sampler;           /* 0 nearest, 1 bilinear (:5-6)               */
    int32_t  width, height;
} Tex;
typedef struct ShadedMesh {
    const Vec3*     vertices;   /* device pointers */
    const uint32_t* indices;    /* 4 per triangle  */
    const Vec3*     normals;
    const Vec2*     texcoords;
    Tex tex_kd, tex_ks, tex_ns;
} ShadedMesh;
typedef struct TriHit { int32_t id; Vec2 uv; } TriHit;

/* bench_interface(mesh, tri_hits, in_dirs, out_dirs, colors, n), tools/bench_interface/bench_interface.impala:137-143,
 * called at tools/bench_interface/bench_interface.cpp:183: for every hit, the shader input (interpolated point, normals,
 * texture coordinates, three texture look-ups, local frame: :91-123) and the diffuse BSDF evaluated on it (:125-135).
 * `mesh` is a HOST struct holding DEVICE pointers; the four arrays are device pointers on device 0, as the reference's
 * gpu_iterate (:56-65) has it; returns after the device finished. */
void bench_interface(const ShadedMesh* mesh, const TriHit* tri_hits, const Vec3* in_dirs, const Vec3* out_dirs,
                     Color* colors, int32_t n);

/* ======================================================================== *
 * Shading micro-benchmark (tools/bench_shading)
 * ======================================================================== */

/* src/render/driver.impala:24-52: one array per field, all of the same capacity. */
typedef struct RayStream {
    int32_t* id;
    float *org_x, *org_y, *org_z, *dir_x, *dir_y, *dir_z, *tmin, *tmax;
} RayStream;
typedef struct PrimaryStream {
    RayStream rays;
    int32_t  *geom_id, *prim_id;
    float    *t, *u, *v;
    uint32_t *rnd;
    float    *mis, *contrib_r, *contrib_g, *contrib_b;
    int32_t  *depth;
    int32_t  size, pad;
} PrimaryStream;

/* Drop-in for the call site of
 *   cpu_bench_shading(primary_in, primary_out, vertices, normals, face_normals, texcoords, indices, pixels,
 *                     width, height, begins, ends, num_tris, num_iters)
 *   (tools/bench_shading/bench_shading.impala:22-104, called at tools/bench_shading/bench_shading.cpp:207-222):
 * identical signature, HOST pointers.  For every ray of the four geometry ranges [begins[g], ends[g]) the surface
 * element, a material made of a diffuse and a Phong lobe mixed by luminance (constant or bilinear, repeat-border
 * texture colours depending on the geometry id), one BSDF sample, and the bounced ray + path state written to
 * primary_out; repeated num_iters times.  Inputs are uploaded, outputs downloaded on every call. */
void b200_bench_shading(const PrimaryStream* primary_in, PrimaryStream* primary_out,
                        const Vec3* vertices, const Vec3* normals, const Vec3* face_normals, const Vec2* texcoords,
                        const int32_t* indices, const uint32_t* pixels, int32_t width, int32_t height,
                        const int32_t* begins, const int32_t* ends, int32_t num_tris, int32_t num_iters);

#ifdef __cplusplus
}
#endif

#endif /* RODENT_B200_H */
