// librodent_b200_refnames.so -- the B200 entry points under the names the reference's own objects export
// (`extern fn` of tools/bench_traversal/bench_traversal.impala:159-493 and tools/bench_shading/bench_shading.impala:22).
//
// librodent_b200.so keeps names of its own (b200_*, cuda_*) so that it can be linked NEXT to the reference's generated
// traversal object.  A build that wants the reference's programs to run on the B200 without touching their sources
// links this shim INSTEAD of that object (tools/bench_traversal/CMakeLists.txt:24: drop ${TRAVERSAL_OBJ}, add the two
// libraries): every cpu_* call site of bench_traversal.cpp:44-122 and the nvvm_* call sites of :124-135 then land here.
// Same signatures, same array layouts, host pointers for cpu_*, device pointers for nvvm_* -- as in the reference.
#include "../../include/rodent_b200.h"

extern "C" {

void cpu_intersect_hybrid_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_intersect_hybrid_ray4_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_hybrid_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_occluded_hybrid_ray4_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_packet_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_intersect_packet_ray4_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_packet_ray4_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_occluded_packet_ray4_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_hybrid_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_intersect_hybrid_ray8_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_hybrid_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_occluded_hybrid_ray8_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_packet_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_intersect_packet_ray8_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_packet_ray8_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_occluded_packet_ray8_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    b200_intersect_single_ray1_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_single_ray1_bvh4_tri4(const Node4* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    b200_occluded_single_ray1_bvh4_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_hybrid_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_intersect_hybrid_ray4_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_hybrid_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_occluded_hybrid_ray4_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_packet_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_intersect_packet_ray4_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_packet_ray4_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray4* rays, Hit4* hits, int32_t num_packets) {
    b200_occluded_packet_ray4_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_hybrid_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_intersect_hybrid_ray8_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_hybrid_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_occluded_hybrid_ray8_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_packet_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_intersect_packet_ray8_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_packet_ray8_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray8* rays, Hit8* hits, int32_t num_packets) {
    b200_occluded_packet_ray8_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_intersect_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    b200_intersect_single_ray1_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}
void cpu_occluded_single_ray1_bvh8_tri4(const Node8* nodes, const Tri4* tris, const Ray1* rays, Hit1* hits, int32_t num_packets) {
    b200_occluded_single_ray1_bvh8_tri4(nodes, tris, rays, hits, num_packets);
}

// bench_traversal.impala:459-493: the reference GPU path's own layout and semantics (BVH2 / Tri1), device pointers
void nvvm_intersect_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    cuda_intersect_single_ray1_bvh2_tri1(dev, nodes, tris, rays, hits, num_rays);
}
void nvvm_occluded_single_ray1_bvh2_tri1(int32_t dev, const Node2* nodes, const Tri1* tris, const Ray1* rays, Hit1* hits, int32_t num_rays) {
    cuda_occluded_single_ray1_bvh2_tri1(dev, nodes, tris, rays, hits, num_rays);
}

// bench_shading.impala:22-35, called at bench_shading.cpp:207-222
void cpu_bench_shading(const PrimaryStream* primary_in, PrimaryStream* primary_out,
                       const Vec3* vertices, const Vec3* normals, const Vec3* face_normals, const Vec2* texcoords,
                       const int32_t* indices, const uint32_t* pixels, int32_t width, int32_t height,
                       const int32_t* begins, const int32_t* ends, int32_t num_tris, int32_t num_iters) {
    b200_bench_shading(primary_in, primary_out, vertices, normals, face_normals, texcoords, indices, pixels, width, height,
                       begins, ends, num_tris, num_iters);
}

}  // extern "C"
