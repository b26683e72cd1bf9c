// bvh_extractor -- writes the multi-block .bvh file that bench_traversal reads (tools/common/load_bvh.h:8-42) from an
// OBJ scene: a BVH8_TRI4 block and a BVH4_TRI4 block, as the reference's tools/bvh_extractor/extract_bvh4_8.cpp:9-42
// does.  The reference builds them with Embree (absent here); these come from this repo's binned-SAH builder collapsed
// to arity 8 / 4 (rodent_b200/csrc/scene.cpp), the one the renderer uses.
//   bvh_extractor scene.obj output.bvh
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>

#include "formats.h"

namespace {
template <typename NodeT>
bool write_block(rb200::File& out, uint32_t type, const NodeT* nodes, uint32_t num_nodes, const Tri4* tris, uint32_t num_tris) {
    const uint64_t size = sizeof(uint32_t) * 3 + sizeof(NodeT) * uint64_t(num_nodes) + sizeof(Tri4) * uint64_t(num_tris);
    return out.write(&size, 8) && out.write(&type, 4) && out.write(&num_nodes, 4) && out.write(&num_tris, 4) &&
           out.write(nodes, sizeof(NodeT) * size_t(num_nodes)) && out.write(tris, sizeof(Tri4) * size_t(num_tris));
}
}  // namespace

int main(int argc, char** argv) {
    if (argc != 3 || !std::strcmp(argv[1], "--help")) {
        std::cout << "Usage: bvh_extractor scene.obj output.bvh" << std::endl;
        return argc == 2 ? 0 : 1;
    }
    RodentScene* scene = rodent_b200_scene_load_obj(argv[1]);
    if (!scene) return 1;
    RodentSceneView view;
    rodent_b200_scene_view(scene, &view);
    const Node4* nodes4; const Tri4* tris4; int32_t n4, t4;
    rodent_b200_scene_bvh4(scene, &nodes4, &n4, &tris4, &t4);
    rb200::File out(argv[2], "wb");
    const uint32_t magic = rb200::kBvhMagic;
    if (!out || !out.write(&magic, 4) ||
        !write_block(out, rb200::kBvh8Tri4, view.nodes, uint32_t(view.num_nodes), view.tris, uint32_t(view.num_tri4)) ||
        !write_block(out, rb200::kBvh4Tri4, nodes4, uint32_t(n4), tris4, uint32_t(t4))) {
        std::cerr << "Cannot write " << argv[2] << std::endl;
        return 1;
    }
    std::cout << "BVH8: " << view.num_nodes << " nodes, " << view.num_tri4 << " Tri4; BVH4: " << n4 << " nodes, " << t4 << " Tri4 ("
              << view.num_tris << " triangles)" << std::endl;
    rodent_b200_scene_free(scene);
    return 0;
}
