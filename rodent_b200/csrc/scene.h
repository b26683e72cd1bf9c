// Host-side scene container behind the opaque RodentScene handle.
#pragma once

#include <array>
#include <string>
#include <vector>

#include "../../include/rodent_b200.h"

namespace rb200 {

struct Scene {
    std::vector<float> vertices, normals, face_normals, texcoords;   // float4 stride
    std::vector<int32_t> indices;                                     // int4 per triangle
    std::vector<int32_t> light_ids;
    std::vector<RodentMaterial> materials;
    std::vector<std::string> material_names;
    std::vector<RodentLight> lights;
    std::vector<Node8> nodes;
    std::vector<Tri4> tris;
    std::vector<Node4> nodes4;            // built on demand (rodent_b200_scene_bvh4)
    std::vector<Tri4> tris4;
    std::vector<Node2> nodes2;            // built or set on demand (rodent_b200_scene_{build,set}_bvh2)
    std::vector<Tri1> tris1;
    std::vector<RodentTexture> textures;  // images of map_Kd / map_Ks, pixels back to back
    std::vector<uint32_t> texture_pixels;

    int add_texture(const uint32_t* rgba, int width, int height);   // returns 1 + index
};

// src/driver/image.cpp:25-93 (image.cpp of this directory)
bool load_png(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why);
// src/driver/image.cpp:186-238 (jpeg.cpp of this directory)
bool load_jpg(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why);
// tga.cpp of this directory (the reference names a load_tga, converter.cpp:759-762, that no device provides)
bool load_tga(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why);

Scene* load_obj_scene(const std::string& path);
Scene* load_data_dir(const std::string& dir, const std::string& obj_path);
bool write_data_dir(Scene& scene, const std::string& dir, int arity, bool padded, const std::string& obj_path);
void build_bvh4(Scene& scene);
void build_bvh2(Scene& scene);
void rebuild_bvh8(Scene& scene);
bool set_bvh2(Scene& scene, const Node2* nodes, int num_nodes, const Tri1* tris, int num_tri1);
Scene* scene_from_bvh8(const Node8* nodes, int num_nodes, const Tri4* tris, int num_tri4,
                       const RodentMaterial* materials, int num_materials, const int32_t* material_of_prim, int num_prims);

}  // namespace rb200
