// converter -- command-line compatible stand-in for the reference's scene converter (src/driver/converter.cpp:973-1092).
//
// The reference turns an OBJ file into (a) a generated main.impala that hard-codes device, camera, materials, lights and
// shaders, to be compiled by the AnyDSL toolchain, and (b) a data/ directory with the mesh, the BVH and the images
// (:403-438, 682-768).  This repository evaluates the scene description at run time, so there is no (a); this tool writes
// (b) -- byte-compatible buffers and BVH container, the layout and padding of the chosen target -- plus data/render.cfg
// holding the options the reference would have baked into the code (spp, max path length, device).  `rodent --data data`
// renders from it.  Options as the reference's: -t / --target, -d / --device, --max-path-len, -spp, --fusion (accepted
// for the megakernel targets; materials stay a table here, so it changes nothing that is written).
#include <sys/stat.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "../include/rodent_b200.h"

namespace {
void usage() {
    std::cout << "converter [options] file\n"
              << "Available options:\n"
              << "    -h     --help                Shows this message\n"
              << "    -t     --target              Sets the target platform whose data layout is written (default: avx2)\n"
              << "    -d     --device              Sets the device to use on the selected platform (default: 0)\n"
              << "           --max-path-len        Sets the maximum path length (default: 64)\n"
              << "    -spp   --samples-per-pixel   Sets the number of samples per pixel (default: 4)\n"
              << "           --fusion              Accepted for the megakernel targets (no effect on the data written)\n"
              << "    -o     --output              Directory to write (default: data)\n"
              << "Available targets:\n"
              << "    generic, sse42, avx, avx2, asimd,\n"
              << "    nvvm = nvvm-streaming, nvvm-megakernel,\n"
              << "    amdgpu = amdgpu-streaming, amdgpu-megakernel\n"
              << std::flush;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc <= 1) { std::cerr << "Not enough arguments. Run with --help to get a list of options." << std::endl; return 1; }
    std::string obj_file, out_dir = "data", target = "avx2";
    long dev = 0, spp = 4, max_path_len = 64;
    bool fusion = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&]() -> const char* {
            if (i + 1 >= argc) { std::cerr << "Missing argument for '" << a << "'. Aborting." << std::endl; std::exit(1); }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "-t" || a == "--target") target = value();
        else if (a == "-d" || a == "--device") dev = std::strtol(value(), nullptr, 10);
        else if (a == "-spp" || a == "--samples-per-pixel") spp = std::strtol(value(), nullptr, 10);
        else if (a == "--max-path-len") max_path_len = std::strtol(value(), nullptr, 10);
        else if (a == "--fusion") fusion = true;
        else if (a == "-o" || a == "--output") out_dir = value();
        else if (a[0] == '-') { std::cerr << "Unknown option '" << a << "'. Aborting." << std::endl; return 1; }
        else if (!obj_file.empty()) { std::cerr << "Only one OBJ file can be converted. Aborting." << std::endl; return 1; }
        else obj_file = a;
    }
    // layouts per target: converter.cpp:630-633 (padding), 716-739 (BVH arity)
    int arity; bool padded, megakernel = false;
    if (target == "generic" || target == "sse42" || target == "asimd") { arity = 4; padded = false; }
    else if (target == "avx" || target == "avx2" || target == "avx2-embree") { arity = 8; padded = false; }
    else if (target == "nvvm" || target == "nvvm-streaming" || target == "amdgpu" || target == "amdgpu-streaming") { arity = 2; padded = true; }
    else if (target == "nvvm-megakernel" || target == "amdgpu-megakernel") { arity = 2; padded = true; megakernel = true; }
    else { std::cerr << "Unknown target '" << target << "'. Aborting." << std::endl; return 1; }
    if (fusion && !megakernel) { std::cerr << "Fusion is only available for megakernel targets. Aborting." << std::endl; return 1; }
    if (obj_file.empty()) { std::cerr << "Please specify an OBJ file to convert. Aborting." << std::endl; return 1; }

    RodentScene* scene = rodent_b200_scene_load_obj(obj_file.c_str());
    if (!scene) return 1;
    mkdir(out_dir.c_str(), 0777);
    if (!rodent_b200_scene_write_data(scene, out_dir.c_str(), arity, padded ? 1 : 0, obj_file.c_str())) return 1;
    std::ofstream cfg(out_dir + "/render.cfg");
    cfg << "spp " << spp << "\nmax_path_len " << max_path_len << "\ndevice " << dev << "\ntarget " << target << "\n";
    RodentSceneView v;
    rodent_b200_scene_view(scene, &v);
    std::cout << "Converted '" << obj_file << "' into '" << out_dir << "': " << v.num_tris << " triangle(s), " << v.num_materials
              << " material(s), " << v.num_lights << " light(s), BVH" << arity << std::endl;
    rodent_b200_scene_free(scene);
    return 0;
}
