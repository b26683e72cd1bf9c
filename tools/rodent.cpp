// rodent -- command-line compatible stand-in for the reference's renderer front end
// (src/driver/driver.cpp: options :169-232, frame loop :279-303, PNG with 2.2 gamma :138-162,
// min/med/max Msamples/s :341-348), without the SDL viewer (the DISABLE_GUI build, which is what the
// reference's benchmarks use: benchmarks/bench.sh:47).  It drives the reference's own entry points
// setup_interface / render / get_pixels / clear_pixels / get_spp / cleanup_interface, which
// librodent_b200.so exports with a B200 wavefront path tracer behind them.
//
// What the reference bakes in at build time is given at run time here:
//   --scene file.obj  (SCENE_FILE)   --spp n (SPP, default 4)   --max-path-len n (MAX_PATH_LEN, default 64)
//   --dev k           (TARGET_DEVICE)      --gpus n  (this repository's addition: the film spread over n devices)
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "formats.h"

namespace {

[[noreturn]] void error(const std::string& msg) { std::cerr << msg << std::endl; std::exit(1); }

struct F3 { float x, y, z; };
F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
F3 cross(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
F3 normalize(F3 a) { return a * (1.0f / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }

void check_arg(int argc, char** argv, int arg, int n) {
    if (arg + n >= argc) error(std::string("Option '") + argv[arg] + "' expects " + std::to_string(n) + " arguments, got " + std::to_string(argc - arg));
}

void usage() {
    std::cout << "Usage: rodent [options]\n"
              << "Available options:\n"
              << "   --help              Shows this message\n"
              << "   --scene  file.obj   Scene to render (the reference's SCENE_FILE)\n"
              << "   --data   dir        Render from a converter-written data/ directory (materials from --scene, else from dir/bvh.stamp)\n"
              << "   --spp    n          Samples per pixel and iteration (the reference's SPP, default 4)\n"
              << "   --max-path-len n    Maximum path length (the reference's MAX_PATH_LEN, default 64)\n"
              << "   --dev    k          CUDA device index\n"
              << "   --gpus   n          Spread the film over devices k .. k+n-1 (row bands of 8, one ncclReduce per frame)\n"
              << "   --width  pixels     Sets the viewport horizontal dimension (in pixels)\n"
              << "   --height pixels     Sets the viewport vertical dimension (in pixels)\n"
              << "   --eye    x y z      Sets the position of the camera\n"
              << "   --dir    x y z      Sets the direction vector of the camera\n"
              << "   --up     x y z      Sets the up vector of the camera\n"
              << "   --fov    degrees    Sets the horizontal field of view (in degrees)\n"
              << "   --bench  iterations Enables benchmarking mode and sets the number of iterations\n"
              << "   -o       image.png  Writes the output image to a file" << std::endl;
}

// save_image, driver.cpp:138-162
void save_image(const std::string& out_file, size_t width, size_t height, uint32_t iter) {
    std::vector<uint8_t> rgba(width * height * 4);
    const float* film = get_pixels();
    const float inv_iter = 1.0f / iter, inv_gamma = 1.0f / 2.2f;
    auto tone = [&](float v) { return uint8_t(std::min(std::max(std::pow(v * inv_iter, inv_gamma), 0.0f), 1.0f) * 255.0f); };
    for (size_t i = 0; i < width * height; i++) {
        rgba[4 * i + 0] = tone(film[3 * i + 0]);
        rgba[4 * i + 1] = tone(film[3 * i + 1]);
        rgba[4 * i + 2] = tone(film[3 * i + 2]);
        rgba[4 * i + 3] = 255;
    }
    if (!rb200::write_png_rgba(out_file, rgba, uint32_t(width), uint32_t(height))) error("Failed to save PNG file '" + out_file + "'");
}

}  // namespace

int main(int argc, char** argv) {
    std::string out_file, scene_file, data_dir;
    size_t bench_iter = 0, width = 1080, height = 720;
    int spp = 4, max_path_len = 64, dev = 0, gpus = 1;
    float fov = 60.0f;
    F3 eye{0, 0, 0}, dir{0, 0, 1}, up{0, 1, 0};
    for (int i = 1; i < argc; ++i) {
        if (argv[i][0] != '-') error(std::string("Unexpected argument '") + argv[i] + "'");
        auto f3 = [&](F3& v) { check_arg(argc, argv, i, 3); v.x = strtof(argv[++i], nullptr); v.y = strtof(argv[++i], nullptr); v.z = strtof(argv[++i], nullptr); };
        if (!strcmp(argv[i], "--width")) { check_arg(argc, argv, i, 1); width = strtoul(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "--height")) { check_arg(argc, argv, i, 1); height = strtoul(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "--eye")) f3(eye);
        else if (!strcmp(argv[i], "--dir")) f3(dir);
        else if (!strcmp(argv[i], "--up")) f3(up);
        else if (!strcmp(argv[i], "--fov")) { check_arg(argc, argv, i, 1); fov = strtof(argv[++i], nullptr); }
        else if (!strcmp(argv[i], "--bench")) { check_arg(argc, argv, i, 1); bench_iter = strtoul(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "-o")) { check_arg(argc, argv, i, 1); out_file = argv[++i]; }
        else if (!strcmp(argv[i], "--scene")) { check_arg(argc, argv, i, 1); scene_file = argv[++i]; }
        else if (!strcmp(argv[i], "--data")) { check_arg(argc, argv, i, 1); data_dir = argv[++i]; }
        else if (!strcmp(argv[i], "--spp")) { check_arg(argc, argv, i, 1); spp = int(strtol(argv[++i], nullptr, 10)); }
        else if (!strcmp(argv[i], "--max-path-len")) { check_arg(argc, argv, i, 1); max_path_len = int(strtol(argv[++i], nullptr, 10)); }
        else if (!strcmp(argv[i], "--dev")) { check_arg(argc, argv, i, 1); dev = int(strtol(argv[++i], nullptr, 10)); }
        else if (!strcmp(argv[i], "--gpus")) { check_arg(argc, argv, i, 1); gpus = int(strtol(argv[++i], nullptr, 10)); }
        else if (!strcmp(argv[i], "--help")) { usage(); return 0; }
        else error(std::string("Unknown option '") + argv[i] + "'");
    }
    if (scene_file.empty() && data_dir.empty()) error("No scene: pass --scene file.obj (or --data dir)");
    if (bench_iter == 0) error("No display in this build (the reference's DISABLE_GUI): pass --bench iterations");
    if (width == 0 || height == 0 || spp <= 0) error("Invalid image size or sample count");

    // Camera::Camera, driver.cpp:31-38
    const float pi = 3.14159265359f;
    const F3 d = normalize(dir), right = normalize(cross(d, up)), u = normalize(cross(right, d));
    const float w = std::tan(fov * pi / 360.0f), h = w / (float(width) / float(height));

    RodentScene* scene = data_dir.empty() ? rodent_b200_scene_load_obj(scene_file.c_str())
                                          : rodent_b200_scene_load_data(data_dir.c_str(), scene_file.empty() ? nullptr : scene_file.c_str());
    if (!scene) return 1;
    if (gpus < 1 || dev < 0 || rodent_b200_device_count() < dev + gpus) error("No such CUDA device");
    if (gpus > 1) {
        std::vector<int32_t> devs(gpus);
        for (int k = 0; k < gpus; k++) devs[k] = dev + k;
        rodent_b200_bind_multi(scene, devs.data(), gpus, spp, max_path_len);
    } else {
        rodent_b200_bind(scene, dev, spp, max_path_len);
    }
    setup_interface(width, height);

    std::vector<double> samples_sec;
    uint32_t iter = 0;
    for (;;) {
        if (iter == 0) clear_pixels();
        Settings settings{Vec3{eye.x, eye.y, eye.z}, Vec3{d.x, d.y, d.z}, Vec3{u.x, u.y, u.z}, Vec3{right.x, right.y, right.z}, w, h};
        const auto ticks = std::chrono::high_resolution_clock::now();
        render(&settings, int32_t(iter++));
        // the reference rounds to whole milliseconds (driver.cpp:297); a B200 frame can take less than one
        const double elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - ticks).count();
        samples_sec.emplace_back(1000.0 * double(get_spp()) * double(width) * double(height) / elapsed_ms);
        if (samples_sec.size() == bench_iter) break;
    }
    if (!out_file.empty()) {
        save_image(out_file, width, height, iter);
        std::cout << "Image saved to '" << out_file << "'" << std::endl;
    }
    cleanup_interface();
    rodent_b200_scene_free(scene);
    std::sort(samples_sec.begin(), samples_sec.end());
    std::cout << "# " << samples_sec.front() * 1e-6 << "/" << samples_sec[samples_sec.size() / 2] * 1e-6 << "/" << samples_sec.back() * 1e-6
              << " (min/med/max Msamples/s)" << std::endl;
    return 0;
}
