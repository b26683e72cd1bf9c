// bench_interface -- the reference's shading-interface micro-benchmark (tools/bench_interface/bench_interface.cpp) on
// librodent_b200.so: the same quad, textures, 1 Mi random hits (mt19937, seed 42) and output line.  No arguments, as there;
// `--iters n` shortens the 1000 timed calls, `--check` prints the colour of the first hits.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <vector>

#include "../include/rodent_b200.h"

namespace {
template <typename T>
T* upload(const std::vector<T>& host) {
    T* p = static_cast<T*>(rodent_b200_alloc_device(0, host.size() * sizeof(T)));
    rodent_b200_copy_to_device(0, p, host.data(), host.size() * sizeof(T));
    return p;
}
}  // namespace

int main(int argc, char** argv) {
    size_t iters = 1000;                                                    // BENCH_CUDA, bench_interface.cpp:174-178
    bool check = false;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--iters") && i + 1 < argc) iters = std::max(1l, std::strtol(argv[++i], nullptr, 10));
        else if (!std::strcmp(argv[i], "--check")) check = true;
        else { std::cerr << "Invalid argument '" << argv[i] << "'" << std::endl; return 1; }
    }
    if (rodent_b200_device_count() < 1) { std::cerr << "No CUDA device" << std::endl; return 1; }

    // the quad, :57-94
    const std::vector<Vec3> vertices{{-1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}};
    const std::vector<Vec3> normals(4, Vec3{0, 0, 1});
    const std::vector<Vec2> texcoords{{-1, 1}, {-1, -1}, {1, -1}, {1, 1}};
    const std::vector<uint32_t> indices{0, 1, 2, uint32_t(-1), 2, 3, 0, uint32_t(-1)};
    // three constant images, :96-130
    const int width = 1024, height = 1024;
    auto image = [&](Color c) { return upload(std::vector<Color>(size_t(width) * height, c)); };
    const Tex tex_kd{image({0.1f, 0.2f, 0.3f}), {0.0f, 0.0f, 0.0f}, 0, 1, width, height};
    const Tex tex_ks{image({1.0f, 0.5f, 0.1f}), {0.5f, 1.0f, 0.2f}, 2, 0, width, height};
    const Tex tex_ns{image({0.1f, 0.5f, 1.0f}), {0.0f, 0.0f, 0.0f}, 1, 1, width, height};
    const ShadedMesh mesh{upload(vertices), upload(indices), upload(normals), upload(texcoords), tex_kd, tex_ks, tex_ns};

    // hits and directions, :143-165
    const size_t N = 1024 * 1024;
    std::vector<Vec3> in_dirs(N), out_dirs(N);
    std::vector<TriHit> tri_hits(N);
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> rnd(0.0f, 1.0f);
    auto unit = [&] {
        const float x = rnd(gen), y = rnd(gen), z = rnd(gen);
        const float inv = 1.0f / std::sqrt(x * x + y * y + z * z);
        return Vec3{x * inv, y * inv, z * inv};
    };
    for (size_t i = 0; i < N; i++) {
        tri_hits[i].id = int32_t(i % 2);
        tri_hits[i].uv.x = rnd(gen);
        tri_hits[i].uv.y = rnd(gen);
        in_dirs[i] = unit();
        out_dirs[i] = unit();
    }
    Vec3* d_in = upload(in_dirs);
    Vec3* d_out = upload(out_dirs);
    TriHit* d_hits = upload(tri_hits);
    Color* d_colors = static_cast<Color*>(rodent_b200_alloc_device(0, N * sizeof(Color)));

    std::vector<double> times;
    for (size_t i = 0; i < iters; i++) {                                    // :179-187
        const auto t0 = std::chrono::high_resolution_clock::now();
        bench_interface(&mesh, d_hits, d_in, d_out, d_colors, int32_t(N));
        const auto t1 = std::chrono::high_resolution_clock::now();
        times.push_back(double(std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count()));
    }
    std::sort(times.begin(), times.end());
    std::cout << double(N) / times[iters / 2] << " Mrays/s" << std::endl;   // :188
    if (check) {
        std::vector<Color> colors(4);
        rodent_b200_copy_to_host(0, colors.data(), d_colors, colors.size() * sizeof(Color));
        for (const Color& c : colors) std::cout << c.r << " " << c.g << " " << c.b << std::endl;
    }
    return 0;
}
