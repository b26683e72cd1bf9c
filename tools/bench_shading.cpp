// bench_shading -- the reference's shading micro-benchmark (tools/bench_shading/bench_shading.cpp) on librodent_b200.so:
// the same quad, checkerboard image, 4096 rays in four geometry ranges (mt19937, seed 42) and output line, with
// b200_bench_shading in place of cpu_bench_shading.  No arguments, as there; `--bench n` / `--iters n` shorten the run.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <random>
#include <vector>

#include "../include/rodent_b200.h"

namespace {
// get_primary_stream, bench_shading.cpp:27-54
void carve(PrimaryStream& s, float* ptr, size_t capacity) {
    s.rays.id = reinterpret_cast<int32_t*>(ptr);
    s.rays.org_x = ptr + 1 * capacity; s.rays.org_y = ptr + 2 * capacity; s.rays.org_z = ptr + 3 * capacity;
    s.rays.dir_x = ptr + 4 * capacity; s.rays.dir_y = ptr + 5 * capacity; s.rays.dir_z = ptr + 6 * capacity;
    s.rays.tmin = ptr + 7 * capacity; s.rays.tmax = ptr + 8 * capacity;
    s.geom_id = reinterpret_cast<int32_t*>(ptr) + 9 * capacity; s.prim_id = reinterpret_cast<int32_t*>(ptr) + 10 * capacity;
    s.t = ptr + 11 * capacity; s.u = ptr + 12 * capacity; s.v = ptr + 13 * capacity;
    s.rnd = reinterpret_cast<uint32_t*>(ptr) + 14 * capacity;
    s.mis = ptr + 15 * capacity; s.contrib_r = ptr + 16 * capacity; s.contrib_g = ptr + 17 * capacity; s.contrib_b = ptr + 18 * capacity;
    s.depth = reinterpret_cast<int32_t*>(ptr) + 19 * capacity;
    s.size = 0; s.pad = 0;
}
}  // namespace

int main(int argc, char** argv) {
    size_t num_iters = 1000, num_bench = 100;                               // :202-203
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--iters") && i + 1 < argc) num_iters = std::max(1l, std::strtol(argv[++i], nullptr, 10));
        else if (!std::strcmp(argv[i], "--bench") && i + 1 < argc) num_bench = std::max(1l, std::strtol(argv[++i], nullptr, 10));
        else { std::cerr << "Invalid argument '" << argv[i] << "'" << std::endl; return 1; }
    }
    if (rodent_b200_device_count() < 1) { std::cerr << "No CUDA device" << std::endl; return 1; }

    const std::vector<Vec3> vertices{{-1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}};      // :74-101
    const std::vector<Vec3> normals(4, Vec3{0, 0, 1}), face_normals(2, Vec3{0, 0, 1});
    const std::vector<Vec2> texcoords{{-1, 1}, {-1, -1}, {1, -1}, {1, 1}};
    const std::vector<int32_t> indices{0, 1, 2, -1, 2, 3, 0, -1};
    const size_t width = 1024, height = 1024;                                               // :103-111
    std::vector<uint32_t> pixels(width * height);
    for (size_t y = 0; y < height; y++)
        for (size_t x = 0; x < width; x++) pixels[y * width + x] = (x + y) % 2 != 0 ? uint32_t(-1) : 0;

    const size_t num_rays = 4096, num_geometries = 4, rays_per_geom = num_rays / num_geometries;   // :113-131
    std::vector<float> in_data(20 * num_rays), out_data(20 * num_rays);
    PrimaryStream in, out;
    carve(in, in_data.data(), num_rays);
    carve(out, out_data.data(), num_rays);
    const Vec3 org{0.0f, 0.0f, -1.0f};
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> rnd(0.0f, 1.0f);
    std::vector<int32_t> begins(num_geometries), ends(num_geometries);
    for (size_t geom = 0, cur = 0; geom < num_geometries; geom++, cur += rays_per_geom) {
        begins[geom] = int32_t(cur);
        ends[geom] = int32_t(cur + rays_per_geom);
        for (size_t i = 0; i < rays_per_geom; i++) {                                        // :132-170
            const int32_t prim_id = rnd(gen) < 0.5f ? 0 : 1;
            float u = rnd(gen), v = rnd(gen);
            if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
            const Vec3 &a = vertices[indices[prim_id * 4]], &b = vertices[indices[prim_id * 4 + 1]], &c = vertices[indices[prim_id * 4 + 2]];
            const float w = 1.0f - u - v;
            const Vec3 p{w * a.x + u * b.x + v * c.x, w * a.y + u * b.y + v * c.y, w * a.z + u * b.z + v * c.z};
            const size_t k = cur + i;
            in.rays.id[k] = int32_t(k);
            in.rays.org_x[k] = org.x; in.rays.org_y[k] = org.y; in.rays.org_z[k] = org.z;
            in.rays.dir_x[k] = p.x - org.x; in.rays.dir_y[k] = p.y - org.y; in.rays.dir_z[k] = p.z - org.z;
            in.rays.tmin[k] = 0.0f; in.rays.tmax[k] = std::numeric_limits<float>::max();
            in.geom_id[k] = int32_t(geom); in.prim_id[k] = prim_id;
            in.t[k] = 1.0f; in.u[k] = u; in.v[k] = v;
            in.rnd[k] = uint32_t(33 * geom + i);
            in.mis[k] = 0.5f; in.contrib_r[k] = in.contrib_g[k] = in.contrib_b[k] = 1.0f;
            in.depth[k] = 0;
        }
    }
    in.size = int32_t(num_rays);

    std::vector<double> us;
    for (size_t i = 0; i < num_bench; i++) {                                                // :205-225
        const auto t0 = std::chrono::high_resolution_clock::now();
        b200_bench_shading(&in, &out, vertices.data(), normals.data(), face_normals.data(), texcoords.data(), indices.data(),
                           pixels.data(), int32_t(width), int32_t(height), begins.data(), ends.data(), 2, int32_t(num_iters));
        const auto t1 = std::chrono::high_resolution_clock::now();
        us.push_back(double(std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count()));
    }
    std::sort(us.begin(), us.end());
    std::cout << double(num_rays * num_iters) / us[us.size() / 2] << " Mrays/s" << std::endl;   // :227
    return 0;
}
