// ray_gen -- ray-set generator, command-line compatible with the reference's
// tools/ray_gen/ray_gen.cpp (modes primary / shadow / random, :113-134,146-226).
//
// The arithmetic is written out so that the two test sets of the reference are
// regenerated bit for bit (tests/test_fixtures.py checks this when the reference
// tree is present):
//   testing/sponza-primary.rays = ray_gen primary -928.012 483.962 -31.5451  1 0 0  0 1 0  60 1024 1024
//   testing/sponza-random.rays  = ray_gen random sponza.bvh 1048576 42
// In `random` mode the reference draws three floats inside a constructor call
// (ray_gen.cpp:99-100); the build that produced the shipped file evaluated them
// right to left, so the first draw is z.  That order is made explicit here.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "formats.h"

namespace {

struct V3 { float x, y, z; };
V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
V3 normalize(V3 a) { const float k = 1.0f / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return k * a; }

bool put(rb200::File& out, V3 a, V3 b) { const float v[6] = {a.x, a.y, a.z, b.x, b.y, b.z}; return out.write(v, sizeof v); }

// Pinhole camera, rows top to bottom (ray_gen.cpp:20-58).
int primary(int argc, char** argv) {
    if (argc != 15) { std::cerr << "Incorrect number of arguments in primary mode" << std::endl; return 1; }
    auto f = [&](int i) { return std::strtof(argv[i], nullptr); };
    const V3 eye{f(2), f(3), f(4)}, dir0{f(5), f(6), f(7)}, up0{f(8), f(9), f(10)};
    const float fov = f(11);
    const int width = int(std::strtol(argv[12], nullptr, 10)), height = int(std::strtol(argv[13], nullptr, 10));
    const V3 dir = normalize(dir0);
    V3 right = normalize(cross(dir0, up0));
    V3 up = normalize(cross(right, dir0));
    const double scale = std::tan(fov * (M_PI / 360.0f));       // double, as in the reference
    right = float(scale) * right;
    up = float((float(height) / float(width)) * scale) * up;
    rb200::File out(argv[14], "wb");
    if (!out) { std::cerr << "Cannot open output file" << std::endl; return 1; }
    const float sx = 2.0f / width, sy = 2.0f / height;
    for (int i = height - 1; i >= 0; i--)
        for (int j = 0; j < width; j++) {
            const float kx = sx * (j + 0.5f) - 1.0f, ky = sy * (i + 0.5f) - 1.0f;
            put(out, eye, dir + kx * right + ky * up);
        }
    return 0;
}

// Rays from a point light to the primary hit points (ray_gen.cpp:60-85).
int shadow(int argc, char** argv) {
    if (argc != 10) { std::cerr << "Incorrect number of arguments in shadow mode" << std::endl; return 1; }
    auto f = [&](int i) { return std::strtof(argv[i], nullptr); };
    const V3 light{f(2), f(3), f(4)};
    std::vector<Ray1> rays;
    if (!rb200::read_rays(argv[5], 0.0f, 1.0f, rays)) { std::cerr << "Cannot load rays" << std::endl; return 1; }
    std::vector<float> t(rays.size());
    rb200::File fbuf(argv[6], "rb");
    if (!fbuf || !fbuf.read(t.data(), t.size() * 4)) { std::cerr << "Cannot load result of traversal" << std::endl; return 1; }
    rb200::File out(argv[9], "wb");
    if (!out) { std::cerr << "Cannot open output file" << std::endl; return 1; }
    for (size_t i = 0; i < rays.size(); i++) {
        const V3 org{rays[i].org[0], rays[i].org[1], rays[i].org[2]}, dir{rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]};
        put(out, light, (org + t[i] * dir) - light);
    }
    return 0;
}

// Scene bounds = union of the root node's child boxes (ray_gen.cpp:134-144 uses
// the BVH4 block; the BVH8 root gives the same box and is accepted as a fallback).
template <typename NodeT, int N>
bool root_bounds(const std::string& path, rb200::BlockType type, V3& lo, V3& hi) {
    std::vector<NodeT> nodes; std::vector<Tri4> tris;
    if (!rb200::read_bvh(path, type, nodes, tris) || nodes.empty()) return false;
    lo = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    hi = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int i = 0; i < N; i++) {
        const auto& b = nodes[0].bounds;
        lo = {std::fmin(lo.x, b[0][i]), std::fmin(lo.y, b[2][i]), std::fmin(lo.z, b[4][i])};
        hi = {std::fmax(hi.x, b[1][i]), std::fmax(hi.y, b[3][i]), std::fmax(hi.z, b[5][i])};
    }
    return true;
}

// Segments between two uniform points of the scene box (ray_gen.cpp:87-111).
int random_rays(int argc, char** argv) {
    if (argc != 6) { std::cerr << "Incorrect number of arguments in random mode" << std::endl; return 1; }
    const long count = std::strtol(argv[3], nullptr, 10), seed = std::strtol(argv[4], nullptr, 10);
    V3 lo, hi;
    if (!root_bounds<Node4, 4>(argv[2], rb200::kBvh4Tri4, lo, hi) && !root_bounds<Node8, 8>(argv[2], rb200::kBvh8Tri4, lo, hi)) {
        std::cerr << "Cannot extract scene bounds" << std::endl; return 1;
    }
    rb200::File out(argv[5], "wb");
    if (!out) { std::cerr << "Cannot open output file" << std::endl; return 1; }
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<float> dis(0.0f, 1.0f);
    const V3 ext = hi - lo;
    auto point = [&] { V3 p; p.z = dis(gen); p.y = dis(gen); p.x = dis(gen); return lo + ext * p; };
    for (long i = 0; i < count; i++) {
        const V3 a = point(), b = point();
        put(out, a, b - a);
    }
    return 0;
}

void usage() {
    std::cout << "Usage: ray_gen mode arguments output\n"
                 "  primary eye-x eye-y eye-z dir-x dir-y dir-z up-x up-y up-z fov width height\n"
                 "  shadow  light-x light-y light-z ray-file fbuf-file width height\n"
                 "  random  bvh-file ray-count seed\n";
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "Not enough arguments" << std::endl; return 1; }
    if (!std::strcmp(argv[1], "primary")) return primary(argc, argv);
    if (!std::strcmp(argv[1], "shadow")) return shadow(argc, argv);
    if (!std::strcmp(argv[1], "random")) return random_rays(argc, argv);
    if (!std::strcmp(argv[1], "-h") || !std::strcmp(argv[1], "--help")) { usage(); return 0; }
    std::cerr << "Unknown mode" << std::endl;
    return 1;
}
