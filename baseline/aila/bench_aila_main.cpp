// Driver of the Aila-Laine comparator: this repository's stand-in for the reference's
// tools/bench_aila/bench_aila.cpp (which needs the AnyDSL runtime for its arrays and loaders).  Same options
// (:26-37), same report lines (:140-150), same preparation of the node array before setup_traversal (:39-50: the
// kernel wants the bounds of a Node2 as c0.x c1.x c0.y c1.y | c0.z c1.z, i.e. the z and second-child y entries swapped
// in place).  The three functions below are defined by the patched kepler_dynamic_fetch.cu (build.py).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "formats.h"

void setup_traversal(const Node2* nodes, size_t num_nodes, const Tri1* tris, size_t num_tris);
void shutdown_traversal();
void bench_traversal(const Ray1* rays, Hit1* hits, int num_rays, double* timings, int ntimes, bool any);

int main(int argc, char** argv) {
    std::string ray_file, bvh_file, out_file;
    float tmin = 0.0f, tmax = 1e9f;
    int iters = 1, warmup = 0;
    bool any_hit = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&]() -> const char* {
            if (i + 1 >= argc) { std::cerr << "Missing argument for " << a << std::endl; std::exit(1); }
            return argv[++i];
        };
        if (a == "-bvh" || a == "--bvh-file") bvh_file = value();
        else if (a == "-ray" || a == "--ray-file") ray_file = value();
        else if (a == "--tmin") tmin = std::strtof(value(), nullptr);
        else if (a == "--tmax") tmax = std::strtof(value(), nullptr);
        else if (a == "--bench") iters = int(std::strtol(value(), nullptr, 10));
        else if (a == "--warmup") warmup = int(std::strtol(value(), nullptr, 10));
        else if (a == "-any") any_hit = true;
        else if (a == "-o" || a == "--output") out_file = value();
        else { std::cerr << "Unknown option '" << a << "'" << std::endl; return 1; }
    }
    if (bvh_file.empty()) { std::cerr << "No BVH file specified" << std::endl; return 1; }
    if (ray_file.empty()) { std::cerr << "No ray file specified" << std::endl; return 1; }

    std::vector<Node2> nodes; std::vector<Tri1> tris;
    if (!rb200::read_bvh(bvh_file, rb200::kBvh2Tri1, nodes, tris)) { std::cerr << "Cannot load BVH file" << std::endl; return 1; }
    for (Node2& n : nodes) {                      // bench_aila.cpp:39-50
        float* b = reinterpret_cast<float*>(&n);
        const float y1lo = b[6], y1hi = b[7], z0lo = b[8], z0hi = b[9], z1a = b[4], z1b = b[5];
        b[4] = y1lo; b[5] = y1hi; b[6] = z0lo; b[7] = z0hi; b[8] = z1a; b[9] = z1b;
    }
    std::vector<Ray1> rays;
    if (!rb200::read_rays(ray_file, tmin, tmax, rays)) { std::cerr << "Cannot load rays" << std::endl; return 1; }
    std::vector<Hit1> hits(rays.size());
    std::vector<double> timings(iters);

    setup_traversal(nodes.data(), nodes.size(), tris.data(), tris.size());
    bench_traversal(rays.data(), hits.data(), int(rays.size()), nullptr, warmup, any_hit);
    bench_traversal(rays.data(), hits.data(), int(rays.size()), timings.data(), iters, any_hit);
    shutdown_traversal();

    size_t intr = 0;
    for (auto& hit : hits) intr += (hit.tri_id >= 0);
    if (!out_file.empty()) rb200::write_fbuf(out_file, hits);

    std::sort(timings.begin(), timings.end());
    const double sum = std::accumulate(timings.begin(), timings.end(), 0.0);
    std::cout << sum << "ms for " << iters << " iteration(s)" << std::endl;
    std::cout << rays.size() * iters / (1000.0 * sum) << " Mrays/sec" << std::endl;
    std::cout << "# Average: " << sum / timings.size() << " ms" << std::endl;
    std::cout << "# Median: " << timings[timings.size() / 2] << " ms" << std::endl;
    std::cout << "# Min: " << timings.front() << " ms" << std::endl;
    std::cout << intr << " intersection(s)" << std::endl;
    return 0;
}
