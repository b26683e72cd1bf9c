// bench_traversal -- command-line compatible stand-in for the reference's
// tools/bench_traversal/bench_traversal.cpp (options :25-42, report :381-391), with the
// traversal running on a B200 through the C ABI of include/rodent_b200.h.
//
//   -gpu cuda [-dev k]   device-resident arrays, cuda_{intersect,occluded}_single_ray1_bvh8_tri4,
//                        timed with CUDA events (the role of `-gpu nvvm` + anydsl_get_kernel_time)
//   -s [--bvh-width 4|8] the CPU single-ray call site (bench_cpu_single, :76-82 / :60-66) served by the
//                        host-pointer drop-ins b200_{intersect,occluded}_single_ray1_bvh{4,8}_tri4,
//                        timed with the host clock around the call, copies included
// Packet/hybrid variants and BVH2 inputs are not provided by this library; asking for
// them ends like the reference's variant_not_available() (bench_traversal.impala:15-21).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "formats.h"

namespace {

struct Options {
    std::string bvh_file, ray_file, out_file, gpu;
    float tmin = 0.0f, tmax = 1e9f;
    int iters = 1, warmup = 0, dev = 0, bvh_width = 4, ray_width = 8;
    bool any_hit = false, single = false, packet = false, bvh_width_given = false;
};

[[noreturn]] void fail(const std::string& msg) { std::cerr << msg << std::endl; std::exit(1); }

void usage() {
    std::cout << "Usage: bench_traversal [options]\n"
                 "  -bvh  --bvh-file   BVH file (BVH8_TRI4 block is used)\n"
                 "  -ray  --ray-file   ray file\n"
                 "        --tmin / --tmax   ray interval (default 0 / 1e9)\n"
                 "        --bench / --warmup  timed / untimed iterations (default 1 / 0)\n"
                 "  -gpu  cuda         device-resident arrays on a B200\n"
                 "  -dev  k            CUDA device index\n"
                 "  -any               exit at the first intersection\n"
                 "  -s    --single     host-buffer single-ray entry point\n"
                 "        --bvh-width  4 or 8 (default 4) ; --ray-width 4 or 8 (default 8)\n"
                 "  -o    --output     write hit distances as .fbuf\n";
}

Options parse(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&]() -> const char* {
            if (i + 1 >= argc) fail("Missing argument for " + a);
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); std::exit(0); }
        else if (a == "-bvh" || a == "--bvh-file") o.bvh_file = value();
        else if (a == "-ray" || a == "--ray-file") o.ray_file = value();
        else if (a == "--tmin") o.tmin = std::strtof(value(), nullptr);
        else if (a == "--tmax") o.tmax = std::strtof(value(), nullptr);
        else if (a == "--bench" || a == "--bench-iters") o.iters = int(std::strtol(value(), nullptr, 10));
        else if (a == "--warmup" || a == "--warmup-iters") o.warmup = int(std::strtol(value(), nullptr, 10));
        else if (a == "-gpu" || a == "--gpu-platform") o.gpu = value();
        else if (a == "-dev" || a == "--gpu-device") o.dev = int(std::strtol(value(), nullptr, 10));
        else if (a == "-any") o.any_hit = true;
        else if (a == "-s" || a == "--single") o.single = true;
        else if (a == "-p" || a == "--packet") o.packet = true;
        else if (a == "--bvh-width") { o.bvh_width = int(std::strtol(value(), nullptr, 10)); o.bvh_width_given = true; }
        else if (a == "--ray-width") o.ray_width = int(std::strtol(value(), nullptr, 10));
        else if (a == "-o" || a == "--output") o.out_file = value();
        else if (a[0] == '-') fail("Unknown option '" + a + "'");
        else fail("Invalid argument '" + a + "'");
    }
    if (!o.gpu.empty() && o.gpu != "cuda") fail("Unknown GPU platform '" + o.gpu + "'");
    if (o.bvh_file.empty()) fail("No BVH file specified");
    if (o.ray_file.empty()) fail("No ray file specified");
    if (!o.gpu.empty() && o.single) fail("Options '--gpu' and '--single' are incompatible");
    if (o.single && o.packet) fail("Options '--packet' and '--single' are incompatible");
    if (o.bvh_width != 4 && o.bvh_width != 8) fail("Invalid BVH width");
    if (o.ray_width != 4 && o.ray_width != 8) fail("Invalid ray width");
    return o;
}

[[noreturn]] void variant_not_available(const std::string& name) {
    std::cerr << name << " is not provided by rodent_b200 (use -gpu cuda, or -s)" << std::endl;
    std::abort();
}

// The benchmark proper, for Node8 (BVH8) or Node4 (BVH4) input.
template <typename NodeT, typename DevFn, typename HostFn>
int run(const Options& o, rb200::BlockType block, DevFn dev_intersect, DevFn dev_occluded, HostFn host_intersect, HostFn host_occluded) {
    const bool use_gpu = !o.gpu.empty();
    std::vector<NodeT> nodes; std::vector<Tri4> tris; std::vector<Ray1> rays;
    if (!rb200::read_bvh(o.bvh_file, block, nodes, tris)) fail("Cannot load BVH file");
    if (!rb200::read_rays(o.ray_file, o.tmin, o.tmax, rays)) fail("Cannot load rays");
    const size_t ray_count = rays.size();
    std::cout << ray_count << " ray(s) in the distribution file." << std::endl;
    std::vector<Hit1> hits(ray_count, Hit1{-1, 0.0f, 0.0f, 0.0f});

    std::function<double()> bench;
    NodeT* d_nodes = nullptr; Tri4* d_tris = nullptr; Ray1* d_rays = nullptr; Hit1* d_hits = nullptr;
    if (use_gpu) {
        if (o.dev < 0 || o.dev >= rodent_b200_device_count()) fail("Invalid GPU device");
        auto upload = [&](const void* src, size_t bytes) {
            void* p = rodent_b200_alloc_device(o.dev, bytes);
            rodent_b200_copy_to_device(o.dev, p, src, bytes);
            return p;
        };
        d_nodes = static_cast<NodeT*>(upload(nodes.data(), nodes.size() * sizeof(NodeT)));
        d_tris = static_cast<Tri4*>(upload(tris.data(), tris.size() * sizeof(Tri4)));
        d_rays = static_cast<Ray1*>(upload(rays.data(), rays.size() * sizeof(Ray1)));
        d_hits = static_cast<Hit1*>(upload(hits.data(), hits.size() * sizeof(Hit1)));
        bench = [&] {
            (o.any_hit ? dev_occluded : dev_intersect)(o.dev, d_nodes, d_tris, d_rays, d_hits, int32_t(ray_count));
            return rodent_b200_last_kernel_ms(o.dev);
        };
    } else {
        rodent_b200_set_device(o.dev);
        bench = [&] {
            const auto t0 = std::chrono::steady_clock::now();
            (o.any_hit ? host_occluded : host_intersect)(nodes.data(), tris.data(), rays.data(), hits.data(), int32_t(ray_count));
            return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
    }

    for (int i = 0; i < o.warmup; i++) bench();
    std::vector<double> timings;
    for (int i = 0; i < o.iters; i++) timings.push_back(bench());

    if (use_gpu) rodent_b200_copy_to_host(o.dev, hits.data(), d_hits, hits.size() * sizeof(Hit1));
    size_t intr = 0;
    for (const Hit1& h : hits) intr += h.tri_id >= 0;
    if (!o.out_file.empty() && !rb200::write_fbuf(o.out_file, hits)) fail("Cannot write output file");

    std::sort(timings.begin(), timings.end());
    const double sum = std::accumulate(timings.begin(), timings.end(), 0.0);
    std::cout << sum << "ms for " << o.iters << " iteration(s)" << std::endl;
    std::cout << ray_count * o.iters / (1000.0 * sum) << " Mrays/sec" << std::endl;
    std::cout << "# Average: " << sum / timings.size() << " ms" << std::endl;
    std::cout << "# Median: " << timings[timings.size() / 2] << " ms" << std::endl;
    std::cout << "# Min: " << timings.front() << " ms" << std::endl;
    std::cout << intr << " intersection(s)" << std::endl;
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const Options o = parse(argc, argv);
    const bool use_gpu = !o.gpu.empty();
    if (!use_gpu && !o.single) {
        const std::string kind = o.packet ? "packet" : "hybrid";
        variant_not_available(std::string("cpu_") + (o.any_hit ? "occluded_" : "intersect_") + kind + "_ray" +
                              std::to_string(o.ray_width) + "_bvh" + std::to_string(o.bvh_width) + "_tri4");
    }
    // -gpu cuda traces the BVH8 block unless --bvh-width 4 is asked for; -s follows --bvh-width (default 4) as the reference does
    const int width = use_gpu && !o.bvh_width_given ? 8 : o.bvh_width;
    if (width == 8)
        return run<Node8>(o, rb200::kBvh8Tri4, cuda_intersect_single_ray1_bvh8_tri4, cuda_occluded_single_ray1_bvh8_tri4,
                          b200_intersect_single_ray1_bvh8_tri4, b200_occluded_single_ray1_bvh8_tri4);
    return run<Node4>(o, rb200::kBvh4Tri4, cuda_intersect_single_ray1_bvh4_tri4, cuda_occluded_single_ray1_bvh4_tri4,
                      b200_intersect_single_ray1_bvh4_tri4, b200_occluded_single_ray1_bvh4_tri4);
}
