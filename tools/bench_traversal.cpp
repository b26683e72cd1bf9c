// bench_traversal -- command-line compatible stand-in for the reference's
// tools/bench_traversal/bench_traversal.cpp (options :25-42, report :381-391), with the
// traversal running on a B200 through the C ABI of include/rodent_b200.h.
//
//   -gpu cuda [-dev k]   device-resident arrays, cuda_{intersect,occluded}_single_ray1_bvh8_tri4 (--bvh-width 4: _bvh4_tri4,
//                        --bvh-width 2: _bvh2_tri1, the block and semantics of the reference's `-gpu nvvm`),
//                        timed with CUDA events (the role of `-gpu nvvm` + anydsl_get_kernel_time)
//   -s [--bvh-width 4|8] the CPU single-ray call site (bench_cpu_single, :76-82 / :60-66) served by the
//                        host-pointer drop-ins b200_{intersect,occluded}_single_ray1_bvh{4,8}_tri4,
//                        timed with the host clock around the call, copies included
//   (default) / -p       the hybrid / packet call sites (:44-74, :84-122) served by b200_*_{hybrid,packet}_ray{4,8}_bvh{4,8}_tri4:
//                        every ray of a packet is traced by the single-ray kernel
// BVH2 input is not provided by this library.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <numeric>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "formats.h"

namespace {

struct Options {
    std::string bvh_file, ray_file, out_file, gpu;
    float tmin = 0.0f, tmax = 1e9f;
    int iters = 1, warmup = 0, dev = 0, bvh_width = 4, ray_width = 8, gpus = 1;
    bool any_hit = false, single = false, packet = false, bvh_width_given = false, pinned = false, packet_order = false;
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
                 "        --gpus n     with -s: cut every call into n contiguous ray ranges over devices dev .. dev+n-1 (BVH replicated)\n"
                 "  -any               exit at the first intersection\n"
                 "  -s    --single     host-buffer single-ray entry point\n"
                 "        --pinned     with -s: page-lock the ray and hit arrays in place (rodent_b200_pin_host)\n"
                 "        --packet-order  without -s: walk in the reference's packet order (records of its packet / hybrid kernels to the bit; slower)\n"
                 "        --bvh-width  4 or 8 (default 4; 2 with -gpu: the reference GPU path's BVH2 block) ; --ray-width 4 or 8 (default 8)\n"
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
        else if (a == "--gpus") o.gpus = int(std::strtol(value(), nullptr, 10));
        else if (a == "-any") o.any_hit = true;
        else if (a == "--pinned") o.pinned = true;
        else if (a == "--packet-order") o.packet_order = true;
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
    if (o.bvh_width != 4 && o.bvh_width != 8 && !(o.bvh_width == 2 && !o.gpu.empty())) fail("Invalid BVH width");
    if (o.ray_width != 4 && o.ray_width != 8) fail("Invalid ray width");
    if (o.gpus < 1 || (o.gpus > 1 && !o.single)) fail("Option '--gpus' needs '--single' (the host-pointer entry points shard a call over the devices)");
    return o;
}

template <typename F> struct FnArgs;
template <typename R, typename... A> struct FnArgs<R (*)(A...)> { using type = std::tuple<A...>; };

// The benchmark proper, for Node8 (BVH8), Node4 (BVH4) or -- the reference GPU path's own layout -- Node2 / Tri1 input.
template <typename NodeT, typename TriT = Tri4, typename DevFn, typename HostFn>
int run(const Options& o, rb200::BlockType block, DevFn dev_intersect, DevFn dev_occluded, HostFn host_intersect, HostFn host_occluded) {
    const bool use_gpu = !o.gpu.empty();
    std::vector<NodeT> nodes; std::vector<TriT> tris; std::vector<Ray1> rays;
    if (!rb200::read_bvh(o.bvh_file, block, nodes, tris)) fail("Cannot load BVH file");
    if (!rb200::read_rays(o.ray_file, o.tmin, o.tmax, rays)) fail("Cannot load rays");
    const size_t ray_count = rays.size();
    std::cout << ray_count << " ray(s) in the distribution file." << std::endl;
    std::vector<Hit1> hits(ray_count, Hit1{-1, 0.0f, 0.0f, 0.0f});

    std::function<double()> bench;
    NodeT* d_nodes = nullptr; TriT* d_tris = nullptr; Ray1* d_rays = nullptr; Hit1* d_hits = nullptr;
    if (use_gpu) {
        if (o.dev < 0 || o.dev >= rodent_b200_device_count()) fail("Invalid GPU device");
        auto upload = [&](const void* src, size_t bytes) {
            void* p = rodent_b200_alloc_device(o.dev, bytes);
            rodent_b200_copy_to_device(o.dev, p, src, bytes);
            return p;
        };
        d_nodes = static_cast<NodeT*>(upload(nodes.data(), nodes.size() * sizeof(NodeT)));
        d_tris = static_cast<TriT*>(upload(tris.data(), tris.size() * sizeof(TriT)));
        d_rays = static_cast<Ray1*>(upload(rays.data(), rays.size() * sizeof(Ray1)));
        d_hits = static_cast<Hit1*>(upload(hits.data(), hits.size() * sizeof(Hit1)));
        bench = [&] {
            (o.any_hit ? dev_occluded : dev_intersect)(o.dev, d_nodes, d_tris, d_rays, d_hits, int32_t(ray_count));
            return rodent_b200_last_kernel_ms(o.dev);
        };
    } else {
        if (o.dev < 0 || o.dev + o.gpus > rodent_b200_device_count()) fail("Invalid GPU device");
        if (o.gpus > 1) {
            std::vector<int32_t> devs(o.gpus);
            std::iota(devs.begin(), devs.end(), o.dev);
            rodent_b200_set_devices(devs.data(), o.gpus);
        } else {
            rodent_b200_set_device(o.dev);
        }
        // the two lines a host adds to get one launch per call and no staging: its own arrays, page-locked in place
        if (o.pinned && ray_count > 0 &&
            (rodent_b200_pin_host(rays.data(), ray_count * sizeof(Ray1)) != 0 || rodent_b200_pin_host(hits.data(), ray_count * sizeof(Hit1)) != 0))
            fail("Cannot page-lock the ray / hit arrays");
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
    if (!use_gpu && o.pinned && ray_count > 0) { rodent_b200_unpin_host(rays.data()); rodent_b200_unpin_host(hits.data()); }
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

// Packet / hybrid variants: RayW packets in, HitW packets out (bench_cpu_packet / bench_cpu_hybrid, bench_traversal.cpp:44-122).
template <typename NodeT, int W, typename Fn>
int run_packets(const Options& o, rb200::BlockType block, Fn intersect, Fn occluded) {
    std::vector<NodeT> nodes; std::vector<Tri4> tris; std::vector<Ray1> rays;
    if (!rb200::read_bvh(o.bvh_file, block, nodes, tris)) fail("Cannot load BVH file");
    if (!rb200::read_rays(o.ray_file, o.tmin, o.tmax, rays)) fail("Cannot load rays");
    const size_t num_packets = rays.size() / W, ray_count = num_packets * W;        // whole packets only, load_rays.h:74-76
    std::cout << ray_count << " ray(s) in the distribution file." << std::endl;
    std::vector<float> packets(num_packets * 8 * W), hits(num_packets * 4 * W, 0.0f);
    for (size_t p = 0; p < num_packets; p++)
        for (int j = 0; j < W; j++) {
            const Ray1& r = rays[p * W + j];
            float* q = &packets[p * 8 * W + j];
            for (int c = 0; c < 3; c++) { q[c * W] = r.org[c]; q[(3 + c) * W] = r.dir[c]; }
            q[6 * W] = r.tmin; q[7 * W] = r.tmax;
        }
    rodent_b200_set_device(o.dev);
    rodent_b200_set_packet_order(o.packet_order ? 1 : 0);
    using RayW = std::remove_pointer_t<std::tuple_element_t<2, typename FnArgs<Fn>::type>>;
    using HitW = std::remove_pointer_t<std::tuple_element_t<3, typename FnArgs<Fn>::type>>;
    auto bench = [&] {
        const auto t0 = std::chrono::steady_clock::now();
        (o.any_hit ? occluded : intersect)(nodes.data(), tris.data(), reinterpret_cast<RayW*>(packets.data()),
                                           reinterpret_cast<HitW*>(hits.data()), int32_t(num_packets));
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    for (int i = 0; i < o.warmup; i++) bench();
    std::vector<double> timings;
    for (int i = 0; i < o.iters; i++) timings.push_back(bench());
    size_t intr = 0;
    std::vector<Hit1> flat(ray_count);
    for (size_t p = 0; p < num_packets; p++)
        for (int j = 0; j < W; j++) {
            const float* q = &hits[p * 4 * W + j];
            int32_t id; std::memcpy(&id, q, 4);
            flat[p * W + j] = Hit1{id, q[W], q[2 * W], q[3 * W]};
            intr += id >= 0;
        }
    if (!o.out_file.empty() && !rb200::write_fbuf(o.out_file, flat)) fail("Cannot write output file");
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
        // the packet / hybrid call sites (the reference's default is hybrid, ray width 8, BVH width 4)
        if (o.bvh_width == 4) {
            if (o.packet) { if (o.ray_width == 4) return run_packets<Node4, 4>(o, rb200::kBvh4Tri4, b200_intersect_packet_ray4_bvh4_tri4, b200_occluded_packet_ray4_bvh4_tri4);
                            return run_packets<Node4, 8>(o, rb200::kBvh4Tri4, b200_intersect_packet_ray8_bvh4_tri4, b200_occluded_packet_ray8_bvh4_tri4); }
            if (o.ray_width == 4) return run_packets<Node4, 4>(o, rb200::kBvh4Tri4, b200_intersect_hybrid_ray4_bvh4_tri4, b200_occluded_hybrid_ray4_bvh4_tri4);
            return run_packets<Node4, 8>(o, rb200::kBvh4Tri4, b200_intersect_hybrid_ray8_bvh4_tri4, b200_occluded_hybrid_ray8_bvh4_tri4);
        }
        if (o.packet) { if (o.ray_width == 4) return run_packets<Node8, 4>(o, rb200::kBvh8Tri4, b200_intersect_packet_ray4_bvh8_tri4, b200_occluded_packet_ray4_bvh8_tri4);
                        return run_packets<Node8, 8>(o, rb200::kBvh8Tri4, b200_intersect_packet_ray8_bvh8_tri4, b200_occluded_packet_ray8_bvh8_tri4); }
        if (o.ray_width == 4) return run_packets<Node8, 4>(o, rb200::kBvh8Tri4, b200_intersect_hybrid_ray4_bvh8_tri4, b200_occluded_hybrid_ray4_bvh8_tri4);
        return run_packets<Node8, 8>(o, rb200::kBvh8Tri4, b200_intersect_hybrid_ray8_bvh8_tri4, b200_occluded_hybrid_ray8_bvh8_tri4);
    }
    // -gpu cuda traces the BVH8 block unless --bvh-width 4 is asked for; -s follows --bvh-width (default 4) as the reference does
    const int width = use_gpu && !o.bvh_width_given ? 8 : o.bvh_width;
    if (width == 2) {       // what `-gpu nvvm` traces in the reference: the BVH2 / Tri1 block (bench_traversal.cpp:250-262)
        using HostFn = void (*)(const Node2*, const Tri1*, const Ray1*, Hit1*, int32_t);
        return run<Node2, Tri1>(o, rb200::kBvh2Tri1, cuda_intersect_single_ray1_bvh2_tri1, cuda_occluded_single_ray1_bvh2_tri1,
                                HostFn(nullptr), HostFn(nullptr));
    }
    if (width == 8)
        return run<Node8>(o, rb200::kBvh8Tri4, cuda_intersect_single_ray1_bvh8_tri4, cuda_occluded_single_ray1_bvh8_tri4,
                          b200_intersect_single_ray1_bvh8_tri4, b200_occluded_single_ray1_bvh8_tri4);
    return run<Node4>(o, rb200::kBvh4Tri4, cuda_intersect_single_ray1_bvh4_tri4, cuda_occluded_single_ray1_bvh4_tri4,
                      b200_intersect_single_ray1_bvh4_tri4, b200_occluded_single_ray1_bvh4_tri4);
}
