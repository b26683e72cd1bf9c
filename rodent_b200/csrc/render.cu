// Wavefront path tracer for sm_100a: the B200 counterpart of gpu_streaming_trace
// (src/render/mapping_gpu.impala:308-369) behind render() / the rodent_b200_render* C ABI.
//
// One wavefront = up to 2 Mi rays (twice the reference's stream capacity, :319) and seven launches, all
// fed from device-side state: the host never waits for a wavefront (the reference blocks on four
// copies per wavefront, :201,208,279,298).  It enqueues wavefronts a few ahead and learns from a
// small asynchronous copy, some wavefronts later, that the loop has ended:
//
//   plan              one thread: survivors of the last wavefront -> size of this one, how many
//                     camera rays fill the stream's free tail, end of the loop         (host code of :321-366)
//   generate          refill the free tail of the primary stream with camera rays      (:223-265)
//   traverse_primary  persistent closest-hit traversal; writes the hit record and
//                     counts rays per material in shared memory                       (:18-30 + count pass of :191-199)
//   scan              exclusive scan of the per-material counts on the device          (host scan of :201-208)
//   scatter           counting sort of the hit rays by material -- of their INDICES     (:210-220); misses are dropped here
//   shade             gathers its rays in material order; surface element + emission + next-event estimation + bounce,
//                     material looked up in a table; surviving paths and shadow rays are
//                     written compacted (ballot ranks, one atomicAdd per warp), which
//                     replaces the separate compaction pass                            (:82-134 + :267-300)
//   traverse_shadow   persistent any-hit traversal + atomic film accumulation          (:47-80, :32-45)
//
// Streams are structure-of-arrays in HBM with 16-byte elements where fields travel
// together (origin+tmin, direction+tmax, hit record, contribution+mis): 80 bytes per
// primary ray and 52 per shadow ray, as in the reference (src/render/driver.impala:24-61).
#include <algorithm>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>

#include "scene.h"
#include "shading.cuh"
#include "traverse_sched.cuh"
#include "traverse_bvh2.cuh"
#include "nccl_dyn.h"

extern "C" void rodent_b200_count_launches(int64_t n);   // traverse.cu: the library-wide launch counter

namespace rb200 {

static int g_capacity = 1 << 21;             // rays per stream of renderers created from now on (rodent_b200_tune "render_capacity"): twice the reference's
                                              // 1 Mi (mapping_gpu.impala:319) -- fewer, larger wavefronts, each with one tail: +2 % on Sponza, +0.8 % on Cornell
constexpr int kRBlock = 128;
constexpr int kRSmemStack = 24;
constexpr int kRefillMin = 16;                // idle lanes that trigger a refill of the warp (traverse_sched.cuh)
constexpr int kMaxBins = 1025;                // <= 1024 geometries + the miss bin (mapping_gpu.impala:202)

struct PrimaryStream {                        // src/render/driver.impala:36-52
    int* pixel;                               // rays.id
    float4* ray_o;                            // org.xyz, tmin
    float4* ray_d;                            // dir.xyz, tmax
    float4* hit;                              // prim_id (bits), t, u, v
    int* geom;                                // geom_id; num_geometries for a miss
    float4* contrib_mis;                      // contrib.rgb, mis
    uint2* rnd_depth;                         // rnd, depth
};
struct ShadowStream {                         // src/render/driver.impala:54-61
    int* pixel;
    float4* ray_o;
    float4* ray_d;
    float4* color;                            // rgb, unused
};

enum Counter { kWorkPrimary = 0, kWorkShadow, kHitCount, kSurvivors, kShadows, kNumCounters = 8 };

// Device-resident state of one pipeline's wavefront loop (the host variables `id` and `size` of
// gpu_streaming_trace, mapping_gpu.impala:321-323, plus the statistics).
struct LoopState {
    long long next_id, total;                 // next camera sample to generate / samples of this render call
    long long gen_first_id;                   // generate: first sample id,
    int gen_first_dst, gen_n;                 //           where it goes in the stream, how many
    int size;                                 // rays in the primary stream of the current wavefront
    int done;                                 // nothing left: every later wavefront is empty
    long long n_primary, n_shadow, n_waves;
};

struct SceneDev {
    const Node8* nodes; const Tri4* tris;
    const Node2* nodes2; const Tri1* tris1;          // null unless the scene carries a BVH2 (closest-hit rays then use it)
    const Node2q* nodes2q; QuantGrid grid;           // the same tree in 32-byte nodes (render_quant), null when switched off
    const float4* normals; const float4* face_normals; const int4* indices; const int* light_ids;
    const RodentMaterial* materials; const RodentLight* lights;
    const float4* texcoords; const RodentTexture* textures; const unsigned* texture_pixels;   // null without textured materials
    int num_materials, num_lights;
};
struct CameraDev { shade::V3 eye, dir, up, right; float w, h; };

// ---- generate (gpu_generate_rays + make_camera_emitter, renderer.impala:26-40, camera.impala:35-44) ----
// ---- plan: the host part of the loop body (mapping_gpu.impala:325-366), one thread on the device ----
// `prev`: the counters the previous wavefront's shade kernel left (null for the first wavefront of a render call);
// `cur`: this wavefront's counters, reset here.
__global__ void plan_wavefront(LoopState* __restrict__ st, const int* __restrict__ prev, int* __restrict__ cur, int capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int size = st->size;
    if (prev) { size = prev[kSurvivors]; st->n_shadow += prev[kShadows]; }
    const long long left = st->total - st->next_id;
    const int n = int(left < (long long)(capacity - size) ? left : (long long)(capacity - size));     // :326-333
    st->gen_first_id = st->next_id; st->gen_first_dst = size; st->gen_n = n;
    st->next_id += n;
    size += n;
    st->size = size;
    st->done = size == 0;                                                                              // :323
    if (size > 0) { st->n_primary += size; st->n_waves += 1; }
    for (int k = 0; k < kNumCounters; k++) cur[k] = 0;
}

__global__ void __launch_bounds__(256)
generate_rays(PrimaryStream s, const LoopState* __restrict__ st, CameraDev cam, int width, int height, int spp, int iter,
              const int* __restrict__ rows) {
    using namespace shade;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= st->gen_n) return;
    const long long ray_id = st->gen_first_id + gid;
    const int dst = st->gen_first_dst + gid;
    const int sample = int(ray_id % spp), local_pixel = int(ray_id / spp);
    const int ly = local_pixel / width, x = local_pixel - ly * width;
    const int y = __ldg(rows + ly);
    unsigned rnd = fnv_hash(fnv_hash(fnv_hash(fnv_hash(0x811C9DC5u, unsigned(sample)), unsigned(iter)), unsigned(x)), unsigned(y));
    const float kx = 2.0f * (float(x) + randf(rnd)) / float(width) - 1.0f;
    const float ky = 1.0f - 2.0f * (float(y) + randf(rnd)) / float(height);
    const V3 d = normalize(cam.right * (cam.w * kx) + cam.up * (cam.h * ky) + cam.dir);
    s.pixel[dst] = y * width + x;
    s.ray_o[dst] = make_float4(cam.eye.x, cam.eye.y, cam.eye.z, 0.0f);
    s.ray_d[dst] = make_float4(d.x, d.y, d.z, kFltMax);
    s.contrib_mis[dst] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    s.rnd_depth[dst] = make_uint2(rnd, 0u);
}

// ---- persistent traversal over a stream -----------------------------------------------------
// SHADOW = false: closest hit, hit record + geometry id + per-material count.
// SHADOW = true : any hit; unoccluded rays add their colour to the film.
template <bool SHADOW, bool WIDE = false, bool FMA = false>
__global__ void __launch_bounds__(kRBlock, 5)
traverse_stream(const Node8* __restrict__ nodes, const Tri4* __restrict__ tris,
                const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const int* __restrict__ count_ptr, int count_max,
                float4* __restrict__ hit_out, int* __restrict__ geom_out, int num_geoms, int* __restrict__ histogram,
                const int* __restrict__ pixels, const float4* __restrict__ colors, float* __restrict__ film, float inv_spp,
                int* __restrict__ work_counter, int refill_min) {
    __shared__ StackEntry smem_stack[kRSmemStack][kRBlock];
    __shared__ int hist[SHADOW ? 1 : kMaxBins];
    const int num_rays = count_ptr ? min(*count_ptr, count_max) : count_max;
    if (!SHADOW) {
        for (int b = threadIdx.x; b <= num_geoms; b += kRBlock) hist[b] = 0;
        __syncthreads();
    }
    int* const hist_bins = hist;
    // the vote-scheduled persistent loop of traverse_sched.cuh, fed from / draining into the SoA streams
    traverse_vote_scheduled<SHADOW, !SHADOW, kRSmemStack, kRBlock, 8, WIDE, FMA>(      // WIDE: 256-bit record loads (the scene arrays are cudaMalloc'ed)
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min,
        [ray_o, ray_d](int i, float4& r0, float4& r1) { r0 = __ldg(ray_o + i); r1 = __ldg(ray_d + i); },
        [=](int i, const HitRecord& h) {
            if (SHADOW) {
                if (h.prim < 0) {                                              // gpu_accumulate, mapping_gpu.impala:32-45
                    const int p = __ldg(pixels + i);
                    const float4 c = __ldg(colors + i);
                    atomicAdd(film + 3 * p + 0, c.x * inv_spp);
                    atomicAdd(film + 3 * p + 1, c.y * inv_spp);
                    atomicAdd(film + 3 * p + 2, c.z * inv_spp);
                }
            } else {
                // make_primary_stream_hit_writer, driver.impala:106-115
                const int g = h.prim < 0 ? num_geoms : h.geom;
                hit_out[i] = make_float4(__int_as_float(h.prim), h.t, h.u, h.v);
                geom_out[i] = g;
                atomicAdd(hist_bins + g, 1);
            }
        });
    if (!SHADOW) {
        __syncthreads();
        for (int b = threadIdx.x; b <= num_geoms; b += kRBlock)
            if (hist[b]) atomicAdd(histogram + b, hist[b]);
    }
}

// The same two kernels on the scene's BVH2 / Tri1 -- the layout and traversal of the reference's GPU renderer
// (gpu_traverse_primary / gpu_traverse_secondary over make_gpu_bvh2_tri1, mapping_gpu.impala:18-80,505-509) -- same stream
// contract as above.
// STACK: levels of the id stack kept in shared memory (deeper ones go to a thread-local array).  The per-material counts
// live in dynamic shared memory, (num_geoms + 1) ints for the closest-hit form: what a CTA does not take as shared memory
// the SM keeps as L1, which serves two thirds of this kernel's record reads.
template <bool SHADOW, int STACK, bool FMA = false, bool QUANT = false>
__global__ void __launch_bounds__(kRBlock, 8)
traverse_stream_bvh2(const void* __restrict__ nodes, QuantGrid grid, const Tri1* __restrict__ tris,
                     const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const int* __restrict__ count_ptr, int count_max,
                     float4* __restrict__ hit_out, int* __restrict__ geom_out, int num_geoms, int* __restrict__ histogram,
                     const int* __restrict__ pixels, const float4* __restrict__ colors, float* __restrict__ film, float inv_spp,
                     int* __restrict__ work_counter, int refill_min, int streak_min, int leaf_streak_min) {
    __shared__ int smem_stack[STACK][kRBlock];
    extern __shared__ int hist[];
    const int num_rays = count_ptr ? min(*count_ptr, count_max) : count_max;
    if (!SHADOW) {
        for (int b = threadIdx.x; b <= num_geoms; b += kRBlock) hist[b] = 0;
        __syncthreads();
    }
    int* const hist_bins = hist;
    traverse_bvh2_scheduled<SHADOW, STACK, kRBlock, FMA, QUANT>(
        nodes, tris, &smem_stack[0][threadIdx.x], num_rays, work_counter, refill_min, streak_min,
        [ray_o, ray_d](int i, float4& r0, float4& r1) { r0 = __ldg(ray_o + i); r1 = __ldg(ray_d + i); },
        [=](int i, const HitRecord& h) {
            if (SHADOW) {
                if (h.prim < 0) {
                    const int p = __ldg(pixels + i);
                    const float4 c = __ldg(colors + i);
                    atomicAdd(film + 3 * p + 0, c.x * inv_spp);
                    atomicAdd(film + 3 * p + 1, c.y * inv_spp);
                    atomicAdd(film + 3 * p + 2, c.z * inv_spp);
                }
            } else {
                const int g = h.prim < 0 ? num_geoms : h.geom;
                hit_out[i] = make_float4(__int_as_float(h.prim), h.t, h.u, h.v);
                geom_out[i] = g;
                atomicAdd(hist_bins + g, 1);
            }
        }, leaf_streak_min, grid);
    if (!SHADOW) {
        __syncthreads();
        for (int b = threadIdx.x; b <= num_geoms; b += kRBlock)
            if (hist[b]) atomicAdd(histogram + b, hist[b]);
    }
}

// ---- scan: per-material begins, number of hit rays (the host scan of mapping_gpu.impala:201-208) ----
__global__ void scan_bins(int* __restrict__ histogram, int* __restrict__ cursor, int num_geoms, int* __restrict__ counters) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0;
    for (int b = 0; b <= num_geoms; b++) {
        cursor[b] = n;
        if (b == num_geoms) counters[kHitCount] = n;     // the miss bin comes last and is not shaded (:357)
        n += histogram[b];
        histogram[b] = 0;
    }
}

// ---- scatter: counting sort by material (gpu_sort_primary third pass, :210-220) ----
// The reference moves the whole 80-byte ray to its sorted position (copy_primary_ray, :136-164).  Here only the
// ray's INDEX is moved: order[d] = i, 8 bytes of traffic per ray instead of 164, and the shade kernel gathers its
// inputs through `order`.  Slots are handed out in two levels: a CTA counts its rays per material in shared memory (one
// shared-memory atomic per warp and material, peers found with __match_any_sync), takes ONE range per material from the
// global cursor, and its rays fill that range -- with a handful of materials (the Cornell box has four) the global
// cursors would otherwise serialise a hundred thousand atomics per wavefront on four addresses (33 us of a 1 Mi-ray
// wavefront in round 1, a seventh of the Cornell render).
constexpr int kScatterBlock = 512;
__global__ void __launch_bounds__(kScatterBlock)
scatter_by_material(PrimaryStream src, int* __restrict__ order, const LoopState* __restrict__ st, int num_geoms, int* __restrict__ cursor) {
    extern __shared__ int bins[];                              // [num_geoms] counts, then [num_geoms] bases
    const int size = st->size;
    if (int(blockIdx.x) * kScatterBlock >= size) return;       // whole CTA beyond the stream
    int* const count = bins;
    int* const base = bins + num_geoms;
    for (int b = threadIdx.x; b < num_geoms; b += kScatterBlock) count[b] = 0;
    __syncthreads();
    const int i = blockIdx.x * kScatterBlock + threadIdx.x;
    const int g = i < size ? src.geom[i] : num_geoms;
    const bool live = g < num_geoms;                           // misses (the last bin) are dropped here
    const unsigned peers = __match_any_sync(0xffffffffu, live ? g : -1 - int(lane_id()));
    const int leader = __ffs(peers) - 1;
    int rank = 0;
    if (live && int(lane_id()) == leader) rank = atomicAdd(count + g, __popc(peers));
    rank = __shfl_sync(0xffffffffu, rank, leader) + __popc(peers & lanemask_lt());
    __syncthreads();
    for (int b = threadIdx.x; b < num_geoms; b += kScatterBlock)
        base[b] = count[b] ? atomicAdd(cursor + b, count[b]) : 0;
    __syncthreads();
    if (live) order[base[g] + rank] = i;
}

// ---- shade (gpu_shade, mapping_gpu.impala:82-134, with the path tracer of renderer.impala:62-162) ----
// MIN_BLOCKS: resident CTAs per SM asked of ptxas (6: 74 registers, 8: 64 with 16 bytes spilled, 10: 48 with 90): the kernel
// waits on gathered loads (66 % long-scoreboard stalls at 22 warps per SM), so it trades registers for warps.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS)
shade_rays(PrimaryStream in, const int* __restrict__ order, PrimaryStream out, ShadowStream shadow, SceneDev sc,
           int* __restrict__ counters, float* __restrict__ film, float inv_spp, int max_path_len) {
    using namespace shade;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;       // position in material order
    const bool live = k < counters[kHitCount];
    const int i = live ? __ldg(order + k) : 0;                 // the ray's place in the stream
    bool emit_shadow = false, bounce = false;
    int pixel = 0;
    float4 sh_o, sh_d, sh_c, b_o, b_d, b_c;
    uint2 b_r;
    if (live) {
        pixel = in.pixel[i];
        const float4 ro = in.ray_o[i], rd = in.ray_d[i], h = in.hit[i], cm = in.contrib_mis[i];
        const uint2 rdp = in.rnd_depth[i];
        RodentMaterial mat = sc.materials[in.geom[i]];
        apply_textures(mat, sc.texcoords, sc.indices, sc.textures, sc.texture_pixels, __float_as_int(h.x), h.z, h.w);
        const V3 org = v3(ro.x, ro.y, ro.z), dir = v3(rd.x, rd.y, rd.z);
        const int prim = __float_as_int(h.x);
        const float t = h.y;
        Col contrib = col(cm.x, cm.y, cm.z);
        const float mis = cm.w;
        unsigned rnd = rdp.x;
        const int depth = int(rdp.y);
        const Surf surf = surface_element(sc.normals, sc.face_normals, sc.indices, org, dir, prim, t, h.z, h.w);
        const V3 out_dir = -dir;
        const float pdf_lightpick = 1.0f / float(sc.num_lights);                              // renderer.impala:65

        // on_hit, renderer.impala:111-127
        if (mat.is_emissive && surf.is_entering) {
            const RodentLight& l = sc.lights[sc.light_ids[prim]];
            const float pdf_dir = cosine_hemisphere_pdf(dot(v3(l.n[0], l.n[1], l.n[2]), out_dir));
            Col intensity = col(0, 0, 0); float pdf_area = 1.0f;                               // make_emission_value, light.impala:94-108
            if (pdf_dir > 0.0f) { intensity = col(mat.ke[0], mat.ke[1], mat.ke[2]); pdf_area = l.inv_area; }   // (= the light's colour; map_Ke sampled at the hit by apply_textures)
            const float next_mis = mis * t * t / dot(out_dir, surf.local.c2);
            const float w = 1.0f / (1.0f + next_mis * pdf_lightpick * pdf_area);
            const Col c = (contrib * intensity) * w;
            atomicAdd(film + 3 * pixel + 0, c.r * inv_spp);
            atomicAdd(film + 3 * pixel + 1, c.g * inv_spp);
            atomicAdd(film + 3 * pixel + 2, c.b * inv_spp);
        }

        // on_shadow, renderer.impala:69-109
        if (!is_specular(mat) && sc.num_lights > 0) {
            const int light_id = (xorshift(rnd) & 0x7FFFFFFF) % sc.num_lights;
            const RodentLight& l = sc.lights[light_id];
            float u = randf(rnd), v = randf(rnd);                                              // sample_triangle, random.impala:51-61
            if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
            const V3 pos = v3(l.v0[0], l.v0[1], l.v0[2]) * (1.0f - v - u) + v3(l.v1[0], l.v1[1], l.v1[2]) * u + v3(l.v2[0], l.v2[1], l.v2[2]) * v;
            const V3 from_dir = surf.point - pos;
            float cos_l = dot(from_dir, v3(l.n[0], l.n[1], l.n[2])) / length(from_dir);
            Col intensity = light_color(l, sc.texcoords, sc.indices, sc.textures, sc.texture_pixels, u, v);
            float pdf_area = l.inv_area;
            if (!(pdf_area > 0.0f && cosine_hemisphere_pdf(cos_l) > 0.0f && cos_l > 0.0f)) {   // make_direct_sample, light.impala:76-92
                intensity = col(0, 0, 0); pdf_area = 1.0f; cos_l = 0.0f;
            }
            const V3 light_dir = pos - surf.point;
            const float vis = dot(light_dir, surf.local.c2);
            if (vis > 0.0f && cos_l > 0.0f) {
                const float inv_d = 1.0f / length(light_dir), inv_d2 = inv_d * inv_d;
                const V3 in_dir = light_dir * inv_d;
                const float pdf_e = bsdf_pdf(mat, surf, in_dir, out_dir);
                const float inv_pdf_l = 1.0f / (pdf_area * pdf_lightpick);
                const float cos_e = vis * inv_d;
                const float w = 1.0f / (1.0f + pdf_e * cos_l * inv_d2 * inv_pdf_l);
                const float geom_factor = cos_e * cos_l * inv_d2 * inv_pdf_l;
                const Col c = (intensity * (contrib * bsdf_eval(mat, surf, in_dir, out_dir))) * (geom_factor * w);
                emit_shadow = true;
                sh_o = make_float4(surf.point.x, surf.point.y, surf.point.z, kOffset);
                sh_d = make_float4(light_dir.x, light_dir.y, light_dir.z, 1.0f - kOffset);
                sh_c = make_float4(c.r, c.g, c.b, 0.0f);
            }
        }

        // on_bounce, renderer.impala:129-152
        float rr = 2.0f * luminance(contrib);                                                   // russian_roulette, random.impala:128-131
        if (rr > 0.75f) rr = 0.75f;
        if (!(depth >= max_path_len || randf(rnd) >= rr)) {
            const BsdfSample bs = bsdf_sample(mat, surf, rnd, out_dir);
            contrib = (contrib * bs.color) * (bs.cos / (bs.pdf * rr));
            bounce = true;
            b_o = make_float4(surf.point.x, surf.point.y, surf.point.z, kOffset);
            b_d = make_float4(bs.in_dir.x, bs.in_dir.y, bs.in_dir.z, kFltMax);
            b_c = make_float4(contrib.r, contrib.g, contrib.b, is_specular(mat) ? 0.0f : 1.0f / bs.pdf);
            b_r = make_uint2(rnd, unsigned(depth + 1));
        }
    }
    // compacted output: ranks from ballots, one atomicAdd per warp and stream
    const unsigned lane = lane_id();
    const unsigned mb = __ballot_sync(0xffffffffu, bounce), ms = __ballot_sync(0xffffffffu, emit_shadow);
    int base_b = 0, base_s = 0;
    if (lane == 0) {
        if (mb) base_b = atomicAdd(counters + kSurvivors, __popc(mb));
        if (ms) base_s = atomicAdd(counters + kShadows, __popc(ms));
    }
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    base_s = __shfl_sync(0xffffffffu, base_s, 0);
    if (bounce) {
        const int d = base_b + __popc(mb & lanemask_lt());
        out.pixel[d] = pixel; out.ray_o[d] = b_o; out.ray_d[d] = b_d; out.contrib_mis[d] = b_c; out.rnd_depth[d] = b_r;
    }
    if (emit_shadow) {
        const int d = base_s + __popc(ms & lanemask_lt());
        shadow.pixel[d] = pixel; shadow.ray_o[d] = sh_o; shadow.ray_d[d] = sh_d; shadow.color[d] = sh_c;
    }
}

// ---- host side --------------------------------------------------------------------------------
constexpr int kLookahead = 3;                   // wavefronts enqueued before the host looks at the loop state again

struct Renderer {
    int dev = 0, width = 0, height = 0, spp = 1, max_path_len = 64;
    int capacity = 1 << 21;                     // rays per stream, fixed at creation
    std::vector<int> rows;                      // image rows owned by this renderer
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_shaded = nullptr, ev_shadow_done = nullptr, ev_done = nullptr;
    PrimaryStream prim[2]{};
    ShadowStream shadow{};
    SceneDev scene{};
    std::vector<void*> allocations;
    int* counters = nullptr; int* histogram = nullptr; int* cursor = nullptr; int* d_rows = nullptr; int* order = nullptr;
    LoopState* state = nullptr;                 // device
    LoopState* h_state = nullptr;               // pinned: kLookahead snapshots + the final one
    cudaEvent_t ev_state[kLookahead] = {};
    float* film = nullptr; float* own_film = nullptr; float* h_film = nullptr;
    int sm_count = 0, occ_primary = 0, occ_shadow = 0, occ_primary2 = 0, occ_shadow2 = 0;
    int64_t stats[5] = {0, 0, 0, 0, 0};
    double last_ms = 0.0;
    // per render call
    int64_t head = 0, tail = 0, n_kernels = 0;  // wavefronts enqueued / whose state snapshot was read
    bool finished = false, generation_over = false;
    int bound = 0;                              // no wavefront from now on holds more rays than this
    // A renderer with several LANES drives that many independent wavefront pipelines (own streams, ray streams and
    // counters, interleaved row bands) into the same film: see render_device.
    std::vector<Renderer*> lanes;
    Renderer* parent = nullptr;
    // Multi-device (rodent_b200_renderer_create_multi): this renderer is the one of devs[0] and owns part 0 of the row
    // bands; `peers` are the renderers of the other devices, `comms` one NCCL communicator per device (this one first).
    std::vector<Renderer*> peers;
    std::vector<ncclComm_t> comms;

    template <typename T>
    T* alloc(size_t n) {
        void* p = nullptr;
        RB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
        allocations.push_back(p);
        return static_cast<T*>(p);
    }
    template <typename T>
    const T* upload(const T* src, size_t n) {
        T* p = alloc<T>(n);
        if (n) RB_CUDA_CHECK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }
};

static void alloc_stream(Renderer& r, PrimaryStream& s) {
    const size_t n = size_t(r.capacity);
    s.pixel = r.alloc<int>(n); s.ray_o = r.alloc<float4>(n); s.ray_d = r.alloc<float4>(n);
    s.hit = r.alloc<float4>(n); s.geom = r.alloc<int>(n); s.contrib_mis = r.alloc<float4>(n);
    s.rnd_depth = r.alloc<uint2>(n);
}

static int g_render_leaf_streak_min = 2;                        // triangle steps follow each other while this many lanes want one (0: as for node steps)
static int g_render_refill_min = 20, g_render_streak_min = 8;   // BVH2 stream kernels: refill threshold, step-streak threshold (swept: profiles/r01_experiments.md)
static int g_render_wide = 0;          // 256-bit record loads in the BVH8 stream kernels (rodent_b200_tune "render_wide")
static int g_render_shadow_bvh2 = 1;   // ... and the shadow rays too (rodent_b200_tune "render_shadow_bvh2"; 0: BVH8 any hit)
static int g_render_bvh2 = 1;      // closest-hit rays through the scene's BVH2 when it has one (rodent_b200_tune "render_bvh2")
static int g_render_lanes = 3;     // pipelines per renderer (rodent_b200_tune "render_lanes")
static int g_render_shade_blocks = 8;   // shade_rays: resident CTAs per SM asked of ptxas (rodent_b200_tune "render_shade_blocks": 6, 8 or 10)
static int g_render_quant = 1;     // BVH2 stream kernels walk 32-byte quantised nodes (rodent_b200_tune "render_quant")
static int g_render_fma = 1;       // contracted slab / triangle arithmetic in the stream kernels (rodent_b200_tune "render_fma")
static int g_render_poly_trig = 0; // test switch: sin / cos from poly_trig.h (rodent_b200_tune "render_poly_trig")
constexpr int kBvh2Stack = 16;     // BVH2 stream kernels: stack levels in shared memory (8 / 16 / 24 / 32 measured 518 / 518 / 518 / 513 Msamples/s)

// Streams, events, ray streams and counters of one wavefront pipeline; `r->rows` must be set.
static void alloc_pipeline(Renderer* r) {
    RB_CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    RB_CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream2, cudaStreamNonBlocking));
    RB_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_shaded, cudaEventDisableTiming));
    RB_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_shadow_done, cudaEventDisableTiming));
    RB_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming));
    for (auto& e : r->ev_state) RB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!r->ev0) { RB_CUDA_CHECK(cudaEventCreate(&r->ev0)); RB_CUDA_CHECK(cudaEventCreate(&r->ev1)); }
    alloc_stream(*r, r->prim[0]);
    alloc_stream(*r, r->prim[1]);
    const size_t n = size_t(r->capacity);
    r->shadow.pixel = r->alloc<int>(n); r->shadow.ray_o = r->alloc<float4>(n);
    r->shadow.ray_d = r->alloc<float4>(n); r->shadow.color = r->alloc<float4>(n);
    r->counters = r->alloc<int>(2 * kNumCounters); r->histogram = r->alloc<int>(kMaxBins); r->cursor = r->alloc<int>(kMaxBins);
    r->order = r->alloc<int>(n);
    r->state = r->alloc<LoopState>(1);
    RB_CUDA_CHECK(cudaMemset(r->histogram, 0, kMaxBins * sizeof(int)));
    r->d_rows = const_cast<int*>(r->upload(r->rows.data(), r->rows.size()));
    RB_CUDA_CHECK(cudaMallocHost(&r->h_state, (kLookahead + 1) * sizeof(LoopState)));
}

// Node2 -> Node2q on a 16-bit grid over the union of all boxes.  Lower bounds are rounded down and upper bounds up, then
// moved two more steps outwards (the kernel evaluates origin + code * step with its own rounding, see set_grid), so a
// quantised box always contains the original.  Returns false (the renderer then walks the 64-byte nodes) when a coordinate is not finite.
static bool quantise_bvh2(const std::vector<Node2>& nodes, std::vector<Node2q>& out, QuantGrid& grid) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (const Node2& n : nodes)
        for (int k = 0; k < 2; k++)
            for (int a = 0; a < 3; a++) {
                const double l = n.bounds[6 * k + 2 * a], h = n.bounds[6 * k + 2 * a + 1];
                if (!std::isfinite(l) || !std::isfinite(h)) return false;
                lo[a] = std::min(lo[a], l); hi[a] = std::max(hi[a], h);
            }
    double step[3];
    for (int a = 0; a < 3; a++) {
        step[a] = std::max((hi[a] - lo[a]) / 65529.0, 1e-30);          // codes 3 .. 65532 span the extent
        grid.origin[a] = float(lo[a] - 3.0 * step[a]);
        grid.step[a] = float(step[a]);
        if (!(grid.step[a] > 0.0f) || !std::isfinite(grid.origin[a])) return false;
        // the margin of one step must cover the kernel's fp32 rounding of origin + code * step: not so for a small scene
        // far from the coordinate origin
        if (std::max(std::fabs(lo[a]), std::fabs(hi[a])) * 2.4e-7 > step[a]) return false;
    }
    out.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
        for (int k = 0; k < 2; k++)
            for (int a = 0; a < 3; a++) {
                const double l = nodes[i].bounds[6 * k + 2 * a], h = nodes[i].bounds[6 * k + 2 * a + 1];
                const double ql = std::floor((l - double(grid.origin[a])) / double(grid.step[a])) - 2.0;
                const double qh = std::ceil((h - double(grid.origin[a])) / double(grid.step[a])) + 2.0;
                out[i].b[6 * k + 2 * a] = uint16_t(std::min(std::max(ql, 0.0), 65535.0));
                out[i].b[6 * k + 2 * a + 1] = uint16_t(std::min(std::max(qh, 0.0), 65535.0));
            }
        out[i].child[0] = nodes[i].child[0]; out[i].child[1] = nodes[i].child[1];
    }
    return true;
}

static Renderer* create_renderer(const Scene& sc, int dev, int width, int height, int spp, int max_path_len, int part, int num_parts, int band) {
    if (sc.materials.size() + 1 > size_t(kMaxBins)) { std::fprintf(stderr, "rodent_b200: more than 1024 materials\n"); return nullptr; }
    if (width <= 0 || height <= 0 || spp <= 0 || num_parts <= 0 || part < 0 || part >= num_parts || band <= 0) return nullptr;
    bool textured = false;
    for (auto& m : sc.materials) {
        if (m.map_kd < 0 || m.map_ks < 0 || m.map_ke < 0 || size_t(m.map_kd) > sc.textures.size() || size_t(m.map_ks) > sc.textures.size() ||
            size_t(m.map_ke) > sc.textures.size()) {
            std::fprintf(stderr, "rodent_b200: a material names texture %d / %d, the scene has %zu\n", m.map_kd, m.map_ks, sc.textures.size());
            return nullptr;
        }
        textured |= (m.map_kd | m.map_ks | m.map_ke) != 0;
    }
    // the stream kernels index per-material tables and shared-memory bins with the ids stored in the BVHs
    const int num_materials = int(sc.materials.size());
    for (const Tri4& t : sc.tris)
        for (int j = 0; j < 4; j++)
            if (t.prim_id[j] != -1 && (t.geom_id[j] < 0 || t.geom_id[j] >= num_materials)) {
                std::fprintf(stderr, "rodent_b200: the scene's BVH8 holds geometry id %d, it has %d materials\n", t.geom_id[j], num_materials);
                return nullptr;
            }
    for (const Tri1& t : sc.tris1)
        if (t.geom_id < 0 || t.geom_id >= num_materials) {
            std::fprintf(stderr, "rodent_b200: the scene's BVH2 holds geometry id %d, it has %d materials\n", t.geom_id, num_materials);
            return nullptr;
        }
    RB_CUDA_CHECK(cudaSetDevice(dev));
    RB_CUDA_CHECK(cudaMemcpyToSymbol(shade::g_poly_trig, &g_render_poly_trig, sizeof(int)));
    auto r = new Renderer();
    r->dev = dev; r->width = width; r->height = height; r->spp = spp; r->max_path_len = max_path_len;
    r->capacity = g_capacity;
    for (int y = 0; y < height; y++)
        if ((y / band) % num_parts == part) r->rows.push_back(y);
    cudaDeviceProp prop;
    RB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    r->sm_count = prop.multiProcessorCount;
    r->film = r->own_film = r->alloc<float>(size_t(width) * height * 3);
    RB_CUDA_CHECK(cudaMemset(r->film, 0, size_t(width) * height * 3 * sizeof(float)));
    RB_CUDA_CHECK(cudaMallocHost(&r->h_film, size_t(width) * height * 3 * sizeof(float)));
    std::memset(r->h_film, 0, size_t(width) * height * 3 * sizeof(float));
    SceneDev& d = r->scene;
    d.nodes = r->upload(sc.nodes.data(), sc.nodes.size());
    d.tris = r->upload(sc.tris.data(), sc.tris.size());
    d.normals = reinterpret_cast<const float4*>(r->upload(sc.normals.data(), sc.normals.size()));
    d.face_normals = reinterpret_cast<const float4*>(r->upload(sc.face_normals.data(), sc.face_normals.size()));
    d.indices = reinterpret_cast<const int4*>(r->upload(sc.indices.data(), sc.indices.size()));
    d.light_ids = r->upload(sc.light_ids.data(), sc.light_ids.size());
    d.materials = r->upload(sc.materials.data(), sc.materials.size());
    d.lights = r->upload(sc.lights.data(), sc.lights.size());
    if (textured) {
        d.texcoords = reinterpret_cast<const float4*>(r->upload(sc.texcoords.data(), sc.texcoords.size()));
        d.textures = r->upload(sc.textures.data(), sc.textures.size());
        d.texture_pixels = r->upload(sc.texture_pixels.data(), sc.texture_pixels.size());
    }
    d.num_materials = num_materials; d.num_lights = int(sc.lights.size());
    if (!sc.nodes2.empty() && g_render_bvh2) {
        d.nodes2 = r->upload(sc.nodes2.data(), sc.nodes2.size());
        d.tris1 = r->upload(sc.tris1.data(), sc.tris1.size());
        if (g_render_quant) {
            std::vector<Node2q> q;
            if (quantise_bvh2(sc.nodes2, q, d.grid)) d.nodes2q = r->upload(q.data(), q.size());
        }
    }
    // persistent grids are sized from the occupancy of the instantiations that are launched
    const size_t bins = (sc.materials.size() + 1) * sizeof(int);
    if (g_render_fma && d.nodes2q) {
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary2, traverse_stream_bvh2<false, kBvh2Stack, true, true>, kRBlock, bins));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow2, traverse_stream_bvh2<true, kBvh2Stack, true, true>, kRBlock, sizeof(int)));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary, traverse_stream<false, false, true>, kRBlock, 0));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow, traverse_stream<true, false, true>, kRBlock, 0));
    } else if (g_render_fma) {
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary2, traverse_stream_bvh2<false, kBvh2Stack, true>, kRBlock, bins));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow2, traverse_stream_bvh2<true, kBvh2Stack, true>, kRBlock, sizeof(int)));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary, traverse_stream<false, false, true>, kRBlock, 0));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow, traverse_stream<true, false, true>, kRBlock, 0));
    } else {
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary2, traverse_stream_bvh2<false, kBvh2Stack, false>, kRBlock, bins));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow2, traverse_stream_bvh2<true, kBvh2Stack, false>, kRBlock, sizeof(int)));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_primary, traverse_stream<false, false, false>, kRBlock, 0));
        RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r->occ_shadow, traverse_stream<true, false, false>, kRBlock, 0));
    }
    // lanes: the rows of this renderer dealt out in bands of eight; small images keep a single pipeline
    const int bands = int((r->rows.size() + 7) / 8);
    const int num_lanes = std::max(1, std::min(g_render_lanes, bands / 4));
    RB_CUDA_CHECK(cudaEventCreate(&r->ev0));
    RB_CUDA_CHECK(cudaEventCreate(&r->ev1));
    if (num_lanes == 1) {
        alloc_pipeline(r);
    } else {
        RB_CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));      // film copies, timing events
        for (int j = 0; j < num_lanes; j++) {
            auto lane = new Renderer();
            lane->parent = r;
            lane->dev = dev; lane->width = width; lane->height = height; lane->spp = spp; lane->max_path_len = max_path_len;
            lane->capacity = r->capacity;
            lane->sm_count = r->sm_count; lane->occ_primary = r->occ_primary; lane->occ_primary2 = r->occ_primary2; lane->occ_shadow2 = r->occ_shadow2; lane->occ_shadow = r->occ_shadow;
            lane->scene = r->scene;
            for (size_t k = 0; k < r->rows.size(); k++)
                if (int(k / 8) % num_lanes == j) lane->rows.push_back(r->rows[k]);
            alloc_pipeline(lane);
            r->lanes.push_back(lane);
        }
    }
    return r;
}

static void destroy_renderer(Renderer* r) {
    if (!r) return;
    for (ncclComm_t c : r->comms) RB_NCCL_CHECK(Nccl::get().CommDestroy(c));
    for (Renderer* peer : r->peers) destroy_renderer(peer);
    RB_CUDA_CHECK(cudaSetDevice(r->dev));
    for (Renderer* lane : r->lanes) destroy_renderer(lane);
    if (r->stream) RB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    if (r->stream2) RB_CUDA_CHECK(cudaStreamSynchronize(r->stream2));
    for (void* p : r->allocations) RB_CUDA_CHECK(cudaFree(p));
    if (r->h_film) RB_CUDA_CHECK(cudaFreeHost(r->h_film));
    if (r->h_state) RB_CUDA_CHECK(cudaFreeHost(r->h_state));
    for (cudaEvent_t e : {r->ev0, r->ev1, r->ev_shaded, r->ev_shadow_done, r->ev_done}) if (e) RB_CUDA_CHECK(cudaEventDestroy(e));
    for (cudaEvent_t e : r->ev_state) if (e) RB_CUDA_CHECK(cudaEventDestroy(e));
    if (r->stream) RB_CUDA_CHECK(cudaStreamDestroy(r->stream));
    if (r->stream2) RB_CUDA_CHECK(cudaStreamDestroy(r->stream2));
    delete r;
}

// ---- gpu_streaming_trace, mapping_gpu.impala:308-369, driven without waiting for the device ----
// The reference's loop needs `size` (the survivors) on the host before it can launch the next wavefront.  Here the
// loop variables live on the device (LoopState, plan_wavefront) and every kernel reads its counts from there, so the
// host can enqueue wavefront k + 1 while wavefront k is still running.  What the host cannot know is when the loop ends:
// each wavefront copies the loop state to pinned memory right after its plan step, and the host reads the snapshot of
// wavefront k - kLookahead before it enqueues wavefront k.  Once a snapshot says `done`, the wavefronts already in the
// queue find an empty stream and fall through.  Grids: the persistent traversal kernels have theirs; generate / scatter /
// shade get a grid for `bound` rays, the full stream until a snapshot shows that all camera rays are out, after which
// the stream can only shrink and the last seen size bounds it.

static void begin_pipeline(Renderer& r, cudaEvent_t start, int64_t total) {
    r.head = r.tail = 0; r.n_kernels = 0; r.finished = false; r.generation_over = false; r.bound = r.capacity;
    LoopState init{};
    init.total = total;
    r.h_state[kLookahead] = init;                         // (pinned scratch for the upload)
    RB_CUDA_CHECK(cudaStreamWaitEvent(r.stream, start, 0));
    RB_CUDA_CHECK(cudaMemcpyAsync(r.state, &r.h_state[kLookahead], sizeof(LoopState), cudaMemcpyHostToDevice, r.stream));
}

template <bool FMA>
static void enqueue_wavefront(Renderer& r, float* film, const CameraDev& cam, int iter) {
    const int parity = int(r.head & 1);
    int* counters = r.counters + parity * kNumCounters;
    const int* prev = r.head ? r.counters + (parity ^ 1) * kNumCounters : nullptr;
    const float inv_spp = 1.0f / float(r.spp);
    const int num_geoms = r.scene.num_materials;
    PrimaryStream& P = r.prim[0];
    PrimaryStream& Q = r.prim[1];
    const int cap = r.capacity, bound = r.bound;
    // Two streams: everything up to the shade kernel runs on `s`; the shadow-ray traversal of a wavefront runs on
    // `s2` and overlaps the generation and closest-hit traversal of the NEXT wavefront.  Both traversals are
    // persistent kernels whose last rays straggle: the CTAs that drain early make room for the other kernel.  The two
    // kernels only share the film (atomics); counters alternate between two sets, and the shade kernel of wavefront
    // k waits for the shadow pass of wavefront k - 1, so a set is never reset under a pass that still reads it.
    cudaStream_t s = r.stream, s2 = r.stream2;
    plan_wavefront<<<1, 32, 0, s>>>(r.state, prev, counters, cap);
    const int slot = int(r.head % kLookahead);
    RB_CUDA_CHECK(cudaMemcpyAsync(&r.h_state[slot], r.state, sizeof(LoopState), cudaMemcpyDeviceToHost, s));
    RB_CUDA_CHECK(cudaEventRecord(r.ev_state[slot], s));
    if (!r.generation_over)
        generate_rays<<<(cap + 255) / 256, 256, 0, s>>>(P, r.state, cam, r.width, r.height, r.spp, iter, r.d_rows);
    if (r.scene.nodes2) {
        const int grid_p = std::min((bound + kRBlock - 1) / kRBlock, r.sm_count * r.occ_primary2);
        const bool quant = r.scene.nodes2q != nullptr;
        auto kernel = quant ? traverse_stream_bvh2<false, kBvh2Stack, FMA, true> : traverse_stream_bvh2<false, kBvh2Stack, FMA, false>;
        kernel<<<grid_p, kRBlock, (num_geoms + 1) * sizeof(int), s>>>(
            quant ? static_cast<const void*>(r.scene.nodes2q) : static_cast<const void*>(r.scene.nodes2), r.scene.grid, r.scene.tris1, P.ray_o, P.ray_d, &r.state->size, cap, P.hit, P.geom, num_geoms,
            r.histogram, nullptr, nullptr, nullptr, 0.0f, counters + kWorkPrimary, g_render_refill_min, g_render_streak_min, g_render_leaf_streak_min);
    } else {
        const int grid_p = std::min((bound + kRBlock - 1) / kRBlock, r.sm_count * r.occ_primary);
        auto kernel = g_render_wide ? traverse_stream<false, true, FMA> : traverse_stream<false, false, FMA>;
        kernel<<<grid_p, kRBlock, 0, s>>>(r.scene.nodes, r.scene.tris, P.ray_o, P.ray_d, &r.state->size, cap, P.hit, P.geom, num_geoms,
                                          r.histogram, nullptr, nullptr, nullptr, 0.0f, counters + kWorkPrimary, kRefillMin);
    }
    scan_bins<<<1, 32, 0, s>>>(r.histogram, r.cursor, num_geoms, counters);
    scatter_by_material<<<(bound + kScatterBlock - 1) / kScatterBlock, kScatterBlock, 2 * num_geoms * sizeof(int), s>>>(P, r.order, r.state, num_geoms, r.cursor);
    RB_CUDA_CHECK(cudaStreamWaitEvent(s, r.ev_shadow_done, 0));      // the previous shadow pass has read the shadow stream
    auto shade = g_render_shade_blocks >= 10 ? shade_rays<10> : g_render_shade_blocks >= 8 ? shade_rays<8> : shade_rays<6>;
    shade<<<(bound + 127) / 128, 128, 0, s>>>(P, r.order, Q, r.shadow, r.scene, counters, film, inv_spp, r.max_path_len);
    RB_CUDA_CHECK(cudaEventRecord(r.ev_shaded, s));
    std::swap(r.prim[0], r.prim[1]);                        // the survivors (in Q) are the next wavefront's stream
    RB_CUDA_CHECK(cudaStreamWaitEvent(s2, r.ev_shaded, 0));
    if (r.scene.nodes2 && g_render_shadow_bvh2) {
        const int grid_s = std::min((bound + kRBlock - 1) / kRBlock, r.sm_count * r.occ_shadow2);
        const bool quant = r.scene.nodes2q != nullptr;
        auto kernel = quant ? traverse_stream_bvh2<true, kBvh2Stack, FMA, true> : traverse_stream_bvh2<true, kBvh2Stack, FMA, false>;
        kernel<<<grid_s, kRBlock, sizeof(int), s2>>>(
            quant ? static_cast<const void*>(r.scene.nodes2q) : static_cast<const void*>(r.scene.nodes2), r.scene.grid, r.scene.tris1, r.shadow.ray_o, r.shadow.ray_d, counters + kShadows, cap,
            nullptr, nullptr, num_geoms, nullptr, r.shadow.pixel, r.shadow.color, film, inv_spp,
            counters + kWorkShadow, g_render_refill_min, g_render_streak_min, g_render_leaf_streak_min);
    } else {
        const int grid_s = std::min((bound + kRBlock - 1) / kRBlock, r.sm_count * r.occ_shadow);
        auto kernel = g_render_wide ? traverse_stream<true, true, FMA> : traverse_stream<true, false, FMA>;
        kernel<<<grid_s, kRBlock, 0, s2>>>(r.scene.nodes, r.scene.tris, r.shadow.ray_o, r.shadow.ray_d, counters + kShadows, cap,
                                           nullptr, nullptr, num_geoms, nullptr, r.shadow.pixel, r.shadow.color, film, inv_spp,
                                           counters + kWorkShadow, kRefillMin);
    }
    RB_CUDA_CHECK(cudaEventRecord(r.ev_shadow_done, s2));
    RB_CUDA_CHECK(cudaGetLastError());
    r.n_kernels += r.generation_over ? 6 : 7;
    r.head++;
}

// Reads the snapshots that have arrived (all of them up to `upto` when `wait`); returns true when one was read.
static bool retire_snapshots(Renderer& r, bool wait) {
    bool any = false;
    while (r.tail < r.head) {
        const int slot = int(r.tail % kLookahead);
        if (wait) RB_CUDA_CHECK(cudaEventSynchronize(r.ev_state[slot]));
        else {
            const cudaError_t q = cudaEventQuery(r.ev_state[slot]);
            if (q == cudaErrorNotReady) break;
            RB_CUDA_CHECK(q);
        }
        const LoopState& st = r.h_state[slot];
        if (st.done) { r.finished = true; r.h_state[kLookahead] = st; }
        if (st.next_id >= st.total) { r.generation_over = true; r.bound = std::max(st.size, 1); }
        r.tail++;
        any = true;
        wait = false;                                           // one blocking wait per call is enough
    }
    return any;
}

// One render(settings, iter) call.  With lanes, the pipelines' kernels interleave on the device, so the stragglers at
// the end of one pipeline's traversal kernels are covered by the other pipelines' work; one host thread feeds them all.
static void render_device(Renderer& r, const Settings& st, int iter) {
    RB_CUDA_CHECK(cudaSetDevice(r.dev));
    const CameraDev cam{{st.eye.x, st.eye.y, st.eye.z}, {st.dir.x, st.dir.y, st.dir.z}, {st.up.x, st.up.y, st.up.z},
                        {st.right.x, st.right.y, st.right.z}, st.width, st.height};
    std::vector<Renderer*> pipes = r.lanes.empty() ? std::vector<Renderer*>{&r} : r.lanes;
    RB_CUDA_CHECK(cudaEventRecord(r.ev0, r.stream));
    for (Renderer* p : pipes) begin_pipeline(*p, r.ev0, int64_t(p->spp) * p->width * int64_t(p->rows.size()));
    for (;;) {
        bool all_finished = true, progressed = false;
        for (Renderer* p : pipes) {
            if (p->finished) continue;
            progressed |= retire_snapshots(*p, false);
            if (p->finished) continue;
            all_finished = false;
            if (p->head - p->tail < kLookahead) {
                if (g_render_fma) enqueue_wavefront<true>(*p, r.film, cam, iter); else enqueue_wavefront<false>(*p, r.film, cam, iter);
                progressed = true;
            }
        }
        if (all_finished) break;
        if (!progressed) {                                      // every queue is full: wait for the oldest snapshot of one of them
            Renderer* oldest = nullptr;
            for (Renderer* p : pipes) if (!p->finished && (!oldest || p->tail < oldest->tail)) oldest = p;
            retire_snapshots(*oldest, true);
        }
    }
    // the wavefronts still in the queues are empty; the render ends when the last shadow pass has added its light
    for (Renderer* p : pipes) {
        RB_CUDA_CHECK(cudaStreamWaitEvent(p->stream, p->ev_shadow_done, 0));
        if (p != &r) {
            RB_CUDA_CHECK(cudaEventRecord(p->ev_done, p->stream));
            RB_CUDA_CHECK(cudaStreamWaitEvent(r.stream, p->ev_done, 0));
        }
    }
    RB_CUDA_CHECK(cudaEventRecord(r.ev1, r.stream));
    RB_CUDA_CHECK(cudaEventSynchronize(r.ev1));
    float ms = 0.0f;
    RB_CUDA_CHECK(cudaEventElapsedTime(&ms, r.ev0, r.ev1));
    r.last_ms = ms;
    for (int k = 0; k < 5; k++) r.stats[k] = 0;
    for (Renderer* p : pipes) {
        const LoopState& fin = p->h_state[kLookahead];
        r.stats[0] += fin.total; r.stats[1] += fin.n_primary; r.stats[2] += fin.n_shadow;
        r.stats[3] = std::max<int64_t>(r.stats[3], fin.n_waves);           // wavefronts: the longest pipeline
        r.stats[4] += p->n_kernels;
        p->tail = p->head;                                                 // (snapshots of the empty wavefronts are not read)
    }
    rodent_b200_count_launches(r.stats[4]);
}

// One render call on several devices of this process: every device renders its row bands (own host thread, own wavefront
// loops), then ONE ncclReduce sums the films onto the first device -- the only communication, as in the one-process-per-GPU
// form (rodent_b200/sharding.py).  The peers' films are cleared after the reduce: what they held is now part of the
// first device's film, which keeps accumulating over iterations like a single renderer's.
static void render_multi(Renderer& r, const Settings& st, int iter) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (Renderer* p : r.peers) threads.emplace_back([p, &st, iter] { render_device(*p, st, iter); });
    render_device(r, st, iter);
    for (auto& t : threads) t.join();
    const size_t count = size_t(r.width) * r.height * 3;
    Nccl& nccl = Nccl::get();
    RB_NCCL_CHECK(nccl.GroupStart());
    for (size_t k = 0; k < r.comms.size(); k++) {
        Renderer& p = k == 0 ? r : *r.peers[k - 1];
        RB_CUDA_CHECK(cudaSetDevice(p.dev));
        RB_NCCL_CHECK(nccl.Reduce(p.film, p.film, count, ncclFloat, ncclSum, 0, r.comms[k], p.stream));
    }
    RB_NCCL_CHECK(nccl.GroupEnd());
    for (Renderer* p : r.peers) {
        RB_CUDA_CHECK(cudaSetDevice(p->dev));
        RB_CUDA_CHECK(cudaMemsetAsync(p->film, 0, count * sizeof(float), p->stream));
        RB_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    }
    RB_CUDA_CHECK(cudaSetDevice(r.dev));
    RB_CUDA_CHECK(cudaStreamSynchronize(r.stream));
    r.last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (Renderer* p : r.peers) {
        for (int k : {0, 1, 2, 4}) r.stats[k] += p->stats[k];
        r.stats[3] = std::max(r.stats[3], p->stats[3]);
    }
}

static void render_any(Renderer& r, const Settings& st, int iter) {
    if (r.peers.empty()) render_device(r, st, iter); else render_multi(r, st, iter);
}

static void present(Renderer& r) {
    RB_CUDA_CHECK(cudaSetDevice(r.dev));
    RB_CUDA_CHECK(cudaMemcpyAsync(r.h_film, r.film, size_t(r.width) * r.height * 3 * sizeof(float), cudaMemcpyDeviceToHost, r.stream));
    RB_CUDA_CHECK(cudaStreamSynchronize(r.stream));
}

// state behind the reference driver's global entry points
static const Scene* g_bound_scene = nullptr;
static int g_bound_dev = 0, g_bound_spp = 4, g_bound_max_path_len = 64;
static std::vector<int> g_bound_devs;          // rodent_b200_bind_multi: setup_interface then spreads the film over these devices
static Renderer* g_current = nullptr;

}  // namespace rb200

using namespace rb200;

extern "C" {

RodentRenderer* rodent_b200_renderer_create(const RodentScene* scene, int32_t dev, int32_t width, int32_t height, int32_t spp,
                                            int32_t max_path_len, int32_t part, int32_t num_parts, int32_t band) {
    return reinterpret_cast<RodentRenderer*>(create_renderer(*reinterpret_cast<const Scene*>(scene), dev, width, height, spp, max_path_len, part, num_parts, band));
}
void rodent_b200_renderer_free(RodentRenderer* r) { destroy_renderer(reinterpret_cast<Renderer*>(r)); }
RodentRenderer* rodent_b200_renderer_create_multi(const RodentScene* scene, const int32_t* devs, int32_t num_devs, int32_t width, int32_t height,
                                                  int32_t spp, int32_t max_path_len, int32_t band) {
    if (!devs || num_devs <= 0) return nullptr;
    const Scene& sc = *reinterpret_cast<const Scene*>(scene);
    Renderer* root = create_renderer(sc, devs[0], width, height, spp, max_path_len, 0, num_devs, band);
    if (!root || num_devs == 1) return reinterpret_cast<RodentRenderer*>(root);
    for (int k = 1; k < num_devs; k++) {
        Renderer* p = create_renderer(sc, devs[k], width, height, spp, max_path_len, k, num_devs, band);
        if (!p) { destroy_renderer(root); return nullptr; }
        root->peers.push_back(p);
    }
    root->comms.resize(num_devs);
    std::vector<int> d(devs, devs + num_devs);
    RB_NCCL_CHECK(Nccl::get().CommInitAll(root->comms.data(), num_devs, d.data()));
    RB_CUDA_CHECK(cudaSetDevice(devs[0]));
    return reinterpret_cast<RodentRenderer*>(root);
}
void rodent_b200_render(RodentRenderer* r, const Settings* settings, int32_t iter) {
    render_any(*reinterpret_cast<Renderer*>(r), *settings, iter);
    present(*reinterpret_cast<Renderer*>(r));
}
void rodent_b200_render_device(RodentRenderer* r, const Settings* settings, int32_t iter) { render_any(*reinterpret_cast<Renderer*>(r), *settings, iter); }
void rodent_b200_present(RodentRenderer* r) { present(*reinterpret_cast<Renderer*>(r)); }
float* rodent_b200_film(RodentRenderer* r) { return reinterpret_cast<Renderer*>(r)->h_film; }
void* rodent_b200_film_device(RodentRenderer* r) { return reinterpret_cast<Renderer*>(r)->film; }
void rodent_b200_renderer_bind_film(RodentRenderer* rr, float* device_film) {
    Renderer& r = *reinterpret_cast<Renderer*>(rr);
    RB_CUDA_CHECK(cudaSetDevice(r.dev));
    RB_CUDA_CHECK(cudaStreamSynchronize(r.stream));
    r.film = device_film ? device_film : r.own_film;
}
void rodent_b200_clear(RodentRenderer* rr) {
    Renderer& r = *reinterpret_cast<Renderer*>(rr);
    for (Renderer* p : r.peers) {
        RB_CUDA_CHECK(cudaSetDevice(p->dev));
        RB_CUDA_CHECK(cudaMemsetAsync(p->film, 0, size_t(p->width) * p->height * 3 * sizeof(float), p->stream));
        RB_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    }
    RB_CUDA_CHECK(cudaSetDevice(r.dev));
    RB_CUDA_CHECK(cudaMemsetAsync(r.film, 0, size_t(r.width) * r.height * 3 * sizeof(float), r.stream));
    RB_CUDA_CHECK(cudaStreamSynchronize(r.stream));
    std::memset(r.h_film, 0, size_t(r.width) * r.height * 3 * sizeof(float));
}
void rodent_b200_render_stats(const RodentRenderer* r, int64_t out[5]) { std::memcpy(out, reinterpret_cast<const Renderer*>(r)->stats, 5 * sizeof(int64_t)); }
double rodent_b200_render_last_ms(const RodentRenderer* r) { return reinterpret_cast<const Renderer*>(r)->last_ms; }

// Thresholds are clamped to [1, 33] (below 1 a streak loop would never end, and the refill leader would be lane -1);
// the stream capacity is read when a renderer is created and kept by it, so changing it later cannot outgrow buffers.
void rodent_b200_render_tune(const char* key, int32_t value) {
    auto clamp = [](int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); };
    if (!std::strcmp(key, "render_lanes")) g_render_lanes = clamp(value, 1, 16);
    else if (!std::strcmp(key, "render_bvh2")) g_render_bvh2 = value != 0;
    else if (!std::strcmp(key, "render_shadow_bvh2")) g_render_shadow_bvh2 = value != 0;
    else if (!std::strcmp(key, "render_wide")) g_render_wide = value != 0;
    else if (!std::strcmp(key, "render_refill_min")) g_render_refill_min = clamp(value, 1, 32);
    else if (!std::strcmp(key, "render_bvh2_stack")) {}        // fixed at 16 since round 2 (8 .. 32 measured the same)
    else if (!std::strcmp(key, "render_fma")) g_render_fma = value != 0;
    else if (!std::strcmp(key, "render_quant")) g_render_quant = value != 0;
    else if (!std::strcmp(key, "render_shade_blocks")) g_render_shade_blocks = clamp(value, 6, 10);
    else if (!std::strcmp(key, "render_poly_trig")) g_render_poly_trig = value != 0;
    else if (!std::strcmp(key, "render_capacity")) g_capacity = clamp(value, 1024, 1 << 24);
    else if (!std::strcmp(key, "render_streak_min")) g_render_streak_min = clamp(value, 1, 33);
    else if (!std::strcmp(key, "render_leaf_streak_min")) g_render_leaf_streak_min = clamp(value, 0, 33);
    else { std::fprintf(stderr, "rodent_b200_tune: unknown key '%s'\n", key); std::abort(); }
}
void rodent_b200_bind(const RodentScene* scene, int32_t dev, int32_t spp, int32_t max_path_len) {
    g_bound_scene = reinterpret_cast<const Scene*>(scene); g_bound_dev = dev; g_bound_spp = spp; g_bound_max_path_len = max_path_len;
    g_bound_devs.clear();
}
void rodent_b200_bind_multi(const RodentScene* scene, const int32_t* devs, int32_t num_devs, int32_t spp, int32_t max_path_len) {
    rodent_b200_bind(scene, num_devs > 0 ? devs[0] : 0, spp, max_path_len);
    g_bound_devs.assign(devs, devs + std::max(num_devs, 0));
}
void setup_interface(size_t width, size_t height) {
    if (!g_bound_scene) { std::fprintf(stderr, "rodent_b200: setup_interface() without rodent_b200_bind()\n"); std::abort(); }
    destroy_renderer(g_current);
    if (g_bound_devs.size() > 1)
        g_current = reinterpret_cast<Renderer*>(rodent_b200_renderer_create_multi(reinterpret_cast<const RodentScene*>(g_bound_scene), g_bound_devs.data(),
                                                                                   int32_t(g_bound_devs.size()), int32_t(width), int32_t(height),
                                                                                   g_bound_spp, g_bound_max_path_len, 8));
    else
        g_current = create_renderer(*g_bound_scene, g_bound_dev, int(width), int(height), g_bound_spp, g_bound_max_path_len, 0, 1, 1);
    if (!g_current) std::abort();
}
void cleanup_interface(void) { destroy_renderer(g_current); g_current = nullptr; }
float* get_pixels(void) { return g_current ? g_current->h_film : nullptr; }
void clear_pixels(void) { if (g_current) rodent_b200_clear(reinterpret_cast<RodentRenderer*>(g_current)); }
int32_t get_spp(void) { return g_bound_spp; }
void render(const Settings* settings, int32_t iter) {
    if (!g_current) { std::fprintf(stderr, "rodent_b200: render() before setup_interface()\n"); std::abort(); }
    rodent_b200_render(reinterpret_cast<RodentRenderer*>(g_current), settings, iter);
}

}  // extern "C"
