// Microbenchmarks that size the traversal kernel's design on B200 (sm_100a):
//  (1) issue throughput of the fp32 / integer-minmax instructions the slab test is made of,
//      scalar vs packed f32x2;
//  (2) L1/L2 throughput of node fetches: per-lane divergent LDG.128 (thread-per-ray),
//      8-lanes-per-line cooperative fetch, and the broadcast (coherent) case.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, float a, float b) {
    float x[8];
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3f + i; p[i] = make_float2(x[i], x[i] + 0.5f); }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    int n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) n[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) x[i] = __fmul_rn(x[i], a);
            if (OP == 1) x[i] = __fadd_rn(x[i], b);
            if (OP == 2) x[i] = __fmaf_rn(x[i], a, b);
            if (OP == 3) p[i] = __fmul2_rn(p[i], a2);
            if (OP == 4) p[i] = __fadd2_rn(p[i], b2);
            if (OP == 5) p[i] = __ffma2_rn(p[i], a2, b2);
            if (OP == 6) { x[i] = __fmul_rn(x[i], a); x[(i + 4) & 7] = __fadd_rn(x[(i + 4) & 7], b); }       // scalar mul+add, independent
            if (OP == 7) { p[i] = __fmul2_rn(p[i], a2); p[(i + 4) & 7] = __fadd2_rn(p[(i + 4) & 7], b2); }   // packed mul+add, independent
            if (OP == 8) n[i] = __vimax3_s32(n[i], n[(i + 1) & 7], it);
            if (OP == 9) n[i] = max(n[i], it + i);
            if (OP == 10) n[i] = n[i] * 3 + it;                                                              // IMAD
            if (OP == 11) { x[i] = __fmul_rn(x[i], a); n[i] = max(n[i], it + i); }                           // fma pipe + alu pipe
        }
    }
    float s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i] + p[i].x + p[i].y; t += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

// ---- node fetch patterns ------------------------------------------------------------------
__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: every lane fetches its own node with 14 LDG.128 (thread-per-ray)
// MODE 1: same but all lanes of a warp use the same node (coherent rays, broadcast)
// MODE 2: cooperative: 8 lanes fetch one 128-byte line; 16 LDG.128 per warp bring 32 nodes (2 lines each) into shared memory,
//         then every lane reads its node back with 14 LDS.128 (row stride padded to 272 bytes)
// MODE 3: as 0 but lanes of a quad share a node (8 distinct nodes per warp)
template <int MODE>
__global__ void __launch_bounds__(256) fetch_kernel(const float4* __restrict__ nodes, int num_nodes, int iters, float* out) {
    extern __shared__ float4 stage_raw[];
    float4 (*stage)[32 * 17] = reinterpret_cast<float4 (*)[32 * 17]>(stage_raw);
    unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned r = lcg(seed) % unsigned(num_nodes);
        if (MODE == 1) r = __shfl_sync(0xffffffffu, r, 0);
        if (MODE == 3) r = __shfl_sync(0xffffffffu, r, lane & ~3u);
        if (MODE == 2) {
            float4 v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const unsigned src = j * 2 + (lane >> 4);                 // ray whose node this lane helps to fetch
                const unsigned node = __shfl_sync(0xffffffffu, r, src);
                v[j] = __ldg(nodes + node * 16 + (lane & 15));
            }
#pragma unroll
            for (int j = 0; j < 16; j++) stage[warp][(j * 2 + (lane >> 4)) * 17 + (lane & 15)] = v[j];
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 14; c++) { const float4 q = stage[warp][lane * 17 + c]; acc += q.x + q.y + q.z + q.w; }
            __syncwarp();
        } else {
            const float4* nb = nodes + r * 16;
            float4 v[14];
#pragma unroll
            for (int c = 0; c < 14; c++) v[c] = __ldg(nb + c);
#pragma unroll
            for (int c = 0; c < 14; c++) acc += v[c].x + v[c].y + v[c].z + v[c].w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("%s, %d SMs, %d kHz\n", prop.name, sms, clk_khz);
    float* out; CK(cudaMalloc(&out, size_t(sms) * 8 * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[] = {"FMUL", "FADD", "FFMA", "FMUL2", "FADD2", "FFMA2", "FMUL+FADD", "FMUL2+FADD2", "VIMNMX3", "IMNMX", "IMAD", "FMUL+IMNMX"};
    const int per_iter[] = {8, 8, 8, 8, 8, 8, 16, 16, 8, 8, 8, 16};
    auto run_pipe = [&](int op, int ctas_per_sm) {
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            switch (op) {
#define C(k) case k: pipe_kernel<k><<<sms * ctas_per_sm, 256>>>(out, 1.0000001f, 1e-7f); break;
                C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11)
#undef C
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double warp_inst = double(sms) * ctas_per_sm * 8 * ITERS * per_iter[op];
        const double cycles = ms * 1e-3 * clk_khz * 1e3;
        printf("%-12s ctas/sm %d: %.3f ms, %.2f warp-inst/cycle/SM\n", names[op], ctas_per_sm, ms, warp_inst / cycles / sms);
    };
    for (int op = 0; op < 12; op++) run_pipe(op, 4);

    const int num_nodes = 15054;
    CK(cudaFuncSetAttribute(fetch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 17 * 16));
    std::vector<float> h(size_t(num_nodes) * 64, 1.0f);
    float4* nodes; CK(cudaMalloc(&nodes, h.size() * 4)); CK(cudaMemcpy(nodes, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const char* fnames[] = {"divergent 14xLDG.128/lane", "broadcast (one node per warp)", "cooperative 8 lanes/line + smem", "quad-shared (8 nodes per warp)"};
    for (int nn : {num_nodes, 256}) {
        for (int mode = 0; mode < 4; mode++) {
            const int iters = 2000, ctas = sms * 4;
            float ms = 0;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0));
                if (mode == 0) fetch_kernel<0><<<ctas, 256>>>(nodes, nn, iters, out);
                if (mode == 1) fetch_kernel<1><<<ctas, 256>>>(nodes, nn, iters, out);
                if (mode == 2) fetch_kernel<2><<<ctas, 256, 8 * 32 * 17 * 16>>>(nodes, nn, iters, out);
                if (mode == 3) fetch_kernel<3><<<ctas, 256>>>(nodes, nn, iters, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            const double fetches = double(ctas) * 256 * iters;
            const double cycles = ms * 1e-3 * clk_khz * 1e3;
            printf("nodes %5d  %-34s: %.3f ms, %.1f Gnode-fetches/s, %.1f cycles per warp-fetch per SM\n", nn, fnames[mode], ms,
                   fetches / ms / 1e6, cycles / (fetches / 32 / sms));
        }
    }
    return 0;
}
