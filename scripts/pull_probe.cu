// Developer probe: how fast can SMs pull a pinned host buffer into device memory themselves (no copy engine), and push
// one back, against cudaMemcpyAsync.  nvcc -arch=sm_100a -O3 -o /tmp/pull_probe scripts/pull_probe.cu && /tmp/pull_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

template <int UNROLL>
__global__ void pull(const float4* __restrict__ src, float4* __restrict__ dst, size_t n16) {
    const size_t stride = size_t(gridDim.x) * blockDim.x * UNROLL;
    for (size_t base = size_t(blockIdx.x) * blockDim.x * UNROLL; base < n16; base += stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) { const size_t i = base + k * blockDim.x + threadIdx.x; if (i < n16) v[k] = __ldcv(src + i); }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) { const size_t i = base + k * blockDim.x + threadIdx.x; if (i < n16) dst[i] = v[k]; }
    }
}

int main() {
    const size_t bytes = 32u << 20, n16 = bytes / 16;
    float4 *h, *h2, *d;
    CK(cudaMallocHost(&h, bytes)); CK(cudaMallocHost(&h2, bytes)); CK(cudaMalloc(&d, bytes));
    for (size_t i = 0; i < n16; i++) h[i] = make_float4(float(i), 1, 2, 3);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](auto f) { float best = 1e9f; for (int r = 0; r < 6; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms; } return best; };
    float ms = time([&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); });
    std::printf("cudaMemcpyAsync H2D 32 MB: %.3f ms = %.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = time([&] { CK(cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost)); });
    std::printf("cudaMemcpyAsync D2H 32 MB: %.3f ms = %.1f GB/s\n", ms, bytes / ms / 1e6);
    for (int ctas : {37, 74, 148, 296, 592}) for (int threads : {64, 128, 256}) {
        ms = time([&] { pull<4><<<ctas, threads>>>(h, d, n16); });
        float ms8 = time([&] { pull<8><<<ctas, threads>>>(h, d, n16); });
        float push = time([&] { pull<4><<<ctas, threads>>>(d, h2, n16); });
        std::printf("SM pull, %3d CTAs x %3d threads: unroll 4 %.3f ms = %.1f GB/s, unroll 8 %.3f ms = %.1f GB/s;  SM push unroll 4 %.3f ms = %.1f GB/s\n",
                    ctas, threads, ms, bytes / ms / 1e6, ms8, bytes / ms8 / 1e6, push, bytes / push / 1e6);
    }
    // both directions at once
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    ms = time([&] { pull<4><<<148, 128, 0, s1>>>(h, d, n16); pull<4><<<148, 128, 0, s2>>>(d + n16 / 2, h2, n16 / 2); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
    std::printf("SM pull 32 MB + SM push 16 MB at once: %.3f ms\n", ms);
    ms = time([&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1)); pull<4><<<148, 128, 0, s2>>>(d + n16 / 2, h2, n16 / 2); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
    std::printf("copy-engine H2D 32 MB + SM push 16 MB at once: %.3f ms\n", ms);
    ms = time([&] { pull<4><<<148, 128, 0, s1>>>(h, d, n16); CK(cudaMemcpyAsync(h2, d + n16 / 2, bytes / 2, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
    std::printf("SM pull 32 MB + copy-engine D2H 16 MB at once: %.3f ms\n", ms);
    ms = time([&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(h2, d + n16 / 2, bytes / 2, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
    std::printf("copy-engine H2D 32 MB + copy-engine D2H 16 MB at once: %.3f ms\n", ms);
    for (int ctas : {8, 16, 37}) {
        ms = time([&] { pull<4><<<ctas, 64, 0, s1>>>(h, d, n16); pull<4><<<ctas, 64, 0, s2>>>(d + n16 / 2, h2, n16 / 2); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
        std::printf("SM pull 32 MB + SM push 16 MB at once, %d CTAs x 64 each: %.3f ms\n", ctas, ms);
    }
    CK(cudaMemcpy(h2, d, bytes, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n16; i += 4097) if (h2[i].x != h[i].x) { std::printf("MISMATCH at %zu\n", i); return 1; }
    std::printf("ok\n");
    return 0;
}
