// Shared device/host helpers for the sm_100a kernels.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/rodent_b200.h"

// Unrecoverable errors abort, as the reference does (src/driver/common.h:43-46,
// tools/bench_aila/kepler_dynamic_fetch.cu:382-389).
#define RB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            std::fprintf(stderr, "rodent_b200: CUDA error '%s' at %s:%d (%s)\n",              \
                         cudaGetErrorString(err__), __FILE__, __LINE__, #expr);               \
            std::abort();                                                                     \
        }                                                                                     \
    } while (0)

namespace rb200 {

constexpr float kFltMax = 3.4028234664e+38f;   // src/core/common.impala:4

// ---- the arithmetic contract (DESIGN.md "Parity contract") ------------------
// IEEE fp32 round-to-nearest, never contracted into FMA: every product and sum of
// the reference's expressions is one __fmul_rn/__fadd_rn.  (The file is also
// compiled with -fmad=false; the intrinsics make the contract local and explicit.)
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
// a*b + c*d + e*f in the reference's left-to-right order (src/core/vector.impala:60)
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return add(add(mul(ax, bx), mul(ay, by)), mul(az, bz));
}

// The same expressions with the contraction switched on by a template flag.  FMA = true is used by the renderer's stream
// kernels only (rodent_b200_tune "render_fma"): their films are held to statistical parity, and half of a slab test's
// instructions go away; the bench_traversal entry points, whose records are compared bit for bit, never set it.
template <bool FMA> __device__ __forceinline__ float madd(float a, float b, float c) { return FMA ? __fmaf_rn(a, b, c) : add(mul(a, b), c); }
template <bool FMA> __device__ __forceinline__ float msub(float a, float b, float c, float d) {          // a*b - c*d
    return FMA ? __fmaf_rn(a, b, -mul(c, d)) : sub(mul(a, b), mul(c, d));
}
template <bool FMA> __device__ __forceinline__ float dot3f(float ax, float ay, float az, float bx, float by, float bz) {
    return FMA ? __fmaf_rn(az, bz, __fmaf_rn(ay, by, mul(ax, bx))) : dot3(ax, ay, az, bx, by, bz);
}

// src/core/common.impala:78-80
__device__ __forceinline__ float prodsign(float x, float y) {
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(y) & int(0x80000000u)));
}
// src/core/common.impala:82-85
__device__ __forceinline__ float safe_rcp(float x) {
    const float ax = x > 0.0f ? x : -x;
    return ax < 1e-8f ? prodsign(kFltMax, x) : __fdiv_rn(1.0f, x);
}
// Integer min/max of the float bits (src/traversal/mapping_cpu.impala:123-133);
// three-input forms map to one VIMNMX3 on sm_100a.
__device__ __forceinline__ float imin3(float a, float b, float c) {
    return __int_as_float(__vimin3_s32(__float_as_int(a), __float_as_int(b), __float_as_int(c)));
}
__device__ __forceinline__ float imax3(float a, float b, float c) {
    return __int_as_float(__vimax3_s32(__float_as_int(a), __float_as_int(b), __float_as_int(c)));
}
__device__ __forceinline__ float imin2(float a, float b) { return __int_as_float(min(__float_as_int(a), __float_as_int(b))); }
__device__ __forceinline__ float imax2(float a, float b) { return __int_as_float(max(__float_as_int(a), __float_as_int(b))); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ int4 ldg4(const int4* p) { return __ldg(p); }

// One 256-bit load (LDG.E.256 on sm_100a): a whole row of a Node8, half a Node2, two rows of a Tri4.  On the divergent
// record fetches of the traversal kernels a load instruction costs the L1 data pipe about one wavefront per lane whatever
// its width (profiles/r01_microbench_pipes.txt), so half as many instructions is half as many wavefronts.
// `p` must be 32-byte aligned.
struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg8(const float4* p) {
    F8 r;
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

}  // namespace rb200
