// Shared helpers for the gsb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsb200.h"

namespace gsb200 {

extern thread_local char g_err[512];
extern unsigned long long g_launches;

inline int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

template <typename... Args>
inline int failf(int code, const char *fmt, Args... args) {
    snprintf(g_err, sizeof(g_err), fmt, args...);
    return code;
}

inline int check_launch(const char *what) {
    ++g_launches;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "gsb200: %s launch failed: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        return 100;
    }
    return 0;
}

#define GSB_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            snprintf(gsb200::g_err, sizeof(gsb200::g_err), "gsb200: %s: %s", #expr,                \
                     cudaGetErrorString(_e));                                                      \
            return 101;                                                                            \
        }                                                                                          \
    } while (0)

constexpr int kThreads = 256;

__host__ __device__ inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline uint32_t bit_width(uint32_t v) {  // floor(log2(v)) + 1 for v >= 1  (isect_tiles.cu:155-156)
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// 128-bit streaming loads / stores
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stg4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// vector reduction to global memory (sm_90+): one 16-byte red instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// predicated form: no branch around the reduction (the address is computed by every lane)
__device__ __forceinline__ void red_add_if(float *addr, float a, bool on) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p red.global.add.f32 [%0], %1;\n}" ::"l"(addr), "f"(a), "r"((uint32_t)on)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gsb200
