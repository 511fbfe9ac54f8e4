// Small device-wide exclusive scan (int32) used by the stream-compaction and tile-offset passes.
#pragma once
#include "common.cuh"

namespace gsb200 {

constexpr int kScanItems = 8;                      // items per thread
constexpr int kScanTile = kThreads * kScanItems;   // 2048 items per block

// Exclusive scan of `nb` values in place by ONE block; the grand total goes to data[nb].
static __global__ void __launch_bounds__(1024) scan_single_block_kernel(uint32_t nb, int32_t *__restrict__ data) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const int v = i < nb ? data[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int excl = s_carry + s_w[warp] + incl - v;
        if (i < nb) data[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[nb] = s_carry;
}

// per-block sums of kScanTile consecutive items
static __global__ void __launch_bounds__(kThreads) scan_block_sums_kernel(uint32_t n, const int32_t *__restrict__ in,
                                                                          int32_t *__restrict__ sums) {
    __shared__ int s_w[kThreads / 32];
    const uint32_t base = blockIdx.x * kScanTile;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const uint32_t i = base + k * kThreads + threadIdx.x;
        acc += i < n ? in[i] : 0;
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

// out[i] = block_offsets[block] + exclusive prefix of in[] inside the block; optional second copy
static __global__ void __launch_bounds__(kThreads)
scan_apply_kernel(uint32_t n, const int32_t *in, const int32_t *__restrict__ block_offsets, int32_t *out,
                  int32_t *out2) {  // `in` may alias `out`: each thread reads its items before writing them
    __shared__ int s_w[kThreads / 32];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;  // blocked arrangement
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) wbase += w < warp ? s_w[w] : 0;
    int run = block_offsets[blockIdx.x] + wbase + incl - sum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) {
            out[base + k] = run;
            if (out2 != nullptr) out2[base + k] = run;
        }
        run += v[k];
    }
}

}  // namespace gsb200
