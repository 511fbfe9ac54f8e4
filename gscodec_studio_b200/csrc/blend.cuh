// Shared device helpers of the blend kernels (blend.cu: generic per-tile kernels; blend_stream.cu: the
// producer/consumer kernels for 16x16 tiles).  See blend.cu for the reference citations.
#pragma once
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace gsb200 {

constexpr int kBatch = 256;
// minimum resident CTAs per SM the register allocator must allow (occupancy knobs)
#ifndef GSB_STAGE_PER_THREAD
#define GSB_STAGE_PER_THREAD 2  // Gaussians staged per thread and barrier (batch = threads x this)
#endif
constexpr int kStagePerThread = GSB_STAGE_PER_THREAD;
#ifndef GSB_FWD_BLOCKS
#define GSB_FWD_BLOCKS 6
#endif
#ifndef GSB_BWD_BLOCKS
#define GSB_BWD_BLOCKS 5
#endif
constexpr float kAlphaMax = 0.999f;
constexpr float kAlphaMin = 1.f / 255.f;
constexpr float kTStop = 1e-4f;  // reference compares against the double 1e-4; identical set of floats

struct BlendGeom {
    uint32_t C, W, H, ts, tw, th;
    int64_t n_isects;
};

// pixel owned by this thread inside its tile: 8x4 warp blocks for 16x16 tiles, row-major otherwise
__device__ __forceinline__ void pixel_in_tile(uint32_t ts, uint32_t tid, uint32_t &lx, uint32_t &ly) {
    if (ts == 16) {
        const uint32_t w = tid >> 5, l = tid & 31;
        lx = (w & 1) * 8 + (l & 7);
        ly = (w >> 1) * 4 + (l >> 3);
    } else {
        lx = tid % ts;
        ly = tid / ts;  // >= ts for the padding threads of a block rounded up to a full warp
    }
}

// Cull parameter of one Gaussian: q_max = 2 ln(255 * opac) (+ margin).  A pixel at offset d from the
// centre can pass the reference's alpha >= 1/255 test only if d^T A d <= q_max.  FLT_MAX (always
// evaluate) when the conic is not positive definite, negative (never) when no pixel can pass.
__device__ __forceinline__ float cull_qmax(float a, float b, float c, float opac) {
    const float tau = __logf(255.f * opac);
    if (!(opac > 0.f) || !(tau >= 0.f)) return (opac * 255.f >= 1.f) ? FLT_MAX : -1.f;  // first case: NaN / inf guard
    if (!(a * c - b * b > 0.f) || !(a > 0.f) || !(c > 0.f)) return FLT_MAX;
    return 2.f * (tau * 1.001f + 1e-3f);
}

// Exact test "does the ellipse { d : d^T A d <= qmax } around (mx,my) meet the pixel-centre rectangle
// [x0,x1] x [y0,y1]?" -- the minimum of the convex form over the rectangle is 0 if the centre is inside,
// otherwise it is attained on one of the four edges (a clamped 1-D parabola each).  A margin proportional
// to the magnitude of the terms keeps the test conservative under float32 rounding, so it never drops a
// (pixel, Gaussian) pair the reference arithmetic would accept.
__device__ __forceinline__ bool ellipse_hits_rect(float mx, float my, float a, float b, float c, float qmax, float x0,
                                                  float x1, float y0, float y1) {
    if (!(qmax >= 0.f)) return false;
    if (qmax == FLT_MAX) return true;
    const float lx0 = x0 - mx, lx1 = x1 - mx, ly0 = y0 - my, ly1 = y1 - my;
    if (lx0 <= 0.f && lx1 >= 0.f && ly0 <= 0.f && ly1 >= 0.f) return true;
    const float ia = __frcp_rn(a), ic = __frcp_rn(c);
    float q = FLT_MAX;
    {
        float dy = fminf(fmaxf(-b * lx0 * ic, ly0), ly1);
        q = fminf(q, a * lx0 * lx0 + 2.f * b * lx0 * dy + c * dy * dy);
        dy = fminf(fmaxf(-b * lx1 * ic, ly0), ly1);
        q = fminf(q, a * lx1 * lx1 + 2.f * b * lx1 * dy + c * dy * dy);
        float dx = fminf(fmaxf(-b * ly0 * ia, lx0), lx1);
        q = fminf(q, a * dx * dx + 2.f * b * dx * ly0 + c * ly0 * ly0);
        dx = fminf(fmaxf(-b * ly1 * ia, lx0), lx1);
        q = fminf(q, a * dx * dx + 2.f * b * dx * ly1 + c * ly1 * ly1);
    }
    const float ex = fmaxf(fabsf(lx0), fabsf(lx1)), ey = fmaxf(fabsf(ly0), fabsf(ly1));
    const float slack = 1e-5f * (a * ex * ex + 2.f * fabsf(b) * ex * ey + c * ey * ey) + 1e-4f;
    return q <= qmax + slack;
}

// shared-memory accesses through 32-bit shared-window addresses (keeps the address arithmetic out of
// the inner loop; generic pointers make the compiler rebuild the window base every iteration)
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds1i(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int CDIM> struct ColorPad { static constexpr int value = (CDIM + 3) / 4 * 4; };

// ---- warp reductions that leave each of NV totals in a distinct lane group --------------------
// After reduce_spread<NV>(v): the total of value `slot` is held in v[0] by every lane whose
// (lane >> shift) & (NV-1), bit-reversed over log2(NV) bits, equals slot -- see spread_slot().
template <int NV>
__device__ __forceinline__ void reduce_spread(float *v, uint32_t lane) {
    // halving steps: partner distance 16, 8, ... while more than one value is alive
    int alive = NV;
    int dist = 16;
#pragma unroll
    for (; alive > 1; alive >>= 1, dist >>= 1) {
        const bool upper = (lane & dist) != 0;
        const int half = alive >> 1;
#pragma unroll
        for (int q = 0; q < half; ++q) {
            const float keep = upper ? v[half + q] : v[q];
            const float send = upper ? v[q] : v[half + q];
            v[q] = keep + __shfl_xor_sync(0xffffffffu, send, dist);
        }
    }
#pragma unroll
    for (; dist > 0; dist >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], dist);
}
// value slot held by `lane` after reduce_spread<NV>
template <int NV>
__device__ __forceinline__ uint32_t spread_slot(uint32_t lane) {
    uint32_t slot = 0, alive = NV, dist = 16;
    for (; alive > 1; alive >>= 1, dist >>= 1) slot += (lane & dist) ? (alive >> 1) : 0;
    return slot;
}

// packed gradient record (floats): [rgb(CDIM) | conic(3) | xy(2) | opac(1) | abs(2)] rounded up to 4
__host__ __device__ constexpr uint32_t grad_record(uint32_t cdim) { return (cdim + 8 + 3) / 4 * 4; }


int launch_tile_order(uint32_t total_tiles, int64_t n_isects, const int32_t *offsets, int32_t *order, cudaStream_t s);

// ---- blend_stream.cu entry points (16x16 tiles, <= 4 colour channels) ------------------------------------
bool blend_stream_supported(uint32_t tile_size, uint32_t cdim, int64_t n_isects);
// true: the stream backward leaves [rgb | sum w dx^2, sum w dx dy, sum w dy^2, sum w dx, sum w dy | sum w] (w = -v_sigma) in the
// gradient record instead of [rgb | v_conic | v_xy | v_opacity]; unpack_grads_kernel applies the per-Gaussian linear map
bool blend_stream_bwd_moments(bool absgrad);
int blend_stream_fwd(const BlendGeom &G, uint32_t cdim, const float *means2d, const float *conics, const float *colors,
                     const float *opacities, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                     const int32_t *flatten_ids, float *render_colors, float *render_alphas, int32_t *last_ids,
                     int32_t *tile_order, float *blend_stream, int32_t *tile_counts, cudaStream_t s);
int blend_stream_bwd(const BlendGeom &G, uint32_t cdim, uint32_t M, const float *backgrounds, const uint8_t *masks,
                     const int32_t *tile_offsets, const float *render_alphas, const int32_t *last_ids,
                     const float *v_render_colors, const float *v_render_alphas, const int32_t *tile_order,
                     const float *blend_stream, const int32_t *tile_counts, float *grad_scratch, bool absgrad, cudaStream_t s);

}  // namespace gsb200
