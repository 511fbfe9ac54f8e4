// Stage 3 / 4a: per-tile front-to-back alpha blend, forward and backward (sm_100a).
//
// Replaces _C.rasterize_to_pixels_fwd / _bwd (gsplat/cuda/csrc/rasterize_to_pixels_fwd.cu:16-185,
// rasterize_to_pixels_bwd.cu:17-277).  Per-pixel arithmetic (pixel centre +0.5, sigma expression,
// __expf, min(0.999), 1/255 cut, exclusive 1e-4 transmittance stop, last_ids, T restored by division
// in the backward) is the reference's, so the same (pixel, Gaussian) pairs contribute.
//
// What is different:
//   * a 16x16 tile is owned by 8 warps of 8x4 pixels (the reference uses 16x2 strips); each warp
//     culls the staged Gaussians against ITS pixel rectangle 32 at a time -- lane l tests Gaussian l's
//     1/255-alpha ellipse bounding box, one ballot, and only survivors are evaluated.  The box is a
//     conservative superset (margin below), so per-pixel results are unchanged;
//   * warps stop independently as soon as all of their pixels are saturated (the reference only
//     stops when the whole tile is);
//   * colours are staged in shared memory with the geometry (the reference re-reads them from global
//     memory per pixel in the forward);
//   * backward: the 9+D per-Gaussian partial gradients of a warp are reduced with a halving butterfly
//     (8 values in 9 shuffles instead of 8x5) that leaves each total in a different lane, and those
//     lanes issue ONE coalesced red.global to a packed per-Gaussian gradient record (a single 32-byte
//     sector for D=3) instead of 9+D scalar atomics to four tensors; an unpack pass then writes the
//     reference's v_means2d / v_conics / v_colors / v_opacities layout;
//   * backward skips staging of batches that lie behind every pixel's last contributing Gaussian.
#include <stdlib.h>
#include <string.h>

#include "blend.cuh"

namespace gsb200 {


// BIG: tiles of 17..32 pixels, i.e. up to 1024 threads per CTA (the reference launches tile_size^2 threads for any size,
// rasterize_to_pixels_fwd.cu:228, 244-263); one Gaussian staged per thread there, so the buffers stay within 227 KB
template <int CDIM, bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : kBatch, (!BIG && CDIM <= 4) ? GSB_FWD_BLOCKS : 1)
blend_fwd_kernel(BlendGeom G, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                 const float *__restrict__ colors, const float *__restrict__ opacities,
                 const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks,
                 const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                 float *__restrict__ render_colors, float *__restrict__ render_alphas, int32_t *__restrict__ last_ids,
                 const int32_t *__restrict__ tile_order) {
    constexpr int CP = ColorPad<CDIM>::value;
    constexpr int SP = BIG ? 1 : kStagePerThread;
    extern __shared__ float4 smem4[];
    const uint32_t nthreads = blockDim.x, tid = threadIdx.x;
    const uint32_t batch = nthreads * SP;
    float4 *sA = smem4;                                 // x, y, opacity, qmax
    float4 *sB = sA + batch;                            // conic a, b, c, -
    float *sC = reinterpret_cast<float *>(sB + batch);  // [batch][CP]
    const uint32_t aA = (uint32_t)__cvta_generic_to_shared(sA), aB = (uint32_t)__cvta_generic_to_shared(sB),
                   aC = (uint32_t)__cvta_generic_to_shared(sC);

    // 1-D grid; with a work order the longest tiles are scheduled first (no straggler tail at the end)
    const uint32_t ct_lin = tile_order ? (uint32_t)tile_order[blockIdx.x] : blockIdx.x;
    const uint32_t cam = ct_lin / (G.tw * G.th), tile_id = ct_lin % (G.tw * G.th);
    const uint32_t tile_y = tile_id / G.tw, tile_x = tile_id % G.tw;
    uint32_t lx, ly;
    pixel_in_tile(G.ts, tid, lx, ly);
    const uint32_t i = tile_y * G.ts + ly, j = tile_x * G.ts + lx;
    const bool inside = ly < G.ts && i < G.H && j < G.W;
    const size_t pix = ((size_t)cam * G.H + i) * G.W + j;
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;

    if (masks != nullptr && !masks[ct]) {  // rasterize_to_pixels_fwd.cu:70-76
        if (inside)
#pragma unroll
            for (int k = 0; k < CDIM; ++k) render_colors[pix * CDIM + k] = bg ? bg[k] : 0.f;
        return;
    }
    const int32_t range_start = tile_offsets[ct];
    const int32_t range_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];

    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const unsigned wmask = 0xffffffffu;  // blocks are launched with whole warps
    const uint32_t lane = tid & 31;
    // pixel-centre rectangle of the warp (inside pixels only)
    const int jmin = __reduce_min_sync(wmask, inside ? (int)j : INT_MAX), jmax = __reduce_max_sync(wmask, inside ? (int)j : -1);
    const int imin = __reduce_min_sync(wmask, inside ? (int)i : INT_MAX), imax = __reduce_max_sync(wmask, inside ? (int)i : -1);
    const float wx0 = (float)jmin + 0.5f, wx1 = (float)jmax + 0.5f, wy0 = (float)imin + 0.5f, wy1 = (float)imax + 0.5f;

    float T = 1.f;
    uint32_t cur_idx = 0;
    float pix_out[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) pix_out[k] = 0.f;
    bool done = !inside;
    bool warp_done = jmax < 0;

    for (int32_t batch_start = range_start; batch_start < range_end; batch_start += (int32_t)batch) {
        // barrier protects the staging buffers; stop when every pixel of the tile is saturated
        if (__syncthreads_count(done) >= (int)nthreads) break;
#pragma unroll
        for (int sp = 0; sp < SP; ++sp) {
            const uint32_t e = tid + sp * nthreads;
            const int32_t idx = batch_start + (int32_t)e;
            if (idx < range_end) {
                const int32_t g = flatten_ids[idx];
                const float2 xy = __ldg(means2d + g);
                const float opac = __ldg(opacities + g);
                const float ca = __ldg(conics + (size_t)g * 3), cb = __ldg(conics + (size_t)g * 3 + 1),
                            cc = __ldg(conics + (size_t)g * 3 + 2);
                sA[e] = make_float4(xy.x, xy.y, opac, cull_qmax(ca, cb, cc, opac));
                sB[e] = make_float4(ca, cb, cc, 0.f);
#pragma unroll
                for (int k = 0; k < CDIM; ++k) sC[e * CP + k] = __ldg(colors + (size_t)g * CDIM + k);
            }
        }
        __syncthreads();
        if (warp_done) continue;
        const uint32_t bsize = min(batch, (uint32_t)(range_end - batch_start));
        for (uint32_t c0 = 0; c0 < bsize; c0 += 32) {
            const uint32_t e = c0 + lane;
            bool hit = false;
            if (e < bsize) {
                const float4 A = lds4(aA + e * 16), B = lds4(aB + e * 16);
                hit = ellipse_hits_rect(A.x, A.y, B.x, B.y, B.z, A.w, wx0, wx1, wy0, wy1);
            }
            unsigned m = __ballot_sync(wmask, hit);
            while (m) {
                const uint32_t t = c0 + (uint32_t)__ffs(m) - 1u;
                m &= m - 1;
                if (!done) {
                    const float4 A = lds4(aA + t * 16);
                    const float4 B = lds4(aB + t * 16);
                    const float dx = A.x - px, dy = A.y - py;
                    const float sigma = 0.5f * (B.x * dx * dx + B.z * dy * dy) + B.y * dx * dy;
                    const float alpha = fminf(kAlphaMax, A.z * __expf(-sigma));
                    if (!(sigma < 0.f || alpha < kAlphaMin)) {
                        const float next_T = T * (1.f - alpha);
                        if (next_T <= kTStop) {
                            done = true;
                        } else {
                            const float vis = alpha * T;
#pragma unroll
                            for (int k4 = 0; k4 < CP; k4 += 4) {
                                const float4 c4 = lds4(aC + (t * CP + k4) * 4);
                                pix_out[k4] += c4.x * vis;
                                if (k4 + 1 < CDIM) pix_out[k4 + 1] += c4.y * vis;
                                if (k4 + 2 < CDIM) pix_out[k4 + 2] += c4.z * vis;
                                if (k4 + 3 < CDIM) pix_out[k4 + 3] += c4.w * vis;
                            }
                            cur_idx = (uint32_t)batch_start + t;
                            T = next_T;
                        }
                    }
                }
            }
            if (__ballot_sync(wmask, !done) == 0) { warp_done = true; break; }
        }
    }
    if (inside) {
        render_alphas[pix] = 1.f - T;
#pragma unroll
        for (int k = 0; k < CDIM; ++k) render_colors[pix * CDIM + k] = bg ? (pix_out[k] + T * bg[k]) : pix_out[k];
        last_ids[pix] = (int32_t)cur_idx;
    }
}

template <int CDIM, bool ABS, bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : kBatch, (!BIG && CDIM <= 4) ? GSB_BWD_BLOCKS : 1)
blend_bwd_kernel(BlendGeom G, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                 const float *__restrict__ colors, const float *__restrict__ opacities,
                 const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks,
                 const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                 const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                 const float *__restrict__ v_render_colors, const float *__restrict__ v_render_alphas,
                 const int32_t *__restrict__ tile_order, float *__restrict__ grad_rec) {
    constexpr uint32_t REC = grad_record(CDIM);
    constexpr int CP = ColorPad<CDIM>::value;
    // values reduced per (warp, Gaussian): [rgb(CDIM) conic(3) xy(2)] in groups of 8, then the tail
    // [opac] (one butterfly) or [opac absx absy 0] (one 4-wide spread) when abs-gradients are wanted
    constexpr int NMAIN = CDIM + 5;
    constexpr int NGRP = (NMAIN + 7) / 8;
    constexpr int SP = BIG ? 1 : kStagePerThread;
    extern __shared__ float4 smem4[];
    __shared__ int s_binmax;
    const uint32_t nthreads = blockDim.x, tid = threadIdx.x;
    const uint32_t batch = nthreads * SP;
    float4 *sA = smem4;                                 // x, y, opacity, qmax
    float4 *sB = sA + batch;                            // conic a, b, c, flatten id (bits)
    float *sC = reinterpret_cast<float *>(sB + batch);  // [batch][CP]
    const uint32_t aA = (uint32_t)__cvta_generic_to_shared(sA), aB = (uint32_t)__cvta_generic_to_shared(sB),
                   aC = (uint32_t)__cvta_generic_to_shared(sC);

    const uint32_t ct_lin = tile_order ? (uint32_t)tile_order[blockIdx.x] : blockIdx.x;
    const uint32_t cam = ct_lin / (G.tw * G.th), tile_id = ct_lin % (G.tw * G.th);
    const uint32_t tile_y = tile_id / G.tw, tile_x = tile_id % G.tw;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    if (masks != nullptr && !masks[ct]) return;
    uint32_t lx, ly;
    pixel_in_tile(G.ts, tid, lx, ly);
    const uint32_t i = tile_y * G.ts + ly, j = tile_x * G.ts + lx;
    const bool inside = ly < G.ts && i < G.H && j < G.W;
    const size_t pix = inside ? ((size_t)cam * G.H + i) * G.W + j : 0;
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;
    const int32_t range_start = tile_offsets[ct];
    int32_t range_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];

    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const unsigned wmask = 0xffffffffu;  // blocks are launched with whole warps
    const uint32_t lane = tid & 31;
    const int jmin = __reduce_min_sync(wmask, inside ? (int)j : INT_MAX), jmax = __reduce_max_sync(wmask, inside ? (int)j : -1);
    const int imin = __reduce_min_sync(wmask, inside ? (int)i : INT_MAX), imax = __reduce_max_sync(wmask, inside ? (int)i : -1);
    const float wx0 = (float)jmin + 0.5f, wx1 = (float)jmax + 0.5f, wy0 = (float)imin + 0.5f, wy1 = (float)imax + 0.5f;

    const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
    float T = T_final;
    float buffer[CDIM], v_c[CDIM];
    float bg_dot = 0.f;  // sum_k bg_k * v_render_c_k
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        buffer[k] = 0.f;
        v_c[k] = inside ? v_render_colors[pix * CDIM + k] : 0.f;
        if (bg) bg_dot += bg[k] * v_c[k];
    }
    const float v_a = inside ? v_render_alphas[pix] : 0.f;
    const int32_t bin_final = inside ? last_ids[pix] : -1;
    const int32_t warp_bin_final = __reduce_max_sync(wmask, bin_final);

    // tile-wide last contributing index: nothing behind it needs staging
    if (tid == 0) s_binmax = -1;
    __syncthreads();
    if (lane == 0) atomicMax(&s_binmax, warp_bin_final);
    __syncthreads();
    range_end = min(range_end, s_binmax + 1);

    // lanes that publish totals after reduce_spread<8> / <4>
    const uint32_t slot8 = spread_slot<8>(lane), slot4 = spread_slot<4>(lane);
    const bool pub8 = (lane & 3) == 0, pub4 = (lane & 7) == 0;

    for (int32_t batch_end = range_end - 1; batch_end >= range_start; batch_end -= (int32_t)batch) {
        __syncthreads();
        const int32_t bsize = min((int32_t)batch, batch_end + 1 - range_start);
#pragma unroll
        for (int sp = 0; sp < SP; ++sp) {
            const int32_t e = (int32_t)(tid + sp * nthreads);
            if (e < bsize) {
                const int32_t g = flatten_ids[batch_end - e];
                const float2 xy = __ldg(means2d + g);
                const float opac = __ldg(opacities + g);
                const float ca = __ldg(conics + (size_t)g * 3), cb = __ldg(conics + (size_t)g * 3 + 1),
                            cc = __ldg(conics + (size_t)g * 3 + 2);
                sA[e] = make_float4(xy.x, xy.y, opac, cull_qmax(ca, cb, cc, opac));
                sB[e] = make_float4(ca, cb, cc, __int_as_float(g));
#pragma unroll
                for (int k = 0; k < CDIM; ++k) sC[e * CP + k] = __ldg(colors + (size_t)g * CDIM + k);
            }
        }
        __syncthreads();
        if (warp_bin_final < 0) continue;
        // entry t <-> sorted index batch_end - t ; start at the warp's last contributing Gaussian
        const int32_t t_first = max(0, batch_end - warp_bin_final);
        for (int32_t c0 = t_first; c0 < bsize; c0 += 32) {
            const int32_t e = c0 + (int32_t)lane;
            bool hit = false;
            if (e < bsize) {
                const float4 A = lds4(aA + e * 16), B = lds4(aB + e * 16);
                hit = ellipse_hits_rect(A.x, A.y, B.x, B.y, B.z, A.w, wx0, wx1, wy0, wy1);
            }
            unsigned m = __ballot_sync(wmask, hit);
            while (m) {
                const int32_t t = c0 + __ffs(m) - 1;
                m &= m - 1;
                const float4 A = lds4(aA + t * 16);
                const float4 B = lds4(aB + t * 16);
                const float dx = A.x - px, dy = A.y - py;
                const float opac = A.z;
                const float sigma = 0.5f * (B.x * dx * dx + B.z * dy * dy) + B.y * dx * dy;
                const float vis = __expf(-sigma);
                const float alpha = fminf(kAlphaMax, opac * vis);
                const bool valid = inside && (batch_end - t <= bin_final) && !(sigma < 0.f || alpha < kAlphaMin);
                if (!__any_sync(wmask, valid)) continue;
                float vals[NGRP * 8];
                float tail[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < NGRP * 8; ++k) vals[k] = 0.f;
                if (valid) {
                    const float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
#pragma unroll
                    for (int k4 = 0; k4 < CP; k4 += 4) {
                        const float4 c4 = lds4(aC + (t * CP + k4) * 4);
                        const float cs[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int k = k4 + q;
                            if (k < CDIM) {
                                vals[k] = fac * v_c[k];
                                v_alpha += (cs[q] * T - buffer[k] * ra) * v_c[k];
                                buffer[k] += cs[q] * fac;
                            }
                        }
                    }
                    v_alpha += T_final * ra * v_a;
                    if (bg) v_alpha += -T_final * ra * bg_dot;
                    if (opac * vis <= kAlphaMax) {
                        const float v_sigma = -opac * vis * v_alpha;
                        vals[CDIM + 0] = 0.5f * v_sigma * dx * dx;
                        vals[CDIM + 1] = v_sigma * dx * dy;
                        vals[CDIM + 2] = 0.5f * v_sigma * dy * dy;
                        const float gx = v_sigma * (B.x * dx + B.y * dy);
                        const float gy = v_sigma * (B.y * dx + B.z * dy);
                        vals[CDIM + 3] = gx;
                        vals[CDIM + 4] = gy;
                        tail[0] = vis * v_alpha;
                        if (ABS) { tail[1] = fabsf(gx); tail[2] = fabsf(gy); }
                    }
                }
                // all reductions first (independent shuffle chains overlap), then the REDs: the asm memory
                // clobber of red_add would otherwise serialise the chains
#pragma unroll
                for (int gi = 0; gi < NGRP; ++gi) reduce_spread<8>(vals + gi * 8, lane);
                float tail_sum;
                if (ABS) {
                    reduce_spread<4>(tail, lane);
                    tail_sum = tail[0];
                } else {
                    tail_sum = warp_sum(tail[0]);
                }
                float *rec = grad_rec + (size_t)__float_as_int(B.w) * REC;
#pragma unroll
                for (int gi = 0; gi < NGRP; ++gi) {
                    const uint32_t off = gi * 8 + slot8;
                    if (pub8 && off < (uint32_t)NMAIN) red_add(rec + off, vals[gi * 8]);
                }
                if (ABS) {
                    if (pub4 && slot4 < 3) red_add(rec + NMAIN + slot4, tail_sum);
                } else {
                    if (lane == 0) red_add(rec + NMAIN, tail_sum);
                }
            }
        }
    }
}

// Work order for the blend kernels: (camera, tile) indices bucketed by floor(log2(#intersections)),
// longest first.  One block; the order inside a bucket is irrelevant.
__global__ void __launch_bounds__(1024)
tile_order_kernel(uint32_t total_tiles, int64_t n_isects, const int32_t *__restrict__ offsets, int32_t *__restrict__ order) {
    __shared__ int s_hist[33];
    __shared__ int s_cursor[33];
    if (threadIdx.x < 33) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < total_tiles; t += blockDim.x) {
        const int64_t end = t + 1 == total_tiles ? n_isects : (int64_t)offsets[t + 1];
        const uint32_t n = (uint32_t)(end - offsets[t]);
        atomicAdd(&s_hist[n == 0 ? 32 : __clz(n)], 1);  // clz ascending == length descending
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < 33; ++b) { s_cursor[b] = run; run += s_hist[b]; }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < total_tiles; t += blockDim.x) {
        const int64_t end = t + 1 == total_tiles ? n_isects : (int64_t)offsets[t + 1];
        const uint32_t n = (uint32_t)(end - offsets[t]);
        order[atomicAdd(&s_cursor[n == 0 ? 32 : __clz(n)], 1)] = (int32_t)t;
    }
}

int launch_tile_order(uint32_t total_tiles, int64_t n_isects, const int32_t *offsets, int32_t *order, cudaStream_t s) {
    tile_order_kernel<<<1, 1024, 0, s>>>(total_tiles, n_isects, offsets, order);
    return check_launch("tile_order");
}

// packed gradient records -> the reference's tensors
// MOMENTS: the record holds the moments of w = -v_sigma about the Gaussian's centre (blend_stream_bwd_kernel):
//   v_conic = -(1/2 Sxx, Sxy, 1/2 Syy),  v_xy = -(a Sx + b Sy, b Sx + c Sy),  v_opacity = S0 / opacity
// (rasterize_to_pixels_bwd.cu:225-241 summed over the pixels: every term is linear in w)
template <int CDIM, bool MOMENTS>
__global__ void __launch_bounds__(kThreads)
unpack_grads_kernel(uint32_t M, const float *__restrict__ rec, const float *__restrict__ conics, const float *__restrict__ opacities,
                    float *__restrict__ v_abs, float *__restrict__ v_means2d,
                    float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opac) {
    constexpr uint32_t REC = grad_record(CDIM);
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= M) return;
    float r[REC];
#pragma unroll
    for (uint32_t k = 0; k < REC; k += 4) {
        const float4 v = ldg4(rec + (size_t)g * REC + k);
        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < CDIM; ++k) v_colors[(size_t)g * CDIM + k] = r[k];
    if (MOMENTS) {
        const float sxx = r[CDIM], sxy = r[CDIM + 1], syy = r[CDIM + 2], sx = r[CDIM + 3], sy = r[CDIM + 4], s0 = r[CDIM + 5];
        const bool touched = sxx != 0.f || sxy != 0.f || syy != 0.f || sx != 0.f || sy != 0.f || s0 != 0.f;
        float a = 0.f, b = 0.f, c = 0.f, o = 1.f;
        if (touched) {  // (rows no pixel wrote to stay zero without reading their attributes)
            a = __ldg(conics + (size_t)g * 3); b = __ldg(conics + (size_t)g * 3 + 1); c = __ldg(conics + (size_t)g * 3 + 2);
            o = __ldg(opacities + g);
        }
        r[CDIM] = -0.5f * sxx; r[CDIM + 1] = -sxy; r[CDIM + 2] = -0.5f * syy;
        r[CDIM + 3] = -(a * sx + b * sy); r[CDIM + 4] = -(b * sx + c * sy);
        r[CDIM + 5] = s0 != 0.f ? s0 / o : 0.f;
    }
    v_conics[(size_t)g * 3] = r[CDIM]; v_conics[(size_t)g * 3 + 1] = r[CDIM + 1]; v_conics[(size_t)g * 3 + 2] = r[CDIM + 2];
    *reinterpret_cast<float2 *>(v_means2d + (size_t)g * 2) = make_float2(r[CDIM + 3], r[CDIM + 4]);
    v_opac[g] = r[CDIM + 5];
    if (v_abs != nullptr) *reinterpret_cast<float2 *>(v_abs + (size_t)g * 2) = make_float2(r[CDIM + 6], r[CDIM + 7]);
}

template <int CDIM>
static int launch_unpack(uint32_t M, const float *grad_scratch, float *v_abs, float *v_means2d, float *v_conics, float *v_colors,
                         float *v_opacities, cudaStream_t s, bool moments = false, const float *conics = nullptr,
                         const float *opacities = nullptr) {
    if (moments)
        unpack_grads_kernel<CDIM, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(M, grad_scratch, conics, opacities, v_abs, v_means2d,
                                                                                v_conics, v_colors, v_opacities);
    else
        unpack_grads_kernel<CDIM, false><<<cdiv(M, kThreads), kThreads, 0, s>>>(M, grad_scratch, nullptr, nullptr, v_abs, v_means2d,
                                                                                 v_conics, v_colors, v_opacities);
    return check_launch("unpack_grads");
}

template <int CDIM>
static int launch_fwd(const BlendGeom &G, const float *means2d, const float *conics, const float *colors,
                      const float *opacities, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                      const int32_t *flatten_ids, float *render_colors, float *render_alphas, int32_t *last_ids,
                      int32_t *tile_order, cudaStream_t s) {
    const uint32_t nthreads = (G.ts * G.ts + 31) / 32 * 32;
    const bool big = nthreads > (uint32_t)kBatch;
    const uint32_t batch = nthreads * (big ? 1 : kStagePerThread);
    const size_t smem = (size_t)batch * (32 + 4 * ColorPad<CDIM>::value);
    if (smem > 40 * 1024) {  // (the kernels also hold a little static shared memory)
        if (big) GSB_CUDA_OK(cudaFuncSetAttribute(blend_fwd_kernel<CDIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else GSB_CUDA_OK(cudaFuncSetAttribute(blend_fwd_kernel<CDIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const uint32_t T = G.C * G.tw * G.th;
    if (tile_order != nullptr)
        if (int e = launch_tile_order(T, G.n_isects, tile_offsets, tile_order, s)) return e;
    if (big)
        blend_fwd_kernel<CDIM, true><<<T, nthreads, smem, s>>>(G, reinterpret_cast<const float2 *>(means2d), conics, colors,
                                                                opacities, backgrounds, masks, tile_offsets, flatten_ids,
                                                                render_colors, render_alphas, last_ids, tile_order);
    else
        blend_fwd_kernel<CDIM, false><<<T, nthreads, smem, s>>>(G, reinterpret_cast<const float2 *>(means2d), conics, colors,
                                                                 opacities, backgrounds, masks, tile_offsets, flatten_ids,
                                                                 render_colors, render_alphas, last_ids, tile_order);
    return check_launch("blend_fwd");
}

template <int CDIM>
static int launch_bwd(const BlendGeom &G, uint32_t M, const float *means2d, const float *conics, const float *colors,
                      const float *opacities, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                      const int32_t *flatten_ids, const float *render_alphas, const int32_t *last_ids,
                      const float *v_render_colors, const float *v_render_alphas, const int32_t *tile_order,
                      float *grad_scratch, float *v_abs,
                      float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, cudaStream_t s) {
    const uint32_t nthreads = (G.ts * G.ts + 31) / 32 * 32;
    const bool big = nthreads > (uint32_t)kBatch;
    const uint32_t batch = nthreads * (big ? 1 : kStagePerThread);
    const size_t smem = (size_t)batch * (32 + 4 * ColorPad<CDIM>::value);
#define GSB_BWD_GO(ABSV, BIGV)                                                                                              \
    do {                                                                                                                    \
        if (smem > 40 * 1024)                                                                                               \
            GSB_CUDA_OK(cudaFuncSetAttribute(blend_bwd_kernel<CDIM, ABSV, BIGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                                   \
        blend_bwd_kernel<CDIM, ABSV, BIGV><<<T, nthreads, smem, s>>>(                                                        \
            G, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, backgrounds, masks, tile_offsets,      \
            flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas, tile_order, grad_scratch);              \
    } while (0)
    GSB_CUDA_OK(cudaMemsetAsync(grad_scratch, 0, sizeof(float) * (size_t)M * grad_record(CDIM), s));
    if (G.n_isects > 0) {
        const uint32_t T = G.C * G.tw * G.th;
        if (v_abs != nullptr) { if (big) GSB_BWD_GO(true, true); else GSB_BWD_GO(true, false); }
        else { if (big) GSB_BWD_GO(false, true); else GSB_BWD_GO(false, false); }
        if (int e = check_launch("blend_bwd")) return e;
    }
#undef GSB_BWD_GO
    return launch_unpack<CDIM>(M, grad_scratch, v_abs, v_means2d, v_conics, v_colors, v_opacities, s);
}

}  // namespace gsb200

using namespace gsb200;

#define GSB_DISPATCH_CDIM(D, CALL)                                                                 \
    switch (D) {                                                                                   \
        case 1: { constexpr int CD = 1; return CALL; }                                             \
        case 2: { constexpr int CD = 2; return CALL; }                                             \
        case 3: { constexpr int CD = 3; return CALL; }                                             \
        case 4: { constexpr int CD = 4; return CALL; }                                             \
        case 5: { constexpr int CD = 5; return CALL; }                                             \
        case 8: { constexpr int CD = 8; return CALL; }                                             \
        case 9: { constexpr int CD = 9; return CALL; }                                             \
        case 16: { constexpr int CD = 16; return CALL; }                                           \
        case 17: { constexpr int CD = 17; return CALL; }                                           \
        case 32: { constexpr int CD = 32; return CALL; }                                           \
        case 33: { constexpr int CD = 33; return CALL; }                                           \
        default: return fail(2, "gsb200_rasterize: unsupported number of channels");               \
    }

extern "C" uint32_t gsb200_grad_record(uint32_t D) { return grad_record(D); }

// GSB200_BLEND=tile (or gsb200_set_blend_mode(1)) selects the generic per-tile kernels for every configuration
// (A/B timing and the stream-vs-tile parity tests)
static int g_blend_mode = [] {
    const char *v = getenv("GSB200_BLEND");
    return (v != nullptr && strcmp(v, "tile") == 0) ? 1 : 0;
}();
static bool stream_blend_enabled() { return g_blend_mode == 0; }

extern "C" int gsb200_set_blend_mode(int mode) {
    const int old = g_blend_mode;
    g_blend_mode = mode;
    return old;
}

extern "C" uint32_t gsb200_blend_stream_record(uint32_t D, uint32_t tile_size) {
    return stream_blend_enabled() && blend_stream_supported(tile_size, D, 0) ? 12u : 0u;
}

static int check_geom(uint32_t tile_size, uint32_t W, uint32_t H, uint32_t tw, uint32_t th) {
    if (tile_size == 0 || tile_size > 32) return fail(2, "gsb200_rasterize: tile_size must be in [1, 32]");
    if (tw * tile_size < W || th * tile_size < H)
        return fail(2, "gsb200_rasterize: tile grid does not cover the image (Assert Failed: tile_height * tile_size >= image_height)");
    return 0;
}

extern "C" int gsb200_rasterize_fwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                                    const float *conics, const float *colors, const float *opacities,
                                    const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H, uint32_t tile_size,
                                    uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                                    const int32_t *flatten_ids, float *render_colors, float *render_alphas, int32_t *last_ids,
                                    int32_t *tile_order, float *blend_stream, int32_t *tile_counts,
                                    gsb200_stream_t stream) {
    (void)M;
    if (int e = check_geom(tile_size, W, H, tile_width, tile_height)) return e;
    if (C == 0 || W == 0 || H == 0) return 0;
    BlendGeom G{C, W, H, tile_size, tile_width, tile_height, n_isects};
    cudaStream_t s = (cudaStream_t)stream;
    if (stream_blend_enabled() && blend_stream_supported(tile_size, D, n_isects))
        return blend_stream_fwd(G, D, means2d, conics, colors, opacities, backgrounds, masks, tile_offsets, flatten_ids,
                                render_colors, render_alphas, last_ids, tile_order, blend_stream, tile_counts, s);
    GSB_DISPATCH_CDIM(D, (launch_fwd<CD>(G, means2d, conics, colors, opacities, backgrounds, masks, tile_offsets, flatten_ids,
                                         render_colors, render_alphas, last_ids, tile_order, s)));
}

extern "C" int gsb200_rasterize_bwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                                    const float *conics, const float *colors, const float *opacities,
                                    const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H, uint32_t tile_size,
                                    uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                                    const int32_t *flatten_ids, const float *render_alphas, const int32_t *last_ids,
                                    const float *v_render_colors, const float *v_render_alphas, const int32_t *tile_order,
                                    const float *blend_stream, const int32_t *tile_counts,
                                    float *grad_scratch, float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                                    gsb200_stream_t stream) {
    if (int e = check_geom(tile_size, W, H, tile_width, tile_height)) return e;
    if (M == 0) return 0;
    BlendGeom G{C, W, H, tile_size, tile_width, tile_height, n_isects};
    cudaStream_t s = (cudaStream_t)stream;
    if (blend_stream != nullptr && tile_counts != nullptr && stream_blend_enabled() &&
        blend_stream_supported(tile_size, D, n_isects)) {
        GSB_CUDA_OK(cudaMemsetAsync(grad_scratch, 0, sizeof(float) * (size_t)M * grad_record(D), s));
        if (n_isects > 0)
            if (int e = blend_stream_bwd(G, D, M, backgrounds, masks, tile_offsets, render_alphas, last_ids, v_render_colors,
                                         v_render_alphas, tile_order, blend_stream, tile_counts, grad_scratch,
                                         v_means2d_abs != nullptr, s))
                return e;
        const bool moments = blend_stream_bwd_moments(v_means2d_abs != nullptr);
        GSB_DISPATCH_CDIM(D, (launch_unpack<CD>(M, grad_scratch, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, s, moments,
                                                conics, opacities)));
    }
    GSB_DISPATCH_CDIM(D, (launch_bwd<CD>(G, M, means2d, conics, colors, opacities, backgrounds, masks, tile_offsets,
                                         flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas, tile_order,
                                         grad_scratch, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, s)));
}
