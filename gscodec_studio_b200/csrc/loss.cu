// Photometric loss next to the hot path: L1 + SSIM of the rendered image against the
// target, forward and backward, on the renderer's own [B,H,W,CH] layout.
//
// Reference call site: examples/simple_trainer.py:948-953 (also simple_trainer_dyngs.py:647, simple_trainer_STG.py:694)
//     l1loss   = F.l1_loss(colors, pixels)
//     ssimloss = 1.0 - fused_ssim(colors.permute(0, 3, 1, 2), pixels.permute(0, 3, 1, 2), padding="valid")
//     loss     = l1loss * (1.0 - ssim_lambda) + ssimloss * ssim_lambda
// `fused_ssim` is a third-party dependency that is NOT vendored in the reference tree (examples/requirements.txt:26
// pins rahul-goel/fused-ssim@1272e21a); its published algorithm is restated here: SSIM with an 11-tap Gaussian window
// (sigma 1.5, zero padding at the image border), C1 = 0.01^2, C2 = 0.03^2, per channel; padding="valid" averages the
// SSIM map without its 5-pixel border; the gradient flows to the first image only.
//
// Here: one forward kernel computes |x - y|, the five windowed moments (separable, staged in shared memory), the
// SSIM map value and the three partial-derivative maps the backward needs, reduces both sums and -- in the last
// block to finish -- writes the scalar loss; one backward kernel convolves the derivative maps and adds the L1
// sign term.  The permuted NCHW view, its .contiguous() copy and the separate L1 kernels of the reference chain
// (5 elementwise launches) disappear.
#include <math.h>

#include "common.cuh"

namespace gsb200 {

constexpr int kWin = 11, kHalo = 5;
constexpr int kTW = 32, kTH = 32;                                // output tile of one CTA
constexpr int kSpanX = kTW + 2 * kHalo, kSpanY = kTH + 2 * kHalo;  // 42 x 42 staged inputs
constexpr int kLossThreads = 256;
constexpr int kRowsPerThread = kTH / (kLossThreads / kTW);  // vertical pass: 4 output rows per thread
constexpr int kColsPerItem = 4;                             // horizontal pass: 4 output columns per work item
constexpr int kTaps = kColsPerItem + kWin - 1;              // 14 inputs feed 4 outputs
static_assert(kRowsPerThread == 4 && kTW % kColsPerItem == 0, "tiling");

struct SsimWindow {
    float w[kWin];
};

struct LossGeom {
    uint32_t B, H, W, CH;
    int valid;  // 1: padding="valid" (SSIM averaged over the interior), 0: "same"
    float coef_l1, coef_ssim, bias;  // loss = coef_l1 * mean|x-y| + coef_ssim * mean(ssim) + bias
};

struct LossWork {  // zeroed before the launch
    double l1_sum, ssim_sum;
    unsigned int done;
};

__device__ __forceinline__ bool ssim_counts(const LossGeom &G, int y, int x) {
    if (!G.valid) return true;
    return y >= kHalo && x >= kHalo && y < (int)G.H - kHalo && x < (int)G.W - kHalo;
}

// Stage one channel of a (kSpanY x kSpanX) window around the tile, zero outside the image.
template <int NA>
__device__ __forceinline__ void stage_window(const LossGeom &G, size_t img_base, int x0, int y0, uint32_t ch,
                                             const float *const (&src)[NA], float (*const (&dst)[NA])[kSpanX + 1]) {
    for (int i = threadIdx.x; i < kSpanY * kSpanX; i += kLossThreads) {
        const int r = i / kSpanX, c = i - r * kSpanX;
        const int gy = y0 - kHalo + r, gx = x0 - kHalo + c;
        const bool in = gy >= 0 && gy < (int)G.H && gx >= 0 && gx < (int)G.W;
        const size_t o = (img_base + (size_t)gy * G.W + gx) * G.CH + ch;
#pragma unroll
        for (int a = 0; a < NA; ++a) dst[a][r][c] = in ? __ldg(src[a] + o) : 0.f;
    }
}

// The separable 11-tap window is applied with register tiling, because the kernel is bound by shared-memory
// wavefronts, not by FMAs: a horizontal work item reads 14 inputs for 4 adjacent outputs (lanes of a warp own
// different rows, odd row stride: conflict-free), a thread of the vertical pass reads 14 rows for its 4 output rows
// (lanes own adjacent columns).  ~1.4 shared-memory wavefronts per pixel and channel instead of 6 for the
// one-output-per-thread form.
__global__ void __launch_bounds__(kLossThreads)
photometric_fwd_kernel(LossGeom G, SsimWindow win, const float *__restrict__ img1, const float *__restrict__ img2,
                       float *__restrict__ map_mu, float *__restrict__ map_s1, float *__restrict__ map_s12, LossWork *work,
                       float *__restrict__ out, float *__restrict__ stats) {
    __shared__ float sX[kSpanY][kSpanX + 1], sY[kSpanY][kSpanX + 1];
    __shared__ float sH[5][kSpanY][kTW + 1];
    __shared__ double s_red[2][kLossThreads / 32];
    const int tid = threadIdx.x, lx = tid % kTW, rg = tid / kTW;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, b = blockIdx.z;
    const int px = x0 + lx, py0 = y0 + rg * kRowsPerThread;
    const size_t img_base = (size_t)b * G.H * G.W;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const bool with_ssim = G.coef_ssim != 0.f;
    float l1_acc = 0.f, ssim_acc = 0.f;

    if (!with_ssim) {  // L1 only: no window, nothing staged
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
            const int py = py0 + j;
            if (px < (int)G.W && py < (int)G.H) {
                const size_t o = (img_base + (size_t)py * G.W + px) * G.CH;
                for (uint32_t ch = 0; ch < G.CH; ++ch) l1_acc += fabsf(__ldg(img1 + o + ch) - __ldg(img2 + o + ch));
            }
        }
    }
    for (uint32_t ch = 0; with_ssim && ch < G.CH; ++ch) {
        __syncthreads();  // the previous channel's readers are done with the staging arrays
        {
            const float *const src[2] = {img1, img2};
            float(*const dst[2])[kSpanX + 1] = {sX, sY};
            stage_window<2>(G, img_base, x0, y0, ch, src, dst);
        }
        __syncthreads();
        // horizontal pass: item = (row r, column group cg); the lanes of a warp cover consecutive rows
        for (int it = tid; it < (kTW / kColsPerItem) * kSpanY; it += kLossThreads) {
            const int cg = it / kSpanY, r = it - cg * kSpanY;
            float h[kColsPerItem][5];
#pragma unroll
            for (int j = 0; j < kColsPerItem; ++j)
#pragma unroll
                for (int q = 0; q < 5; ++q) h[j][q] = 0.f;
#pragma unroll
            for (int k = 0; k < kTaps; ++k) {  // every input and its three products are formed once
                const float a = sX[r][cg * kColsPerItem + k], d = sY[r][cg * kColsPerItem + k];
                const float aa = a * a, dd = d * d, ad = a * d;
#pragma unroll
                for (int j = 0; j < kColsPerItem; ++j) {
                    if (k - j >= 0 && k - j < kWin) {
                        const float w = win.w[k - j];
                        h[j][0] += w * a; h[j][1] += w * d; h[j][2] += w * aa; h[j][3] += w * dd; h[j][4] += w * ad;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kColsPerItem; ++j) {
                const int c = cg * kColsPerItem + j;
                sH[0][r][c] = h[j][0]; sH[1][r][c] = h[j][1]; sH[2][r][c] = h[j][2]; sH[3][r][c] = h[j][3]; sH[4][r][c] = h[j][4];
            }
        }
        __syncthreads();
        // vertical pass: this thread's column lx, output rows rg*4 .. rg*4+3
        float acc[kRowsPerThread][5];
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j)
#pragma unroll
            for (int q = 0; q < 5; ++q) acc[j][q] = 0.f;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const int row = rg * kRowsPerThread + k;
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = sH[q][row][lx];
#pragma unroll
            for (int j = 0; j < kRowsPerThread; ++j) {
                if (k - j >= 0 && k - j < kWin) {
                    const float w = win.w[k - j];
#pragma unroll
                    for (int q = 0; q < 5; ++q) acc[j][q] += w * v[q];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
            const int py = py0 + j;
            if (px < (int)G.W && py < (int)G.H) {
                const float mu1 = acc[j][0], mu2 = acc[j][1], e11 = acc[j][2], e22 = acc[j][3], e12 = acc[j][4];
                const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
                const float A = mu1 * mu1 + mu2 * mu2 + C1, Bq = s1 + s2 + C2;
                const float Cq = 2.f * mu1 * mu2 + C1, D = 2.f * s12 + C2;
                const float iA = 1.f / A, iB = 1.f / Bq;
                const float m = Cq * D * iA * iB;
                // partials with (mu1, sigma1^2, sigma12) independent, then the chain through sigma1^2 = E11 - mu1^2 and
                // sigma12 = E12 - mu1 mu2 folded into the mu term:  dm/dx(p) = w(q - p) [ M + 2 x(p) S1 + y(p) S12 ]
                const float dS1 = -m * iB;
                const float dS12 = 2.f * Cq * iA * iB;
                const float dMu = 2.f * D * iA * iB * (mu2 - mu1 * Cq * iA);
                const bool counts = ssim_counts(G, py, px);
                const size_t o = (img_base + (size_t)py * G.W + px) * G.CH + ch;
                if (map_mu != nullptr) {
                    map_mu[o] = counts ? dMu - 2.f * mu1 * dS1 - mu2 * dS12 : 0.f;
                    map_s1[o] = counts ? dS1 : 0.f;
                    map_s12[o] = counts ? dS12 : 0.f;
                }
                if (counts) ssim_acc += m;
                const int r = rg * kRowsPerThread + j + kHalo;
                l1_acc += fabsf(sX[r][lx + kHalo] - sY[r][lx + kHalo]);
            }
        }
    }
    // block reduction (double: 6 M terms at 1080p), then the last block to arrive writes the scalars
    double d1 = (double)l1_acc, d2 = (double)ssim_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = d1; s_red[1][tid >> 5] = d2; }
    __syncthreads();
    if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
        atomicAdd(&work->l1_sum, t1);
        atomicAdd(&work->ssim_sum, t2);
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&work->done, 1u) == total - 1) {
            __threadfence();
            const double l1 = atomicAdd(&work->l1_sum, 0.0), ss = atomicAdd(&work->ssim_sum, 0.0);
            const double n_all = (double)G.B * G.H * G.W * G.CH;
            const double n_ssim = G.valid ? (double)G.B * G.CH * (double)(G.H - 2 * kHalo) * (double)(G.W - 2 * kHalo) : n_all;
            const double l1_mean = l1 / n_all, ssim_mean = ss / n_ssim;
            out[0] = (float)((double)G.coef_l1 * l1_mean + (double)G.coef_ssim * ssim_mean + (double)G.bias);
            if (stats != nullptr) { stats[0] = (float)l1_mean; stats[1] = (float)ssim_mean; }
        }
    }
}

// v_img1 = v_loss * [ coef_l1 / n_all * sign(x - y) + coef_ssim / n_ssim * (conv(M) + 2 x conv(S1) + y conv(S12)) ]
__global__ void __launch_bounds__(kLossThreads)
photometric_bwd_kernel(LossGeom G, SsimWindow win, const float *__restrict__ img1, const float *__restrict__ img2,
                       const float *__restrict__ map_mu, const float *__restrict__ map_s1, const float *__restrict__ map_s12,
                       const float *__restrict__ v_loss, float *__restrict__ v_img1) {
    __shared__ float sM[3][kSpanY][kSpanX + 1];
    __shared__ float sH[3][kSpanY][kTW + 1];
    const int tid = threadIdx.x, lx = tid % kTW, rg = tid / kTW;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, b = blockIdx.z;
    const int px = x0 + lx, py0 = y0 + rg * kRowsPerThread;
    const size_t img_base = (size_t)b * G.H * G.W;
    const double n_all = (double)G.B * G.H * G.W * G.CH;
    const double n_ssim = G.valid ? (double)G.B * G.CH * (double)(G.H - 2 * kHalo) * (double)(G.W - 2 * kHalo) : n_all;
    const float up = __ldg(v_loss);
    const float k_l1 = up * (float)((double)G.coef_l1 / n_all), k_ssim = up * (float)((double)G.coef_ssim / n_ssim);
    const bool with_ssim = G.coef_ssim != 0.f;

    if (!with_ssim) {
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
            const int py = py0 + j;
            if (px < (int)G.W && py < (int)G.H) {
                const size_t o = (img_base + (size_t)py * G.W + px) * G.CH;
                for (uint32_t ch = 0; ch < G.CH; ++ch) {
                    const float d = __ldg(img1 + o + ch) - __ldg(img2 + o + ch);
                    v_img1[o + ch] = k_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
                }
            }
        }
        return;
    }
    for (uint32_t ch = 0; ch < G.CH; ++ch) {
        __syncthreads();
        {
            const float *const src[3] = {map_mu, map_s1, map_s12};
            float(*const dst[3])[kSpanX + 1] = {sM[0], sM[1], sM[2]};
            stage_window<3>(G, img_base, x0, y0, ch, src, dst);
        }
        __syncthreads();
        for (int it = tid; it < (kTW / kColsPerItem) * kSpanY; it += kLossThreads) {
            const int cg = it / kSpanY, r = it - cg * kSpanY;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float ma[kTaps];
#pragma unroll
                for (int k = 0; k < kTaps; ++k) ma[k] = sM[q][r][cg * kColsPerItem + k];
#pragma unroll
                for (int j = 0; j < kColsPerItem; ++j) {
                    float h = 0.f;
#pragma unroll
                    for (int k = 0; k < kWin; ++k) h += win.w[k] * ma[j + k];
                    sH[q][r][cg * kColsPerItem + j] = h;
                }
            }
        }
        __syncthreads();
        float acc[kRowsPerThread][3];
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; }
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const int row = rg * kRowsPerThread + k;
            const float v0 = sH[0][row][lx], v1 = sH[1][row][lx], v2 = sH[2][row][lx];
#pragma unroll
            for (int j = 0; j < kRowsPerThread; ++j) {
                if (k - j >= 0 && k - j < kWin) {
                    const float w = win.w[k - j];
                    acc[j][0] += w * v0; acc[j][1] += w * v1; acc[j][2] += w * v2;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
            const int py = py0 + j;
            if (px < (int)G.W && py < (int)G.H) {
                const size_t o = (img_base + (size_t)py * G.W + px) * G.CH + ch;
                const float x = __ldg(img1 + o), y = __ldg(img2 + o);
                const float d = x - y;
                const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                v_img1[o] = k_l1 * sgn + k_ssim * (acc[j][0] + 2.f * x * acc[j][1] + y * acc[j][2]);
            }
        }
    }
}

static SsimWindow make_window() {
    SsimWindow w;
    double g[kWin], sum = 0.0;
    for (int k = 0; k < kWin; ++k) {
        const double d = (double)(k - kHalo);
        g[k] = exp(-(d * d) / (2.0 * 1.5 * 1.5));
        sum += g[k];
    }
    for (int k = 0; k < kWin; ++k) w.w[k] = (float)(g[k] / sum);
    return w;
}

static int check_loss_geom(const char *who, uint32_t B, uint32_t H, uint32_t W, uint32_t CH, int valid) {
    if (CH == 0 || CH > 64) return failf(2, "%s: channels must be in [1, 64]", who);
    if (valid && (H <= 2 * kHalo || W <= 2 * kHalo)) return failf(2, "%s: padding=valid needs an image larger than 10x10", who);
    if (B > 65535) return failf(2, "%s: at most 65535 images per call", who);
    return 0;
}

}  // namespace gsb200

using namespace gsb200;

extern "C" uint64_t gsb200_photometric_workspace_bytes(void) { return sizeof(LossWork); }

extern "C" int gsb200_photometric_loss_fwd(uint32_t B, uint32_t H, uint32_t W, uint32_t CH, const float *img1, const float *img2,
                                           float coef_l1, float coef_ssim, float bias, int padding_valid, float *map_mu,
                                           float *map_s1, float *map_s12, void *workspace, float *out, float *stats,
                                           gsb200_stream_t stream) {
    const char *who = "gsb200_photometric_loss_fwd";
    if (int rc = check_loss_geom(who, B, H, W, CH, padding_valid)) return rc;
    if (img1 == nullptr || img2 == nullptr || workspace == nullptr || out == nullptr) return failf(2, "%s: null pointer", who);
    if ((map_mu == nullptr) != (map_s1 == nullptr) || (map_mu == nullptr) != (map_s12 == nullptr))
        return failf(2, "%s: pass all three derivative maps or none", who);
    if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return failf(2, "%s: workspace must be 8-byte aligned", who);
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(LossWork), s));
    if (B == 0 || H == 0 || W == 0) return 0;
    const LossGeom G{B, H, W, CH, padding_valid ? 1 : 0, coef_l1, coef_ssim, bias};
    const dim3 grid(cdiv(W, kTW), cdiv(H, kTH), B);
    photometric_fwd_kernel<<<grid, kLossThreads, 0, s>>>(G, make_window(), img1, img2, map_mu, map_s1, map_s12,
                                                          reinterpret_cast<LossWork *>(workspace), out, stats);
    return check_launch("photometric_loss_fwd");
}

extern "C" int gsb200_photometric_loss_bwd(uint32_t B, uint32_t H, uint32_t W, uint32_t CH, const float *img1, const float *img2,
                                           float coef_l1, float coef_ssim, int padding_valid, const float *map_mu,
                                           const float *map_s1, const float *map_s12, const float *v_loss, float *v_img1,
                                           gsb200_stream_t stream) {
    const char *who = "gsb200_photometric_loss_bwd";
    if (int rc = check_loss_geom(who, B, H, W, CH, padding_valid)) return rc;
    if (B == 0 || H == 0 || W == 0) return 0;
    if (img1 == nullptr || img2 == nullptr || v_loss == nullptr || v_img1 == nullptr) return failf(2, "%s: null pointer", who);
    if (coef_ssim != 0.f && (map_mu == nullptr || map_s1 == nullptr || map_s12 == nullptr))
        return failf(2, "%s: the derivative maps of the forward are required when coef_ssim != 0", who);
    const LossGeom G{B, H, W, CH, padding_valid ? 1 : 0, coef_l1, coef_ssim, 0.f};
    const dim3 grid(cdiv(W, kTW), cdiv(H, kTH), B);
    photometric_bwd_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(G, make_window(), img1, img2, map_mu, map_s1, map_s12,
                                                                             v_loss, v_img1);
    return check_launch("photometric_loss_bwd");
}
