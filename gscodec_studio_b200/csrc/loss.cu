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
constexpr int kTile = 16;
constexpr int kSpan = kTile + 2 * kHalo;  // 26

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

__global__ void __launch_bounds__(kTile * kTile)
photometric_fwd_kernel(LossGeom G, SsimWindow win, const float *__restrict__ img1, const float *__restrict__ img2,
                       float *__restrict__ map_mu, float *__restrict__ map_s1, float *__restrict__ map_s12, LossWork *work,
                       float *__restrict__ out, float *__restrict__ stats) {
    __shared__ float sX[kSpan][kSpan + 1], sY[kSpan][kSpan + 1];
    __shared__ float sH[5][kSpan][kTile + 1];
    __shared__ double s_red[2][kTile * kTile / 32];
    const int tid = threadIdx.x, lx = tid % kTile, ly = tid / kTile;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, b = blockIdx.z;
    const int px = x0 + lx, py = y0 + ly;
    const bool inside = px < (int)G.W && py < (int)G.H;
    const size_t img_base = (size_t)b * G.H * G.W;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float l1_acc = 0.f, ssim_acc = 0.f;

    for (uint32_t ch = 0; ch < G.CH; ++ch) {
        __syncthreads();  // previous channel's readers are done with the staging arrays
        for (int i = tid; i < kSpan * kSpan; i += kTile * kTile) {
            const int r = i / kSpan, c = i - r * kSpan;
            const int gy = y0 - kHalo + r, gx = x0 - kHalo + c;
            float a = 0.f, d = 0.f;
            if (gy >= 0 && gy < (int)G.H && gx >= 0 && gx < (int)G.W) {
                const size_t o = (img_base + (size_t)gy * G.W + gx) * G.CH + ch;
                a = __ldg(img1 + o);
                d = __ldg(img2 + o);
            }
            sX[r][c] = a;
            sY[r][c] = d;
        }
        __syncthreads();
        for (int i = tid; i < kSpan * kTile; i += kTile * kTile) {  // horizontal pass
            const int r = i / kTile, c = i - r * kTile;
            float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float a = sX[r][c + k], d = sY[r][c + k], w = win.w[k];
                h0 += w * a; h1 += w * d; h2 += w * a * a; h3 += w * d * d; h4 += w * a * d;
            }
            sH[0][r][c] = h0; sH[1][r][c] = h1; sH[2][r][c] = h2; sH[3][r][c] = h3; sH[4][r][c] = h4;
        }
        __syncthreads();
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {  // vertical pass
            const float w = win.w[k];
            mu1 += w * sH[0][ly + k][lx]; mu2 += w * sH[1][ly + k][lx];
            e11 += w * sH[2][ly + k][lx]; e22 += w * sH[3][ly + k][lx]; e12 += w * sH[4][ly + k][lx];
        }
        if (inside) {
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
            l1_acc += fabsf(sX[ly + kHalo][lx + kHalo] - sY[ly + kHalo][lx + kHalo]);
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
        for (int w = 0; w < kTile * kTile / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
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
__global__ void __launch_bounds__(kTile * kTile)
photometric_bwd_kernel(LossGeom G, SsimWindow win, const float *__restrict__ img1, const float *__restrict__ img2,
                       const float *__restrict__ map_mu, const float *__restrict__ map_s1, const float *__restrict__ map_s12,
                       const float *__restrict__ v_loss, float *__restrict__ v_img1) {
    __shared__ float sM[3][kSpan][kSpan + 1];
    __shared__ float sH[3][kSpan][kTile + 1];
    const int tid = threadIdx.x, lx = tid % kTile, ly = tid / kTile;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, b = blockIdx.z;
    const int px = x0 + lx, py = y0 + ly;
    const bool inside = px < (int)G.W && py < (int)G.H;
    const size_t img_base = (size_t)b * G.H * G.W;
    const double n_all = (double)G.B * G.H * G.W * G.CH;
    const double n_ssim = G.valid ? (double)G.B * G.CH * (double)(G.H - 2 * kHalo) * (double)(G.W - 2 * kHalo) : n_all;
    const float up = __ldg(v_loss);
    const float k_l1 = up * (float)((double)G.coef_l1 / n_all), k_ssim = up * (float)((double)G.coef_ssim / n_ssim);

    for (uint32_t ch = 0; ch < G.CH; ++ch) {
        __syncthreads();
        for (int i = tid; i < kSpan * kSpan; i += kTile * kTile) {
            const int r = i / kSpan, c = i - r * kSpan;
            const int gy = y0 - kHalo + r, gx = x0 - kHalo + c;
            float a = 0.f, d = 0.f, e = 0.f;
            if (gy >= 0 && gy < (int)G.H && gx >= 0 && gx < (int)G.W) {
                const size_t o = (img_base + (size_t)gy * G.W + gx) * G.CH + ch;
                a = __ldg(map_mu + o); d = __ldg(map_s1 + o); e = __ldg(map_s12 + o);
            }
            sM[0][r][c] = a; sM[1][r][c] = d; sM[2][r][c] = e;
        }
        __syncthreads();
        for (int i = tid; i < kSpan * kTile; i += kTile * kTile) {
            const int r = i / kTile, c = i - r * kTile;
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float w = win.w[k];
                h0 += w * sM[0][r][c + k]; h1 += w * sM[1][r][c + k]; h2 += w * sM[2][r][c + k];
            }
            sH[0][r][c] = h0; sH[1][r][c] = h1; sH[2][r][c] = h2;
        }
        __syncthreads();
        float cM = 0.f, cS1 = 0.f, cS12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[k];
            cM += w * sH[0][ly + k][lx]; cS1 += w * sH[1][ly + k][lx]; cS12 += w * sH[2][ly + k][lx];
        }
        if (inside) {
            const size_t o = (img_base + (size_t)py * G.W + px) * G.CH + ch;
            const float x = __ldg(img1 + o), y = __ldg(img2 + o);
            const float d = x - y;
            const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            v_img1[o] = k_l1 * sgn + k_ssim * (cM + 2.f * x * cS1 + y * cS12);
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
    const dim3 grid(cdiv(W, kTile), cdiv(H, kTile), B);
    photometric_fwd_kernel<<<grid, kTile * kTile, 0, s>>>(G, make_window(), img1, img2, map_mu, map_s1, map_s12,
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
    if (img1 == nullptr || img2 == nullptr || map_mu == nullptr || map_s1 == nullptr || map_s12 == nullptr || v_loss == nullptr ||
        v_img1 == nullptr)
        return failf(2, "%s: null pointer", who);
    const LossGeom G{B, H, W, CH, padding_valid ? 1 : 0, coef_l1, coef_ssim, 0.f};
    const dim3 grid(cdiv(W, kTile), cdiv(H, kTile), B);
    photometric_bwd_kernel<<<grid, kTile * kTile, 0, (cudaStream_t)stream>>>(G, make_window(), img1, img2, map_mu, map_s1, map_s12,
                                                                             v_loss, v_img1);
    return check_launch("photometric_loss_bwd");
}
