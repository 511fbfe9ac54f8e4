// Optimizer step next to the hot path: Adam over the per-Gaussian parameter tensors.
//
//   * gsb200_selective_adam_update replaces _C.selective_adam_update (gsplat/cuda/csrc/adam.cu:13-83, bound
//     at ext.cpp; called by gsplat/optimizers/selective_adam.py:42-94): rows whose visibility flag is clear
//     keep parameter AND moments untouched; no bias correction (the reference has none).
//   * gsb200_adam_multi is torch.optim.Adam's update (bias-corrected, eps added after the sqrt; the optimizer
//     examples/simple_trainer.py:396-408 builds per attribute) for up to GSB200_ADAM_MAX_TENSORS tensors in ONE
//     launch: the trainer steps six optimizers per iteration, each a chain of foreach kernels that re-read the
//     same arrays; here every element of (param, grad, m, v) is read once and (param, m, v) written once --
//     28 B of traffic per float, the HBM floor of the update.
//
// Both are streaming kernels: 128-bit loads / stores, four independent vectors in flight per thread.
#include "common.cuh"

namespace gsb200 {

constexpr int kAdamUnroll = 4;
constexpr int kAdamChunk = kThreads * kAdamUnroll;  // float4 vectors (or scalars, unaligned form) per block

struct AdamTable {
    gsb200_adam_tensor t[GSB200_ADAM_MAX_TENSORS];
    uint32_t chunk_end[GSB200_ADAM_MAX_TENSORS];  // exclusive prefix of blocks, per tensor
    uint32_t vec[GSB200_ADAM_MAX_TENSORS];        // 1: all four arrays 16-byte aligned
    uint32_t n;
};

__device__ __forceinline__ void adam_elem(float &p, float g, float &m, float &v, const gsb200_adam_tensor &h) {
    m = h.beta1 * m + h.one_minus_beta1 * g;  // 1 - beta is rounded from double on the host, as torch does
    v = h.beta2 * v + h.one_minus_beta2 * g * g;
    const float denom = sqrtf(v) * h.inv_sqrt_bias2 + h.eps;
    p -= h.step_size * (m / denom);
}

__global__ void __launch_bounds__(kThreads) adam_multi_kernel(const __grid_constant__ AdamTable T) {
    uint32_t ti = 0;
    while (ti + 1 < T.n && blockIdx.x >= T.chunk_end[ti]) ++ti;
    const gsb200_adam_tensor &h = T.t[ti];
    const uint32_t chunk = blockIdx.x - (ti == 0 ? 0u : T.chunk_end[ti - 1]);
    if (T.vec[ti]) {
        const int64_t nvec = h.n / 4;
        const int64_t base = (int64_t)chunk * kAdamChunk + threadIdx.x;
        float4 p[kAdamUnroll], g[kAdamUnroll], m[kAdamUnroll], v[kAdamUnroll];
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            if (i < nvec) {
                p[u] = *reinterpret_cast<const float4 *>(h.param + i * 4);
                g[u] = ldg4(h.grad + i * 4);
                m[u] = *reinterpret_cast<const float4 *>(h.exp_avg + i * 4);
                v[u] = *reinterpret_cast<const float4 *>(h.exp_avg_sq + i * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            if (i < nvec) {
                adam_elem(p[u].x, g[u].x, m[u].x, v[u].x, h);
                adam_elem(p[u].y, g[u].y, m[u].y, v[u].y, h);
                adam_elem(p[u].z, g[u].z, m[u].z, v[u].z, h);
                adam_elem(p[u].w, g[u].w, m[u].w, v[u].w, h);
                stg4(h.param + i * 4, p[u]);
                stg4(h.exp_avg + i * 4, m[u]);
                stg4(h.exp_avg_sq + i * 4, v[u]);
            }
        }
        // the n % 4 trailing scalars belong to the tensor's first block
        if (chunk == 0 && threadIdx.x < (uint32_t)(h.n - nvec * 4)) {
            const int64_t i = nvec * 4 + threadIdx.x;
            float pp = h.param[i], mm = h.exp_avg[i], vv = h.exp_avg_sq[i];
            adam_elem(pp, h.grad[i], mm, vv, h);
            h.param[i] = pp; h.exp_avg[i] = mm; h.exp_avg_sq[i] = vv;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            const int64_t i = (int64_t)chunk * kAdamChunk + (int64_t)u * kThreads + threadIdx.x;
            if (i < h.n) {
                float pp = h.param[i], mm = h.exp_avg[i], vv = h.exp_avg_sq[i];
                adam_elem(pp, h.grad[i], mm, vv, h);
                h.param[i] = pp; h.exp_avg[i] = mm; h.exp_avg_sq[i] = vv;
            }
        }
    }
}

// adam.cu:31-40: no bias correction, eps after the sqrt
__device__ __forceinline__ void selective_elem(float &p, float g, float &m, float &v, float lr, float b1, float b2, float eps) {
    m = b1 * m + (1.0f - b1) * g;
    v = b2 * v + (1.0f - b2) * g * g;
    p += -lr * m / (sqrtf(v) + eps);
}

// row-major [N, M] parameter, one flag per row gates the whole update of that row (parameter and moments)
__global__ void __launch_bounds__(kThreads)
selective_adam_vec_kernel(float *__restrict__ param, const float *__restrict__ grad, float *__restrict__ exp_avg,
                          float *__restrict__ exp_avg_sq, const uint8_t *__restrict__ visible, float lr, float b1, float b2,
                          float eps, uint64_t total, uint32_t M) {  // total % 4 == 0
    const uint64_t i0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (i0 >= total) return;
    bool on[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) on[q] = visible == nullptr || visible[(i0 + q) / M] != 0;
    if (!(on[0] || on[1] || on[2] || on[3])) return;
    float4 p = *reinterpret_cast<const float4 *>(param + i0);
    const float4 g = ldg4(grad + i0);
    float4 m = *reinterpret_cast<const float4 *>(exp_avg + i0), v = *reinterpret_cast<const float4 *>(exp_avg_sq + i0);
    if (on[0]) selective_elem(p.x, g.x, m.x, v.x, lr, b1, b2, eps);
    if (on[1]) selective_elem(p.y, g.y, m.y, v.y, lr, b1, b2, eps);
    if (on[2]) selective_elem(p.z, g.z, m.z, v.z, lr, b1, b2, eps);
    if (on[3]) selective_elem(p.w, g.w, m.w, v.w, lr, b1, b2, eps);
    stg4(param + i0, p);
    stg4(exp_avg + i0, m);
    stg4(exp_avg_sq + i0, v);
}

__global__ void __launch_bounds__(kThreads)
selective_adam_scalar_kernel(float *__restrict__ param, const float *__restrict__ grad, float *__restrict__ exp_avg,
                             float *__restrict__ exp_avg_sq, const uint8_t *__restrict__ visible, float lr, float b1, float b2,
                             float eps, uint64_t total, uint32_t M) {
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    if (visible != nullptr && visible[i / M] == 0) return;
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    selective_elem(p, grad[i], m, v, lr, b1, b2, eps);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
}

// MCMC relocation (gsplat/cuda/csrc/compute_relocation.cu:13-40; Eq. 9 of "3D Gaussian Splatting as Markov Chain
// Monte Carlo"): opacity and scale of a Gaussian that is split into `ratio` copies.
__global__ void __launch_bounds__(kThreads)
relocation_kernel(uint32_t N, const float *__restrict__ opacities, const float *__restrict__ scales, const int32_t *__restrict__ ratios,
                  const float *__restrict__ binoms, int n_max, float *__restrict__ new_opacities, float *__restrict__ new_scales) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    const int n = ratios[g];
    const float opac = opacities[g];
    const float new_opac = 1.0f - powf(1.0f - opac, 1.0f / (float)n);
    new_opacities[g] = new_opac;
    float denom = 0.0f;
    for (int i = 1; i <= n; ++i) {
        float pw = new_opac;  // new_opac^(k+1)
        for (int k = 0; k <= i - 1; ++k) {
            const float term = ((k & 1) ? -1.0f : 1.0f) / sqrtf((float)(k + 1)) * pw;
            denom += __ldg(binoms + (size_t)(i - 1) * n_max + k) * term;
            pw *= new_opac;
        }
    }
    const float coeff = opac / denom;
#pragma unroll
    for (int k = 0; k < 3; ++k) new_scales[(size_t)g * 3 + k] = coeff * scales[(size_t)g * 3 + k];
}

static bool aligned16(const void *a, const void *b, const void *c, const void *d) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_selective_adam_update(float *param, const float *param_grad, float *exp_avg, float *exp_avg_sq,
                                            const uint8_t *visibility, float lr, float b1, float b2, float eps, uint32_t N,
                                            uint32_t M, gsb200_stream_t stream) {
    if (N == 0 || M == 0) return 0;
    if (param == nullptr || param_grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr)
        return fail(2, "gsb200_selective_adam_update: null pointer");
    const uint64_t total = (uint64_t)N * M;
    cudaStream_t s = (cudaStream_t)stream;
    if (total % 4 == 0 && aligned16(param, param_grad, exp_avg, exp_avg_sq)) {
        const uint64_t blocks = (total / 4 + kThreads - 1) / kThreads;
        selective_adam_vec_kernel<<<(uint32_t)blocks, kThreads, 0, s>>>(param, param_grad, exp_avg, exp_avg_sq, visibility, lr, b1,
                                                                         b2, eps, total, M);
    } else {
        const uint64_t blocks = (total + kThreads - 1) / kThreads;
        selective_adam_scalar_kernel<<<(uint32_t)blocks, kThreads, 0, s>>>(param, param_grad, exp_avg, exp_avg_sq, visibility, lr,
                                                                            b1, b2, eps, total, M);
    }
    return check_launch("selective_adam_update");
}

extern "C" int gsb200_adam_multi(uint32_t n_tensors, const gsb200_adam_tensor *tensors, gsb200_stream_t stream) {
    if (n_tensors > GSB200_ADAM_MAX_TENSORS) return failf(2, "gsb200_adam_multi: at most %d tensors per call", GSB200_ADAM_MAX_TENSORS);
    if (n_tensors > 0 && tensors == nullptr) return fail(2, "gsb200_adam_multi: null table");
    AdamTable T{};
    uint64_t blocks = 0;
    for (uint32_t i = 0; i < n_tensors; ++i) {
        const gsb200_adam_tensor &h = tensors[i];
        if (h.n <= 0) continue;
        if (h.param == nullptr || h.grad == nullptr || h.exp_avg == nullptr || h.exp_avg_sq == nullptr)
            return fail(2, "gsb200_adam_multi: null pointer in the tensor table");
        const bool vec = aligned16(h.param, h.grad, h.exp_avg, h.exp_avg_sq);
        const uint64_t units = vec ? (uint64_t)(h.n / 4) : (uint64_t)h.n;
        blocks += (units + kAdamChunk - 1) / kAdamChunk;
        if (vec && units == 0) blocks += 1;  // fewer than four elements: the tail branch of block 0
        if (blocks > 0x7fffffffull) return fail(2, "gsb200_adam_multi: too many elements for one launch");
        T.t[T.n] = h;
        T.vec[T.n] = vec ? 1u : 0u;
        T.chunk_end[T.n] = (uint32_t)blocks;
        ++T.n;
    }
    if (T.n == 0) return 0;
    adam_multi_kernel<<<(uint32_t)blocks, kThreads, 0, (cudaStream_t)stream>>>(T);
    return check_launch("adam_multi");
}

extern "C" int gsb200_compute_relocation(uint32_t N, const float *opacities, const float *scales, const int32_t *ratios,
                                         const float *binoms, int n_max, float *new_opacities, float *new_scales,
                                         gsb200_stream_t stream) {
    if (N == 0) return 0;
    if (opacities == nullptr || scales == nullptr || ratios == nullptr || binoms == nullptr || new_opacities == nullptr ||
        new_scales == nullptr)
        return fail(2, "gsb200_compute_relocation: null pointer");
    if (n_max <= 0) return fail(2, "gsb200_compute_relocation: n_max must be positive");
    relocation_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, opacities, scales, ratios, binoms, n_max,
                                                                                 new_opacities, new_scales);
    return check_launch("compute_relocation");
}
