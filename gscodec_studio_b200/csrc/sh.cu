// Colour stage: real spherical harmonics (degree <= 4) -> RGB, forward and backward (sm_100a).
//
// Replaces _C.compute_sh_fwd / _C.compute_sh_bwd (gsplat/cuda/csrc/compute_sh_fwd.cu:11-37,
// compute_sh_bwd.cu:13-51, basis constants as in gsplat/cuda/include/spherical_harmonics.cuh:21-94).
//
// Design: the reference runs one thread per (element, channel) -- three threads each re-derive the
// basis and read stride-3 scalars.  Here one thread owns one element: the basis is evaluated once
// as a stream of (Y_k, grad Y_k) terms, coefficients move as 128-bit loads when the row is 16-byte
// aligned (K = 16: 192 B rows), only (degree+1)^2 bases are touched, masked rows cost one byte of
// read, and the backward writes each v_coeffs row exactly once (zeros included) so no separate
// memset pass over the [n,K,3] gradient is needed.
#include "common.cuh"
#include "quant_load.cuh"

namespace gsb200 {

// Calls f(k, Y_k, dY_k/dx, dY_k/dy, dY_k/dz) for every basis k < (deg+1)^2 at the unit direction
// (x,y,z).  Gradients are those of the polynomial extension; the caller projects them onto the
// tangent plane, which removes the dependence on the extension chosen.
template <typename F>
__device__ __forceinline__ void sh_terms(uint32_t deg, float x, float y, float z, F &&f) {
    f(0, 0.2820947917738781f, 0.f, 0.f, 0.f);
    if (deg < 1) return;
    const float a = 0.48860251190292f;
    f(1, -a * y, 0.f, -a, 0.f);
    f(2, a * z, 0.f, 0.f, a);
    f(3, -a * x, -a, 0.f, 0.f);
    if (deg < 2) return;
    const float z2 = z * z;
    const float c1 = x * x - y * y, s1 = 2.f * x * y;
    const float b = 0.5462742152960395f, d = -1.092548430592079f;
    const float Y6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    const float Y6z = 2.f * 0.9461746957575601f * z;
    f(4, b * s1, b * 2.f * y, b * 2.f * x, 0.f);
    f(5, d * z * y, 0.f, d * z, d * y);
    f(6, Y6, 0.f, 0.f, Y6z);
    f(7, d * z * x, d * z, 0.f, d * x);
    f(8, b * c1, b * 2.f * x, -b * 2.f * y, 0.f);
    if (deg < 3) return;
    const float c2 = x * c1 - y * s1, s2 = x * s1 + y * c1;
    const float h = -0.5900435899266435f, p = 1.445305721320277f;
    const float q = -2.285228997322329f;
    const float t1 = q * z2 + 0.4570457994644658f;
    const float Y12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float Y12z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    f(9, h * s2, h * 3.f * s1, h * 3.f * c1, 0.f);
    f(10, p * z * s1, p * z * 2.f * y, p * z * 2.f * x, p * s1);
    f(11, t1 * y, 0.f, t1, 2.f * q * z * y);
    f(12, Y12, 0.f, 0.f, Y12z);
    f(13, t1 * x, t1, 0.f, 2.f * q * z * x);
    f(14, p * z * c1, p * z * 2.f * x, -p * z * 2.f * y, p * c1);
    f(15, h * c2, h * 3.f * c1, -h * 3.f * s1, 0.f);
    if (deg < 4) return;
    const float c3 = x * c2 - y * s2, s3 = x * s2 + y * c2;
    const float w = 0.6258357354491763f, m = -1.770130769779931f;
    const float A = 3.31161143515146f, D = -4.683325804901025f, E = 2.007139630671868f;
    const float u1 = A * z2 - 0.47308734787878f;
    const float u0 = z * (D * z2 + E);
    const float u0z = 3.f * D * z2 + E;
    const float Fc = 1.984313483298443f, Gc = 1.006230589874905f;
    f(16, w * s3, w * 4.f * s2, w * 4.f * c2, 0.f);
    f(17, m * z * s2, m * z * 3.f * s1, m * z * 3.f * c1, m * s2);
    f(18, u1 * s1, u1 * 2.f * y, u1 * 2.f * x, 2.f * A * z * s1);
    f(19, u0 * y, 0.f, u0, u0z * y);
    f(20, Fc * z * Y12 - Gc * Y6, 0.f, 0.f, Fc * (Y12 + z * Y12z) - Gc * Y6z);
    f(21, u0 * x, u0, 0.f, u0z * x);
    f(22, u1 * c1, u1 * 2.f * x, -u1 * 2.f * y, 2.f * A * z * c1);
    f(23, m * z * c2, m * z * 3.f * c1, -m * z * 3.f * s1, m * c2);
    f(24, w * c3, w * 4.f * c2, -w * 4.f * s2, 0.f);
}

constexpr int kMaxBases = 25;

template <int DEG>
__global__ void __launch_bounds__(kThreads)
sh_fwd_kernel(uint32_t n, uint32_t K, const float *__restrict__ dirs, const float *__restrict__ coeffs,
              const uint8_t *__restrict__ masks, float *__restrict__ colors, bool vec) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float r = 0.f, g = 0.f, b = 0.f;
    if (masks == nullptr || masks[i]) {
        constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
        float c[nb * 3 + 3];
        {
            const float *row = coeffs + (size_t)i * K * 3;
            if (vec) {
#pragma unroll
                for (uint32_t k = 0; k < nb * 3; k += 4) {
                    float4 v = ldg4(row + k);
                    c[k] = v.x; c[k + 1] = v.y; c[k + 2] = v.z; c[k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (uint32_t k = 0; k < nb * 3; ++k) c[k] = __ldg(row + k);
            }
        }
        float x = 0.f, y = 0.f, z = 0.f;
        if (DEG > 0) {
            x = dirs[(size_t)i * 3]; y = dirs[(size_t)i * 3 + 1]; z = dirs[(size_t)i * 3 + 2];
            const float inorm = rsqrtf(x * x + y * y + z * z);
            x *= inorm; y *= inorm; z *= inorm;
        }
        sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
            r += Y * c[k * 3]; g += Y * c[k * 3 + 1]; b += Y * c[k * 3 + 2];
        });
    }
    colors[(size_t)i * 3] = r; colors[(size_t)i * 3 + 1] = g; colors[(size_t)i * 3 + 2] = b;
}

template <int DEG>
__global__ void __launch_bounds__(kThreads)
sh_bwd_kernel(uint32_t n, uint32_t K, const float *__restrict__ dirs, const float *__restrict__ coeffs,
              const uint8_t *__restrict__ masks, const float *__restrict__ v_colors, float *__restrict__ v_coeffs,
              float *__restrict__ v_dirs, bool vec) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    float *vrow = v_coeffs + (size_t)i * K * 3;
    const uint32_t nfl = K * 3;
    const bool on = masks == nullptr || masks[i];
    float vc[nb * 3 + 3];
#pragma unroll
    for (uint32_t k = 0; k < nb * 3 + 3; ++k) vc[k] = 0.f;
    float vdx = 0.f, vdy = 0.f, vdz = 0.f;
    if (on) {
        const float vr = v_colors[(size_t)i * 3], vg = v_colors[(size_t)i * 3 + 1], vb = v_colors[(size_t)i * 3 + 2];
        float x = 0.f, y = 0.f, z = 0.f, inorm = 1.f;
        if (DEG > 0) {
            x = dirs[(size_t)i * 3]; y = dirs[(size_t)i * 3 + 1]; z = dirs[(size_t)i * 3 + 2];
            inorm = rsqrtf(x * x + y * y + z * z);
            x *= inorm; y *= inorm; z *= inorm;
        }
        if (v_dirs != nullptr && DEG > 0) {
            float c[nb * 3 + 3];
            const float *row = coeffs + (size_t)i * K * 3;
            if (vec) {
#pragma unroll
                for (uint32_t k = 0; k < nb * 3; k += 4) {
                    float4 v = ldg4(row + k);
                    c[k] = v.x; c[k + 1] = v.y; c[k + 2] = v.z; c[k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (uint32_t k = 0; k < nb * 3; ++k) c[k] = __ldg(row + k);
            }
            float gx = 0.f, gy = 0.f, gz = 0.f;
            sh_terms(DEG, x, y, z, [&](int k, float Y, float dx, float dy, float dz) {
                vc[k * 3] = Y * vr; vc[k * 3 + 1] = Y * vg; vc[k * 3 + 2] = Y * vb;
                const float wk = c[k * 3] * vr + c[k * 3 + 1] * vg + c[k * 3 + 2] * vb;
                gx += dx * wk; gy += dy * wk; gz += dz * wk;
            });
            const float dotp = gx * x + gy * y + gz * z;
            vdx = (gx - dotp * x) * inorm; vdy = (gy - dotp * y) * inorm; vdz = (gz - dotp * z) * inorm;
        } else {
            sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
                vc[k * 3] = Y * vr; vc[k * 3 + 1] = Y * vg; vc[k * 3 + 2] = Y * vb;
            });
        }
    }
    // write the whole row once: active bases, then zeros for the unused tail
    if (vec) {  // nb*3 % 4 == 0 here
#pragma unroll
        for (uint32_t k = 0; k + 3 < nb * 3; k += 4) stg4(vrow + k, make_float4(vc[k], vc[k + 1], vc[k + 2], vc[k + 3]));
        for (uint32_t k = nb * 3; k < nfl; k += 4) stg4(vrow + k, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
#pragma unroll
        for (uint32_t k = 0; k < nb * 3; ++k) vrow[k] = vc[k];
        for (uint32_t k = nb * 3; k < nfl; ++k) vrow[k] = 0.f;
    }
    if (v_dirs != nullptr) {
        v_dirs[(size_t)i * 3] = vdx; v_dirs[(size_t)i * 3 + 1] = vdy; v_dirs[(size_t)i * 3 + 2] = vdz;
    }
}

// ------------------------------------------------------------------------------------------------
// View-dependent colour of every (camera, Gaussian) pair in one pass: dirs = mean - camera position,
// SH -> RGB, + 0.5, clamp at 0 (gsplat/rendering.py:370-392 does this with a batched inverse, a
// subtraction, an expand, the SH op, an add and a clamp -- six kernels and an [C,N,K,3] broadcast whose
// backward is a 193 MB reduction).  One thread owns one Gaussian and loops over the cameras: the
// coefficient row is read once, the per-camera gradients are summed in registers and each v_coeffs row
// is written exactly once.
// Where the coefficient rows live.  Single GPU: one [N, K, 3] array (n_block == 0).  View-parallel multi-GPU: rank r
// owns rows [r * n_block, (r + 1) * n_block) of the gathered set and publishes them in peer-accessible memory;
// base[r] is rank r's block as mapped into THIS process (NVLink peer mapping), so a kernel pulls the rows it
// needs -- only the visible ones -- straight from the owner instead of receiving an all-gathered copy.
constexpr int kMaxShards = GSB200_MAX_SHARDS;
struct CoeffShards {
    const float *base[kMaxShards];
    uint32_t n_block;
};

__device__ __forceinline__ const float *coeff_row(const CoeffShards &sh, uint32_t g, uint32_t nfl) {
    if (sh.n_block == 0) return sh.base[0] + (size_t)g * nfl;
    const uint32_t r = g / sh.n_block;
    return sh.base[r] + (size_t)(g - r * sh.n_block) * nfl;
}

// Where the coefficient GRADIENT rows go.  Single GPU (rows[0] == nullptr): the dense local array v_coeffs.
// Multi-GPU: rank `src` PUSHES the row of gathered Gaussian g into slot [src][g % n_block] of the staging array of
// g's owner (remote stores over NVLink are posted, nobody waits on a round trip), and a 0/1 flag per row next to
// it, so that the owner's reduction afterwards reads local memory only and skips rows nobody wrote.
struct PushTargets {
    float *rows[kMaxShards];     // owner o's staging array [world, n_block, nfl]
    int32_t *flags[kMaxShards];  // owner o's flags [world, n_block]
    uint32_t src;
};

__device__ __forceinline__ size_t push_slot(const PushTargets &t, uint32_t g, uint32_t n_block, uint32_t &owner) {
    owner = g / n_block;
    return (size_t)t.src * n_block + (g - owner * n_block);
}

constexpr int kCoopThreads = 128;

// Cooperative row staging of the COOP kernels: a warp moves the rows of its visible Gaussians between global
// (possibly remote) memory and shared memory with fully coalesced 16-byte accesses -- L consecutive lanes cover one
// row, 32 / L rows per instruction -- instead of every lane walking its own row at a 4 * nfl byte stride, which
// over NVLink fetches every 32-byte sector twice.  Row stride in shared memory is an odd number of 16-byte units,
// so the per-lane row reads afterwards are conflict-free.
template <int DEG>
struct CoopRows {
    static constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    static constexpr uint32_t L = nb * 3 / 4;            // float4 per row
    static constexpr uint32_t RPI = 32 / L;              // rows per warp-wide instruction
    static constexpr uint32_t U = (L & 1) ? L : L + 1;   // padded row stride, in float4
    static constexpr uint32_t kWarpFloat4 = 32 * U;
};

template <int DEG>
__device__ __forceinline__ uint32_t coop_list_visible(bool any, uint8_t *list) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mask = __ballot_sync(0xffffffffu, any);
    if (any) list[__popc(mask & ((1u << lane) - 1u))] = (uint8_t)lane;
    __syncwarp();
    return __popc(mask);
}

template <int DEG>
__device__ __forceinline__ void coop_pull_rows(const CoeffShards &coeffs, uint32_t g0, uint32_t nfl, uint32_t nvis,
                                               const uint8_t *list, float4 *rows, float *row_cache) {
    using R = CoopRows<DEG>;
    const uint32_t lane = threadIdx.x & 31, sub = lane / R::L, kk = lane - sub * R::L;
#pragma unroll 4
    for (uint32_t q0 = 0; q0 < nvis; q0 += R::RPI) {
        const uint32_t q = q0 + sub;
        if (sub < R::RPI && q < nvis) {
            const uint32_t j = list[q];
            const float4 v = ldg4(coeff_row(coeffs, g0 + j, nfl) + kk * 4);
            rows[j * R::U + kk] = v;
            if (row_cache != nullptr) stg4(row_cache + (size_t)(g0 + j) * nfl + kk * 4, v);  // local copy for the backward
        }
    }
    __syncwarp();
}

template <int DEG>
__device__ __forceinline__ void coop_push_rows(const PushTargets &t, uint32_t g0, uint32_t n_block, uint32_t nfl, uint32_t nvis,
                                               const uint8_t *list, const float4 *rows) {
    using R = CoopRows<DEG>;
    const uint32_t lane = threadIdx.x & 31, sub = lane / R::L, kk = lane - sub * R::L;
#pragma unroll 4
    for (uint32_t q0 = 0; q0 < nvis; q0 += R::RPI) {
        const uint32_t q = q0 + sub;
        if (sub < R::RPI && q < nvis) {
            const uint32_t j = list[q];
            uint32_t owner;
            const size_t slot = push_slot(t, g0 + j, n_block, owner);
            stg4(t.rows[owner] + slot * nfl + kk * 4, rows[j * R::U + kk]);
        }
    }
}

template <int DEG, bool COOP>
__global__ void __launch_bounds__(COOP ? kCoopThreads : kThreads)
sh_view_fwd_kernel(uint32_t C, uint32_t N, uint32_t K, const float *__restrict__ means, const float *__restrict__ campos,
                   const CoeffShards coeffs, const int32_t *__restrict__ radii, float *__restrict__ colors,
                   float *__restrict__ row_cache, bool vec) {
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (!COOP && g >= N) return;
    bool any = false;
    if (g < N)
        for (uint32_t c = 0; c < C; ++c) any |= radii[(size_t)c * N + g] > 0;
    float cf[nb * 3 + 3];
    float mx = 0.f, my = 0.f, mz = 0.f;
    if constexpr (COOP) {
        using R = CoopRows<DEG>;
        __shared__ float4 s_rows[(kCoopThreads / 32) * R::kWarpFloat4];
        __shared__ uint8_t s_list[kCoopThreads];
        float4 *rows = s_rows + (threadIdx.x >> 5) * R::kWarpFloat4;
        uint8_t *list = s_list + (threadIdx.x & ~31u);
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t nvis = coop_list_visible<DEG>(any, list);
        coop_pull_rows<DEG>(coeffs, g - lane, K * 3, nvis, list, rows, row_cache);
        if (any) {
#pragma unroll
            for (uint32_t k = 0; k < R::L; ++k) {
                const float4 v = rows[lane * R::U + k];
                cf[4 * k] = v.x; cf[4 * k + 1] = v.y; cf[4 * k + 2] = v.z; cf[4 * k + 3] = v.w;
            }
        }
        if (g >= N) return;
    } else if (any) {
        const float *row = coeff_row(coeffs, g, K * 3);
        float *keep = row_cache != nullptr ? row_cache + (size_t)g * K * 3 : nullptr;
        if (vec) {
#pragma unroll
            for (uint32_t k = 0; k < nb * 3; k += 4) {
                const float4 v = ldg4(row + k);
                cf[k] = v.x; cf[k + 1] = v.y; cf[k + 2] = v.z; cf[k + 3] = v.w;
                if (keep != nullptr) stg4(keep + k, v);
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < nb * 3; ++k) {
                cf[k] = __ldg(row + k);
                if (keep != nullptr) keep[k] = cf[k];
            }
        }
    }
    if (any) { mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2]; }
    for (uint32_t c = 0; c < C; ++c) {
        const size_t idx = (size_t)c * N + g;
        float r = 0.f, gg = 0.f, b = 0.f;
        if (radii[idx] > 0) {
            float x = mx - campos[c * 3], y = my - campos[c * 3 + 1], z = mz - campos[c * 3 + 2];
            if (DEG > 0) {
                const float inorm = rsqrtf(x * x + y * y + z * z);
                x *= inorm; y *= inorm; z *= inorm;
            }
            sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
                r += Y * cf[k * 3]; gg += Y * cf[k * 3 + 1]; b += Y * cf[k * 3 + 2];
            });
            r = fmaxf(r + 0.5f, 0.f); gg = fmaxf(gg + 0.5f, 0.f); b = fmaxf(b + 0.5f, 0.f);
        }
        colors[idx * 3] = r; colors[idx * 3 + 1] = gg; colors[idx * 3 + 2] = b;
    }
}

template <int DEG, bool COOP>
__global__ void __launch_bounds__(COOP ? kCoopThreads : kThreads)
sh_view_bwd_kernel(uint32_t C, uint32_t N, uint32_t K, const float *__restrict__ means, const float *__restrict__ campos,
                   const CoeffShards coeffs, const int32_t *__restrict__ radii, const float *__restrict__ v_colors,
                   float *__restrict__ v_coeffs, const PushTargets push, uint32_t push_block, float *__restrict__ v_means,
                   float *__restrict__ v_campos, bool vec) {
    __shared__ float s_red[kThreads / 32][3];
    using R = CoopRows<COOP ? DEG : 3>;
    __shared__ float4 s_rows[COOP ? (kCoopThreads / 32) * R::kWarpFloat4 : 1];
    __shared__ uint8_t s_list[COOP ? kCoopThreads : 1];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    const uint32_t nfl = K * 3;
    float vc[nb * 3 + 3];
#pragma unroll
    for (uint32_t k = 0; k < nb * 3 + 3; ++k) vc[k] = 0.f;
    float vmx = 0.f, vmy = 0.f, vmz = 0.f;
    bool any = false;
    if (g < N)
        for (uint32_t c = 0; c < C; ++c) any |= radii[(size_t)c * N + g] > 0;
    float cf[nb * 3 + 3];
    float mx = 0.f, my = 0.f, mz = 0.f;
    float4 *rows = s_rows + (COOP ? (threadIdx.x >> 5) * R::kWarpFloat4 : 0);
    uint8_t *list = s_list + (COOP ? (threadIdx.x & ~31u) : 0);
    uint32_t nvis = 0;
    if constexpr (COOP) {
        nvis = coop_list_visible<DEG>(any, list);
        coop_pull_rows<DEG>(coeffs, g - lane, nfl, nvis, list, rows, nullptr);
        if (any) {
#pragma unroll
            for (uint32_t k = 0; k < R::L; ++k) {
                const float4 v = rows[lane * R::U + k];
                cf[4 * k] = v.x; cf[4 * k + 1] = v.y; cf[4 * k + 2] = v.z; cf[4 * k + 3] = v.w;
            }
        }
        __syncwarp();  // rows[] is reused for the gradient rows below
    } else if (any) {
        const float *row = coeff_row(coeffs, g, nfl);
        if (vec) {
#pragma unroll
            for (uint32_t k = 0; k < nb * 3; k += 4) {
                const float4 v = ldg4(row + k);
                cf[k] = v.x; cf[k + 1] = v.y; cf[k + 2] = v.z; cf[k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < nb * 3; ++k) cf[k] = __ldg(row + k);
        }
    }
    if (any) { mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2]; }
    for (uint32_t c = 0; c < C; ++c) {
        float dxv = 0.f, dyv = 0.f, dzv = 0.f;
        if (any && radii[(size_t)c * N + g] > 0) {
            const size_t idx = (size_t)c * N + g;
            float x = mx - campos[c * 3], y = my - campos[c * 3 + 1], z = mz - campos[c * 3 + 2];
            float inorm = 1.f;
            if (DEG > 0) {
                inorm = rsqrtf(x * x + y * y + z * z);
                x *= inorm; y *= inorm; z *= inorm;
            }
            // recompute the pre-clamp colour for the clamp gate (torch: gradient passes where x + 0.5 >= 0)
            float r = 0.f, gg = 0.f, b = 0.f;
            sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
                r += Y * cf[k * 3]; gg += Y * cf[k * 3 + 1]; b += Y * cf[k * 3 + 2];
            });
            const float vr = (r + 0.5f >= 0.f) ? v_colors[idx * 3] : 0.f;
            const float vg = (gg + 0.5f >= 0.f) ? v_colors[idx * 3 + 1] : 0.f;
            const float vb = (b + 0.5f >= 0.f) ? v_colors[idx * 3 + 2] : 0.f;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            sh_terms(DEG, x, y, z, [&](int k, float Y, float dx, float dy, float dz) {
                vc[k * 3] += Y * vr; vc[k * 3 + 1] += Y * vg; vc[k * 3 + 2] += Y * vb;
                const float wk = cf[k * 3] * vr + cf[k * 3 + 1] * vg + cf[k * 3 + 2] * vb;
                gx += dx * wk; gy += dy * wk; gz += dz * wk;
            });
            if (DEG > 0) {
                const float dotp = gx * x + gy * y + gz * z;
                dxv = (gx - dotp * x) * inorm; dyv = (gy - dotp * y) * inorm; dzv = (gz - dotp * z) * inorm;
                vmx += dxv; vmy += dyv; vmz += dzv;
            }
        }
        if (v_campos != nullptr) {  // d dirs / d campos = -1: block reduction, one atomic set per block and camera
            const float sx = warp_sum(dxv), sy = warp_sum(dyv), sz = warp_sum(dzv);
            __syncthreads();
            if (lane == 0) { s_red[threadIdx.x >> 5][0] = sx; s_red[threadIdx.x >> 5][1] = sy; s_red[threadIdx.x >> 5][2] = sz; }
            __syncthreads();
            if (threadIdx.x < 3) {
                float t = 0.f;
                for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += s_red[w][threadIdx.x];
                if (t != 0.f) atomicAdd(v_campos + c * 3 + threadIdx.x, -t);
            }
        }
    }
    const bool pushing = push.rows[0] != nullptr;
    if constexpr (COOP) {  // always pushing; nb * 3 == nfl
        if (any) {
#pragma unroll
            for (uint32_t k = 0; k < R::L; ++k)
                rows[lane * R::U + k] = make_float4(vc[4 * k], vc[4 * k + 1], vc[4 * k + 2], vc[4 * k + 3]);
        }
        __syncwarp();
        coop_push_rows<DEG>(push, g - lane, push_block, nfl, nvis, list, rows);
    }
    if (g >= N) return;
    float *vrow = nullptr;
    if (pushing) {  // only visible rows travel; the flag tells the owner which
        uint32_t owner;
        const size_t slot = push_slot(push, g, push_block, owner);
        push.flags[owner][slot] = any ? 1 : 0;
        if (!COOP && any) vrow = push.rows[owner] + slot * nfl;
    } else if (v_coeffs != nullptr) {
        vrow = v_coeffs + (size_t)g * nfl;
    }
    if (vrow != nullptr) {
        if (vec) {
#pragma unroll
            for (uint32_t k = 0; k + 3 < nb * 3; k += 4) stg4(vrow + k, make_float4(vc[k], vc[k + 1], vc[k + 2], vc[k + 3]));
            for (uint32_t k = nb * 3; k < nfl; k += 4) stg4(vrow + k, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
#pragma unroll
            for (uint32_t k = 0; k < nb * 3; ++k) vrow[k] = vc[k];
            for (uint32_t k = nb * 3; k < nfl; ++k) vrow[k] = 0.f;
        }
    }
    if (v_means != nullptr) {
        v_means[(size_t)g * 3] = vmx; v_means[(size_t)g * 3 + 1] = vmy; v_means[(size_t)g * 3 + 2] = vmz;
    }
}

// Owner-side half of the sparse reduce-scatter: sums, in source-rank order (reproducible), the rows the ranks pushed
// into this rank's staging array [world, n_block, row] -- local memory only, rows nobody flagged are skipped.
__global__ void __launch_bounds__(kThreads)
shard_reduce_rows_kernel(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_vec4, const float *__restrict__ staging,
                         const int32_t *__restrict__ flags, float *__restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (size_t)n_rows * row_vec4) return;
    const uint32_t i = idx / row_vec4, k = idx - (size_t)i * row_vec4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t p0 = 0; p0 < world; p0 += 4) {
        float4 v[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const uint32_t p = p0 + u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < world && __ldg(flags + (size_t)p * n_block + i) != 0)
                v[u] = ldg4(staging + (((size_t)p * n_block + i) * row_vec4 + k) * 4);
        }
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    stg4(out + idx * 4, acc);
}

// Same reduction, written to TWO tensors: the first `split` floats of a row go to out0 [n_rows, split], the rest to
// out1 [n_rows, row - split] (the trainers' sh0 / shN gradients; a (K-1)*12-byte row is not 16-byte aligned -> scalar stores)
__global__ void __launch_bounds__(kThreads)
shard_reduce_rows_split_kernel(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_vec4, uint32_t split,
                               const float *__restrict__ staging, const int32_t *__restrict__ flags, float *__restrict__ out0,
                               float *__restrict__ out1) {
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (size_t)n_rows * row_vec4) return;
    const uint32_t i = idx / row_vec4, k = idx - (size_t)i * row_vec4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t p = 0; p < world; ++p)
        if (__ldg(flags + (size_t)p * n_block + i) != 0) {
            const float4 v = ldg4(staging + (((size_t)p * n_block + i) * row_vec4 + k) * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
    const uint32_t rest = row_vec4 * 4 - split;
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
        const uint32_t f = k * 4 + u;
        if (f < split) out0[(size_t)i * split + f] = vals[u];
        else out1[(size_t)i * rest + (f - split)] = vals[u];
    }
}

// All-gather by pull: every rank copies the published block of every peer (NVLink peer mappings) into its own
// gathered array.  One launch, 64 bytes in flight per thread, no staging copies, no protocol overhead.
struct BlockShards {
    const float *base[kMaxShards];
};

__global__ void __launch_bounds__(kThreads)
shard_gather_rows_kernel(uint64_t block_vec4, const BlockShards blocks, float *__restrict__ out) {
    const float *src = blocks.base[blockIdx.y];
    float *dst = out + (size_t)blockIdx.y * block_vec4 * 4;
    const uint64_t i0 = ((uint64_t)blockIdx.x * kThreads) * 4 + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
        if (i0 + u * kThreads < block_vec4) v[u] = ldg4(src + (i0 + u * kThreads) * 4);
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
        if (i0 + u * kThreads < block_vec4) stg4(dst + (i0 + u * kThreads) * 4, v[u]);
}


// ---- geometry exchange over peer memory: gather-and-split forward, push + owner-side reduce backward ------------------------
// The published geometry block of every rank is [n_block, w] (means | quats | scales | opacity, or means | covars |
// opacity).  The forward pulls all blocks and writes the attribute tensors of the gathered set directly (no gathered
// [N, w] intermediate, no column-split copies); the backward pushes each owner's rows of the attribute gradients into that
// owner's staging slot [source rank][row][w] (posted NVLink stores, 4-byte coalesced), and after a barrier every owner sums
// its world slots in rank order -- replaces torch.cat + ncclReduceScatter on a padded [N, w] tensor.
struct GeoParts {
    float *ptr[4];
    uint32_t width[4], col0[4];  // part k holds columns [col0, col0 + width)
    uint32_t n_parts, w;
};
struct PeerRows {
    float *base[kMaxShards];
};

// Block / staging-slot layout: the parts one after the other, part k = its [n_block, width_k] rows, dense, at float offset
// n_block * col0_k (a multiple of 4: n_block is).  Every kernel moves 16-byte vectors that are contiguous on BOTH sides
// (several in flight per thread -- NVLink latency is ~1.8 k cycles); nothing is scattered element-wise.
constexpr int kGeoUnroll = 4;

__device__ __forceinline__ void geo_vec_locate(const GeoParts &P, uint32_t n_block, uint64_t e0, uint32_t &part, uint64_t &e) {
    part = 0;
#pragma unroll
    for (uint32_t k = 1; k < 4; ++k)
        if (k < P.n_parts && e0 >= (uint64_t)n_block * P.col0[k]) part = k;
    e = e0 - (uint64_t)n_block * P.col0[part];
}

__device__ __forceinline__ float4 load4_maybe(const float *p, bool vec) {
    return vec ? ldg4(p) : make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// gather: peer p's block (float4 loads over NVLink) -> rows p * n_block.. of the attribute tensors
__global__ void __launch_bounds__(kThreads)
shard_gather_geo_kernel(uint32_t n_block, const BlockShards blocks, const __grid_constant__ GeoParts out) {
    const uint32_t p = blockIdx.y;
    const uint64_t n_vec = (uint64_t)n_block * out.w / 4;
    const uint64_t i0 = (uint64_t)blockIdx.x * kThreads * kGeoUnroll + threadIdx.x;
    float4 v[kGeoUnroll];
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u)
        if (i0 + (uint64_t)u * kThreads < n_vec) v[u] = ldg4(blocks.base[p] + (i0 + (uint64_t)u * kThreads) * 4);
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u) {
        const uint64_t i = i0 + (uint64_t)u * kThreads;
        if (i >= n_vec) continue;
        uint32_t part;
        uint64_t e;
        geo_vec_locate(out, n_block, i * 4, part, e);
        stg4(out.ptr[part] + (size_t)p * n_block * out.width[part] + e, v[u]);
    }
}

// push: rows p * n_block.. of the gradient tensors -> owner p's staging slot `rank` (float4 stores over NVLink)
__global__ void __launch_bounds__(kThreads)
shard_push_geo_kernel(uint32_t n_block, uint32_t rank, uint64_t slot_floats, bool vec, const __grid_constant__ GeoParts grads,
                      const PeerRows peers) {
    const uint32_t p = blockIdx.y;
    const uint64_t n_vec = (uint64_t)n_block * grads.w / 4;
    const uint64_t i0 = (uint64_t)blockIdx.x * kThreads * kGeoUnroll + threadIdx.x;
    float *dst = peers.base[p] + (size_t)rank * slot_floats;
    float4 v[kGeoUnroll];
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u) {
        const uint64_t i = i0 + (uint64_t)u * kThreads;
        if (i >= n_vec) continue;
        uint32_t part;
        uint64_t e;
        geo_vec_locate(grads, n_block, i * 4, part, e);
        v[u] = grads.ptr[part] ? load4_maybe(grads.ptr[part] + (size_t)p * n_block * grads.width[part] + e, vec)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u)
        if (i0 + (uint64_t)u * kThreads < n_vec) stg4(dst + (i0 + (uint64_t)u * kThreads) * 4, v[u]);
}

// publish: this rank's attribute tensors -> its geometry block (rows >= n_local of every part get the inert pad row) and
// its coefficient tensor(s) -> its coefficient block [n_local, row_floats] (the trainers' sh0 [n,1,3] next to shN
// [n,K-1,3]), one launch, every store a coalesced 16-byte vector (replaces two strided torch copies + torch.cat).
// Blocks [0, n_geo_blocks) handle the geometry vectors, the rest the coefficient vectors.
__global__ void __launch_bounds__(kThreads)
shard_publish_kernel(uint32_t n_local, uint32_t n_block, uint32_t n_geo_blocks, const __grid_constant__ GeoParts parts,
                     const float *__restrict__ pad_row, float *__restrict__ geo_block, const float *__restrict__ first,
                     const float *__restrict__ rest, uint32_t first_floats, uint32_t row_floats, float *__restrict__ coeff_block) {
    if (blockIdx.x < n_geo_blocks) {
        const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
        if (i >= (uint64_t)n_block * parts.w / 4) return;
        uint32_t part;
        uint64_t e;
        geo_vec_locate(parts, n_block, i * 4, part, e);
        const uint32_t wk = parts.width[part];
        const uint64_t n_own = (uint64_t)n_local * wk;
        const float *src = parts.ptr[part];
        float vals[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            vals[u] = e + u < n_own ? __ldg(src + e + u) : (pad_row ? pad_row[parts.col0[part] + (uint32_t)((e + u) % wk)] : 0.f);
        stg4(geo_block + i * 4, make_float4(vals[0], vals[1], vals[2], vals[3]));
        return;
    }
    if (first == nullptr) return;
    const uint32_t vec_per_row = row_floats / 4, rest_floats = row_floats - first_floats;
    const uint64_t i = (uint64_t)(blockIdx.x - n_geo_blocks) * kThreads + threadIdx.x;
    if (i >= (uint64_t)n_local * vec_per_row) return;
    const uint64_t row = i / vec_per_row;
    const uint32_t e0 = (uint32_t)(i % vec_per_row) * 4;
    float vals[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t e = e0 + u;
        vals[u] = e < first_floats ? __ldg(first + row * first_floats + e) : __ldg(rest + row * rest_floats + (e - first_floats));
    }
    stg4(coeff_block + i * 4, make_float4(vals[0], vals[1], vals[2], vals[3]));
}

// Stream-ordered barrier over the ranks of one node: every rank owns a signal pad (kMaxShards words, in the symmetric
// allocation); thread p stores this barrier's epoch into peer p's pad (release, system scope, after a system fence that
// orders the stream's earlier writes) and spins until peer p's epoch has arrived in the own pad (acquire).  Epochs only
// grow, so nothing is ever reset.  A peer that does not arrive within ~20 s traps the context instead of hanging the GPU.
struct PeerPads {
    uint32_t *base[kMaxShards];
};
__global__ void peer_barrier_kernel(uint32_t rank, uint32_t epoch, const PeerPads pads) {
    const uint32_t p = threadIdx.x;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pads.base[p] + rank), "r"(epoch) : "memory");
    const uint32_t *mine = pads.base[rank] + p;
    uint64_t t0 = 0;
    for (uint32_t polls = 0;; ++polls) {
        uint32_t seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - epoch) >= 0) break;
        if ((polls & 1023u) == 1023u) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) __trap();
        }
    }
    __threadfence_system();
}

// owner-side reduce: sum of the world slots of the local staging array, in rank order -> the local gradient tensors
__global__ void __launch_bounds__(kThreads)
shard_reduce_geo_kernel(uint32_t world, uint32_t n_block, uint32_t n_rows, uint64_t slot_floats, const float *__restrict__ staging,
                        const __grid_constant__ GeoParts out) {
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (uint64_t)n_block * out.w / 4) return;
    uint32_t part;
    uint64_t e;
    geo_vec_locate(out, n_block, i * 4, part, e);
    const uint64_t n_own = (uint64_t)n_rows * out.width[part];
    if (out.ptr[part] == nullptr || e >= n_own) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t p = 0; p < world; ++p) {
        const float4 v = ldg4(staging + (size_t)p * slot_floats + i * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float *dst = out.ptr[part] + e;
    if (e + 4 <= n_own) {
        stg4(dst, acc);
    } else {
        const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
        for (uint32_t u = 0; e + u < n_own; ++u) dst[u] = vals[u];
    }
}

// ------------------------------------------------------------------------------------------------
// Split-coefficient form of the fused view-colour stage.  The trainers keep the DC band and the higher bands as
// two parameters, sh0 [N,1,3] and shN [N,K-1,3], and concatenate them every step only to hand one [N,K,3] tensor
// to rasterization() (examples/simple_trainer.py:776: a 193 MB copy forward and two strided copies backward at
// 1 M Gaussians).  These kernels read / write the two arrays in place: the visible rows are read by their owner
// thread (4-byte loads -- a (K-1)*12-byte row is not 16-byte aligned), the gradient rows of a block -- one
// contiguous [kSplitThreads, (K-1)*3] slab of v_shN, zeros for invisible Gaussians and unused bands included --
// are staged in shared memory and written with coalesced 16-byte stores, exactly once.
constexpr int kSplitThreads = 128;

// Where a (view, Gaussian) pair's visibility, camera centre, colour and colour gradient live.  LocalViews: the viewer
// evaluates its own views over all N Gaussians (single GPU, or gathered coefficients).  PeerViews: the OWNER of a shard
// evaluates its rows for every view of the job -- masks and colour gradients were pushed into local staging by the
// viewers, the colours are stored straight into the viewers' colour tensors over NVLink (12 bytes per visible pair
// instead of the 16x3-float coefficient row travelling the other way).
struct LocalViews {
    const float *campos;
    const int32_t *radii;
    float *colors;
    const float *v_colors;
    uint32_t N;
    static constexpr bool kWriteInvisible = true;  // the local colour tensor is torch.empty: invisible rows are zeroed
    __device__ __forceinline__ bool visible(uint32_t c, uint32_t g) const { return radii[(size_t)c * N + g] > 0; }
    __device__ __forceinline__ const float *cam(uint32_t c) const { return campos + c * 3; }
    __device__ __forceinline__ float *color(uint32_t c, uint32_t g) const { return colors + ((size_t)c * N + g) * 3; }
    __device__ __forceinline__ const float *v_color(uint32_t c, uint32_t g) const { return v_colors + ((size_t)c * N + g) * 3; }
};
constexpr int kMaxViews = GSB200_MAX_VIEWS;
struct PeerViews {
    const float *campos;       // local [V,3]: every view's camera centre (pushed by the viewers)
    const int32_t *masks;      // local [V, n_block]: radii of this owner's rows in view v (pushed by the viewers)
    const float *v_colors;     // local [V, n_block, 3]: colour gradients of this owner's rows (pushed by the viewers)
    float *colors[kMaxViews];  // view v: this owner's [n_block, 3] slice of the viewer's colour tensor (peer memory)
    uint32_t n_block;
    static constexpr bool kWriteInvisible = false;  // nothing travels for an invisible pair (the blend never reads it)
    __device__ __forceinline__ bool visible(uint32_t c, uint32_t g) const { return masks[(size_t)c * n_block + g] > 0; }
    __device__ __forceinline__ const float *cam(uint32_t c) const { return campos + c * 3; }
    __device__ __forceinline__ float *color(uint32_t c, uint32_t g) const { return colors[c] + (size_t)g * 3; }
    __device__ __forceinline__ const float *v_color(uint32_t c, uint32_t g) const {
        return v_colors + ((size_t)c * n_block + g) * 3;
    }
};

template <int DEG, typename VIEWS>
__global__ void __launch_bounds__(kSplitThreads)
sh_view_split_fwd_kernel(uint32_t C, uint32_t N, uint32_t RF, const float *__restrict__ means, const float *__restrict__ sh0,
                         const float *__restrict__ shN, const __grid_constant__ VIEWS views, const __grid_constant__ QuantLoad q0) {
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    const uint32_t g = blockIdx.x * kSplitThreads + threadIdx.x;
    if (g >= N) return;
    bool any = false;
    for (uint32_t c = 0; c < C; ++c) any |= views.visible(c, g);
    float cf[nb * 3];
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (any) {
        cf[0] = __ldg(sh0 + (size_t)g * 3); cf[1] = __ldg(sh0 + (size_t)g * 3 + 1); cf[2] = __ldg(sh0 + (size_t)g * 3 + 2);
        if (q0.on) {  // sh0 is the RAW parameter: fake-quantise it here (quant_load.cuh)
            const float raw[3] = {cf[0], cf[1], cf[2]};
            float qv[3];
            quant_load<3>(q0, (int64_t)g * 3, raw, qv);
            cf[0] = qv[0]; cf[1] = qv[1]; cf[2] = qv[2];
        }
        if (DEG > 0) {
            const float *row = shN + (size_t)g * RF;
#pragma unroll
            for (uint32_t k = 0; k + 3 < nb * 3; ++k) cf[3 + k] = __ldg(row + k);
        }
        mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2];
    }
    for (uint32_t c = 0; c < C; ++c) {
        float r = 0.f, gg = 0.f, b = 0.f;
        const bool vis = any && views.visible(c, g);
        if (vis) {
            const float *cam = views.cam(c);
            float x = mx - cam[0], y = my - cam[1], z = mz - cam[2];
            if (DEG > 0) {
                const float inorm = rsqrtf(x * x + y * y + z * z);
                x *= inorm; y *= inorm; z *= inorm;
            }
            sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
                r += Y * cf[k * 3]; gg += Y * cf[k * 3 + 1]; b += Y * cf[k * 3 + 2];
            });
            r = fmaxf(r + 0.5f, 0.f); gg = fmaxf(gg + 0.5f, 0.f); b = fmaxf(b + 0.5f, 0.f);
        }
        if (vis || VIEWS::kWriteInvisible) {
            float *col = views.color(c, g);
            col[0] = r; col[1] = gg; col[2] = b;
        }
    }
}

template <int DEG, typename VIEWS>
__global__ void __launch_bounds__(kSplitThreads)
sh_view_split_bwd_kernel(uint32_t C, uint32_t N, uint32_t RF, const float *__restrict__ means, const float *__restrict__ sh0,
                         const float *__restrict__ shN, const __grid_constant__ VIEWS views, float *__restrict__ v_sh0,
                         float *__restrict__ v_shN, float *__restrict__ v_means, float *__restrict__ v_campos,
                         const __grid_constant__ QuantLoad q0) {
    extern __shared__ float s_slab[];  // [kSplitThreads][RF | 1]: the block's v_shN rows
    __shared__ float s_red[kSplitThreads / 32][3];
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    constexpr uint32_t nact = nb * 3 - 3;  // floats of an shN row that the active bands cover
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t g = blockIdx.x * kSplitThreads + tid;
    float vc[nb * 3];
#pragma unroll
    for (uint32_t k = 0; k < nb * 3; ++k) vc[k] = 0.f;
    float vmx = 0.f, vmy = 0.f, vmz = 0.f;
    bool any = false;
    if (g < N)
        for (uint32_t c = 0; c < C; ++c) any |= views.visible(c, g);
    float cf[nb * 3];
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (any) {
        cf[0] = __ldg(sh0 + (size_t)g * 3); cf[1] = __ldg(sh0 + (size_t)g * 3 + 1); cf[2] = __ldg(sh0 + (size_t)g * 3 + 2);
        if (q0.on) {  // sh0 is the RAW parameter: fake-quantise it here (quant_load.cuh)
            const float raw[3] = {cf[0], cf[1], cf[2]};
            float qv[3];
            quant_load<3>(q0, (int64_t)g * 3, raw, qv);
            cf[0] = qv[0]; cf[1] = qv[1]; cf[2] = qv[2];
        }
        if (DEG > 0) {
            const float *row = shN + (size_t)g * RF;
#pragma unroll
            for (uint32_t k = 0; k + 3 < nb * 3; ++k) cf[3 + k] = __ldg(row + k);
        }
        mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2];
    }
    for (uint32_t c = 0; c < C; ++c) {
        float dxv = 0.f, dyv = 0.f, dzv = 0.f;
        if (any && views.visible(c, g)) {
            const float *cam = views.cam(c), *vcol = views.v_color(c, g);
            float x = mx - cam[0], y = my - cam[1], z = mz - cam[2];
            float inorm = 1.f;
            if (DEG > 0) {
                inorm = rsqrtf(x * x + y * y + z * z);
                x *= inorm; y *= inorm; z *= inorm;
            }
            float r = 0.f, gg = 0.f, b = 0.f;  // pre-clamp colour: the gradient passes where x + 0.5 >= 0
            sh_terms(DEG, x, y, z, [&](int k, float Y, float, float, float) {
                r += Y * cf[k * 3]; gg += Y * cf[k * 3 + 1]; b += Y * cf[k * 3 + 2];
            });
            const float vr = (r + 0.5f >= 0.f) ? vcol[0] : 0.f;
            const float vg = (gg + 0.5f >= 0.f) ? vcol[1] : 0.f;
            const float vb = (b + 0.5f >= 0.f) ? vcol[2] : 0.f;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            sh_terms(DEG, x, y, z, [&](int k, float Y, float dx, float dy, float dz) {
                vc[k * 3] += Y * vr; vc[k * 3 + 1] += Y * vg; vc[k * 3 + 2] += Y * vb;
                const float wk = cf[k * 3] * vr + cf[k * 3 + 1] * vg + cf[k * 3 + 2] * vb;
                gx += dx * wk; gy += dy * wk; gz += dz * wk;
            });
            if (DEG > 0) {
                const float dotp = gx * x + gy * y + gz * z;
                dxv = (gx - dotp * x) * inorm; dyv = (gy - dotp * y) * inorm; dzv = (gz - dotp * z) * inorm;
                vmx += dxv; vmy += dyv; vmz += dzv;
            }
        }
        if (v_campos != nullptr) {  // d dirs / d campos = -1: block reduction, one atomic set per block and camera
            const float sx = warp_sum(dxv), sy = warp_sum(dyv), sz = warp_sum(dzv);
            __syncthreads();
            if (lane == 0) { s_red[tid >> 5][0] = sx; s_red[tid >> 5][1] = sy; s_red[tid >> 5][2] = sz; }
            __syncthreads();
            if (tid < 3) {
                float t = 0.f;
                for (uint32_t w = 0; w < kSplitThreads / 32; ++w) t += s_red[w][tid];
                if (t != 0.f) atomicAdd(v_campos + c * 3 + tid, -t);
            }
        }
    }
    if (g < N) {
        if (q0.on && any) {  // gradient to the RAW sh0: the clamp gate of the noise-mode fake quantisation (ops.py:44-47)
#pragma unroll
            for (int k = 0; k < 3; ++k) vc[k] = quant_grad(q0, vc[k], __ldg(sh0 + (size_t)g * 3 + k), 0.f);
        }
        v_sh0[(size_t)g * 3] = vc[0]; v_sh0[(size_t)g * 3 + 1] = vc[1]; v_sh0[(size_t)g * 3 + 2] = vc[2];
        if (v_means != nullptr) {
            v_means[(size_t)g * 3] = vmx; v_means[(size_t)g * 3 + 1] = vmy; v_means[(size_t)g * 3 + 2] = vmz;
        }
    }
    if (RF == 0) return;
    // a block none of whose Gaussians is visible in any view (71 % of the Gaussians of the headline frame are culled, and
    // neighbours in memory are neighbours in space): its slab is all zeros -- plain 16-byte stores, no staging
    if (!__syncthreads_or(any)) {
        const uint32_t row0z = blockIdx.x * kSplitThreads;
        const uint32_t totalz = min((uint32_t)kSplitThreads, N - row0z) * RF;
        float *dstz = v_shN + (size_t)row0z * RF;
        for (uint32_t v = tid; v < totalz / 4; v += kSplitThreads) stg4(dstz + (size_t)v * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        const uint32_t iz = (totalz / 4) * 4 + tid;
        if (iz < totalz) dstz[iz] = 0.f;
        return;
    }
    const uint32_t stride = RF | 1u;  // odd: the per-thread row writes below are bank-conflict free
    if (DEG > 0) {
#pragma unroll
        for (uint32_t k = 0; k < nact; ++k) s_slab[tid * stride + k] = vc[3 + k];
    }
    __syncthreads();
    // coalesced copy of the slab: element i of the block's region is (row i / RF, column i % RF); columns past the
    // active bands are zero.  The region starts at a multiple of kSplitThreads * RF floats, i.e. 16-byte aligned.
    const uint32_t row0 = blockIdx.x * kSplitThreads;
    const uint32_t nrows = min((uint32_t)kSplitThreads, N - row0);
    const uint32_t total = nrows * RF, nvec = total / 4;
    float *dst = v_shN + (size_t)row0 * RF;
    uint32_t row = (4 * tid) / RF, col = (4 * tid) - row * RF;
    const uint32_t step_row = (4 * kSplitThreads) / RF, step_col = (4 * kSplitThreads) - step_row * RF;
    for (uint32_t v = tid; v < nvec; v += kSplitThreads) {
        float e[4];
        uint32_t r = row, cc = col;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[q] = cc < nact ? s_slab[r * stride + cc] : 0.f;
            if (++cc == RF) { cc = 0; ++r; }
        }
        stg4(dst + (size_t)v * 4, make_float4(e[0], e[1], e[2], e[3]));
        row += step_row; col += step_col;
        if (col >= RF) { col -= RF; ++row; }
    }
    const uint32_t i = nvec * 4 + tid;  // ragged tail of the last block
    if (i < total) {
        const uint32_t r = i / RF, cc = i - r * RF;
        dst[i] = cc < nact ? s_slab[r * stride + cc] : 0.f;
    }
}


// viewer -> owners: slot c of src ([n_slots][world * n_block, width]; rows p * n_block.. belong to owner p) is stored as one
// dense [n_block * width] block into owner p's staging at slot (slot0 + c), 16-byte vectors over NVLink; block (0, p, c)
// also delivers the slot's `extra_w` extra floats (the view's camera centre) to owner p.
__global__ void __launch_bounds__(kThreads)
shard_push_rows_kernel(uint32_t n_block, uint32_t width, uint32_t slot0, uint64_t slot_floats, const float *__restrict__ src,
                       uint64_t src_slot_floats, const PeerRows peers, const float *__restrict__ extra_src, uint32_t extra_w,
                       const PeerRows peer_extra) {
    const uint32_t p = blockIdx.y, c = blockIdx.z;
    const uint64_t n_vec = (uint64_t)n_block * width / 4;
    const float *from = src + (size_t)c * src_slot_floats + (size_t)p * n_block * width;
    float *dst = peers.base[p] + (size_t)(slot0 + c) * slot_floats;
    const uint64_t i0 = (uint64_t)blockIdx.x * kThreads * kGeoUnroll + threadIdx.x;
    float4 v[kGeoUnroll];
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u)
        if (i0 + (uint64_t)u * kThreads < n_vec) v[u] = ldg4(from + (i0 + (uint64_t)u * kThreads) * 4);
#pragma unroll
    for (int u = 0; u < kGeoUnroll; ++u)
        if (i0 + (uint64_t)u * kThreads < n_vec) stg4(dst + (i0 + (uint64_t)u * kThreads) * 4, v[u]);
    if (blockIdx.x == 0 && extra_src != nullptr && threadIdx.x < extra_w)
        peer_extra.base[p][(size_t)(slot0 + c) * extra_w + threadIdx.x] = extra_src[(size_t)c * extra_w + threadIdx.x];
}

// a few floats per slot (cameras) to every peer: peer p gets src [n_slots, w] at float offset slot0 * w
__global__ void peer_push_small_kernel(uint32_t n_floats, uint64_t dst_offset, const float *__restrict__ src, const PeerRows peers) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_floats) peers.base[blockIdx.y][dst_offset + i] = src[i];
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_sh_fwd(uint32_t n, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                             const uint8_t *masks, float *colors, gsb200_stream_t stream) {
    if (degree > 4) return fail(2, "gsb200_sh_fwd: degree must be <= 4");
    if ((degree + 1) * (degree + 1) > K) return fail(2, "gsb200_sh_fwd: coeffs hold fewer than (degree+1)^2 bases");
    if (n == 0) return 0;
    const bool vec = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0) &&
                     (((degree + 1) * (degree + 1) * 3) % 4 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t grid = cdiv(n, kThreads);
    switch (degree) {
        case 0: sh_fwd_kernel<0><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, colors, false); break;
        case 1: sh_fwd_kernel<1><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, colors, vec); break;
        case 2: sh_fwd_kernel<2><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, colors, false); break;
        case 3: sh_fwd_kernel<3><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, colors, vec); break;
        default: sh_fwd_kernel<4><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, colors, false); break;
    }
    return check_launch("sh_fwd");
}

extern "C" int gsb200_sh_bwd(uint32_t n, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                             const uint8_t *masks, const float *v_colors, float *v_coeffs, float *v_dirs,
                             gsb200_stream_t stream) {
    if (degree > 4) return fail(2, "gsb200_sh_bwd: degree must be <= 4");
    if ((degree + 1) * (degree + 1) > K) return fail(2, "gsb200_sh_bwd: coeffs hold fewer than (degree+1)^2 bases");
    if (n == 0) return 0;
    // vector path: rows 16-byte aligned (K*3 % 4 == 0) for both reads and writes; reads additionally
    // need (degree+1)^2*3 % 4 == 0 only for the coefficient prefix, which the kernel loads in 4s.
    const bool vec = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(v_coeffs) & 15) == 0) && (((degree + 1) * (degree + 1) * 3) % 4 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t grid = cdiv(n, kThreads);
    switch (degree) {
        case 0: sh_bwd_kernel<0><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, false); break;
        case 1: sh_bwd_kernel<1><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, vec); break;
        case 2: sh_bwd_kernel<2><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, false); break;
        case 3: sh_bwd_kernel<3><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, vec); break;
        default: sh_bwd_kernel<4><<<grid, kThreads, 0, s>>>(n, K, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, false); break;
    }
    return check_launch("sh_bwd");
}

static bool sh_vec_ok(uint32_t K, uint32_t degree, const void *a, const void *b) {
    return ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0) &&
           (((degree + 1) * (degree + 1) * 3) % 4 == 0);
}

template <int DEG, bool COOP>
static void launch_view_fwd(cudaStream_t s, uint32_t C, uint32_t N, uint32_t K, const float *means, const float *campos,
                            const CoeffShards &sh, const int32_t *radii, float *colors, float *row_cache, bool vec) {
    const uint32_t threads = COOP ? kCoopThreads : kThreads;
    sh_view_fwd_kernel<DEG, COOP><<<cdiv(N, threads), threads, 0, s>>>(C, N, K, means, campos, sh, radii, colors, row_cache, vec);
}

template <int DEG, bool COOP>
static void launch_view_bwd(cudaStream_t s, uint32_t C, uint32_t N, uint32_t K, const float *means, const float *campos,
                            const CoeffShards &sh, const int32_t *radii, const float *v_colors, float *v_coeffs,
                            const PushTargets &push, uint32_t push_block, float *v_means, float *v_campos, bool vec) {
    const uint32_t threads = COOP ? kCoopThreads : kThreads;
    sh_view_bwd_kernel<DEG, COOP><<<cdiv(N, threads), threads, 0, s>>>(C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push,
                                                                        push_block, v_means, v_campos, vec);
}

// coop: warp-cooperative row staging (multi-GPU form, 16-byte rows, every coefficient of the row used)
static bool coop_ok(uint32_t K, uint32_t degree, bool vec) { return vec && (degree == 1 || degree == 3) && (degree + 1) * (degree + 1) == K; }

static int view_fwd(const char *who, uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                    const CoeffShards &sh, bool vec, bool coop, const int32_t *radii, float *colors, float *row_cache, cudaStream_t s) {
    if (coop && degree == 3) launch_view_fwd<3, true>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec);
    else if (coop) launch_view_fwd<1, true>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec);
    else switch (degree) {
        case 0: launch_view_fwd<0, false>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec); break;
        case 1: launch_view_fwd<1, false>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec); break;
        case 2: launch_view_fwd<2, false>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec); break;
        case 3: launch_view_fwd<3, false>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec); break;
        default: launch_view_fwd<4, false>(s, C, N, K, means, campos, sh, radii, colors, row_cache, vec); break;
    }
    return check_launch(who);
}

static int view_bwd(const char *who, uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                    const CoeffShards &sh, bool vec, bool coop, const int32_t *radii, const float *v_colors, float *v_coeffs,
                    const PushTargets &push, uint32_t push_block, float *v_means, float *v_campos, cudaStream_t s) {
    if (v_campos != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_campos, 0, sizeof(float) * 3 * C, s));
    if (coop && degree == 3) launch_view_bwd<3, true>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec);
    else if (coop) launch_view_bwd<1, true>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec);
    else switch (degree) {
        case 0: launch_view_bwd<0, false>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec); break;
        case 1: launch_view_bwd<1, false>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec); break;
        case 2: launch_view_bwd<2, false>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec); break;
        case 3: launch_view_bwd<3, false>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec); break;
        default: launch_view_bwd<4, false>(s, C, N, K, means, campos, sh, radii, v_colors, v_coeffs, push, push_block, v_means, v_campos, vec); break;
    }
    return check_launch(who);
}

static int check_view_args(const char *who, uint32_t K, uint32_t degree) {
    if (degree > 4) return failf(2, "%s: degree must be <= 4", who);
    if ((degree + 1) * (degree + 1) > K) return failf(2, "%s: coeffs hold fewer than (degree+1)^2 bases", who);
    return 0;
}

// fills the shard table from a host array of `world` device pointers; the vector path needs every block 16-byte aligned
static int make_shards(const char *who, const float *const *shard_coeffs, uint32_t world, uint32_t n_block, uint32_t N,
                       CoeffShards &sh, bool &aligned) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "%s: world must be in [1, %d]", who, kMaxShards);
    if (n_block == 0 || (uint64_t)n_block * world < N) return failf(2, "%s: world * n_block must cover N", who);
    aligned = true;
    for (uint32_t r = 0; r < (uint32_t)kMaxShards; ++r) {
        sh.base[r] = r < world ? shard_coeffs[r] : nullptr;
        if (r < world && sh.base[r] == nullptr) return failf(2, "%s: null shard pointer", who);
        aligned = aligned && (reinterpret_cast<uintptr_t>(sh.base[r]) & 15) == 0;
    }
    sh.n_block = n_block;
    return 0;
}

extern "C" int gsb200_sh_view_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                  const float *coeffs, const int32_t *radii, float *colors, gsb200_stream_t stream) {
    if (int rc = check_view_args("gsb200_sh_view_fwd", K, degree)) return rc;
    if (C == 0 || N == 0) return 0;
    CoeffShards sh{};
    sh.base[0] = coeffs;
    return view_fwd("sh_view_fwd", C, N, K, degree, means, campos, sh, sh_vec_ok(K, degree, coeffs, nullptr), false, radii, colors,
                    nullptr, (cudaStream_t)stream);
}

extern "C" int gsb200_sh_view_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                  const float *coeffs, const int32_t *radii, const float *v_colors, float *v_coeffs,
                                  float *v_means, float *v_campos, gsb200_stream_t stream) {
    if (int rc = check_view_args("gsb200_sh_view_bwd", K, degree)) return rc;
    if (C == 0 || N == 0) return 0;
    CoeffShards sh{};
    sh.base[0] = coeffs;
    const PushTargets none{};
    return view_bwd("sh_view_bwd", C, N, K, degree, means, campos, sh, sh_vec_ok(K, degree, coeffs, v_coeffs), false, radii, v_colors,
                    v_coeffs, none, 0, v_means, v_campos, (cudaStream_t)stream);
}

extern "C" int gsb200_sh_view_fwd_sharded(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                          const float *campos, const float *const *shard_coeffs, uint32_t world, uint32_t n_block,
                                          const int32_t *radii, float *colors, float *row_cache, gsb200_stream_t stream) {
    if (int rc = check_view_args("gsb200_sh_view_fwd_sharded", K, degree)) return rc;
    if (C == 0 || N == 0) return 0;
    CoeffShards sh{};
    bool aligned = false;
    if (int rc = make_shards("gsb200_sh_view_fwd_sharded", shard_coeffs, world, n_block, N, sh, aligned)) return rc;
    const bool vec = aligned && sh_vec_ok(K, degree, row_cache, nullptr);
    return view_fwd("sh_view_fwd_sharded", C, N, K, degree, means, campos, sh, vec, coop_ok(K, degree, vec), radii, colors,
                    row_cache, (cudaStream_t)stream);
}

extern "C" int gsb200_sh_view_bwd_sharded(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                          const float *campos, const float *const *shard_coeffs, uint32_t world, uint32_t n_block,
                                          uint32_t rank, const float *row_cache, const int32_t *radii, const float *v_colors,
                                          float *const *peer_rows, int32_t *const *peer_flags, float *v_means, float *v_campos,
                                          gsb200_stream_t stream) {
    const char *who = "gsb200_sh_view_bwd_sharded";
    if (int rc = check_view_args(who, K, degree)) return rc;
    if (peer_rows == nullptr || peer_flags == nullptr) return fail(2, "gsb200_sh_view_bwd_sharded: null peer tables");
    if (C == 0 || N == 0) return 0;
    CoeffShards sh{};
    bool aligned = false;
    if (int rc = make_shards(who, shard_coeffs, world, n_block, N, sh, aligned)) return rc;
    if (rank >= world) return fail(2, "gsb200_sh_view_bwd_sharded: rank >= world");
    PushTargets push{};
    push.src = rank;
    for (uint32_t r = 0; r < world; ++r) {
        push.rows[r] = peer_rows[r];
        push.flags[r] = peer_flags[r];
        if (push.rows[r] == nullptr || push.flags[r] == nullptr) return fail(2, "gsb200_sh_view_bwd_sharded: null peer pointer");
        aligned = aligned && (reinterpret_cast<uintptr_t>(push.rows[r]) & 15) == 0;
    }
    if (row_cache != nullptr) {  // the forward kept a local copy of every visible row: no second pull over the link
        sh = CoeffShards{};
        sh.base[0] = row_cache;
        aligned = aligned && (reinterpret_cast<uintptr_t>(row_cache) & 15) == 0;
    }
    const bool vec = aligned && sh_vec_ok(K, degree, nullptr, nullptr);
    return view_bwd("sh_view_bwd_sharded", C, N, K, degree, means, campos, sh, vec, coop_ok(K, degree, vec), radii, v_colors, nullptr,
                    push, n_block, v_means, v_campos, (cudaStream_t)stream);
}

extern "C" int gsb200_shard_reduce_rows(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_floats, const float *staging,
                                        const int32_t *flags, float *out, gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "gsb200_shard_reduce_rows: world must be in [1, %d]", kMaxShards);
    if (n_rows > n_block) return fail(2, "gsb200_shard_reduce_rows: n_rows > n_block");
    if (row_floats % 4 != 0) return fail(2, "gsb200_shard_reduce_rows: row_floats must be a multiple of 4");
    if (n_rows == 0) return 0;
    if (staging == nullptr || flags == nullptr || out == nullptr || ((reinterpret_cast<uintptr_t>(staging) | reinterpret_cast<uintptr_t>(out)) & 15) != 0)
        return fail(2, "gsb200_shard_reduce_rows: pointers must be non-null and 16-byte aligned");
    const uint32_t row_vec4 = row_floats / 4;
    const size_t total = (size_t)n_rows * row_vec4;
    shard_reduce_rows_kernel<<<(uint32_t)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        world, n_block, n_rows, row_vec4, staging, flags, out);
    return check_launch("shard_reduce_rows");
}

extern "C" int gsb200_shard_reduce_rows_split(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_floats, uint32_t split,
                                              const float *staging, const int32_t *flags, float *out0, float *out1,
                                              gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "gsb200_shard_reduce_rows_split: world must be in [1, %d]", kMaxShards);
    if (n_rows > n_block) return fail(2, "gsb200_shard_reduce_rows_split: n_rows > n_block");
    if (row_floats % 4 != 0 || split == 0 || split >= row_floats) return fail(2, "gsb200_shard_reduce_rows_split: bad row / split size");
    if (n_rows == 0) return 0;
    if (staging == nullptr || flags == nullptr || out0 == nullptr || out1 == nullptr || (reinterpret_cast<uintptr_t>(staging) & 15) != 0)
        return fail(2, "gsb200_shard_reduce_rows_split: pointers must be non-null, staging 16-byte aligned");
    const uint32_t row_vec4 = row_floats / 4;
    const size_t total = (size_t)n_rows * row_vec4;
    shard_reduce_rows_split_kernel<<<(uint32_t)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        world, n_block, n_rows, row_vec4, split, staging, flags, out0, out1);
    return check_launch("shard_reduce_rows_split");
}

extern "C" int gsb200_shard_gather_rows(uint32_t world, uint64_t block_floats, const float *const *blocks, float *out,
                                        gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "gsb200_shard_gather_rows: world must be in [1, %d]", kMaxShards);
    if (block_floats % 4 != 0) return fail(2, "gsb200_shard_gather_rows: block_floats must be a multiple of 4");
    if (block_floats == 0) return 0;
    BlockShards sh{};
    for (uint32_t r = 0; r < world; ++r) {
        sh.base[r] = blocks[r];
        if (sh.base[r] == nullptr || (reinterpret_cast<uintptr_t>(sh.base[r]) & 15) != 0)
            return fail(2, "gsb200_shard_gather_rows: block pointers must be non-null and 16-byte aligned");
    }
    if (out == nullptr || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return fail(2, "gsb200_shard_gather_rows: out must be 16-byte aligned");
    const uint64_t block_vec4 = block_floats / 4;
    const dim3 grid((uint32_t)((block_vec4 + kThreads * 4 - 1) / (kThreads * 4)), world);
    shard_gather_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(block_vec4, sh, out);
    return check_launch("shard_gather_rows");
}


static int geo_parts(const char *who, uint32_t n_parts, float *const *ptrs, const uint32_t *widths, bool need_ptrs, GeoParts &P) {
    if (n_parts == 0 || n_parts > 4) return failf(2, "%s: 1..4 attribute tensors", who);
    P = GeoParts{};
    P.n_parts = n_parts;
    uint32_t col = 0;
    for (uint32_t k = 0; k < n_parts; ++k) {
        if (widths[k] == 0) return failf(2, "%s: zero-width attribute", who);
        if (need_ptrs && ptrs[k] == nullptr) return failf(2, "%s: null attribute pointer", who);
        P.ptr[k] = ptrs[k];
        P.width[k] = widths[k];
        P.col0[k] = col;
        col += widths[k];
    }
    P.w = col;
    return 0;
}

extern "C" int gsb200_shard_gather_geo(uint32_t world, uint32_t n_block, uint32_t n_parts, const uint32_t *widths,
                                       const float *const *blocks, float *const *outs, gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "gsb200_shard_gather_geo: world must be in [1, %d]", kMaxShards);
    GeoParts P;
    if (int e = geo_parts("gsb200_shard_gather_geo", n_parts, outs, widths, true, P)) return e;
    if (n_block % 4 != 0) return fail(2, "gsb200_shard_gather_geo: n_block must be a multiple of 4");
    if (n_block == 0) return 0;
    for (uint32_t k = 0; k < n_parts; ++k)
        if ((reinterpret_cast<uintptr_t>(outs[k]) & 15) != 0) return fail(2, "gsb200_shard_gather_geo: outputs must be 16-byte aligned");
    BlockShards sh{};
    for (uint32_t r = 0; r < world; ++r) {
        sh.base[r] = blocks[r];
        if (sh.base[r] == nullptr || (reinterpret_cast<uintptr_t>(sh.base[r]) & 15) != 0)
            return fail(2, "gsb200_shard_gather_geo: block pointers must be non-null and 16-byte aligned");
    }
    const uint64_t n_vec = (uint64_t)n_block * P.w / 4;
    const dim3 grid((uint32_t)((n_vec + kThreads * kGeoUnroll - 1) / (kThreads * kGeoUnroll)), world);
    shard_gather_geo_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n_block, sh, P);
    return check_launch("shard_gather_geo");
}

// camera centres of rigid world-to-camera matrices: -R^T t (one thread per camera; replaces a batched matmul + negate,
// two launches and ~55 us of host time per step for a few dozen flops)
__global__ void camera_positions_kernel(uint32_t C, const float *__restrict__ viewmats, float *__restrict__ campos) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float *V = viewmats + (size_t)c * 16;
    const float tx = V[3], ty = V[7], tz = V[11];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        campos[c * 3 + j] = -__fadd_rn(__fadd_rn(__fmul_rn(V[j], tx), __fmul_rn(V[4 + j], ty)), __fmul_rn(V[8 + j], tz));
}

extern "C" int gsb200_camera_positions(uint32_t C, const float *viewmats, float *campos, gsb200_stream_t stream) {
    if (C == 0) return 0;
    if (viewmats == nullptr || campos == nullptr) return fail(2, "gsb200_camera_positions: null pointer");
    camera_positions_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, viewmats, campos);
    return check_launch("camera_positions");
}

extern "C" int gsb200_shard_publish(uint32_t n_local, uint32_t n_block, uint32_t n_parts, const uint32_t *widths,
                                    float *const *parts, const float *pad_row, float *geo_block, const float *first,
                                    const float *rest, uint32_t first_floats, uint32_t row_floats, float *coeff_block,
                                    gsb200_stream_t stream) {
    if (n_local > n_block) return fail(2, "gsb200_shard_publish: n_local > n_block");
    GeoParts P{};
    uint32_t n_geo_blocks = 0;
    if (n_parts != 0) {
        if (int e = geo_parts("gsb200_shard_publish", n_parts, parts, widths, true, P)) return e;
        if (geo_block == nullptr || (reinterpret_cast<uintptr_t>(geo_block) & 15) != 0)
            return fail(2, "gsb200_shard_publish: geo_block must be non-null and 16-byte aligned");
        if (n_block % 4 != 0) return fail(2, "gsb200_shard_publish: n_block must be a multiple of 4");
        n_geo_blocks = (uint32_t)(((uint64_t)n_block * P.w / 4 + kThreads - 1) / kThreads);
    }
    uint64_t n_coeff_blocks = 0;
    if (first != nullptr && n_local != 0) {
        if (row_floats == 0 || row_floats % 4 != 0 || first_floats == 0 || first_floats > row_floats ||
            (first_floats < row_floats && rest == nullptr))
            return fail(2, "gsb200_shard_publish: row_floats must be a positive multiple of 4 covered by first (+ rest)");
        if (coeff_block == nullptr || (reinterpret_cast<uintptr_t>(coeff_block) & 15) != 0)
            return fail(2, "gsb200_shard_publish: coeff_block must be non-null and 16-byte aligned");
        n_coeff_blocks = ((uint64_t)n_local * (row_floats / 4) + kThreads - 1) / kThreads;
    }
    const uint64_t grid = n_geo_blocks + n_coeff_blocks;
    if (grid == 0) return 0;
    if (grid > 0x7fffffffull) return fail(2, "gsb200_shard_publish: too many rows for one launch");
    shard_publish_kernel<<<(uint32_t)grid, kThreads, 0, (cudaStream_t)stream>>>(
        n_local, n_block, n_geo_blocks, P, pad_row, geo_block, n_coeff_blocks ? first : nullptr, rest, first_floats, row_floats,
        coeff_block);
    return check_launch("shard_publish");
}

extern "C" int gsb200_peer_barrier(uint32_t world, uint32_t rank, uint32_t *const *pads, uint32_t epoch, gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards || rank >= world)
        return failf(2, "gsb200_peer_barrier: world must be in [1, %d] and rank < world", kMaxShards);
    PeerPads P{};
    for (uint32_t r = 0; r < world; ++r) {
        P.base[r] = pads[r];
        if (P.base[r] == nullptr) return fail(2, "gsb200_peer_barrier: null signal pad");
    }
    peer_barrier_kernel<<<1, world, 0, (cudaStream_t)stream>>>(rank, epoch, P);
    return check_launch("peer_barrier");
}

extern "C" int gsb200_shard_push_geo(uint32_t world, uint32_t n_block, uint32_t rank, uint64_t slot_floats, uint32_t n_parts,
                                     const uint32_t *widths, float *const *grads, float *const *peer_staging,
                                     gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards || rank >= world)
        return failf(2, "gsb200_shard_push_geo: world must be in [1, %d] and rank < world", kMaxShards);
    GeoParts P;
    if (int e = geo_parts("gsb200_shard_push_geo", n_parts, grads, widths, false, P)) return e;
    if ((uint64_t)n_block * P.w > slot_floats || slot_floats % 4 != 0)
        return fail(2, "gsb200_shard_push_geo: staging slot too small or not a multiple of 4 floats");
    if (n_block % 4 != 0) return fail(2, "gsb200_shard_push_geo: n_block must be a multiple of 4");
    if (n_block == 0) return 0;
    bool vec = true;  // 16-byte loads of the gradient tensors only if every one of them is aligned
    for (uint32_t k = 0; k < n_parts; ++k) vec = vec && (reinterpret_cast<uintptr_t>(grads[k]) & 15) == 0;
    PeerRows peers{};
    for (uint32_t r = 0; r < world; ++r) {
        peers.base[r] = peer_staging[r];
        if (peers.base[r] == nullptr || (reinterpret_cast<uintptr_t>(peers.base[r]) & 15) != 0)
            return fail(2, "gsb200_shard_push_geo: staging pointers must be non-null and 16-byte aligned");
    }
    const uint64_t n_vec = (uint64_t)n_block * P.w / 4;
    const dim3 grid((uint32_t)((n_vec + kThreads * kGeoUnroll - 1) / (kThreads * kGeoUnroll)), world);
    shard_push_geo_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n_block, rank, slot_floats, vec, P, peers);
    return check_launch("shard_push_geo");
}

extern "C" int gsb200_shard_reduce_geo(uint32_t world, uint32_t n_block, uint32_t n_rows, uint64_t slot_floats, uint32_t n_parts,
                                       const uint32_t *widths, const float *staging, float *const *outs, gsb200_stream_t stream) {
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "gsb200_shard_reduce_geo: world must be in [1, %d]", kMaxShards);
    if (n_rows > n_block) return fail(2, "gsb200_shard_reduce_geo: n_rows > n_block");
    GeoParts P;
    if (int e = geo_parts("gsb200_shard_reduce_geo", n_parts, outs, widths, false, P)) return e;
    if ((uint64_t)n_block * P.w > slot_floats || slot_floats % 4 != 0 || staging == nullptr ||
        (reinterpret_cast<uintptr_t>(staging) & 15) != 0)
        return fail(2, "gsb200_shard_reduce_geo: bad staging");
    if (n_block % 4 != 0) return fail(2, "gsb200_shard_reduce_geo: n_block must be a multiple of 4");
    if (n_rows == 0) return 0;
    for (uint32_t k = 0; k < n_parts; ++k)
        if ((reinterpret_cast<uintptr_t>(outs[k]) & 15) != 0) return fail(2, "gsb200_shard_reduce_geo: outputs must be 16-byte aligned");
    const uint64_t n_vec = (uint64_t)n_block * P.w / 4;
    shard_reduce_geo_kernel<<<(uint32_t)((n_vec + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        world, n_block, n_rows, slot_floats, staging, P);
    return check_launch("shard_reduce_geo");
}

// ---- split-coefficient entry points (sh0 [N,1,3] + shN [N,K-1,3] in place of one [N,K,3] tensor) -------------
static int check_split_args(const char *who, uint32_t K, uint32_t degree, const float *sh0, const float *shN) {
    if (int rc = check_view_args(who, K, degree)) return rc;
    if (K - 1 > GSB200_SH_SPLIT_MAX_REST) return failf(2, "%s: at most %d higher-band coefficients per Gaussian", who, GSB200_SH_SPLIT_MAX_REST);
    if (sh0 == nullptr || (degree > 0 && shN == nullptr)) return failf(2, "%s: null coefficient pointer", who);
    return 0;
}

template <typename VIEWS>
static void launch_split_fwd(cudaStream_t s, uint32_t degree, uint32_t C, uint32_t N, uint32_t K, const float *means, const float *sh0,
                             const float *shN, const VIEWS &views, const QuantLoad &q0 = QuantLoad{nullptr, 0, 0, 0.f, 0.f, 0.f, 0, 0}) {
    const uint32_t grid = cdiv(N, kSplitThreads), RF = (K - 1) * 3;
    switch (degree) {
        case 0: sh_view_split_fwd_kernel<0, VIEWS><<<grid, kSplitThreads, 0, s>>>(C, N, RF, means, sh0, shN, views, q0); break;
        case 1: sh_view_split_fwd_kernel<1, VIEWS><<<grid, kSplitThreads, 0, s>>>(C, N, RF, means, sh0, shN, views, q0); break;
        case 2: sh_view_split_fwd_kernel<2, VIEWS><<<grid, kSplitThreads, 0, s>>>(C, N, RF, means, sh0, shN, views, q0); break;
        case 3: sh_view_split_fwd_kernel<3, VIEWS><<<grid, kSplitThreads, 0, s>>>(C, N, RF, means, sh0, shN, views, q0); break;
        default: sh_view_split_fwd_kernel<4, VIEWS><<<grid, kSplitThreads, 0, s>>>(C, N, RF, means, sh0, shN, views, q0); break;
    }
}

template <typename VIEWS>
static void launch_split_bwd(cudaStream_t s, uint32_t degree, uint32_t C, uint32_t N, uint32_t K, const float *means, const float *sh0,
                             const float *shN, const VIEWS &views, float *v_sh0, float *v_shN, float *v_means, float *v_campos,
                             const QuantLoad &q0 = QuantLoad{nullptr, 0, 0, 0.f, 0.f, 0.f, 0, 0}) {
    const uint32_t grid = cdiv(N, kSplitThreads), RF = (K - 1) * 3;
    const size_t smem = sizeof(float) * kSplitThreads * (RF | 1u);
#define GSB_SPLIT_BWD(D)                                                                                                      \
    sh_view_split_bwd_kernel<D, VIEWS><<<grid, kSplitThreads, smem, s>>>(C, N, RF, means, sh0, shN, views, v_sh0, v_shN, v_means, \
                                                                         v_campos, q0)
    switch (degree) {
        case 0: GSB_SPLIT_BWD(0); break;
        case 1: GSB_SPLIT_BWD(1); break;
        case 2: GSB_SPLIT_BWD(2); break;
        case 3: GSB_SPLIT_BWD(3); break;
        default: GSB_SPLIT_BWD(4); break;
    }
#undef GSB_SPLIT_BWD
}

extern "C" int gsb200_sh_view_split_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                        const float *sh0, const float *shN, const int32_t *radii, float *colors,
                                        gsb200_stream_t stream) {
    if (int rc = check_split_args("gsb200_sh_view_split_fwd", K, degree, sh0, shN)) return rc;
    if (C == 0 || N == 0) return 0;
    const LocalViews views{campos, radii, colors, nullptr, N};
    launch_split_fwd((cudaStream_t)stream, degree, C, N, K, means, sh0, shN, views);
    return check_launch("sh_view_split_fwd");
}

extern "C" int gsb200_sh_view_split_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                        const float *sh0, const float *shN, const int32_t *radii, const float *v_colors,
                                        float *v_sh0, float *v_shN, float *v_means, float *v_campos, gsb200_stream_t stream) {
    const char *who = "gsb200_sh_view_split_bwd";
    if (int rc = check_split_args(who, K, degree, sh0, shN)) return rc;
    if (C == 0 || N == 0) return 0;
    if (v_sh0 == nullptr || (K > 1 && v_shN == nullptr)) return failf(2, "%s: null gradient pointer", who);
    if ((reinterpret_cast<uintptr_t>(v_shN) & 15) != 0) return failf(2, "%s: v_shN must be 16-byte aligned", who);
    cudaStream_t s = (cudaStream_t)stream;
    if (v_campos != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_campos, 0, sizeof(float) * 3 * C, s));
    const LocalViews views{campos, radii, nullptr, v_colors, N};
    launch_split_bwd(s, degree, C, N, K, means, sh0, shN, views, v_sh0, v_shN, v_means, v_campos);
    return check_launch("sh_view_split_bwd");
}

// ---- the same two calls with sh0 fake-quantised where it is loaded (noise mode; quant_load.cuh): `sh0` is the RAW
// parameter, v_sh0 its gradient (clamp gate applied).  q_sh0 == NULL: identical to the calls above.
extern "C" int gsb200_sh_view_split_quant_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                              const float *campos, const float *sh0, const gsb200_quant_load *q_sh0, const float *shN,
                                              const int32_t *radii, float *colors, gsb200_stream_t stream) {
    const char *who = "gsb200_sh_view_split_quant_fwd";
    if (int rc = check_split_args(who, K, degree, sh0, shN)) return rc;
    QuantLoad q0;
    if (int rc = make_quant_load(who, q_sh0, q0)) return rc;
    if (q0.on && q0.act != 0) return failf(2, "%s: sh0 takes no activation", who);
    if (C == 0 || N == 0) return 0;
    const LocalViews views{campos, radii, colors, nullptr, N};
    launch_split_fwd((cudaStream_t)stream, degree, C, N, K, means, sh0, shN, views, q0);
    return check_launch("sh_view_split_quant_fwd");
}

extern "C" int gsb200_sh_view_split_quant_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                              const float *campos, const float *sh0, const gsb200_quant_load *q_sh0, const float *shN,
                                              const int32_t *radii, const float *v_colors, float *v_sh0, float *v_shN,
                                              float *v_means, float *v_campos, gsb200_stream_t stream) {
    const char *who = "gsb200_sh_view_split_quant_bwd";
    if (int rc = check_split_args(who, K, degree, sh0, shN)) return rc;
    QuantLoad q0;
    if (int rc = make_quant_load(who, q_sh0, q0)) return rc;
    if (q0.on && q0.act != 0) return failf(2, "%s: sh0 takes no activation", who);
    if (C == 0 || N == 0) return 0;
    if (v_sh0 == nullptr || (K > 1 && v_shN == nullptr)) return failf(2, "%s: null gradient pointer", who);
    if ((reinterpret_cast<uintptr_t>(v_shN) & 15) != 0) return failf(2, "%s: v_shN must be 16-byte aligned", who);
    cudaStream_t s = (cudaStream_t)stream;
    if (v_campos != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_campos, 0, sizeof(float) * 3 * C, s));
    const LocalViews views{campos, radii, nullptr, v_colors, N};
    launch_split_bwd(s, degree, C, N, K, means, sh0, shN, views, v_sh0, v_shN, v_means, v_campos, q0);
    return check_launch("sh_view_split_quant_bwd");
}

static int make_peer_views(const char *who, uint32_t V, uint32_t n_local, uint32_t n_block, const float *campos, const int32_t *masks,
                           const float *v_colors, float *const *view_colors, PeerViews &views) {
    if (V == 0 || V > (uint32_t)kMaxViews) return failf(2, "%s: 1..%d views", who, kMaxViews);
    if (n_local > n_block) return failf(2, "%s: n_local > n_block", who);
    if (campos == nullptr || masks == nullptr) return failf(2, "%s: null campos / masks", who);
    views = PeerViews{};
    views.campos = campos;
    views.masks = masks;
    views.v_colors = v_colors;
    views.n_block = n_block;
    for (uint32_t v = 0; v < V; ++v) views.colors[v] = view_colors ? view_colors[v] : nullptr;
    return 0;
}

extern "C" int gsb200_sh_owner_fwd(uint32_t V, uint32_t n_local, uint32_t n_block, uint32_t K, uint32_t degree, const float *means,
                                   const float *campos, const float *sh0, const float *shN, const int32_t *masks,
                                   float *const *view_colors, gsb200_stream_t stream) {
    const char *who = "gsb200_sh_owner_fwd";
    if (int rc = check_split_args(who, K, degree, sh0, shN)) return rc;
    PeerViews views;
    if (int rc = make_peer_views(who, V, n_local, n_block, campos, masks, nullptr, view_colors, views)) return rc;
    if (view_colors == nullptr) return failf(2, "%s: null colour table", who);
    for (uint32_t v = 0; v < V; ++v)
        if (views.colors[v] == nullptr) return failf(2, "%s: null colour pointer", who);
    if (n_local == 0) return 0;
    launch_split_fwd((cudaStream_t)stream, degree, V, n_local, K, means, sh0, shN, views);
    return check_launch("sh_owner_fwd");
}

extern "C" int gsb200_sh_owner_bwd(uint32_t V, uint32_t n_local, uint32_t n_block, uint32_t K, uint32_t degree, const float *means,
                                   const float *campos, const float *sh0, const float *shN, const int32_t *masks,
                                   const float *v_colors, float *v_sh0, float *v_shN, float *v_means, gsb200_stream_t stream) {
    const char *who = "gsb200_sh_owner_bwd";
    if (int rc = check_split_args(who, K, degree, sh0, shN)) return rc;
    PeerViews views;
    if (int rc = make_peer_views(who, V, n_local, n_block, campos, masks, v_colors, nullptr, views)) return rc;
    if (v_colors == nullptr) return failf(2, "%s: null v_colors", who);
    if (n_local == 0) return 0;
    if (v_sh0 == nullptr || (K > 1 && v_shN == nullptr)) return failf(2, "%s: null gradient pointer", who);
    if ((reinterpret_cast<uintptr_t>(v_shN) & 15) != 0) return failf(2, "%s: v_shN must be 16-byte aligned", who);
    launch_split_bwd((cudaStream_t)stream, degree, V, n_local, K, means, sh0, shN, views, v_sh0, v_shN, v_means, nullptr);
    return check_launch("sh_owner_bwd");
}

extern "C" int gsb200_shard_push_rows(uint32_t world, uint32_t n_block, uint32_t width, uint32_t n_slots, uint32_t slot0,
                                      uint64_t slot_floats, const float *src, float *const *peer_staging, const float *extra_src,
                                      uint32_t extra_w, float *const *peer_extra, gsb200_stream_t stream) {
    const char *who = "gsb200_shard_push_rows";
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "%s: world must be in [1, %d]", who, kMaxShards);
    if (n_block % 4 != 0 || width == 0) return failf(2, "%s: n_block must be a multiple of 4 and width > 0", who);
    if ((uint64_t)n_block * width > slot_floats || slot_floats % 4 != 0) return failf(2, "%s: staging slot too small or unaligned", who);
    if (n_slots == 0 || n_slots > 65535u) return failf(2, "%s: 1..65535 slots per call", who);
    if (extra_src != nullptr && (extra_w == 0 || extra_w > (uint32_t)kThreads || peer_extra == nullptr))
        return failf(2, "%s: bad extra payload", who);
    if (n_block == 0) return 0;
    if (src == nullptr || (reinterpret_cast<uintptr_t>(src) & 15) != 0) return failf(2, "%s: src must be non-null and 16-byte aligned", who);
    PeerRows peers{}, extras{};
    for (uint32_t r = 0; r < world; ++r) {
        peers.base[r] = peer_staging[r];
        if (peers.base[r] == nullptr || (reinterpret_cast<uintptr_t>(peers.base[r]) & 15) != 0)
            return failf(2, "%s: staging pointers must be non-null and 16-byte aligned", who);
        if (extra_src != nullptr) {
            extras.base[r] = peer_extra[r];
            if (extras.base[r] == nullptr) return failf(2, "%s: null extra pointer", who);
        }
    }
    const uint64_t n_vec = (uint64_t)n_block * width / 4;
    const dim3 grid((uint32_t)((n_vec + kThreads * kGeoUnroll - 1) / (kThreads * kGeoUnroll)), world, n_slots);
    shard_push_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n_block, width, slot0, slot_floats, src,
                                                                        (uint64_t)world * n_block * width, peers, extra_src, extra_w,
                                                                        extras);
    return check_launch("shard_push_rows");
}

extern "C" int gsb200_peer_push_small(uint32_t world, uint32_t slot0, uint32_t n_slots, uint32_t w, const float *src,
                                      float *const *peer_dst, gsb200_stream_t stream) {
    const char *who = "gsb200_peer_push_small";
    if (world == 0 || world > (uint32_t)kMaxShards) return failf(2, "%s: world must be in [1, %d]", who, kMaxShards);
    if ((uint64_t)n_slots * w > 65536u) return failf(2, "%s: at most 65536 floats", who);
    if (n_slots == 0 || w == 0) return 0;
    if (src == nullptr) return failf(2, "%s: null source", who);
    PeerRows peers{};
    for (uint32_t r = 0; r < world; ++r) {
        peers.base[r] = peer_dst[r];
        if (peers.base[r] == nullptr) return failf(2, "%s: null destination", who);
    }
    const uint32_t n = n_slots * w;
    peer_push_small_kernel<<<dim3(cdiv(n, 128u), world), 128, 0, (cudaStream_t)stream>>>(n, (uint64_t)slot0 * w, src, peers);
    return check_launch("peer_push_small");
}
