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
template <int DEG>
__global__ void __launch_bounds__(kThreads)
sh_view_fwd_kernel(uint32_t C, uint32_t N, uint32_t K, const float *__restrict__ means, const float *__restrict__ campos,
                   const float *__restrict__ coeffs, const int32_t *__restrict__ radii, float *__restrict__ colors, bool vec) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    constexpr uint32_t nb = (DEG + 1) * (DEG + 1);
    bool any = false;
    for (uint32_t c = 0; c < C; ++c) any |= radii[(size_t)c * N + g] > 0;
    float cf[nb * 3 + 3];
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (any) {
        const float *row = coeffs + (size_t)g * K * 3;
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
        mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2];
    }
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

template <int DEG>
__global__ void __launch_bounds__(kThreads)
sh_view_bwd_kernel(uint32_t C, uint32_t N, uint32_t K, const float *__restrict__ means, const float *__restrict__ campos,
                   const float *__restrict__ coeffs, const int32_t *__restrict__ radii, const float *__restrict__ v_colors,
                   float *__restrict__ v_coeffs, float *__restrict__ v_means, float *__restrict__ v_campos, bool vec) {
    __shared__ float s_red[kThreads / 32][3];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
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
    if (any) {
        const float *row = coeffs + (size_t)g * K * 3;
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
        mx = means[(size_t)g * 3]; my = means[(size_t)g * 3 + 1]; mz = means[(size_t)g * 3 + 2];
    }
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
            if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = sx; s_red[threadIdx.x >> 5][1] = sy; s_red[threadIdx.x >> 5][2] = sz; }
            __syncthreads();
            if (threadIdx.x < 3) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < kThreads / 32; ++w) t += s_red[w][threadIdx.x];
                if (t != 0.f) atomicAdd(v_campos + c * 3 + threadIdx.x, -t);
            }
        }
    }
    if (g >= N) return;
    float *vrow = v_coeffs + (size_t)g * nfl;
    if (vec) {
#pragma unroll
        for (uint32_t k = 0; k + 3 < nb * 3; k += 4) stg4(vrow + k, make_float4(vc[k], vc[k + 1], vc[k + 2], vc[k + 3]));
        for (uint32_t k = nb * 3; k < nfl; k += 4) stg4(vrow + k, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
#pragma unroll
        for (uint32_t k = 0; k < nb * 3; ++k) vrow[k] = vc[k];
        for (uint32_t k = nb * 3; k < nfl; ++k) vrow[k] = 0.f;
    }
    if (v_means != nullptr) {
        v_means[(size_t)g * 3] = vmx; v_means[(size_t)g * 3 + 1] = vmy; v_means[(size_t)g * 3 + 2] = vmz;
    }
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

extern "C" int gsb200_sh_view_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                  const float *coeffs, const int32_t *radii, float *colors, gsb200_stream_t stream) {
    if (degree > 4) return fail(2, "gsb200_sh_view_fwd: degree must be <= 4");
    if ((degree + 1) * (degree + 1) > K) return fail(2, "gsb200_sh_view_fwd: coeffs hold fewer than (degree+1)^2 bases");
    if (C == 0 || N == 0) return 0;
    const bool vec = sh_vec_ok(K, degree, coeffs, nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t grid = cdiv(N, kThreads);
    switch (degree) {
        case 0: sh_view_fwd_kernel<0><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, colors, vec); break;
        case 1: sh_view_fwd_kernel<1><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, colors, vec); break;
        case 2: sh_view_fwd_kernel<2><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, colors, vec); break;
        case 3: sh_view_fwd_kernel<3><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, colors, vec); break;
        default: sh_view_fwd_kernel<4><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, colors, vec); break;
    }
    return check_launch("sh_view_fwd");
}

extern "C" int gsb200_sh_view_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means, const float *campos,
                                  const float *coeffs, const int32_t *radii, const float *v_colors, float *v_coeffs,
                                  float *v_means, float *v_campos, gsb200_stream_t stream) {
    if (degree > 4) return fail(2, "gsb200_sh_view_bwd: degree must be <= 4");
    if ((degree + 1) * (degree + 1) > K) return fail(2, "gsb200_sh_view_bwd: coeffs hold fewer than (degree+1)^2 bases");
    if (C == 0 || N == 0) return 0;
    const bool vec = sh_vec_ok(K, degree, coeffs, v_coeffs);
    cudaStream_t s = (cudaStream_t)stream;
    if (v_campos != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_campos, 0, sizeof(float) * 3 * C, s));
    const uint32_t grid = cdiv(N, kThreads);
    switch (degree) {
        case 0: sh_view_bwd_kernel<0><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, v_colors, v_coeffs, v_means, v_campos, vec); break;
        case 1: sh_view_bwd_kernel<1><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, v_colors, v_coeffs, v_means, v_campos, vec); break;
        case 2: sh_view_bwd_kernel<2><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, v_colors, v_coeffs, v_means, v_campos, vec); break;
        case 3: sh_view_bwd_kernel<3><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, v_colors, v_coeffs, v_means, v_campos, vec); break;
        default: sh_view_bwd_kernel<4><<<grid, kThreads, 0, s>>>(C, N, K, means, campos, coeffs, radii, v_colors, v_coeffs, v_means, v_campos, vec); break;
    }
    return check_launch("sh_view_bwd");
}
