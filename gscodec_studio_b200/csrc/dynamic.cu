// Temporal pre-transform of Spacetime-Gaussian (STG) splats: the caller side of the hot path for the dynamic
// trainer (BASELINE config 4).
//
// Reference: examples/simple_trainer_dyngs.py:476-520 (same chain in simple_trainer_STG.py and :1180-1215) builds
// the per-timestamp splats with ~20 elementwise torch kernels forward and as many again backward:
//     d        = t - trbf_center                        tforpoly = d.detach()
//     trbf     = exp(-(d / (sqrt(2) exp(trbf_scale)))^2)
//     opacity  = sigmoid(opacities) * trbf
//     means_t  = means + motion[0:3] d + motion[3:6] d^2 + motion[6:9] d^3        (d detached)
//     quats_t  = normalize(quats + d * omega)                                     (d detached)
//     scales   = exp(scales)
// Here: one forward and one backward kernel, every attribute read once and every output / gradient written once.
#include "common.cuh"

namespace gsb200 {

__global__ void __launch_bounds__(kThreads)
stg_fwd_kernel(uint32_t N, float t, const float *__restrict__ means, const float *__restrict__ quats,
               const float *__restrict__ scales_raw, const float *__restrict__ opac_raw, const float *__restrict__ trbf_center,
               const float *__restrict__ trbf_scale_raw, const float *__restrict__ motion, const float *__restrict__ omega,
               float *__restrict__ means_t, float *__restrict__ quats_t, float *__restrict__ scales, float *__restrict__ opacities,
               float *__restrict__ trbf) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    const float d = t - __ldg(trbf_center + g);
    const float sigma = expf(__ldg(trbf_scale_raw + g));
    const float x = d / (1.4142135623730951f * sigma);
    const float rbf = expf(-x * x);
    trbf[g] = rbf;
    opacities[g] = rbf / (1.f + expf(-__ldg(opac_raw + g)));
    const float d2 = d * d, d3 = d2 * d;
    const float *mo = motion + (size_t)g * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        means_t[(size_t)g * 3 + k] = __ldg(means + (size_t)g * 3 + k) + __ldg(mo + k) * d + __ldg(mo + 3 + k) * d2 + __ldg(mo + 6 + k) * d3;
        scales[(size_t)g * 3 + k] = expf(__ldg(scales_raw + (size_t)g * 3 + k));
    }
    const float4 q = ldg4(quats + (size_t)g * 4), w = ldg4(omega + (size_t)g * 4);
    const float ux = q.x + d * w.x, uy = q.y + d * w.y, uz = q.z + d * w.z, uw = q.w + d * w.w;
    const float inv = 1.f / fmaxf(sqrtf(ux * ux + uy * uy + uz * uz + uw * uw), 1e-12f);  // F.normalize
    stg4(quats_t + (size_t)g * 4, make_float4(ux * inv, uy * inv, uz * inv, uw * inv));
}

__global__ void __launch_bounds__(kThreads)
stg_bwd_kernel(uint32_t N, float t, const float *__restrict__ quats, const float *__restrict__ scales_raw,
               const float *__restrict__ opac_raw, const float *__restrict__ trbf_center, const float *__restrict__ trbf_scale_raw,
               const float *__restrict__ omega, const float *__restrict__ v_means_t, const float *__restrict__ v_quats_t,
               const float *__restrict__ v_scales, const float *__restrict__ v_opacities, float *__restrict__ v_means,
               float *__restrict__ v_quats, float *__restrict__ v_scales_raw, float *__restrict__ v_opac_raw,
               float *__restrict__ v_trbf_center, float *__restrict__ v_trbf_scale_raw, float *__restrict__ v_motion,
               float *__restrict__ v_omega) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    const float d = t - __ldg(trbf_center + g);
    const float sigma = expf(__ldg(trbf_scale_raw + g));
    const float inv_s = 1.f / (1.4142135623730951f * sigma);
    const float x = d * inv_s;
    const float rbf = expf(-x * x);
    const float o = 1.f / (1.f + expf(-__ldg(opac_raw + g)));
    const float vo = __ldg(v_opacities + g);
    v_opac_raw[g] = vo * rbf * o * (1.f - o);
    const float v_x = vo * o * rbf * (-2.f * x);
    v_trbf_center[g] = -v_x * inv_s;   // d = t - center, only through the RBF (the polynomial uses d.detach())
    v_trbf_scale_raw[g] = -v_x * x;    // x = d / (sqrt2 e^s)  ->  dx/ds = -x
    const float d2 = d * d, d3 = d2 * d;
    float *vm = v_motion + (size_t)g * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = __ldg(v_means_t + (size_t)g * 3 + k);
        v_means[(size_t)g * 3 + k] = v;
        vm[k] = v * d; vm[3 + k] = v * d2; vm[6 + k] = v * d3;
        v_scales_raw[(size_t)g * 3 + k] = __ldg(v_scales + (size_t)g * 3 + k) * expf(__ldg(scales_raw + (size_t)g * 3 + k));
    }
    const float4 q = ldg4(quats + (size_t)g * 4), w = ldg4(omega + (size_t)g * 4), vr = ldg4(v_quats_t + (size_t)g * 4);
    const float ux = q.x + d * w.x, uy = q.y + d * w.y, uz = q.z + d * w.z, uw = q.w + d * w.w;
    const float nrm = sqrtf(ux * ux + uy * uy + uz * uz + uw * uw);
    float4 vu;
    if (nrm >= 1e-12f) {
        const float inv = 1.f / nrm;
        const float rx = ux * inv, ry = uy * inv, rz = uz * inv, rw = uw * inv;
        const float dot = vr.x * rx + vr.y * ry + vr.z * rz + vr.w * rw;
        vu = make_float4((vr.x - dot * rx) * inv, (vr.y - dot * ry) * inv, (vr.z - dot * rz) * inv, (vr.w - dot * rw) * inv);
    } else {  // clamped denominator: the map is linear there
        vu = make_float4(vr.x * 1e12f, vr.y * 1e12f, vr.z * 1e12f, vr.w * 1e12f);
    }
    stg4(v_quats + (size_t)g * 4, vu);
    stg4(v_omega + (size_t)g * 4, make_float4(vu.x * d, vu.y * d, vu.z * d, vu.w * d));
}

}  // namespace gsb200

using namespace gsb200;

static bool stg_aligned(const void *a, const void *b, const void *c) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

extern "C" int gsb200_stg_temporal_fwd(uint32_t N, float timestamp, const float *means, const float *quats, const float *scales_raw,
                                       const float *opacities_raw, const float *trbf_center, const float *trbf_scale_raw,
                                       const float *motion, const float *omega, float *means_t, float *quats_t, float *scales,
                                       float *opacities, float *trbf, gsb200_stream_t stream) {
    if (N == 0) return 0;
    if (!means || !quats || !scales_raw || !opacities_raw || !trbf_center || !trbf_scale_raw || !motion || !omega || !means_t ||
        !quats_t || !scales || !opacities || !trbf)
        return fail(2, "gsb200_stg_temporal_fwd: null pointer");
    if (!stg_aligned(quats, omega, quats_t)) return fail(2, "gsb200_stg_temporal_fwd: quats / omega / quats_t must be 16-byte aligned");
    stg_fwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, timestamp, means, quats, scales_raw, opacities_raw,
                                                                             trbf_center, trbf_scale_raw, motion, omega, means_t,
                                                                             quats_t, scales, opacities, trbf);
    return check_launch("stg_temporal_fwd");
}

extern "C" int gsb200_stg_temporal_bwd(uint32_t N, float timestamp, const float *quats, const float *scales_raw,
                                       const float *opacities_raw, const float *trbf_center, const float *trbf_scale_raw,
                                       const float *omega, const float *v_means_t, const float *v_quats_t, const float *v_scales,
                                       const float *v_opacities, float *v_means, float *v_quats, float *v_scales_raw,
                                       float *v_opacities_raw, float *v_trbf_center, float *v_trbf_scale_raw, float *v_motion,
                                       float *v_omega, gsb200_stream_t stream) {
    if (N == 0) return 0;
    if (!quats || !scales_raw || !opacities_raw || !trbf_center || !trbf_scale_raw || !omega || !v_means_t || !v_quats_t || !v_scales ||
        !v_opacities || !v_means || !v_quats || !v_scales_raw || !v_opacities_raw || !v_trbf_center || !v_trbf_scale_raw || !v_motion ||
        !v_omega)
        return fail(2, "gsb200_stg_temporal_bwd: null pointer");
    if (!stg_aligned(quats, omega, v_quats_t) || !stg_aligned(v_quats, v_omega, nullptr))
        return fail(2, "gsb200_stg_temporal_bwd: quaternion arrays must be 16-byte aligned");
    stg_bwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        N, timestamp, quats, scales_raw, opacities_raw, trbf_center, trbf_scale_raw, omega, v_means_t, v_quats_t, v_scales, v_opacities,
        v_means, v_quats, v_scales_raw, v_opacities_raw, v_trbf_center, v_trbf_scale_raw, v_motion, v_omega);
    return check_launch("stg_temporal_bwd");
}
