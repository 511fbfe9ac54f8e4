// Un-fused public geometry ops: the pieces of the projection stage as separate,
// differentiable operators, for callers that compose them by hand (gsplat/__init__.py:7-20 exports them; the
// reference's tests/test_basic.py:53-173 exercise them against _torch_impl.py).
//
//   gsb200_quat_scale_to_covar_preci_{fwd,bwd}  replaces _C.quat_scale_to_covar_preci_{fwd,bwd}
//       (gsplat/cuda/csrc/quat_scale_to_covar_preci_{fwd,bwd}.cu, include/quat_scale_to_covar_preci.cuh:9-124)
//   gsb200_proj_{fwd,bwd}                       replaces _C.proj_{fwd,bwd}   (csrc/proj_{fwd,bwd}.cu, include/proj.cuh)
//   gsb200_world_to_cam_{fwd,bwd}               replaces _C.world_to_cam_{fwd,bwd} (csrc/world_to_cam_{fwd,bwd}.cu,
//       include/transform.cuh:9-71)
//
// Covariances here are full row-major 3x3 (2x2) matrices that need not be symmetric, as in the reference ops.
// All kernels are one-thread-per-element streaming kernels; the world->camera backward gives one thread a
// Gaussian and loops over the cameras, so the sums over cameras stay in registers (no atomics) and the sums
// over Gaussians (v_viewmats) are block-reduced before one atomic per block, camera and entry.
#include "common.cuh"
#include "geom.cuh"

namespace gsb200 {

__device__ __forceinline__ void load9(const float *p, float m[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m[i][j] = __ldg(p + i * 3 + j);
}

__device__ __forceinline__ void store_covar(float *out, const Sym3 &S, bool triu) {
    if (triu) {
        out[0] = S.xx; out[1] = S.xy; out[2] = S.xz; out[3] = S.yy; out[4] = S.yz; out[5] = S.zz;
    } else {
        out[0] = S.xx; out[1] = S.xy; out[2] = S.xz;
        out[3] = S.xy; out[4] = S.yy; out[5] = S.yz;
        out[6] = S.xz; out[7] = S.yz; out[8] = S.zz;
    }
}

__global__ void __launch_bounds__(kThreads)
quat_scale_fwd_kernel(uint32_t N, const float *__restrict__ quats, const float *__restrict__ scales, bool triu,
                      float *__restrict__ covars, float *__restrict__ precis) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    const float4 q = ldg4(quats + (size_t)g * 4);
    const Mat3 Rq = quat_to_rot(q.x, q.y, q.z, q.w, nullptr);
    const float sx = __ldg(scales + (size_t)g * 3), sy = __ldg(scales + (size_t)g * 3 + 1), sz = __ldg(scales + (size_t)g * 3 + 2);
    const uint32_t stride = triu ? 6 : 9;
    if (covars != nullptr) store_covar(covars + (size_t)g * stride, covar_from_quat_scale(Rq, sx, sy, sz), triu);
    if (precis != nullptr) store_covar(precis + (size_t)g * stride, covar_from_quat_scale(Rq, 1.f / sx, 1.f / sy, 1.f / sz), triu);
}

// gradient of a (possibly triu-packed) symmetric-product output: V + V^T
__device__ __forceinline__ void load_sym_grad(const float *v, bool triu, float G[3][3]) {
    float V[3][3];
    if (triu) {  // the packed entry stands for the mean of the two mirrored entries (quat_scale_to_covar_preci_bwd.cu:56-68)
        V[0][0] = v[0]; V[0][1] = V[1][0] = 0.5f * v[1]; V[0][2] = V[2][0] = 0.5f * v[2];
        V[1][1] = v[3]; V[1][2] = V[2][1] = 0.5f * v[4]; V[2][2] = v[5];
    } else {
        load9(v, V);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) G[i][j] = V[i][j] + V[j][i];
}

// D = M M^T with M = Rq diag(s): accumulates v_Rq and returns d/ds_j = sum_i Rq[i][j] vM[i][j]
__device__ __forceinline__ void mmt_vjp(const Mat3 &Rq, const float s[3], const float G[3][3], float vRq[3][3], float vs[3]) {
    float M[3][3], vM[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i][j] = Rq.m[i][j] * s[j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vM[i][j] = G[i][0] * M[0][j] + G[i][1] * M[1][j] + G[i][2] * M[2][j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        vs[j] = Rq.m[0][j] * vM[0][j] + Rq.m[1][j] * vM[1][j] + Rq.m[2][j] * vM[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i) vRq[i][j] += vM[i][j] * s[j];
    }
}

__global__ void __launch_bounds__(kThreads)
quat_scale_bwd_kernel(uint32_t N, const float *__restrict__ quats, const float *__restrict__ scales, bool triu,
                      const float *__restrict__ v_covars, const float *__restrict__ v_precis, float *__restrict__ v_quats,
                      float *__restrict__ v_scales) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= N) return;
    const float4 q = ldg4(quats + (size_t)g * 4);
    float inv_norm;
    const Mat3 Rq = quat_to_rot(q.x, q.y, q.z, q.w, &inv_norm);
    const float s[3] = {__ldg(scales + (size_t)g * 3), __ldg(scales + (size_t)g * 3 + 1), __ldg(scales + (size_t)g * 3 + 2)};
    const uint32_t stride = triu ? 6 : 9;
    float vRq[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float vs[3] = {0.f, 0.f, 0.f};
    float G[3][3];
    if (v_covars != nullptr) {
        load_sym_grad(v_covars + (size_t)g * stride, triu, G);
        float d[3];
        mmt_vjp(Rq, s, G, vRq, d);
        vs[0] += d[0]; vs[1] += d[1]; vs[2] += d[2];
    }
    if (v_precis != nullptr) {
        load_sym_grad(v_precis + (size_t)g * stride, triu, G);
        const float r[3] = {1.f / s[0], 1.f / s[1], 1.f / s[2]};
        float d[3];
        mmt_vjp(Rq, r, G, vRq, d);
        vs[0] += -r[0] * r[0] * d[0]; vs[1] += -r[1] * r[1] * d[1]; vs[2] += -r[2] * r[2] * d[2];
    }
    // rotation matrix of the NORMALISED quaternion, then the normalisation (quat.cuh:35-57)
    const float w = q.x * inv_norm, x = q.y * inv_norm, y = q.z * inv_norm, z = q.w * inv_norm;
    const float vw = 2.f * (x * (vRq[2][1] - vRq[1][2]) + y * (vRq[0][2] - vRq[2][0]) + z * (vRq[1][0] - vRq[0][1]));
    const float vx = 2.f * (-2.f * x * (vRq[1][1] + vRq[2][2]) + y * (vRq[1][0] + vRq[0][1]) + z * (vRq[2][0] + vRq[0][2]) +
                            w * (vRq[2][1] - vRq[1][2]));
    const float vy = 2.f * (x * (vRq[1][0] + vRq[0][1]) - 2.f * y * (vRq[0][0] + vRq[2][2]) + z * (vRq[2][1] + vRq[1][2]) +
                            w * (vRq[0][2] - vRq[2][0]));
    const float vz = 2.f * (x * (vRq[2][0] + vRq[0][2]) + y * (vRq[2][1] + vRq[1][2]) - 2.f * z * (vRq[0][0] + vRq[1][1]) +
                            w * (vRq[1][0] - vRq[0][1]));
    const float dotp = vw * w + vx * x + vy * y + vz * z;
    stg4(v_quats + (size_t)g * 4, make_float4((vw - dotp * w) * inv_norm, (vx - dotp * x) * inv_norm, (vy - dotp * y) * inv_norm,
                                               (vz - dotp * z) * inv_norm));
    v_scales[(size_t)g * 3] = vs[0]; v_scales[(size_t)g * 3 + 1] = vs[1]; v_scales[(size_t)g * 3 + 2] = vs[2];
}

// ------------------------------------------------------------------------------------------ proj
__global__ void __launch_bounds__(kThreads)
proj_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ Ks, float W, float H, int model, float *__restrict__ means2d,
                float *__restrict__ covars2d) {
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (size_t)C * N) return;
    const uint32_t c = idx / N;
    const float *K = Ks + (size_t)c * 9;
    const float x = __ldg(means + idx * 3), y = __ldg(means + idx * 3 + 1), z = __ldg(means + idx * 3 + 2);
    float S[3][3];
    load9(covars + idx * 9, S);
    const CamProj P = project_point(model, x, y, z, __ldg(K), __ldg(K + 4), __ldg(K + 2), __ldg(K + 5), W, H);
    float B[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[r][j] = P.J[r][0] * S[0][j] + P.J[r][1] * S[1][j] + P.J[r][2] * S[2][j];
    float o[4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) o[r * 2 + q] = B[r][0] * P.J[q][0] + B[r][1] * P.J[q][1] + B[r][2] * P.J[q][2];
    *reinterpret_cast<float2 *>(means2d + idx * 2) = make_float2(P.mx, P.my);
    stg4(covars2d + idx * 4, make_float4(o[0], o[1], o[2], o[3]));
}

__global__ void __launch_bounds__(kThreads)
proj_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ Ks, float W, float H, int model, const float *__restrict__ v_means2d,
                const float *__restrict__ v_covars2d, float *__restrict__ v_means, float *__restrict__ v_covars) {
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (size_t)C * N) return;
    const uint32_t c = idx / N;
    const float *K = Ks + (size_t)c * 9;
    const float fx = __ldg(K), fy = __ldg(K + 4), cx = __ldg(K + 2), cy = __ldg(K + 5);
    const float x = __ldg(means + idx * 3), y = __ldg(means + idx * 3 + 1), z = __ldg(means + idx * 3 + 2);
    float S[3][3];
    load9(covars + idx * 9, S);
    const CamProj P = project_point(model, x, y, z, fx, fy, cx, cy, W, H);
    const float4 g4 = ldg4(v_covars2d + idx * 4);
    const float G[2][2] = {{g4.x, g4.y}, {g4.z, g4.w}};
    const float vm0 = __ldg(v_means2d + idx * 2), vm1 = __ldg(v_means2d + idx * 2 + 1);
    float GJ[2][3], GtJ[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        GJ[0][j] = G[0][0] * P.J[0][j] + G[0][1] * P.J[1][j];
        GJ[1][j] = G[1][0] * P.J[0][j] + G[1][1] * P.J[1][j];
        GtJ[0][j] = G[0][0] * P.J[0][j] + G[1][0] * P.J[1][j];
        GtJ[1][j] = G[0][1] * P.J[0][j] + G[1][1] * P.J[1][j];
    }
    // v_Sigma = J^T G J
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) v_covars[idx * 9 + i * 3 + j] = P.J[0][i] * GJ[0][j] + P.J[1][i] * GJ[1][j];
    // v_J = G J Sigma^T + G^T J Sigma
    float vJ[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vJ[r][j] = GJ[r][0] * S[j][0] + GJ[r][1] * S[j][1] + GJ[r][2] * S[j][2] + GtJ[r][0] * S[0][j] + GtJ[r][1] * S[1][j] +
                       GtJ[r][2] * S[2][j];
    float vmc[3];
    if (model == GSB200_CAMERA_PINHOLE) {  // proj.cuh:121-199
        const float rz = P.rz, rz2 = rz * rz, rz3 = rz2 * rz;
        vmc[0] = fx * rz * vm0;
        vmc[1] = fy * rz * vm1;
        vmc[2] = -(fx * x * vm0 + fy * y * vm1) * rz2;
        if (!P.clip_x) vmc[0] += -fx * rz2 * vJ[0][2];
        else vmc[2] += -fx * rz3 * vJ[0][2] * P.tx;
        if (!P.clip_y) vmc[1] += -fy * rz2 * vJ[1][2];
        else vmc[2] += -fy * rz3 * vJ[1][2] * P.ty;
        vmc[2] += -fx * rz2 * vJ[0][0] - fy * rz2 * vJ[1][1] + 2.f * fx * P.tx * rz3 * vJ[0][2] + 2.f * fy * P.ty * rz3 * vJ[1][2];
    } else if (model == GSB200_CAMERA_FISHEYE) {  // exact derivative of the fisheye map by forward-mode duals
        Dual3 m[2], Jd[2][3];
        fisheye_dual(x, y, z, fx, fy, cx, cy, m, Jd);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float acc = vm0 * m[0].d[k] + vm1 * m[1].d[k];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc += vJ[r][j] * Jd[r][j].d[k];
            vmc[k] = acc;
        }
    } else {
        vmc[0] = fx * vm0; vmc[1] = fy * vm1; vmc[2] = 0.f;
    }
    v_means[idx * 3] = vmc[0]; v_means[idx * 3 + 1] = vmc[1]; v_means[idx * 3 + 2] = vmc[2];
}

// ------------------------------------------------------------------------------------------ world -> camera
__global__ void __launch_bounds__(kThreads)
world_to_cam_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                        const float *__restrict__ viewmats, float *__restrict__ means_c, float *__restrict__ covars_c) {
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (size_t)C * N) return;
    const uint32_t c = idx / N, g = idx - (size_t)c * N;
    const float *V = viewmats + (size_t)c * 16;
    float R[3][3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        R[i][0] = __ldg(V + i * 4); R[i][1] = __ldg(V + i * 4 + 1); R[i][2] = __ldg(V + i * 4 + 2); t[i] = __ldg(V + i * 4 + 3);
    }
    const float p[3] = {__ldg(means + (size_t)g * 3), __ldg(means + (size_t)g * 3 + 1), __ldg(means + (size_t)g * 3 + 2)};
#pragma unroll
    for (int i = 0; i < 3; ++i) means_c[idx * 3 + i] = R[i][0] * p[0] + R[i][1] * p[1] + R[i][2] * p[2] + t[i];
    float S[3][3], A[3][3];
    load9(covars + (size_t)g * 9, S);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = R[i][0] * S[0][j] + R[i][1] * S[1][j] + R[i][2] * S[2][j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int l = 0; l < 3; ++l) covars_c[idx * 9 + i * 3 + l] = A[i][0] * R[l][0] + A[i][1] * R[l][1] + A[i][2] * R[l][2];
}

__global__ void __launch_bounds__(kThreads)
world_to_cam_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                        const float *__restrict__ viewmats, const float *__restrict__ v_means_c,
                        const float *__restrict__ v_covars_c, float *__restrict__ v_means, float *__restrict__ v_covars,
                        float *__restrict__ v_viewmats) {
    __shared__ float s_red[kThreads / 32][12];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool on = g < N;
    float p[3] = {0.f, 0.f, 0.f}, S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i][j] = 0.f;
    if (on) {
        p[0] = __ldg(means + (size_t)g * 3); p[1] = __ldg(means + (size_t)g * 3 + 1); p[2] = __ldg(means + (size_t)g * 3 + 2);
        load9(covars + (size_t)g * 9, S);
    }
    float vp[3] = {0.f, 0.f, 0.f}, vS[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vS[i][j] = 0.f;
    for (uint32_t c = 0; c < C; ++c) {
        const float *Vm = viewmats + (size_t)c * 16;
        float R[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { R[i][0] = __ldg(Vm + i * 4); R[i][1] = __ldg(Vm + i * 4 + 1); R[i][2] = __ldg(Vm + i * 4 + 2); }
        float vmc[3] = {0.f, 0.f, 0.f}, V[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) V[i][j] = 0.f;
        if (on) {
            const size_t idx = (size_t)c * N + g;
            vmc[0] = __ldg(v_means_c + idx * 3); vmc[1] = __ldg(v_means_c + idx * 3 + 1); vmc[2] = __ldg(v_means_c + idx * 3 + 2);
            load9(v_covars_c + idx * 9, V);
        }
        // v_p += R^T v_mc ; v_Sigma += R^T V R
#pragma unroll
        for (int j = 0; j < 3; ++j) vp[j] += R[0][j] * vmc[0] + R[1][j] * vmc[1] + R[2][j] * vmc[2];
        float VR[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) VR[i][j] = V[i][0] * R[0][j] + V[i][1] * R[1][j] + V[i][2] * R[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) vS[i][j] += R[0][i] * VR[0][j] + R[1][i] * VR[1][j] + R[2][i] * VR[2][j];
        if (v_viewmats != nullptr) {
            // v_R = v_mc p^T + V R Sigma^T + V^T R Sigma ; v_t = v_mc
            float VtR[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) VtR[i][j] = V[0][i] * R[0][j] + V[1][i] * R[1][j] + V[2][i] * R[2][j];
            float vals[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    vals[i * 4 + j] = warp_sum(vmc[i] * p[j] + VR[i][0] * S[j][0] + VR[i][1] * S[j][1] + VR[i][2] * S[j][2] +
                                               VtR[i][0] * S[0][j] + VtR[i][1] * S[1][j] + VtR[i][2] * S[2][j]);
                vals[i * 4 + 3] = warp_sum(vmc[i]);
            }
            __syncthreads();
            if (lane == 0)
#pragma unroll
                for (int k = 0; k < 12; ++k) s_red[warp][k] = vals[k];
            __syncthreads();
            if (threadIdx.x < 12) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kThreads / 32; ++w) s += s_red[w][threadIdx.x];
                if (s != 0.f) atomicAdd(v_viewmats + (size_t)c * 16 + threadIdx.x, s);
            }
        }
    }
    if (!on) return;
    if (v_means != nullptr) { v_means[(size_t)g * 3] = vp[0]; v_means[(size_t)g * 3 + 1] = vp[1]; v_means[(size_t)g * 3 + 2] = vp[2]; }
    if (v_covars != nullptr)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v_covars[(size_t)g * 9 + i * 3 + j] = vS[i][j];
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_quat_scale_to_covar_preci_fwd(uint32_t N, const float *quats, const float *scales, int triu, float *covars,
                                                    float *precis, gsb200_stream_t stream) {
    if (N == 0 || (covars == nullptr && precis == nullptr)) return 0;
    if (quats == nullptr || scales == nullptr) return fail(2, "gsb200_quat_scale_to_covar_preci_fwd: null input");
    if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0) return fail(2, "gsb200_quat_scale_to_covar_preci_fwd: quats must be 16-byte aligned");
    quat_scale_fwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, quats, scales, triu != 0, covars, precis);
    return check_launch("quat_scale_to_covar_preci_fwd");
}

extern "C" int gsb200_quat_scale_to_covar_preci_bwd(uint32_t N, const float *quats, const float *scales, int triu,
                                                    const float *v_covars, const float *v_precis, float *v_quats, float *v_scales,
                                                    gsb200_stream_t stream) {
    if (N == 0) return 0;
    if (quats == nullptr || scales == nullptr || v_quats == nullptr || v_scales == nullptr)
        return fail(2, "gsb200_quat_scale_to_covar_preci_bwd: null pointer");
    if (((reinterpret_cast<uintptr_t>(quats) | reinterpret_cast<uintptr_t>(v_quats)) & 15) != 0)
        return fail(2, "gsb200_quat_scale_to_covar_preci_bwd: quats / v_quats must be 16-byte aligned");
    quat_scale_bwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, quats, scales, triu != 0, v_covars, v_precis,
                                                                                    v_quats, v_scales);
    return check_launch("quat_scale_to_covar_preci_bwd");
}

static int check_cn(const char *who, uint32_t C, uint32_t N, int model) {
    if ((uint64_t)C * N > 0x7fffffffull * kThreads) return failf(2, "%s: C * N too large", who);
    if (model != GSB200_CAMERA_PINHOLE && model != GSB200_CAMERA_ORTHO && model != GSB200_CAMERA_FISHEYE)
        return failf(2, "%s: unknown camera model %d", who, model);
    return 0;
}

extern "C" int gsb200_proj_fwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks, uint32_t width,
                               uint32_t height, int camera_model, float *means2d, float *covars2d, gsb200_stream_t stream) {
    if (int rc = check_cn("gsb200_proj_fwd", C, N, camera_model)) return rc;
    if (C == 0 || N == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(covars2d) & 15) != 0 || (reinterpret_cast<uintptr_t>(means2d) & 7) != 0)
        return fail(2, "gsb200_proj_fwd: outputs must be 16-byte aligned");
    const uint64_t total = (uint64_t)C * N;
    proj_fwd_kernel<<<(uint32_t)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, Ks, (float)width, (float)height, camera_model, means2d, covars2d);
    return check_launch("proj_fwd");
}

extern "C" int gsb200_proj_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks, uint32_t width,
                               uint32_t height, int camera_model, const float *v_means2d, const float *v_covars2d, float *v_means,
                               float *v_covars, gsb200_stream_t stream) {
    if (int rc = check_cn("gsb200_proj_bwd", C, N, camera_model)) return rc;
    if (C == 0 || N == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(v_covars2d) & 15) != 0) return fail(2, "gsb200_proj_bwd: v_covars2d must be 16-byte aligned");
    const uint64_t total = (uint64_t)C * N;
    proj_bwd_kernel<<<(uint32_t)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, Ks, (float)width, (float)height, camera_model, v_means2d, v_covars2d, v_means, v_covars);
    return check_launch("proj_bwd");
}

extern "C" int gsb200_world_to_cam_fwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *viewmats,
                                       float *means_c, float *covars_c, gsb200_stream_t stream) {
    if (int rc = check_cn("gsb200_world_to_cam_fwd", C, N, GSB200_CAMERA_PINHOLE)) return rc;
    if (C == 0 || N == 0) return 0;
    const uint64_t total = (uint64_t)C * N;
    world_to_cam_fwd_kernel<<<(uint32_t)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, viewmats, means_c, covars_c);
    return check_launch("world_to_cam_fwd");
}

extern "C" int gsb200_world_to_cam_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *viewmats,
                                       const float *v_means_c, const float *v_covars_c, float *v_means, float *v_covars,
                                       float *v_viewmats, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (v_viewmats != nullptr && C > 0) GSB_CUDA_OK(cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * C, s));
    if (N == 0) return 0;
    if (C == 0) {
        if (v_means != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_means, 0, sizeof(float) * 3 * N, s));
        if (v_covars != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_covars, 0, sizeof(float) * 9 * N, s));
        return 0;
    }
    world_to_cam_bwd_kernel<<<cdiv(N, kThreads), kThreads, 0, s>>>(C, N, means, covars, viewmats, v_means_c, v_covars_c, v_means,
                                                                   v_covars, v_viewmats);
    return check_launch("world_to_cam_bwd");
}
