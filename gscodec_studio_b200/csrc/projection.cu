// Stage 1 / 4b: fused 3D->2D projection of Gaussians, forward and backward (sm_100a).
//
// Behaviour follows the reference ops _C.fully_fused_projection_{fwd,bwd}
// (gsplat/cuda/csrc/fully_fused_projection_fwd.cu:22-196, fully_fused_projection_bwd.cu:24-263 and the
// device math in gsplat/cuda/include/{quat,transform,proj,utils,quat_scale_to_covar_preci}.cuh).
//
// Design differences (B200-first, not a translation):
//   * one thread owns one Gaussian and loops over the C cameras: the 40 B of Gaussian parameters
//     and the quaternion -> covariance work are read / done once instead of C times, and the
//     backward needs NO atomics and no warp-group partitions -- the sum over cameras is a register
//     accumulation followed by one plain store (deterministic);
//   * camera matrices live in shared memory (the reference re-reads them from global per thread);
//   * Gaussian parameters are read lazily: quats / scales only for Gaussians that survive the near/far
//     test in some camera (forward) or are visible in some camera (backward);
//   * culled entries are zero-filled, so downstream consumers never see uninitialised memory;
//   * v_viewmats is block-reduced in shared memory, one atomic set per block and camera.
#include "common.cuh"
#include "geom.cuh"
#include "quant_load.cuh"
#include "scan.cuh"

namespace gsb200 {

constexpr int kMaxSmemCams = 64;  // cameras cached in shared memory per pass

// MODE 0: dense outputs [C,N,...].  MODE 1 / 2: the two passes of the packed (COO) form -- memory proportional to the
// number of visible (camera, Gaussian) pairs, never C * N (reference: fully_fused_projection_packed_fwd.cu:200-247):
//   1  only counts the visible Gaussians of every (camera, block) -> pk_counts[c * n_blocks + block]
//   2  recomputes the projection and writes row pk_counts[c * n_blocks + block] (now the exclusive scan) + rank inside
//      the block; rows therefore come out in (camera, Gaussian) lexicographic order, like the reference's
// owner mode (MODE 3, multi-GPU): where view v's rows of THIS rank live inside the viewer's tensors (peer memory)
struct ProjOwnerOut {
    int32_t *radii[GSB200_MAX_VIEWS];
    float *means2d[GSB200_MAX_VIEWS];
    float *depths[GSB200_MAX_VIEWS];
    float *conics[GSB200_MAX_VIEWS];
    float *opacities[GSB200_MAX_VIEWS];
};

// MODE 4 (quantise on load): MODE 0 whose quats / scales / opacities are the trainer's RAW parameters -- the noise-mode fake
// quantisation of compression_simulation and the exp / sigmoid that follow it are applied where the values are read
// (quant_load.cuh); the activated opacities [N] (a blend input) are written once per Gaussian.
struct ProjQuant {
    QuantLoad quats, scales, opac;
    const float *opac_raw;    // [N] raw (logit) opacities, or NULL
    float *opac_act;          // fwd out / bwd in: sigmoid(fake-quantised opacity) [N]
    const float *v_opac_act;  // bwd in: gradient of opac_act [N]
    float *v_opac_raw;        // bwd out [N]
};

template <int MODE>
__global__ void __launch_bounds__(kThreads)
projection_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ quats, const float *__restrict__ scales,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, float W, float H,
                      float eps2d, float near_plane, float far_plane, float radius_clip, int model,
                      int packed_rule, int32_t *__restrict__ radii, float *__restrict__ means2d,
                      float *__restrict__ depths, float *__restrict__ conics, float *__restrict__ comps,
                      int32_t *__restrict__ pk_counts, int64_t *__restrict__ pk_camera_ids, int64_t *__restrict__ pk_gaussian_ids,
                      uint32_t ldN, const float *__restrict__ opacities, const __grid_constant__ ProjOwnerOut peer,
                      const __grid_constant__ ProjQuant Q) {
    __shared__ float s_vm[kMaxSmemCams][12];
    __shared__ float s_k[kMaxSmemCams][4];
    __shared__ int s_wcnt[kThreads / 32];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (g < N) { px = __ldg(means + (size_t)g * 3); py = __ldg(means + (size_t)g * 3 + 1); pz = __ldg(means + (size_t)g * 3 + 2); }
    if (MODE == 4 && g < N && Q.opac_raw != nullptr) {
        const float raw[1] = {__ldg(Q.opac_raw + g)};
        float act[1];
        quant_load<1>(Q.opac, (int64_t)g, raw, act);
        Q.opac_act[g] = act[0];
    }
    // the world-space covariance is built lazily, the first time a camera keeps this Gaussian after the
    // near/far test -- for most Gaussians of a large scene that never happens and quats/scales are never read
    Sym3 S = {0, 0, 0, 0, 0, 0};
    bool have_S = false;
    for (uint32_t c0 = 0; c0 < C; c0 += kMaxSmemCams) {
        const uint32_t nc = min((uint32_t)kMaxSmemCams, C - c0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nc * 12; i += blockDim.x) s_vm[i / 12][i % 12] = viewmats[(size_t)(c0 + i / 12) * 16 + i % 12];
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
            const float *K = Ks + (size_t)(c0 + i) * 9;
            s_k[i][0] = K[0]; s_k[i][1] = K[4]; s_k[i][2] = K[2]; s_k[i][3] = K[5];
        }
        __syncthreads();
        if ((MODE == 0 || MODE == 4) && g >= N) continue;  // (the packed passes meet block barriers inside the camera loop)
        if (MODE == 3 && g >= ldN) continue;  // owner mode: the padding rows [N, ldN) of the block publish radius 0
        for (uint32_t ci = 0; ci < nc; ++ci) {
            const float *R = s_vm[ci];
            const size_t idx = (size_t)(c0 + ci) * (MODE == 3 ? ldN : N) + g;
            float mx = R[0] * px + R[1] * py + R[2] * pz + R[3];
            float my = R[4] * px + R[5] * py + R[6] * pz + R[7];
            float mz = R[8] * px + R[9] * py + R[10] * pz + R[11];
            int32_t rad = 0;
            float o_mx = 0.f, o_my = 0.f, o_d = 0.f, o_a = 0.f, o_b = 0.f, o_c = 0.f, o_comp = 0.f;
            // (written so that a NaN depth -- the inert padding splats of ragged distributed shards -- is culled for every
            // camera model; equivalent to the reference's `z < near || z > far` test for every other value)
            if (g < N && mz >= near_plane && mz <= far_plane) {
                if (!have_S) {
                    have_S = true;
                    if (covars != nullptr) {
                        const float *c = covars + (size_t)g * 6;
                        S.xx = c[0]; S.xy = c[1]; S.xz = c[2]; S.yy = c[3]; S.yz = c[4]; S.zz = c[5];
                    } else {
                        const float4 q = ldg4(quats + (size_t)g * 4);
                        float qv[4] = {q.x, q.y, q.z, q.w};
                        float sv[3] = {__ldg(scales + (size_t)g * 3), __ldg(scales + (size_t)g * 3 + 1), __ldg(scales + (size_t)g * 3 + 2)};
                        if (MODE == 4) {
                            const float qr[4] = {qv[0], qv[1], qv[2], qv[3]}, sr[3] = {sv[0], sv[1], sv[2]};
                            quant_load<4>(Q.quats, (int64_t)g * 4, qr, qv);
                            quant_load<3>(Q.scales, (int64_t)g * 3, sr, sv);
                        }
                        const Mat3 Rq = quat_to_rot(qv[0], qv[1], qv[2], qv[3], nullptr);
                        S = covar_from_quat_scale(Rq, sv[0], sv[1], sv[2]);
                    }
                }
                const Mat3 Sc = covar_to_cam_full(R, S);
                CamProj P = project_point(model, mx, my, mz, s_k[ci][0], s_k[ci][1], s_k[ci][2], s_k[ci][3], W, H);
                float c00, c01, c10, c11;
                project_covar_full(P, Sc, c00, c01, c10, c11);
                // utils.cuh:32-40 and :8-21 with glm's [column][row] indexing: M[0][1] is (row 1, col 0)
                float det_orig = c00 * c11 - c10 * c01;
                c00 += eps2d; c11 += eps2d;
                float det = c00 * c11 - c10 * c01;
                float compensation = sqrtf(fmaxf(0.f, det_orig / det));
                if (det > 0.f) {
                    float inv_det = 1.f / det;
                    float b = 0.5f * (c00 + c11);
                    float radius;
                    if (packed_rule) {
                        float s = sqrtf(fmaxf(0.1f, b * b - det));
                        radius = ceilf(3.f * sqrtf(fmaxf(b + s, b - s)));
                    } else {
                        float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
                        radius = ceilf(3.f * sqrtf(v1));
                    }
                    bool ok = !(radius <= radius_clip);
                    ok = ok && !(P.mx + radius <= 0 || P.mx - radius >= W || P.my + radius <= 0 || P.my - radius >= H);
                    if (ok) {
                        rad = (int32_t)radius;
                        o_mx = P.mx; o_my = P.my; o_d = mz;
                        o_a = c11 * inv_det; o_b = -c10 * inv_det; o_c = c00 * inv_det;
                        o_comp = compensation;
                    }
                }
            }
            if (MODE == 0 || MODE == 4) {
                radii[idx] = rad;
                *reinterpret_cast<float2 *>(means2d + idx * 2) = make_float2(o_mx, o_my);
                depths[idx] = o_d;
                conics[idx * 3] = o_a; conics[idx * 3 + 1] = o_b; conics[idx * 3 + 2] = o_c;
                if (comps != nullptr) comps[idx] = o_comp;
            } else if (MODE == 3) {
                // owner mode: view c0 + ci belongs to another rank.  Its radius always travels (the viewer's tile
                // intersection reads every row), everything else only for a visible pair; a local copy of radius and conic
                // stays here for the backward.
                const uint32_t v = c0 + ci;
                radii[idx] = rad;
                peer.radii[v][g] = rad;
                if (rad > 0) {
                    conics[idx * 3] = o_a; conics[idx * 3 + 1] = o_b; conics[idx * 3 + 2] = o_c;
                    *reinterpret_cast<float2 *>(peer.means2d[v] + (size_t)g * 2) = make_float2(o_mx, o_my);
                    peer.depths[v][g] = o_d;
                    float *pc = peer.conics[v] + (size_t)g * 3;
                    pc[0] = o_a; pc[1] = o_b; pc[2] = o_c;
                    peer.opacities[v][g] = __ldg(opacities + g);
                }
            } else if (MODE == 1) {
                const int cnt = __syncthreads_count(rad > 0);
                if (threadIdx.x == 0) pk_counts[(size_t)(c0 + ci) * gridDim.x + blockIdx.x] = cnt;
            } else {
                const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                const unsigned bal = __ballot_sync(0xffffffffu, rad > 0);
                if (lane == 0) s_wcnt[warp] = __popc(bal);
                __syncthreads();
                int before = 0;
#pragma unroll
                for (uint32_t w = 0; w < kThreads / 32; ++w) before += w < warp ? s_wcnt[w] : 0;
                if (rad > 0) {
                    const size_t row = (size_t)pk_counts[(size_t)(c0 + ci) * gridDim.x + blockIdx.x] + before +
                                       __popc(bal & ((1u << lane) - 1u));
                    pk_camera_ids[row] = c0 + ci;
                    pk_gaussian_ids[row] = g;
                    radii[row] = rad;
                    *reinterpret_cast<float2 *>(means2d + row * 2) = make_float2(o_mx, o_my);
                    depths[row] = o_d;
                    conics[row * 3] = o_a; conics[row * 3 + 1] = o_b; conics[row * 3 + 2] = o_c;
                    if (comps != nullptr) comps[row] = o_comp;
                }
                __syncthreads();
            }
        }
    }
}

// PACKED: radii / conics / comps / v_* are [nnz, ...] rows in (camera, Gaussian) order and `pk_offsets` / `pk_gaussian_ids`
// (the forward's per-(camera, block) row offsets and the rows' Gaussian indices) tell every block which of its threads
// own a row for the current camera: same ownership as the dense form -- no atomics, no dense scatter of the gradients.
// QUANT: quats / scales (/ opacities) are the RAW parameters of a MODE 4 forward; the gradients are those of the raw parameters
template <bool VIEWGRAD, bool FISHEYE, bool PACKED, bool QUANT = false>
__global__ void __launch_bounds__(kThreads, FISHEYE ? 1 : 3)
projection_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ quats, const float *__restrict__ scales,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, float W, float H,
                      float eps2d, int model, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                      const float *__restrict__ comps, const float *__restrict__ v_means2d,
                      const float *__restrict__ v_depths, const float *__restrict__ v_conics,
                      const float *__restrict__ v_comps, float *__restrict__ v_means, float *__restrict__ v_covars,
                      float *__restrict__ v_quats, float *__restrict__ v_scales, float *__restrict__ v_viewmats,
                      const int32_t *__restrict__ pk_offsets, const int64_t *__restrict__ pk_gaussian_ids, uint32_t ldN,
                      const __grid_constant__ ProjQuant Q) {
    __shared__ float s_vm[kMaxSmemCams][12];
    __shared__ float s_k[kMaxSmemCams][4];
    __shared__ float s_red[kThreads / 32][12];
    __shared__ int s_row[PACKED ? kThreads : 1];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float px = 0.f, py = 0.f, pz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    Sym3 S = {0, 0, 0, 0, 0, 0};
    Mat3 Rq;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Rq.m[i][j] = 0.f;
    float inv_norm = 1.f;
    bool have_S = false;  // Gaussian parameters are read lazily: only Gaussians visible in some camera need them
    float vmean[3] = {0.f, 0.f, 0.f};
    float vS[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};  // full (non-symmetrised) v_Sigma
    for (uint32_t c0 = 0; c0 < C; c0 += kMaxSmemCams) {
        const uint32_t nc = min((uint32_t)kMaxSmemCams, C - c0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nc * 12; i += blockDim.x) s_vm[i / 12][i % 12] = viewmats[(size_t)(c0 + i / 12) * 16 + i % 12];
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
            const float *K = Ks + (size_t)(c0 + i) * 9;
            s_k[i][0] = K[0]; s_k[i][1] = K[4]; s_k[i][2] = K[2]; s_k[i][3] = K[5];
        }
        __syncthreads();
        for (uint32_t ci = 0; ci < nc; ++ci) {
            const float *R = s_vm[ci];
            size_t idx = (size_t)(c0 + ci) * ldN + g;  // ldN = N, or the block stride of the owner-side staging
            float vR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            float vt[3] = {0.f, 0.f, 0.f};
            bool active;
            if (PACKED) {
                // rows of this block for this camera: [lo, hi); their Gaussian indices say which thread owns which row
                const size_t ob = (size_t)(c0 + ci) * gridDim.x + blockIdx.x;
                const int lo = pk_offsets[ob], hi = pk_offsets[ob + 1];
                __syncthreads();
                s_row[threadIdx.x] = -1;
                __syncthreads();
                if ((int)threadIdx.x < hi - lo) s_row[(uint32_t)pk_gaussian_ids[lo + threadIdx.x] - blockIdx.x * kThreads] = lo + (int)threadIdx.x;
                __syncthreads();
                const int row = s_row[threadIdx.x];
                active = row >= 0;
                idx = active ? (size_t)row : 0;
            } else {
                active = (g < N) && radii[idx] > 0;
            }
            if (active) {
                if (!have_S) {
                    have_S = true;
                    px = __ldg(means + (size_t)g * 3); py = __ldg(means + (size_t)g * 3 + 1); pz = __ldg(means + (size_t)g * 3 + 2);
                    if (covars != nullptr) {
                        const float *c = covars + (size_t)g * 6;
                        S.xx = c[0]; S.xy = c[1]; S.xz = c[2]; S.yy = c[3]; S.yz = c[4]; S.zz = c[5];
                    } else {
                        q = ldg4(quats + (size_t)g * 4);
                        sx = __ldg(scales + (size_t)g * 3); sy = __ldg(scales + (size_t)g * 3 + 1); sz = __ldg(scales + (size_t)g * 3 + 2);
                        if (QUANT) {  // the values the forward saw (same uniforms: counter-based RNG)
                            const float qr[4] = {q.x, q.y, q.z, q.w}, sr[3] = {sx, sy, sz};
                            float qv[4], sv[3];
                            quant_load<4>(Q.quats, (int64_t)g * 4, qr, qv);
                            quant_load<3>(Q.scales, (int64_t)g * 3, sr, sv);
                            q = make_float4(qv[0], qv[1], qv[2], qv[3]);
                            sx = sv[0]; sy = sv[1]; sz = sv[2];
                        }
                        Rq = quat_to_rot(q.x, q.y, q.z, q.w, &inv_norm);
                        S = covar_from_quat_scale(Rq, sx, sy, sz);
                    }
                }
                const float ca = conics[idx * 3], cb = conics[idx * 3 + 1], cc = conics[idx * 3 + 2];
                const float va = v_conics[idx * 3], vb = 0.5f * v_conics[idx * 3 + 1], vc = v_conics[idx * 3 + 2];
                // v_cov2d = -Cn V Cn  (utils.cuh:24-29)
                float t00 = ca * va + cb * vb, t01 = ca * vb + cb * vc;
                float t10 = cb * va + cc * vb, t11 = cb * vb + cc * vc;
                float G00 = -(t00 * ca + t01 * cb), G01 = -(t00 * cb + t01 * cc);
                float G10 = -(t10 * ca + t11 * cb), G11 = -(t10 * cb + t11 * cc);
                if (v_comps != nullptr) {  // utils.cuh:42-73
                    const float comp = comps[idx], v_comp = v_comps[idx];
                    float det_conic = ca * cc - cb * cb;
                    float v_sqr = v_comp * 0.5f / (comp + 1e-6f);
                    float omc = 1.f - comp * comp;
                    G00 += v_sqr * (omc * ca - eps2d * det_conic);
                    G01 += v_sqr * (omc * cb);
                    G10 += v_sqr * (omc * cb);
                    G11 += v_sqr * (omc * cc - eps2d * det_conic);
                }
                float mx = R[0] * px + R[1] * py + R[2] * pz + R[3];
                float my = R[4] * px + R[5] * py + R[6] * pz + R[7];
                float mz = R[8] * px + R[9] * py + R[10] * pz + R[11];
                Sym3 Sc = covar_to_cam(R, S);
                const float fx = s_k[ci][0], fy = s_k[ci][1];
                CamProj P = project_point(model, mx, my, mz, fx, fy, s_k[ci][2], s_k[ci][3], W, H);
                const float vm0 = v_means2d[idx * 2], vm1 = v_means2d[idx * 2 + 1];
                // v_Sigma_c = J^T G J
                float GJ[2][3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    GJ[0][j] = G00 * P.J[0][j] + G01 * P.J[1][j];
                    GJ[1][j] = G10 * P.J[0][j] + G11 * P.J[1][j];
                }
                float vSc[3][3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) vSc[i][j] = P.J[0][i] * GJ[0][j] + P.J[1][i] * GJ[1][j];
                float vmc[3];
                if (model == GSB200_CAMERA_PINHOLE) {
                    const float rz = P.rz, rz2 = rz * rz, rz3 = rz2 * rz;
                    vmc[0] = fx * rz * vm0;
                    vmc[1] = fy * rz * vm1;
                    vmc[2] = -(fx * mx * vm0 + fy * my * vm1) * rz2;
                    // v_J = G J Sigma_c^T + G^T J Sigma_c  (only the entries the pinhole Jacobian depends on)
                    float ScM[3][3] = {{Sc.xx, Sc.xy, Sc.xz}, {Sc.xy, Sc.yy, Sc.yz}, {Sc.xz, Sc.yz, Sc.zz}};
                    float GtJ[2][3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        GtJ[0][j] = G00 * P.J[0][j] + G10 * P.J[1][j];
                        GtJ[1][j] = G01 * P.J[0][j] + G11 * P.J[1][j];
                    }
                    float vJ[2][3];
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            vJ[r][j] = (GJ[r][0] + GtJ[r][0]) * ScM[0][j] + (GJ[r][1] + GtJ[r][1]) * ScM[1][j] +
                                       (GJ[r][2] + GtJ[r][2]) * ScM[2][j];
                    if (!P.clip_x) vmc[0] += -fx * rz2 * vJ[0][2];
                    else vmc[2] += -fx * rz3 * vJ[0][2] * P.tx;
                    if (!P.clip_y) vmc[1] += -fy * rz2 * vJ[1][2];
                    else vmc[2] += -fy * rz3 * vJ[1][2] * P.ty;
                    vmc[2] += -fx * rz2 * vJ[0][0] - fy * rz2 * vJ[1][1] + 2.f * fx * P.tx * rz3 * vJ[0][2] +
                              2.f * fy * P.ty * rz3 * vJ[1][2];
                } else if (FISHEYE && model == GSB200_CAMERA_FISHEYE) {
                    // exact derivative of the fisheye map: duals give d(mean2d)/d(xyz) and dJ/d(xyz)
                    Dual3 m[2], Jd[2][3];
                    fisheye_dual(mx, my, mz, fx, fy, s_k[ci][2], s_k[ci][3], m, Jd);
                    float ScM[3][3] = {{Sc.xx, Sc.xy, Sc.xz}, {Sc.xy, Sc.yy, Sc.yz}, {Sc.xz, Sc.yz, Sc.zz}};
                    float vJ[2][3];
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const float g0 = GJ[r][0] + (r == 0 ? G00 * P.J[0][0] + G10 * P.J[1][0] : G01 * P.J[0][0] + G11 * P.J[1][0]);
                            const float g1 = GJ[r][1] + (r == 0 ? G00 * P.J[0][1] + G10 * P.J[1][1] : G01 * P.J[0][1] + G11 * P.J[1][1]);
                            const float g2 = GJ[r][2] + (r == 0 ? G00 * P.J[0][2] + G10 * P.J[1][2] : G01 * P.J[0][2] + G11 * P.J[1][2]);
                            vJ[r][j] = g0 * ScM[0][j] + g1 * ScM[1][j] + g2 * ScM[2][j];
                        }
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
                if (v_depths != nullptr) vmc[2] += v_depths[idx];
                // world -> camera VJPs (transform.cuh:20-71)
#pragma unroll
                for (int j = 0; j < 3; ++j) vmean[j] += R[j] * vmc[0] + R[4 + j] * vmc[1] + R[8 + j] * vmc[2];
                // v_Sigma += R^T vSc R
                float Tm[3][3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) Tm[i][j] = vSc[i][0] * R[j] + vSc[i][1] * R[4 + j] + vSc[i][2] * R[8 + j];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) vS[i][j] += R[i] * Tm[0][j] + R[4 + i] * Tm[1][j] + R[8 + i] * Tm[2][j];
                if (VIEWGRAD) {
                    const float p[3] = {px, py, pz};
                    const float SM[3][3] = {{S.xx, S.xy, S.xz}, {S.xy, S.yy, S.yz}, {S.xz, S.yz, S.zz}};
                    // v_R = outer(vmc, p) + vSc R Sigma^T + vSc^T R Sigma = outer + (vSc + vSc^T) R Sigma
                    float RS[3][3];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) RS[i][j] = R[i * 4] * SM[0][j] + R[i * 4 + 1] * SM[1][j] + R[i * 4 + 2] * SM[2][j];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        vt[i] = vmc[i];
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            vR[i][j] = vmc[i] * p[j] + (vSc[i][0] + vSc[0][i]) * RS[0][j] + (vSc[i][1] + vSc[1][i]) * RS[1][j] +
                                       (vSc[i][2] + vSc[2][i]) * RS[2][j];
                    }
                }
            }
            if (VIEWGRAD) {  // block reduction over Gaussians, one atomic set per block and camera
                float vals[12];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    vals[i * 4 + 0] = warp_sum(vR[i][0]); vals[i * 4 + 1] = warp_sum(vR[i][1]);
                    vals[i * 4 + 2] = warp_sum(vR[i][2]); vals[i * 4 + 3] = warp_sum(vt[i]);
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
                    if (s != 0.f) atomicAdd(v_viewmats + (size_t)(c0 + ci) * 16 + threadIdx.x, s);
                }
            }
        }
    }
    if (g >= N) return;
    if (QUANT && Q.v_opac_raw != nullptr)  // sigmoid'(.) and the clamp gate of the opacity (the blend's gradient arrives summed over cameras)
        Q.v_opac_raw[g] = Q.opac.on ? quant_grad(Q.opac, __ldg(Q.v_opac_act + g), __ldg(Q.opac_raw + g), __ldg(Q.opac_act + g))
                                    : __ldg(Q.v_opac_act + g);
    if (!have_S) {  // invisible everywhere: all gradients are zero
        v_means[(size_t)g * 3] = 0.f; v_means[(size_t)g * 3 + 1] = 0.f; v_means[(size_t)g * 3 + 2] = 0.f;
        if (covars != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) v_covars[(size_t)g * 6 + k] = 0.f;
        } else {
            stg4(v_quats + (size_t)g * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            v_scales[(size_t)g * 3] = 0.f; v_scales[(size_t)g * 3 + 1] = 0.f; v_scales[(size_t)g * 3 + 2] = 0.f;
        }
        return;
    }
    v_means[(size_t)g * 3] = vmean[0]; v_means[(size_t)g * 3 + 1] = vmean[1]; v_means[(size_t)g * 3 + 2] = vmean[2];
    if (covars != nullptr) {
        float *o = v_covars + (size_t)g * 6;
        o[0] = vS[0][0]; o[1] = vS[0][1] + vS[1][0]; o[2] = vS[0][2] + vS[2][0];
        o[3] = vS[1][1]; o[4] = vS[1][2] + vS[2][1]; o[5] = vS[2][2];
    } else {
        // Sigma = M M^T, M = Rq diag(s)  (quat_scale_to_covar_preci.cuh:44-81, quat.cuh:35-57)
        float M[3][3], vM[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { M[i][0] = Rq.m[i][0] * sx; M[i][1] = Rq.m[i][1] * sy; M[i][2] = Rq.m[i][2] * sz; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                vM[i][j] = (vS[i][0] + vS[0][i]) * M[0][j] + (vS[i][1] + vS[1][i]) * M[1][j] + (vS[i][2] + vS[2][i]) * M[2][j];
        float vRq[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { vRq[i][0] = vM[i][0] * sx; vRq[i][1] = vM[i][1] * sy; vRq[i][2] = vM[i][2] * sz; }
        float vsx = Rq.m[0][0] * vM[0][0] + Rq.m[1][0] * vM[1][0] + Rq.m[2][0] * vM[2][0];
        float vsy = Rq.m[0][1] * vM[0][1] + Rq.m[1][1] * vM[1][1] + Rq.m[2][1] * vM[2][1];
        float vsz = Rq.m[0][2] * vM[0][2] + Rq.m[1][2] * vM[1][2] + Rq.m[2][2] * vM[2][2];
        const float w = q.x * inv_norm, x = q.y * inv_norm, y = q.z * inv_norm, z = q.w * inv_norm;
        float vw = 2.f * (x * (vRq[2][1] - vRq[1][2]) + y * (vRq[0][2] - vRq[2][0]) + z * (vRq[1][0] - vRq[0][1]));
        float vx = 2.f * (-2.f * x * (vRq[1][1] + vRq[2][2]) + y * (vRq[1][0] + vRq[0][1]) + z * (vRq[2][0] + vRq[0][2]) +
                          w * (vRq[2][1] - vRq[1][2]));
        float vy = 2.f * (x * (vRq[1][0] + vRq[0][1]) - 2.f * y * (vRq[0][0] + vRq[2][2]) + z * (vRq[2][1] + vRq[1][2]) +
                          w * (vRq[0][2] - vRq[2][0]));
        float vz = 2.f * (x * (vRq[2][0] + vRq[0][2]) + y * (vRq[2][1] + vRq[1][2]) - 2.f * z * (vRq[0][0] + vRq[1][1]) +
                          w * (vRq[1][0] - vRq[0][1]));
        float dotp = vw * w + vx * x + vy * y + vz * z;
        float4 vq = make_float4((vw - dotp * w) * inv_norm, (vx - dotp * x) * inv_norm, (vy - dotp * y) * inv_norm,
                                (vz - dotp * z) * inv_norm);
        if (QUANT) {  // back to the raw parameters: clamp gates, exp'(.) = the activated scale itself
            if (Q.quats.on) {
                const float4 qr = ldg4(quats + (size_t)g * 4);
                vq.x = quant_grad(Q.quats, vq.x, qr.x, q.x); vq.y = quant_grad(Q.quats, vq.y, qr.y, q.y);
                vq.z = quant_grad(Q.quats, vq.z, qr.z, q.z); vq.w = quant_grad(Q.quats, vq.w, qr.w, q.w);
            }
            if (Q.scales.on) {
                vsx = quant_grad(Q.scales, vsx, __ldg(scales + (size_t)g * 3), sx);
                vsy = quant_grad(Q.scales, vsy, __ldg(scales + (size_t)g * 3 + 1), sy);
                vsz = quant_grad(Q.scales, vsz, __ldg(scales + (size_t)g * 3 + 2), sz);
            }
        }
        stg4(v_quats + (size_t)g * 4, vq);
        v_scales[(size_t)g * 3] = vsx; v_scales[(size_t)g * 3 + 1] = vsy; v_scales[(size_t)g * 3 + 2] = vsz;
    }
}

// ------------------------------------------------------------------ packed (COO) helpers
constexpr int kCompactItems = 8;  // 2048 elements per block

__global__ void __launch_bounds__(kThreads) compact_count_kernel(uint32_t n, const int32_t *__restrict__ radii,
                                                                 int32_t *__restrict__ block_counts) {
    __shared__ int s_w[kThreads / 32];
    const uint32_t base = blockIdx.x * kThreads * kCompactItems;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) {
        uint32_t i = base + k * kThreads + threadIdx.x;
        cnt += (i < n && radii[i] > 0) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_w[w];
        block_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kThreads)
compact_fill_kernel(uint32_t C, uint32_t N, const int32_t *__restrict__ radii, const float *__restrict__ means2d,
                    const float *__restrict__ depths, const float *__restrict__ conics, const float *__restrict__ comps,
                    const int32_t *__restrict__ block_offsets, int32_t *__restrict__ indptr,
                    int64_t *__restrict__ camera_ids, int64_t *__restrict__ gaussian_ids, int32_t *__restrict__ p_radii,
                    float *__restrict__ p_means2d, float *__restrict__ p_depths, float *__restrict__ p_conics,
                    float *__restrict__ p_comps) {
    __shared__ int s_w[kThreads / 32];
    const uint32_t n = C * N;
    const uint32_t base = blockIdx.x * kThreads * kCompactItems;
    int running = block_offsets[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < kCompactItems; ++k) {
        const uint32_t i = base + k * kThreads + threadIdx.x;
        const bool vis = i < n && radii[i] > 0;
        const unsigned m = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            const int c = s_w[w];
            before += w < warp ? c : 0;
            total += c;
        }
        const int row = running + before + __popc(m & ((1u << lane) - 1u));
        if (vis) {
            camera_ids[row] = i / N;
            gaussian_ids[row] = i % N;
            p_radii[row] = radii[i];
            p_means2d[(size_t)row * 2] = means2d[(size_t)i * 2];
            p_means2d[(size_t)row * 2 + 1] = means2d[(size_t)i * 2 + 1];
            p_depths[row] = depths[i];
            p_conics[(size_t)row * 3] = conics[(size_t)i * 3];
            p_conics[(size_t)row * 3 + 1] = conics[(size_t)i * 3 + 1];
            p_conics[(size_t)row * 3 + 2] = conics[(size_t)i * 3 + 2];
            if (p_comps != nullptr) p_comps[row] = comps[i];
        }
        // indptr[c] = number of visible rows before element c*N: element i == c*N exactly starts camera c
        if (i < n && i % N == 0) indptr[i / N] = row;
        running += total;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) indptr[C] = block_offsets[gridDim.x];
}

__global__ void __launch_bounds__(kThreads)
scatter_rows_kernel(uint32_t N, int64_t nnz, const int64_t *__restrict__ camera_ids, const int64_t *__restrict__ gaussian_ids,
                    const float *__restrict__ pm, const float *__restrict__ pd, const float *__restrict__ pc,
                    const float *__restrict__ pk, float *__restrict__ vm, float *__restrict__ vd, float *__restrict__ vc,
                    float *__restrict__ vk) {
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= nnz) return;
    const size_t i = (size_t)camera_ids[r] * N + (size_t)gaussian_ids[r];
    vm[i * 2] = pm[r * 2]; vm[i * 2 + 1] = pm[r * 2 + 1];
    vd[i] = pd[r];
    vc[i * 3] = pc[r * 3]; vc[i * 3 + 1] = pc[r * 3 + 1]; vc[i * 3 + 2] = pc[r * 3 + 2];
    if (vk != nullptr) vk[i] = pk[r];
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_projection_fwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *quats,
                                     const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                     float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                                     int packed_radius_rule, int32_t *radii, float *means2d, float *depths, float *conics,
                                     float *compensations, gsb200_stream_t stream) {
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return fail(2, "gsb200_projection_fwd: camera_model must be pinhole, ortho or fisheye");
    if (C == 0 || N == 0) return 0;  // (empty tensors have NULL data pointers)
    if (covars == nullptr && (quats == nullptr || scales == nullptr))
        return fail(2, "gsb200_projection_fwd: covars or (quats, scales) should be provided");
    projection_fwd_kernel<0><<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, near_plane, far_plane, radius_clip,
        camera_model, packed_radius_rule, radii, means2d, depths, conics, compensations, nullptr, nullptr, nullptr, N, nullptr,
        ProjOwnerOut{}, ProjQuant{});
    return check_launch("projection_fwd");
}

// ---- quantise on load: the dense projection on the trainer's RAW parameters (gsplat/compression_simulation fused into the
// attribute load).  q_* == NULL: that attribute is read as it is.  opacities_raw / opacities_act may both be NULL.
static int make_proj_quant(const char *who, const gsb200_quant_load *q_quats, const gsb200_quant_load *q_scales,
                           const gsb200_quant_load *q_opac, ProjQuant &Q) {
    Q = ProjQuant{};
    if (int e = make_quant_load(who, q_quats, Q.quats)) return e;
    if (int e = make_quant_load(who, q_scales, Q.scales)) return e;
    if (int e = make_quant_load(who, q_opac, Q.opac)) return e;
    if (Q.quats.on && Q.quats.act != 0) return failf(2, "%s: quats take no activation", who);
    return 0;
}

extern "C" int gsb200_projection_quant_fwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                           const float *opacities_raw, const gsb200_quant_load *q_quats,
                                           const gsb200_quant_load *q_scales, const gsb200_quant_load *q_opacities,
                                           const float *viewmats, const float *Ks, uint32_t W, uint32_t H, float eps2d,
                                           float near_plane, float far_plane, float radius_clip, int camera_model, int32_t *radii,
                                           float *means2d, float *depths, float *conics, float *compensations,
                                           float *opacities_act, gsb200_stream_t stream) {
    const char *who = "gsb200_projection_quant_fwd";
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return failf(2, "%s: camera_model must be pinhole, ortho or fisheye", who);
    ProjQuant Q;
    if (int e = make_proj_quant(who, q_quats, q_scales, q_opacities, Q)) return e;
    if (C == 0 || N == 0) return 0;
    if (quats == nullptr || scales == nullptr) return failf(2, "%s: quats and scales are required", who);
    if ((opacities_raw == nullptr) != (opacities_act == nullptr)) return failf(2, "%s: opacities_raw and opacities_act go together", who);
    Q.opac_raw = opacities_raw;
    Q.opac_act = opacities_act;
    projection_fwd_kernel<4><<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, nullptr, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, near_plane, far_plane, radius_clip,
        camera_model, 0, radii, means2d, depths, conics, compensations, nullptr, nullptr, nullptr, N, nullptr, ProjOwnerOut{}, Q);
    return check_launch("projection_quant_fwd");
}

extern "C" int gsb200_projection_quant_bwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                           const float *opacities_raw, const gsb200_quant_load *q_quats,
                                           const gsb200_quant_load *q_scales, const gsb200_quant_load *q_opacities,
                                           const float *viewmats, const float *Ks, uint32_t W, uint32_t H, float eps2d,
                                           int camera_model, const int32_t *radii, const float *conics, const float *compensations,
                                           const float *v_means2d, const float *v_depths, const float *v_conics,
                                           const float *v_compensations, const float *opacities_act, const float *v_opacities_act,
                                           float *v_means, float *v_quats, float *v_scales, float *v_opacities_raw,
                                           gsb200_stream_t stream) {
    const char *who = "gsb200_projection_quant_bwd";
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return failf(2, "%s: camera_model must be pinhole, ortho or fisheye", who);
    ProjQuant Q;
    if (int e = make_proj_quant(who, q_quats, q_scales, q_opacities, Q)) return e;
    if (C == 0 || N == 0) return 0;
    if (!means || !quats || !scales || !radii || !conics || !v_means2d || !v_conics || !v_means || !v_quats || !v_scales)
        return failf(2, "%s: null pointer", who);
    if (v_opacities_raw != nullptr && (v_opacities_act == nullptr || (Q.opac.on && (opacities_raw == nullptr || opacities_act == nullptr))))
        return failf(2, "%s: the opacity gradient needs opacities_raw, opacities_act and v_opacities_act", who);
    Q.opac_raw = opacities_raw;
    Q.opac_act = const_cast<float *>(opacities_act);
    Q.v_opac_act = v_opacities_act;
    Q.v_opac_raw = v_opacities_raw;
    const dim3 grid(cdiv(N, kThreads));
    cudaStream_t s = (cudaStream_t)stream;
#define GSB_LAUNCH_PROJ_QBWD(FE)                                                                                       \
    projection_bwd_kernel<false, FE, false, true><<<grid, kThreads, 0, s>>>(                                            \
        C, N, means, nullptr, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, camera_model, radii, conics,      \
        compensations, v_means2d, v_depths, v_conics, v_compensations, v_means, nullptr, v_quats, v_scales, nullptr,    \
        nullptr, nullptr, N, Q)
    if (camera_model == GSB200_CAMERA_FISHEYE) GSB_LAUNCH_PROJ_QBWD(true); else GSB_LAUNCH_PROJ_QBWD(false);
#undef GSB_LAUNCH_PROJ_QBWD
    return check_launch("projection_quant_bwd");
}

// ---- packed (COO) projection: two passes, nothing of size C * N ------------------------------------------------------------
extern "C" int gsb200_projection_packed_count(uint32_t C, uint32_t N, const float *means, const float *covars, const float *quats,
                                              const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                              float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                                              int32_t *block_offsets, int64_t *nnz_host, gsb200_stream_t stream) {
    if (nnz_host != nullptr) *nnz_host = 0;
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return fail(2, "gsb200_projection_packed_count: camera_model must be pinhole, ortho or fisheye");
    if (C == 0 || N == 0) return 0;
    if (covars == nullptr && (quats == nullptr || scales == nullptr))
        return fail(2, "gsb200_projection_packed_count: covars or (quats, scales) should be provided");
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nb = cdiv(N, kThreads), n = C * nb;
    projection_fwd_kernel<1><<<nb, kThreads, 0, s>>>(C, N, means, covars, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d,
                                                     near_plane, far_plane, radius_clip, camera_model, 1, nullptr, nullptr, nullptr,
                                                     nullptr, nullptr, block_offsets, nullptr, nullptr, N, nullptr, ProjOwnerOut{}, ProjQuant{});
    if (int e = check_launch("projection_packed_count")) return e;
    // exclusive scan over (camera, block): block_offsets[0..n) in place, total -> block_offsets[n]; scratch behind it
    int32_t *scratch = block_offsets + n + 1;
    const uint32_t nsb = cdiv(n, kScanTile);
    scan_block_sums_kernel<<<nsb, kThreads, 0, s>>>(n, block_offsets, scratch);
    if (int e = check_launch("scan_block_sums")) return e;
    scan_single_block_kernel<<<1, 1024, 0, s>>>(nsb, scratch);
    if (int e = check_launch("scan_single_block")) return e;
    scan_apply_kernel<<<nsb, kThreads, 0, s>>>(n, block_offsets, scratch, block_offsets, nullptr);
    if (int e = check_launch("scan_apply")) return e;
    GSB_CUDA_OK(cudaMemcpyAsync(block_offsets + n, scratch + nsb, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (nnz_host != nullptr) {
        int32_t total = 0;
        GSB_CUDA_OK(cudaMemcpyAsync(&total, scratch + nsb, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        GSB_CUDA_OK(cudaStreamSynchronize(s));
        *nnz_host = total;
    }
    return 0;
}

__global__ void __launch_bounds__(kThreads)
packed_indptr_kernel(uint32_t C, uint32_t nb, const int32_t *__restrict__ block_offsets, int32_t *__restrict__ indptr) {
    const uint32_t c = blockIdx.x * kThreads + threadIdx.x;
    if (c <= C) indptr[c] = block_offsets[(size_t)c * nb];  // c == C: the total behind the prefixes
}

extern "C" int gsb200_projection_packed_fill(uint32_t C, uint32_t N, const float *means, const float *covars, const float *quats,
                                             const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                             float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                                             const int32_t *block_offsets, int32_t *indptr, int64_t *camera_ids,
                                             int64_t *gaussian_ids, int32_t *radii, float *means2d, float *depths, float *conics,
                                             float *compensations, gsb200_stream_t stream) {
    if (C == 0 || N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nb = cdiv(N, kThreads);
    projection_fwd_kernel<2><<<nb, kThreads, 0, s>>>(C, N, means, covars, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d,
                                                     near_plane, far_plane, radius_clip, camera_model, 1, radii, means2d, depths,
                                                     conics, compensations, const_cast<int32_t *>(block_offsets), camera_ids,
                                                     gaussian_ids, N, nullptr, ProjOwnerOut{}, ProjQuant{});
    if (int e = check_launch("projection_packed_fill")) return e;
    packed_indptr_kernel<<<cdiv(C + 1, kThreads), kThreads, 0, s>>>(C, nb, block_offsets, indptr);
    return check_launch("packed_indptr");
}

extern "C" int gsb200_projection_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *quats,
                                     const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                     float eps2d, int camera_model, const int32_t *radii, const float *conics,
                                     const float *compensations, const float *v_means2d, const float *v_depths,
                                     const float *v_conics, const float *v_compensations, float *v_means, float *v_covars,
                                     float *v_quats, float *v_scales, float *v_viewmats, gsb200_stream_t stream) {
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return fail(2, "gsb200_projection_bwd: camera_model must be pinhole, ortho or fisheye");
    if (C == 0 || N == 0) return 0;
    if (v_viewmats != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * C, (cudaStream_t)stream));
    const dim3 grid(cdiv(N, kThreads));
    cudaStream_t s = (cudaStream_t)stream;
#define GSB_LAUNCH_PROJ_BWD(VG, FE)                                                                                  \
    projection_bwd_kernel<VG, FE, false><<<grid, kThreads, 0, s>>>(                                                   \
        C, N, means, covars, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, camera_model, radii, conics,     \
        compensations, v_means2d, v_depths, v_conics, v_compensations, v_means, v_covars, v_quats, v_scales, v_viewmats, \
        nullptr, nullptr, N, ProjQuant{})
    const bool fe = camera_model == GSB200_CAMERA_FISHEYE;
    if (v_viewmats != nullptr) { if (fe) GSB_LAUNCH_PROJ_BWD(true, true); else GSB_LAUNCH_PROJ_BWD(true, false); }
    else { if (fe) GSB_LAUNCH_PROJ_BWD(false, true); else GSB_LAUNCH_PROJ_BWD(false, false); }
#undef GSB_LAUNCH_PROJ_BWD
    return check_launch("projection_bwd");
}

/* packed backward: conics / compensations / v_* are the [nnz, ...] rows of gsb200_projection_packed_fill; block_offsets and
 * gaussian_ids are the forward's.  Gradient outputs are dense [N, ...] (sum over cameras), written once, no atomics. */
extern "C" int gsb200_projection_packed_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *quats,
                                            const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                            float eps2d, int camera_model, const int32_t *block_offsets,
                                            const int64_t *gaussian_ids, const float *conics, const float *compensations,
                                            const float *v_means2d, const float *v_depths, const float *v_conics,
                                            const float *v_compensations, float *v_means, float *v_covars, float *v_quats,
                                            float *v_scales, float *v_viewmats, gsb200_stream_t stream) {
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return fail(2, "gsb200_projection_packed_bwd: camera_model must be pinhole, ortho or fisheye");
    if (C == 0 || N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (v_viewmats != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * C, s));
    const dim3 grid(cdiv(N, kThreads));
#define GSB_LAUNCH_PROJ_PBWD(VG, FE)                                                                                 \
    projection_bwd_kernel<VG, FE, true><<<grid, kThreads, 0, s>>>(                                                    \
        C, N, means, covars, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, camera_model, nullptr, conics,   \
        compensations, v_means2d, v_depths, v_conics, v_compensations, v_means, v_covars, v_quats, v_scales, v_viewmats, \
        block_offsets, gaussian_ids, N, ProjQuant{})
    const bool fe = camera_model == GSB200_CAMERA_FISHEYE;
    if (v_viewmats != nullptr) { if (fe) GSB_LAUNCH_PROJ_PBWD(true, true); else GSB_LAUNCH_PROJ_PBWD(true, false); }
    else { if (fe) GSB_LAUNCH_PROJ_PBWD(false, true); else GSB_LAUNCH_PROJ_PBWD(false, false); }
#undef GSB_LAUNCH_PROJ_PBWD
    return check_launch("projection_packed_bwd");
}

extern "C" int gsb200_compact_count(uint32_t C, uint32_t N, const int32_t *radii, int32_t *block_offsets,
                                    int64_t *nnz_host, gsb200_stream_t stream) {
    const uint32_t n = C * N;
    const uint32_t nb = cdiv(n, kThreads * kCompactItems);
    *nnz_host = 0;
    if (n == 0) return 0;
    compact_count_kernel<<<nb, kThreads, 0, (cudaStream_t)stream>>>(n, radii, block_offsets);
    if (int e = check_launch("compact_count")) return e;
    scan_single_block_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(nb, block_offsets);
    if (int e = check_launch("scan_blocks")) return e;
    int32_t total = 0;
    GSB_CUDA_OK(cudaMemcpyAsync(&total, block_offsets + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GSB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    *nnz_host = total;
    return 0;
}

extern "C" int gsb200_compact_fill(uint32_t C, uint32_t N, const int32_t *radii, const float *means2d, const float *depths,
                                   const float *conics, const float *compensations, const int32_t *block_offsets,
                                   int32_t *indptr, int64_t *camera_ids, int64_t *gaussian_ids, int32_t *p_radii,
                                   float *p_means2d, float *p_depths, float *p_conics, float *p_compensations,
                                   gsb200_stream_t stream) {
    const uint32_t n = C * N;
    if (n == 0) return 0;
    const uint32_t nb = cdiv(n, kThreads * kCompactItems);
    compact_fill_kernel<<<nb, kThreads, 0, (cudaStream_t)stream>>>(C, N, radii, means2d, depths, conics, compensations,
                                                                    block_offsets, indptr, camera_ids, gaussian_ids, p_radii,
                                                                    p_means2d, p_depths, p_conics, p_compensations);
    return check_launch("compact_fill");
}

extern "C" int gsb200_scatter_rows(uint32_t C, uint32_t N, int64_t nnz, const int64_t *camera_ids,
                                   const int64_t *gaussian_ids, const float *p_v_means2d, const float *p_v_depths,
                                   const float *p_v_conics, const float *p_v_compensations, float *v_means2d, float *v_depths,
                                   float *v_conics, float *v_compensations, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)C * N;
    GSB_CUDA_OK(cudaMemsetAsync(v_means2d, 0, n * 2 * sizeof(float), s));
    GSB_CUDA_OK(cudaMemsetAsync(v_depths, 0, n * sizeof(float), s));
    GSB_CUDA_OK(cudaMemsetAsync(v_conics, 0, n * 3 * sizeof(float), s));
    if (v_compensations != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_compensations, 0, n * sizeof(float), s));
    if (nnz == 0) return 0;
    scatter_rows_kernel<<<(unsigned)((nnz + kThreads - 1) / kThreads), kThreads, 0, s>>>(
        N, nnz, camera_ids, gaussian_ids, p_v_means2d, p_v_depths, p_v_conics, p_v_compensations, v_means2d, v_depths,
        v_conics, v_compensations);
    return check_launch("scatter_rows");
}


// ---- owner-side projection (view-parallel rendering over NVLink peer memory) ------------------------------------------------
// Every rank projects ITS rows for all V views of the job (the cameras were exchanged first) and stores the results into the
// viewers' tensors; the backward runs on the owner as well, on gradients the viewers pushed into local staging [V, n_block, .]
// -- the Gaussians' attributes and their gradients never cross the link.
extern "C" int gsb200_projection_owner_fwd(uint32_t V, uint32_t n_local, uint32_t n_block, const float *means, const float *quats,
                                           const float *scales, const float *opacities, const float *viewmats, const float *Ks,
                                           uint32_t W, uint32_t H, float eps2d, float near_plane, float far_plane,
                                           float radius_clip, int camera_model, int32_t *radii_local, float *conics_local,
                                           int32_t *const *view_radii, float *const *view_means2d, float *const *view_depths,
                                           float *const *view_conics, float *const *view_opacities, gsb200_stream_t stream) {
    const char *who = "gsb200_projection_owner_fwd";
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return failf(2, "%s: camera_model must be pinhole, ortho or fisheye", who);
    if (V == 0 || V > GSB200_MAX_VIEWS) return failf(2, "%s: 1..%d views", who, GSB200_MAX_VIEWS);
    if (n_local > n_block) return failf(2, "%s: n_local > n_block", who);
    if (n_block == 0) return 0;
    if (n_local != 0 && (means == nullptr || quats == nullptr || scales == nullptr || opacities == nullptr))
        return failf(2, "%s: null attribute pointer", who);
    if (radii_local == nullptr || conics_local == nullptr || viewmats == nullptr || Ks == nullptr)
        return failf(2, "%s: null pointer", who);
    ProjOwnerOut out{};
    for (uint32_t v = 0; v < V; ++v) {
        out.radii[v] = view_radii[v]; out.means2d[v] = view_means2d[v]; out.depths[v] = view_depths[v];
        out.conics[v] = view_conics[v]; out.opacities[v] = view_opacities[v];
        if (!out.radii[v] || !out.means2d[v] || !out.depths[v] || !out.conics[v] || !out.opacities[v])
            return failf(2, "%s: null view pointer", who);
        if ((reinterpret_cast<uintptr_t>(out.means2d[v]) & 7) != 0) return failf(2, "%s: means2d slices must be 8-byte aligned", who);
    }
    projection_fwd_kernel<3><<<cdiv(n_block, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        V, n_local, means, nullptr, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, near_plane, far_plane, radius_clip,
        camera_model, 0, radii_local, nullptr, nullptr, conics_local, nullptr, nullptr, nullptr, nullptr, n_block, opacities, out, ProjQuant{});
    return check_launch("projection_owner_fwd");
}

extern "C" int gsb200_projection_owner_bwd(uint32_t V, uint32_t n_local, uint32_t n_block, const float *means, const float *quats,
                                           const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                           float eps2d, int camera_model, const int32_t *radii_local, const float *conics_local,
                                           const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means,
                                           float *v_quats, float *v_scales, gsb200_stream_t stream) {
    const char *who = "gsb200_projection_owner_bwd";
    if (camera_model < GSB200_CAMERA_PINHOLE || camera_model > GSB200_CAMERA_FISHEYE)
        return failf(2, "%s: camera_model must be pinhole, ortho or fisheye", who);
    if (V == 0 || n_local > n_block) return failf(2, "%s: bad sizes", who);
    if (n_local == 0) return 0;
    if (!means || !quats || !scales || !viewmats || !Ks || !radii_local || !conics_local || !v_means2d || !v_conics || !v_means ||
        !v_quats || !v_scales)
        return failf(2, "%s: null pointer", who);
    const dim3 grid(cdiv(n_local, kThreads));
    cudaStream_t s = (cudaStream_t)stream;
#define GSB_LAUNCH_OWNER_BWD(FE)                                                                                              \
    projection_bwd_kernel<false, FE, false><<<grid, kThreads, 0, s>>>(                                                          \
        V, n_local, means, nullptr, quats, scales, viewmats, Ks, (float)W, (float)H, eps2d, camera_model, radii_local,         \
        conics_local, nullptr, v_means2d, v_depths, v_conics, nullptr, v_means, nullptr, v_quats, v_scales, nullptr, nullptr,   \
        nullptr, n_block, ProjQuant{})
    if (camera_model == GSB200_CAMERA_FISHEYE) GSB_LAUNCH_OWNER_BWD(true); else GSB_LAUNCH_OWNER_BWD(false);
#undef GSB_LAUNCH_OWNER_BWD
    return check_launch("projection_owner_bwd");
}
