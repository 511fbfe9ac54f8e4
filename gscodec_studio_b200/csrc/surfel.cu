// 2D Gaussian (surfel) family, forward and backward (sm_100a): ray-splat transform + bounding box per (camera, Gaussian),
// per-tile front-to-back blend with ray-splat intersection, normals, depth distortion and median depth, and the
// per-pixel contributor lists of the torch rasterizer.
//
// Behaviour follows the reference ops
//   _C.fully_fused_projection_{fwd,bwd}_2dgs     gsplat/cuda/csrc/fully_fused_projection_2dgs_{fwd,bwd}.cu, include/2dgs.cuh
//   _C.rasterize_to_pixels_{fwd,bwd}_2dgs        gsplat/cuda/csrc/rasterize_to_pixels_2dgs_{fwd,bwd}.cu
//   _C.rasterize_to_indices_in_range_2dgs        gsplat/cuda/csrc/rasterize_to_indices_in_range_2dgs.cu
// The per-element expressions restate the reference's (so that radii, tile lists and every alpha >= 1/255 decision agree);
// the organisation is this library's:
//   * projection: one thread owns one Gaussian and loops over the cameras (cameras in shared memory, Gaussian read once,
//     quaternion / scale read only if some camera keeps it); the backward sums over cameras in registers and writes each
//     gradient once -- no atomics, deterministic; culled entries are zero-filled instead of left uninitialised;
//   * blend: the tile's Gaussians are staged as four float4 per Gaussian; the backward reduces its CDIM + 15 (+ 2)
//     per-Gaussian partial sums with the halving butterfly of blend.cuh (8 values per group: 9 shuffles instead of 40)
//     and one vector of atomics per group;
//   * the densification signal (`v_densify`, read-after-atomic inside the reference's kernel, i.e. racy there) is
//     computed AFTER the blend backward from the final sums: v_densify[g] = (v_ray_transforms[g][0][2],
//     v_ray_transforms[g][1][2]) * ray_transforms[g][2][2] -- the value the reference converges to when its last writer
//     happens to see the complete sums.
#include "blend.cuh"
#include "geom.cuh"

namespace gsb200 {

constexpr int kSurfelMaxCams = 64;
constexpr float kFilterInvSquare = 2.0f;  // 2dgs.cuh:7

// ------------------------------------------------------------------------------------------------ projection forward
__global__ void __launch_bounds__(kThreads)
surfel_projection_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ quats,
                             const float *__restrict__ scales, const float *__restrict__ viewmats, const float *__restrict__ Ks,
                             float W, float H, float near_plane, float far_plane, float radius_clip, int32_t *__restrict__ radii,
                             float *__restrict__ means2d, float *__restrict__ depths, float *__restrict__ ray_transforms,
                             float *__restrict__ normals) {
    __shared__ float s_vm[kSurfelMaxCams][12];
    __shared__ float s_k[kSurfelMaxCams][4];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (g < N) { px = __ldg(means + (size_t)g * 3); py = __ldg(means + (size_t)g * 3 + 1); pz = __ldg(means + (size_t)g * 3 + 2); }
    Mat3 Rq;
    float s0 = 0.f, s1 = 0.f;
    bool have = false;
    for (uint32_t c0 = 0; c0 < C; c0 += kSurfelMaxCams) {
        const uint32_t nc = min((uint32_t)kSurfelMaxCams, C - c0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nc * 12; i += blockDim.x) s_vm[i / 12][i % 12] = viewmats[(size_t)(c0 + i / 12) * 16 + i % 12];
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
            const float *K = Ks + (size_t)(c0 + i) * 9;
            s_k[i][0] = K[0]; s_k[i][1] = K[4]; s_k[i][2] = K[2]; s_k[i][3] = K[5];
        }
        __syncthreads();
        if (g >= N) continue;
        for (uint32_t ci = 0; ci < nc; ++ci) {
            const float *R = s_vm[ci];
            const size_t idx = (size_t)(c0 + ci) * N + g;
            const float fx = s_k[ci][0], fy = s_k[ci][1], cx = s_k[ci][2], cy = s_k[ci][3];
            const float mcx = R[0] * px + R[1] * py + R[2] * pz + R[3];
            const float mcy = R[4] * px + R[5] * py + R[6] * pz + R[7];
            const float mcz = R[8] * px + R[9] * py + R[10] * pz + R[11];
            int32_t rad = 0;
            float o_m[2] = {0.f, 0.f}, o_d = 0.f, o_T[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o_n[3] = {0.f, 0.f, 0.f};
            if (!(mcz < near_plane || mcz > far_plane)) {  // fully_fused_projection_2dgs_fwd.cu:130-133
                if (!have) {
                    have = true;
                    const float4 q = ldg4(quats + (size_t)g * 4);
                    Rq = quat_to_rot(q.x, q.y, q.z, q.w, nullptr);
                    s0 = __ldg(scales + (size_t)g * 3); s1 = __ldg(scales + (size_t)g * 3 + 1);
                }
                // The expressions below keep the SHAPE of the reference's glm products, zero terms included: which product of a
                // sum becomes the FMUL and which the FFMAs is decided by that shape, and radii / tile lists must agree bit for bit.
                // RS_camera = ((R_cam Rq) diag(s0, s1, 1)): columns a0 (tangent u), a1 (tangent v), a2 (normal)  (:138-142)
                float a0[3], a1[3], a2[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float r0 = R[r * 4], r1 = R[r * 4 + 1], r2 = R[r * 4 + 2];
                    const float ab0 = r0 * Rq.m[0][0] + r1 * Rq.m[1][0] + r2 * Rq.m[2][0];
                    const float ab1 = r0 * Rq.m[0][1] + r1 * Rq.m[1][1] + r2 * Rq.m[2][1];
                    const float ab2 = r0 * Rq.m[0][2] + r1 * Rq.m[1][2] + r2 * Rq.m[2][2];
                    a0[r] = ab0 * s0 + ab1 * 0.f + ab2 * 0.f;
                    a1[r] = ab0 * 0.f + ab1 * s1 + ab2 * 0.f;
                    a2[r] = ab0 * 0.f + ab1 * 0.f + ab2 * 1.f;
                }
                // M = [a0 a1 mean_c]^T K^T (:144-160): its columns are the rows of the ray-splat transform K [a0 a1 mean_c]
                const float wh[3][3] = {{a0[0], a0[1], a0[2]}, {a1[0], a1[1], a1[2]}, {mcx, mcy, mcz}};  // wh[c] = column c of WH
                float M0[3], M1[3], M2[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    M0[r] = wh[r][0] * fx + wh[r][1] * 0.f + wh[r][2] * cx;
                    M1[r] = wh[r][0] * 0.f + wh[r][1] * fy + wh[r][2] * cy;
                    M2[r] = wh[r][0] * 0.f + wh[r][1] * 0.f + wh[r][2] * 1.f;
                }
                // bounding box (:195-212): temp_point = (1, 1, -1), f = temp_point / distance
                const float distance = (1.f * M2[0]) * M2[0] + (1.f * M2[1]) * M2[1] + (-1.f * M2[2]) * M2[2];
                if (distance != 0.f) {
                    const float fi = 1.f / distance;
                    const float f0 = fi * 1.f, f1 = fi * 1.f, f2 = fi * -1.f;
                    const float mx = (f0 * M0[0]) * M2[0] + (f1 * M0[1]) * M2[1] + (f2 * M0[2]) * M2[2];
                    const float my = (f0 * M1[0]) * M2[0] + (f1 * M1[1]) * M2[1] + (f2 * M1[2]) * M2[2];
                    const float tx = (f0 * M0[0]) * M0[0] + (f1 * M0[1]) * M0[1] + (f2 * M0[2]) * M0[2];
                    const float ty = (f0 * M1[0]) * M1[0] + (f1 * M1[1]) * M1[1] + (f2 * M1[2]) * M1[2];
                    const float hx = mx * mx - tx, hy = my * my - ty;
                    // (the reference's `max(1e-4, .)` promotes to double: square root, scaling and ceil run in float64, :214-215)
                    const float radius = (float)ceil(3.0 * sqrt(fmax(1e-4, (double)fmaxf(hx, hy))));
                    bool ok = !(radius <= radius_clip);
                    ok = ok && !(mx + radius <= 0 || mx - radius >= W || my + radius <= 0 || my - radius >= H);
                    if (ok) {
                        rad = (int32_t)radius;
                        o_m[0] = mx; o_m[1] = my; o_d = mcz;
#pragma unroll
                        for (int k = 0; k < 3; ++k) { o_T[k] = M0[k]; o_T[3 + k] = M1[k]; o_T[6 + k] = M2[k]; }
                        // dual-visible normal (:228-232)
                        const float facing = -(a2[0] * mcx + a2[1] * mcy + a2[2] * mcz);
                        const float sgn = facing > 0.f ? 1.f : -1.f;
                        o_n[0] = a2[0] * sgn; o_n[1] = a2[1] * sgn; o_n[2] = a2[2] * sgn;
                    }
                }
            }
            radii[idx] = rad;
            *reinterpret_cast<float2 *>(means2d + idx * 2) = make_float2(o_m[0], o_m[1]);
            depths[idx] = o_d;
#pragma unroll
            for (int k = 0; k < 9; ++k) ray_transforms[idx * 9 + k] = o_T[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) normals[idx * 3 + k] = o_n[k];
        }
    }
}

// ----------------------------------------------------------------------------------------------- projection backward
// v_T [C,N,9] is dL/d ray_transforms (row-major, from the blend); compute_ray_transforms_aabb_vjp (2dgs.cuh:12-75).
__global__ void __launch_bounds__(kThreads)
surfel_projection_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ quats,
                             const float *__restrict__ scales, const float *__restrict__ viewmats, const float *__restrict__ Ks,
                             const int32_t *__restrict__ radii, const float *__restrict__ ray_transforms,
                             const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                             const float *__restrict__ v_normals, const float *__restrict__ v_T, float *__restrict__ v_means,
                             float *__restrict__ v_quats, float *__restrict__ v_scales) {
    __shared__ float s_vm[kSurfelMaxCams][12];
    __shared__ float s_k[kSurfelMaxCams][4];
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    float px = 0.f, py = 0.f, pz = 0.f, s0 = 0.f, s1 = 0.f, inv_norm = 1.f;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    Mat3 Rq;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Rq.m[i][j] = 0.f;
    bool have = false;
    float vmean[3] = {0.f, 0.f, 0.f};
    float vRq[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};  // dL/dRq, row-major
    float vs0 = 0.f, vs1 = 0.f;
    for (uint32_t c0 = 0; c0 < C; c0 += kSurfelMaxCams) {
        const uint32_t nc = min((uint32_t)kSurfelMaxCams, C - c0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nc * 12; i += blockDim.x) s_vm[i / 12][i % 12] = viewmats[(size_t)(c0 + i / 12) * 16 + i % 12];
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
            const float *K = Ks + (size_t)(c0 + i) * 9;
            s_k[i][0] = K[0]; s_k[i][1] = K[4]; s_k[i][2] = K[2]; s_k[i][3] = K[5];
        }
        __syncthreads();
        if (g >= N) continue;
        for (uint32_t ci = 0; ci < nc; ++ci) {
            const size_t idx = (size_t)(c0 + ci) * N + g;
            if (radii[idx] <= 0) continue;
            const float *R = s_vm[ci];
            const float fx = s_k[ci][0], fy = s_k[ci][1], cx = s_k[ci][2], cy = s_k[ci][3];
            if (!have) {
                have = true;
                px = __ldg(means + (size_t)g * 3); py = __ldg(means + (size_t)g * 3 + 1); pz = __ldg(means + (size_t)g * 3 + 2);
                q = ldg4(quats + (size_t)g * 4);
                s0 = __ldg(scales + (size_t)g * 3); s1 = __ldg(scales + (size_t)g * 3 + 1);
                Rq = quat_to_rot(q.x, q.y, q.z, q.w, &inv_norm);
            }
            const float mcx = R[0] * px + R[1] * py + R[2] * pz + R[3];
            const float mcy = R[4] * px + R[5] * py + R[6] * pz + R[7];
            const float mcz = R[8] * px + R[9] * py + R[10] * pz + R[11];
            float T[9], V[9];  // V[3 i + j] = dL / dT[i][j]
#pragma unroll
            for (int k = 0; k < 9; ++k) { T[k] = ray_transforms[idx * 9 + k]; V[k] = v_T[idx * 9 + k]; }
            V[8] += v_depths[idx];  // depth = T[2][2]  (fully_fused_projection_2dgs_bwd.cu:95)
            const float vm0 = v_means2d[idx * 2], vm1 = v_means2d[idx * 2 + 1];
            if (vm0 != 0.f || vm1 != 0.f) {  // 2dgs.cuh:27-53 (the reference's own, partial, derivative of the bounding-box centre)
                const float distance = T[6] * T[6] + T[7] * T[7] - T[8] * T[8];
                const float f = 1.f / distance;
                const float d6 = f * T[6], d7 = f * T[7], d8 = -f * T[8];
                const float k6 = f - 2.f * f * f * T[6] * T[6], k7 = f - 2.f * f * f * T[7] * T[7];
                const float k8 = f + 2.f * f * f * T[8] * T[8];
                V[0] += vm0 * d6; V[1] += vm0 * d7; V[2] += vm0 * d8;
                V[3] += vm1 * d6; V[4] += vm1 * d7; V[5] += vm1 * d8;
                V[6] += vm0 * (T[0] * k6) + vm1 * (T[3] * k6);
                V[7] += vm0 * (T[1] * k7) + vm1 * (T[4] * k7);
                V[8] += vm0 * (-T[2] * k8) + vm1 * (-T[5] * k8);
            }
            // dL/dWH = K^T V (T = K WH): column c of WH is (a0 | a1 | mean_c)
            float G[3][3];  // G[r][c] = dL / dWH[r][c]
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                G[0][c] = fx * V[c];
                G[1][c] = fy * V[3 + c];
                G[2][c] = cx * V[c] + cy * V[3 + c] + V[6 + c];
            }
            // back through the camera rotation: v_RS = R_cam^T G  (columns: d/d(Rq col 0 * s0), d/d(Rq col 1 * s1), d/d mean)
            float vRS[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) vRS[r][c] = R[r] * G[0][c] + R[4 + r] * G[1][c] + R[8 + r] * G[2][c];
            // normal: tn = R_cam Rq[:,2], flipped towards the camera (2dgs.cuh:59-66)
            float tn[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tn[r] = R[r * 4] * Rq.m[0][2] + R[r * 4 + 1] * Rq.m[1][2] + R[r * 4 + 2] * Rq.m[2][2];
            const float facing = -(tn[0] * mcx + tn[1] * mcy + tn[2] * mcz);
            const float sgn = facing > 0.f ? 1.f : -1.f;
            const float vn0 = v_normals[idx * 3], vn1 = v_normals[idx * 3 + 1], vn2 = v_normals[idx * 3 + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float vtn = (R[r] * vn0 + R[4 + r] * vn1 + R[8 + r] * vn2) * sgn;
                vRq[r][0] += vRS[r][0] * s0;
                vRq[r][1] += vRS[r][1] * s1;
                vRq[r][2] += vtn;
                vmean[r] += vRS[r][2];
            }
            vs0 += vRS[0][0] * Rq.m[0][0] + vRS[1][0] * Rq.m[1][0] + vRS[2][0] * Rq.m[2][0];
            vs1 += vRS[0][1] * Rq.m[0][1] + vRS[1][1] * Rq.m[1][1] + vRS[2][1] * Rq.m[2][1];
        }
    }
    if (g >= N) return;
    v_means[(size_t)g * 3] = vmean[0]; v_means[(size_t)g * 3 + 1] = vmean[1]; v_means[(size_t)g * 3 + 2] = vmean[2];
    v_scales[(size_t)g * 3] = vs0; v_scales[(size_t)g * 3 + 1] = vs1; v_scales[(size_t)g * 3 + 2] = 0.f;
    if (!have) {
        stg4(v_quats + (size_t)g * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    // quat.cuh:35-57 (the quaternion is normalised inside quat_to_rot)
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
}

// ------------------------------------------------------------------------------------------------------------ blend
struct SurfelGeom {
    uint32_t C, W, H, ts, tw, th;
    int64_t n_isects;
};

// one (pixel, Gaussian) evaluation: rasterize_to_pixels_2dgs_fwd.cu:232-266
struct SurfelHit {
    float hu[3], hv[3], rc[3];  // homogeneous plane parameters and their cross product (the ray of intersection)
    float sx, sy, dx, dy;       // intersection in splat coordinates; offset of the projected centre
    float gw3, gw2, vis, alpha;
    bool ok;
};
__device__ __forceinline__ SurfelHit surfel_eval(float px, float py, const float4 &A, const float4 &B, const float4 &Cq,
                                                 const float4 &D) {
    SurfelHit h;
    const float u[3] = {B.x, B.y, B.z}, v[3] = {B.w, Cq.x, Cq.y}, w[3] = {Cq.z, Cq.w, D.x};
#pragma unroll
    for (int k = 0; k < 3; ++k) { h.hu[k] = px * w[k] - u[k]; h.hv[k] = py * w[k] - v[k]; }
    h.rc[0] = h.hu[1] * h.hv[2] - h.hv[1] * h.hu[2];
    h.rc[1] = h.hu[2] * h.hv[0] - h.hv[2] * h.hu[0];
    h.rc[2] = h.hu[0] * h.hv[1] - h.hv[0] * h.hu[1];
    h.ok = h.rc[2] != 0.f;
    h.sx = h.rc[0] / h.rc[2];
    h.sy = h.rc[1] / h.rc[2];
    h.gw3 = h.sx * h.sx + h.sy * h.sy;
    h.dx = A.x - px; h.dy = A.y - py;
    h.gw2 = kFilterInvSquare * (h.dx * h.dx + h.dy * h.dy);
    const float sigma = 0.5f * fminf(h.gw3, h.gw2);
    h.vis = __expf(-sigma);
    h.alpha = fminf(0.999f, A.z * h.vis);
    if (sigma < 0.f || h.alpha < 1.f / 255.f) h.ok = false;
    return h;
}

// staging: A = (x, y, opacity, flatten id), B = (u.xyz, v.x), Cq = (v.yz, w.xy), D = (w.z, normal.xyz)
__device__ __forceinline__ void surfel_stage(float4 *sA, uint32_t slot, uint32_t block_size, int32_t g, const float2 *means2d,
                                             const float *ray_transforms, const float *opacities, const float *normals) {
    const float2 xy = __ldg(means2d + g);
    const float *T = ray_transforms + (size_t)g * 9;
    const float *n = normals + (size_t)g * 3;
    sA[slot] = make_float4(xy.x, xy.y, __ldg(opacities + g), __int_as_float(g));
    sA[block_size + slot] = make_float4(__ldg(T), __ldg(T + 1), __ldg(T + 2), __ldg(T + 3));
    sA[2 * block_size + slot] = make_float4(__ldg(T + 4), __ldg(T + 5), __ldg(T + 6), __ldg(T + 7));
    sA[3 * block_size + slot] = make_float4(__ldg(T + 8), n ? __ldg(n) : 0.f, n ? __ldg(n + 1) : 0.f, n ? __ldg(n + 2) : 0.f);
}

template <int CDIM>
__global__ void surfel_blend_fwd_kernel(SurfelGeom G, const float2 *__restrict__ means2d, const float *__restrict__ ray_transforms,
                                        const float *__restrict__ colors, const float *__restrict__ opacities,
                                        const float *__restrict__ normals, const float *__restrict__ backgrounds,
                                        const uint8_t *__restrict__ masks, const int32_t *__restrict__ tile_offsets,
                                        const int32_t *__restrict__ flatten_ids, float *__restrict__ render_colors,
                                        float *__restrict__ render_alphas, float *__restrict__ render_normals,
                                        float *__restrict__ render_distort, float *__restrict__ render_median,
                                        int32_t *__restrict__ last_ids, int32_t *__restrict__ median_ids) {
    extern __shared__ float4 smem4[];
    const uint32_t block_size = blockDim.x * blockDim.y;
    const uint32_t cam = blockIdx.x, tile_y = blockIdx.y, tile_x = blockIdx.z;
    const uint32_t tile_id = tile_y * G.tw + tile_x;
    const uint32_t tr = threadIdx.y * blockDim.x + threadIdx.x;
    const uint32_t i = tile_y * G.ts + threadIdx.y, j = tile_x * G.ts + threadIdx.x;
    const bool inside = i < G.H && j < G.W;
    const size_t pix = ((size_t)cam * G.H + i) * G.W + j;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;
    if (masks != nullptr && !masks[ct]) {  // :150-156 (the other outputs of a masked tile are zero-filled by the host)
        if (inside)
#pragma unroll
            for (int k = 0; k < CDIM; ++k) render_colors[pix * CDIM + k] = bg ? bg[k] : 0.f;
        return;
    }
    const int32_t range_start = tile_offsets[ct];
    const int32_t range_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];
    const uint32_t num_batches = ((uint32_t)(range_end - range_start) + block_size - 1) / block_size;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    bool done = !inside;
    float T = 1.f, distort = 0.f, accum_vis_depth = 0.f, median_depth = 0.f;
    uint32_t cur_idx = 0, median_idx = 0;
    float pix_out[CDIM], normal_out[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < CDIM; ++k) pix_out[k] = 0.f;
    for (uint32_t b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= (int)block_size) break;
        const uint32_t batch_start = (uint32_t)range_start + block_size * b;
        if (batch_start + tr < (uint32_t)range_end)
            surfel_stage(smem4, tr, block_size, flatten_ids[batch_start + tr], means2d, ray_transforms, opacities, normals);
        __syncthreads();
        const uint32_t batch_size = min(block_size, (uint32_t)range_end - batch_start);
        for (uint32_t t = 0; t < batch_size && !done; ++t) {
            const float4 A = smem4[t], D = smem4[3 * block_size + t];
            const SurfelHit h = surfel_eval(px, py, A, smem4[block_size + t], smem4[2 * block_size + t], D);
            if (!h.ok) continue;
            const float next_T = T * (1.f - h.alpha);
            if (next_T <= 1e-4f) { done = true; break; }
            const float vis = h.alpha * T;
            const float *c_ptr = colors + (size_t)__float_as_int(A.w) * CDIM;
            float depth = 0.f;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) {
                const float c = __ldg(c_ptr + k);
                pix_out[k] += c * vis;
                if (k == CDIM - 1) depth = c;
            }
            normal_out[0] += D.y * vis; normal_out[1] += D.z * vis; normal_out[2] += D.w * vis;
            {  // depth distortion (Mip-NeRF 360) on the last channel, :283-291
                const float d0 = vis * depth * (1.f - T), d1 = vis * accum_vis_depth;
                distort += 2.f * (d0 - d1);
                accum_vis_depth += vis * depth;
            }
            if (T > 0.5f) { median_depth = depth; median_idx = batch_start + t; }
            cur_idx = batch_start + t;
            T = next_T;
        }
    }
    if (inside) {
        render_alphas[pix] = 1.f - T;
#pragma unroll
        for (int k = 0; k < CDIM; ++k) render_colors[pix * CDIM + k] = bg ? (pix_out[k] + T * bg[k]) : pix_out[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) render_normals[pix * 3 + k] = normal_out[k];
        last_ids[pix] = (int32_t)cur_idx;
        render_distort[pix] = distort;
        render_median[pix] = median_depth;
        median_ids[pix] = (int32_t)median_idx;
    }
}

// where value `off` of a Gaussian's partial sums [rgb(CDIM) | normal(3) | v_T(9) | xy(2) | opacity | abs xy(2)] is accumulated
struct SurfelGradOut {
    float *v_colors, *v_normals, *v_T, *v_means2d, *v_opacities, *v_abs;
};
template <int CDIM>
__device__ __forceinline__ float *surfel_grad_addr(const SurfelGradOut &O, uint32_t off, size_t g) {
    if (off < (uint32_t)CDIM) return O.v_colors + g * CDIM + off;
    off -= CDIM;
    if (off < 3) return O.v_normals + g * 3 + off;
    off -= 3;
    if (off < 9) return O.v_T + g * 9 + off;
    off -= 9;
    if (off < 2) return O.v_means2d + g * 2 + off;
    off -= 2;
    if (off < 1) return O.v_opacities + g;
    off -= 1;
    return O.v_abs ? O.v_abs + g * 2 + off : nullptr;
}

template <int CDIM, bool ABS>
__global__ void surfel_blend_bwd_kernel(SurfelGeom G, const float2 *__restrict__ means2d, const float *__restrict__ ray_transforms,
                                        const float *__restrict__ colors, const float *__restrict__ normals,
                                        const float *__restrict__ opacities, const float *__restrict__ backgrounds,
                                        const uint8_t *__restrict__ masks, const int32_t *__restrict__ tile_offsets,
                                        const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_colors,
                                        const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                                        const int32_t *__restrict__ median_ids, const float *__restrict__ v_render_colors,
                                        const float *__restrict__ v_render_alphas, const float *__restrict__ v_render_normals,
                                        const float *__restrict__ v_render_distort, const float *__restrict__ v_render_median,
                                        const SurfelGradOut O) {
    constexpr int NV = CDIM + 15 + (ABS ? 2 : 0);
    constexpr int NGRP = (NV + 7) / 8;
    extern __shared__ float4 smem4[];
    const uint32_t block_size = blockDim.x * blockDim.y;
    float *s_rgb = reinterpret_cast<float *>(smem4 + 4 * block_size);  // [block_size][CDIM]
    const uint32_t cam = blockIdx.x, tile_y = blockIdx.y, tile_x = blockIdx.z;
    const uint32_t tile_id = tile_y * G.tw + tile_x;
    const uint32_t tr = threadIdx.y * blockDim.x + threadIdx.x, lane = tr & 31;
    const uint32_t i = tile_y * G.ts + threadIdx.y, j = tile_x * G.ts + threadIdx.x;
    const bool inside = i < G.H && j < G.W;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    if (masks != nullptr && !masks[ct]) return;
    const size_t pix = (size_t)cam * G.H * G.W + min((size_t)i * G.W + j, (size_t)G.W * G.H - 1);
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;
    const int32_t range_start = tile_offsets[ct];
    const int32_t range_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];
    const uint32_t num_batches = ((uint32_t)(range_end - range_start) + block_size - 1) / block_size;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;

    const float T_final = 1.f - render_alphas[pix];
    float T = T_final;
    float buffer[CDIM], buffer_n[3] = {0.f, 0.f, 0.f}, v_c[CDIM], v_n[3];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        buffer[k] = 0.f;
        v_c[k] = v_render_colors[pix * CDIM + k];
        if (bg) bg_dot += bg[k] * v_c[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) v_n[k] = v_render_normals[pix * 3 + k];
    const float v_a = v_render_alphas[pix];
    const int32_t bin_final = inside ? last_ids[pix] : 0;
    const int32_t median_idx = inside ? median_ids[pix] : 0;
    // distortion state (:196-209)
    const float v_distort = v_render_distort[pix];
    float accum_d_buffer = render_colors[pix * CDIM + CDIM - 1], accum_w_buffer = render_alphas[pix];
    const float accum_d = accum_d_buffer, accum_w = accum_w_buffer;
    float distort_buffer = 0.f;
    const float v_median = v_render_median[pix];
    const int32_t warp_bin_final = __reduce_max_sync(0xffffffffu, bin_final);
    // which of the NV totals this lane publishes after the group reductions
    const uint32_t slot8 = spread_slot<8>(lane);
    const bool pub8 = (lane & 3) == 0;

    for (uint32_t b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int32_t batch_end = range_end - 1 - (int32_t)(block_size * b);
        const int32_t batch_size = min((int32_t)block_size, batch_end + 1 - range_start);
        const int32_t idx = batch_end - (int32_t)tr;
        if (idx >= range_start) {
            const int32_t g = flatten_ids[idx];
            surfel_stage(smem4, tr, block_size, g, means2d, ray_transforms, opacities, normals);
#pragma unroll
            for (int k = 0; k < CDIM; ++k) s_rgb[tr * CDIM + k] = __ldg(colors + (size_t)g * CDIM + k);
        }
        __syncthreads();
        for (int32_t t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && !(batch_end - t > bin_final);
            const float4 A = smem4[t], D = smem4[3 * block_size + t];
            SurfelHit h;
            if (valid) {
                h = surfel_eval(px, py, A, smem4[block_size + t], smem4[2 * block_size + t], D);
                valid = h.ok;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float vals[NGRP * 8];
#pragma unroll
            for (int k = 0; k < NGRP * 8; ++k) vals[k] = 0.f;
            if (valid) {
                const float opac = A.z;
                const float *rgb = s_rgb + t * CDIM;
                if (batch_end - t == median_idx) vals[CDIM - 1] += v_median;
                const float ra = 1.f / (1.f - h.alpha);
                T *= ra;
                const float fac = h.alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < CDIM; ++k) {
                    vals[k] += fac * v_c[k];
                    v_alpha += (rgb[k] * T - buffer[k] * ra) * v_c[k];
                }
                const float nrm[3] = {D.y, D.z, D.w};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    vals[CDIM + k] = fac * v_n[k];
                    v_alpha += (nrm[k] * T - buffer_n[k] * ra) * v_n[k];
                }
                v_alpha += T_final * ra * v_a;
                if (bg) v_alpha += -T_final * ra * bg_dot;
                {  // distortion, :338-353
                    const float depth = rgb[CDIM - 1];
                    const float dl_dw = 2.f * (2.f * (depth * accum_w_buffer - accum_d_buffer) + (accum_d - depth * accum_w));
                    v_alpha += (dl_dw * T - distort_buffer * ra) * v_distort;
                    accum_d_buffer -= fac * depth;
                    accum_w_buffer -= fac;
                    distort_buffer += dl_dw * fac;
                    vals[CDIM - 1] += 2.f * fac * (2.f - 2.f * T - accum_w + fac) * v_distort;
                }
                if (opac * h.vis <= 0.999f) {
                    const float v_G = opac * v_alpha;
                    if (h.gw3 <= h.gw2) {  // the ray-splat intersection carried the weight: gradient to the transform, :361-385
                        const float vsx = v_G * -h.vis * h.sx, vsy = v_G * -h.vis * h.sy;
                        const float a = vsx / h.rc[2], bq = vsy / h.rc[2];
                        const float vrc[3] = {a, bq, -(a * h.sx + bq * h.sy)};
                        // v_h_u = h_v x v_rc, v_h_v = v_rc x h_u
                        const float vhu[3] = {h.hv[1] * vrc[2] - vrc[1] * h.hv[2], h.hv[2] * vrc[0] - vrc[2] * h.hv[0],
                                              h.hv[0] * vrc[1] - vrc[0] * h.hv[1]};
                        const float vhv[3] = {vrc[1] * h.hu[2] - h.hu[1] * vrc[2], vrc[2] * h.hu[0] - h.hu[2] * vrc[0],
                                              vrc[0] * h.hu[1] - h.hu[0] * vrc[1]};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            vals[CDIM + 3 + k] = -vhu[k];
                            vals[CDIM + 6 + k] = -vhv[k];
                            vals[CDIM + 9 + k] = px * vhu[k] + py * vhv[k];
                        }
                    } else {  // the low-pass filter around the projected centre did, :386-395
                        const float gx = v_G * (-h.vis * kFilterInvSquare * h.dx), gy = v_G * (-h.vis * kFilterInvSquare * h.dy);
                        vals[CDIM + 12] = gx;
                        vals[CDIM + 13] = gy;
                        if (ABS) { vals[CDIM + 15] = fabsf(gx); vals[CDIM + 16] = fabsf(gy); }
                    }
                    vals[CDIM + 14] = h.vis * v_alpha;
                }
#pragma unroll
                for (int k = 0; k < CDIM; ++k) buffer[k] += rgb[k] * fac;
#pragma unroll
                for (int k = 0; k < 3; ++k) buffer_n[k] += nrm[k] * fac;
            }
            const size_t g = (size_t)__float_as_int(A.w);
#pragma unroll
            for (int gi = 0; gi < NGRP; ++gi) {
                reduce_spread<8>(vals + gi * 8, lane);
                const uint32_t off = gi * 8 + slot8;
                if (pub8 && off < (uint32_t)NV && vals[gi * 8] != 0.f) atomicAdd(surfel_grad_addr<CDIM>(O, off, g), vals[gi * 8]);
            }
        }
    }
}

// v_densify [M,2] from the finished sums (see the header of this file; reference: rasterize_to_pixels_2dgs_bwd.cu:608-615)
__global__ void __launch_bounds__(kThreads)
surfel_densify_kernel(uint32_t M, const float *__restrict__ ray_transforms, const float *__restrict__ v_T,
                      float *__restrict__ v_densify) {
    const uint32_t g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= M) return;
    const float depth = ray_transforms[(size_t)g * 9 + 8];
    *reinterpret_cast<float2 *>(v_densify + (size_t)g * 2) = make_float2(v_T[(size_t)g * 9 + 2] * depth, v_T[(size_t)g * 9 + 5] * depth);
}

// -------------------------------------------------------------------------------------- per-pixel contributor lists
struct SurfelRange {
    uint32_t C, N, W, H, ts, tw, th, range_start, range_end;
    int64_t n_isects;
};
template <bool FILL>
__global__ void surfel_indices_in_range_kernel(SurfelRange G, const float2 *__restrict__ means2d,
                                               const float *__restrict__ ray_transforms, const float *__restrict__ opacities,
                                               const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                                               const float *__restrict__ transmittances, const int64_t *__restrict__ chunk_starts,
                                               int32_t *__restrict__ chunk_cnts, int64_t *__restrict__ gaussian_ids,
                                               int64_t *__restrict__ pixel_ids) {
    extern __shared__ float4 smem4[];
    const uint32_t block_size = G.ts * G.ts;
    const uint32_t cam = blockIdx.x, tile_y = blockIdx.y, tile_x = blockIdx.z;
    const uint32_t tile_id = tile_y * G.tw + tile_x;
    const uint32_t tr = threadIdx.y * G.ts + threadIdx.x;
    const uint32_t i = tile_y * G.ts + threadIdx.y, j = tile_x * G.ts + threadIdx.x;
    const bool inside = i < G.H && j < G.W;
    const size_t pix = ((size_t)cam * G.H + i) * G.W + j;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    const int32_t isect_start = tile_offsets[ct];
    const int32_t isect_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];
    const uint32_t num_batches = ((uint32_t)(isect_end - isect_start) + block_size - 1) / block_size;
    if (G.range_start >= num_batches) {
        if (!FILL && inside) chunk_cnts[pix] = 0;
        return;
    }
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    float trans = inside ? transmittances[pix] : 0.f;
    const int64_t base = (FILL && inside) ? chunk_starts[pix] : 0;
    bool done = !inside;
    int32_t cnt = 0;
    const uint32_t b_end = min(G.range_end, num_batches);
    for (uint32_t b = G.range_start; b < b_end; ++b) {
        if (__syncthreads_count(done) >= (int)block_size) break;
        const uint32_t batch_start = (uint32_t)isect_start + block_size * b;
        if (batch_start + tr < (uint32_t)isect_end)
            surfel_stage(smem4, tr, block_size, flatten_ids[batch_start + tr], means2d, ray_transforms, opacities, nullptr);
        __syncthreads();
        const uint32_t batch_size = min(block_size, (uint32_t)isect_end - batch_start);
        for (uint32_t t = 0; t < batch_size && !done; ++t) {
            const float4 A = smem4[t];
            const SurfelHit h = surfel_eval(px, py, A, smem4[block_size + t], smem4[2 * block_size + t], smem4[3 * block_size + t]);
            if (!h.ok) continue;
            const float next_trans = trans * (1.f - h.alpha);
            if (next_trans <= 1e-4f) { done = true; break; }
            if (FILL) {
                gaussian_ids[base + cnt] = (int64_t)((uint32_t)__float_as_int(A.w) % G.N);
                pixel_ids[base + cnt] = (int64_t)pix;
            }
            ++cnt;
            trans = next_trans;
        }
    }
    if (!FILL && inside) chunk_cnts[pix] = cnt;
}

static int surfel_check_geom(const char *who, uint32_t tile_size, uint32_t W, uint32_t H, uint32_t tw, uint32_t th) {
    if (tile_size == 0 || tile_size > 32) return failf(2, "%s: tile_size must be in [1, 32]", who);
    if (tw * tile_size < W || th * tile_size < H)
        return failf(2, "%s: tile grid does not cover the image (Assert Failed: tile_height * tile_size >= image_height)", who);
    if (th > 65535 || tw > 65535) return failf(2, "%s: tile grid too large", who);
    return 0;
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_projection_2dgs_fwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                          const float *viewmats, const float *Ks, uint32_t W, uint32_t H, float near_plane,
                                          float far_plane, float radius_clip, int32_t *radii, float *means2d, float *depths,
                                          float *ray_transforms, float *normals, gsb200_stream_t stream) {
    if (C == 0 || N == 0) return 0;
    if (!means || !quats || !scales || !viewmats || !Ks || !radii || !means2d || !depths || !ray_transforms || !normals)
        return fail(2, "gsb200_projection_2dgs_fwd: null pointer");
    surfel_projection_fwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, viewmats, Ks, (float)W, (float)H, near_plane, far_plane, radius_clip, radii, means2d, depths,
        ray_transforms, normals);
    return check_launch("projection_2dgs_fwd");
}

extern "C" int gsb200_projection_2dgs_bwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                          const float *viewmats, const float *Ks, const int32_t *radii,
                                          const float *ray_transforms, const float *v_means2d, const float *v_depths,
                                          const float *v_normals, const float *v_ray_transforms, float *v_means, float *v_quats,
                                          float *v_scales, gsb200_stream_t stream) {
    if (C == 0 || N == 0) return 0;
    if (!means || !quats || !scales || !viewmats || !Ks || !radii || !ray_transforms || !v_means2d || !v_depths || !v_normals ||
        !v_ray_transforms || !v_means || !v_quats || !v_scales)
        return fail(2, "gsb200_projection_2dgs_bwd: null pointer");
    surfel_projection_bwd_kernel<<<cdiv(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, viewmats, Ks, radii, ray_transforms, v_means2d, v_depths, v_normals, v_ray_transforms, v_means,
        v_quats, v_scales);
    return check_launch("projection_2dgs_bwd");
}

#define GSB_SURFEL_CDIM(D, CALL)                                                                  \
    switch (D) {                                                                                  \
        case 1: { constexpr int CD = 1; CALL; break; }                                            \
        case 2: { constexpr int CD = 2; CALL; break; }                                            \
        case 3: { constexpr int CD = 3; CALL; break; }                                            \
        case 4: { constexpr int CD = 4; CALL; break; }                                            \
        case 8: { constexpr int CD = 8; CALL; break; }                                            \
        case 16: { constexpr int CD = 16; CALL; break; }                                          \
        case 32: { constexpr int CD = 32; CALL; break; }                                          \
        default: return failf(2, "%s: channels must be one of 1, 2, 3, 4, 8, 16, 32 (the Python wrapper pads)", who); \
    }

extern "C" int gsb200_rasterize_2dgs_fwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                                         const float *ray_transforms, const float *colors, const float *opacities,
                                         const float *normals, const float *backgrounds, const uint8_t *masks, uint32_t W,
                                         uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                         const int32_t *tile_offsets, const int32_t *flatten_ids, float *render_colors,
                                         float *render_alphas, float *render_normals, float *render_distort, float *render_median,
                                         int32_t *last_ids, int32_t *median_ids, gsb200_stream_t stream) {
    const char *who = "gsb200_rasterize_2dgs_fwd";
    (void)M;
    if (int e = surfel_check_geom(who, tile_size, W, H, tile_width, tile_height)) return e;
    if (C == 0 || W == 0 || H == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t P = (size_t)C * H * W;
    if (masks != nullptr) {  // a masked tile only gets its colours written (the reference leaves the rest uninitialised)
        GSB_CUDA_OK(cudaMemsetAsync(render_alphas, 0, sizeof(float) * P, s));
        GSB_CUDA_OK(cudaMemsetAsync(render_normals, 0, sizeof(float) * 3 * P, s));
        GSB_CUDA_OK(cudaMemsetAsync(render_distort, 0, sizeof(float) * P, s));
        GSB_CUDA_OK(cudaMemsetAsync(render_median, 0, sizeof(float) * P, s));
        GSB_CUDA_OK(cudaMemsetAsync(last_ids, 0, sizeof(int32_t) * P, s));
        GSB_CUDA_OK(cudaMemsetAsync(median_ids, 0, sizeof(int32_t) * P, s));
    }
    const SurfelGeom G{C, W, H, tile_size, tile_width, tile_height, n_isects};
    const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
    const size_t smem = (size_t)tile_size * tile_size * 4 * sizeof(float4);
#define GSB_SURFEL_FWD                                                                                                      \
    do {                                                                                                                    \
        if (smem > 48 * 1024)                                                                                               \
            GSB_CUDA_OK(cudaFuncSetAttribute(surfel_blend_fwd_kernel<CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        surfel_blend_fwd_kernel<CD><<<blocks, threads, smem, s>>>(                                                           \
            G, reinterpret_cast<const float2 *>(means2d), ray_transforms, colors, opacities, normals, backgrounds, masks,    \
            tile_offsets, flatten_ids, render_colors, render_alphas, render_normals, render_distort, render_median, last_ids, \
            median_ids);                                                                                                     \
    } while (0)
    GSB_SURFEL_CDIM(D, GSB_SURFEL_FWD)
#undef GSB_SURFEL_FWD
    return check_launch("rasterize_2dgs_fwd");
}

extern "C" int gsb200_rasterize_2dgs_bwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                                         const float *ray_transforms, const float *colors, const float *opacities,
                                         const float *normals, const float *backgrounds, const uint8_t *masks, uint32_t W,
                                         uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                         const int32_t *tile_offsets, const int32_t *flatten_ids, const float *render_colors,
                                         const float *render_alphas, const int32_t *last_ids, const int32_t *median_ids,
                                         const float *v_render_colors, const float *v_render_alphas, const float *v_render_normals,
                                         const float *v_render_distort, const float *v_render_median, float *v_means2d_abs,
                                         float *v_means2d, float *v_ray_transforms, float *v_colors, float *v_opacities,
                                         float *v_normals, float *v_densify, gsb200_stream_t stream) {
    const char *who = "gsb200_rasterize_2dgs_bwd";
    if (int e = surfel_check_geom(who, tile_size, W, H, tile_width, tile_height)) return e;
    if (M == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_OK(cudaMemsetAsync(v_means2d, 0, sizeof(float) * 2 * (size_t)M, s));
    GSB_CUDA_OK(cudaMemsetAsync(v_ray_transforms, 0, sizeof(float) * 9 * (size_t)M, s));
    GSB_CUDA_OK(cudaMemsetAsync(v_colors, 0, sizeof(float) * D * (size_t)M, s));
    GSB_CUDA_OK(cudaMemsetAsync(v_opacities, 0, sizeof(float) * (size_t)M, s));
    GSB_CUDA_OK(cudaMemsetAsync(v_normals, 0, sizeof(float) * 3 * (size_t)M, s));
    if (v_means2d_abs != nullptr) GSB_CUDA_OK(cudaMemsetAsync(v_means2d_abs, 0, sizeof(float) * 2 * (size_t)M, s));
    if (C != 0 && W != 0 && H != 0 && n_isects > 0) {
        const SurfelGeom G{C, W, H, tile_size, tile_width, tile_height, n_isects};
        const SurfelGradOut O{v_colors, v_normals, v_ray_transforms, v_means2d, v_opacities, v_means2d_abs};
        const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
        const size_t smem = (size_t)tile_size * tile_size * (4 * sizeof(float4) + D * sizeof(float));
        if ((tile_size * tile_size) % 32 != 0)
            return failf(2, "%s: tile_size^2 must be a multiple of 32 (warp reductions)", who);
#define GSB_SURFEL_BWD_T(ABSV)                                                                                              \
    do {                                                                                                                    \
        if (smem > 48 * 1024)                                                                                               \
            GSB_CUDA_OK(cudaFuncSetAttribute(surfel_blend_bwd_kernel<CD, ABSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                                   \
        surfel_blend_bwd_kernel<CD, ABSV><<<blocks, threads, smem, s>>>(                                                     \
            G, reinterpret_cast<const float2 *>(means2d), ray_transforms, colors, normals, opacities, backgrounds, masks,    \
            tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, median_ids, v_render_colors, v_render_alphas, \
            v_render_normals, v_render_distort, v_render_median, O);                                                         \
    } while (0)
#define GSB_SURFEL_BWD                                                              \
    do {                                                                            \
        if (v_means2d_abs != nullptr) GSB_SURFEL_BWD_T(true); else GSB_SURFEL_BWD_T(false); \
    } while (0)
        GSB_SURFEL_CDIM(D, GSB_SURFEL_BWD)
#undef GSB_SURFEL_BWD
#undef GSB_SURFEL_BWD_T
        if (int e = check_launch("rasterize_2dgs_bwd")) return e;
    }
    if (v_densify != nullptr) {
        surfel_densify_kernel<<<cdiv(M, kThreads), kThreads, 0, s>>>(M, ray_transforms, v_ray_transforms, v_densify);
        return check_launch("surfel_densify");
    }
    return 0;
}

extern "C" int gsb200_indices_in_range_2dgs_count(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                                  const float *means2d, const float *ray_transforms, const float *opacities,
                                                  uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                                                  uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                                  const float *transmittances, int32_t *chunk_cnts, gsb200_stream_t stream) {
    const char *who = "gsb200_indices_in_range_2dgs_count";
    if (int rc = surfel_check_geom(who, tile_size, W, H, tile_width, tile_height)) return rc;
    if (C == 0 || W == 0 || H == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_isects == 0 || tile_width * tile_height == 0) {
        GSB_CUDA_OK(cudaMemsetAsync(chunk_cnts, 0, sizeof(int32_t) * (size_t)C * H * W, s));
        return 0;
    }
    const SurfelRange G{C, N, W, H, tile_size, tile_width, tile_height, range_start, range_end, n_isects};
    const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
    const size_t smem = (size_t)tile_size * tile_size * 4 * sizeof(float4);
    if (smem > 48 * 1024)
        GSB_CUDA_OK(cudaFuncSetAttribute(surfel_indices_in_range_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    surfel_indices_in_range_kernel<false><<<blocks, threads, smem, s>>>(G, reinterpret_cast<const float2 *>(means2d), ray_transforms,
                                                                        opacities, tile_offsets, flatten_ids, transmittances,
                                                                        nullptr, chunk_cnts, nullptr, nullptr);
    return check_launch("indices_in_range_2dgs_count");
}

extern "C" int gsb200_indices_in_range_2dgs_fill(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                                 const float *means2d, const float *ray_transforms, const float *opacities,
                                                 uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                                                 uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                                 const float *transmittances, const int64_t *chunk_starts, int64_t *gaussian_ids,
                                                 int64_t *pixel_ids, gsb200_stream_t stream) {
    const char *who = "gsb200_indices_in_range_2dgs_fill";
    if (int rc = surfel_check_geom(who, tile_size, W, H, tile_width, tile_height)) return rc;
    if (C == 0 || W == 0 || H == 0 || n_isects == 0 || tile_width * tile_height == 0) return 0;
    if (chunk_starts == nullptr || gaussian_ids == nullptr || pixel_ids == nullptr) return failf(2, "%s: null pointer", who);
    const SurfelRange G{C, N, W, H, tile_size, tile_width, tile_height, range_start, range_end, n_isects};
    const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
    const size_t smem = (size_t)tile_size * tile_size * 4 * sizeof(float4);
    if (smem > 48 * 1024)
        GSB_CUDA_OK(cudaFuncSetAttribute(surfel_indices_in_range_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    surfel_indices_in_range_kernel<true><<<blocks, threads, smem, (cudaStream_t)stream>>>(
        G, reinterpret_cast<const float2 *>(means2d), ray_transforms, opacities, tile_offsets, flatten_ids, transmittances,
        chunk_starts, nullptr, gaussian_ids, pixel_ids);
    return check_launch("indices_in_range_2dgs_fill");
}
