// Device math shared by the projection kernels (projection.cu) and the un-fused geometry ops (geom_ops.cu):
// quaternion -> rotation, covariance construction / transforms, camera projection Jacobians
// (reference device code: gsplat/cuda/include/{quat,transform,proj,quat_scale_to_covar_preci}.cuh).
#pragma once
#include "common.cuh"

namespace gsb200 {

struct Sym3 { float xx, xy, xz, yy, yz, zz; };
struct Mat3 { float m[3][3]; };  // row-major

__device__ __forceinline__ Mat3 quat_to_rot(float w, float x, float y, float z, float *inv_norm_out) {
    float inv_norm = rsqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm; y *= inv_norm; z *= inv_norm; w *= inv_norm;
    if (inv_norm_out) *inv_norm_out = inv_norm;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    Mat3 R;
    R.m[0][0] = 1.f - 2.f * (y2 + z2); R.m[0][1] = 2.f * (xy - wz); R.m[0][2] = 2.f * (xz + wy);
    R.m[1][0] = 2.f * (xy + wz); R.m[1][1] = 1.f - 2.f * (x2 + z2); R.m[1][2] = 2.f * (yz - wx);
    R.m[2][0] = 2.f * (xz - wy); R.m[2][1] = 2.f * (yz + wx); R.m[2][2] = 1.f - 2.f * (x2 + y2);
    return R;
}

// Sigma = (Rq diag(s)) (Rq diag(s))^T
__device__ __forceinline__ Sym3 covar_from_quat_scale(const Mat3 &Rq, float sx, float sy, float sz) {
    float M[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { M[i][0] = Rq.m[i][0] * sx; M[i][1] = Rq.m[i][1] * sy; M[i][2] = Rq.m[i][2] * sz; }
    Sym3 S;
    S.xx = M[0][0] * M[0][0] + M[0][1] * M[0][1] + M[0][2] * M[0][2];
    S.xy = M[0][0] * M[1][0] + M[0][1] * M[1][1] + M[0][2] * M[1][2];
    S.xz = M[0][0] * M[2][0] + M[0][1] * M[2][1] + M[0][2] * M[2][2];
    S.yy = M[1][0] * M[1][0] + M[1][1] * M[1][1] + M[1][2] * M[1][2];
    S.yz = M[1][0] * M[2][0] + M[1][1] * M[2][1] + M[1][2] * M[2][2];
    S.zz = M[2][0] * M[2][0] + M[2][1] * M[2][1] + M[2][2] * M[2][2];
    return S;
}

// Sigma_c = R Sigma R^T, association (R Sigma) R^T as in transform.cuh:40-49
__device__ __forceinline__ Sym3 covar_to_cam(const float *R /*3x4 rows*/, const Sym3 &S) {
    float A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float r0 = R[i * 4 + 0], r1 = R[i * 4 + 1], r2 = R[i * 4 + 2];
        A[i][0] = r0 * S.xx + r1 * S.xy + r2 * S.xz;
        A[i][1] = r0 * S.xy + r1 * S.yy + r2 * S.yz;
        A[i][2] = r0 * S.xz + r1 * S.yz + r2 * S.zz;
    }
    Sym3 O;
    O.xx = A[0][0] * R[0] + A[0][1] * R[1] + A[0][2] * R[2];
    O.xy = A[0][0] * R[4] + A[0][1] * R[5] + A[0][2] * R[6];
    O.xz = A[0][0] * R[8] + A[0][1] * R[9] + A[0][2] * R[10];
    O.yy = A[1][0] * R[4] + A[1][1] * R[5] + A[1][2] * R[6];
    O.yz = A[1][0] * R[8] + A[1][1] * R[9] + A[1][2] * R[10];
    O.zz = A[2][0] * R[8] + A[2][1] * R[9] + A[2][2] * R[10];
    return O;
}

// Same product keeping all nine entries: (R Sigma) R^T is symmetric in exact arithmetic but NOT bit for bit -- entry (i, j)
// is sum_k A(i,k) R(j,k) and (j, i) is sum_k A(j,k) R(i,k).  The reference's glm products carry the full matrix into the
// projection (transform.cuh:40-49, type_mat3x3.inl:529-541), so the forward does too: conics then match the reference's
// in every bit, which keeps the alpha >= 1/255 / transmittance decisions of the blend identical from the RAW parameters.
__device__ __forceinline__ Mat3 covar_to_cam_full(const float *R /*3x4 rows*/, const Sym3 &S) {
    float A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float r0 = R[i * 4 + 0], r1 = R[i * 4 + 1], r2 = R[i * 4 + 2];
        A[i][0] = r0 * S.xx + r1 * S.xy + r2 * S.xz;
        A[i][1] = r0 * S.xy + r1 * S.yy + r2 * S.yz;
        A[i][2] = r0 * S.xz + r1 * S.yz + r2 * S.zz;
    }
    Mat3 O;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) O.m[i][j] = A[i][0] * R[j * 4 + 0] + A[i][1] * R[j * 4 + 1] + A[i][2] * R[j * 4 + 2];
    return O;
}

// Forward-mode dual number over (x, y, z): used to differentiate the fisheye projection exactly as written
// (the reference hand-expands the same derivatives, proj.cuh:244-362).
struct Dual3 {
    float v, d[3];
};
__device__ __forceinline__ Dual3 dconst(float c) { return {c, {0.f, 0.f, 0.f}}; }
__device__ __forceinline__ Dual3 dvar(float c, int k) { Dual3 r = dconst(c); r.d[k] = 1.f; return r; }
__device__ __forceinline__ Dual3 operator+(Dual3 a, Dual3 b) { return {a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
__device__ __forceinline__ Dual3 operator-(Dual3 a, Dual3 b) { return {a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
__device__ __forceinline__ Dual3 operator+(Dual3 a, float b) { a.v += b; return a; }
__device__ __forceinline__ Dual3 operator*(Dual3 a, Dual3 b) {
    return {a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
__device__ __forceinline__ Dual3 operator*(Dual3 a, float b) { return {a.v * b, {a.d[0] * b, a.d[1] * b, a.d[2] * b}}; }
__device__ __forceinline__ Dual3 operator/(Dual3 a, Dual3 b) {
    const float inv = 1.f / b.v, q = a.v * inv;
    return {q, {(a.d[0] - q * b.d[0]) * inv, (a.d[1] - q * b.d[1]) * inv, (a.d[2] - q * b.d[2]) * inv}};
}
__device__ __forceinline__ Dual3 dsqrt(Dual3 a) {
    const float r = sqrtf(a.v), h = 0.5f / r;
    return {r, {a.d[0] * h, a.d[1] * h, a.d[2] * h}};
}
__device__ __forceinline__ Dual3 datan2(Dual3 p, Dual3 q) {
    const float inv = 1.f / (p.v * p.v + q.v * q.v);
    return {atan2f(p.v, q.v), {(q.v * p.d[0] - p.v * q.d[0]) * inv, (q.v * p.d[1] - p.v * q.d[1]) * inv,
                                (q.v * p.d[2] - p.v * q.d[2]) * inv}};
}

// fisheye projection (proj.cuh:202-241) evaluated on duals: projected mean and the 2x3 matrix J
__device__ __forceinline__ void fisheye_dual(float x, float y, float z, float fx, float fy, float cx, float cy, Dual3 m[2],
                                             Dual3 J[2][3]) {
    const float eps = 0.0000001f;
    const Dual3 X = dvar(x, 0), Y = dvar(y, 1), Z = dvar(z, 2);
    const Dual3 xy_len = dsqrt(X * X + Y * Y) + eps;
    const Dual3 theta = datan2(xy_len, Z + eps);
    m[0] = X * fx * theta / xy_len + cx;
    m[1] = Y * fy * theta / xy_len + cy;
    const Dual3 x2 = X * X + eps, y2 = Y * Y, xy = X * Y;
    const Dual3 x2y2 = x2 + y2;
    const Dual3 inv = dconst(1.f) / (x2y2 + Z * Z);
    const Dual3 b = datan2(xy_len, Z) / xy_len / x2y2;
    const Dual3 a = Z * inv / x2y2;
    J[0][0] = (x2 * a + y2 * b) * fx; J[0][1] = xy * (a - b) * fx; J[0][2] = X * inv * (-fx);
    J[1][0] = xy * (a - b) * fy; J[1][1] = (y2 * a + x2 * b) * fy; J[1][2] = Y * inv * (-fy);
}

struct CamProj {
    float J[2][3];
    float mx, my;
    float tx, ty, rz;     // pinhole only
    bool clip_x, clip_y;  // pinhole only: x/z (y/z) was clamped to the fov limits
};

// Jacobian of the camera projection and projected mean (proj.cuh:9-36, 80-119)
__device__ __forceinline__ CamProj project_point(int model, float x, float y, float z, float fx, float fy,
                                                 float cx, float cy, float W, float H) {
    CamProj P;
    P.clip_x = P.clip_y = false;
    P.tx = P.ty = 0.f; P.rz = 0.f;
    if (model == GSB200_CAMERA_PINHOLE) {
        float tan_fovx = 0.5f * W / fx, tan_fovy = 0.5f * H / fy;
        float lim_x_pos = (W - cx) / fx + 0.3f * tan_fovx, lim_x_neg = cx / fx + 0.3f * tan_fovx;
        float lim_y_pos = (H - cy) / fy + 0.3f * tan_fovy, lim_y_neg = cy / fy + 0.3f * tan_fovy;
        float rz = 1.f / z, rz2 = rz * rz;
        float xz = x * rz, yz = y * rz;
        float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, xz));
        float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, yz));
        P.clip_x = !(xz <= lim_x_pos && xz >= -lim_x_neg);
        P.clip_y = !(yz <= lim_y_pos && yz >= -lim_y_neg);
        P.J[0][0] = fx * rz; P.J[0][1] = 0.f; P.J[0][2] = -fx * tx * rz2;
        P.J[1][0] = 0.f; P.J[1][1] = fy * rz; P.J[1][2] = -fy * ty * rz2;
        P.mx = fx * x * rz + cx; P.my = fy * y * rz + cy;
        P.tx = tx; P.ty = ty; P.rz = rz;
    } else if (model == GSB200_CAMERA_FISHEYE) {
        Dual3 m[2], J[2][3];
        fisheye_dual(x, y, z, fx, fy, cx, cy, m, J);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) P.J[r][j] = J[r][j].v;
        P.mx = m[0].v; P.my = m[1].v;
    } else {  // ORTHO
        P.J[0][0] = fx; P.J[0][1] = 0.f; P.J[0][2] = 0.f;
        P.J[1][0] = 0.f; P.J[1][1] = fy; P.J[1][2] = 0.f;
        P.mx = fx * x + cx; P.my = fy * y + cy;
    }
    return P;
}

// cov2d = (J Sigma_c) J^T -> (c00, c01, c11)
__device__ __forceinline__ void project_covar(const CamProj &P, const Sym3 &S, float &c00, float &c01, float &c11) {
    float B00 = P.J[0][0] * S.xx + P.J[0][1] * S.xy + P.J[0][2] * S.xz;
    float B01 = P.J[0][0] * S.xy + P.J[0][1] * S.yy + P.J[0][2] * S.yz;
    float B02 = P.J[0][0] * S.xz + P.J[0][1] * S.yz + P.J[0][2] * S.zz;
    float B10 = P.J[1][0] * S.xx + P.J[1][1] * S.xy + P.J[1][2] * S.xz;
    float B11 = P.J[1][0] * S.xy + P.J[1][1] * S.yy + P.J[1][2] * S.yz;
    float B12 = P.J[1][0] * S.xz + P.J[1][1] * S.yz + P.J[1][2] * S.zz;
    c00 = B00 * P.J[0][0] + B01 * P.J[0][1] + B02 * P.J[0][2];
    c01 = B00 * P.J[1][0] + B01 * P.J[1][1] + B02 * P.J[1][2];
    c11 = B10 * P.J[1][0] + B11 * P.J[1][1] + B12 * P.J[1][2];
}

// cov2d = (J Sigma_c) J^T with the full Sigma_c, all four entries (type_mat3x2.inl:451-486): c01 = (row 0, col 1),
// c10 = (row 1, col 0)
__device__ __forceinline__ void project_covar_full(const CamProj &P, const Mat3 &C, float &c00, float &c01, float &c10, float &c11) {
    float T[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[r][j] = P.J[r][0] * C.m[0][j] + P.J[r][1] * C.m[1][j] + P.J[r][2] * C.m[2][j];
    c00 = T[0][0] * P.J[0][0] + T[0][1] * P.J[0][1] + T[0][2] * P.J[0][2];
    c10 = T[1][0] * P.J[0][0] + T[1][1] * P.J[0][1] + T[1][2] * P.J[0][2];
    c01 = T[0][0] * P.J[1][0] + T[0][1] * P.J[1][1] + T[0][2] * P.J[1][2];
    c11 = T[1][0] * P.J[1][0] + T[1][1] * P.J[1][1] + T[1][2] * P.J[1][2];
}

}  // namespace gsb200
