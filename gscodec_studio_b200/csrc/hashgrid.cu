// Multi-resolution hash-grid encoder of the reference's alternative ("gaussian_model") in-loop entropy model
// (SURVEY.md section 8f.3).
//
// Replaces _gridencoder.grid_encode_forward / grid_encode_backward (third_party/gridencoder/src/gridencoder.cu:100-884,
// bindings.cpp) for the way gsplat/compression_simulation/gaussian_distribution_model.py:60-182 calls them
// (binary_vxl = None, PV = 0, all levels), with the STE_binary embedding transform (:19-38) fused in:
//   * positions x in [0,1]^D are scaled to x (R - 2) + 0.5 per level; the 2^D corners of the cell are looked up in the
//     level's table (dense row-major index while it fits, XOR-of-primes hash otherwise, gridencoder.cu:45-88);
//   * corners on the outer shell of the grid (a coordinate 0 or R-1) are dropped and the remaining weights
//     renormalised; inputs outside [0,1] give zeros;
//   * backward: embeddings receive w / wn (scatter-add), inputs receive the plain multilinear derivative with factor
//     (R - 2) (NOT the derivative of the renormalised forward -- the reference's dy_dx, gridencoder.cu:360-660).
// One thread per (point, level); outputs are written level-major [N, L*F] directly (the reference writes [L,N,F] and
// permutes); no dy_dx buffer: the backward recomputes the Jacobian instead of storing N*L*D*F floats.
#include "common.cuh"

namespace gsb200 {

constexpr int kGridMaxLevels = GSB200_GRID_MAX_LEVELS;

struct GridLevels {
    uint32_t offset[kGridMaxLevels + 1];  // rows, exclusive prefix
    uint32_t resolution[kGridMaxLevels];
    uint32_t n_levels;
};

template <int D>
__device__ __forceinline__ uint32_t grid_row(const uint32_t (&p)[D], uint32_t table_rows, uint32_t resolution) {
    constexpr uint32_t primes[3] = {1u, 2654435761u, 805459861u};
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (stride <= table_rows) {
            index += p[d] * stride;
            stride *= resolution;
        }
    }
    if (stride > table_rows) {
        index = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) index ^= p[d] * primes[d];
    }
    return index % table_rows;
}

template <int D>
__device__ __forceinline__ bool on_shell(const uint32_t (&p)[D], uint32_t resolution) {
    bool s = false;
#pragma unroll
    for (int d = 0; d < D; ++d) s |= (p[d] == 0u) || (p[d] == resolution - 1u);
    return s;
}

// embedding value as the encoder sees it: sign with +1 at zero when the binary straight-through transform is on
__device__ __forceinline__ float embed_value(float p, bool ste_binary) { return ste_binary ? (p >= 0.f ? 1.f : -1.f) : p; }

template <int D, int F>
__global__ void __launch_bounds__(kThreads)
grid_encode_fwd_kernel(uint32_t N, GridLevels L, bool ste_binary, const float *__restrict__ inputs, const float *__restrict__ params,
                       float *__restrict__ outputs) {
    const uint32_t n = blockIdx.x * kThreads + threadIdx.x, lvl = blockIdx.y;
    if (n >= N) return;
    const uint32_t R = L.resolution[lvl], rows = L.offset[lvl + 1] - L.offset[lvl];
    const float *table = params + (size_t)L.offset[lvl] * F;
    float *out = outputs + (size_t)n * L.n_levels * F + lvl * F;
    float x[D];
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        x[d] = inputs[(size_t)n * D + d];
        oob |= (x[d] < 0.f) || (x[d] > 1.f);
    }
    float res[F];
#pragma unroll
    for (int f = 0; f < F; ++f) res[f] = 0.f;
    if (!oob) {
        float frac[D];
        uint32_t cell[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float p = __fadd_rn(__fmul_rn(x[d], (float)(R - 2u)), 0.5f);  // two roundings, as the reference (gridencoder.cu:183): no FMA
            const float fl = floorf(p);
            cell[d] = (uint32_t)fl;
            frac[d] = p - fl;
        }
        float wn = 0.f;
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
        for (int c = 0; c < (1 << D); ++c) {
            float w = 1.f;
            uint32_t pg[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const bool hi = (c >> d) & 1;
                w *= hi ? frac[d] : 1.f - frac[d];
                pg[d] = hi ? min(cell[d] + 1u, R - 1u) : cell[d];
            }
            if (!on_shell<D>(pg, R)) {
                const uint32_t row = grid_row<D>(pg, rows, R);
                wn += w;
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] += w * embed_value(__ldg(table + (size_t)row * F + f), ste_binary);
            }
        }
        if (wn == 0.f) wn += 1e-9f;
        const float wn_re = 1.0f / wn;
#pragma unroll
        for (int f = 0; f < F; ++f) res[f] = acc[f] * wn_re;
    }
#pragma unroll
    for (int f = 0; f < F; ++f) out[f] = res[f];
}

template <int D, int F>
__global__ void __launch_bounds__(kThreads)
grid_encode_bwd_kernel(uint32_t N, GridLevels L, bool ste_binary, const float *__restrict__ inputs, const float *__restrict__ params,
                       const float *__restrict__ v_outputs, float *__restrict__ v_params, float *__restrict__ v_inputs) {
    const uint32_t n = blockIdx.x * kThreads + threadIdx.x, lvl = blockIdx.y;
    if (n >= N) return;
    const uint32_t R = L.resolution[lvl], rows = L.offset[lvl + 1] - L.offset[lvl];
    const float *table = params + (size_t)L.offset[lvl] * F;
    float *v_table = v_params + (size_t)L.offset[lvl] * F;
    float g[F];
#pragma unroll
    for (int f = 0; f < F; ++f) g[f] = v_outputs[(size_t)n * L.n_levels * F + lvl * F + f];
    float x[D];
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        x[d] = inputs[(size_t)n * D + d];
        oob |= (x[d] < 0.f) || (x[d] > 1.f);
    }
    if (oob) return;  // zero output, zero Jacobian
    float frac[D];
    uint32_t cell[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float p = __fadd_rn(__fmul_rn(x[d], (float)(R - 2u)), 0.5f);  // two roundings, as the reference (gridencoder.cu:183): no FMA
        const float fl = floorf(p);
        cell[d] = (uint32_t)fl;
        frac[d] = p - fl;
    }
    // embedding gradient: w / wn to every kept corner (through the straight-through gate |p| <= 1 when binarised)
    float wn = 0.f;
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
        float w = 1.f;
        uint32_t pg[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const bool hi = (c >> d) & 1;
            w *= hi ? frac[d] : 1.f - frac[d];
            pg[d] = hi ? min(cell[d] + 1u, R - 1u) : cell[d];
        }
        if (!on_shell<D>(pg, R)) wn += w;
    }
    if (wn == 0.f) wn += 1e-9f;
    const float wn_re = 1.0f / wn;
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
        float w = 1.f;
        uint32_t pg[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const bool hi = (c >> d) & 1;
            w *= hi ? frac[d] : 1.f - frac[d];
            pg[d] = hi ? min(cell[d] + 1u, R - 1u) : cell[d];
        }
        if (!on_shell<D>(pg, R)) {
            const uint32_t row = grid_row<D>(pg, rows, R);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float p = __ldg(table + (size_t)row * F + f);
                if (!ste_binary || (p >= -1.f && p <= 1.f)) red_add(v_table + (size_t)row * F + f, w * wn_re * g[f]);
            }
        }
    }
    // input gradient: multilinear derivative along every axis, dropped corners read as zero
    if (v_inputs != nullptr) {
#pragma unroll
        for (int gd = 0; gd < D; ++gd) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < (1 << (D - 1)); ++c) {
                float w = (float)(R - 2u);
                uint32_t pg[D];
                int bit = 0;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if (d == gd) continue;
                    const bool hi = (c >> bit) & 1;
                    ++bit;
                    w *= hi ? frac[d] : 1.f - frac[d];
                    pg[d] = hi ? min(cell[d] + 1u, R - 1u) : cell[d];
                }
                float diff[F];
#pragma unroll
                for (int f = 0; f < F; ++f) diff[f] = 0.f;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    pg[gd] = side ? min(cell[gd] + 1u, R - 1u) : cell[gd];
                    if (!on_shell<D>(pg, R)) {
                        const uint32_t row = grid_row<D>(pg, rows, R);
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float v = embed_value(__ldg(table + (size_t)row * F + f), ste_binary);
                            diff[f] += side ? v : -v;
                        }
                    }
                }
#pragma unroll
                for (int f = 0; f < F; ++f) acc += w * diff[f] * g[f];
            }
            red_add(v_inputs + (size_t)n * D + gd, acc);
        }
    }
}

static int make_levels(const char *who, uint32_t n_levels, const int32_t *offsets, const int32_t *resolutions, GridLevels &L) {
    if (n_levels == 0 || n_levels > (uint32_t)kGridMaxLevels) return failf(2, "%s: n_levels must be in [1, %d]", who, kGridMaxLevels);
    if (offsets == nullptr || resolutions == nullptr) return failf(2, "%s: null level tables", who);
    for (uint32_t l = 0; l <= n_levels; ++l) {
        if (offsets[l] < 0 || (l > 0 && offsets[l] <= offsets[l - 1])) return failf(2, "%s: offsets must increase", who);
        L.offset[l] = (uint32_t)offsets[l];
    }
    for (uint32_t l = 0; l < n_levels; ++l) {
        if (resolutions[l] < 3) return failf(2, "%s: resolutions must be >= 3", who);
        L.resolution[l] = (uint32_t)resolutions[l];
    }
    L.n_levels = n_levels;
    return 0;
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_grid_encode_fwd(uint32_t N, uint32_t num_dim, uint32_t n_features, uint32_t n_levels, const int32_t *offsets,
                                      const int32_t *resolutions, int ste_binary, const float *inputs, const float *params,
                                      float *outputs, gsb200_stream_t stream) {
    const char *who = "gsb200_grid_encode_fwd";
    GridLevels L{};
    if (int rc = make_levels(who, n_levels, offsets, resolutions, L)) return rc;
    if (n_features != 2 || (num_dim != 2 && num_dim != 3)) return failf(2, "%s: only 2 features per level and 2 or 3 input dimensions", who);
    if (N == 0) return 0;
    if (inputs == nullptr || params == nullptr || outputs == nullptr) return failf(2, "%s: null pointer", who);
    const dim3 grid(cdiv(N, kThreads), n_levels);
    cudaStream_t s = (cudaStream_t)stream;
    if (num_dim == 3) grid_encode_fwd_kernel<3, 2><<<grid, kThreads, 0, s>>>(N, L, ste_binary != 0, inputs, params, outputs);
    else grid_encode_fwd_kernel<2, 2><<<grid, kThreads, 0, s>>>(N, L, ste_binary != 0, inputs, params, outputs);
    return check_launch("grid_encode_fwd");
}

extern "C" int gsb200_grid_encode_bwd(uint32_t N, uint32_t num_dim, uint32_t n_features, uint32_t n_levels, const int32_t *offsets,
                                      const int32_t *resolutions, int ste_binary, const float *inputs, const float *params,
                                      const float *v_outputs, float *v_params, float *v_inputs, gsb200_stream_t stream) {
    const char *who = "gsb200_grid_encode_bwd";
    GridLevels L{};
    if (int rc = make_levels(who, n_levels, offsets, resolutions, L)) return rc;
    if (n_features != 2 || (num_dim != 2 && num_dim != 3)) return failf(2, "%s: only 2 features per level and 2 or 3 input dimensions", who);
    if (v_params == nullptr) return failf(2, "%s: null v_params", who);
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_OK(cudaMemsetAsync(v_params, 0, sizeof(float) * (size_t)L.offset[n_levels] * n_features, s));
    if (v_inputs != nullptr && N > 0) GSB_CUDA_OK(cudaMemsetAsync(v_inputs, 0, sizeof(float) * (size_t)N * num_dim, s));
    if (N == 0) return 0;
    if (inputs == nullptr || params == nullptr || v_outputs == nullptr) return failf(2, "%s: null pointer", who);
    const dim3 grid(cdiv(N, kThreads), n_levels);
    if (num_dim == 3) grid_encode_bwd_kernel<3, 2><<<grid, kThreads, 0, s>>>(N, L, ste_binary != 0, inputs, params, v_outputs, v_params, v_inputs);
    else grid_encode_bwd_kernel<2, 2><<<grid, kThreads, 0, s>>>(N, L, ste_binary != 0, inputs, params, v_outputs, v_params, v_inputs);
    return check_launch("grid_encode_bwd");
}
