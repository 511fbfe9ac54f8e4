// In-loop compression simulation: fused fake-quantisation (+ activation) and the factorized-prior
// entropy model, forward and backward (sm_100a).  HBM-bound elementwise kernels.
//
// fake quantisation replaces the torch op chain of gsplat/compression_simulation/ops.py:39-75
// (clamp, uniform_ noise, mul, add -- 4-6 kernels per attribute, each a full read+write) with one
// pass that can also apply the activation the trainer applies next (exp for scales, sigmoid for
// opacities; examples/simple_trainer.py:758-776).
//
// The entropy model replaces Entropy_factorized_optimized_refactor.forward
// (gsplat/compression_simulation/entropy_model.py:201-254): 3-4 bmm + ~12 elementwise kernels per
// attribute become one kernel that evaluates the per-channel monotone MLP CDF in registers.  The
// reference's reshape [2ch,1,Npad] -> [64ch,1,Npad/32] with weights tiled by repeat(64,1,1) means
// element (n, c) is evaluated with parameter set p = (32*c + n / (Npad/32)) mod ch for BOTH the
// lower and the upper logit; blocks are laid out so p is block-uniform.
#include "common.cuh"
#include "quant_load.cuh"  // Philox4x32-10 and u01_centered, shared with the quantise-on-load kernels

namespace gsb200 {

// the activations the trainer applies next (examples/simple_trainer.py:758-776), with torch.exp / torch.sigmoid's exact
// arithmetic -- libdevice expf and an IEEE division; this file is built without --use_fast_math (see _build.py)
__device__ __forceinline__ float activate(int act, float v) {
    if (act == 1) return expf(v);
    if (act == 2) return 1.f / (1.f + expf(-v));
    return v;
}

// 4 consecutive elements starting at `base` (one Philox call yields 4 uniforms)
__device__ __forceinline__ void fake_quantize_fwd4(int64_t base, int64_t n, float *__restrict__ x, float lo, float hi, float q_step,
                                                   float qs_norm, int mode, int act, const float *__restrict__ noise, uint64_t seed,
                                                   uint64_t offset, float *__restrict__ y) {
    if (base >= n) return;
    float u[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == 0) {
        if (noise != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = base + k < n ? noise[base + k] : 0.f;
        } else {
            const uint64_t c = offset + (uint64_t)(base >> 2);
            const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            u[0] = u01_centered(r.x); u[1] = u01_centered(r.y); u[2] = u01_centered(r.z); u[3] = u01_centered(r.w);
        }
    }
    // whole, 16-byte aligned quads move as one vector each way; the ragged tail and unaligned tensors element-wise
    const bool vec = base + 4 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    float in[4] = {0.f, 0.f, 0.f, 0.f}, cl4[4], out4[4];
    if (vec) {
        const float4 v4 = *reinterpret_cast<const float4 *>(x + base);
        in[0] = v4.x; in[1] = v4.y; in[2] = v4.z; in[3] = v4.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (base + k < n) in[k] = x[base + k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float cl = fminf(fmaxf(in[k], lo), hi);
        float out;
        if (mode == 0) {
            out = __fadd_rn(cl, __fmul_rn(u[k], q_step));  // ops.py:44-47 (two roundings, like the torch ops)
        } else {
            const float nrm = __fdiv_rn(cl - lo, hi - lo);
            out = __fadd_rn(__fmul_rn(__fmul_rn(rintf(__fdiv_rn(nrm, qs_norm)), qs_norm), hi - lo), lo);  // ops.py:65-69
        }
        cl4[k] = cl;
        out4[k] = activate(act, out);
    }
    if (vec) {
        if (mode != 0) *reinterpret_cast<float4 *>(x + base) = make_float4(cl4[0], cl4[1], cl4[2], cl4[3]);  // ops.py:63
        *reinterpret_cast<float4 *>(y + base) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (base + k >= n) break;
            if (mode != 0) x[base + k] = cl4[k];  // in-place clamp of the parameter, ops.py:63
            y[base + k] = out4[k];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
fake_quantize_fwd_kernel(int64_t n, float *__restrict__ x, float lo, float hi, float q_step, float qs_norm, int mode, int act,
                         const float *__restrict__ noise, uint64_t seed, uint64_t offset, float *__restrict__ y) {
    fake_quantize_fwd4(((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4, n, x, lo, hi, q_step, qs_norm, mode, act, noise, seed,
                       offset, y);
}

__device__ __forceinline__ float fake_quantize_bwd_elem(float g, float xv, float yv, float lo, float hi, int mode, int act) {
    if (act == 1) g *= yv;
    else if (act == 2) g *= yv * (1.f - yv);
    if (mode == 0) g = (xv >= lo && xv <= hi) ? g : 0.f;  // clamp gate; round mode is a pure STE
    return g;
}

// 4 consecutive elements starting at `base`: one 16-byte vector per array when the quad is whole and aligned
__device__ __forceinline__ void fake_quantize_bwd4(int64_t base, int64_t n, const float *__restrict__ x, const float *__restrict__ y,
                                                   float lo, float hi, int mode, int act, const float *__restrict__ v_y,
                                                   float *__restrict__ v_x) {
    if (base >= n) return;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(v_y) |
                           reinterpret_cast<uintptr_t>(v_x);
    if (base + 4 <= n && (bits & 15) == 0) {
        const float4 g = *reinterpret_cast<const float4 *>(v_y + base);
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv;
        if (mode == 0) xv = *reinterpret_cast<const float4 *>(x + base);
        if (act != 0) yv = *reinterpret_cast<const float4 *>(y + base);
        *reinterpret_cast<float4 *>(v_x + base) =
            make_float4(fake_quantize_bwd_elem(g.x, xv.x, yv.x, lo, hi, mode, act), fake_quantize_bwd_elem(g.y, xv.y, yv.y, lo, hi, mode, act),
                        fake_quantize_bwd_elem(g.z, xv.z, yv.z, lo, hi, mode, act), fake_quantize_bwd_elem(g.w, xv.w, yv.w, lo, hi, mode, act));
        return;
    }
    for (int k = 0; k < 4 && base + k < n; ++k) {
        const int64_t i = base + k;
        v_x[i] = fake_quantize_bwd_elem(v_y[i], mode == 0 ? x[i] : 0.f, act != 0 ? y[i] : 0.f, lo, hi, mode, act);
    }
}

__global__ void __launch_bounds__(kThreads)
fake_quantize_bwd_kernel(int64_t n, const float *__restrict__ x, const float *__restrict__ y, float lo, float hi, int mode,
                         int act, const float *__restrict__ v_y, float *__restrict__ v_x) {
    fake_quantize_bwd4(((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4, n, x, y, lo, hi, mode, act, v_y, v_x);
}

// ---- all attributes of a step in ONE launch per direction (the policy quantises 4-7 tensors per step; one launch each
// costs more host time than device time at 1 M Gaussians) -- same per-element arithmetic as the single-tensor kernels
struct FqTable {
    gsb200_fq_tensor t[GSB200_FQ_MAX_TENSORS];
    float q_step[GSB200_FQ_MAX_TENSORS], qs_norm[GSB200_FQ_MAX_TENSORS];
    uint32_t block_end[GSB200_FQ_MAX_TENSORS];  // exclusive prefix of blocks, per tensor
    uint32_t n;
};

__global__ void __launch_bounds__(kThreads) fake_quantize_multi_fwd_kernel(const __grid_constant__ FqTable T) {
    uint32_t ti = 0;
    while (ti + 1 < T.n && blockIdx.x >= T.block_end[ti]) ++ti;
    const gsb200_fq_tensor &h = T.t[ti];
    const uint32_t blk = blockIdx.x - (ti == 0 ? 0u : T.block_end[ti - 1]);
    fake_quantize_fwd4(((int64_t)blk * kThreads + threadIdx.x) * 4, h.n, h.x, h.lo, h.hi, T.q_step[ti], T.qs_norm[ti], h.mode, h.act,
                       h.noise, h.seed, h.offset, h.y);
}

__global__ void __launch_bounds__(kThreads) fake_quantize_multi_bwd_kernel(const __grid_constant__ FqTable T) {
    uint32_t ti = 0;
    while (ti + 1 < T.n && blockIdx.x >= T.block_end[ti]) ++ti;
    const gsb200_fq_tensor &h = T.t[ti];
    const uint32_t blk = blockIdx.x - (ti == 0 ? 0u : T.block_end[ti - 1]);
    fake_quantize_bwd4(((int64_t)blk * kThreads + threadIdx.x) * 4, h.n, h.x, h.y, h.lo, h.hi, h.mode, h.act, h.v_y, h.v_x);
}

// ---------------------------------------------------------------- factorized entropy model
// activated parameter set layout (floats), NH hidden layers of width 3:
//   layer 0      : W[3] (3x1), b[3], tf[3]
//   layer 1..NH-1: W[9] (3x3 row-major [out][in]), b[3], tf[3]
//   last         : W[3] (1x3), b[1]
template <int NH> struct EntropyDims { static constexpr int PS = 9 + (NH - 1) * 15 + 4; };

// (this file is built without --use_fast_math for the torch-exact activations above; the entropy model keeps the fast
// intrinsics it was validated with -- an IEEE division here is 8 extra instructions, 22 times per element)
__device__ __forceinline__ float fast_tanh(float v) {
    const float e = __expf(2.f * v);
    return 1.f - __fdividef(2.f, e + 1.f);  // tanh(v); saturates correctly for |v| large (2 / inf = 0)
}
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

template <int NH>
__device__ __forceinline__ float entropy_logit(const float *__restrict__ P, float u, float (*h_pre)[3], float (*h_tanh)[3]) {
    // returns the logit; optionally records per-layer pre-activations and tanh values for the VJP
    float h[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const float z = P[o] * u + P[3 + o];
        const float t = fast_tanh(z);
        if (h_pre) { h_pre[0][o] = z; h_tanh[0][o] = t; }
        h[o] = z + P[6 + o] * t;
    }
#pragma unroll
    for (int l = 1; l < NH; ++l) {
        const float *L = P + 9 + (l - 1) * 15;
        float hn[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float z = L[o * 3] * h[0] + L[o * 3 + 1] * h[1] + L[o * 3 + 2] * h[2] + L[9 + o];
            const float t = fast_tanh(z);
            if (h_pre) { h_pre[l][o] = z; h_tanh[l][o] = t; }
            hn[o] = z + L[12 + o] * t;
        }
        h[0] = hn[0]; h[1] = hn[1]; h[2] = hn[2];
    }
    const float *L = P + 9 + (NH - 1) * 15;
    return L[0] * h[0] + L[1] * h[1] + L[2] * h[2] + L[3];
}

// VJP of entropy_logit w.r.t. u and the activated parameters; vP accumulates (+=)
template <int NH>
__device__ __forceinline__ float entropy_logit_vjp(const float *__restrict__ P, float u, float v_logit, float *vP) {
    float h_pre[NH][3], h_tanh[NH][3];
    (void)entropy_logit<NH>(P, u, h_pre, h_tanh);
    // layer outputs h_l = z + tf * tanh(z)
    const float *LL = P + 9 + (NH - 1) * 15;
    float *vLL = vP + 9 + (NH - 1) * 15;
    float vh[3];
    {
        float hl[3];
        const float *tfp = NH == 1 ? P + 6 : P + 9 + (NH - 2) * 15 + 12;
#pragma unroll
        for (int o = 0; o < 3; ++o) hl[o] = h_pre[NH - 1][o] + tfp[o] * h_tanh[NH - 1][o];
#pragma unroll
        for (int o = 0; o < 3; ++o) { vLL[o] += v_logit * hl[o]; vh[o] = v_logit * LL[o]; }
        vLL[3] += v_logit;
    }
#pragma unroll
    for (int l = NH - 1; l >= 1; --l) {
        const float *L = P + 9 + (l - 1) * 15;
        float *vL = vP + 9 + (l - 1) * 15;
        // input to layer l = output of layer l-1
        float hin[3];
        const float *tfp = l == 1 ? P + 6 : P + 9 + (l - 2) * 15 + 12;
#pragma unroll
        for (int o = 0; o < 3; ++o) hin[o] = h_pre[l - 1][o] + tfp[o] * h_tanh[l - 1][o];
        float vin[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float t = h_tanh[l][o];
            vL[12 + o] += vh[o] * t;                              // d/d tf
            const float vz = vh[o] * (1.f + L[12 + o] * (1.f - t * t));
            vL[9 + o] += vz;                                      // bias
#pragma unroll
            for (int k = 0; k < 3; ++k) { vL[o * 3 + k] += vz * hin[k]; vin[k] += vz * L[o * 3 + k]; }
        }
        vh[0] = vin[0]; vh[1] = vin[1]; vh[2] = vin[2];
    }
    float vu = 0.f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const float t = h_tanh[0][o];
        vP[6 + o] += vh[o] * t;
        const float vz = vh[o] * (1.f + P[6 + o] * (1.f - t * t));
        vP[3 + o] += vz;
        vP[o] += vz * u;
        vu += vz * P[o];
    }
    return vu;
}

// ---- activated parameter table of the entropy model in ONE launch each way (softplus of the matrices, tanh of the factors,
// biases as they are: entropy_model.py:216-232; as torch ops that is ~10 tiny kernels forward and ~25 backward per model, three
// models per step -- 1 ms of host time).  One thread per (channel, table column).
struct EntropyRaw {
    const float *matrices[4], *bias[4], *factor[3];
    float *v_matrices[4], *v_bias[4], *v_factor[3];
};

// column of the table row -> (layer, kind 0 matrix / 1 bias / 2 factor, element offset inside the channel's block, block size)
__device__ __forceinline__ void entropy_column(uint32_t NH, uint32_t col, uint32_t &layer, uint32_t &kind, uint32_t &idx, uint32_t &size) {
    const uint32_t mid_end = 9 + (NH - 1) * 15;
    if (col < 9) {
        layer = 0; kind = col / 3; idx = col % 3; size = 3;
    } else if (col < mid_end) {
        layer = 1 + (col - 9) / 15;
        const uint32_t r = (col - 9) % 15;
        if (r < 9) { kind = 0; idx = r; size = 9; }
        else if (r < 12) { kind = 1; idx = r - 9; size = 3; }
        else { kind = 2; idx = r - 12; size = 3; }
    } else {
        layer = NH;
        const uint32_t r = col - mid_end;
        if (r < 3) { kind = 0; idx = r; size = 3; }
        else { kind = 1; idx = 0; size = 1; }
    }
}

template <bool BWD>
__global__ void entropy_params_kernel(uint32_t ch, uint32_t NH, uint32_t PS, const __grid_constant__ EntropyRaw R, float *__restrict__ params,
                                      const float *__restrict__ v_params) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ch * PS) return;
    const uint32_t c = t / PS, col = t % PS;
    uint32_t layer, kind, idx, size;
    entropy_column(NH, col, layer, kind, idx, size);
    const size_t at = (size_t)c * size + idx;
    if (!BWD) {
        float v;
        if (kind == 0) { const float x = R.matrices[layer][at]; v = x > 20.f ? x : log1pf(expf(x)); }  // F.softplus
        else if (kind == 1) v = R.bias[layer][at];
        else v = tanhf(R.factor[layer][at]);
        params[t] = v;
    } else {
        const float g = v_params[t];
        if (kind == 0) {
            const float x = R.matrices[layer][at];
            if (R.v_matrices[layer]) R.v_matrices[layer][at] = x > 20.f ? g : g / (1.f + expf(-x));
        } else if (kind == 1) {
            if (R.v_bias[layer]) R.v_bias[layer][at] = g;
        } else if (R.v_factor[layer]) {
            const float th = tanhf(R.factor[layer][at]);
            R.v_factor[layer][at] = g * (1.f - th * th);
        }
    }
}

constexpr int kEntropyThreads = 128;

// grid: (blocks per segment, 32 chunks, ch); element n = j*L + local, local grid-strided inside the segment
template <int NH, bool BWD>
__global__ void __launch_bounds__(kEntropyThreads)
entropy_kernel(int64_t N, uint32_t ch, int64_t L, const float *__restrict__ x, float q, const float *__restrict__ params,
               float bound, float *__restrict__ bits, const float *__restrict__ v_bits, float *__restrict__ v_x,
               float *__restrict__ v_params) {
    constexpr int PS = EntropyDims<NH>::PS;
    __shared__ float sP[PS];
    __shared__ float sV[PS];
    const uint32_t j = blockIdx.y, c = blockIdx.z;
    const uint32_t p = (32u * c + j) % ch;
    for (int k = threadIdx.x; k < PS; k += blockDim.x) { sP[k] = params[(size_t)p * PS + k]; if (BWD) sV[k] = 0.f; }
    __syncthreads();
    float vP[PS];
    if (BWD) {
#pragma unroll
        for (int k = 0; k < PS; ++k) vP[k] = 0.f;
    }
    // forward: a thread evaluates the same tiny network for all its elements -- its parameters live in registers (the
    // backward already holds PS gradient accumulators there and reads the parameters from shared memory)
    float rP[BWD ? 1 : PS];
    if (!BWD) {
#pragma unroll
        for (int k = 0; k < PS; ++k) rP[k] = sP[k];
    }
    const float *P = BWD ? sP : rP;
    // grid-stride over the segment: a thread owns many elements, so the parameter-gradient reduction
    // below runs once per thread instead of once per element
    for (int64_t local = (int64_t)blockIdx.x * kEntropyThreads + threadIdx.x; local < L;
         local += (int64_t)gridDim.x * kEntropyThreads) {
        const int64_t n = (int64_t)j * L + local;
        if (n >= N) break;
        const float xv = x[n * ch + c];
        const float ul = xv - 0.5f * q, uu = xv + 0.5f * q;
        const float lower = entropy_logit<NH>(P, ul, nullptr, nullptr);
        const float upper = entropy_logit<NH>(P, uu, nullptr, nullptr);
        const float s = lower + upper;
        const float sign = s > 0.f ? -1.f : (s < 0.f ? 1.f : 0.f);
        const float su = sigmoidf_(sign * upper), sl = sigmoidf_(sign * lower);
        const float diff = su - sl;
        const float lik_raw = fabsf(diff);
        const float lik = fmaxf(lik_raw, bound);
        if (!BWD) {
            bits[n * ch + c] = -__log2f(lik);
        } else {
            const float vb = v_bits[n * ch + c];
            float v_lik = -vb / (lik * 0.6931471805599453f);
            if (!(lik_raw >= bound || v_lik < 0.f)) v_lik = 0.f;  // LowerBound gradient rule, entropy_model.py:351-353
            const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            const float v_upper = v_lik * sd * su * (1.f - su) * sign;
            const float v_lower = -v_lik * sd * sl * (1.f - sl) * sign;
            float vx = entropy_logit_vjp<NH>(sP, uu, v_upper, vP);
            vx += entropy_logit_vjp<NH>(sP, ul, v_lower, vP);
            v_x[n * ch + c] = vx;
        }
    }
    if (BWD) {
        // block reduction of the PS parameter gradients: warp butterfly, then shared, then one atomic each
#pragma unroll
        for (int k = 0; k < PS; ++k) {
            const float t = warp_sum(vP[k]);
            if ((threadIdx.x & 31) == 0 && t != 0.f) atomicAdd(&sV[k], t);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < PS; k += blockDim.x)
            if (sV[k] != 0.f) atomicAdd(v_params + (size_t)p * PS + k, sV[k]);
    }
}

// ---------------------------------------------------------------- on-disk quantise-to-grid (PngCompression)
// gsplat/compression/png_compression.py:166-360: every attribute is normalised per channel to [0, 1] over the whole grid,
// rounded to `bits` levels (numpy's round-half-even on the float32 product) and stored as 8-bit image planes -- k-bit values
// left-aligned in the byte (`img << (8 - k)`), 16-bit values as a low and a high plane.  The reference does this with five
// full-size torch / numpy passes on the host; here it is one pass each way on the device (this file is built without
// --use_fast_math: the division and the products are the IEEE ones numpy performs).
__global__ void __launch_bounds__(kThreads)
grid_quantize_kernel(int64_t n_elems, uint32_t ch, const float *__restrict__ x, const float *__restrict__ mins,
                     const float *__restrict__ maxs, int bits, uint8_t *__restrict__ lo, uint8_t *__restrict__ hi) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_elems) return;
    const uint32_t c = (uint32_t)(i % ch);
    const float mn = mins[c], mx = maxs[c];
    const float norm = __fdiv_rn(__fsub_rn(x[i], mn), __fsub_rn(mx, mn));  // (grid - mins) / (maxs - mins)
    const float levels = (float)((1u << bits) - 1u);
    const float q = rintf(__fmul_rn(norm, levels));  // np.round: half to even
    // astype(np.uint8 / np.uint16) of an out-of-range or NaN value (maxs == mins) is unspecified in numpy; saturate
    const uint32_t v = q >= 0.f ? (q <= levels ? (uint32_t)q : (uint32_t)levels) : 0u;
    if (bits <= 8) {
        lo[i] = (uint8_t)(v << (8 - bits));
    } else {
        lo[i] = (uint8_t)(v & 0xFFu);
        hi[i] = (uint8_t)(v >> 8);
    }
}

__global__ void __launch_bounds__(kThreads)
grid_dequantize_kernel(int64_t n_elems, uint32_t ch, const uint8_t *__restrict__ lo, const uint8_t *__restrict__ hi,
                       const float *__restrict__ mins, const float *__restrict__ maxs, int bits, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_elems) return;
    const uint32_t c = (uint32_t)(i % ch);
    const uint32_t v = bits <= 8 ? ((uint32_t)lo[i] >> (8 - bits)) : (((uint32_t)hi[i] << 8) + (uint32_t)lo[i]);
    // the reference decodes in float64 (numpy `img / (2**k - 1)`, then a float64 * float32 torch product) and casts at the end
    const double norm = (double)v / (double)((1u << bits) - 1u);
    const float span = __fsub_rn(maxs[c], mins[c]);
    out[i] = (float)(norm * (double)span + (double)mins[c]);
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_grid_quantize(int64_t n_elems, uint32_t ch, const float *x, const float *mins, const float *maxs, int bits,
                                    uint8_t *lo, uint8_t *hi, gsb200_stream_t stream) {
    const char *who = "gsb200_grid_quantize";
    if (bits < 1 || bits > 16) return failf(2, "%s: bits must be in [1, 16]", who);
    if (ch == 0) return failf(2, "%s: zero channels", who);
    if (n_elems == 0) return 0;
    if (!x || !mins || !maxs || !lo || (bits > 8 && !hi)) return failf(2, "%s: null pointer", who);
    grid_quantize_kernel<<<(unsigned)((n_elems + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(n_elems, ch, x, mins,
                                                                                                                maxs, bits, lo, hi);
    return check_launch("grid_quantize");
}

extern "C" int gsb200_grid_dequantize(int64_t n_elems, uint32_t ch, const uint8_t *lo, const uint8_t *hi, const float *mins,
                                      const float *maxs, int bits, float *out, gsb200_stream_t stream) {
    const char *who = "gsb200_grid_dequantize";
    if (bits < 1 || bits > 16) return failf(2, "%s: bits must be in [1, 16]", who);
    if (ch == 0) return failf(2, "%s: zero channels", who);
    if (n_elems == 0) return 0;
    if (!out || !mins || !maxs || !lo || (bits > 8 && !hi)) return failf(2, "%s: null pointer", who);
    grid_dequantize_kernel<<<(unsigned)((n_elems + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(n_elems, ch, lo, hi,
                                                                                                                  mins, maxs, bits, out);
    return check_launch("grid_dequantize");
}


extern "C" int gsb200_fake_quantize_fwd(int64_t n, float *x, float lo, float hi, int bitwidth, int mode, int act,
                                        const float *noise, uint64_t seed, uint64_t offset, float *y, gsb200_stream_t stream) {
    if (mode != 0 && mode != 1) return fail(2, "gsb200_fake_quantize_fwd: mode must be 0 (noise) or 1 (round)");
    if (act < 0 || act > 2) return fail(2, "gsb200_fake_quantize_fwd: act must be 0, 1 or 2");
    if (n == 0) return 0;
    const double levels = (double)((1ull << bitwidth) - 1ull);
    const float q_step = (float)(((double)hi - (double)lo) / levels);
    const float qs_norm = (float)(1.0 / levels);
    const int64_t nthreads = (n + 3) / 4;
    fake_quantize_fwd_kernel<<<(unsigned)((nthreads + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n, x, lo, hi, q_step, qs_norm, mode, act, noise, seed, offset, y);
    return check_launch("fake_quantize_fwd");
}

static int fq_table(const char *who, uint32_t n_tensors, const gsb200_fq_tensor *tensors, bool backward, FqTable &T, uint64_t &blocks) {
    if (n_tensors > GSB200_FQ_MAX_TENSORS) return failf(2, "%s: at most %d tensors per call", who, GSB200_FQ_MAX_TENSORS);
    if (n_tensors > 0 && tensors == nullptr) return failf(2, "%s: null table", who);
    blocks = 0;
    for (uint32_t i = 0; i < n_tensors; ++i) {
        const gsb200_fq_tensor &h = tensors[i];
        if (h.n <= 0) continue;
        if (h.mode != 0 && h.mode != 1) return failf(2, "%s: mode must be 0 (noise) or 1 (round)", who);
        if (h.act < 0 || h.act > 2) return failf(2, "%s: act must be 0, 1 or 2", who);
        if (h.x == nullptr || h.y == nullptr || (backward && (h.v_y == nullptr || h.v_x == nullptr)))
            return failf(2, "%s: null pointer in the tensor table", who);
        const double levels = (double)((1ull << h.bitwidth) - 1ull);
        T.q_step[T.n] = (float)(((double)h.hi - (double)h.lo) / levels);
        T.qs_norm[T.n] = (float)(1.0 / levels);
        const uint64_t units = (uint64_t)(h.n + 3) / 4;  // four elements per thread, both directions
        blocks += (units + kThreads - 1) / kThreads;
        if (blocks > 0x7fffffffull) return failf(2, "%s: too many elements for one launch", who);
        T.t[T.n] = h;
        T.block_end[T.n] = (uint32_t)blocks;
        ++T.n;
    }
    return 0;
}

extern "C" int gsb200_fake_quantize_multi_fwd(uint32_t n_tensors, const gsb200_fq_tensor *tensors, gsb200_stream_t stream) {
    FqTable T{};
    uint64_t blocks = 0;
    if (int e = fq_table("gsb200_fake_quantize_multi_fwd", n_tensors, tensors, false, T, blocks)) return e;
    if (T.n == 0) return 0;
    fake_quantize_multi_fwd_kernel<<<(uint32_t)blocks, kThreads, 0, (cudaStream_t)stream>>>(T);
    return check_launch("fake_quantize_multi_fwd");
}

extern "C" int gsb200_fake_quantize_multi_bwd(uint32_t n_tensors, const gsb200_fq_tensor *tensors, gsb200_stream_t stream) {
    FqTable T{};
    uint64_t blocks = 0;
    if (int e = fq_table("gsb200_fake_quantize_multi_bwd", n_tensors, tensors, true, T, blocks)) return e;
    if (T.n == 0) return 0;
    fake_quantize_multi_bwd_kernel<<<(uint32_t)blocks, kThreads, 0, (cudaStream_t)stream>>>(T);
    return check_launch("fake_quantize_multi_bwd");
}

extern "C" int gsb200_fake_quantize_bwd(int64_t n, const float *x, const float *y, float lo, float hi, int mode, int act,
                                        const float *v_y, float *v_x, gsb200_stream_t stream) {
    if (n == 0) return 0;
    fake_quantize_bwd_kernel<<<(unsigned)(((n + 3) / 4 + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n, x, y, lo, hi, mode, act, v_y, v_x);
    return check_launch("fake_quantize_bwd");
}

static int entropy_launch(bool bwd, int64_t N, uint32_t ch, uint32_t n_hidden, const float *x, float q, const float *params,
                          float bound, float *bits, const float *v_bits, float *v_x, float *v_params, cudaStream_t s) {
    if (n_hidden != 2 && n_hidden != 3) return fail(2, "gsb200_entropy: n_hidden must be 2 or 3 (filters (3,3) or (3,3,3))");
    if (ch == 0 || ch > 65535) return fail(2, "gsb200_entropy: bad channel count");
    if (N == 0) return 0;
    const int64_t Npad = N + (32 - N % 32);  // entropy_model.py:223 (a full extra 32 when N % 32 == 0)
    const int64_t L = Npad / 32;
    const int64_t full = (L + kEntropyThreads - 1) / kEntropyThreads;
    // backward: few blocks per segment so that each thread accumulates ~16 elements before the reduction
    const int64_t per_seg = bwd ? (full < 16 ? full : 16) : full;
    dim3 grid((unsigned)per_seg, 32, ch);
    if (!bwd) {
        if (n_hidden == 2) entropy_kernel<2, false><<<grid, kEntropyThreads, 0, s>>>(N, ch, L, x, q, params, bound, bits, nullptr, nullptr, nullptr);
        else entropy_kernel<3, false><<<grid, kEntropyThreads, 0, s>>>(N, ch, L, x, q, params, bound, bits, nullptr, nullptr, nullptr);
    } else {
        if (n_hidden == 2) entropy_kernel<2, true><<<grid, kEntropyThreads, 0, s>>>(N, ch, L, x, q, params, bound, nullptr, v_bits, v_x, v_params);
        else entropy_kernel<3, true><<<grid, kEntropyThreads, 0, s>>>(N, ch, L, x, q, params, bound, nullptr, v_bits, v_x, v_params);
    }
    return check_launch(bwd ? "entropy_bwd" : "entropy_fwd");
}

extern "C" int gsb200_entropy_fwd(int64_t N, uint32_t ch, uint32_t n_hidden, const float *x, float q, const float *params,
                                  float likelihood_bound, float *bits, gsb200_stream_t stream) {
    return entropy_launch(false, N, ch, n_hidden, x, q, params, likelihood_bound, bits, nullptr, nullptr, nullptr,
                          (cudaStream_t)stream);
}

extern "C" int gsb200_entropy_bwd(int64_t N, uint32_t ch, uint32_t n_hidden, const float *x, float q, const float *params,
                                  float likelihood_bound, const float *v_bits, float *v_x, float *v_params,
                                  gsb200_stream_t stream) {
    return entropy_launch(true, N, ch, n_hidden, x, q, params, likelihood_bound, nullptr, v_bits, v_x, v_params,
                          (cudaStream_t)stream);
}

static int entropy_raw(const char *who, uint32_t n_hidden, const float *const *matrices, const float *const *bias,
                       const float *const *factor, EntropyRaw &R) {
    if (n_hidden != 2 && n_hidden != 3) return failf(2, "%s: n_hidden must be 2 or 3", who);
    R = EntropyRaw{};
    for (uint32_t i = 0; i <= n_hidden; ++i) {
        if (matrices == nullptr || matrices[i] == nullptr) return failf(2, "%s: null matrix pointer", who);
        R.matrices[i] = matrices[i];
        if (bias != nullptr) R.bias[i] = bias[i];
        if (i < n_hidden) {
            if (factor == nullptr || factor[i] == nullptr) return failf(2, "%s: null factor pointer", who);
            R.factor[i] = factor[i];
        }
    }
    return 0;
}

extern "C" int gsb200_entropy_params_fwd(uint32_t ch, uint32_t n_hidden, const float *const *matrices, const float *const *bias,
                                         const float *const *factor, float *params, gsb200_stream_t stream) {
    const char *who = "gsb200_entropy_params_fwd";
    EntropyRaw R;
    if (int e = entropy_raw(who, n_hidden, matrices, bias, factor, R)) return e;
    if (ch == 0) return 0;
    for (uint32_t i = 0; i <= n_hidden; ++i)
        if (R.bias[i] == nullptr) return failf(2, "%s: null bias pointer", who);
    if (params == nullptr) return failf(2, "%s: null output", who);
    const uint32_t PS = 9 + (n_hidden - 1) * 15 + 4, n = ch * PS;
    entropy_params_kernel<false><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ch, n_hidden, PS, R, params, nullptr);
    return check_launch("entropy_params_fwd");
}

extern "C" int gsb200_entropy_params_bwd(uint32_t ch, uint32_t n_hidden, const float *const *matrices, const float *const *factor,
                                         const float *v_params, float *const *v_matrices, float *const *v_bias,
                                         float *const *v_factor, gsb200_stream_t stream) {
    const char *who = "gsb200_entropy_params_bwd";
    EntropyRaw R;
    if (int e = entropy_raw(who, n_hidden, matrices, nullptr, factor, R)) return e;
    if (ch == 0) return 0;
    if (v_params == nullptr) return failf(2, "%s: null v_params", who);
    for (uint32_t i = 0; i <= n_hidden; ++i) {
        R.v_matrices[i] = v_matrices ? v_matrices[i] : nullptr;
        R.v_bias[i] = v_bias ? v_bias[i] : nullptr;
        if (i < n_hidden) R.v_factor[i] = v_factor ? v_factor[i] : nullptr;
    }
    const uint32_t PS = 9 + (n_hidden - 1) * 15 + 4, n = ch * PS;
    entropy_params_kernel<true><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ch, n_hidden, PS, R, nullptr, v_params);
    return check_launch("entropy_params_bwd");
}
