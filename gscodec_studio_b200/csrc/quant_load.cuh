// Fake-quantisation applied where an attribute is LOADED (noise mode of gsplat/compression_simulation/ops.py:39-54 plus the
// activation the trainer applies next, examples/simple_trainer.py:758-776): the projection / colour kernels read the RAW
// parameter and see clamp(x, lo, hi) + u * q_step (then exp / sigmoid), the backward kernels hand the gradient straight back to
// the raw parameter (clamp gate, activation derivative).  No quantised copy of the attribute is written to or read from HBM.
//
// Bit-compatibility with the stand-alone kernels of quant.cu (which other consumers -- the entropy model, tests -- still
// use): same Philox counter scheme (one counter per 4 consecutive elements of the tensor), same two roundings for
// clamp + u * q_step, and torch's exp / sigmoid arithmetic.  quant.cu is built without --use_fast_math for that; the
// projection / SH translation units are fast-math, so the libdevice entry point and the IEEE division are named explicitly.
#pragma once
#include "common.cuh"

extern "C" __device__ float __nv_expf(float);  // libdevice's accurate expf (what `expf` is outside --use_fast_math)

namespace gsb200 {

// ---------------------------------------------------------------- Philox4x32-10 (counter RNG)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01_centered(uint32_t r) {  // uniform in (-0.5, 0.5)
    return ((float)(r >> 8) + 0.5f) * (1.f / 16777216.f) - 0.5f;
}

// device-side copy of gsb200_quant_load with the step precomputed; on == 0: the attribute is read as it is
struct QuantLoad {
    const float *noise;
    uint64_t seed, offset;
    float lo, hi, q_step;
    int act, on;
};

// the K uniforms of elements [base, base + K) of the tensor (K <= 4: at most two Philox counters)
template <int K>
__device__ __forceinline__ void quant_uniforms(const QuantLoad &q, int64_t base, float (&u)[K]) {
    if (q.noise != nullptr) {
#pragma unroll
        for (int k = 0; k < K; ++k) u[k] = __ldg(q.noise + base + k);
        return;
    }
    const uint2 key = make_uint2((uint32_t)q.seed, (uint32_t)(q.seed >> 32));
    uint64_t c = q.offset + (uint64_t)(base >> 2);
    uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
    const int first = (int)(base & 3);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int lane = (first + k) & 3;
        if (k > 0 && lane == 0) {  // crossed into the next counter
            ++c;
            r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
        }
        const uint32_t w = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
        u[k] = u01_centered(w);
    }
}

// value the consumer sees for raw parameter x with uniform u (same arithmetic as fake_quantize_fwd4 + activate in quant.cu)
__device__ __forceinline__ float quant_value(const QuantLoad &q, float x, float u) {
    const float cl = fminf(fmaxf(x, q.lo), q.hi);
    const float v = __fadd_rn(cl, __fmul_rn(u, q.q_step));
    if (q.act == 1) return __nv_expf(v);
    if (q.act == 2) return __fdiv_rn(1.f, __fadd_rn(1.f, __nv_expf(-v)));
    return v;
}

// d value / d raw parameter, given the value y the forward produced (fake_quantize_bwd_elem in quant.cu)
__device__ __forceinline__ float quant_grad(const QuantLoad &q, float g, float x, float y) {
    if (q.act == 1) g *= y;
    else if (q.act == 2) g *= y * (1.f - y);
    return (x >= q.lo && x <= q.hi) ? g : 0.f;
}

// K consecutive elements starting at `base`: raw values in, consumer values out
template <int K>
__device__ __forceinline__ void quant_load(const QuantLoad &q, int64_t base, const float (&x)[K], float (&y)[K]) {
    if (!q.on) {
#pragma unroll
        for (int k = 0; k < K; ++k) y[k] = x[k];
        return;
    }
    float u[K];
    quant_uniforms<K>(q, base, u);
#pragma unroll
    for (int k = 0; k < K; ++k) y[k] = quant_value(q, x[k], u[k]);
}

// host: C-ABI descriptor -> device descriptor (NULL: attribute is read as it is)
inline int make_quant_load(const char *who, const gsb200_quant_load *h, QuantLoad &q) {
    q = QuantLoad{nullptr, 0, 0, 0.f, 0.f, 0.f, 0, 0};
    if (h == nullptr) return 0;
    if (h->act < 0 || h->act > 2) return failf(2, "%s: quant_load.act must be 0, 1 or 2", who);
    if (h->bitwidth < 1 || h->bitwidth > 32) return failf(2, "%s: quant_load.bitwidth must be in [1, 32]", who);
    const double levels = (double)((1ull << h->bitwidth) - 1ull);
    q.noise = h->noise; q.seed = h->seed; q.offset = h->offset; q.lo = h->lo; q.hi = h->hi;
    q.q_step = (float)(((double)h->hi - (double)h->lo) / levels);  // as gsb200_fake_quantize_fwd computes it
    q.act = h->act; q.on = 1;
    return 0;
}

}  // namespace gsb200
