// Stage 3 / 4a for 16x16 tiles and <= 4 colour channels: producer/consumer alpha blend over a packed,
// pre-culled per-tile record stream (sm_100a).
//
// Same per-pixel arithmetic as blend.cu (i.e. the reference's, rasterize_to_pixels_fwd.cu:121-166 and
// rasterize_to_pixels_bwd.cu:192-241), different organisation:
//
//   * one CTA per tile = 4/NP consumer warps + 1 producer warp.  A consumer thread owns NP pairs of pixels
//     (x, y) / (x, y + 4) of an 8x8 block and evaluates each pair with packed FP32 (FFMA2 / FMUL2 / FADD2):
//     one instruction stream for two pixels, per-lane results bit-identical to the scalar expression;
//   * forward: the producer warp gathers the tile's Gaussians through flatten_ids (software-pipelined,
//     L2-resident attributes), tests each 1/255-alpha ellipse against the four 8x8 blocks of the tile
//     (exact, conservative -- blend.cuh), drops the ones that miss the tile, and appends a 48-byte record
//     {x, y, opacity, (local index << 4 | block mask)} {conic a, b, c, flatten id} {colour[4]} to a ring of
//     shared-memory stages guarded by mbarriers -- and to a per-tile stream in global memory.  Consumers
//     never stage, never cull and never meet a block-wide barrier: one LDS + ballot per 32 records;
//   * backward: the stream the forward left behind is contiguous per tile, so one elected thread stages it
//     back-to-front with cp.async.bulk (TMA, mbarrier complete_tx) -- no gather, no re-culling; the per-warp
//     reductions and the packed-record REDs are the ones of blend.cu, amortised over 64 pixels per warp
//     instead of 32.
#include <stdlib.h>

#include "blend.cuh"

namespace gsb200 {

constexpr int kStageRecs = 64;  // records per pipeline stage
#ifndef GSB_STREAM_STAGES
#define GSB_STREAM_STAGES 6
#endif
constexpr int kStages = GSB_STREAM_STAGES;
// The forward's ring may use smaller stages than the backward's bulk copies (a consumer can start on a stage only when the
// producer has completed it, and half of the tiles hold fewer than 128 records).  Measured on B200: 64 / 48 / 32 records per
// stage all 292-296 us -- 64 stays (tools/gpu_r2_visit43.sh)
#ifndef GSB_FWD_STAGE_RECS
#define GSB_FWD_STAGE_RECS 64
#endif
constexpr int kFStageRecs = GSB_FWD_STAGE_RECS;
constexpr int kFStages = kStages * kStageRecs / kFStageRecs;  // same ring capacity
constexpr int kMaxStages = kFStages > kStages ? kFStages : kStages;
static_assert(kFStageRecs >= 32 && kFStageRecs <= 64 && (kStages * kStageRecs) % kFStageRecs == 0, "a chunk of 32 candidates completes at most one stage");
#ifndef GSB_SFWD_BLOCKS
#define GSB_SFWD_BLOCKS 6  // minimum resident CTAs per SM the register allocator must allow (forward, NP = 1)
#endif
#ifndef GSB_SBWD_BLOCKS
#define GSB_SBWD_BLOCKS 7  // (backward, NP = 1, one record per trip -- GSB_BWD_PAIR = 0; that walk is launched with a 5-CTA cap, see bwd_t)
#endif
#ifndef GSB_SBWD_PAIR_BLOCKS
#define GSB_SBWD_PAIR_BLOCKS 8  // two-record walk, final kernel: 8 CTAs / 62 registers 495 us, 7 / 69: 515-524, 6 / 76: 512 (tools/gpu_r2_visit50.sh); no spills in any instantiation
#endif
#ifndef GSB_PRODUCER_SLEEP_NS
#define GSB_PRODUCER_SLEEP_NS 256  // ring full: the consumers need ~10^3 cycles per stage
#endif
#ifndef GSB_CONSUMER_SLEEP_NS
#define GSB_CONSUMER_SLEEP_NS 32
#endif
#ifndef GSB_BWD_MOMENTS
#define GSB_BWD_MOMENTS 1  // backward accumulates moments of -v_sigma (finished per Gaussian in unpack_grads_kernel); 0: round-2 form
#endif
#ifndef GSB_BWD_PIPELINE
#define GSB_BWD_PIPELINE 0  // retire record i's sums inside the iteration of record i + 1 (see blend_stream_bwd_kernel); measured SLOWER (761 vs 668 us): off
#endif
#ifndef GSB_FWD_GROUP
#define GSB_FWD_GROUP 2  // records a forward consumer warp evaluates per trip (NP = 1); 1: one at a time
#endif
#ifndef GSB_BWD_PAIR
#define GSB_BWD_PAIR 1  // backward walks a warp's records two at a time (NP = 1): see blend_stream_bwd_kernel
#endif
constexpr int kRecBytes = 48;
constexpr uint32_t kEndFlag = 0x100;
constexpr float kNegLog2e = -1.4426950408889634f;

// ---- mbarrier / bulk-copy primitives (32-bit shared-window addresses) ----------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// A waiting warp must not burn issue slots of the warps it waits for (the first profile had 19 % of the forward's
// instructions in this loop): back off with nanosleep between probes.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    while (true) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy (TMA, non-tensor form); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds1u(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float2 bc(float v) { return make_float2(v, v); }
// per-thread gradient partials: with one pixel pair per thread the first (only) contribution is an assignment
template <bool ASSIGN>
__device__ __forceinline__ void acc(float &dst, float v) {
    if (ASSIGN) dst = v; else dst += v;
}


// Which of the four 8x8 blocks of a tile does the ellipse { d : d^T A d <= qmax } around (mx, my) meet?  Same exact,
// conservative test as ellipse_hits_rect (blend.cuh) -- the minimum of the convex form over a block's pixel-centre
// rectangle is 0 if the centre is inside, else it lies on one of the four edges -- evaluated for the 2 x 2 blocks at
// once: the unconstrained minimiser along each of the 4 vertical / 4 horizontal grid lines is shared by the two blocks
// the line bounds.  xs / ys: pixel-centre coordinates of the block edges {lo0, hi0, lo1, hi1}; an interval that lies
// outside the image has hi < lo (block never hit).  ~150 instructions instead of 4 x 100.
__device__ __forceinline__ uint32_t blocks_hit_mask(float mx, float my, float a, float b, float c, float qmax, const float *xs,
                                                    const float *ys) {
    if (!(qmax >= 0.f)) return 0u;
    uint32_t valid = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (xs[2 * (k & 1) + 1] >= xs[2 * (k & 1)] && ys[2 * (k >> 1) + 1] >= ys[2 * (k >> 1)]) valid |= 1u << k;
    if (qmax == FLT_MAX) return valid;
    float lx[4], ly[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { lx[k] = xs[k] - mx; ly[k] = ys[k] - my; }
    const float nbc = -b * __frcp_rn(c), nba = -b * __frcp_rn(a), b2 = 2.f * b;
    // qv[k][j]: minimum along the vertical line x = xs[k] over y-interval j;  qh[k][i]: horizontal line y = ys[k], x-interval i
    float qv[4][2], qh[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float ystar = nbc * lx[k], a2 = a * lx[k] * lx[k], bl = b2 * lx[k];
        const float xstar = nba * ly[k], c2 = c * ly[k] * ly[k], bt = b2 * ly[k];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float dy = fminf(fmaxf(ystar, ly[2 * j]), ly[2 * j + 1]);
            qv[k][j] = (c * dy + bl) * dy + a2;
            const float dx = fminf(fmaxf(xstar, lx[2 * j]), lx[2 * j + 1]);
            qh[k][j] = (a * dx + bt) * dx + c2;
        }
    }
    // rounding margin, as in ellipse_hits_rect, with the extents of the whole tile
    const float ex = fmaxf(fabsf(lx[0]), fabsf(lx[3])), ey = fmaxf(fabsf(ly[0]), fabsf(ly[3]));
    const float bound = qmax + 1e-5f * (a * ex * ex + 2.f * fabsf(b) * ex * ey + c * ey * ey) + 1e-4f;
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = k & 1, j = k >> 1;  // x-interval i, y-interval j
        const bool inside = lx[2 * i] <= 0.f && lx[2 * i + 1] >= 0.f && ly[2 * j] <= 0.f && ly[2 * j + 1] >= 0.f;
        const float q = fminf(fminf(qv[2 * i][j], qv[2 * i + 1][j]), fminf(qh[2 * j][i], qh[2 * j + 1][i]));
        if (inside || q <= bound) mask |= 1u << k;
    }
    return mask & valid;
}

struct StreamShared {
    float4 rec[kStages * kStageRecs * 3];
    unsigned long long full[kMaxStages];
    unsigned long long empty[kMaxStages];
    int count[kMaxStages];
    int done;        // consumer warps that need no more records (forward)
    int last_stage;  // last stage in which any pixel of the tile took a contribution (forward)
    int emitted;     // records the producer emitted (forward)
    int finished;    // warps that have left their loop (forward): the last one publishes the tile's record count
};

// pixel pair p of warp w: block b = w * NP + p at (b & 1, b >> 1) * 8; lane -> column lane & 7, rows lane >> 3 and + 4
template <int NP>
__device__ __forceinline__ void pair_pixels(uint32_t w, uint32_t lane, uint32_t p, uint32_t tile_x, uint32_t tile_y, uint32_t &j,
                                            uint32_t &i0) {
    const uint32_t b = w * NP + p;
    j = tile_x * 16 + (b & 1) * 8 + (lane & 7);
    i0 = tile_y * 16 + (b >> 1) * 8 + (lane >> 3);
}

// ---------------------------------------------------------------------------------------------- forward
template <int CDIM, int NP>
__global__ void __launch_bounds__((4 / NP + 1) * 32, NP == 1 ? GSB_SFWD_BLOCKS : 8)
blend_stream_fwd_kernel(BlendGeom G, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                        const float *__restrict__ colors, const float *__restrict__ opacities,
                        const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks,
                        const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                        float *__restrict__ render_colors, float *__restrict__ render_alphas, int32_t *__restrict__ last_ids,
                        const int32_t *__restrict__ tile_order, float4 *__restrict__ stream, int32_t *__restrict__ tile_counts) {
    constexpr uint32_t NCONS = 4 / NP;
    __shared__ __align__(128) StreamShared S;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ct_lin = tile_order ? (uint32_t)tile_order[blockIdx.x] : blockIdx.x;
    const uint32_t cam = ct_lin / (G.tw * G.th), tile_id = ct_lin % (G.tw * G.th);
    const uint32_t tile_y = tile_id / G.tw, tile_x = tile_id % G.tw;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;

    if (masks != nullptr && !masks[ct]) {  // rasterize_to_pixels_fwd.cu:70-76
        if (warp < NCONS) {
#pragma unroll
            for (uint32_t p = 0; p < NP; ++p) {
                uint32_t j, i0;
                pair_pixels<NP>(warp, lane, p, tile_x, tile_y, j, i0);
#pragma unroll
                for (uint32_t r = 0; r < 2; ++r) {
                    const uint32_t i = i0 + 4 * r;
                    if (i < G.H && j < G.W) {
                        const size_t pix = ((size_t)cam * G.H + i) * G.W + j;
#pragma unroll
                        for (int k = 0; k < CDIM; ++k) render_colors[pix * CDIM + k] = bg ? bg[k] : 0.f;
                    }
                }
            }
        }
        if (tid == 0 && tile_counts) tile_counts[ct] = 0;
        return;
    }
    const int32_t range_start = tile_offsets[ct];
    const int32_t range_end = (cam == G.C - 1 && tile_id == G.tw * G.th - 1) ? (int32_t)G.n_isects : tile_offsets[ct + 1];
    const uint32_t n = (uint32_t)(range_end - range_start);

    const uint32_t a_rec = (uint32_t)__cvta_generic_to_shared(S.rec);
    const uint32_t a_full = (uint32_t)__cvta_generic_to_shared(S.full), a_empty = (uint32_t)__cvta_generic_to_shared(S.empty);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kFStages; ++s) {
            mbar_init(a_full + 8 * s, 32);      // every producer lane arrives
            mbar_init(a_empty + 8 * s, NCONS);  // lane 0 of every consumer warp arrives
        }
        S.done = 0;
        S.last_stage = -1;
        S.emitted = 0;
        S.finished = 0;
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == NCONS) {
        // ------------------------------------------------------------------ producer warp
        // pixel-centre coordinates of the block edges {lo0, hi0, lo1, hi1} along x and y, clipped to the image (an
        // interval outside the image: hi < lo)
        float xs[4], ys[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int xl = (int)(tile_x * 16 + i * 8), yl = (int)(tile_y * 16 + i * 8);
            xs[2 * i] = (float)xl + 0.5f;
            xs[2 * i + 1] = (float)min(xl + 7, (int)G.W - 1) + 0.5f;
            ys[2 * i] = (float)yl + 0.5f;
            ys[2 * i + 1] = (float)min(yl + 7, (int)G.H - 1) + 0.5f;
        }
        struct Cand {
            float2 xy;
            float a, b, c, o;
            float col[4];
            int32_t g;
        };
        auto load = [&](int32_t g) {
            Cand q;
            q.g = g;
            q.col[0] = q.col[1] = q.col[2] = q.col[3] = 0.f;
            if (g >= 0) {
                q.xy = __ldg(means2d + g);
                q.o = __ldg(opacities + g);
                q.a = __ldg(conics + (size_t)g * 3);
                q.b = __ldg(conics + (size_t)g * 3 + 1);
                q.c = __ldg(conics + (size_t)g * 3 + 2);
#pragma unroll
                for (int k = 0; k < CDIM; ++k) q.col[k] = __ldg(colors + (size_t)g * CDIM + k);
            } else {
                q.xy = make_float2(0.f, 0.f);
                q.a = q.b = q.c = q.o = 0.f;
            }
            return q;
        };
        auto ids = [&](uint32_t at) { return at + lane < n ? __ldg(flatten_ids + range_start + at + lane) : -1; };

        uint32_t total = 0;            // records emitted so far
        uint32_t acquired = kFStages;   // stages [0, acquired) may be written (the first ring round is free)
        auto acquire_to = [&](uint32_t stage) {
            while (acquired <= stage) {
                mbar_wait<GSB_PRODUCER_SLEEP_NS>(a_empty + 8 * (acquired % kFStages), ((acquired / kFStages) - 1) & 1);
                ++acquired;
            }
        };
        const uint32_t lt_mask = (1u << lane) - 1u;
        int32_t g_next = ids(32);
        Cand cur = load(ids(0));
        for (uint32_t base = 0; base < n; base += 32) {
            if (*reinterpret_cast<volatile int *>(&S.done) >= (int)NCONS) break;
            // software pipeline: attributes of the next chunk and ids of the one after it are in flight while this
            // chunk is tested
            const Cand nxt = load(g_next);
            g_next = ids(base + 64);

            uint32_t hitmask = 0;
            if (cur.g >= 0)
                hitmask = blocks_hit_mask(cur.xy.x, cur.xy.y, cur.a, cur.b, cur.c, cull_qmax(cur.a, cur.b, cur.c, cur.o), xs, ys);
            const unsigned ballot = __ballot_sync(0xffffffffu, hitmask != 0);
            const uint32_t cnt = __popc(ballot);
            if (cnt) {
                acquire_to((total + cnt - 1) / kFStageRecs);
                if (hitmask) {
                    const uint32_t pos = total + __popc(ballot & lt_mask);
                    const uint32_t slot = pos % (kFStages * kFStageRecs);
                    const float4 q0 = make_float4(cur.xy.x, cur.xy.y, cur.o, __uint_as_float(((base + lane) << 4) | hitmask));
                    const float4 q1 = make_float4(cur.a, cur.b, cur.c, __int_as_float(cur.g));
                    const float4 q2 = make_float4(cur.col[0], cur.col[1], cur.col[2], cur.col[3]);
                    const uint32_t dst = a_rec + slot * kRecBytes;
                    sts4(dst, q0);
                    sts4(dst + 16, q1);
                    sts4(dst + 32, q2);
                    if (stream != nullptr) {
                        float4 *out = stream + ((size_t)range_start + pos) * 3;
                        out[0] = q0;
                        out[1] = q1;
                        out[2] = q2;
                    }
                }
                const uint32_t st_old = total / kFStageRecs;
                total += cnt;
                if (total / kFStageRecs != st_old) {  // stage st_old is complete: publish it
                    if (lane == 0) S.count[st_old % kFStages] = kFStageRecs;
                    __syncwarp();
                    mbar_arrive(a_full + 8 * (st_old % kFStages));
                }
            }
            cur = nxt;
        }
        // last (possibly empty) stage carries the end flag
        const uint32_t st_last = total / kFStageRecs;
        acquire_to(st_last);
        if (lane == 0) {
            S.count[st_last % kFStages] = (int)(total % kFStageRecs) | (int)kEndFlag;
            S.emitted = (int)total;
        }
        __syncwarp();
        mbar_arrive(a_full + 8 * (st_last % kFStages));
    } else {
        // ------------------------------------------------------------------ consumer warps
        uint32_t pj[NP], pi0[NP];
        float px[NP];
        float2 npy[NP];  // minus the pixel-centre rows
        // T <= 0 marks a pixel that takes no more Gaussians (saturated: -T is its transmittance; outside the image: -1)
        float2 T[NP], out[NP][CDIM];
        uint32_t cur0[NP], cur1[NP];  // record word (local index << 4 | mask, never 0) of the last contribution
        bool all_done = true;
#pragma unroll
        for (uint32_t p = 0; p < NP; ++p) {
            pair_pixels<NP>(warp, lane, p, tile_x, tile_y, pj[p], pi0[p]);
            px[p] = (float)pj[p] + 0.5f;
            npy[p] = make_float2(-((float)pi0[p] + 0.5f), -((float)(pi0[p] + 4) + 0.5f));
            const bool in0 = pj[p] < G.W && pi0[p] < G.H, in1 = pj[p] < G.W && pi0[p] + 4 < G.H;
            all_done = all_done && !in0 && !in1;
            T[p] = make_float2(in0 ? 1.f : -1.f, in1 ? 1.f : -1.f);
            cur0[p] = cur1[p] = 0;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) out[p][k] = make_float2(0.f, 0.f);
        }
        bool warp_done = __ballot_sync(0xffffffffu, !all_done) == 0;
        if (warp_done && lane == 0) atomicAdd(&S.done, 1);
        const uint32_t hitsel = ((1u << NP) - 1u) << (warp * NP);
        int my_last_stage = -1;

        for (uint32_t k = 0;; ++k) {
            const uint32_t slot = k % kFStages;
            mbar_wait<GSB_CONSUMER_SLEEP_NS>(a_full + 8 * slot, (k / kFStages) & 1);
            const int cf = *reinterpret_cast<volatile int *>(&S.count[slot]);
            const uint32_t nrec = (uint32_t)cf & 0xffu;
            if (!warp_done) {
                uint32_t cur_sig = 0;  // changes iff some pixel of this thread takes a Gaussian of this stage
#pragma unroll
                for (uint32_t p = 0; p < NP; ++p) cur_sig += cur0[p] + cur1[p];
                const uint32_t a_stage = a_rec + slot * (kFStageRecs * kRecBytes);
                for (uint32_t c0 = 0; c0 < nrec; c0 += 32) {
                    const uint32_t e = c0 + lane;
                    const uint32_t word = e < nrec ? lds1u(a_stage + e * kRecBytes + 12) : 0u;
                    unsigned m = __ballot_sync(0xffffffffu, (word & hitsel) != 0);
#if GSB_FWD_GROUP > 1
                    // GSB_FWD_GROUP records per trip (NP = 1): their sigma / exp2 chains are independent, only T, the
                    // last-contributor word and the colour sums pass from one record to the next -- same per-lane arithmetic in
                    // the same order, evaluated as one basic block so that the chains interleave
                    while (NP == 1 && m) {
                        constexpr int K = GSB_FWD_GROUP;
                        uint32_t t[K];
                        bool have[K];
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            have[r] = m != 0;
                            t[r] = have[r] ? c0 + (uint32_t)__ffs(m) - 1u : t[0];
                            m &= m - 1;
                        }
                        float4 A[K], B[K], C4[K];
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            A[r] = lds4(a_stage + t[r] * kRecBytes);
                            B[r] = lds4(a_stage + t[r] * kRecBytes + 16);
                        }
#pragma unroll
                        for (int r = 0; r < K; ++r) C4[r] = lds4(a_stage + t[r] * kRecBytes + 32);
                        float2 sigma[K], alpha[K];
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            const float dx = A[r].x - px[0];
                            const float2 dy = __fadd2_rn(bc(A[r].y), npy[0]);
                            float2 tt = __fmul2_rn(bc(B[r].z), dy);
                            const float u = B[r].x * dx, v = B[r].y * dx;
                            tt = __fmul2_rn(dy, tt);
                            float2 uu = __ffma2_rn(bc(dx), bc(u), tt);
                            uu = __fmul2_rn(uu, bc(0.5f));
                            sigma[r] = __ffma2_rn(dy, bc(v), uu);
                            const float2 ee = __fmul2_rn(sigma[r], bc(kNegLog2e));
                            const float2 ov = __fmul2_rn(bc(A[r].z), make_float2(exp2f(ee.x), exp2f(ee.y)));
                            // no such record: alpha = 0 fails the 1/255 test for both pixels, the record is a no-op
                            alpha[r] = make_float2((r == 0 || have[r]) ? fminf(kAlphaMax, ov.x) : 0.f,
                                                   (r == 0 || have[r]) ? fminf(kAlphaMax, ov.y) : 0.f);
                        }
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            const float cs[4] = {C4[r].x, C4[r].y, C4[r].z, C4[r].w};
                            const uint32_t wt = __float_as_uint(A[r].w);
                            const float2 next_T = __fmul2_rn(T[0], __ffma2_rn(alpha[r], bc(-1.f), bc(1.f)));
                            const bool ok0 = !(sigma[r].x < 0.f || alpha[r].x < kAlphaMin), ok1 = !(sigma[r].y < 0.f || alpha[r].y < kAlphaMin);
                            const bool take0 = ok0 && !(next_T.x <= kTStop), take1 = ok1 && !(next_T.y <= kTStop);
                            float2 vis = __fmul2_rn(alpha[r], T[0]);
                            vis.x = take0 ? vis.x : 0.f;
                            vis.y = take1 ? vis.y : 0.f;
#pragma unroll
                            for (int c = 0; c < CDIM; ++c) out[0][c] = __ffma2_rn(bc(cs[c]), vis, out[0][c]);
                            T[0].x = take0 ? next_T.x : (ok0 ? -fabsf(T[0].x) : T[0].x);
                            T[0].y = take1 ? next_T.y : (ok1 ? -fabsf(T[0].y) : T[0].y);
                            cur0[0] = take0 ? wt : cur0[0];
                            cur1[0] = take1 ? wt : cur1[0];
                        }
                    }
#endif
                    while (m) {
                        const uint32_t t = c0 + (uint32_t)__ffs(m) - 1u;
                        m &= m - 1;
                        const float4 A = lds4(a_stage + t * kRecBytes);
                        const float4 B = lds4(a_stage + t * kRecBytes + 16);
                        const float4 C4 = lds4(a_stage + t * kRecBytes + 32);
                        const float cs[4] = {C4.x, C4.y, C4.z, C4.w};
                        const uint32_t wt = __float_as_uint(A.w);
#pragma unroll
                        for (uint32_t p = 0; p < NP; ++p) {
                            if (NP > 1 && !((wt >> (warp * NP + p)) & 1u)) continue;
                            // the scalar expression of blend.cu, two rows at a time (contraction spelled out)
                            const float dx = A.x - px[p];
                            const float2 dy = __fadd2_rn(bc(A.y), npy[p]);
                            float2 tt = __fmul2_rn(bc(B.z), dy);
                            const float u = B.x * dx, v = B.y * dx;
                            tt = __fmul2_rn(dy, tt);
                            float2 uu = __ffma2_rn(bc(dx), bc(u), tt);
                            uu = __fmul2_rn(uu, bc(0.5f));
                            const float2 sigma = __ffma2_rn(dy, bc(v), uu);
                            const float2 ee = __fmul2_rn(sigma, bc(kNegLog2e));
                            const float2 ov = __fmul2_rn(bc(A.z), make_float2(exp2f(ee.x), exp2f(ee.y)));
                            const float2 alpha = make_float2(fminf(kAlphaMax, ov.x), fminf(kAlphaMax, ov.y));
                            const float2 next_T = __fmul2_rn(T[p], __ffma2_rn(alpha, bc(-1.f), bc(1.f)));
                            const bool ok0 = !(sigma.x < 0.f || alpha.x < kAlphaMin), ok1 = !(sigma.y < 0.f || alpha.y < kAlphaMin);
                            // a finished pixel has T <= 0, hence next_T <= 0 <= 1e-4: it can never take
                            const bool take0 = ok0 && !(next_T.x <= kTStop), take1 = ok1 && !(next_T.y <= kTStop);
                            float2 vis = __fmul2_rn(alpha, T[p]);
                            vis.x = take0 ? vis.x : 0.f;
                            vis.y = take1 ? vis.y : 0.f;
#pragma unroll
                            for (int c = 0; c < CDIM; ++c) out[p][c] = __ffma2_rn(bc(cs[c]), vis, out[p][c]);
                            T[p].x = take0 ? next_T.x : (ok0 ? -fabsf(T[p].x) : T[p].x);
                            T[p].y = take1 ? next_T.y : (ok1 ? -fabsf(T[p].y) : T[p].y);
                            cur0[p] = take0 ? wt : cur0[p];
                            cur1[p] = take1 ? wt : cur1[p];
                        }
                    }
                    bool pending = false;
#pragma unroll
                    for (uint32_t p = 0; p < NP; ++p) pending = pending || T[p].x > 0.f || T[p].y > 0.f;
                    if (__ballot_sync(0xffffffffu, pending) == 0) {
                        warp_done = true;
                        if (lane == 0) atomicAdd(&S.done, 1);
                        break;
                    }
                }
                uint32_t cur_now = 0;
#pragma unroll
                for (uint32_t p = 0; p < NP; ++p) cur_now += cur0[p] + cur1[p];
                if (__any_sync(0xffffffffu, cur_now != cur_sig)) my_last_stage = (int)k;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(a_empty + 8 * slot);
            if (cf & (int)kEndFlag) break;
        }
        if (lane == 0 && my_last_stage >= 0) atomicMax(&S.last_stage, my_last_stage);

#pragma unroll
        for (uint32_t p = 0; p < NP; ++p) {
#pragma unroll
            for (uint32_t r = 0; r < 2; ++r) {
                const uint32_t i = pi0[p] + 4 * r;
                if (i < G.H && pj[p] < G.W) {
                    const size_t pix = ((size_t)cam * G.H + i) * G.W + pj[p];
                    const float Tf = fabsf(r ? T[p].y : T[p].x);
                    render_alphas[pix] = 1.f - Tf;
#pragma unroll
                    for (int k = 0; k < CDIM; ++k) {
                        const float o = r ? out[p][k].y : out[p][k].x;
                        render_colors[pix * CDIM + k] = bg ? (o + Tf * bg[k]) : o;
                    }
                    const uint32_t cl = r ? cur1[p] : cur0[p];
                    // reference: index into the sorted list of the last contributing Gaussian, 0 if none
                    last_ids[pix] = cl ? range_start + (int32_t)(cl >> 4) : 0;
                }
            }
        }
    }
    // the backward needs the stream up to the last stage that contributed to any pixel.  No block barrier here: warps leave
    // as they finish and the LAST one to arrive (it sees every other warp's S.emitted / S.last_stage update) writes the count
    __syncwarp();
    if (lane == 0 && tile_counts != nullptr) {
        __threadfence_block();
        if (atomicAdd(&S.finished, 1) == (int)NCONS) {
            __threadfence_block();
            const int emitted = *reinterpret_cast<volatile int *>(&S.emitted), last = *reinterpret_cast<volatile int *>(&S.last_stage);
            tile_counts[ct] = min(emitted, (last + 1) * kFStageRecs);
        }
    }
}

// --------------------------------------------------------------------------------------------- backward
// No producer warp here: staging is one bulk copy per stage, so the LAST consumer warp to release a ring slot refills it
// (a shared-memory counter per slot plays the "empty" barrier) and every warp of the CTA does blend work.
template <int CDIM, int NP, bool ABS>
__global__ void __launch_bounds__((4 / NP) * 32, NP == 1 ? (GSB_BWD_PAIR ? GSB_SBWD_PAIR_BLOCKS : GSB_SBWD_BLOCKS) : 8)
blend_stream_bwd_kernel(BlendGeom G, const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks,
                        const int32_t *__restrict__ tile_offsets, const float *__restrict__ render_alphas,
                        const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                        const float *__restrict__ v_render_alphas, const int32_t *__restrict__ tile_order,
                        const float4 *__restrict__ stream, const int32_t *__restrict__ tile_counts,
                        float *__restrict__ grad_rec) {
    constexpr uint32_t NCONS = 4 / NP;
    constexpr uint32_t REC = grad_record(CDIM);
    constexpr int NMAIN = CDIM + 5;  // [rgb(CDIM) conic(3) xy(2)] reduced in groups of 8, then [opac (absx absy)]
    constexpr int NGRP = (NMAIN + 7) / 8;
    __shared__ __align__(128) StreamShared S;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ct_lin = tile_order ? (uint32_t)tile_order[blockIdx.x] : blockIdx.x;
    const uint32_t cam = ct_lin / (G.tw * G.th), tile_id = ct_lin % (G.tw * G.th);
    const uint32_t tile_y = tile_id / G.tw, tile_x = tile_id % G.tw;
    const size_t ct = (size_t)cam * G.tw * G.th + tile_id;
    if (masks != nullptr && !masks[ct]) return;
    const int32_t cnt = tile_counts[ct];
    if (cnt <= 0) return;
    const int32_t range_start = tile_offsets[ct];
    const uint32_t nst = ((uint32_t)cnt + kStageRecs - 1) / kStageRecs;
    const float *bg = backgrounds ? backgrounds + (size_t)cam * CDIM : nullptr;

    const uint32_t a_rec = (uint32_t)__cvta_generic_to_shared(S.rec);
    const uint32_t a_full = (uint32_t)__cvta_generic_to_shared(S.full);
    const char *src = reinterpret_cast<const char *>(stream) + (size_t)range_start * kRecBytes;
    // stage j of the walk (back to front) = records [st * 64, st * 64 + nrec) of the tile's stream, st = nst - 1 - j
    auto issue = [&](uint32_t j) {
        const uint32_t st = nst - 1 - j, slot = j % kStages;
        const uint32_t bytes = min((uint32_t)kStageRecs, (uint32_t)cnt - st * kStageRecs) * kRecBytes;
        mbar_arrive_expect_tx(a_full + 8 * slot, bytes);
        bulk_g2s(a_rec + slot * (kStageRecs * kRecBytes), src + (size_t)st * kStageRecs * kRecBytes, bytes, a_full + 8 * slot);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(a_full + 8 * s, 1);  // the issuing thread's arrive.expect_tx (+ the copy's bytes)
            S.count[s] = 0;                // warps that have released the slot
        }
        fence_barrier_init();
        for (uint32_t j = 0; j < min(nst, (uint32_t)kStages); ++j) issue(j);
    }
    __syncthreads();
    // ---------------------------------------------------------------------- consumer warps
    float px[NP];
    float2 npy[NP], T[NP], Tfin[NP], v_a[NP], bg_dot[NP];
    float2 Bsum[NP], v_c[NP][CDIM];
    int32_t bin0[NP], bin1[NP];
    int32_t warp_bin = -1;
#pragma unroll
    for (uint32_t p = 0; p < NP; ++p) {
        uint32_t j, i0;
        pair_pixels<NP>(warp, lane, p, tile_x, tile_y, j, i0);
        px[p] = (float)j + 0.5f;
        npy[p] = make_float2(-((float)i0 + 0.5f), -((float)(i0 + 4) + 0.5f));
        const bool in0 = j < G.W && i0 < G.H, in1 = j < G.W && i0 + 4 < G.H;
        const size_t pix0 = ((size_t)cam * G.H + i0) * G.W + j, pix1 = pix0 + (size_t)4 * G.W;
        Tfin[p] = make_float2(in0 ? 1.f - render_alphas[pix0] : 1.f, in1 ? 1.f - render_alphas[pix1] : 1.f);
        T[p] = Tfin[p];
        v_a[p] = make_float2(in0 ? v_render_alphas[pix0] : 0.f, in1 ? v_render_alphas[pix1] : 0.f);
        bg_dot[p] = make_float2(0.f, 0.f);
        Bsum[p] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CDIM; ++k) {
            v_c[p][k] = make_float2(in0 ? v_render_colors[pix0 * CDIM + k] : 0.f, in1 ? v_render_colors[pix1 * CDIM + k] : 0.f);
            if (bg) bg_dot[p] = __ffma2_rn(bc(bg[k]), v_c[p][k], bg_dot[p]);
        }
        // local (per-tile) index of the last contributing Gaussian; -1: nothing to do for this pixel
        bin0[p] = in0 ? last_ids[pix0] - range_start : -1;
        bin1[p] = in1 ? last_ids[pix1] - range_start : -1;
        warp_bin = max(warp_bin, max(bin0[p], bin1[p]));
    }
    warp_bin = __reduce_max_sync(0xffffffffu, warp_bin);
    const uint32_t hitsel = ((1u << NP) - 1u) << (warp * NP);
    const uint32_t slot8 = spread_slot<8>(lane), slot4 = spread_slot<4>(lane);
    const bool pub8 = (lane & 3) == 0, pub4 = (lane & 7) == 0;
    const uint32_t slot16 = spread_slot<16>(lane);
    const bool pub16 = (lane & 1) == 0;

    // Software pipeline over the records (GSB_BWD_PIPELINE): the partial sums of record i are kept in registers and
    // retired -- warp reduction + REDs -- inside the iteration of record i + 1, right after that record's loads have been
    // issued: the five dependent shuffle stages (~150 cycles of latency for THIS warp) then overlap with the sigma / exp /
    // reciprocal chain of the next record instead of preceding it.  A tile's warps walk their records serially, so this
    // latency is the critical path of the long tiles that finish last (sm__cycles_active was 79 % of elapsed).
    float pend[NGRP * 8], pend_tail[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NGRP * 8; ++k) pend[k] = 0.f;
    float *pend_rec = nullptr;  // warp-uniform: record the pending sums belong to (nullptr: nothing pending)
    auto retire = [&](float *vals, float *tail, float *rec, bool live) {
#pragma unroll
        for (int gi = 0; gi < NGRP; ++gi) reduce_spread<8>(vals + gi * 8, lane);
        float tail_sum;
        if (ABS) {
            reduce_spread<4>(tail, lane);
            tail_sum = tail[0];
        } else {
            tail_sum = warp_sum(tail[0]);
        }
        // (`live` folds into the lane predicates of the REDs: no branch, so that the compiler can schedule the shuffle
        // stages between the arithmetic of the record whose loads precede this call)
#pragma unroll
        for (int gi = 0; gi < NGRP; ++gi) {
            const uint32_t off = gi * 8 + slot8;
            if (live && pub8 && off < (uint32_t)NMAIN) red_add(rec + off, vals[gi * 8]);
        }
        if (ABS) {
            if (live && pub4 && slot4 < 3) red_add(rec + NMAIN + slot4, tail_sum);
        } else {
            if (live && lane == 0) red_add(rec + NMAIN, tail_sum);
        }
    };

    for (uint32_t j = 0; j < nst; ++j) {
        const uint32_t st = nst - 1 - j, slot = j % kStages;
        mbar_wait<GSB_CONSUMER_SLEEP_NS>(a_full + 8 * slot, (j / kStages) & 1);
        const int32_t nrec = (int32_t)min((uint32_t)kStageRecs, (uint32_t)cnt - st * kStageRecs);
        const uint32_t a_stage = a_rec + slot * (kStageRecs * kRecBytes);
        if (warp_bin >= 0) {
            for (int32_t c0 = 0; c0 < nrec; c0 += 32) {
                const int32_t r = nrec - 1 - c0 - (int32_t)lane;  // records back to front
                const uint32_t word = r >= 0 ? lds1u(a_stage + (uint32_t)r * kRecBytes + 12) : 0u;
                unsigned m = __ballot_sync(0xffffffffu, (word & hitsel) != 0 && (int32_t)(word >> 4) <= warp_bin);
#if GSB_BWD_PAIR
                // Two records per trip (NP = 1).  A warp's iteration is one long dependent chain -- record loads, sigma,
                // exp2, the reciprocal, v_alpha, five shuffle levels -- and with 5 resident CTAs a scheduler has ~4.4 warps
                // to hide it behind (issue slots 68 % busy, stalls 'wait' + 'short scoreboard' 2.9 of 6.4 cycles per
                // instruction, profiles/r02b_blend_stream_ncu.md).  Everything of two consecutive records is independent
                // except the running T and B (one multiply / one FMA per record each), so the two are evaluated as ONE
                // basic block: the arithmetic of both interleaves, and their 2 x 9 sums go through one 16-value + one 2-value
                // halving butterfly (21 shuffles instead of 28).  Per-lane arithmetic of a record is unchanged.
                while (NP == 1 && m) {
                    const uint32_t base_t = (uint32_t)(nrec - 1 - c0);
                    const uint32_t t1 = base_t - ((uint32_t)__ffs(m) - 1u);
                    m &= m - 1;
                    const bool two = m != 0;
                    const uint32_t t2 = two ? base_t - ((uint32_t)__ffs(m) - 1u) : t1;
                    m &= m - 1;
                    const float4 A1 = lds4(a_stage + t1 * kRecBytes), B1 = lds4(a_stage + t1 * kRecBytes + 16);
                    const float4 A2 = lds4(a_stage + t2 * kRecBytes), B2 = lds4(a_stage + t2 * kRecBytes + 16);
                    const float4 C1 = lds4(a_stage + t1 * kRecBytes + 32), C2 = lds4(a_stage + t2 * kRecBytes + 32);
                    // Validity is folded into the operands where it is decided: an invalid pixel continues with alpha = 0 -- then
                    // 1 / (alpha - 1) = -1 exactly, T * 1 = T and fac = 0 * T = 0, i.e. the record is the identity for its T and B
                    // -- and a pixel outside the gate (invalid, or opacity * vis clamped at 0.999) with opacity * vis = 0, hence
                    // -v_sigma = 0.  No predicate outlives the vote, no select follows it.
                    struct P1 {
                        float dx, u, v;
                        float2 dy, gvis, gov, alpha;
                    };
                    auto phase1 = [&](const float4 &A, const float4 &B, bool live, bool &any) {
                        P1 q;
                        // (no such record: an index beyond every pixel's last contributor makes both pixels invalid)
                        const int32_t idx = live ? (int32_t)(__float_as_uint(A.w) >> 4) : INT_MAX;
                        q.dx = A.x - px[0];
                        q.dy = __fadd2_rn(bc(A.y), npy[0]);
                        float2 tt = __fmul2_rn(bc(B.z), q.dy);
                        q.u = B.x * q.dx;
                        q.v = B.y * q.dx;
                        tt = __fmul2_rn(q.dy, tt);
                        float2 uu = __ffma2_rn(bc(q.dx), bc(q.u), tt);
                        uu = __fmul2_rn(uu, bc(0.5f));
                        const float2 sigma = __ffma2_rn(q.dy, bc(q.v), uu);
                        const float2 ee = __fmul2_rn(sigma, bc(kNegLog2e));
                        const float2 vis = make_float2(exp2f(ee.x), exp2f(ee.y));
                        const float2 ov = __fmul2_rn(bc(A.z), vis);
                        const float2 alpha = make_float2(fminf(kAlphaMax, ov.x), fminf(kAlphaMax, ov.y));
                        const bool valid0 = idx <= bin0[0] && sigma.x >= 0.f && alpha.x >= kAlphaMin;
                        const bool valid1 = idx <= bin1[0] && sigma.y >= 0.f && alpha.y >= kAlphaMin;
                        any = __any_sync(0xffffffffu, valid0 || valid1);
                        q.alpha = make_float2(valid0 ? alpha.x : 0.f, valid1 ? alpha.y : 0.f);
                        const bool gate0 = valid0 && ov.x <= kAlphaMax, gate1 = valid1 && ov.y <= kAlphaMax;
                        q.gov = make_float2(gate0 ? ov.x : 0.f, gate1 ? ov.y : 0.f);
                        q.gvis = make_float2(gate0 ? vis.x : 0.f, gate1 ? vis.y : 0.f);  // (ABS / non-moment form only)
                        return q;
                    };
                    bool any1, any2;
                    const P1 q1 = phase1(A1, B1, true, any1), q2 = phase1(A2, B2, two, any2);
                    if (!(any1 || any2)) continue;
                    // per-thread sums of one record (the arithmetic of the single-record loop below); advances T and B
                    auto phase2 = [&](const float4 &B, const float4 &C4, const P1 &q, float *vals, float *tail) {
                        const float cs[4] = {C4.x, C4.y, C4.z, C4.w};
                        const float2 nom = __fadd2_rn(q.alpha, bc(-1.f));
                        const float2 nra = make_float2(1.f / nom.x, 1.f / nom.y);
                        const float2 ra = make_float2(-nra.x, -nra.y);
                        const float2 Tn = __fmul2_rn(T[0], ra);
                        T[0] = Tn;
                        const float2 fac = __fmul2_rn(q.alpha, Tn);
                        float2 Acv = __fmul2_rn(bc(cs[0]), v_c[0][0]);
#pragma unroll
                        for (int k = 1; k < CDIM; ++k) Acv = __ffma2_rn(bc(cs[k]), v_c[0][k], Acv);
                        float2 v_alpha = __ffma2_rn(Tn, Acv, __fmul2_rn(nra, Bsum[0]));
                        Bsum[0] = __ffma2_rn(fac, Acv, Bsum[0]);
#pragma unroll
                        for (int k = 0; k < CDIM; ++k) {
                            const float2 gk = __fmul2_rn(fac, v_c[0][k]);
                            vals[k] = gk.x + gk.y;
                        }
                        v_alpha = __ffma2_rn(__fmul2_rn(Tfin[0], ra), v_a[0], v_alpha);
                        if (bg) v_alpha = __ffma2_rn(__fmul2_rn(Tfin[0], nra), bg_dot[0], v_alpha);
                        const float2 nvs = __fmul2_rn(q.gov, v_alpha);  // -v_sigma
                        if (ABS || !GSB_BWD_MOMENTS) {
                            const float2 hs = __fmul2_rn(nvs, bc(0.5f));
                            const float2 ga = __fmul2_rn(__fmul2_rn(hs, bc(q.dx)), bc(q.dx));
                            const float2 gb = __fmul2_rn(__fmul2_rn(nvs, bc(q.dx)), q.dy);
                            const float2 gc = __fmul2_rn(__fmul2_rn(hs, q.dy), q.dy);
                            const float2 gx = __fmul2_rn(nvs, __ffma2_rn(bc(B.y), q.dy, bc(q.u)));
                            const float2 gy = __fmul2_rn(nvs, __ffma2_rn(bc(B.z), q.dy, bc(q.v)));
                            vals[CDIM + 0] = -ga.x - ga.y;
                            vals[CDIM + 1] = -gb.x - gb.y;
                            vals[CDIM + 2] = -gc.x - gc.y;
                            vals[CDIM + 3] = -gx.x - gx.y;
                            vals[CDIM + 4] = -gy.x - gy.y;
                            const float2 go = __fmul2_rn(q.gvis, v_alpha);
                            tail[0] = go.x + go.y;
                            if (ABS) {
                                tail[1] = fabsf(gx.x) + fabsf(gx.y);
                                tail[2] = fabsf(gy.x) + fabsf(gy.y);
                            }
                        } else {
                            const float n01 = nvs.x + nvs.y;
                            const float2 nd = __fmul2_rn(nvs, q.dy);
                            const float2 ndd = __fmul2_rn(nd, q.dy);
                            const float sy = nd.x + nd.y;
                            const float sx = q.dx * n01;
                            vals[CDIM + 0] = q.dx * sx;
                            vals[CDIM + 1] = q.dx * sy;
                            vals[CDIM + 2] = ndd.x + ndd.y;
                            vals[CDIM + 3] = sx;
                            vals[CDIM + 4] = sy;
                            tail[0] = n01;
                        }
                    };
                    float red[NGRP][16], tp[8];
                    {
                        float va[NGRP * 8], vb[NGRP * 8], ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int k = 0; k < NGRP * 8; ++k) va[k] = vb[k] = 0.f;
                        phase2(B1, C1, q1, va, ta);  // t1 is further back in the tile's list: it comes first in the walk
                        phase2(B2, C2, q2, vb, tb);
#pragma unroll
                        for (int gi = 0; gi < NGRP; ++gi)
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                red[gi][k] = va[gi * 8 + k];
                                red[gi][8 + k] = vb[gi * 8 + k];
                            }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tp[k] = ABS ? ta[k] : (k == 0 ? ta[0] : (k == 1 ? tb[0] : 0.f));
                            tp[4 + k] = ABS ? tb[k] : 0.f;
                        }
                    }
#pragma unroll
                    for (int gi = 0; gi < NGRP; ++gi) reduce_spread<16>(red[gi], lane);
                    if (ABS) reduce_spread<8>(tp, lane); else reduce_spread<2>(tp, lane);
                    // (a record none of whose pixels is valid -- or the stand-in for a missing second record -- has all-zero sums:
                    // adding them is harmless and keeps the publishing lanes free of warp-uniform predicates.)  After the first
                    // halving step the upper half-warp holds the second record's values in every butterfly: one select of the id
                    float *rec = grad_rec + (size_t)((lane & 16u) ? __float_as_int(B2.w) : __float_as_int(B1.w)) * REC;
#pragma unroll
                    for (int gi = 0; gi < NGRP; ++gi)
                        red_add_if(rec + gi * 8 + (slot16 & 7u), red[gi][0], pub16 && gi * 8 + (slot16 & 7u) < (uint32_t)NMAIN);
                    if (ABS)
                        red_add_if(rec + NMAIN + (slot8 & 3u), tp[0], pub8 && (slot8 & 3u) < 3);
                    else
                        red_add_if(rec + NMAIN, tp[0], (lane & 15u) == 0);
                }
#endif
                while (m) {
                    const uint32_t t = (uint32_t)(nrec - 1 - c0) - ((uint32_t)__ffs(m) - 1u);
                    m &= m - 1;
                    const float4 A = lds4(a_stage + t * kRecBytes);
                    const float4 B = lds4(a_stage + t * kRecBytes + 16);
                    const float4 C4 = lds4(a_stage + t * kRecBytes + 32);
                    const float cs[4] = {C4.x, C4.y, C4.z, C4.w};
                    const uint32_t wt = __float_as_uint(A.w);
                    const int32_t idx = (int32_t)(wt >> 4);
                    const float opac = A.z;
                    if (GSB_BWD_PIPELINE) {  // (unconditional shuffles, predicated REDs: one basic block with the arithmetic below)
                        retire(pend, pend_tail, pend_rec, pend_rec != nullptr);
                        pend_rec = nullptr;
                    }
                    // (pipelined: the sums are formed directly in the pending registers the retire above has just consumed)
                    float vals_local[NGRP * 8], tail_local[4];
                    float *vals = GSB_BWD_PIPELINE ? pend : vals_local, *tail = GSB_BWD_PIPELINE ? pend_tail : tail_local;
#pragma unroll
                    for (int k = 0; k < 4; ++k) tail[k] = 0.f;
#pragma unroll
                    for (int k = 0; k < NGRP * 8; ++k) vals[k] = 0.f;
                    bool any_valid = false;
#pragma unroll
                    for (uint32_t p = 0; p < NP; ++p) {
                        if (NP > 1 && !((wt >> (warp * NP + p)) & 1u)) continue;
                        const float dx = A.x - px[p];
                        const float2 dy = __fadd2_rn(bc(A.y), npy[p]);
                        float2 tt = __fmul2_rn(bc(B.z), dy);
                        const float u = B.x * dx, v = B.y * dx;
                        tt = __fmul2_rn(dy, tt);
                        float2 uu = __ffma2_rn(bc(dx), bc(u), tt);
                        uu = __fmul2_rn(uu, bc(0.5f));
                        const float2 sigma = __ffma2_rn(dy, bc(v), uu);
                        const float2 ee = __fmul2_rn(sigma, bc(kNegLog2e));
                        const float2 vis = make_float2(exp2f(ee.x), exp2f(ee.y));
                        const float2 ov = __fmul2_rn(bc(opac), vis);
                        const float2 alpha = make_float2(fminf(kAlphaMax, ov.x), fminf(kAlphaMax, ov.y));
                        const bool valid0 = idx <= bin0[p] && !(sigma.x < 0.f || alpha.x < kAlphaMin);
                        const bool valid1 = idx <= bin1[p] && !(sigma.y < 0.f || alpha.y < kAlphaMin);
                        if (!__any_sync(0xffffffffu, valid0 || valid1)) continue;
                        any_valid = true;
                        // packed operands have no negate modifier: carry the sign in ra (nra = -1 / (1 - alpha), exact) and in
                        // the v_sigma chain (nvs = -v_sigma) and restore it in the per-thread sums, where negation is free
                        const float2 nom = __fadd2_rn(alpha, bc(-1.f));
                        const float2 nra = make_float2(1.f / nom.x, 1.f / nom.y);
                        const float2 ra = make_float2(-nra.x, -nra.y);
                        const float2 Tn = __fmul2_rn(T[p], ra);
                        T[p].x = valid0 ? Tn.x : T[p].x;
                        T[p].y = valid1 ? Tn.y : T[p].y;
                        float2 fac = __fmul2_rn(alpha, Tn);
                        fac.x = valid0 ? fac.x : 0.f;
                        fac.y = valid1 ? fac.y : 0.f;
                        // sum_k (c_k T - buffer_k ra) v_k  =  T * A - ra * B  with A = sum_k c_k v_k and B = sum_k buffer_k v_k;
                        // buffer_k only ever enters through B, and buffer_k += c_k fac means B += fac * A: one running
                        // scalar per pixel instead of CDIM accumulated colours (rasterize_to_pixels_bwd.cu:211-216)
                        float2 A2 = __fmul2_rn(bc(cs[0]), v_c[p][0]);
#pragma unroll
                        for (int k = 1; k < CDIM; ++k) A2 = __ffma2_rn(bc(cs[k]), v_c[p][k], A2);
                        float2 v_alpha = __ffma2_rn(Tn, A2, __fmul2_rn(nra, Bsum[p]));
                        Bsum[p] = __ffma2_rn(fac, A2, Bsum[p]);
#pragma unroll
                        for (int k = 0; k < CDIM; ++k) {
                            const float2 gk = __fmul2_rn(fac, v_c[p][k]);
                            acc<NP == 1>(vals[k], gk.x + gk.y);
                        }
                        v_alpha = __ffma2_rn(__fmul2_rn(Tfin[p], ra), v_a[p], v_alpha);
                        if (bg) v_alpha = __ffma2_rn(__fmul2_rn(Tfin[p], nra), bg_dot[p], v_alpha);
                        const bool gate0 = valid0 && ov.x <= kAlphaMax, gate1 = valid1 && ov.y <= kAlphaMax;
                        float2 nvs = __fmul2_rn(ov, v_alpha);  // -v_sigma
                        nvs.x = gate0 ? nvs.x : 0.f;
                        nvs.y = gate1 ? nvs.y : 0.f;
                        if (ABS || !GSB_BWD_MOMENTS) {
                            const float2 hs = __fmul2_rn(nvs, bc(0.5f));
                            const float2 ga = __fmul2_rn(__fmul2_rn(hs, bc(dx)), bc(dx));
                            const float2 gb = __fmul2_rn(__fmul2_rn(nvs, bc(dx)), dy);
                            const float2 gc = __fmul2_rn(__fmul2_rn(hs, dy), dy);
                            const float2 gx = __fmul2_rn(nvs, __ffma2_rn(bc(B.y), dy, bc(u)));
                            const float2 gy = __fmul2_rn(nvs, __ffma2_rn(bc(B.z), dy, bc(v)));
                            acc<NP == 1>(vals[CDIM + 0], -ga.x - ga.y);
                            acc<NP == 1>(vals[CDIM + 1], -gb.x - gb.y);
                            acc<NP == 1>(vals[CDIM + 2], -gc.x - gc.y);
                            acc<NP == 1>(vals[CDIM + 3], -gx.x - gx.y);
                            acc<NP == 1>(vals[CDIM + 4], -gy.x - gy.y);
                            const float2 go = __fmul2_rn(vis, v_alpha);
                            acc<NP == 1>(tail[0], (gate0 ? go.x : 0.f) + (gate1 ? go.y : 0.f));
                            if (ABS) {
                                acc<NP == 1>(tail[1], fabsf(gx.x) + fabsf(gx.y));
                                acc<NP == 1>(tail[2], fabsf(gy.x) + fabsf(gy.y));
                            }
                        } else {
                            // moment form: every geometric gradient is linear in the five moments of w = -v_sigma about the
                            // Gaussian's centre, sum w {dx, dy, dx^2, dx dy, dy^2}, and v_opacity = sum w / opacity (w = opacity
                            // * vis * v_alpha under the same gate).  The record accumulates the moments; the linear map to
                            // (v_conic, v_xy, v_opacity) runs once per Gaussian in unpack_grads_kernel.  dx is shared by the
                            // thread's two pixels: 2 packed + 6 scalar operations instead of 12 packed + 6 scalar.
                            const float n01 = nvs.x + nvs.y;
                            const float2 nd = __fmul2_rn(nvs, dy);
                            const float2 ndd = __fmul2_rn(nd, dy);
                            const float sy = nd.x + nd.y;
                            const float sx = dx * n01;
                            acc<NP == 1>(vals[CDIM + 0], dx * sx);
                            acc<NP == 1>(vals[CDIM + 1], dx * sy);
                            acc<NP == 1>(vals[CDIM + 2], ndd.x + ndd.y);
                            acc<NP == 1>(vals[CDIM + 3], sx);
                            acc<NP == 1>(vals[CDIM + 4], sy);
                            acc<NP == 1>(tail[0], n01);
                        }
                    }
                    if (!any_valid) continue;
                    float *rec = grad_rec + (size_t)__float_as_int(B.w) * REC;
                    if (GSB_BWD_PIPELINE) {
                        pend_rec = rec;
                    } else {
                        retire(vals, tail, rec, true);
                    }
                }
            }
        }
        // release the slot; the last warp to do so refills it with the stage kStages further along the walk.  The warp's
        // shared-memory reads of the slot have been consumed (their values are in registers) before this point.
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&S.count[slot], 1) == (int)NCONS - 1) {
                S.count[slot] = 0;
                if (j + kStages < nst) issue(j + kStages);
            }
        }
    }
    if (GSB_BWD_PIPELINE && pend_rec != nullptr) retire(pend, pend_tail, pend_rec, true);  // the last record of the walk
}

// ------------------------------------------------------------------------------------------------ host
// pixel pairs per consumer thread: 1 (four 8x8 warps per tile) or 2 (two 16x8 warps); GSB200_STREAM_NP overrides
#ifndef GSB_STREAM_NP
#define GSB_STREAM_NP 1
#endif
// Resident CTAs per SM can be capped from the environment (A/B experiments on the long-tile tail): GSB200_SFWD_CTAS /
// GSB200_SBWD_CTAS = n makes every CTA ask for 1/n of the SM's shared memory as (unused) dynamic shared memory.
static size_t occupancy_pad(const char *env, size_t static_bytes, int default_ctas = 0) {
    const char *v = getenv(env);
    const int n = v != nullptr ? atoi(v) : default_ctas;
    if (n < 1 || n > 16) return 0;
    const size_t per_cta = (size_t)(227 * 1024) / (size_t)n - 1024;  // 1 KB per CTA is reserved by the runtime
    return per_cta > static_bytes ? per_cta - static_bytes : 0;
}

static int stream_np() {
    static const int np = [] {
        const char *v = getenv("GSB200_STREAM_NP");
        return (v != nullptr && (v[0] == '1' || v[0] == '2')) ? v[0] - '0' : GSB_STREAM_NP;
    }();
    return np;
}

bool blend_stream_bwd_moments(bool absgrad) { return GSB_BWD_MOMENTS && !absgrad; }

bool blend_stream_supported(uint32_t tile_size, uint32_t cdim, int64_t n_isects) {
    return tile_size == 16 && cdim >= 1 && cdim <= 4 && n_isects < (int64_t(1) << 28);
}

template <int CDIM>
static int fwd_t(const BlendGeom &G, const float *means2d, const float *conics, const float *colors, const float *opacities,
                 const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets, const int32_t *flatten_ids,
                 float *render_colors, float *render_alphas, int32_t *last_ids, int32_t *tile_order, float *blend_stream,
                 int32_t *tile_counts, cudaStream_t s) {
    const uint32_t T = G.C * G.tw * G.th;
    if (tile_order != nullptr)
        if (int e = launch_tile_order(T, G.n_isects, tile_offsets, tile_order, s)) return e;
    if (stream_np() == 2)
        blend_stream_fwd_kernel<CDIM, 2><<<T, 3 * 32, 0, s>>>(
            G, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, backgrounds, masks, tile_offsets, flatten_ids,
            render_colors, render_alphas, last_ids, tile_order, reinterpret_cast<float4 *>(blend_stream), tile_counts);
    else {
        const size_t pad = occupancy_pad("GSB200_SFWD_CTAS", sizeof(StreamShared));
        if (pad > 0)
            GSB_CUDA_OK(cudaFuncSetAttribute(blend_stream_fwd_kernel<CDIM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
        blend_stream_fwd_kernel<CDIM, 1><<<T, 5 * 32, pad, s>>>(
            G, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, backgrounds, masks, tile_offsets, flatten_ids,
            render_colors, render_alphas, last_ids, tile_order, reinterpret_cast<float4 *>(blend_stream), tile_counts);
    }
    return check_launch("blend_stream_fwd");
}

int blend_stream_fwd(const BlendGeom &G, uint32_t cdim, const float *means2d, const float *conics, const float *colors,
                     const float *opacities, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                     const int32_t *flatten_ids, float *render_colors, float *render_alphas, int32_t *last_ids,
                     int32_t *tile_order, float *blend_stream, int32_t *tile_counts, cudaStream_t s) {
#define GSB_FWD_CASE(D)                                                                                                \
    case D:                                                                                                            \
        return fwd_t<D>(G, means2d, conics, colors, opacities, backgrounds, masks, tile_offsets, flatten_ids, render_colors, \
                        render_alphas, last_ids, tile_order, blend_stream, tile_counts, s);
    switch (cdim) {
        GSB_FWD_CASE(1) GSB_FWD_CASE(2) GSB_FWD_CASE(3) GSB_FWD_CASE(4)
        default: return fail(2, "gsb200_rasterize: stream blend supports 1..4 channels");
    }
#undef GSB_FWD_CASE
}

template <int CDIM>
static int bwd_t(const BlendGeom &G, const float *backgrounds, const uint8_t *masks, const int32_t *tile_offsets,
                 const float *render_alphas, const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                 const int32_t *tile_order, const float *blend_stream, const int32_t *tile_counts, float *grad_scratch,
                 bool absgrad, cudaStream_t s) {
    const uint32_t T = G.C * G.tw * G.th;
    const float4 *st = reinterpret_cast<const float4 *>(blend_stream);
    // measured on B200 at the headline size.  One record per trip (tools/gpu_r2_visit29.sh): 8 resident CTAs 667 us, 6: 663,
    // 5: 650, 4: 691, 3: 823 -- the long tiles' serial chains wanted fewer co-residents.  Two records per trip
    // (tools/gpu_r2_visit34.sh): 5: 577 us, 6: 568, 8: 566 -- the chains are half as long, no cap
    const size_t pad = occupancy_pad("GSB200_SBWD_CTAS", sizeof(StreamShared), GSB_BWD_PAIR ? 0 : 5);
#define GSB_BWD_LAUNCH(NPV, ABSV)                                                                                     \
    do {                                                                                                              \
        if (pad > 0)                                                                                                  \
            GSB_CUDA_OK(cudaFuncSetAttribute(blend_stream_bwd_kernel<CDIM, NPV, ABSV>,                                \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));                 \
        blend_stream_bwd_kernel<CDIM, NPV, ABSV><<<T, (4 / NPV) * 32, pad, s>>>(G, backgrounds, masks, tile_offsets,  \
                                                                             render_alphas, last_ids, v_render_colors, \
                                                                             v_render_alphas, tile_order, st,         \
                                                                             tile_counts, grad_scratch);              \
    } while (0)
    if (stream_np() == 2) {
        if (absgrad) GSB_BWD_LAUNCH(2, true); else GSB_BWD_LAUNCH(2, false);
    } else {
        if (absgrad) GSB_BWD_LAUNCH(1, true); else GSB_BWD_LAUNCH(1, false);
    }
#undef GSB_BWD_LAUNCH
    return check_launch("blend_stream_bwd");
}

int blend_stream_bwd(const BlendGeom &G, uint32_t cdim, uint32_t M, const float *backgrounds, const uint8_t *masks,
                     const int32_t *tile_offsets, const float *render_alphas, const int32_t *last_ids,
                     const float *v_render_colors, const float *v_render_alphas, const int32_t *tile_order,
                     const float *blend_stream, const int32_t *tile_counts, float *grad_scratch, bool absgrad, cudaStream_t s) {
    (void)M;
#define GSB_BWD_CASE(D)                                                                                                  \
    case D:                                                                                                              \
        return bwd_t<D>(G, backgrounds, masks, tile_offsets, render_alphas, last_ids, v_render_colors, v_render_alphas,   \
                        tile_order, blend_stream, tile_counts, grad_scratch, absgrad, s);
    switch (cdim) {
        GSB_BWD_CASE(1) GSB_BWD_CASE(2) GSB_BWD_CASE(3) GSB_BWD_CASE(4)
        default: return fail(2, "gsb200_rasterize: stream blend supports 1..4 channels");
    }
#undef GSB_BWD_CASE
}

}  // namespace gsb200
