// Stage 2: Gaussian/tile intersection, per-tile depth sort and tile offsets (sm_100a).
//
// Replaces _C.isect_tiles + _C.isect_offset_encode (gsplat/cuda/csrc/isect_tiles.cu:16-389).  The
// reference emits (camera | tile | depth) 64-bit keys in Gaussian order and sorts ALL of them with
// one global cub::DeviceRadixSort (6 digit passes over 12 B/intersection).  Here the structure of
// the key is used instead:
//
//   count    : per element, tile bbox -> tiles_per_gauss, and a histogram of intersections per
//              (camera, tile) with one RED per covered tile
//   scan     : exclusive scan of the histogram -> isect_offsets (the reference needs a separate
//              isect_offset_encode pass over the sorted keys for this)
//   scatter  : every intersection claims a slot inside its tile's segment (atomic cursor) and
//              stores (depth bits << 32 | flatten id)
//   tile sort: one CTA per (camera, tile) sorts its segment in shared memory with a normalized
//              bitonic network on the unique 64-bit (depth bits, flatten id) key and writes the
//              reference's isect_ids / flatten_ids.
//
// Ordering contract: the reference's stable LSD radix sort keeps emission order (ascending flatten
// index) among equal keys, i.e. it sorts by (camera, tile, depth bits, flatten id).  That composite
// key is unique, so any correct sort of it reproduces the reference bit for bit.
#include <stdlib.h>
#include <string.h>

#include "isect.cuh"

namespace gsb200 {

constexpr uint32_t kSortSmall = 2048;   // keys a 256-thread CTA sorts in 16 KB of static shared memory
constexpr uint32_t kSortLarge = 12288;  // keys a 1024-thread CTA sorts in 96 KB of dynamic shared memory

__global__ void __launch_bounds__(kThreads)
isect_count_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                   const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th,
                   uint32_t total_tiles, int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ tile_counts) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    TileBox b = {0, 0, 0, 0};
    uint32_t cam = 0;
    bool has = false;
    if (i < n) {
        const int32_t r = radii[i];
        int32_t tiles = 0;
        if (r > 0) {
            const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
            b = tile_box(m.x, m.y, (float)r, ts, tw, th);
            // the reference multiplies the two uint32 extents (isect_tiles.cu:73-76)
            tiles = (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0));
            has = b.x1 > b.x0 && b.y1 > b.y0;
            cam = camera_ids ? (uint32_t)camera_ids[i] : i / per_cam;
        }
        tiles_per_gauss[i] = tiles;
    }
    // counters are replicated GSB200_ISECT_REPLICAS times (replica = block index mod R) so that the
    // ~500 increments every tile receives are spread over R cache lines instead of hammering one
    const uint32_t n_tiles = tw * th;
    int32_t *mine = tile_counts + (size_t)(blockIdx.x % GSB200_ISECT_REPLICAS) * total_tiles;
    for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
        const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
        if (on) atomicAdd(mine + (size_t)c * n_tiles + tile, 1);
    });
}

// per-tile total over the replicas (input of the exclusive scan that yields isect_offsets)
__global__ void __launch_bounds__(kThreads)
replica_totals_kernel(uint32_t total_tiles, const int32_t *__restrict__ counts, int32_t *__restrict__ totals) {
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total_tiles) return;
    int sum = 0;
#pragma unroll 8
    for (uint32_t r = 0; r < GSB200_ISECT_REPLICAS; ++r) sum += counts[(size_t)r * total_tiles + t];
    totals[t] = sum;
}

// counts[r][t] -> first slot of replica r inside tile t's segment (in place); tiles too long for the
// small sort kernel are appended to a work list (counter at counts[R*T], ids behind it)
__global__ void __launch_bounds__(kThreads)
replica_cursor_kernel(uint32_t total_tiles, const int32_t *__restrict__ offsets, int32_t *__restrict__ counts,
                      uint32_t large_threshold) {
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total_tiles) return;
    int run = offsets[t];
    const int first = run;
#pragma unroll 8
    for (uint32_t r = 0; r < GSB200_ISECT_REPLICAS; ++r) {
        const int c = counts[(size_t)r * total_tiles + t];
        counts[(size_t)r * total_tiles + t] = run;
        run += c;
    }
    if ((uint32_t)(run - first) > large_threshold) {
        int32_t *list = counts + (size_t)GSB200_ISECT_REPLICAS * total_tiles;
        list[1 + atomicAdd(list, 1)] = (int32_t)t;
    }
}

// ---- the three steps above (replica totals -> exclusive scan over the tiles -> replica cursors) in TWO launches ------------
// One warp per tile, lane = replica: the 32 counters of a tile are one strided warp load, their total one warp reduction,
// their exclusive prefix one warp scan.  A block covers 32 consecutive tiles.  Kernel 1 leaves the block totals, kernel 2
// re-derives its own prefix from them (a block reduction over the <= few hundred lower blocks -- no look-back flags, no
// third kernel), scans its 32 tiles and writes offsets, cursors, the long-tile work list and the grand total.  The five
// launches this replaces were latency chains on 1 MB of counters: 36 us at 8160 tiles (profiles/r02_isect_bucket_sort_ncu.md).
static_assert(GSB200_ISECT_REPLICAS == 32, "replica_block_* kernels map one replica to one lane");
constexpr int kTilesPerBlock = 32;

__global__ void __launch_bounds__(1024)
replica_block_totals_kernel(uint32_t total_tiles, const int32_t *__restrict__ counts, int32_t *__restrict__ block_totals) {
    __shared__ int s_tot[kTilesPerBlock];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * kTilesPerBlock + warp;
    const int c = t < total_tiles ? counts[(size_t)lane * total_tiles + t] : 0;
    const int tile_total = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_tot[warp] = tile_total;
    __syncthreads();
    if (warp == 0) {
        const int b = __reduce_add_sync(0xffffffffu, s_tot[lane]);
        if (lane == 0) block_totals[blockIdx.x] = b;
    }
}

__global__ void __launch_bounds__(1024)
replica_block_cursor_kernel(uint32_t total_tiles, const int32_t *__restrict__ block_totals, int32_t *__restrict__ counts,
                            int32_t *__restrict__ offsets, int32_t *__restrict__ total_out, uint32_t large_threshold) {
    __shared__ int s_tot[kTilesPerBlock];
    __shared__ int s_part[32];
    __shared__ int s_prefix;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // exclusive prefix of this block: sum of the totals of all lower blocks
    int part = 0;
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += 1024) part += block_totals[j];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_part[warp] = part;
    const uint32_t t = blockIdx.x * kTilesPerBlock + warp;
    const int c = t < total_tiles ? counts[(size_t)lane * total_tiles + t] : 0;
    int incl = c;  // inclusive scan over the replicas of this tile
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    const int tile_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) s_tot[warp] = tile_total;
    __syncthreads();
    if (warp == 0) {
        const int prefix = __reduce_add_sync(0xffffffffu, s_part[lane]);
        const int v = s_tot[lane];
        int wi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)lane >= o) wi += u;
        }
        s_tot[lane] = wi - v;  // exclusive prefix of tile `lane` inside the block
        if (lane == 31) {
            s_prefix = prefix;
            if (blockIdx.x == gridDim.x - 1) *total_out = prefix + wi;
        }
    }
    __syncthreads();
    if (t >= total_tiles) return;
    const int base = s_prefix + s_tot[warp];
    counts[(size_t)lane * total_tiles + t] = base + incl - c;  // first slot of replica `lane` inside the tile's segment
    if (lane == 0) {
        offsets[t] = base;
        if ((uint32_t)tile_total > large_threshold) {
            int32_t *list = counts + (size_t)GSB200_ISECT_REPLICAS * total_tiles;
            list[1 + atomicAdd(list, 1)] = (int32_t)t;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
isect_scatter_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                     const int32_t *__restrict__ radii, const float *__restrict__ depths, float ts, uint32_t tw, uint32_t th,
                     uint32_t total_tiles, int32_t *__restrict__ tile_cursor, uint64_t *__restrict__ bucket,
                     int64_t capacity) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    TileBox b = {0, 0, 0, 0};
    uint32_t cam = 0, depth_bits = 0;
    bool has = false;
    if (i < n) {
        const int32_t r = radii[i];
        if (r > 0) {
            const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
            b = tile_box(m.x, m.y, (float)r, ts, tw, th);
            has = b.x1 > b.x0 && b.y1 > b.y0;
            cam = camera_ids ? (uint32_t)camera_ids[i] : i / per_cam;
            depth_bits = __float_as_uint(depths[i]);
        }
    }
    const uint32_t n_tiles = tw * th;
    int32_t *mine = tile_cursor + (size_t)(blockIdx.x % GSB200_ISECT_REPLICAS) * total_tiles;  // same replica as in the count pass
    for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
        const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
        const uint32_t d = __shfl_sync(0xffffffffu, depth_bits, src);
        const uint32_t id = __shfl_sync(0xffffffffu, i, src);
        if (on) {
            const int32_t slot = atomicAdd(mine + (size_t)c * n_tiles + tile, 1);
            if (slot < capacity) bucket[slot] = ((uint64_t)d << 32) | (uint64_t)id;
        }
    });
}

// ---- CTA-private tile histograms (opt-in, where the histogram of ALL tiles fits a CTA's shared memory) ---------------
// The two kernels above are bound by L2 atomics: one RED per intersection in the count pass and one RETURNING atomic per
// intersection in the scatter pass (32 + 75 us at the headline size, I = 4 M; 80 MB of traffic would take 12 us).  Here a
// CTA owns the same Gaussians in BOTH passes and keeps a histogram of all C x tiles counters in shared
// memory:
//   count  : shared-memory atomics; the CTA's row is stored to global memory (coalesced) and its NON-ZERO counts are added
//            to the replica (CTA mod 32) of the global histogram -- the scan kernels downstream are unchanged;
//   scatter: the CTA reloads its row and claims a contiguous piece of every tile it touches with ONE returning atomic per
//            (CTA, tile) pair, then hands out the slots of that piece through shared-memory atomics.
// Global atomics per pass: the number of non-zero (CTA, tile) pairs instead of I.  Which CTA's piece comes first inside a
// tile is not deterministic, like the slot order before: the per-tile sort that follows orders the unique (depth, id) keys.
constexpr int kPrivThreads = 512;
constexpr uint32_t kPrivMaxTiles = 24576;  // 96 KB of counters: two CTAs per SM

// Which Gaussians a CTA owns (the same in both passes): groups of 32 consecutive ones -- the tile walk is warp-cooperative --
// dealt round-robin to the CTAs and then to their warps.  Contiguous chunks were measured first and were SLOWER than the
// global-atomic kernels (count 62 -> 128 us): neighbours in memory are neighbours in space, so a few CTAs owned all the
// large foreground Gaussians.
#define GSB_PRIVATE_GROUPS(j, n) \
    for (uint32_t j = blockIdx.x + gridDim.x * (threadIdx.x >> 5); j * 32u < (n); j += gridDim.x * (kPrivThreads / 32))

__global__ void __launch_bounds__(kPrivThreads)
isect_count_private_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                           const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th, uint32_t total_tiles,
                           int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ tile_counts, int32_t *__restrict__ rows) {
    extern __shared__ int32_t s_hist[];
    for (uint32_t t = threadIdx.x; t < total_tiles; t += kPrivThreads) s_hist[t] = 0;
    __syncthreads();
    const uint32_t n_tiles = tw * th, end = n;
    GSB_PRIVATE_GROUPS(j, n) {
        const uint32_t i = j * 32u + (threadIdx.x & 31u);
        TileBox b = {0, 0, 0, 0};
        uint32_t cam = 0;
        bool has = false;
        if (i < end) {
            const int32_t r = radii[i];
            int32_t tiles = 0;
            if (r > 0) {
                const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
                b = tile_box(m.x, m.y, (float)r, ts, tw, th);
                tiles = (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0));  // (isect_tiles.cu:73-76)
                has = b.x1 > b.x0 && b.y1 > b.y0;
                cam = camera_ids ? (uint32_t)camera_ids[i] : i / per_cam;
            }
            tiles_per_gauss[i] = tiles;
        }
        for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
            const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
            if (on) atomicAdd(&s_hist[c * n_tiles + tile], 1);
        });
    }
    __syncthreads();
    int32_t *mine = tile_counts + (size_t)(blockIdx.x % GSB200_ISECT_REPLICAS) * total_tiles;
    int32_t *row = rows + (size_t)blockIdx.x * total_tiles;
    for (uint32_t t = threadIdx.x; t < total_tiles; t += kPrivThreads) {
        const int32_t c = s_hist[t];
        row[t] = c;
        if (c != 0) atomicAdd(mine + t, c);
    }
}

__global__ void __launch_bounds__(kPrivThreads)
isect_scatter_private_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                             const int32_t *__restrict__ radii, const float *__restrict__ depths, float ts, uint32_t tw, uint32_t th,
                             uint32_t total_tiles, int32_t *__restrict__ tile_cursor, const int32_t *__restrict__ rows,
                             uint64_t *__restrict__ bucket, int64_t capacity) {
    extern __shared__ int32_t s_hist[];  // here: the next free slot of this CTA's piece of every tile
    int32_t *mine = tile_cursor + (size_t)(blockIdx.x % GSB200_ISECT_REPLICAS) * total_tiles;  // same replica as in the count pass
    const int32_t *row = rows + (size_t)blockIdx.x * total_tiles;
    for (uint32_t t = threadIdx.x; t < total_tiles; t += kPrivThreads) {
        const int32_t c = row[t];
        s_hist[t] = c != 0 ? atomicAdd(mine + t, c) : 0;
    }
    __syncthreads();
    const uint32_t n_tiles = tw * th, end = n;
    GSB_PRIVATE_GROUPS(j, n) {
        const uint32_t i = j * 32u + (threadIdx.x & 31u);
        TileBox b = {0, 0, 0, 0};
        uint32_t cam = 0, depth_bits = 0;
        bool has = false;
        if (i < end) {
            const int32_t r = radii[i];
            if (r > 0) {
                const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
                b = tile_box(m.x, m.y, (float)r, ts, tw, th);
                has = b.x1 > b.x0 && b.y1 > b.y0;
                cam = camera_ids ? (uint32_t)camera_ids[i] : i / per_cam;
                depth_bits = __float_as_uint(depths[i]);
            }
        }
        for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
            const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
            const uint32_t d = __shfl_sync(0xffffffffu, depth_bits, src);
            const uint32_t id = __shfl_sync(0xffffffffu, i, src);
            if (on) {
                const int32_t slot = atomicAdd(&s_hist[c * n_tiles + tile], 1);
                if (slot < capacity) bucket[slot] = ((uint64_t)d << 32) | (uint64_t)id;
            }
        });
    }
}

// ---- per-tile sort: normalized bitonic network, 64-bit unique keys ------------------------------
// Every comparison moves the smaller key to the lower index, so +inf padding beyond n never moves.
// Sub-steps whose partner distance is < 64 stay inside aligned 64-element blocks; a warp keeps such a
// block in registers (2 keys per lane) and exchanges through shuffles, so only the long-distance
// sub-steps (k >= 128) go through shared memory with a block barrier: 6 barriers for 512 keys
// instead of 45.
constexpr uint64_t kInf = ~0ull;

__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_idx64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void keep(uint64_t &x, uint64_t p, bool lower) { x = lower ? min(x, p) : max(x, p); }

// partner distance j < 32 inside both 32-key halves held as (x0 = key[lane], x1 = key[32 + lane])
__device__ __forceinline__ void warp_steps_below32(uint64_t &x0, uint64_t &x1, uint32_t lane, int j_first) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        if (j > j_first) continue;
        const bool lower = (lane & j) == 0;
        keep(x0, shfl_xor64(x0, j), lower);
        keep(x1, shfl_xor64(x1, j), lower);
    }
}

// full sort of the 64 keys of one block
__device__ __forceinline__ void warp_sort64(uint64_t &x0, uint64_t &x1, uint32_t lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        const bool lower = (lane & (k >> 1)) == 0;  // mirror step: partner = lane ^ (k - 1)
        keep(x0, shfl_xor64(x0, k - 1), lower);
        keep(x1, shfl_xor64(x1, k - 1), lower);
        warp_steps_below32(x0, x1, lane, k >> 2);
    }
    // k = 64 mirror: key i pairs with key 63 - i, which lives in the other half of lane 31 - lane
    const uint64_t p0 = shfl_idx64(x1, 31 - (int)lane), p1 = shfl_idx64(x0, 31 - (int)lane);
    x0 = min(x0, p0);
    x1 = max(x1, p1);
    warp_steps_below32(x0, x1, lane, 16);
}

// tail of a merge (partner distances 32, 16, ..., 1) of one 64-key block
__device__ __forceinline__ void warp_merge_tail64(uint64_t &x0, uint64_t &x1, uint32_t lane) {
    const uint64_t lo = min(x0, x1), hi = max(x0, x1);
    x0 = lo;
    x1 = hi;
    warp_steps_below32(x0, x1, lane, 16);
}

// one long-distance sub-step over a[0..n_pad) in shared or global memory (all distances are powers of two)
__device__ __forceinline__ void block_substep(uint64_t *a, uint32_t n_pad, uint32_t half_pairs, uint32_t log_k, uint32_t j,
                                              bool mirror) {
    for (uint32_t t = threadIdx.x; t < half_pairs; t += blockDim.x) {
        uint32_t lo, hi;
        if (mirror) {
            const uint32_t log_hk = log_k - 1, blk = t >> log_hk, r = t & ((1u << log_hk) - 1u);
            lo = (blk << log_k) + r;
            hi = (blk << log_k) + ((1u << log_k) - 1u - r);
        } else {
            lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            hi = lo | j;
        }
        if (hi < n_pad) {
            const uint64_t x = a[lo], y = a[hi];
            if (x > y) { a[lo] = y; a[hi] = x; }
        }
    }
}

// sort a[0..n_pad), n_pad a multiple of 64, entries in [n, n_pad) hold kInf
__device__ __forceinline__ void block_sort_padded(uint64_t *a, uint32_t n_pad) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t nblk = n_pad >> 6;
    for (uint32_t blk = warp; blk < nblk; blk += nwarps) {
        uint64_t x0 = a[blk * 64 + lane], x1 = a[blk * 64 + 32 + lane];
        warp_sort64(x0, x1, lane);
        a[blk * 64 + lane] = x0;
        a[blk * 64 + 32 + lane] = x1;
    }
    uint32_t P = 64, log_p = 6;
    while (P < n_pad) { P <<= 1; ++log_p; }
    const uint32_t half_pairs = P >> 1;
    for (uint32_t log_k = 7; log_k <= log_p; ++log_k) {
        __syncthreads();
        block_substep(a, n_pad, half_pairs, log_k, 0, true);
        for (uint32_t j = 1u << (log_k - 2); j >= 64; j >>= 1) {
            __syncthreads();
            block_substep(a, n_pad, half_pairs, log_k, j, false);
        }
        __syncthreads();
        for (uint32_t blk = warp; blk < nblk; blk += nwarps) {
            uint64_t x0 = a[blk * 64 + lane], x1 = a[blk * 64 + 32 + lane];
            warp_merge_tail64(x0, x1, lane);
            a[blk * 64 + lane] = x0;
            a[blk * 64 + 32 + lane] = x1;
        }
    }
    __syncthreads();
}

// generic in-place variant for oversized segments in global memory (rare): same network, no padding
__device__ __forceinline__ void block_sort_global(uint64_t *a, uint32_t n) {
    uint32_t P = 1, log_p = 0;
    while (P < n) { P <<= 1; ++log_p; }
    const uint32_t half_pairs = P >> 1;
    for (uint32_t log_k = 1; log_k <= log_p; ++log_k) {
        block_substep(a, n, half_pairs, log_k, 0, true);
        __syncthreads();
        for (uint32_t j = log_k >= 2 ? 1u << (log_k - 2) : 0; j > 0; j >>= 1) {
            block_substep(a, n, half_pairs, log_k, j, false);
            __syncthreads();
        }
    }
}

// One CTA per (camera, tile).  LARGE = false handles segments of up to kSortSmall keys, LARGE = true
// the (few) longer ones -- in shared memory up to kSortLarge keys, in place in global memory (L2
// resident) beyond that.  Both kernels are launched over all tiles and exit early on the other class.
template <bool LARGE>
__global__ void __launch_bounds__(LARGE ? 1024 : kThreads)
tile_sort_kernel(uint32_t n_tiles_total, uint32_t n_tiles, uint32_t tile_bits, int64_t n_isects_or_capacity,
                 const int32_t *__restrict__ total_dev, const int32_t *__restrict__ offsets,
                 const int32_t *__restrict__ large_list, uint64_t *__restrict__ bucket, int64_t *__restrict__ isect_ids,
                 int32_t *__restrict__ flatten_ids) {
    extern __shared__ uint64_t s_dyn[];
    // exact mode: the host knows the count; capacity mode: the count lives on the device and segments that
    // do not fit the (speculatively sized) buffers are skipped -- the host re-runs the fill in that case
    const int64_t n_isects = total_dev ? (int64_t)*total_dev : n_isects_or_capacity;
    const int64_t capacity = n_isects_or_capacity;
    __shared__ uint64_t s_static[LARGE ? 1 : kSortSmall];
    // small: one CTA per tile; large: a few persistent CTAs walk the work list written by replica_cursor_kernel
    const uint32_t n_work = LARGE ? (uint32_t)large_list[0] : n_tiles_total;
    for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const uint32_t ct = LARGE ? (uint32_t)large_list[1 + w] : w;
        const int64_t start = offsets[ct];
        const int64_t end = (ct + 1 == n_tiles_total) ? n_isects : (int64_t)offsets[ct + 1];
        const uint32_t n = (uint32_t)(end - start);
        if (n == 0 || (!LARGE && n > kSortSmall) || end > capacity) continue;
        const uint64_t cam = ct / n_tiles, tile = ct % n_tiles;
        const int64_t hi = (int64_t)((cam << (32 + tile_bits)) | (tile << 32));
        uint64_t *a;
        if (!LARGE || n <= kSortLarge) {
            a = LARGE ? s_dyn : s_static;
            const uint32_t n_pad = (n + 63u) & ~63u;
            for (uint32_t t = threadIdx.x; t < n_pad; t += blockDim.x) a[t] = t < n ? bucket[start + t] : kInf;
            __syncthreads();
            if (n > 1) block_sort_padded(a, n_pad);
        } else {
            a = bucket + start;
            __syncthreads();
            block_sort_global(a, n);
        }
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
            const uint64_t it = a[t];
            // low word: sign-extending OR like the reference's (int64)(int32)depth cast (isect_tiles.cu:88)
            isect_ids[start + t] = hi | (int64_t)(int32_t)(uint32_t)(it >> 32);
            flatten_ids[start + t] = (int32_t)(uint32_t)(it & 0xffffffffu);
        }
        if (LARGE) __syncthreads();  // the staging buffer is reused by the next work item
    }
}

// ---- default short-segment sort: per-tile depth buckets + rank by counting ------------------------------------------------
// The bitonic network above spends ~860 lane-instructions per key.  A tile's keys are (depth bits, flatten id) with depths
// spread between the tile's own nearest and farthest Gaussian, so a linear map of the depth onto ~n/4 buckets (monotone in
// the depth, hence in the key) leaves a handful of keys per bucket: histogram -> scan -> scatter in shared memory, then
// every key counts the smaller keys of its own bucket.  Order = bucket major, exact 64-bit key inside a bucket = the same
// ascending key order as a full sort.  A segment whose fullest bucket exceeds kBucketBail (all depths equal, one outlier
// stretching the range) falls back to the bitonic network; segments of <= 64 keys are sorted by one warp in registers.
constexpr uint32_t kBucketMax = kSortSmall / 4;
constexpr uint32_t kBucketBail = 160;

__device__ __forceinline__ uint32_t depth_bucket(uint64_t key, float d_min, float scale, uint32_t n_buckets) {
    const float f = __fmul_rn(__fsub_rn(__uint_as_float((uint32_t)(key >> 32)), d_min), scale);
    return min((uint32_t)f, n_buckets - 1u);  // f >= 0; float -> uint conversion saturates
}

__device__ __forceinline__ void write_sorted(int64_t hi, int64_t pos, uint64_t key, int64_t *__restrict__ isect_ids,
                                             int32_t *__restrict__ flatten_ids) {
    // low word: sign-extending OR like the reference's (int64)(int32)depth cast (isect_tiles.cu:88)
    isect_ids[pos] = hi | (int64_t)(int32_t)(uint32_t)(key >> 32);
    flatten_ids[pos] = (int32_t)(uint32_t)(key & 0xffffffffu);
}

__global__ void __launch_bounds__(kThreads)
tile_sort_bucket_kernel(uint32_t n_tiles_total, uint32_t n_tiles, uint32_t tile_bits, int64_t n_isects_or_capacity,
                        const int32_t *__restrict__ total_dev, const int32_t *__restrict__ offsets,
                        const uint64_t *__restrict__ bucket, int64_t *__restrict__ isect_ids,
                        int32_t *__restrict__ flatten_ids) {
    constexpr uint32_t KPT = kSortSmall / kThreads, NW = kThreads / 32;
    __shared__ uint64_t s_keys[kSortSmall];
    __shared__ uint32_t s_base[kBucketMax + 1];
    __shared__ uint32_t s_min[NW], s_max[NW], s_sum[NW], s_full[NW];
    const int64_t n_isects = total_dev ? (int64_t)*total_dev : n_isects_or_capacity;
    const int64_t capacity = n_isects_or_capacity;
    const uint32_t ct = blockIdx.x;
    const int64_t start = offsets[ct];
    const int64_t end = (ct + 1 == n_tiles_total) ? n_isects : (int64_t)offsets[ct + 1];
    const uint32_t n = (uint32_t)(end - start);
    if (end <= start || n > kSortSmall || end > capacity) return;  // longer segments belong to tile_sort_kernel<true>
    const uint64_t cam = ct / n_tiles, tile = ct % n_tiles;
    const int64_t hi = (int64_t)((cam << (32 + tile_bits)) | (tile << 32));
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (n <= 64) {
        if (warp == 0) {
            uint64_t x0 = lane < n ? bucket[start + lane] : kInf, x1 = 32 + lane < n ? bucket[start + 32 + lane] : kInf;
            warp_sort64(x0, x1, lane);
            if (lane < n) write_sorted(hi, start + lane, x0, isect_ids, flatten_ids);
            if (32 + lane < n) write_sorted(hi, start + 32 + lane, x1, isect_ids, flatten_ids);
        }
        return;
    }
    uint32_t n_buckets = 16;
    while (n_buckets * 4 < n) n_buckets <<= 1;  // <= kBucketMax
    uint64_t key[KPT];
    uint32_t d_lo = 0xffffffffu, d_hi = 0;  // positive floats order like their bit patterns
#pragma unroll
    for (uint32_t k = 0; k < KPT; ++k) {
        const uint32_t i = threadIdx.x + k * kThreads;
        key[k] = i < n ? bucket[start + i] : kInf;
        if (i < n) {
            d_lo = min(d_lo, (uint32_t)(key[k] >> 32));
            d_hi = max(d_hi, (uint32_t)(key[k] >> 32));
        }
    }
    d_lo = __reduce_min_sync(0xffffffffu, d_lo);
    d_hi = __reduce_max_sync(0xffffffffu, d_hi);
    if (lane == 0) s_min[warp] = d_lo, s_max[warp] = d_hi;
    for (uint32_t t = threadIdx.x; t <= n_buckets; t += kThreads) s_base[t] = 0;
    __syncthreads();
#pragma unroll
    for (uint32_t w = 0; w < NW; ++w) d_lo = min(d_lo, s_min[w]), d_hi = max(d_hi, s_max[w]);
    const float d_min = __uint_as_float(d_lo);
    float scale = d_hi > d_lo ? __fdiv_rn((float)n_buckets, __fsub_rn(__uint_as_float(d_hi), d_min)) : 0.f;
    if (!(scale <= 3.0e38f)) scale = 0.f;  // inf / nan: one bucket, the bail-out below sorts it
    uint32_t where[KPT];  // bucket << 16 | arrival slot inside the bucket
#pragma unroll
    for (uint32_t k = 0; k < KPT; ++k) {
        if (threadIdx.x + k * kThreads < n) {
            const uint32_t b = depth_bucket(key[k], d_min, scale, n_buckets);
            where[k] = b << 16 | atomicAdd(&s_base[b], 1u);
        }
    }
    __syncthreads();
    {   // exclusive scan of the bucket counts (n_buckets <= 2 * kThreads) and the fullest bucket
        const uint32_t per = n_buckets > kThreads ? 2u : 1u, i0 = threadIdx.x * per;
        const uint32_t c0 = i0 < n_buckets ? s_base[i0] : 0u, c1 = per == 2 && i0 + 1 < n_buckets ? s_base[i0 + 1] : 0u;
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        const uint32_t full = __reduce_max_sync(0xffffffffu, max(c0, c1));
        if (lane == 31) s_sum[warp] = incl;
        if (lane == 0) s_full[warp] = full;
        __syncthreads();
        uint32_t before = 0, fullest = 0;
#pragma unroll
        for (uint32_t w = 0; w < NW; ++w) before += w < warp ? s_sum[w] : 0u, fullest = max(fullest, s_full[w]);
        const uint32_t excl = before + incl - c0 - c1;
        if (i0 < n_buckets) s_base[i0] = excl;
        if (per == 2 && i0 + 1 < n_buckets) s_base[i0 + 1] = excl + c0;
        if (threadIdx.x == 0) s_base[n_buckets] = n;
        if (fullest > kBucketBail) {  // block-uniform
            const uint32_t n_pad = (n + 63u) & ~63u;
#pragma unroll
            for (uint32_t k = 0; k < KPT; ++k)
                if (threadIdx.x + k * kThreads < n_pad) s_keys[threadIdx.x + k * kThreads] = key[k];
            __syncthreads();
            block_sort_padded(s_keys, n_pad);
            for (uint32_t t = threadIdx.x; t < n; t += kThreads) write_sorted(hi, start + t, s_keys[t], isect_ids, flatten_ids);
            return;
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0; k < KPT; ++k)
        if (threadIdx.x + k * kThreads < n) s_keys[s_base[where[k] >> 16] + (where[k] & 0xffffu)] = key[k];
    __syncthreads();
    // consecutive lanes take consecutive bucket-ordered slots: neighbours share a bucket, its keys are broadcast reads
#pragma unroll 2
    for (uint32_t k = 0; k < KPT; ++k) {
        const uint32_t i = threadIdx.x + k * kThreads;
        if (i >= n) break;
        const uint64_t mine = s_keys[i];
        const uint32_t b = depth_bucket(mine, d_min, scale, n_buckets);
        const uint32_t lo = s_base[b], up = s_base[b + 1];
        uint32_t rank = lo;
        for (uint32_t j = lo; j < up; ++j) rank += s_keys[j] < mine ? 1u : 0u;
        write_sorted(hi, start + rank, mine, isect_ids, flatten_ids);
    }
}

// Long segments (> kSortSmall keys), same scheme with 1024 threads walking the work list of replica_cursor_kernel.  The
// keys are not held in registers: the three passes (depth range, histogram, scatter) re-read them from the L2-resident
// bucket array, and the scatter takes its slot from a running per-bucket cursor, which leaves s_base[b] = end of bucket b.
constexpr uint32_t kBucketMaxLarge = 4096;
constexpr size_t kBucketLargeSmem = kSortLarge * sizeof(uint64_t) + (kBucketMaxLarge + 1) * sizeof(uint32_t);

__global__ void __launch_bounds__(1024)
tile_sort_bucket_large_kernel(uint32_t n_tiles_total, uint32_t n_tiles, uint32_t tile_bits, int64_t n_isects_or_capacity,
                              const int32_t *__restrict__ total_dev, const int32_t *__restrict__ offsets,
                              const int32_t *__restrict__ large_list, uint64_t *__restrict__ bucket,
                              int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    constexpr uint32_t NT = 1024, NW = NT / 32;
    extern __shared__ uint64_t s_dyn[];
    uint64_t *s_keys = s_dyn;
    uint32_t *s_base = reinterpret_cast<uint32_t *>(s_dyn + kSortLarge);
    __shared__ uint32_t s_min[NW], s_max[NW], s_sum[NW], s_full[NW];
    const int64_t n_isects = total_dev ? (int64_t)*total_dev : n_isects_or_capacity;
    const int64_t capacity = n_isects_or_capacity;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_work = (uint32_t)large_list[0];
    for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const uint32_t ct = (uint32_t)large_list[1 + w];
        const int64_t start = offsets[ct];
        const int64_t end = (ct + 1 == n_tiles_total) ? n_isects : (int64_t)offsets[ct + 1];
        const uint32_t n = (uint32_t)(end - start);
        if (end <= start || end > capacity) continue;
        const uint64_t cam = ct / n_tiles, tile = ct % n_tiles;
        const int64_t hi = (int64_t)((cam << (32 + tile_bits)) | (tile << 32));
        uint64_t *seg = bucket + start;
        if (n > kSortLarge) {  // in place in global memory
            __syncthreads();
            block_sort_global(seg, n);
            for (uint32_t t = threadIdx.x; t < n; t += NT) write_sorted(hi, start + t, seg[t], isect_ids, flatten_ids);
            continue;
        }
        uint32_t n_buckets = NT;
        while (n_buckets * 4 < n) n_buckets <<= 1;  // <= kBucketMaxLarge
        uint32_t d_lo = 0xffffffffu, d_hi = 0;
        for (uint32_t i = threadIdx.x; i < n; i += NT) {
            const uint32_t d = (uint32_t)(seg[i] >> 32);
            d_lo = min(d_lo, d), d_hi = max(d_hi, d);
        }
        d_lo = __reduce_min_sync(0xffffffffu, d_lo);
        d_hi = __reduce_max_sync(0xffffffffu, d_hi);
        if (lane == 0) s_min[warp] = d_lo, s_max[warp] = d_hi;
        for (uint32_t t = threadIdx.x; t <= n_buckets; t += NT) s_base[t] = 0;
        __syncthreads();
#pragma unroll
        for (uint32_t q = 0; q < NW; ++q) d_lo = min(d_lo, s_min[q]), d_hi = max(d_hi, s_max[q]);
        const float d_min = __uint_as_float(d_lo);
        float scale = d_hi > d_lo ? __fdiv_rn((float)n_buckets, __fsub_rn(__uint_as_float(d_hi), d_min)) : 0.f;
        if (!(scale <= 3.0e38f)) scale = 0.f;
        for (uint32_t i = threadIdx.x; i < n; i += NT) atomicAdd(&s_base[depth_bucket(seg[i], d_min, scale, n_buckets)], 1u);
        __syncthreads();
        const uint32_t per = n_buckets / NT, i0 = threadIdx.x * per;  // 1, 2 or 4 counters per thread
        uint32_t c[4] = {0, 0, 0, 0}, mine = 0, full = 0;
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q)
            if (q < per) c[q] = s_base[i0 + q], mine += c[q], full = max(full, c[q]);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        full = __reduce_max_sync(0xffffffffu, full);
        if (lane == 31) s_sum[warp] = incl;
        if (lane == 0) s_full[warp] = full;
        __syncthreads();
        uint32_t run = incl - mine, fullest = 0;
#pragma unroll
        for (uint32_t q = 0; q < NW; ++q) run += q < warp ? s_sum[q] : 0u, fullest = max(fullest, s_full[q]);
        if (fullest > kBucketBail) {  // block-uniform: the bitonic network on the whole segment
            const uint32_t n_pad = (n + 63u) & ~63u;
            for (uint32_t t = threadIdx.x; t < n_pad; t += NT) s_keys[t] = t < n ? seg[t] : kInf;
            __syncthreads();
            block_sort_padded(s_keys, n_pad);
            for (uint32_t t = threadIdx.x; t < n; t += NT) write_sorted(hi, start + t, s_keys[t], isect_ids, flatten_ids);
            __syncthreads();
            continue;
        }
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q)
            if (q < per) s_base[i0 + q] = run, run += c[q];  // running cursor, starts at the bucket's first slot
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += NT) {
            const uint64_t key = seg[i];
            s_keys[atomicAdd(&s_base[depth_bucket(key, d_min, scale, n_buckets)], 1u)] = key;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += NT) {
            const uint64_t key = s_keys[i];
            const uint32_t b = depth_bucket(key, d_min, scale, n_buckets);
            const uint32_t lo = b ? s_base[b - 1] : 0u, up = s_base[b];
            uint32_t rank = lo;
            for (uint32_t j = lo; j < up; ++j) rank += s_keys[j] < key ? 1u : 0u;
            write_sorted(hi, start + rank, key, isect_ids, flatten_ids);
        }
        __syncthreads();  // shared memory is reused by the next work item
    }
}

// ---- EXPERIMENTAL alternative to tile_sort_kernel<false> (opt-in: GSB200_TILE_SORT=radix, not yet run on a GPU) -----------
// The bitonic network above spends ~860 lane-instructions per key (64-bit compare-exchange, 45-66 stages).  Here a CTA
// sorts its segment (<= kSortSmall keys) with a stable LSD radix sort on the depth word -- four 8-bit passes, passes in
// which all keys share the digit are skipped -- and then orders the (rare) runs of exactly equal depth by flatten id,
// which gives the same ascending (depth bits, flatten id) order.  Stability inside a pass comes from processing the
// keys in index order: warp w owns a contiguous chunk, walks it 32 keys at a time, and ranks equal digits inside a
// group with __match_any_sync (rank = number of lower lanes with the same digit).
constexpr uint32_t kRadixWarps = kThreads / 32;

__device__ __forceinline__ void block_sort_radix(uint64_t *a, uint64_t *b, uint32_t n, uint16_t (*hist)[256], uint32_t *s_warp_sums,
                                                 uint32_t *s_flag) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = ((n + kRadixWarps - 1) / kRadixWarps + 31u) & ~31u;  // keys per warp, whole groups of 32
    const uint32_t w_begin = min(n, warp * chunk), w_end = min(n, w_begin + chunk);
    uint64_t *src = a, *dst = b;
    if (threadIdx.x < 4) s_flag[threadIdx.x] = 0;  // one skip flag per pass: never reset while it may still be read
    for (uint32_t pass = 0; pass < 4; ++pass) {
        const uint32_t shift = 32 + 8 * pass;
        for (uint32_t t = threadIdx.x; t < kRadixWarps * 256; t += blockDim.x) hist[t >> 8][t & 255] = 0;
        __syncthreads();
        // (1) per-warp digit histogram
        for (uint32_t g0 = w_begin; g0 < w_end; g0 += 32) {
            const uint32_t i = g0 + lane;
            const bool valid = i < w_end;
            const uint32_t d = valid ? (uint32_t)(src[i] >> shift) & 255u : 256u + lane;  // invalid lanes match nobody
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == (uint32_t)__ffs(peers) - 1u) hist[warp][d] += (uint16_t)__popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // (2) exclusive scan over (digit major, warp minor); thread t owns digit t
        {
            const uint32_t d = threadIdx.x;  // kThreads == 256 digits
            uint32_t total = 0;
#pragma unroll
            for (uint32_t w = 0; w < kRadixWarps; ++w) total += hist[w][d];
            if (total == n) s_flag[pass] = 1;  // every key has this digit: the pass would not move anything
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            if (lane == 31) s_warp_sums[warp] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (uint32_t w = 0; w < warp; ++w) before += s_warp_sums[w];
            uint32_t running = before + incl - total;
#pragma unroll
            for (uint32_t w = 0; w < kRadixWarps; ++w) {
                const uint32_t c = hist[w][d];
                hist[w][d] = (uint16_t)running;
                running += c;
            }
        }
        __syncthreads();
        if (s_flag[pass]) continue;  // uniform: written before the barrier above, never rewritten
        // (3) stable scatter
        for (uint32_t g0 = w_begin; g0 < w_end; g0 += 32) {
            const uint32_t i = g0 + lane;
            const bool valid = i < w_end;
            const uint64_t key = valid ? src[i] : 0ull;
            const uint32_t d = valid ? (uint32_t)(key >> shift) & 255u : 256u + lane;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t pos = 0;
            if (valid) pos = hist[warp][d] + rank;
            __syncwarp();
            if (valid && lane == (uint32_t)__ffs(peers) - 1u) hist[warp][d] += (uint16_t)__popc(peers);
            __syncwarp();
            if (valid) dst[pos] = key;
        }
        __syncthreads();
        uint64_t *tmp = src; src = dst; dst = tmp;
    }
    if (src != a) {
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) a[t] = src[t];
        __syncthreads();
    }
    // (4) runs of exactly equal depth: order by flatten id (insertion sort by the thread that owns the run's first key)
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t depth = (uint32_t)(a[i] >> 32);
        const bool first = i == 0 || (uint32_t)(a[i - 1] >> 32) != depth;
        if (first && i + 1 < n && (uint32_t)(a[i + 1] >> 32) == depth) {
            uint32_t j = i + 1;
            while (j < n && (uint32_t)(a[j] >> 32) == depth) ++j;
            for (uint32_t u = i + 1; u < j; ++u) {
                const uint64_t x = a[u];
                uint32_t v = u;
                while (v > i && a[v - 1] > x) { a[v] = a[v - 1]; --v; }
                a[v] = x;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
tile_sort_radix_kernel(uint32_t n_tiles_total, uint32_t n_tiles, uint32_t tile_bits, int64_t n_isects_or_capacity,
                       const int32_t *__restrict__ total_dev, const int32_t *__restrict__ offsets, uint64_t *__restrict__ bucket,
                       int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    __shared__ uint64_t s_a[kSortSmall], s_b[kSortSmall];
    __shared__ uint16_t s_hist[kRadixWarps][256];
    __shared__ uint32_t s_warp_sums[kRadixWarps], s_flag[4];
    const int64_t n_isects = total_dev ? (int64_t)*total_dev : n_isects_or_capacity;
    const int64_t capacity = n_isects_or_capacity;
    const uint32_t ct = blockIdx.x;
    const int64_t start = offsets[ct];
    const int64_t end = (ct + 1 == n_tiles_total) ? n_isects : (int64_t)offsets[ct + 1];
    const uint32_t n = (uint32_t)(end - start);
    if (n == 0 || n > kSortSmall || end > capacity) return;  // longer segments belong to tile_sort_kernel<true>
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_a[t] = bucket[start + t];
    __syncthreads();
    if (n > 1) block_sort_radix(s_a, s_b, n, s_hist, s_warp_sums, s_flag);
    const uint64_t cam = ct / n_tiles, tile = ct % n_tiles;
    const int64_t hi = (int64_t)((cam << (32 + tile_bits)) | (tile << 32));
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        const uint64_t it = s_a[t];
        isect_ids[start + t] = hi | (int64_t)(int32_t)(uint32_t)(it >> 32);
        flatten_ids[start + t] = (int32_t)(uint32_t)(it & 0xffffffffu);
    }
}

// sort == false variant of the reference op: keys in emission order (ascending flatten index, row-major
// tiles inside each Gaussian), isect_tiles.cu:89-103.  `cum` = exclusive scan of tiles_per_gauss.
__global__ void __launch_bounds__(kThreads)
isect_emit_unsorted_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids,
                           const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                           const float *__restrict__ depths, float ts, uint32_t tw, uint32_t th, uint32_t tile_bits,
                           const int32_t *__restrict__ cum, int64_t *__restrict__ isect_ids,
                           int32_t *__restrict__ flatten_ids) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int32_t r = radii[i];
    if (r <= 0) return;
    const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
    const TileBox b = tile_box(m.x, m.y, (float)r, ts, tw, th);
    const uint64_t cam = camera_ids ? (uint64_t)camera_ids[i] : i / per_cam;
    const int64_t hi = (int64_t)(cam << (32 + tile_bits));
    const int64_t lo = (int64_t)(int32_t)__float_as_uint(depths[i]);
    int64_t at = cum[i];
    for (uint32_t y = b.y0; y < b.y1; ++y)
        for (uint32_t x = b.x0; x < b.x1; ++x, ++at) {
            isect_ids[at] = hi | ((int64_t)(y * tw + x) << 32) | lo;
            flatten_ids[at] = (int32_t)i;
        }
}

// stand-alone offsets from already-sorted ids: one lower_bound per (camera, tile)
__global__ void __launch_bounds__(kThreads)
offset_encode_kernel(int64_t n_isects, const int64_t *__restrict__ isect_ids, uint32_t total_tiles, uint32_t n_tiles,
                     uint32_t tile_bits, int32_t *__restrict__ offsets) {
    const uint32_t ct = blockIdx.x * kThreads + threadIdx.x;
    if (ct >= total_tiles) return;
    const int64_t target = ((int64_t)(ct / n_tiles) << tile_bits) | (int64_t)(ct % n_tiles);
    int64_t lo = 0, hi = n_isects;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((isect_ids[mid] >> 32) < target) lo = mid + 1;
        else hi = mid;
    }
    offsets[ct] = (int32_t)lo;
}

int exclusive_scan(uint32_t n, int32_t *data, int32_t *copy, int32_t *block_scratch, cudaStream_t s) {
    // data[0..n) -> exclusive prefix in place (+ optional copy); total -> block_scratch[nb]
    const uint32_t nb = cdiv(n, kScanTile);
    scan_block_sums_kernel<<<nb, kThreads, 0, s>>>(n, data, block_scratch);
    if (int e = check_launch("scan_block_sums")) return e;
    scan_single_block_kernel<<<1, 1024, 0, s>>>(nb, block_scratch);
    if (int e = check_launch("scan_single_block")) return e;
    scan_apply_kernel<<<nb, kThreads, 0, s>>>(n, data, block_scratch, data, copy);
    return check_launch("scan_apply");
}

}  // namespace gsb200

using namespace gsb200;

// GSB200_ISECT=ordered selects the depth pre-sort + ordered scatter path (csrc/isect_ordered.cu) where it applies.  It is
// bit-exact (the whole GPU suite passes on it) but, as measured on B200 at the headline size, SLOWER than the per-tile sort
// path (580 us vs 316 us: its per-piece walks are serial chains of shared-memory atomics, ~550 cycles per Gaussian), so
// the per-tile sort path stays the default.  Kept because its first half (13x fewer, 32-bit keys to sort) is the right
// starting point for a faster stage 2 -- see DESIGN.md section 6.
static bool use_ordered(uint32_t C, uint32_t n_elems, uint32_t tw, uint32_t th) {
    static const bool on = [] {
        const char *v = getenv("GSB200_ISECT");
        return v != nullptr && strcmp(v, "ordered") == 0;
    }();
    return on && n_elems > 0 && ordered_plan(C, n_elems, tw, th).ok;
}

static size_t tilesort_scratch_ints(uint32_t T) {  // counters | work list | total | scan scratch or block totals
    return (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 2 + max(T / kScanTile + 3, T / 32 + 2);
}
// GSB200_ISECT_SCAN=legacy: the five-launch totals / scan / cursor sequence of round 2 (A/B timing)
static bool fused_replica_scan() {
    static const bool on = [] {
        const char *v = getenv("GSB200_ISECT_SCAN");
        return !(v != nullptr && strcmp(v, "legacy") == 0);
    }();
    return on;
}
static size_t tilesort_total_index(uint32_t T) { return (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 1; }
// CTA-private histograms (isect_count_private_kernel) are opt-in, GSB200_ISECT_PRIVATE=1 (read at every call: a test toggles
// it): bit-exact like the default, but measured no faster at the headline size -- 183 vs 187 us for the stage
// (profiles/r2v45_timing_priv*.log); with round-robin ownership a CTA touches ~80 % of the 8160 tiles, so the global atomics
// only halve (1.9 M (CTA, tile) pairs per pass instead of 4 M intersections) and the row store / reload eats the rest
static uint32_t private_ctas(uint32_t T, uint32_t n_elems) {
    const char *v = getenv("GSB200_ISECT_PRIVATE");
    const bool on = v != nullptr && v[0] == '1';
    if (!on || T > kPrivMaxTiles || n_elems < 8u * kPrivThreads) return 0;
    return min(2u * 148u, cdiv(n_elems, (uint32_t)kPrivThreads));
}
static size_t private_rows_at(uint32_t T) { return (tilesort_scratch_ints(T) + 3) & ~(size_t)3; }
static int private_smem_ok() {
    static int state = -1;  // -1: attributes not set yet
    if (state < 0) {
        const int bytes = (int)(kPrivMaxTiles * sizeof(int32_t));
        const bool ok = cudaFuncSetAttribute(isect_count_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
                        cudaFuncSetAttribute(isect_scatter_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
        state = ok ? 1 : 0;
    }
    return state;
}

extern "C" int gsb200_isect_plan(uint32_t C, uint32_t n_elems, uint32_t tile_width, uint32_t tile_height, int64_t *scratch_ints,
                                 int64_t *total_index) {
    const uint32_t T = max(C * tile_width * tile_height, 1u);
    if (use_ordered(C, n_elems, tile_width, tile_height)) {
        const OrderedPlan P = ordered_plan(C, n_elems, tile_width, tile_height);
        *scratch_ints = (int64_t)P.scratch_ints;
        *total_index = (int64_t)P.total_index;
    } else {
        // (+ one row of T counters per CTA of the private-histogram passes, behind everything else)
        *scratch_ints = (int64_t)(private_rows_at(T) + (size_t)private_ctas(T, n_elems) * T);
        *total_index = (int64_t)tilesort_total_index(T);
    }
    return 0;
}

extern "C" int gsb200_isect_prepare(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera, const int64_t *camera_ids,
                                    const float *means2d, const int32_t *radii, const float *depths, uint32_t tile_size,
                                    uint32_t tile_width, uint32_t tile_height, int32_t *tiles_per_gauss, int32_t *isect_offsets,
                                    int32_t *tile_cursor, int64_t *n_isects_host, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t T = C * tile_width * tile_height;
    if (n_isects_host != nullptr) *n_isects_host = 0;
    if (bit_width(tile_width * tile_height) + bit_width(C) > 32)
        return fail(2, "gsb200_isect_prepare: camera and tile ids do not fit 32 bits");
    if (T == 0) return 0;
    if (use_ordered(C, n_elems, tile_width, tile_height)) {
        if (depths == nullptr) return fail(2, "gsb200_isect_prepare: depths is required");
        if (int e = ordered_prepare(C, n_elems, elems_per_camera, camera_ids, means2d, radii, depths, tile_size, tile_width,
                                    tile_height, tiles_per_gauss, isect_offsets, tile_cursor, s))
            return e;
        if (n_isects_host != nullptr) {
            int32_t total = 0;
            const OrderedPlan P = ordered_plan(C, n_elems, tile_width, tile_height);
            GSB_CUDA_OK(cudaMemcpyAsync(&total, tile_cursor + P.total_index, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            GSB_CUDA_OK(cudaStreamSynchronize(s));
            *n_isects_host = total;
        }
        return 0;
    }
    if (n_elems == 0) {
        GSB_CUDA_OK(cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * T, s));
        GSB_CUDA_OK(cudaMemsetAsync(tile_cursor + (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 1, 0, sizeof(int32_t), s));
        return 0;
    }
    GSB_CUDA_OK(cudaMemsetAsync(tile_cursor, 0, sizeof(int32_t) * ((size_t)T * GSB200_ISECT_REPLICAS + 1), s));
    const uint32_t priv = private_smem_ok() ? private_ctas(T, n_elems) : 0;
    if (priv)
        isect_count_private_kernel<<<priv, kPrivThreads, T * sizeof(int32_t), s>>>(
            n_elems, elems_per_camera ? elems_per_camera : 1, camera_ids, means2d, radii, (float)tile_size, tile_width, tile_height, T,
            tiles_per_gauss, tile_cursor, tile_cursor + private_rows_at(T));
    else
        isect_count_kernel<<<cdiv(n_elems, kThreads), kThreads, 0, s>>>(n_elems, elems_per_camera ? elems_per_camera : 1,
                                                                         camera_ids, means2d, radii, (float)tile_size,
                                                                         tile_width, tile_height, T, tiles_per_gauss, tile_cursor);
    if (int e = check_launch("isect_count")) return e;
    // scratch lives at the end of the caller's tile_cursor buffer (a stream-ordered allocation here would
    // go back to the driver after every host synchronisation -- milliseconds)
    int32_t *scratch = tile_cursor + (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 2;
    int32_t total = 0;
    int32_t *total_dev = tile_cursor + (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 1;
    int e = 0;
    if (fused_replica_scan()) {
        const uint32_t nblk = cdiv(T, kTilesPerBlock);
        replica_block_totals_kernel<<<nblk, 1024, 0, s>>>(T, tile_cursor, scratch);
        e = check_launch("replica_block_totals");
        if (e == 0) {
            replica_block_cursor_kernel<<<nblk, 1024, 0, s>>>(T, scratch, tile_cursor, isect_offsets, total_dev, kSortSmall);
            e = check_launch("replica_block_cursor");
        }
    } else {
        replica_totals_kernel<<<cdiv(T, kThreads), kThreads, 0, s>>>(T, tile_cursor, isect_offsets);
        if (int e2 = check_launch("replica_totals")) return e2;
        const uint32_t nb = cdiv(T, kScanTile);
        e = exclusive_scan(T, isect_offsets, nullptr, scratch, s);
        if (e == 0) {
            replica_cursor_kernel<<<cdiv(T, kThreads), kThreads, 0, s>>>(T, isect_offsets, tile_cursor, kSortSmall);
            e = check_launch("replica_cursor");
        }
        if (e == 0 && cudaMemcpyAsync(total_dev, scratch + nb, sizeof(int32_t), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
            e = fail(101, "gsb200_isect_prepare: device copy of the intersection count failed");
    }
    if (e == 0 && n_isects_host != nullptr) {
        cudaError_t ce = cudaMemcpyAsync(&total, total_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess) e = fail(101, cudaGetErrorString(ce));
    }
    if (e) return e;
    if (n_isects_host != nullptr) *n_isects_host = total;
    return 0;
}

extern "C" int gsb200_isect_count_async(const int32_t *tile_cursor, int64_t total_index, int32_t *count_pinned,
                                        gsb200_stream_t stream) {
    if (tile_cursor == nullptr || count_pinned == nullptr || total_index < 0) return fail(2, "gsb200_isect_count_async: bad argument");
    GSB_CUDA_OK(cudaMemcpyAsync(count_pinned, tile_cursor + total_index, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
    return 0;
}

extern "C" int gsb200_isect_fill(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera, const int64_t *camera_ids,
                                 const float *means2d, const int32_t *radii, const float *depths, uint32_t tile_size,
                                 uint32_t tile_width, uint32_t tile_height, int64_t n_isects, const int32_t *isect_offsets,
                                 int32_t *tile_cursor, uint64_t *scratch, int64_t *isect_ids, int32_t *flatten_ids,
                                 gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_isects == 0 || n_elems == 0) return 0;
    const uint32_t n_tiles = tile_width * tile_height, T = C * n_tiles;
    // n_isects < 0: capacity mode (see include/gsb200.h)
    const int64_t bound = n_isects < 0 ? -n_isects : n_isects;
    if (use_ordered(C, n_elems, tile_width, tile_height))
        return ordered_fill(C, n_elems, elems_per_camera, camera_ids, means2d, radii, tile_size, tile_width, tile_height, bound,
                            isect_offsets, tile_cursor, isect_ids, flatten_ids, s);
    const int32_t *total_dev = n_isects < 0 ? tile_cursor + (size_t)(GSB200_ISECT_REPLICAS + 1) * T + 1 : nullptr;
    const uint32_t priv = private_smem_ok() ? private_ctas(T, n_elems) : 0;
    if (priv)
        isect_scatter_private_kernel<<<priv, kPrivThreads, T * sizeof(int32_t), s>>>(
            n_elems, elems_per_camera ? elems_per_camera : 1, camera_ids, means2d, radii, depths, (float)tile_size, tile_width,
            tile_height, T, tile_cursor, tile_cursor + private_rows_at(T), scratch, bound);
    else
        isect_scatter_kernel<<<cdiv(n_elems, kThreads), kThreads, 0, s>>>(n_elems, elems_per_camera ? elems_per_camera : 1,
                                                                           camera_ids, means2d, radii, depths, (float)tile_size,
                                                                           tile_width, tile_height, T, tile_cursor, scratch, bound);
    if (int e = check_launch("isect_scatter")) return e;
    const int32_t *large_list = tile_cursor + (size_t)GSB200_ISECT_REPLICAS * T;
    static const int small_sort = [] {  // 0 depth buckets (default), 1 bitonic network, 2 radix (both opt-in)
        const char *v = getenv("GSB200_TILE_SORT");
        return v == nullptr ? 0 : strcmp(v, "bitonic") == 0 ? 1 : strcmp(v, "radix") == 0 ? 2 : 0;
    }();
    if (small_sort == 0)
        tile_sort_bucket_kernel<<<T, kThreads, 0, s>>>(T, n_tiles, bit_width(n_tiles), bound, total_dev, isect_offsets, scratch,
                                                        isect_ids, flatten_ids);
    else if (small_sort == 2)
        tile_sort_radix_kernel<<<T, kThreads, 0, s>>>(T, n_tiles, bit_width(n_tiles), bound, total_dev, isect_offsets, scratch,
                                                       isect_ids, flatten_ids);
    else
        tile_sort_kernel<false><<<T, kThreads, 0, s>>>(T, n_tiles, bit_width(n_tiles), bound, total_dev, isect_offsets, large_list,
                                                        scratch, isect_ids, flatten_ids);
    if (int e = check_launch("tile_sort")) return e;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CUDA_OK(cudaFuncSetAttribute(tile_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSortLarge * sizeof(uint64_t))));
        GSB_CUDA_OK(cudaFuncSetAttribute(tile_sort_bucket_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kBucketLargeSmem));
        attr_set = true;
    }
    const uint32_t large_grid = min(T, 2u * 148u);
    if (small_sort == 0)
        tile_sort_bucket_large_kernel<<<large_grid, 1024, kBucketLargeSmem, s>>>(
            T, n_tiles, bit_width(n_tiles), bound, total_dev, isect_offsets, large_list, scratch, isect_ids, flatten_ids);
    else
        tile_sort_kernel<true><<<large_grid, 1024, kSortLarge * sizeof(uint64_t), s>>>(
            T, n_tiles, bit_width(n_tiles), bound, total_dev, isect_offsets, large_list, scratch, isect_ids, flatten_ids);
    return check_launch("tile_sort_large");
}

extern "C" int gsb200_isect_emit_unsorted(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera, const int64_t *camera_ids,
                                          const float *means2d, const int32_t *radii, const float *depths,
                                          uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                          const int32_t *tiles_per_gauss, int32_t *cum_scratch, int64_t *isect_ids,
                                          int32_t *flatten_ids, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_elems == 0) return 0;
    GSB_CUDA_OK(cudaMemcpyAsync(cum_scratch, tiles_per_gauss, sizeof(int32_t) * n_elems, cudaMemcpyDeviceToDevice, s));
    int32_t *scratch = cum_scratch + n_elems;  // block sums behind the prefix array
    int e = exclusive_scan(n_elems, cum_scratch, nullptr, scratch, s);
    if (e) return e;
    const uint32_t n_tiles = tile_width * tile_height;
    isect_emit_unsorted_kernel<<<cdiv(n_elems, kThreads), kThreads, 0, s>>>(
        n_elems, elems_per_camera ? elems_per_camera : 1, camera_ids, means2d, radii, depths, (float)tile_size, tile_width,
        tile_height, bit_width(n_tiles), cum_scratch, isect_ids, flatten_ids);
    return check_launch("isect_emit_unsorted");
}

extern "C" int gsb200_isect_offset_encode(int64_t n_isects, const int64_t *isect_ids, uint32_t C, uint32_t tile_width,
                                          uint32_t tile_height, int32_t *isect_offsets, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t n_tiles = tile_width * tile_height, T = C * n_tiles;
    if (T == 0) return 0;
    if (n_isects == 0) {  // isect_tiles.cu:385-387
        GSB_CUDA_OK(cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * T, s));
        return 0;
    }
    offset_encode_kernel<<<cdiv(T, kThreads), kThreads, 0, s>>>(n_isects, isect_ids, T, n_tiles, bit_width(n_tiles),
                                                                 isect_offsets);
    return check_launch("offset_encode");
}
