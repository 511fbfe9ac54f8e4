// Stage 2, ordered path: depth pre-sort of the VISIBLE Gaussians + ordered scatter -- no per-tile sort (sm_100a).
//
// Same outputs as isect.cu (the reference's isect_tiles + isect_offset_encode, gsplat/cuda/csrc/isect_tiles.cu:16-389,
// bit for bit); the work is organised around one fact: every Gaussian contributes AT MOST ONE intersection to a tile.
// If the visible (camera, Gaussian) entries are visited in (camera, depth bits, flatten index) order -- the order the
// reference's stable radix sort leaves inside every tile -- the rank of an entry's intersection inside tile t is simply
// the number of earlier entries that cover t.  So:
//
//   compact   : entries with a non-empty tile box -> (depth bits, flatten index), in flatten-index order
//   sort      : stable LSD radix sort of those ~0.3 M keys by depth bits (then by camera, for C > 1) -- 13x fewer keys
//               than the 4 M (tile | depth) intersection keys the reference sorts, and 32-bit instead of 64-bit
//   pieces    : the sorted list is cut into pieces of 512 entries; one WARP owns a piece and a private array of tile
//               counters in shared memory: pass 1 counts the piece's intersections per tile (rows of H)
//   column scan: per tile, exclusive prefix over the pieces (in place in H) and the tile total -> isect_offsets after one
//               scan over tiles (the reference needs the sorted keys for this)
//   scatter   : pass 2 walks the piece again IN ORDER; the whole warp visits the tile box of one entry at a time, every
//               lane takes the next slot of its tile (cursor = offsets[tile] + H[piece][tile]) and writes the final
//               isect_ids / flatten_ids entry.  Nothing is sorted afterwards.
#include "isect.cuh"

namespace gsb200 {

constexpr uint32_t kVisItems = 8;
constexpr uint32_t kVisTile = kThreads * kVisItems;  // 2048 elements per block in the compaction
constexpr uint32_t kSortItems = 16;
constexpr uint32_t kSortTile = kThreads * kSortItems;  // 4096 keys per block in a radix pass
constexpr uint32_t kPiece = 512;                        // sorted entries per piece
constexpr uint32_t kPieceSmemMax = 220 * 1024;

constexpr uint32_t kMaxPasses = 8;   // 4 depth digits + up to 4 camera digits
constexpr uint32_t kColChunks = 16;  // the column scan over the pieces runs as 16 independent chunks per tile

struct Layout {
    size_t total, n_vis, n_pieces, block_counts, keys[2], vals[2], hist, cam_start, piece_base, tile_scan, colsum, H, end;
    uint32_t nbv, nbs, pieces_max;
};

static Layout make_layout(uint32_t C, uint32_t n, uint32_t n_tiles) {
    Layout L;
    L.nbv = cdiv(n, kVisTile);
    L.nbs = cdiv(n, kSortTile);
    L.pieces_max = cdiv(n, kPiece) + C;
    size_t at = 0;
    auto take = [&](size_t count) { const size_t o = at; at += (count + 3) & ~size_t(3); return o; };  // 16-byte aligned parts
    L.total = take(1);
    L.n_vis = L.total + 1;
    L.n_pieces = L.total + 2;
    L.block_counts = take(L.nbv + 2);
    L.keys[0] = take(n); L.keys[1] = take(n);
    L.vals[0] = take(n); L.vals[1] = take(n);
    L.hist = take((size_t)kMaxPasses * 256 * L.nbs);  // [pass][block][digit], zeroed per call
    L.cam_start = take(C + 1);
    L.piece_base = take(C + 1);
    L.tile_scan = take((size_t)C * n_tiles / kScanTile + 4);
    L.colsum = take((size_t)kColChunks * C * n_tiles);
    L.H = take((size_t)L.pieces_max * n_tiles);
    L.end = at;
    return L;
}

OrderedPlan ordered_plan(uint32_t C, uint32_t n_elems, uint32_t tw, uint32_t th) {
    OrderedPlan P{false, 0, 0};
    const uint64_t n_tiles = (uint64_t)tw * th;
    if (C == 0 || n_tiles == 0 || n_tiles * 4 > kPieceSmemMax) return P;
    const Layout L = make_layout(C, n_elems, (uint32_t)n_tiles);
    if (L.end > (size_t(3) << 28)) return P;  // > 3 GB of scratch: let the per-tile sort path handle it
    P.ok = true;
    P.scratch_ints = L.end;
    P.total_index = L.total;
    return P;
}

__device__ __forceinline__ bool visible_box(uint32_t i, const float *means2d, const int32_t *radii, float ts, uint32_t tw, uint32_t th,
                                            TileBox &b, int32_t &tiles) {
    const int32_t r = radii[i];
    tiles = 0;
    if (r <= 0) return false;
    const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + i);
    b = tile_box(m.x, m.y, (float)r, ts, tw, th);
    tiles = (int32_t)((b.y1 - b.y0) * (b.x1 - b.x0));  // the reference multiplies the two uint32 extents (isect_tiles.cu:73-76)
    return b.x1 > b.x0 && b.y1 > b.y0;
}

// ---- compaction of the visible entries ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
vis_count_kernel(uint32_t n, const float *__restrict__ means2d, const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th,
                 int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ block_counts) {
    int cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < kVisItems; ++k) {
        const uint32_t i = blockIdx.x * kVisTile + k * kThreads + threadIdx.x;
        bool has = false;
        if (i < n) {
            TileBox b;
            int32_t tiles;
            has = visible_box(i, means2d, radii, ts, tw, th, b, tiles);
            tiles_per_gauss[i] = tiles;
        }
        cnt += __syncthreads_count(has);
    }
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(kThreads)
vis_write_kernel(uint32_t n, const float *__restrict__ means2d, const int32_t *__restrict__ radii, const float *__restrict__ depths,
                 float ts, uint32_t tw, uint32_t th, const int32_t *__restrict__ block_offsets, uint32_t *__restrict__ keys,
                 uint32_t *__restrict__ vals, int32_t *__restrict__ hist0) {
    __shared__ int s_w[kThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = block_offsets[blockIdx.x];
#pragma unroll
    for (uint32_t k = 0; k < kVisItems; ++k) {
        const uint32_t i = blockIdx.x * kVisTile + k * kThreads + threadIdx.x;
        bool has = false;
        if (i < n) {
            TileBox b;
            int32_t tiles;
            has = visible_box(i, means2d, radii, ts, tw, th, b, tiles);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (uint32_t w = 0; w < kThreads / 32; ++w) {
            const int c = s_w[w];
            before += w < warp ? c : 0;
            all += c;
        }
        if (has) {
            const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
            const uint32_t key = __float_as_uint(depths[i]);
            keys[pos] = key;
            vals[pos] = i;
            atomicAdd(hist0 + (size_t)((uint32_t)pos / kSortTile) * 256 + (key & 255u), 1);  // histogram of the first radix pass
        }
        base += all;
        __syncthreads();
    }
}

// ---- stable LSD radix sort, 8-bit digits, n on the device ---------------------------------------------------------------
// KIND 0: digit of the key (depth bits); KIND 1: digit of the entry's camera index
template <int KIND>
__device__ __forceinline__ uint32_t digit_of(uint32_t key, uint32_t val, uint32_t shift, uint32_t per_cam, const int64_t *camera_ids) {
    if (KIND == 0) return (key >> shift) & 255u;
    const uint32_t cam = camera_ids ? (uint32_t)camera_ids[val] : val / per_cam;
    return (cam >> shift) & 255u;
}

// One pass = ONE kernel.  hist [blocks][256] holds this pass's per-block digit counts (written by the previous pass, or by
// vis_write_kernel for the first); every block derives its own scatter bases from it (digit totals + the blocks before
// it), ranks its 4096 keys stably (match_any per 32-key round, per-warp counters) and, while scattering, counts the
// NEXT pass's digits per destination block into hist_next.
__device__ __forceinline__ uint32_t digit_rt(int kind, uint32_t key, uint32_t val, uint32_t shift, uint32_t per_cam,
                                             const int64_t *camera_ids) {
    return kind == 0 ? digit_of<0>(key, val, shift, per_cam, camera_ids) : digit_of<1>(key, val, shift, per_cam, camera_ids);
}

__global__ void __launch_bounds__(kThreads)
radix_pass_kernel(const int32_t *__restrict__ n_dev, const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int kind, uint32_t shift, int next_kind,
                  uint32_t next_shift, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const int32_t *__restrict__ hist,
                  int32_t *__restrict__ hist_next) {
    constexpr uint32_t NW = kThreads / 32;
    __shared__ uint16_t s_cnt[NW][256];
    __shared__ uint32_t s_off[NW][256];
    __shared__ uint32_t s_wsum[NW];
    const uint32_t n = (uint32_t)*n_dev;
    const uint32_t nb = cdiv(n, kSortTile), blk = blockIdx.x;
    if (blk >= nb) return;
    const uint32_t base = blk * kSortTile;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t = threadIdx.x; t < NW * 256; t += kThreads) (&s_cnt[0][0])[t] = 0;
    // scatter base of digit d (= this thread): keys with a smaller digit + keys with digit d in the blocks before this one
    uint32_t tot = 0, pre = 0;
    {
        const uint32_t d = threadIdx.x;
        for (uint32_t bb = 0; bb < nb; ++bb) {
            const uint32_t v = (uint32_t)hist[(size_t)bb * 256 + d];
            pre += bb < blk ? v : 0u;
            tot += v;
        }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (uint32_t w = 0; w < NW; ++w) before += w < warp ? s_wsum[w] : 0u;
        pre += before + incl - tot;  // exclusive prefix over the digits
    }
    // warp w owns keys [base + w * 512, +512) as 16 rounds of 32 consecutive keys: (round, lane) order is the input order
    uint32_t key[kSortItems], val[kSortItems], rd[kSortItems];  // rd: rank inside the warp's slice << 8 | digit
#pragma unroll
    for (uint32_t r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + warp * (32 * kSortItems) + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : 0u;
        val[r] = ok ? vals_in[i] : 0u;
        const uint32_t d = ok ? digit_rt(kind, key[r], val[r], shift, per_cam, camera_ids) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && (int)lane == leader) {
            old = s_cnt[warp][d];
            s_cnt[warp][d] = (uint16_t)(old + __popc(peers));
        }
        __syncwarp();
        old = __shfl_sync(0xffffffffu, old, leader);
        rd[r] = ((old + __popc(peers & ((1u << lane) - 1u))) << 8) | (d & 255u);
    }
    __syncthreads();
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = pre;
#pragma unroll
        for (uint32_t w = 0; w < NW; ++w) {
            s_off[w][d] = run;
            run += s_cnt[w][d];
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + warp * (32 * kSortItems) + r * 32 + lane;
        if (i < n) {
            const uint32_t pos = s_off[warp][rd[r] & 255u] + (rd[r] >> 8);
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
            if (next_kind >= 0)
                atomicAdd(hist_next + (size_t)(pos / kSortTile) * 256 + digit_rt(next_kind, key[r], val[r], next_shift, per_cam, camera_ids), 1);
        }
    }
}

// ---- pieces -----------------------------------------------------------------------------------------------------------
// cam_start[c] = first sorted entry of camera c (cam_start[C] = V); piece_base[c] = pieces before camera c
__global__ void __launch_bounds__(kThreads)
cam_pieces_kernel(uint32_t C, const int32_t *__restrict__ n_dev, const uint32_t *__restrict__ vals, uint32_t per_cam,
                  const int64_t *__restrict__ camera_ids, int32_t *__restrict__ cam_start, int32_t *__restrict__ piece_base,
                  int32_t *__restrict__ n_pieces) {
    const uint32_t n = (uint32_t)*n_dev;
    for (uint32_t c = threadIdx.x; c <= C; c += blockDim.x) {
        uint32_t lo = 0, hi = n;  // first entry whose camera is >= c
        if (c == 0) hi = 0;
        else if (c == C) lo = n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t v = vals[mid];
            const uint32_t cam = camera_ids ? (uint32_t)camera_ids[v] : v / per_cam;
            if (cam < c) lo = mid + 1; else hi = mid;
        }
        cam_start[c] = (int32_t)lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (uint32_t c = 0; c < C; ++c) {
            piece_base[c] = run;
            run += (cam_start[c + 1] - cam_start[c] + (int)kPiece - 1) / (int)kPiece;
        }
        piece_base[C] = run;
        *n_pieces = run;
    }
}

// One warp per piece.  SCATTER = false: count the piece's intersections per tile -> row of H.  SCATTER = true: take
// slots in order and write the final entries.  Shared memory: one int per tile and warp (+ a 32-entry box staging area).
// The entries of a piece are visited in order, one per step: the whole warp walks that entry's tile box (lanes along x
// first, boxes wider / larger than the warp in several patches) and every lane bumps the counter of its tile with a
// shared-memory atomic.  Steps do not depend on one another except through those counters, and a warp's shared-memory
// atomics are performed in issue order, so consecutive steps pipeline (no shuffle, no barrier inside the loop) while the
// slots of a tile are still handed out in (depth, index) order.
constexpr uint32_t kPieceWarpsMax = 4;

template <bool SCATTER>
__global__ void __launch_bounds__(kPieceWarpsMax * 32)
piece_kernel(uint32_t C, uint32_t n_tiles, uint32_t tile_bits, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
             float ts, uint32_t tw, uint32_t th, const int32_t *__restrict__ n_pieces, const int32_t *__restrict__ cam_start,
             const int32_t *__restrict__ piece_base, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
             int32_t *__restrict__ H, const int32_t *__restrict__ offsets, int64_t capacity, int64_t *__restrict__ isect_ids,
             int32_t *__restrict__ flatten_ids) {
    extern __shared__ int s_tiles[];
    __shared__ uint4 s_stage[kPieceWarpsMax][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + warp;
    if (p >= (uint32_t)*n_pieces) return;  // whole warps leave; no block barrier below
    uint32_t c = 0;
    {
        uint32_t lo = 0, hi = C;  // last camera with piece_base[c] <= p
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t)piece_base[mid] <= p) lo = mid; else hi = mid;
        }
        c = lo;
    }
    const uint32_t first = (uint32_t)cam_start[c] + (p - (uint32_t)piece_base[c]) * kPiece;
    const uint32_t last = min((uint32_t)cam_start[c + 1], first + kPiece);
    int *cur = s_tiles + (size_t)warp * n_tiles;
    int32_t *row = H + (size_t)p * n_tiles;
    if (SCATTER) {
        const int32_t *off = offsets + (size_t)c * n_tiles;
        for (uint32_t t = lane; t < n_tiles; t += 32) cur[t] = off[t] + row[t];
    } else {
        for (uint32_t t = lane; t < n_tiles; t += 32) cur[t] = 0;
    }
    __syncwarp();
    const int64_t hi_bits = (int64_t)((uint64_t)c << (32 + tile_bits));
    uint4 *stage = s_stage[warp];
    for (uint32_t k = first; k < last; k += 32) {
        const uint32_t e = k + lane;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (e < last) {
            const uint32_t id = vals[e];
            const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + id);
            const TileBox b = tile_box(m.x, m.y, (float)radii[id], ts, tw, th);
            rec = make_uint4(b.x0 | (b.y0 << 16), (b.x1 - b.x0) | ((b.y1 - b.y0) << 16), keys[e], id);
        }
        stage[lane] = rec;
        __syncwarp();
        const uint32_t cnt = min(32u, last - k);
        for (uint32_t g = 0; g < cnt; ++g) {
            const uint4 q = stage[g];  // broadcast
            const uint32_t x0 = q.x & 0xffffu, y0 = q.x >> 16, nx = q.y & 0xffffu, ny = q.y >> 16;
            for (uint32_t xb = 0; xb < nx; xb += 32) {
                const uint32_t w = min(32u, nx - xb);
                const uint32_t sh = w <= 1 ? 0 : 32 - __clz(w - 1);  // log2 of the padded patch width
                const uint32_t cx = lane & ((1u << sh) - 1u), ry = lane >> sh, rows = 32u >> sh;
                for (uint32_t yb = 0; yb < ny; yb += rows) {
                    if (cx < w && yb + ry < ny) {
                        const uint32_t tile = (y0 + yb + ry) * tw + x0 + xb + cx;
                        if (SCATTER) {
                            const int slot = atomicAdd(&cur[tile], 1);
                            if (slot < capacity) {
                                // low word: sign-extending OR like the reference's (int64)(int32)depth cast (isect_tiles.cu:88)
                                isect_ids[slot] = hi_bits | ((int64_t)tile << 32) | (int64_t)(int32_t)q.z;
                                flatten_ids[slot] = (int32_t)q.w;
                            }
                        } else {
                            atomicAdd(&cur[tile], 1);
                        }
                    }
                }
            }
        }
        __syncwarp();  // the staging area is rewritten by the next batch
    }
    if (!SCATTER) {
        __syncwarp();
        for (uint32_t t = lane; t < n_tiles; t += 32) row[t] = cur[t];
    }
}

// Column scan over the pieces, per (camera, tile), as kColChunks independent chunks: sums of the chunks first, then each
// chunk turns its rows of H into exclusive prefixes (in place) starting from the sum of the chunks before it; the last
// chunk leaves the tile's total.
__device__ __forceinline__ void chunk_range(const int32_t *piece_base, uint32_t c, uint32_t ch, uint32_t &p0, uint32_t &p1) {
    const uint32_t b0 = (uint32_t)piece_base[c], b1 = (uint32_t)piece_base[c + 1];
    const uint32_t len = cdiv(b1 - b0, kColChunks);
    p0 = min(b1, b0 + ch * len);
    p1 = min(b1, p0 + len);
}

__global__ void __launch_bounds__(kThreads)
piece_colsum_kernel(uint32_t C, uint32_t n_tiles, const int32_t *__restrict__ piece_base, const int32_t *__restrict__ H,
                    int32_t *__restrict__ colsum) {
    const uint32_t ct = blockIdx.x * kThreads + threadIdx.x, ch = blockIdx.y;
    if (ct >= C * n_tiles) return;
    const uint32_t c = ct / n_tiles, t = ct % n_tiles;
    uint32_t p0, p1;
    chunk_range(piece_base, c, ch, p0, p1);
    int sum = 0;
#pragma unroll 8
    for (uint32_t p = p0; p < p1; ++p) sum += H[(size_t)p * n_tiles + t];
    colsum[(size_t)ch * C * n_tiles + ct] = sum;
}

__global__ void __launch_bounds__(kThreads)
piece_colscan_kernel(uint32_t C, uint32_t n_tiles, const int32_t *__restrict__ piece_base, int32_t *__restrict__ H,
                     const int32_t *__restrict__ colsum, int32_t *__restrict__ tile_counts) {
    const uint32_t ct = blockIdx.x * kThreads + threadIdx.x, ch = blockIdx.y;
    if (ct >= C * n_tiles) return;
    const uint32_t c = ct / n_tiles, t = ct % n_tiles;
    uint32_t p0, p1;
    chunk_range(piece_base, c, ch, p0, p1);
    int run = 0;
    for (uint32_t k = 0; k < ch; ++k) run += colsum[(size_t)k * C * n_tiles + ct];
    uint32_t p = p0;
    for (; p + 4 <= p1; p += 4) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = H[(size_t)(p + u) * n_tiles + t];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            H[(size_t)(p + u) * n_tiles + t] = run;
            run += v[u];
        }
    }
    for (; p < p1; ++p) {
        const int v = H[(size_t)p * n_tiles + t];
        H[(size_t)p * n_tiles + t] = run;
        run += v;
    }
    if (ch == kColChunks - 1) tile_counts[ct] = run;
}

static uint32_t piece_warps(uint32_t n_tiles) {
    uint32_t w = kPieceSmemMax / (n_tiles * 4);
    return w > kPieceWarpsMax ? kPieceWarpsMax : (w < 1 ? 1 : w);
}

int ordered_prepare(uint32_t C, uint32_t n, uint32_t per_cam, const int64_t *camera_ids, const float *means2d, const int32_t *radii,
                    const float *depths, uint32_t tile_size, uint32_t tw, uint32_t th, int32_t *tiles_per_gauss,
                    int32_t *isect_offsets, int32_t *scratch, cudaStream_t s) {
    const uint32_t n_tiles = tw * th, T = C * n_tiles;
    const Layout L = make_layout(C, n, n_tiles);
    const float ts = (float)tile_size;
    int32_t *block_counts = scratch + L.block_counts;
    uint32_t *keys[2] = {reinterpret_cast<uint32_t *>(scratch + L.keys[0]), reinterpret_cast<uint32_t *>(scratch + L.keys[1])};
    uint32_t *vals[2] = {reinterpret_cast<uint32_t *>(scratch + L.vals[0]), reinterpret_cast<uint32_t *>(scratch + L.vals[1])};
    int32_t *hist = scratch + L.hist;
    // visible entries, in flatten-index order (+ the first pass's per-block digit histogram)
    const uint32_t cam_passes = C > 1 ? cdiv(bit_width(C - 1), 8u) : 0u;
    const uint32_t n_pass = 4 + cam_passes;
    GSB_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t)n_pass * 256 * L.nbs, s));
    vis_count_kernel<<<L.nbv, kThreads, 0, s>>>(n, means2d, radii, ts, tw, th, tiles_per_gauss, block_counts);
    if (int e = check_launch("vis_count")) return e;
    scan_single_block_kernel<<<1, 1024, 0, s>>>(L.nbv, block_counts);
    if (int e = check_launch("vis_scan")) return e;
    vis_write_kernel<<<L.nbv, kThreads, 0, s>>>(n, means2d, radii, depths, ts, tw, th, block_counts, keys[0], vals[0], hist);
    if (int e = check_launch("vis_write")) return e;
    const int32_t *n_vis = block_counts + L.nbv;  // the scan leaves the total behind the prefixes
    // stable sort by depth bits (4 digits), then by camera: one kernel per pass
    int cur = 0;
    for (uint32_t ps = 0; ps < n_pass; ++ps) {
        const int kind = ps < 4 ? 0 : 1, next_kind = ps + 1 < n_pass ? (ps + 1 < 4 ? 0 : 1) : -1;
        const uint32_t shift = ps < 4 ? ps * 8 : (ps - 4) * 8, next_shift = ps + 1 < 4 ? (ps + 1) * 8 : (ps + 1 - 4) * 8;
        radix_pass_kernel<<<L.nbs, kThreads, 0, s>>>(n_vis, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], kind, shift, next_kind,
                                                    next_shift, per_cam ? per_cam : 1, camera_ids, hist + (size_t)ps * 256 * L.nbs,
                                                    hist + (size_t)(ps + 1 < n_pass ? ps + 1 : ps) * 256 * L.nbs);
        if (int e = check_launch("radix_pass")) return e;
        cur ^= 1;
    }
    if (cur != 0) {  // the fill reads keys[0] / vals[0]
        GSB_CUDA_OK(cudaMemcpyAsync(keys[0], keys[1], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_OK(cudaMemcpyAsync(vals[0], vals[1], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
    }
    cam_pieces_kernel<<<1, kThreads, 0, s>>>(C, n_vis, vals[0], per_cam ? per_cam : 1, camera_ids, scratch + L.cam_start,
                                            scratch + L.piece_base, scratch + L.n_pieces);
    if (int e = check_launch("cam_pieces")) return e;
    // pass 1 over the pieces, column scan, tile offsets
    const uint32_t W = piece_warps(n_tiles);
    const size_t smem = (size_t)W * n_tiles * sizeof(int);
    static bool smem_set = false;
    if (!smem_set) {
        GSB_CUDA_OK(cudaFuncSetAttribute(piece_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPieceSmemMax));
        GSB_CUDA_OK(cudaFuncSetAttribute(piece_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPieceSmemMax));
        smem_set = true;
    }
    piece_kernel<false><<<cdiv(L.pieces_max, W), W * 32, smem, s>>>(
        C, n_tiles, bit_width(n_tiles), means2d, radii, ts, tw, th, scratch + L.n_pieces, scratch + L.cam_start,
        scratch + L.piece_base, keys[0], vals[0], scratch + L.H, nullptr, 0, nullptr, nullptr);
    if (int e = check_launch("piece_count")) return e;
    const dim3 cgrid(cdiv(T, kThreads), kColChunks);
    piece_colsum_kernel<<<cgrid, kThreads, 0, s>>>(C, n_tiles, scratch + L.piece_base, scratch + L.H, scratch + L.colsum);
    if (int e = check_launch("piece_colsum")) return e;
    piece_colscan_kernel<<<cgrid, kThreads, 0, s>>>(C, n_tiles, scratch + L.piece_base, scratch + L.H, scratch + L.colsum, isect_offsets);
    if (int e = check_launch("piece_colscan")) return e;
    if (int e = exclusive_scan(T, isect_offsets, nullptr, scratch + L.tile_scan, s)) return e;
    GSB_CUDA_OK(cudaMemcpyAsync(scratch + L.total, scratch + L.tile_scan + cdiv(T, kScanTile), sizeof(int32_t),
                                cudaMemcpyDeviceToDevice, s));
    GSB_CUDA_OK(cudaMemcpyAsync(scratch + L.n_vis, n_vis, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int ordered_fill(uint32_t C, uint32_t n, uint32_t per_cam, const int64_t *camera_ids, const float *means2d, const int32_t *radii,
                 uint32_t tile_size, uint32_t tw, uint32_t th, int64_t capacity, const int32_t *isect_offsets, int32_t *scratch,
                 int64_t *isect_ids, int32_t *flatten_ids, cudaStream_t s) {
    const uint32_t n_tiles = tw * th;
    const Layout L = make_layout(C, n, n_tiles);
    const uint32_t W = piece_warps(n_tiles);
    const size_t smem = (size_t)W * n_tiles * sizeof(int);
    (void)per_cam; (void)camera_ids;
    piece_kernel<true><<<cdiv(L.pieces_max, W), W * 32, smem, s>>>(
        C, n_tiles, bit_width(n_tiles), means2d, radii, (float)tile_size, tw, th,
        scratch + L.n_pieces, scratch + L.cam_start, scratch + L.piece_base,
        reinterpret_cast<const uint32_t *>(scratch + L.keys[0]), reinterpret_cast<const uint32_t *>(scratch + L.vals[0]),
        scratch + L.H, isect_offsets, capacity, isect_ids, flatten_ids);
    return check_launch("piece_scatter");
}

}  // namespace gsb200
