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
#include "common.cuh"
#include "scan.cuh"

namespace gsb200 {

struct TileBox { uint32_t x0, y0, x1, y1; };

// tile bbox of one Gaussian: min inclusive, max exclusive (isect_tiles.cu:59-71).  The float ->
// uint32 conversion saturates (negative -> 0) exactly like the reference's cast.
__device__ __forceinline__ TileBox tile_box(float mx, float my, float radius, float ts, uint32_t tw, uint32_t th) {
    const float tr = radius / ts, tx = mx / ts, ty = my / ts;
    TileBox b;
    b.x0 = min(__float2uint_rz(floorf(tx - tr)), tw);
    b.y0 = min(__float2uint_rz(floorf(ty - tr)), th);
    b.x1 = min(__float2uint_rz(ceilf(tx + tr)), tw);
    b.y1 = min(__float2uint_rz(ceilf(ty + tr)), th);
    return b;
}

// Warp-cooperative walk over the tile boxes of the 32 Gaussians a warp owns: for every Gaussian with a
// non-empty box the whole warp visits its tiles as a (rows x columns) patch -- lanes along x first, so
// one warp-wide atomic touches adjacent counters (one or two 32-byte sectors) instead of 32 unrelated
// ones, and large boxes (up to ~200 tiles) are spread over the lanes instead of serialising one thread.
// f(tile_index_within_camera, lane_is_active) is called convergently by all lanes.
template <typename F>
__device__ __forceinline__ void for_each_tile_coop(bool has, const TileBox &b, uint32_t tw, F &&f) {
    const uint32_t lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, has);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t x0 = __shfl_sync(0xffffffffu, b.x0, src), y0 = __shfl_sync(0xffffffffu, b.y0, src);
        const uint32_t nx = __shfl_sync(0xffffffffu, b.x1, src) - x0, ny = __shfl_sync(0xffffffffu, b.y1, src) - y0;
        for (uint32_t xb = 0; xb < nx; xb += 32) {
            const uint32_t w = min(32u, nx - xb);
            const uint32_t sh = w <= 1 ? 0 : 32 - __clz(w - 1);  // log2 of the padded patch width
            const uint32_t c = lane & ((1u << sh) - 1u), r = lane >> sh, rows = 32u >> sh;
            for (uint32_t yb = 0; yb < ny; yb += rows) {
                const bool on = c < w && yb + r < ny;
                f(src, (y0 + yb + r) * tw + x0 + xb + c, on);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
isect_count_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                   const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th,
                   int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ tile_counts) {
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
    const uint32_t n_tiles = tw * th;
    for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
        const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
        if (on) atomicAdd(tile_counts + (size_t)c * n_tiles + tile, 1);
    });
}

__global__ void __launch_bounds__(kThreads)
isect_scatter_kernel(uint32_t n, uint32_t per_cam, const int64_t *__restrict__ camera_ids, const float *__restrict__ means2d,
                     const int32_t *__restrict__ radii, const float *__restrict__ depths, float ts, uint32_t tw, uint32_t th,
                     int32_t *__restrict__ tile_cursor, uint64_t *__restrict__ bucket) {
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
    for_each_tile_coop(has, b, tw, [&](int src, uint32_t tile, bool on) {
        const uint32_t c = __shfl_sync(0xffffffffu, cam, src);
        const uint32_t d = __shfl_sync(0xffffffffu, depth_bits, src);
        const uint32_t id = __shfl_sync(0xffffffffu, i, src);
        if (on) {
            const int32_t slot = atomicAdd(tile_cursor + (size_t)c * n_tiles + tile, 1);
            bucket[slot] = ((uint64_t)d << 32) | (uint64_t)id;
        }
    });
}

// one compare-exchange sub-step of the normalized bitonic network over a[0..n): every comparison
// moves the smaller key to the lower index, so virtual +inf padding beyond n never moves and pairs
// whose upper index is >= n are simply skipped.
__device__ __forceinline__ void bitonic_substep(uint64_t *a, uint32_t n, uint32_t half_pairs, uint32_t k, uint32_t j,
                                                bool mirror) {
    for (uint32_t t = threadIdx.x; t < half_pairs; t += blockDim.x) {
        uint32_t lo, hi;
        if (mirror) {  // block of size k: r-th element pairs with its mirror image
            const uint32_t hk = k >> 1, blk = t / hk, r = t % hk;
            lo = blk * k + r;
            hi = blk * k + (k - 1 - r);
        } else {
            lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            hi = lo | j;
        }
        if (hi < n) {
            const uint64_t x = a[lo], y = a[hi];
            if (x > y) { a[lo] = y; a[hi] = x; }
        }
    }
}

__device__ __forceinline__ void bitonic_sort(uint64_t *a, uint32_t n) {
    uint32_t P = 1;
    while (P < n) P <<= 1;
    const uint32_t half_pairs = P >> 1;
    for (uint32_t k = 2; k <= P; k <<= 1) {
        bitonic_substep(a, n, half_pairs, k, 0, true);
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            bitonic_substep(a, n, half_pairs, k, j, false);
            __syncthreads();
        }
    }
}

constexpr uint32_t kSortSmall = 2048;   // items a 256-thread CTA sorts in 16 KB of static shared memory
constexpr uint32_t kSortLarge = 12288;  // items a 1024-thread CTA sorts in 96 KB of dynamic shared memory

// One CTA per (camera, tile).  LARGE = false handles segments of up to kSortSmall items, LARGE = true
// the (few) longer ones -- in shared memory up to kSortLarge items, in place in global memory (L2
// resident) beyond that.  Both kernels are launched over all tiles and exit early on the other class.
template <bool LARGE>
__global__ void __launch_bounds__(LARGE ? 1024 : kThreads)
tile_sort_kernel(uint32_t n_tiles_total, uint32_t n_tiles, uint32_t tile_bits, int64_t n_isects,
                 const int32_t *__restrict__ offsets, uint64_t *__restrict__ bucket, int64_t *__restrict__ isect_ids,
                 int32_t *__restrict__ flatten_ids) {
    extern __shared__ uint64_t s_dyn[];
    __shared__ uint64_t s_static[LARGE ? 1 : kSortSmall];
    const uint32_t ct = blockIdx.x;
    const int64_t start = offsets[ct];
    const int64_t end = (ct + 1 == n_tiles_total) ? n_isects : (int64_t)offsets[ct + 1];
    const uint32_t n = (uint32_t)(end - start);
    if (n == 0 || (LARGE ? n <= kSortSmall : n > kSortSmall)) return;
    const uint64_t cam = ct / n_tiles, tile = ct % n_tiles;
    const int64_t hi = (int64_t)((cam << (32 + tile_bits)) | (tile << 32));
    uint64_t *a;
    if (!LARGE || n <= kSortLarge) {
        a = LARGE ? s_dyn : s_static;
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) a[t] = bucket[start + t];
    } else {
        a = bucket + start;
    }
    __syncthreads();
    if (n > 1) bitonic_sort(a, n);
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        const uint64_t it = a[t];
        // low word: sign-extending OR like the reference's (int64)(int32)depth cast (isect_tiles.cu:88)
        isect_ids[start + t] = hi | (int64_t)(int32_t)(uint32_t)(it >> 32);
        flatten_ids[start + t] = (int32_t)(uint32_t)(it & 0xffffffffu);
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

static int exclusive_scan(uint32_t n, int32_t *data, int32_t *copy, int32_t *block_scratch, cudaStream_t s) {
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

extern "C" int gsb200_isect_prepare(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera, const int64_t *camera_ids,
                                    const float *means2d, const int32_t *radii, uint32_t tile_size, uint32_t tile_width,
                                    uint32_t tile_height, int32_t *tiles_per_gauss, int32_t *isect_offsets,
                                    int32_t *tile_cursor, int64_t *n_isects_host, gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t T = C * tile_width * tile_height;
    *n_isects_host = 0;
    if (bit_width(tile_width * tile_height) + bit_width(C) > 32)
        return fail(2, "gsb200_isect_prepare: camera and tile ids do not fit 32 bits");
    if (T == 0) return 0;
    GSB_CUDA_OK(cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * T, s));
    if (n_elems == 0) {
        GSB_CUDA_OK(cudaMemsetAsync(tile_cursor, 0, sizeof(int32_t) * T, s));
        return 0;
    }
    isect_count_kernel<<<cdiv(n_elems, kThreads), kThreads, 0, s>>>(n_elems, elems_per_camera ? elems_per_camera : 1,
                                                                     camera_ids, means2d, radii, (float)tile_size,
                                                                     tile_width, tile_height, tiles_per_gauss, isect_offsets);
    if (int e = check_launch("isect_count")) return e;
    // block sums scratch: reuse the head of tile_cursor is not possible (it receives the copy), so a
    // small side allocation from the stream-ordered pool is used.
    const uint32_t nb = cdiv(T, kScanTile);
    int32_t *scratch = nullptr;
    GSB_CUDA_OK(cudaMallocAsync((void **)&scratch, sizeof(int32_t) * (nb + 2), s));
    int e = exclusive_scan(T, isect_offsets, tile_cursor, scratch, s);
    int32_t total = 0;
    if (e == 0) {
        cudaError_t ce = cudaMemcpyAsync(&total, scratch + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess) e = fail(101, cudaGetErrorString(ce));
    }
    cudaFreeAsync(scratch, s);
    if (e) return e;
    *n_isects_host = total;
    return 0;
}

extern "C" int gsb200_isect_fill(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera, const int64_t *camera_ids,
                                 const float *means2d, const int32_t *radii, const float *depths, uint32_t tile_size,
                                 uint32_t tile_width, uint32_t tile_height, int64_t n_isects, const int32_t *isect_offsets,
                                 int32_t *tile_cursor, uint64_t *scratch, int64_t *isect_ids, int32_t *flatten_ids,
                                 gsb200_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_isects == 0 || n_elems == 0) return 0;
    const uint32_t n_tiles = tile_width * tile_height;
    isect_scatter_kernel<<<cdiv(n_elems, kThreads), kThreads, 0, s>>>(n_elems, elems_per_camera ? elems_per_camera : 1,
                                                                       camera_ids, means2d, radii, depths, (float)tile_size,
                                                                       tile_width, tile_height, tile_cursor, scratch);
    if (int e = check_launch("isect_scatter")) return e;
    tile_sort_kernel<false><<<C * n_tiles, kThreads, 0, s>>>(C * n_tiles, n_tiles, bit_width(n_tiles), n_isects,
                                                              isect_offsets, scratch, isect_ids, flatten_ids);
    if (int e = check_launch("tile_sort")) return e;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CUDA_OK(cudaFuncSetAttribute(tile_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSortLarge * sizeof(uint64_t))));
        attr_set = true;
    }
    tile_sort_kernel<true><<<C * n_tiles, 1024, kSortLarge * sizeof(uint64_t), s>>>(
        C * n_tiles, n_tiles, bit_width(n_tiles), n_isects, isect_offsets, scratch, isect_ids, flatten_ids);
    return check_launch("tile_sort_large");
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
