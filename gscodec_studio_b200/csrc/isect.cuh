// Shared pieces of the tile-intersection stage (isect.cu: per-tile sort path; isect_ordered.cu: depth pre-sort +
// ordered scatter path).
#pragma once
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


// exclusive prefix of data[0..n) in place (+ optional copy); the total goes to block_scratch[cdiv(n, kScanTile)]
int exclusive_scan(uint32_t n, int32_t *data, int32_t *copy, int32_t *block_scratch, cudaStream_t s);

// ---- isect_ordered.cu ---------------------------------------------------------------------------------------------
struct OrderedPlan {
    bool ok;              // the ordered path can serve this configuration
    size_t scratch_ints;  // int32 entries of scratch it needs
    size_t total_index;   // where the intersection count lives inside the scratch
};
OrderedPlan ordered_plan(uint32_t C, uint32_t n_elems, uint32_t tile_width, uint32_t tile_height);
int ordered_prepare(uint32_t C, uint32_t n_elems, uint32_t per_cam, const int64_t *camera_ids, const float *means2d,
                    const int32_t *radii, const float *depths, uint32_t tile_size, uint32_t tw, uint32_t th,
                    int32_t *tiles_per_gauss, int32_t *isect_offsets, int32_t *scratch, cudaStream_t s);
int ordered_fill(uint32_t C, uint32_t n_elems, uint32_t per_cam, const int64_t *camera_ids, const float *means2d,
                 const int32_t *radii, uint32_t tile_size, uint32_t tw, uint32_t th, int64_t capacity,
                 const int32_t *isect_offsets, int32_t *scratch, int64_t *isect_ids, int32_t *flatten_ids, cudaStream_t s);

}  // namespace gsb200
