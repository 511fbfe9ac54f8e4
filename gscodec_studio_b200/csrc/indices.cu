// rasterize_to_indices_in_range: which Gaussians contribute to which pixels.
//
// Replaces _C.rasterize_to_indices_in_range (gsplat/cuda/csrc/rasterize_to_indices_in_range.cu:13-330, bindings.h:217).
// The reference's pure-torch rasterizer (_torch_impl.py:522-617) calls it in a loop over depth slices: slice
// [range_start, range_end) is counted in BATCHES of tile_size^2 intersections per tile, every pixel starts from the
// transmittance the previous slices left (`transmittances`) and stops for good once it would drop to <= 1e-4.
// Same per-pair arithmetic as the blend (pixel centre + 0.5, sigma expression, __expf, min(0.999), 1/255 cut).
//
// Two launches of one kernel, as the data-dependent output size demands: COUNT (hits per pixel) -> the caller
// scans the counts and allocates -> FILL (Gaussian id and global pixel id of every hit, per pixel front to back).
// One thread per pixel, the batch staged in shared memory as two float4 per Gaussian.
#include "common.cuh"

namespace gsb200 {

struct RangeGeom {
    uint32_t C, N, W, H, ts, tw, th;
    uint32_t range_start, range_end;
    int64_t n_isects;
};

template <bool FILL>
__global__ void indices_in_range_kernel(RangeGeom G, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                                        const float *__restrict__ opacities, const int32_t *__restrict__ tile_offsets,
                                        const int32_t *__restrict__ flatten_ids, const float *__restrict__ transmittances,
                                        const int64_t *__restrict__ chunk_starts, int32_t *__restrict__ chunk_cnts,
                                        int64_t *__restrict__ gaussian_ids, int64_t *__restrict__ pixel_ids) {
    extern __shared__ float4 smem4[];
    const uint32_t block_size = G.ts * G.ts;  // the slice unit of the reference: one batch = tile_size^2 intersections
    float4 *sA = smem4;                       // x, y, opacity, flatten id (bits)
    float4 *sB = sA + block_size;             // conic a, b, c, -
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
    if (G.range_start >= num_batches) {  // nothing of this tile lies in the slice
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
        const uint32_t idx = batch_start + tr;
        if (idx < (uint32_t)isect_end) {
            const int32_t g = flatten_ids[idx];
            const float2 xy = __ldg(means2d + g);
            sA[tr] = make_float4(xy.x, xy.y, __ldg(opacities + g), __int_as_float(g));
            sB[tr] = make_float4(__ldg(conics + (size_t)g * 3), __ldg(conics + (size_t)g * 3 + 1), __ldg(conics + (size_t)g * 3 + 2), 0.f);
        }
        __syncthreads();
        const uint32_t batch_size = min(block_size, (uint32_t)isect_end - batch_start);
        for (uint32_t t = 0; t < batch_size && !done; ++t) {
            const float4 A = sA[t], B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = 0.5f * (B.x * dx * dx + B.z * dy * dy) + B.y * dx * dy;
            const float alpha = fminf(0.999f, A.z * __expf(-sigma));
            if (sigma < 0.f || alpha < 1.f / 255.f) continue;
            const float next_trans = trans * (1.0f - alpha);
            if (next_trans <= 1e-4f) {  // this pixel is finished; the Gaussian that would finish it is not recorded
                done = true;
                break;
            }
            if (FILL) {
                gaussian_ids[base + cnt] = (int64_t)((uint32_t)__float_as_int(A.w) % G.N);
                pixel_ids[base + cnt] = (int64_t)pix;  // = pixel + camera * H * W
            }
            ++cnt;
            trans = next_trans;
        }
    }
    if (!FILL && inside) chunk_cnts[pix] = cnt;
}

static int check_range_args(const char *who, uint32_t C, uint32_t tile_size, uint32_t W, uint32_t H, uint32_t tw, uint32_t th) {
    if (tile_size == 0 || tile_size > 32) return failf(2, "%s: tile_size must be in [1, 32]", who);
    if (tw * tile_size < W || th * tile_size < H)
        return failf(2, "%s: tile grid does not cover the image (Assert Failed: tile_height * tile_size >= image_height)", who);
    if (th > 65535 || tw > 65535) return failf(2, "%s: tile grid too large", who);
    (void)C;
    return 0;
}

}  // namespace gsb200

using namespace gsb200;

extern "C" int gsb200_indices_in_range_count(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                             const float *means2d, const float *conics, const float *opacities, uint32_t W,
                                             uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                             const int32_t *tile_offsets, const int32_t *flatten_ids, const float *transmittances,
                                             int32_t *chunk_cnts, gsb200_stream_t stream) {
    if (int rc = check_range_args("gsb200_indices_in_range_count", C, tile_size, W, H, tile_width, tile_height)) return rc;
    if (C == 0 || W == 0 || H == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_isects == 0 || tile_width * tile_height == 0) {
        GSB_CUDA_OK(cudaMemsetAsync(chunk_cnts, 0, sizeof(int32_t) * (size_t)C * H * W, s));
        return 0;
    }
    const RangeGeom G{C, N, W, H, tile_size, tile_width, tile_height, range_start, range_end, n_isects};
    const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
    const size_t smem = (size_t)tile_size * tile_size * 2 * sizeof(float4);
    indices_in_range_kernel<false><<<blocks, threads, smem, s>>>(G, reinterpret_cast<const float2 *>(means2d), conics, opacities,
                                                                 tile_offsets, flatten_ids, transmittances, nullptr, chunk_cnts,
                                                                 nullptr, nullptr);
    return check_launch("indices_in_range_count");
}

extern "C" int gsb200_indices_in_range_fill(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                            const float *means2d, const float *conics, const float *opacities, uint32_t W,
                                            uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                            const int32_t *tile_offsets, const int32_t *flatten_ids, const float *transmittances,
                                            const int64_t *chunk_starts, int64_t *gaussian_ids, int64_t *pixel_ids,
                                            gsb200_stream_t stream) {
    if (int rc = check_range_args("gsb200_indices_in_range_fill", C, tile_size, W, H, tile_width, tile_height)) return rc;
    if (C == 0 || W == 0 || H == 0 || n_isects == 0 || tile_width * tile_height == 0) return 0;
    if (chunk_starts == nullptr || gaussian_ids == nullptr || pixel_ids == nullptr)
        return fail(2, "gsb200_indices_in_range_fill: null pointer");
    const RangeGeom G{C, N, W, H, tile_size, tile_width, tile_height, range_start, range_end, n_isects};
    const dim3 threads(tile_size, tile_size, 1), blocks(C, tile_height, tile_width);
    const size_t smem = (size_t)tile_size * tile_size * 2 * sizeof(float4);
    indices_in_range_kernel<true><<<blocks, threads, smem, (cudaStream_t)stream>>>(
        G, reinterpret_cast<const float2 *>(means2d), conics, opacities, tile_offsets, flatten_ids, transmittances, chunk_starts,
        nullptr, gaussian_ids, pixel_ids);
    return check_launch("indices_in_range_fill");
}
