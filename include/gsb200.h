/*
 * gsb200.h -- C-ABI of the B200-native Gaussian-splat rasterization hot path.
 *
 * This is the drop-in boundary.  In the reference the same role is played by the pybind11 module
 * `_C` (gsplat/cuda/csrc/ext.cpp:3-92, signatures in gsplat/cuda/include/bindings.h:41-502) that
 * gsplat/cuda/_wrapper.py resolves by name (_wrapper.py:9-16).  Every entry point below names the
 * `_C.<symbol>` it replaces.  Differences forced by a C ABI (no torch types):
 *   - the CALLER allocates every output (the reference op allocates torch tensors itself,
 *     e.g. fully_fused_projection_fwd.cu:236-245); data-dependent sizes use a two-call protocol
 *     exactly where the reference synchronises with the host (isect_tiles.cu:199-200);
 *   - all pointers are device pointers to contiguous float32/int32/int64 arrays unless stated;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately,
 *     except gsb200_isect_prepare which returns the intersection count to the host;
 *   - return value: 0 on success, non-zero error code otherwise; gsb200_last_error() gives the
 *     message (the reference raises RuntimeError through TORCH_CHECK, bindings.h:10-16).
 * Gradient outputs are OVERWRITTEN (not accumulated) unless documented otherwise.
 */
#ifndef GSB200_H
#define GSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB200_ABI_VERSION 1

/* camera models: bindings.h:34-39 / ext.cpp:4-8 */
#define GSB200_CAMERA_PINHOLE 0
#define GSB200_CAMERA_ORTHO 1
#define GSB200_CAMERA_FISHEYE 2

/* per-(camera, tile) counters are kept in this many replicas to spread atomic traffic */
#define GSB200_ISECT_REPLICAS 32

typedef void *gsb200_stream_t;

int gsb200_abi_version(void);
const char *gsb200_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t gsb200_launch_count(void);

/* ---------------------------------------------------------------- stage 1: projection
 * replaces _C.fully_fused_projection_fwd (bindings.h:94-115, fully_fused_projection_fwd.cu:22-196).
 * means [N,3]; covars [N,6] (xx,xy,xz,yy,yz,zz) or NULL; quats [N,4] wxyz + scales [N,3] or NULL;
 * viewmats [C,4,4]; Ks [C,3,3].  Outputs radii i32 [C,N], means2d [C,N,2], depths [C,N],
 * conics [C,N,3], compensations [C,N] or NULL.  Culled entries get radii = 0 and ZERO in the other
 * outputs (the reference leaves them uninitialised).  packed_radius_rule != 0 selects the radius
 * formula of the packed reference kernel (fully_fused_projection_packed_fwd.cu:183-186). */
int gsb200_projection_fwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                          const float *quats, const float *scales, const float *viewmats,
                          const float *Ks, uint32_t image_width, uint32_t image_height, float eps2d,
                          float near_plane, float far_plane, float radius_clip, int camera_model,
                          int packed_radius_rule, int32_t *radii, float *means2d, float *depths,
                          float *conics, float *compensations, gsb200_stream_t stream);

/* replaces _C.fully_fused_projection_bwd (bindings.h:117-145, fully_fused_projection_bwd.cu:24-263).
 * v_means [N,3], v_covars [N,6] (covars path) or v_quats [N,4] + v_scales [N,3], v_viewmats [C,4,4]
 * or NULL.  All gradient outputs are overwritten with the sum over cameras. */
int gsb200_projection_bwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                          const float *quats, const float *scales, const float *viewmats,
                          const float *Ks, uint32_t image_width, uint32_t image_height, float eps2d,
                          int camera_model, const int32_t *radii, const float *conics,
                          const float *compensations, const float *v_means2d, const float *v_depths,
                          const float *v_conics, const float *v_compensations, float *v_means,
                          float *v_covars, float *v_quats, float *v_scales, float *v_viewmats,
                          gsb200_stream_t stream);

/* Packed (COO) projection, replaces _C.fully_fused_projection_packed_fwd / _bwd (bindings.h:253-309,
 * fully_fused_projection_packed_fwd.cu:200-247, _packed_bwd.cu:201-283) with memory proportional to nnz -- nothing of size
 * C*N is allocated.  Two passes, as in the reference: _count projects every (camera, Gaussian) pair and only counts the
 * visible ones per (camera, 256-Gaussian block), leaves the exclusive scan of those counts in block_offsets
 * (i32 [C * ceil(N/256) + 1 + ceil(C * ceil(N/256) / 2048) + 3], the total at index C * ceil(N/256)) and returns nnz to
 * the host (synchronises `stream`); _fill projects again and writes the rows in (camera, Gaussian) order: indptr i32 [C+1],
 * camera_ids / gaussian_ids i64 [nnz], radii i32 [nnz], means2d [nnz,2], depths [nnz], conics [nnz,3], compensations
 * [nnz] or NULL.  The packed radius rule (fully_fused_projection_packed_fwd.cu:183-186) applies.
 * _bwd: conics / compensations / v_* are the packed rows; gradients come back dense ([N,3], [N,6] | [N,4] + [N,3], [C,4,4])
 * summed over cameras by the thread that owns the Gaussian -- no atomics, no dense scatter of the incoming gradients. */
int gsb200_projection_packed_count(uint32_t C, uint32_t N, const float *means, const float *covars,
                                   const float *quats, const float *scales, const float *viewmats,
                                   const float *Ks, uint32_t image_width, uint32_t image_height,
                                   float eps2d, float near_plane, float far_plane, float radius_clip,
                                   int camera_model, int32_t *block_offsets, int64_t *nnz_host,
                                   gsb200_stream_t stream);
int gsb200_projection_packed_fill(uint32_t C, uint32_t N, const float *means, const float *covars,
                                  const float *quats, const float *scales, const float *viewmats,
                                  const float *Ks, uint32_t image_width, uint32_t image_height,
                                  float eps2d, float near_plane, float far_plane, float radius_clip,
                                  int camera_model, const int32_t *block_offsets, int32_t *indptr,
                                  int64_t *camera_ids, int64_t *gaussian_ids, int32_t *radii,
                                  float *means2d, float *depths, float *conics, float *compensations,
                                  gsb200_stream_t stream);
int gsb200_projection_packed_bwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                                 const float *quats, const float *scales, const float *viewmats,
                                 const float *Ks, uint32_t image_width, uint32_t image_height,
                                 float eps2d, int camera_model, const int32_t *block_offsets,
                                 const int64_t *gaussian_ids, const float *conics,
                                 const float *compensations, const float *v_means2d,
                                 const float *v_depths, const float *v_conics,
                                 const float *v_compensations, float *v_means, float *v_covars,
                                 float *v_quats, float *v_scales, float *v_viewmats,
                                 gsb200_stream_t stream);

/* packed (COO) projection: replaces _C.fully_fused_projection_packed_fwd
 * (bindings.h:253-276, fully_fused_projection_packed_fwd.cu:22-249).  Two calls:
 *   1. gsb200_compact_count: from dense radii [C,N] count visible pairs; writes block_offsets
 *      (scratch, (C*N+2047)/2048+1 int32) and returns nnz through *nnz_host (synchronises `stream`).
 *   2. gsb200_compact_fill: writes indptr i32 [C+1], camera_ids/gaussian_ids i64 [nnz] and gathers
 *      the dense rows into packed radii/means2d/depths/conics/compensations (row order =
 *      (camera, gaussian) lexicographic, as the reference emits). */
int gsb200_compact_count(uint32_t C, uint32_t N, const int32_t *radii, int32_t *block_offsets,
                         int64_t *nnz_host, gsb200_stream_t stream);
int gsb200_compact_fill(uint32_t C, uint32_t N, const int32_t *radii, const float *means2d,
                        const float *depths, const float *conics, const float *compensations,
                        const int32_t *block_offsets, int32_t *indptr, int64_t *camera_ids,
                        int64_t *gaussian_ids, int32_t *p_radii, float *p_means2d, float *p_depths,
                        float *p_conics, float *p_compensations, gsb200_stream_t stream);
/* ---- 2D Gaussian (surfel) family ----------------------------------------------------------------------
 * replaces _C.fully_fused_projection_fwd_2dgs / _bwd_2dgs (bindings.h:311-349; kernels fully_fused_projection_2dgs_{fwd,bwd}.cu,
 * include/2dgs.cuh).  radii i32 [C,N], means2d [C,N,2], depths [C,N], ray_transforms [C,N,3,3] (rows of K [t_u t_v mean_c]),
 * normals [C,N,3] (camera space, flipped towards the camera).  Culled entries are zero-filled.  `scales` is [N,3]; only the
 * first two columns are used and v_scales[:,2] = 0.  No camera gradients (the reference has none either). */
int gsb200_projection_2dgs_fwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                               const float *viewmats, const float *Ks, uint32_t image_width, uint32_t image_height,
                               float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
                               float *depths, float *ray_transforms, float *normals, gsb200_stream_t stream);
int gsb200_projection_2dgs_bwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                               const float *viewmats, const float *Ks, const int32_t *radii, const float *ray_transforms,
                               const float *v_means2d, const float *v_depths, const float *v_normals,
                               const float *v_ray_transforms, float *v_means, float *v_quats, float *v_scales,
                               gsb200_stream_t stream);

/* replaces _C.rasterize_to_pixels_fwd_2dgs / _bwd_2dgs (bindings.h:351-412; kernels rasterize_to_pixels_2dgs_{fwd,bwd}.cu).
 * M = C*N or nnz rows; D colour channels in {1,2,3,4,8,16,32}, the LAST channel is read as depth by the distortion and
 * median-depth outputs.  Outputs: render_colors [C,H,W,D], render_alphas / render_distort / render_median [C,H,W,1],
 * render_normals [C,H,W,3], last_ids / median_ids i32 [C,H,W].  The backward zero-fills and accumulates the gradient
 * tensors (v_means2d_abs may be NULL); v_densify [M,2] (may be NULL) = (v_ray_transforms[.,0,2], v_ray_transforms[.,1,2]) *
 * ray_transforms[.,2,2] from the finished sums. */
int gsb200_rasterize_2dgs_fwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                              const float *ray_transforms, const float *colors, const float *opacities,
                              const float *normals, const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                              uint32_t image_height, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                              const int32_t *tile_offsets, const int32_t *flatten_ids, float *render_colors,
                              float *render_alphas, float *render_normals, float *render_distort, float *render_median,
                              int32_t *last_ids, int32_t *median_ids, gsb200_stream_t stream);
int gsb200_rasterize_2dgs_bwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                              const float *ray_transforms, const float *colors, const float *opacities,
                              const float *normals, const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                              uint32_t image_height, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                              const int32_t *tile_offsets, const int32_t *flatten_ids, const float *render_colors,
                              const float *render_alphas, const int32_t *last_ids, const int32_t *median_ids,
                              const float *v_render_colors, const float *v_render_alphas, const float *v_render_normals,
                              const float *v_render_distort, const float *v_render_median, float *v_means2d_abs,
                              float *v_means2d, float *v_ray_transforms, float *v_colors, float *v_opacities,
                              float *v_normals, float *v_densify, gsb200_stream_t stream);

/* replaces _C.rasterize_to_indices_in_range_2dgs (bindings.h:414-429): the two-call protocol of
 * gsb200_indices_in_range_count / _fill with the surfel alpha. */
int gsb200_indices_in_range_2dgs_count(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                       const float *means2d, const float *ray_transforms, const float *opacities,
                                       uint32_t image_width, uint32_t image_height, uint32_t tile_size, uint32_t tile_width,
                                       uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                       const float *transmittances, int32_t *chunk_cnts, gsb200_stream_t stream);
int gsb200_indices_in_range_2dgs_fill(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, int64_t n_isects,
                                      const float *means2d, const float *ray_transforms, const float *opacities,
                                      uint32_t image_width, uint32_t image_height, uint32_t tile_size, uint32_t tile_width,
                                      uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                      const float *transmittances, const int64_t *chunk_starts, int64_t *gaussian_ids,
                                      int64_t *pixel_ids, gsb200_stream_t stream);

/* ---- quantise on load: compression_simulation fused into the attribute load --------------------------------------------------
 * The reference fake-quantises every attribute with its own chain of torch ops and hands the quantised copies to
 * rasterization() (compression_simulation/ops.py:39-54 called from simulation.py:242-316; then torch.exp / torch.sigmoid,
 * examples/simple_trainer.py:758-776).  These entry points take the trainer's RAW parameters plus one descriptor per quantised
 * attribute and apply clamp(x, lo, hi) + u * q_step (noise mode; u from Philox (seed, offset) -- one counter per 4 consecutive
 * elements, exactly gsb200_fake_quantize_fwd's -- or from `noise`) and the activation where the value is loaded; the backward
 * returns the gradients of the RAW parameters (clamp gate, activation derivative).  No quantised tensor exists in HBM.
 * A NULL descriptor: that attribute is read as it is.  act: 0 none, 1 exp, 2 sigmoid. */
typedef struct gsb200_quant_load {
    const float *noise;  /* explicit U(-.5,.5) draw with the attribute's shape, or NULL */
    uint64_t seed, offset;
    float lo, hi;
    int32_t bitwidth, act;
} gsb200_quant_load;

/* gsb200_projection_fwd (dense) on raw quats [N,4] / scales [N,3] (+ optionally raw opacities [N] -> opacities_act [N], the
 * blend's opacity input; both NULL to skip). */
int gsb200_projection_quant_fwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                const float *opacities_raw, const gsb200_quant_load *q_quats,
                                const gsb200_quant_load *q_scales, const gsb200_quant_load *q_opacities,
                                const float *viewmats, const float *Ks, uint32_t image_width, uint32_t image_height,
                                float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                                int32_t *radii, float *means2d, float *depths, float *conics, float *compensations,
                                float *opacities_act, gsb200_stream_t stream);
/* gsb200_projection_bwd for the call above; v_opacities_act [N] is the gradient of opacities_act (summed over cameras),
 * v_opacities_raw [N] may be NULL.  No camera gradients in this form. */
int gsb200_projection_quant_bwd(uint32_t C, uint32_t N, const float *means, const float *quats, const float *scales,
                                const float *opacities_raw, const gsb200_quant_load *q_quats,
                                const gsb200_quant_load *q_scales, const gsb200_quant_load *q_opacities,
                                const float *viewmats, const float *Ks, uint32_t image_width, uint32_t image_height,
                                float eps2d, int camera_model, const int32_t *radii, const float *conics,
                                const float *compensations, const float *v_means2d, const float *v_depths,
                                const float *v_conics, const float *v_compensations, const float *opacities_act,
                                const float *v_opacities_act, float *v_means, float *v_quats, float *v_scales,
                                float *v_opacities_raw, gsb200_stream_t stream);

/* gradient scatter back to the dense layout used by gsb200_projection_bwd:
 * dense v_means2d/v_depths/v_conics/v_compensations [C,N,*] are zero-filled then row r of the packed
 * gradients is written at (camera_ids[r], gaussian_ids[r]).
 * (stands in for _C.fully_fused_projection_packed_bwd, bindings.h:278-309, dense-gradient mode). */
int gsb200_scatter_rows(uint32_t C, uint32_t N, int64_t nnz, const int64_t *camera_ids,
                        const int64_t *gaussian_ids, const float *p_v_means2d,
                        const float *p_v_depths, const float *p_v_conics,
                        const float *p_v_compensations, float *v_means2d, float *v_depths,
                        float *v_conics, float *v_compensations, gsb200_stream_t stream);

/* ---------------------------------------------------------------- colour: spherical harmonics
 * replaces _C.compute_sh_fwd / _C.compute_sh_bwd (bindings.h:234-248, compute_sh_fwd.cu:11-37,
 * compute_sh_bwd.cu:13-51).  dirs [n,3] (un-normalised), coeffs [n,K,3], masks u8 [n] or NULL,
 * colors [n,3].  Masked-out rows: colors = 0, v_coeffs = 0, v_dirs = 0.  Launched on `stream`
 * (the reference uses the legacy default stream). */
int gsb200_sh_fwd(uint32_t n, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                  const uint8_t *masks, float *colors, gsb200_stream_t stream);
int gsb200_sh_bwd(uint32_t n, uint32_t K, uint32_t degree, const float *dirs, const float *coeffs,
                  const uint8_t *masks, const float *v_colors, float *v_coeffs, float *v_dirs,
                  gsb200_stream_t stream);

/* fused view-dependent colour for the dense layout: for every camera c and Gaussian g with radii > 0
 *   colors[c,g,:] = max(SH_degree(means[g] - campos[c]) . coeffs[g] + 0.5, 0)      (0 where radii <= 0)
 * i.e. gsplat/rendering.py:370-392 (dirs, expand, compute_sh_fwd, + 0.5, clamp_min) in one pass over
 * coeffs [N,K,3] shared by all cameras.  campos [C,3] are the camera centres.  The backward overwrites
 * v_coeffs [N,K,3] (summed over cameras) and v_means [N,3] (may be NULL) and, when not NULL, v_campos
 * [C,3] (= -sum of the direction gradients). */
int gsb200_sh_view_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                       const float *campos, const float *coeffs, const int32_t *radii, float *colors,
                       gsb200_stream_t stream);
int gsb200_sh_view_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                       const float *campos, const float *coeffs, const int32_t *radii,
                       const float *v_colors, float *v_coeffs, float *v_means, float *v_campos,
                       gsb200_stream_t stream);

/* Split-coefficient form of the two calls above: the DC band sh0 [N,1,3] and the higher bands shN [N,K-1,3]
 * stay two arrays (the trainers' parameters, examples/simple_trainer.py:776 concatenates them every step only
 * to build the [N,K,3] argument of rasterization(); gsplat/rendering.py:386 then consumes that copy).  K counts
 * sh0 too, K - 1 <= GSB200_SH_SPLIT_MAX_REST.  The backward overwrites v_sh0 [N,1,3] and v_shN [N,K-1,3]
 * (16-byte aligned; zero rows for invisible Gaussians and unused bands included), v_means / v_campos as above. */
#define GSB200_SH_SPLIT_MAX_REST 31
int gsb200_sh_view_split_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                             const float *campos, const float *sh0, const float *shN, const int32_t *radii,
                             float *colors, gsb200_stream_t stream);
int gsb200_sh_view_split_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                             const float *campos, const float *sh0, const float *shN, const int32_t *radii,
                             const float *v_colors, float *v_sh0, float *v_shN, float *v_means,
                             float *v_campos, gsb200_stream_t stream);

/* gsb200_sh_view_split_fwd / _bwd with the raw sh0 fake-quantised where it is loaded (q_sh0 NULL: identical to them) */
int gsb200_sh_view_split_quant_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                   const float *campos, const float *sh0, const gsb200_quant_load *q_sh0, const float *shN,
                                   const int32_t *radii, float *colors, gsb200_stream_t stream);
int gsb200_sh_view_split_quant_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                                   const float *campos, const float *sh0, const gsb200_quant_load *q_sh0, const float *shN,
                                   const int32_t *radii, const float *v_colors, float *v_sh0, float *v_shN, float *v_means,
                                   float *v_campos, gsb200_stream_t stream);

/* View-parallel multi-GPU form of the same stage.  The reference all-gathers / all-to-alls whole attribute
 * tensors between ranks around these ops (rendering.py:282-294, 397-478; distributed.py:90-215).  Here the SH
 * coefficients -- 16x3 floats per Gaussian, the bulk of the exchange -- stay on their owner: rank r publishes
 * its block in peer-accessible (NVLink-mapped) memory and the kernels PULL the rows of visible Gaussians only.
 *   shard_coeffs : HOST array of `world` device pointers; shard_coeffs[r] = rank r's [n_block, K, 3] block as
 *                  mapped into this process.  Gathered row g lives at shard_coeffs[g / n_block][g % n_block].
 *   N            : rows of the gathered set, N <= world * n_block.
 *   row_cache    : optional local [N, K, 3] array; the forward stores every row it pulled there (rows of invisible
 *                  Gaussians stay untouched) and the backward, given the same array, reads it instead of pulling
 *                  the rows over the link a second time.
 * The backward PUSHES: rank `rank` stores the coefficient-gradient row of every visible gathered Gaussian g into
 * slot [rank][g % n_block] of the staging array of g's owner, and a 0/1 flag for EVERY row next to it:
 *   peer_rows[o]  : owner o's staging array [world, n_block, K*3] as mapped into this process (HOST array of ptrs)
 *   peer_flags[o] : owner o's flags [world, n_block] (int32)
 * gsb200_shard_reduce_rows is the owner-side half of the reduce-scatter, on LOCAL memory:
 *   out[i] = sum over source ranks p (in rank order) of staging[p][i] where flags[p][i] != 0,  i < n_rows.
 * The caller orders the phases across ranks (publish -> barrier -> pull; push -> barrier -> reduce -> barrier). */
#define GSB200_MAX_SHARDS 16
#define GSB200_MAX_VIEWS 64 /* views of the whole job (ranks x cameras per rank) in the owner-side colour exchange */
int gsb200_sh_view_fwd_sharded(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                               const float *campos, const float *const *shard_coeffs, uint32_t world,
                               uint32_t n_block, const int32_t *radii, float *colors, float *row_cache,
                               gsb200_stream_t stream);
int gsb200_sh_view_bwd_sharded(uint32_t C, uint32_t N, uint32_t K, uint32_t degree, const float *means,
                               const float *campos, const float *const *shard_coeffs, uint32_t world,
                               uint32_t n_block, uint32_t rank, const float *row_cache, const int32_t *radii,
                               const float *v_colors,
                               float *const *peer_rows, int32_t *const *peer_flags, float *v_means,
                               float *v_campos, gsb200_stream_t stream);
int gsb200_shard_reduce_rows(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_floats,
                             const float *staging, const int32_t *flags, float *out, gsb200_stream_t stream);
/* same reduction written to two tensors: floats [0, split) of a row to out0 [n_rows, split], the rest to out1
 * [n_rows, row_floats - split] -- the gradients of the trainers' two SH parameters sh0 / shN (no [n,K,3] detour) */
int gsb200_shard_reduce_rows_split(uint32_t world, uint32_t n_block, uint32_t n_rows, uint32_t row_floats,
                                   uint32_t split, const float *staging, const int32_t *flags, float *out0,
                                   float *out1, gsb200_stream_t stream);
/* All-gather by pull, used for the geometry attributes (needed for EVERY row, so nothing to skip): out
 * [world, block_floats] <- blocks[r][0 .. block_floats) for every rank r (HOST array of peer-mapped device
 * pointers).  Replaces the all-gather of all_gather_tensor_list (distributed.py:102-167) for that tensor. */
int gsb200_shard_gather_rows(uint32_t world, uint64_t block_floats, const float *const *blocks, float *out,
                             gsb200_stream_t stream);
/* Geometry exchange without the gathered [N, w] intermediate.  A rank's published geometry block holds the attribute
 * tensors one after the other (means [n_block,3] | quats [n_block,4] | scales [n_block,3] | opacity [n_block,1]: widths
 * {3,4,3,1}; or means | covars | opacity), n_block a multiple of 4, so that every transfer is a run of 16-byte vectors that
 * is contiguous on both sides; the staging slots of the backward use the same layout.
 *   _gather_geo : pulls every peer's block and writes the n_parts attribute tensors of the gathered set,
 *                 outs[k] [world * n_block, widths[k]] (rank r's rows at r * n_block) -- HOST arrays of device pointers.
 *   _push_geo   : backward, sender side: grads[k] [world * n_block, widths[k]] (NULL = zero gradient); the rows of owner p
 *                 are stored, as one dense [n_block * w] block, into p's staging array peer_staging[p]
 *                 [world, slot_floats] at slot `rank` (slot_floats >= n_block * w, a multiple of 4).
 *   _reduce_geo : backward, owner side (after a cross-rank barrier): sums the world slots of the local staging array in
 *                 rank order (reproducible) into outs[k] [n_rows, widths[k]] (NULL = not wanted).
 * Replaces the reduce-scatter of all_gather_tensor_list's backward (gsplat/distributed.py:102-167) for the geometry. */
/* Owner-side colour exchange (view-parallel rendering, coefficients sharded by owner).  Instead of pulling the
 * 16x3-float coefficient rows of visible Gaussians to the viewer, the OWNER of a shard evaluates its rows for every view
 * of the job and only colours (forward) and colour gradients (backward) -- 12 bytes per visible (view, Gaussian) pair --
 * cross NVLink; the coefficient gradients never leave their owner.  V = ranks x cameras per rank views, view v belongs to
 * rank v / C.  All staging arrays live in peer-mapped memory; barriers between the steps are the caller's
 * (gsb200_peer_barrier).
 *   _shard_push_rows : viewer side.  src [n_slots][world * n_block, width] (a rank's radii [C, N] as bit patterns, or its
 *                      v_colors [C, N, 3]); the rows of owner p travel as one dense block into peer_staging[p] at slot
 *                      (slot0 + c), slot_floats apart; extra_src [n_slots, extra_w] (camera centres; NULL = none) goes to
 *                      peer_extra[p] + (slot0 + c) * extra_w.  HOST arrays of device pointers.
 *   _sh_owner_fwd    : owner side, after a barrier.  means [n_local,3], sh0 [n_local,1,3], shN [n_local,K-1,3] local;
 *                      campos [V,3], masks [V, n_block] (> 0 = visible) local staging; view_colors[v] = address of this
 *                      owner's [n_block,3] slice inside view v's colour tensor (HOST array of V device pointers).
 *                      Invisible pairs are not written.
 *   _sh_owner_bwd    : owner side, after the viewers pushed v_colors [V, n_block, 3] and a barrier.  Overwrites v_sh0,
 *                      v_shN (dense, 16-byte aligned) and v_means [n_local,3] (NULL = not wanted), summing the views in
 *                      order (reproducible).  Camera-centre gradients are not produced on this path.
 * Replaces the coefficient all-to-all of the reference's distributed mode (rendering.py:397-478). */
/* Owner-side projection, same idea for stage 1 / 4b: every rank projects ITS rows for all V views (the cameras were
 * exchanged with gsb200_peer_push_small + a barrier) and stores radius (always), 2-D mean, depth, conic and opacity (visible
 * pairs only) into the viewers' tensors: view_X[v] = address of this owner's n_block rows inside view v's X tensor (HOST
 * arrays of V device pointers into peer memory).  radii_local [V, n_block] / conics_local [V, n_block, 3] keep what the
 * owner's backward needs.  _owner_bwd: v_means2d [V, n_block, 2], v_conics [V, n_block, 3], v_depths [V, n_block] (NULL =
 * none) were pushed by the viewers (gsb200_shard_push_rows); gradients of the local means / quats / scales are written once,
 * summed over the views in order, no atomics.  (quats, scales) parameterisation, no compensations, no camera gradients --
 * the caller falls back to the gather path for anything else.  Replaces fully_fused_projection_{fwd,bwd} + the attribute
 * all-gather / gradient reduce-scatter of the reference's distributed mode (rendering.py:282-294, distributed.py:102-167). */
int gsb200_projection_owner_fwd(uint32_t V, uint32_t n_local, uint32_t n_block, const float *means, const float *quats,
                                const float *scales, const float *opacities, const float *viewmats, const float *Ks,
                                uint32_t W, uint32_t H, float eps2d, float near_plane, float far_plane,
                                float radius_clip, int camera_model, int32_t *radii_local, float *conics_local,
                                int32_t *const *view_radii, float *const *view_means2d, float *const *view_depths,
                                float *const *view_conics, float *const *view_opacities, gsb200_stream_t stream);
int gsb200_projection_owner_bwd(uint32_t V, uint32_t n_local, uint32_t n_block, const float *means, const float *quats,
                                const float *scales, const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                float eps2d, int camera_model, const int32_t *radii_local, const float *conics_local,
                                const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means,
                                float *v_quats, float *v_scales, gsb200_stream_t stream);
/* src [n_slots, w] floats (a rank's cameras) -> peer_dst[p] + slot0 * w on every peer p (HOST array of device pointers). */
int gsb200_peer_push_small(uint32_t world, uint32_t slot0, uint32_t n_slots, uint32_t w, const float *src,
                           float *const *peer_dst, gsb200_stream_t stream);
int gsb200_shard_push_rows(uint32_t world, uint32_t n_block, uint32_t width, uint32_t n_slots, uint32_t slot0,
                           uint64_t slot_floats, const float *src, float *const *peer_staging, const float *extra_src,
                           uint32_t extra_w, float *const *peer_extra, gsb200_stream_t stream);
int gsb200_sh_owner_fwd(uint32_t V, uint32_t n_local, uint32_t n_block, uint32_t K, uint32_t degree, const float *means,
                        const float *campos, const float *sh0, const float *shN, const int32_t *masks,
                        float *const *view_colors, gsb200_stream_t stream);
int gsb200_sh_owner_bwd(uint32_t V, uint32_t n_local, uint32_t n_block, uint32_t K, uint32_t degree, const float *means,
                        const float *campos, const float *sh0, const float *shN, const int32_t *masks,
                        const float *v_colors, float *v_sh0, float *v_shN, float *v_means, gsb200_stream_t stream);
/* Camera centres campos [C,3] = -R^T t of rigid world-to-camera matrices viewmats [C,4,4] (row major): the
 * `torch.inverse(viewmats)[:, :3, 3]` of the reference's colour stage (rendering.py:370) for rigid matrices. */
int gsb200_camera_positions(uint32_t C, const float *viewmats, float *campos, gsb200_stream_t stream);
/* Start of a distributed step, one launch: this rank's attribute tensors parts[k] [n_local, widths[k]] are written side by
 * side into its published geometry block geo_block (n_block * w floats: part k's [n_block, widths[k]] rows, dense, one part
 * after the other; n_block a multiple of 4; rows >= n_local = the matching columns of pad_row [w], an inert splat; NULL = 0),
 * and its coefficient rows -- `first` [n_local, first_floats] next to `rest` [n_local, row_floats - first_floats] (NULL when
 * first covers the row): the trainers' sh0 / shN -- into coeff_block [n_local, row_floats].  n_parts = 0 skips the
 * geometry, first = NULL the coefficients.  Replaces the torch.cat / copies in front of the reference's all-gather
 * (rendering.py:282-294). */
int gsb200_shard_publish(uint32_t n_local, uint32_t n_block, uint32_t n_parts, const uint32_t *widths,
                         float *const *parts, const float *pad_row, float *geo_block, const float *first,
                         const float *rest, uint32_t first_floats, uint32_t row_floats, float *coeff_block,
                         gsb200_stream_t stream);
/* Stream-ordered barrier over the `world` ranks of one node.  pads[r]: rank r's signal pad (GSB200_MAX_SHARDS zero-
 * initialised uint32 words in peer-mapped memory; a HOST array of device pointers); every rank passes the same, strictly
 * increasing `epoch` for the same barrier.  The host does not block; kernels enqueued after it see every peer's writes
 * made before its call.  Stands in for the implicit synchronisation of the reference's NCCL collectives
 * (gsplat/distributed.py:90-215). */
int gsb200_peer_barrier(uint32_t world, uint32_t rank, uint32_t *const *pads, uint32_t epoch, gsb200_stream_t stream);
int gsb200_shard_gather_geo(uint32_t world, uint32_t n_block, uint32_t n_parts, const uint32_t *widths,
                            const float *const *blocks, float *const *outs, gsb200_stream_t stream);
int gsb200_shard_push_geo(uint32_t world, uint32_t n_block, uint32_t rank, uint64_t slot_floats,
                          uint32_t n_parts, const uint32_t *widths, float *const *grads,
                          float *const *peer_staging, gsb200_stream_t stream);
int gsb200_shard_reduce_geo(uint32_t world, uint32_t n_block, uint32_t n_rows, uint64_t slot_floats,
                            uint32_t n_parts, const uint32_t *widths, const float *staging,
                            float *const *outs, gsb200_stream_t stream);

/* ---------------------------------------------------------------- stage 2: tile intersection
 * replaces _C.isect_tiles + _C.isect_offset_encode (bindings.h:147-166, isect_tiles.cu:16-389).  Results are
 * bit-identical to the reference's stable global radix sort.  Two implementations behind the same calls:
 * the per-tile sort path (csrc/isect.cu: tile histogram -> scatter -> one CTA sorts each tile; default) and the ordered
 * path (csrc/isect_ordered.cu: depth pre-sort of the visible Gaussians + ordered scatter; GSB200_ISECT=ordered, validated
 * bit-exact but currently slower).
 *
 *   0. gsb200_isect_plan: size (int32 entries) of the scratch buffer `tile_cursor` the other two calls need for this
 *      configuration, and the index inside it where the intersection count is left on the device.
 *   1. gsb200_isect_prepare: n_elems = C*N (dense) or nnz (packed; camera_ids i64 [nnz] required).
 *      Writes tiles_per_gauss i32 [n_elems] and isect_offsets i32 [C*tile_h*tile_w]; returns the
 *      number of intersections through *n_isects_host (synchronises `stream`, like the reference's
 *      .item() at isect_tiles.cu:200).  With n_isects_host == NULL the call does not synchronise; the
 *      count is then only available on the device, at tile_cursor[total_index].
 *   2. gsb200_isect_fill: writes isect_ids i64 [I] and flatten_ids i32 [I], sorted.  scratch: u64 [I] (only the
 *      per-tile sort path uses it).
 *      n_isects >= 0: exact count (buffers hold I entries).  n_isects < 0: "capacity mode" -- buffers hold
 *      -n_isects entries, the true count is read from the device; if it exceeds the capacity the outputs
 *      are incomplete and the caller must call again with larger buffers once it knows the count
 *      (lets the host size buffers speculatively and fetch the count without idling the GPU). */
int gsb200_isect_plan(uint32_t C, uint32_t n_elems, uint32_t tile_width, uint32_t tile_height,
                      int64_t *scratch_ints, int64_t *total_index);
int gsb200_isect_prepare(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera,
                         const int64_t *camera_ids, const float *means2d, const int32_t *radii,
                         const float *depths, uint32_t tile_size, uint32_t tile_width,
                         uint32_t tile_height, int32_t *tiles_per_gauss, int32_t *isect_offsets,
                         int32_t *tile_cursor, int64_t *n_isects_host, gsb200_stream_t stream);
/* Capacity mode helper: asynchronous 4-byte copy of the intersection count left by gsb200_isect_prepare at
 * tile_cursor[total_index] (gsb200_isect_plan) into PINNED host memory; the host polls the word or waits on an event. */
int gsb200_isect_count_async(const int32_t *tile_cursor, int64_t total_index, int32_t *count_pinned,
                             gsb200_stream_t stream);
int gsb200_isect_fill(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera,
                      const int64_t *camera_ids, const float *means2d, const int32_t *radii,
                      const float *depths, uint32_t tile_size, uint32_t tile_width,
                      uint32_t tile_height, int64_t n_isects, const int32_t *isect_offsets,
                      int32_t *tile_cursor, uint64_t *scratch, int64_t *isect_ids,
                      int32_t *flatten_ids, gsb200_stream_t stream);
/* `sort = false` form of _C.isect_tiles: keys in emission order (ascending flatten index, row-major tiles per
 * Gaussian).  Call after gsb200_isect_prepare (which gives tiles_per_gauss and the count);
 * cum_scratch: i32 [n_elems + n_elems / 2048 + 3]. */
int gsb200_isect_emit_unsorted(uint32_t C, uint32_t n_elems, uint32_t elems_per_camera,
                               const int64_t *camera_ids, const float *means2d, const int32_t *radii,
                               const float *depths, uint32_t tile_size, uint32_t tile_width,
                               uint32_t tile_height, const int32_t *tiles_per_gauss,
                               int32_t *cum_scratch, int64_t *isect_ids, int32_t *flatten_ids,
                               gsb200_stream_t stream);
/* stand-alone _C.isect_offset_encode (isect_tiles.cu:308-389) for callers that hold sorted ids */
int gsb200_isect_offset_encode(int64_t n_isects, const int64_t *isect_ids, uint32_t C,
                               uint32_t tile_width, uint32_t tile_height, int32_t *isect_offsets,
                               gsb200_stream_t stream);

/* ---------------------------------------------------------------- stage 3 / 4a: alpha blend
 * replaces _C.rasterize_to_pixels_fwd (bindings.h:168-184, rasterize_to_pixels_fwd.cu:16-185).
 * means2d [M,2], conics [M,3], colors [M,D], opacities [M] with M = C*N or nnz; backgrounds [C,D] or
 * NULL; masks u8 [C,th,tw] or NULL; tile_offsets i32 [C,th,tw]; flatten_ids i32 [I].
 * Outputs render_colors [C,H,W,D], render_alphas [C,H,W], last_ids i32 [C,H,W].
 * D must be in {1,2,3,4,5,8,9,16,17,32,33} (the host pads others, as _wrapper.py:496-541 does).
 * tile_order (optional, may be NULL): i32 [C*th*tw] scratch.  The forward fills it with the (camera, tile)
 * indices sorted longest-list-first and schedules its CTAs in that order (no straggler tail); pass the same
 * buffer to the backward.
 * blend_stream / tile_counts (optional, may be NULL): float [n_isects * gsb200_blend_stream_record(D, tile_size)] and
 * i32 [C*th*tw].  When the record size is non-zero (16x16 tiles, D <= 4) the forward leaves its pre-culled,
 * front-to-back per-tile record stream there (tile t's records start at tile_offsets[t]) and the backward stages it
 * with bulk copies instead of gathering through flatten_ids again; NULL in the backward selects the gathering kernel. */
uint32_t gsb200_blend_stream_record(uint32_t D, uint32_t tile_size);
/* 0 (default): stream kernels where supported; 1: the generic per-tile kernels everywhere (also GSB200_BLEND=tile).
 * Returns the previous mode.  Process-wide; meant for A/B timing and parity tests. */
int gsb200_set_blend_mode(int mode);
int gsb200_rasterize_fwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                         const float *conics, const float *colors, const float *opacities,
                         const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                         uint32_t image_height, uint32_t tile_size, uint32_t tile_width,
                         uint32_t tile_height, const int32_t *tile_offsets,
                         const int32_t *flatten_ids, float *render_colors, float *render_alphas,
                         int32_t *last_ids, int32_t *tile_order, float *blend_stream,
                         int32_t *tile_counts, gsb200_stream_t stream);

/* replaces _C.rasterize_to_pixels_bwd (bindings.h:186-215, rasterize_to_pixels_bwd.cu:17-277).
 * Gradient outputs v_means2d [M,2], v_conics [M,3], v_colors [M,D], v_opacities [M] and, if not
 * NULL, v_means2d_abs [M,2] are overwritten.  grad_scratch: float [M * gsb200_grad_record(D)],
 * used for the packed per-Gaussian accumulation records. */
uint32_t gsb200_grad_record(uint32_t D);
int gsb200_rasterize_bwd(uint32_t C, uint32_t M, int64_t n_isects, uint32_t D, const float *means2d,
                         const float *conics, const float *colors, const float *opacities,
                         const float *backgrounds, const uint8_t *masks, uint32_t image_width,
                         uint32_t image_height, uint32_t tile_size, uint32_t tile_width,
                         uint32_t tile_height, const int32_t *tile_offsets,
                         const int32_t *flatten_ids, const float *render_alphas,
                         const int32_t *last_ids, const float *v_render_colors,
                         const float *v_render_alphas, const int32_t *tile_order,
                         const float *blend_stream, const int32_t *tile_counts, float *grad_scratch,
                         float *v_means2d_abs,
                         float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                         gsb200_stream_t stream);

/* replaces _C.rasterize_to_indices_in_range (bindings.h:217-232, rasterize_to_indices_in_range.cu:13-330), the op the
 * reference's pure-torch rasterizer loops over (_torch_impl.py:522-617).  The slice [range_start, range_end) counts
 * BATCHES of tile_size^2 intersections per tile; transmittances [C,H,W] is what the previous slices left.  Two calls
 * where the reference synchronises for the output size: _count fills chunk_cnts i32 [C,H,W] (hits per pixel); the
 * caller turns them into exclusive starts (int64 [C,H,W]) and allocates; _fill writes, per pixel front to back,
 * gaussian_ids i64 [M] (index within the camera's N) and pixel_ids i64 [M] (pixel + camera * H * W). */
int gsb200_indices_in_range_count(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N,
                                  int64_t n_isects, const float *means2d, const float *conics,
                                  const float *opacities, uint32_t image_width, uint32_t image_height,
                                  uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                  const int32_t *tile_offsets, const int32_t *flatten_ids,
                                  const float *transmittances, int32_t *chunk_cnts, gsb200_stream_t stream);
int gsb200_indices_in_range_fill(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N,
                                 int64_t n_isects, const float *means2d, const float *conics,
                                 const float *opacities, uint32_t image_width, uint32_t image_height,
                                 uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                 const int32_t *tile_offsets, const int32_t *flatten_ids,
                                 const float *transmittances, const int64_t *chunk_starts,
                                 int64_t *gaussian_ids, int64_t *pixel_ids, gsb200_stream_t stream);

/* ---------------------------------------------------------------- quant-sim
 * replaces the torch elementwise chain of fake_quantize_ste (compression_simulation/ops.py:39-75).
 * mode 0 = "noise": y = act(clamp(x,lo,hi) + noise*q_step), noise [n] supplied or, if NULL, drawn
 *          in-kernel from Philox(seed, offset); mask[i] = lo <= x <= hi (gradient gate).
 * mode 1 = "round": x is clamped IN PLACE (ops.py:63), y = act(round_half_even((x-lo)/(hi-lo)/qs)
 *          * qs * (hi-lo) + lo), qs = 1/(2^bits-1).
 * act: 0 identity, 1 exp, 2 sigmoid -- the activation the trainer applies next
 * (simple_trainer.py:758-776), fused here so the attribute is read once. */
int gsb200_fake_quantize_fwd(int64_t n, float *x, float lo, float hi, int bitwidth, int mode, int act,
                             const float *noise, uint64_t seed, uint64_t offset, float *y,
                             gsb200_stream_t stream);
/* v_x = v_y * act'(y) * (mode==0 ? [lo<=x<=hi] : 1) */
int gsb200_fake_quantize_bwd(int64_t n, const float *x, const float *y, float lo, float hi, int mode,
                             int act, const float *v_y, float *v_x, gsb200_stream_t stream);
/* All attributes of one simulate_compression() call in ONE launch per direction (same arithmetic per element as the
 * single-tensor calls above).  fwd reads x / noise (or Philox(seed, offset)), writes y (and, in round mode, the clamped x);
 * bwd reads x, y, v_y and writes v_x.  Unused pointers of a direction may be NULL. */
#define GSB200_FQ_MAX_TENSORS 8
typedef struct {
    int64_t n; /* elements */
    float *x;
    float *y;
    const float *noise;
    const float *v_y;
    float *v_x;
    uint64_t seed, offset;
    float lo, hi;
    int bitwidth, mode, act;
} gsb200_fq_tensor;
int gsb200_fake_quantize_multi_fwd(uint32_t n_tensors, const gsb200_fq_tensor *tensors, gsb200_stream_t stream);
int gsb200_fake_quantize_multi_bwd(uint32_t n_tensors, const gsb200_fq_tensor *tensors, gsb200_stream_t stream);

/* On-disk quantise-to-grid of PngCompression (gsplat/compression/png_compression.py:166-360: _compress_png / _png_kbit /
 * _png_16bit and their inverses).  x [n_elems] = an attribute viewed as [.., ch]; mins / maxs [ch] are its per-channel extrema.
 * bits <= 8: lo[i] = round((x - min) / (max - min) * (2^bits - 1)) << (8 - bits); bits in (8, 16]: low / high byte planes.
 * dequantize mirrors the reference's float64 decode and writes float32. */
int gsb200_grid_quantize(int64_t n_elems, uint32_t ch, const float *x, const float *mins, const float *maxs, int bits,
                         uint8_t *lo, uint8_t *hi, gsb200_stream_t stream);
int gsb200_grid_dequantize(int64_t n_elems, uint32_t ch, const uint8_t *lo, const uint8_t *hi, const float *mins,
                           const float *maxs, int bits, float *out, gsb200_stream_t stream);

/* factorized-prior entropy model, replaces Entropy_factorized_optimized_refactor.forward
 * (compression_simulation/entropy_model.py:201-254) incl. its channel-scrambled weight tiling.
 * x [N,ch]; params: concatenated per layer i: softplus-ed matrix [ch,f(i+1),f(i)], bias [ch,f(i+1)],
 * tanh-ed factor [ch,f(i+1)] (absent for the last layer), raw layout described in
 * gscodec_studio_b200/compression_simulation/entropy_model.py.  n_hidden in {2,3} layers of width 3.
 * bits [N,ch].  bwd: v_x [N,ch] and v_params (same layout as params, pre-activation chain rule is
 * applied by the host).  v_params must be zero-initialised; it is accumulated atomically. */
int gsb200_entropy_fwd(int64_t N, uint32_t ch, uint32_t n_hidden, const float *x, float q,
                       const float *params, float likelihood_bound, float *bits,
                       gsb200_stream_t stream);
int gsb200_entropy_bwd(int64_t N, uint32_t ch, uint32_t n_hidden, const float *x, float q,
                       const float *params, float likelihood_bound, const float *v_bits, float *v_x,
                       float *v_params, gsb200_stream_t stream);

/* The activated parameter table `params` [ch, PS] of the two calls above from the module's raw parameters, one launch each
 * way: matrices[i] [ch, f(i+1), f(i)] -> softplus, bias[i] [ch, f(i+1), 1] as is, factor[i] [ch, f(i+1), 1] -> tanh
 * (entropy_model.py:216-232), i = 0..n_hidden (no factor for the last layer); HOST arrays of device pointers.  _bwd applies
 * the chain rule to v_params [ch, PS] and OVERWRITES v_matrices[i] / v_bias[i] / v_factor[i] (NULL = not wanted). */
int gsb200_entropy_params_fwd(uint32_t ch, uint32_t n_hidden, const float *const *matrices, const float *const *bias,
                              const float *const *factor, float *params, gsb200_stream_t stream);
int gsb200_entropy_params_bwd(uint32_t ch, uint32_t n_hidden, const float *const *matrices, const float *const *factor,
                              const float *v_params, float *const *v_matrices, float *const *v_bias,
                              float *const *v_factor, gsb200_stream_t stream);

/* ---------------------------------------------------------------- hash-grid encoder -- EXPERIMENTAL
 * Not yet validated on a GPU; the Python layer only reaches it with GSB200_EXPERIMENTAL_HASHGRID=1.
 * replaces _gridencoder.grid_encode_forward / _backward (third_party/gridencoder/src/gridencoder.cu:100-884,
 * bindings.cpp) as called by compression_simulation/gaussian_distribution_model.py:60-182 (no voxel mask, all levels),
 * with the STE_binary embedding transform (:19-38) fused when ste_binary != 0.
 * inputs [N,num_dim] in [0,1] (num_dim 2 or 3); params [offsets[n_levels], 2]; offsets i32 [n_levels+1] (rows) and
 * resolutions i32 [n_levels] are HOST arrays, n_levels <= GSB200_GRID_MAX_LEVELS; outputs [N, n_levels*2] level-major.
 * bwd overwrites v_params (same shape as params) and, when not NULL, v_inputs [N,num_dim]. */
#define GSB200_GRID_MAX_LEVELS 16
int gsb200_grid_encode_fwd(uint32_t N, uint32_t num_dim, uint32_t n_features, uint32_t n_levels,
                           const int32_t *offsets, const int32_t *resolutions, int ste_binary,
                           const float *inputs, const float *params, float *outputs, gsb200_stream_t stream);
int gsb200_grid_encode_bwd(uint32_t N, uint32_t num_dim, uint32_t n_features, uint32_t n_levels,
                           const int32_t *offsets, const int32_t *resolutions, int ste_binary,
                           const float *inputs, const float *params, const float *v_outputs, float *v_params,
                           float *v_inputs, gsb200_stream_t stream);

/* ---------------------------------------------------------------- optimizer step
 * gsb200_selective_adam_update replaces _C.selective_adam_update (ext.cpp, adam.cu:13-83; caller
 * gsplat/optimizers/selective_adam.py:42-94): param / param_grad / exp_avg / exp_avg_sq are [N, M] row-major,
 * visibility u8 [N] (NULL = every row); rows with a clear flag keep parameter and moments untouched:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  param += -lr m / (sqrt(v) + eps)      (no bias correction). */
int gsb200_selective_adam_update(float *param, const float *param_grad, float *exp_avg, float *exp_avg_sq,
                                 const uint8_t *visibility, float lr, float b1, float b2, float eps,
                                 uint32_t N, uint32_t M, gsb200_stream_t stream);
/* torch.optim.Adam's update (the optimizer examples/simple_trainer.py:396-408 creates per attribute) for up
 * to GSB200_ADAM_MAX_TENSORS tensors in one launch.  `tensors` is a HOST array; per tensor, with t = step count:
 *   step_size = lr / (1 - beta1^t),  inv_sqrt_bias2 = 1 / sqrt(1 - beta2^t),  one_minus_beta = 1 - beta (all rounded
 *   to float from double arithmetic on the host, as torch does with its Python scalars)
 *   m = beta1 m + (1-beta1) g;  v = beta2 v + (1-beta2) g^2;  param -= step_size m / (sqrt(v) inv_sqrt_bias2 + eps) */
#define GSB200_ADAM_MAX_TENSORS 16
typedef struct gsb200_adam_tensor {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t n; /* elements */
    float step_size, beta1, beta2, eps, inv_sqrt_bias2, one_minus_beta1, one_minus_beta2;
} gsb200_adam_tensor;
int gsb200_adam_multi(uint32_t n_tensors, const gsb200_adam_tensor *tensors, gsb200_stream_t stream);

/* replaces _C.compute_relocation (ext.cpp, compute_relocation.cu:13-74; caller gsplat/relocation.py:10-49, the MCMC
 * strategy): opacities [N], scales [N,3], ratios i32 [N] in [1, n_max], binoms [n_max, n_max] ->
 * new_opacities [N] = 1 - (1 - o)^(1/ratio), new_scales [N,3] = scales * o / sum_{i<=ratio} sum_{k<i} binom(i-1,k)
 * (-1)^k new_o^(k+1) / sqrt(k+1). */
int gsb200_compute_relocation(uint32_t N, const float *opacities, const float *scales, const int32_t *ratios,
                              const float *binoms, int n_max, float *new_opacities, float *new_scales,
                              gsb200_stream_t stream);

/* ---------------------------------------------------------------- photometric loss
 * replaces the loss chain of examples/simple_trainer.py:948-953: F.l1_loss + (1 - fused_ssim(..., padding="valid")),
 * where fused_ssim is the un-vendored dependency pinned at examples/requirements.txt:26 (rahul-goel/fused-ssim@1272e21a;
 * 11-tap Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, gradient to the first image only).
 * img1 (rendered) / img2 (target) are [B,H,W,CH] -- the renderer's layout, no NCHW permute copy.
 *   out[0] = coef_l1 * mean|img1 - img2| + coef_ssim * mean(ssim map) + bias;  stats (2 floats, may be NULL) = the L1
 *   mean and the SSIM mean
 *   (trainer loss: coef_l1 = 1 - lambda, coef_ssim = -lambda, bias = lambda; plain fused_ssim: 0, 1, 0)
 * padding_valid != 0 averages the SSIM map without its 5-pixel border (padding="valid"), else over all pixels.
 * map_mu / map_s1 / map_s12 [B,H,W,CH]: derivative maps saved for the backward (all NULL = inference only).
 * workspace: gsb200_photometric_workspace_bytes() bytes of device memory, 8-byte aligned (zeroed by the call).
 * bwd: v_img1 [B,H,W,CH] is overwritten with v_loss[0] * d out[0] / d img1 (v_loss is a DEVICE scalar). */
uint64_t gsb200_photometric_workspace_bytes(void);
int gsb200_photometric_loss_fwd(uint32_t B, uint32_t H, uint32_t W, uint32_t CH, const float *img1,
                                const float *img2, float coef_l1, float coef_ssim, float bias,
                                int padding_valid, float *map_mu, float *map_s1, float *map_s12,
                                void *workspace, float *out, float *stats, gsb200_stream_t stream);
int gsb200_photometric_loss_bwd(uint32_t B, uint32_t H, uint32_t W, uint32_t CH, const float *img1,
                                const float *img2, float coef_l1, float coef_ssim, int padding_valid,
                                const float *map_mu, const float *map_s1, const float *map_s12,
                                const float *v_loss, float *v_img1, gsb200_stream_t stream);

/* ---------------------------------------------------------------- un-fused geometry ops
 * The projection stage as separate differentiable pieces, exported by gsplat/__init__.py:7-20 and tested by
 * the reference's tests/test_basic.py:53-173.  Covariances are FULL row-major matrices (need not be symmetric).
 *
 * replaces _C.quat_scale_to_covar_preci_fwd / _bwd (bindings.h, quat_scale_to_covar_preci_{fwd,bwd}.cu):
 * quats [N,4] wxyz (normalised inside), scales [N,3]; covars / precis [N,3,3], or [N,6] (xx,xy,xz,yy,yz,zz) when
 * triu != 0; either output may be NULL.  bwd: v_covars / v_precis in the same layout (NULL = no gradient from it);
 * a packed off-diagonal gradient counts half for each mirrored entry.  v_quats [N,4], v_scales [N,3] overwritten. */
int gsb200_quat_scale_to_covar_preci_fwd(uint32_t N, const float *quats, const float *scales, int triu,
                                         float *covars, float *precis, gsb200_stream_t stream);
int gsb200_quat_scale_to_covar_preci_bwd(uint32_t N, const float *quats, const float *scales, int triu,
                                         const float *v_covars, const float *v_precis, float *v_quats,
                                         float *v_scales, gsb200_stream_t stream);
/* replaces _C.proj_fwd / _C.proj_bwd (proj_{fwd,bwd}.cu, include/proj.cuh): camera-space means [C,N,3] and covars
 * [C,N,3,3], Ks [C,3,3] -> means2d [C,N,2], covars2d [C,N,2,2] = J covar J^T for the pinhole (with the reference's
 * field-of-view clamp of the Jacobian), orthographic or fisheye model. */
int gsb200_proj_fwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
                    uint32_t width, uint32_t height, int camera_model, float *means2d, float *covars2d,
                    gsb200_stream_t stream);
int gsb200_proj_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
                    uint32_t width, uint32_t height, int camera_model, const float *v_means2d,
                    const float *v_covars2d, float *v_means, float *v_covars, gsb200_stream_t stream);
/* replaces _C.world_to_cam_fwd / _bwd (world_to_cam_{fwd,bwd}.cu, include/transform.cuh:9-71): means [N,3], covars
 * [N,3,3], viewmats [C,4,4] -> means_c [C,N,3] = R p + t, covars_c [C,N,3,3] = R covar R^T.  bwd: v_means [N,3] and
 * v_covars [N,3,3] (summed over cameras) and v_viewmats [C,4,4] (summed over Gaussians, last row zero) are
 * overwritten; any of the three may be NULL. */
int gsb200_world_to_cam_fwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                            const float *viewmats, float *means_c, float *covars_c, gsb200_stream_t stream);
int gsb200_world_to_cam_bwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                            const float *viewmats, const float *v_means_c, const float *v_covars_c,
                            float *v_means, float *v_covars, float *v_viewmats, gsb200_stream_t stream);

/* ---------------------------------------------------------------- dynamic splats: temporal pre-transform
 * The per-timestamp attribute chain of examples/simple_trainer_dyngs.py:476-520 (BASELINE config 4) in one pass:
 *   d = timestamp - trbf_center;  trbf = exp(-(d / (sqrt(2) exp(trbf_scale_raw)))^2)
 *   opacities = sigmoid(opacities_raw) * trbf;  scales = exp(scales_raw)
 *   means_t = means + motion[0:3] d + motion[3:6] d^2 + motion[6:9] d^3;  quats_t = normalize(quats + d omega)
 * All arrays have N rows: means [N,3], quats / omega [N,4], scales_raw [N,3], opacities_raw / trbf_center /
 * trbf_scale_raw [N], motion [N,9].  In the polynomial and the rotation d is a constant (the trainer detaches it),
 * so trbf_center / trbf_scale_raw receive gradients through the opacity only.  bwd overwrites every v_* output. */
int gsb200_stg_temporal_fwd(uint32_t N, float timestamp, const float *means, const float *quats,
                            const float *scales_raw, const float *opacities_raw, const float *trbf_center,
                            const float *trbf_scale_raw, const float *motion, const float *omega,
                            float *means_t, float *quats_t, float *scales, float *opacities, float *trbf,
                            gsb200_stream_t stream);
int gsb200_stg_temporal_bwd(uint32_t N, float timestamp, const float *quats, const float *scales_raw,
                            const float *opacities_raw, const float *trbf_center, const float *trbf_scale_raw,
                            const float *omega, const float *v_means_t, const float *v_quats_t,
                            const float *v_scales, const float *v_opacities, float *v_means, float *v_quats,
                            float *v_scales_raw, float *v_opacities_raw, float *v_trbf_center,
                            float *v_trbf_scale_raw, float *v_motion, float *v_omega, gsb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H */
