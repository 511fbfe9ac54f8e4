/*
 * TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path
 * (gscodec_studio_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call into this file.
 *
 * Plain-C CPU restatement of the reference's per-tile alpha-blend (stage 3 and stage 4a):
 *   forward : /root/reference/gsplat/cuda/csrc/rasterize_to_pixels_fwd.cu:59-184
 *   backward: /root/reference/gsplat/cuda/csrc/rasterize_to_pixels_bwd.cu:105-275
 * The per-pixel arithmetic is float32 in the reference's expression order.  Gradient sums,
 * which the reference accumulates with float atomics in an unspecified order, are accumulated
 * here per tile in double and added once per (tile, Gaussian) -- a legitimate member of the
 * reference's set of possible summation orders, and a more accurate one.
 *
 * Parity pinning: the reference ships no golden vectors for this path and its torch blend
 * (_torch_impl.py:522-617) needs CUDA + nerfacc, so this file is pinned against the UNMODIFIED
 * reference CUDA kernels (oracle/_ref/gsplat_ref_cuda.so) run on a B200; the resulting vectors
 * are committed under tests/golden/ (see oracle/make_golden_gpu.py).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared -o liboracle.so oracle_cpu.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ALPHA_MAX 0.999f
#define ALPHA_MIN (1.f / 255.f)
#define T_STOP 1e-4 /* double literal on purpose: reference compares float next_T <= 1e-4 (double) */

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline void atomic_addf(float *p, double v) {
    const float f = (float)v;
    if (f != 0.f) {
#pragma omp atomic
        *p += f;
    }
}

/* range of sorted intersections owned by (cam, tile): rasterize_to_pixels_fwd.cu:79-85 */
static inline void tile_range(const int32_t *tile_offsets, int C, int n_tiles, int cam, int tile,
                              int n_isects, int *start, int *end) {
    const int32_t *off = tile_offsets + (int64_t)cam * n_tiles;
    *start = off[tile];
    *end = (cam == C - 1 && tile == n_tiles - 1) ? n_isects : off[tile + 1];
}

/*
 * Stage 3.  Layouts follow the reference op (bindings.h:168-184):
 * means2d [CN,2], conics [CN,3], colors [CN,D], opacities [CN] (CN = C*N or nnz),
 * backgrounds [C,D] or NULL, masks [C,th,tw] (uint8) or NULL, tile_offsets [C,th,tw],
 * flatten_ids [n_isects]; outputs render_colors [C,H,W,D], render_alphas [C,H,W], last_ids [C,H,W].
 * tile_begin/tile_end select a sub-range of flattened (cam,tile) ids (used for bounded CPU timing);
 * pass 0 and C*th*tw for the whole job.
 */
void oracle_blend_fwd(int C, int n_isects, int D, const float *means2d, const float *conics,
                      const float *colors, const float *opacities, const float *backgrounds,
                      const uint8_t *masks, int W, int H, int tile_size, int tile_w, int tile_h,
                      const int32_t *tile_offsets, const int32_t *flatten_ids, float *render_colors,
                      float *render_alphas, int32_t *last_ids, int tile_begin, int tile_end) {
    const int n_tiles = tile_w * tile_h;
#pragma omp parallel for schedule(dynamic, 4)
    for (int ct = tile_begin; ct < tile_end; ++ct) {
        const int cam = ct / n_tiles, tile = ct % n_tiles;
        const int ty = tile / tile_w, tx = tile % tile_w;
        int start, end;
        tile_range(tile_offsets, C, n_tiles, cam, tile, n_isects, &start, &end);
        const float *bg = backgrounds ? backgrounds + (int64_t)cam * D : NULL;
        float *pix_out = (float *)malloc(sizeof(float) * (size_t)D);
        for (int iy = 0; iy < tile_size; ++iy) {
            for (int ix = 0; ix < tile_size; ++ix) {
                const int i = ty * tile_size + iy, j = tx * tile_size + ix;
                if (i >= H || j >= W) continue;
                const int64_t pix = (int64_t)cam * H * W + (int64_t)i * W + j;
                if (masks && !masks[(int64_t)cam * n_tiles + tile]) {
                    /* masked tile: background (or 0) only; alpha / last_ids untouched (fwd.cu:70-76) */
                    for (int k = 0; k < D; ++k) render_colors[pix * D + k] = bg ? bg[k] : 0.f;
                    continue;
                }
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.f;
                int32_t cur = 0;
                for (int k = 0; k < D; ++k) pix_out[k] = 0.f;
                for (int idx = start; idx < end; ++idx) {
                    const int32_t g = flatten_ids[idx];
                    const float dx = means2d[2 * (int64_t)g] - px, dy = means2d[2 * (int64_t)g + 1] - py;
                    const float ca = conics[3 * (int64_t)g], cb = conics[3 * (int64_t)g + 1],
                                cc = conics[3 * (int64_t)g + 2];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    float alpha = opacities[g] * expf(-sigma);
                    if (alpha > ALPHA_MAX) alpha = ALPHA_MAX;
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    const float next_T = T * (1.f - alpha);
                    if ((double)next_T <= T_STOP) break; /* exclusive: this Gaussian is not applied */
                    const float vis = alpha * T;
                    const float *c = colors + (int64_t)g * D;
                    for (int k = 0; k < D; ++k) pix_out[k] += c[k] * vis;
                    cur = idx;
                    T = next_T;
                }
                render_alphas[pix] = 1.f - T;
                for (int k = 0; k < D; ++k)
                    render_colors[pix * D + k] = bg ? (pix_out[k] + T * bg[k]) : pix_out[k];
                last_ids[pix] = cur;
            }
        }
        free(pix_out);
    }
}

/*
 * Stage 4a.  Inputs as the reference op (bindings.h:186-215).  Outputs must be zero-initialised by
 * the caller: v_means2d_abs [CN,2] or NULL, v_means2d [CN,2], v_conics [CN,3], v_colors [CN,D],
 * v_opacities [CN].
 */
void oracle_blend_bwd(int C, int n_isects, int D, const float *means2d, const float *conics,
                      const float *colors, const float *opacities, const float *backgrounds,
                      const uint8_t *masks, int W, int H, int tile_size, int tile_w, int tile_h,
                      const int32_t *tile_offsets, const int32_t *flatten_ids,
                      const float *render_alphas, const int32_t *last_ids,
                      const float *v_render_colors, const float *v_render_alphas,
                      float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors,
                      float *v_opacities, int tile_begin, int tile_end) {
    const int n_tiles = tile_w * tile_h;
    const int S = 8 + D; /* per-Gaussian accumulator: xy(2) abs(2) conic(3) opac(1) rgb(D) */
#pragma omp parallel for schedule(dynamic, 4)
    for (int ct = tile_begin; ct < tile_end; ++ct) {
        const int cam = ct / n_tiles, tile = ct % n_tiles;
        const int ty = tile / tile_w, tx = tile % tile_w;
        if (masks && !masks[(int64_t)cam * n_tiles + tile]) continue;
        int start, end;
        tile_range(tile_offsets, C, n_tiles, cam, tile, n_isects, &start, &end);
        if (end <= start) continue;
        const float *bg = backgrounds ? backgrounds + (int64_t)cam * D : NULL;
        double *acc = (double *)calloc((size_t)(end - start) * (size_t)S, sizeof(double));
        float *buffer = (float *)malloc(sizeof(float) * (size_t)D);
        for (int iy = 0; iy < tile_size; ++iy) {
            for (int ix = 0; ix < tile_size; ++ix) {
                const int i = ty * tile_size + iy, j = tx * tile_size + ix;
                if (i >= H || j >= W) continue;
                const int64_t pix = (int64_t)cam * H * W + (int64_t)i * W + j;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                const float T_final = 1.f - render_alphas[pix];
                float T = T_final;
                const int bin_final = last_ids[pix];
                const float *v_c = v_render_colors + pix * D;
                const float v_a = v_render_alphas[pix];
                for (int k = 0; k < D; ++k) buffer[k] = 0.f;
                /* back to front, starting at the last Gaussian the forward applied (bwd.cu:155-168) */
                int hi = bin_final < end - 1 ? bin_final : end - 1;
                for (int idx = hi; idx >= start; --idx) {
                    const int32_t g = flatten_ids[idx];
                    const float dx = means2d[2 * (int64_t)g] - px, dy = means2d[2 * (int64_t)g + 1] - py;
                    const float ca = conics[3 * (int64_t)g], cb = conics[3 * (int64_t)g + 1],
                                cc = conics[3 * (int64_t)g + 2];
                    const float opac = opacities[g];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    const float vis = expf(-sigma);
                    float alpha = opac * vis;
                    if (alpha > ALPHA_MAX) alpha = ALPHA_MAX;
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    const float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    const float *c = colors + (int64_t)g * D;
                    double *a = acc + (size_t)(idx - start) * (size_t)S;
                    float v_alpha = 0.f;
                    for (int k = 0; k < D; ++k) {
                        a[8 + k] += (double)(fac * v_c[k]);
                        v_alpha += (c[k] * T - buffer[k] * ra) * v_c[k];
                    }
                    v_alpha += T_final * ra * v_a;
                    if (bg) {
                        float accum = 0.f;
                        for (int k = 0; k < D; ++k) accum += bg[k] * v_c[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    if (opac * vis <= ALPHA_MAX) {
                        const float v_sigma = -opac * vis * v_alpha;
                        const float gx = v_sigma * (ca * dx + cb * dy);
                        const float gy = v_sigma * (cb * dx + cc * dy);
                        a[0] += (double)gx;
                        a[1] += (double)gy;
                        a[2] += (double)fabsf(gx);
                        a[3] += (double)fabsf(gy);
                        a[4] += (double)(0.5f * v_sigma * dx * dx);
                        a[5] += (double)(v_sigma * dx * dy);
                        a[6] += (double)(0.5f * v_sigma * dy * dy);
                        a[7] += (double)(vis * v_alpha);
                    }
                    for (int k = 0; k < D; ++k) buffer[k] += c[k] * fac;
                }
            }
        }
        for (int idx = start; idx < end; ++idx) {
            const int64_t g = flatten_ids[idx];
            const double *a = acc + (size_t)(idx - start) * (size_t)S;
            atomic_addf(&v_means2d[2 * g], a[0]);
            atomic_addf(&v_means2d[2 * g + 1], a[1]);
            if (v_means2d_abs) {
                atomic_addf(&v_means2d_abs[2 * g], a[2]);
                atomic_addf(&v_means2d_abs[2 * g + 1], a[3]);
            }
            atomic_addf(&v_conics[3 * g], a[4]);
            atomic_addf(&v_conics[3 * g + 1], a[5]);
            atomic_addf(&v_conics[3 * g + 2], a[6]);
            atomic_addf(&v_opacities[g], a[7]);
            for (int k = 0; k < D; ++k) atomic_addf(&v_colors[g * D + k], a[8 + k]);
        }
        free(acc);
        free(buffer);
    }
}
