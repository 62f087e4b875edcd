/*
 * pnerf_b200.h — C ABI of libpnerf_b200.so: the B200 (sm_100a) implementation of PaletteNeRF's
 * volumetric-rendering hot path.
 *
 * Conventions (all entry points)
 *   - plain device pointers + sizes, no torch types; caller allocates every buffer (SURVEY §8b "Ownership");
 *   - asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - never throws, never allocates device memory, keeps no pointer after return;
 *   - returns PNERF_OK (0) or a negative pnerf_status; pnerf_status_string() explains it;
 *   - "ref:" names the reference interface the function replaces (path:line under the reference repo).
 *
 * The Python packages raymarching/, gridencoder/, shencoder/, freqencoder/ and palette.backend bind exactly
 * these symbols through ctypes (palettenerf_b200/_lib.py); INTEGRATION.md shows the stub.
 */
#ifndef PNERF_B200_H_
#define PNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PNERF_API __attribute__((visibility("default")))
#else
#define PNERF_API
#endif

typedef enum pnerf_status {
    PNERF_OK = 0,
    PNERF_ERR_INVALID_ARG = -1,  /* null pointer, negative size, bad enum */
    PNERF_ERR_UNSUPPORTED = -2,  /* configuration outside what the kernels implement (e.g. n_channel > 128) */
    PNERF_ERR_CUDA = -3          /* a CUDA runtime call / launch failed; see pnerf_last_cuda_error() */
} pnerf_status;

typedef enum pnerf_dtype { PNERF_F16 = 0, PNERF_F32 = 1, PNERF_F64 = 2 } pnerf_dtype;

/* output layout of the grid encoder: the reference kernel writes [L,B,C] and the wrapper permutes
 * (ref: gridencoder/grid.py:41-52); the native layout writes the final [B, L*C] rows directly. */
typedef enum pnerf_grid_layout { PNERF_LAYOUT_LBC = 0, PNERF_LAYOUT_BLC = 1 } pnerf_grid_layout;

PNERF_API const char* pnerf_status_string(int status);
PNERF_API const char* pnerf_last_cuda_error(void);
PNERF_API int pnerf_abi_version(void);
/* zero-fill of a caller-allocated buffer on `stream` (cudaMemsetAsync; a memset node under graph capture): what the reference's
 * wrappers do with torch.zeros / zeros_like for gradient buffers (gridencoder/grid.py:72, raymarching/raymarching.py:283-284) */
PNERF_API int pnerf_zero_fill(void* dst, uint64_t bytes, void* stream);
/* compiled gencode string, e.g. "sm_100a" */
PNERF_API const char* pnerf_build_arch(void);

/* ------------------------------------------------------------------------------------------------
 * raymarching  (ref: raymarching/src/raymarching.h:7-23, bindings.cpp:5-23)
 * all float buffers are fp32, index buffers int32, bitfield uint8
 * ---------------------------------------------------------------------------------------------- */

/* ref: raymarching.cu:95-159 near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars) */
PNERF_API int pnerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                                       uint32_t N, float min_near, float* nears, float* fars, void* stream);

/* ref: raymarching.cu:166-211 sph_from_ray(rays_o, rays_d, radius, N, coords[N,2]) */
PNERF_API int pnerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                                 float* coords, void* stream);

/* ref: raymarching.cu:217-235 morton3D(coords[N,3], N, indices[N]) */
PNERF_API int pnerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);

/* ref: raymarching.cu:240-263 morton3D_invert(indices[N], N, coords[N,3]) */
PNERF_API int pnerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);

/* ref: raymarching.cu:271-303 packbits(grid, N = C*H^3/8 bytes, density_thresh, bitfield[N]) */
PNERF_API int pnerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                             void* stream);

/* ref: raymarching.cu:315-493 march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
 *      nears, fars, xyzs, dirs, deltas, rays, counter, noises)
 * xyzs/dirs [M,3], deltas [M,2] must be zero-initialised by the caller (only occupied slots are written);
 * rays [N,3] = (ray id, sample offset, sample count); counter[2] += (total samples, N).
 * Slot assignment is a deterministic exclusive scan in ray order (one of the orders the reference's
 * atomicAdd race can produce); rays that would overflow M are skipped exactly like raymarching.cu:419.
 * `order` : 0 = deterministic scan (default), 1 = reserved. */
PNERF_API int pnerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                     float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                                     uint32_t H, uint32_t M, const float* nears, const float* fars,
                                     float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                                     const float* noises, void* stream);

/* pnerf_march_rays_train with workspaces: ONE walk of the occupancy grid instead of the reference's two
 * (raymarching.cu:358-409 counts, :421-482 walks again to write). Same outputs, bit for bit.
 *   t_list   [N, max_steps] fp32 scratch (contents ignored): the counting walk records the ray parameter of every
 *            sample; the write pass turns the list into xyzs / dirs / deltas without touching the grid again.
 *   occ_aabb [6] or NULL: bounds of the occupied cells from pnerf_occupied_bounds for THIS grid; the counting walk
 *            stops where the ray leaves them instead of marching the empty tail up to `far` (no lattice point behind
 *            that exit can fall into an occupied cell, so the samples are unchanged).
 *   valid_rows [1] int32 or NULL: receives the number of leading rows of xyzs / dirs / deltas that were written. Rows are
 *            handed out in ray order, so this is min(counter[0], offset of the first ray with offset + num_steps > M):
 *            a caller that does not zero-fill the buffers (static-capacity training step) must stop reading there. */
PNERF_API int pnerf_march_rays_train_ws(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                        float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                        uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                                        float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                        float* t_list, const float* occ_aabb, int32_t* valid_rows, void* stream);

/* occ_aabb[0..6) (device) = (lo xyz, hi xyz): world-space bounds of every occupied cell of the C cascades of `bitfield`
 * ([C*H^3/8] bytes, Morton order, as packbits writes it), padded by one cell; a side reaching the scene bound is
 * +-FLT_MAX; an empty grid gives lo > hi. The buffer must hold pnerf_occupied_bounds_floats() floats (the tail is
 * scratch for the per-CTA partial bounds). Two small launches; recompute when the bitfield changes. */
PNERF_API uint32_t pnerf_occupied_bounds_floats(void);
PNERF_API int pnerf_occupied_bounds(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, float* occ_aabb,
                                    void* stream);

/* NeRFRenderer.mark_untrained_grid (ref: nerf/renderer.py:395-465) as one kernel: poses [B,4,4] row-major camera-to-world,
 * intrinsics fx fy cx cy; density_grid [C, H^3] (Morton order) receives -1 in every cell that no camera sees or that a
 * seeing camera has closer than min_near (filter_close_point: also cells within min_near of any camera centre).
 * n_marked (optional device counter, += number of cells marked). B = 0 marks everything. */
PNERF_API int pnerf_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t C,
                                        uint32_t H, float bound, float min_near, int filter_close_point,
                                        float* density_grid, uint32_t* n_marked, void* stream);

/* ref: raymarching.cu:504-580,647-655 */
PNERF_API int pnerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                                 const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                                 float* weights_sum, float* depth, float* image, void* stream);

/* ref: raymarching.cu:681-761,821-829; grad_sigmas/grad_rgbs must be zero-initialised */
PNERF_API int pnerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                                  const float* sigmas, const float* rgbs, const float* deltas,
                                                  const int32_t* rays, const float* weights_sum,
                                                  const float* image, uint32_t M, uint32_t N, float T_thresh,
                                                  float* grad_sigmas, float* grad_rgbs, void* stream);

/* ref: raymarching.cu:583-645,657-668 (n_channel <= 128, else PNERF_ERR_UNSUPPORTED) */
PNERF_API int pnerf_composite_rays_flex_train_forward(const float* sigmas, const float* input,
                                                      const float* deltas, const int32_t* rays, uint32_t M,
                                                      uint32_t N, uint32_t n_channel, float T_thresh,
                                                      float* output, void* stream);

/* ref: raymarching.cu:764-819,831-844; grad_input must be zero-initialised */
PNERF_API int pnerf_composite_rays_flex_train_backward(const float* grad_output, const float* sigmas,
                                                       const float* input, const float* deltas,
                                                       const int32_t* rays, const float* output, uint32_t M,
                                                       uint32_t N, uint32_t n_channel, float T_thresh,
                                                       float* grad_input, void* stream);

/* ref: raymarching.cu:848-894 spread_ray_to_sample(input[N,c], rays, M, N, n_channel, output[M,c]) */
PNERF_API int pnerf_spread_ray_to_sample(const float* input, const int32_t* rays, uint32_t M, uint32_t N,
                                         uint32_t n_channel, float* output, void* stream);

/* ref: raymarching.cu:907-1021 march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma,
 *      max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, noises); outputs zero-initialised by caller */
PNERF_API int pnerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                               const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                               const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                               const float* noises, void* stream);

/* ref: raymarching.cu:1025-1111,1187-1193 (in-place on rays_alive, rays_t, weights_sum, depth, image) */
PNERF_API int pnerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                                   float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                                   float* weights_sum, float* depth, float* image, void* stream);

/* ref: raymarching.cu:1114-1185,1195-1205 (in-place on output; reads weights_sum) */
PNERF_API int pnerf_composite_rays_flex(uint32_t n_alive, uint32_t n_step, uint32_t n_channel, float T_thresh,
                                        const int32_t* rays_alive, const float* rays_t, const float* sigmas,
                                        const float* input, const float* deltas, const float* weights_sum,
                                        float* output, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gridencoder  (ref: gridencoder/src/gridencoder.h:12-13)
 * inputs fp32 [B,D] in [0,1]; embeddings/outputs/grad of `dtype`; offsets int32 [L+1]
 * S = log2(per_level_scale) as fp32, H = base resolution, gridtype 0 = hash, 1 = tiled
 * dy_dx (optional, may be NULL): [B, L, D, C] of `dtype`
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                                        void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                        uint32_t H, void* dy_dx, uint32_t gridtype, int align_corners,
                                        int dtype, int out_layout, void* stream);

/* grad_embeddings must be zero-initialised (or hold a running sum to accumulate into);
 * grad_inputs (optional, with dy_dx) is overwritten; grad layout follows `grad_layout`. */
PNERF_API int pnerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                         const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                         uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                                         void* grad_inputs, uint32_t gridtype, int align_corners, int dtype,
                                         int grad_layout, void* stream);

/* ------------------------------------------------------------------------------------------------
 * shencoder  (ref: shencoder/src/shencoder.h:9-10)  fp32, D must be 3, degree C in [1,8]
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C,
                                      float* dy_dx, void* stream);
/* grad_inputs is accumulated into (+=), as shencoder.cu:377-380 does */
PNERF_API int pnerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D,
                                       uint32_t C, const float* dy_dx, float* grad_inputs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * freqencoder  (ref: freqencoder/src/freqencoder.h:7-10)  fp32
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                        float* outputs, void* stream);
PNERF_API int pnerf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D,
                                         uint32_t deg, uint32_t C, float* grad_inputs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * palette ops  (ref: palette/src/palette_func.h:7-8, palette/src/bindings.cpp:40-104)
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_rgb_to_hsv(uint32_t n, const float* input, float* output, void* stream);
PNERF_API int pnerf_hsv_to_rgb(uint32_t n, const float* input, float* output, void* stream);
/* host function (CPU pointers): colors_rgb [n,3] f32, weights [n] f32 ->
 * bin_weights [2^(3b)] f64, bin_centers_rgb [2^(3b),3] f32; ref: bindings.cpp:52-91 */
PNERF_API int pnerf_compute_rgb_histogram(const float* colors_rgb, const float* weights, uint64_t n,
                                          int bits_per_channel, double* bin_weights, float* bin_centers_rgb);

/* ------------------------------------------------------------------------------------------------
 * fused palette field + persistent fused renderer (new entry points; no reference counterpart at the kernel level)
 * They replace, as ONE launch each, what the reference does with many:
 *   pnerf_palette_field_forward  <->  PaletteNetwork.forward in eval mode      (ref: palette/network.py:156-280)
 *   pnerf_palette_render_fused   <->  the inference loop of PaletteRenderer.run_cuda (ref: palette/renderer.py:430-523)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnerf_palette_field {
    const void* table_sigma;    /* fp16 [n_entries,2] : encoder.embeddings          */
    const void* table_palette;  /* fp16               : encoder_palette.embeddings  */
    const void* table_clip;     /* fp16               : encoder_clip.embeddings (NULL unless pred_clip) */
    const int32_t* offsets;     /* [L+1]                                             */
    const void* wpack;          /* fp16 mma.m16n8k16 B fragments, palettenerf_b200/fused.py::pack_weights */
    const float* head_bias;     /* [16] offsets_radiance_net.bias (13) + zeros      */
    const float* palette;       /* [4*3] basis_color clamped to [0,1]               */
    uint32_t L, H, pred_clip, clip_dim;
    float S, bound, density_scale, offsets_weight, view_dep_weight;
    /* optional (may be NULL): the density and palette tables interleaved entry by entry, fp16 [n_entries][2 tables][2],
     * so that one 8-byte load fetches the features of both grids at a lattice corner (the two grids share their
     * geometry). The inference kernels prefer it; table_sigma / table_palette must still be valid. */
    const void* table_sigma_palette;
    /* optional (may be NULL): the MLP weights as tcgen05 B operands — per layer fp16 [k-chunk][n][8] (K-major, no swizzle),
     * pnerf_palette_tc_weight_bytes() bytes, palettenerf_b200/fused.py::tc_pack_index. Needed by the *_tc entry points. */
    const void* wpack_tc;
    /* 0: palette model (all of the above); 1: stage-1 NeRF model (ref: nerf/network.py:95-124) — one hash grid (table_sigma),
     * sigma net + colour net only (the same weight image with the palette layers unused); the *_tc renderer then composites
     * the colour net's output and writes image / depth / weights_sum only */
    uint32_t model_kind;
} pnerf_palette_field;

/* xyzs, dirs [M,3] fp32 -> sigma [M], clip [M,clip_dim] (NULL unless pred_clip), omega [M,4], off_rad [M,13],
 * view_dep [M,3], diffuse [M,3]; fp16 tensor-core math with fp32 accumulation */
PNERF_API int pnerf_palette_field_forward(const float* xyzs, const float* dirs, uint32_t M,
                                          const pnerf_palette_field* field, float* sigma, float* clip, float* omega,
                                          float* off_rad, float* view_dep, float* diffuse, void* stream);

/* all outputs zero-initialised by the caller; the five aux maps may all be NULL (gui_mode); queue[68] zeroed
 * (4 counters + a 32-bucket histogram + 32 cursors used to order the rays longest-first):
 * on return queue[1] = number of samples shaded, queue[2] = number of rays with at least one sample,
 * queue[3] = number of 32-sample tiles evaluated (queue[1] / (32 queue[3]) = tile fill).
 * hit_list (int32) is a [2N] scratch buffer (ordered hit list + samples per ray), t_first, t_last (fp32) are [N].
 * occ_aabb (optional, may be NULL): pnerf_occupied_bounds of `bitfield`; the ray pre-pass then stops each walk at the
 * ray's exit from the occupied bounds (same samples; rays that miss the bounds are not walked at all). */
PNERF_API int pnerf_palette_render_fused(const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                                         const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C,
                                         uint32_t Hgrid, uint32_t max_steps, float dt_gamma, float T_thresh,
                                         const pnerf_palette_field* field, float* weights_sum, float* depth, float* image,
                                         float* direct_rgb, float* view_dep_rgb, float* basis_acc, float* basis_rgb,
                                         float* unscaled_basis_rgb, float* clip_feat, uint32_t* queue,
                                         int32_t* hit_list, float* t_first, float* t_last, const float* occ_aabb,
                                         void* stream);

/* The fp16 state of pnerf_palette_field rebuilt from the model's fp32 parameters (csrc/field_cache.cu; no reference
 * counterpart: its MLPs read the fp32 parameters through autocast on every call, palette/network.py:156-280).
 *   pnerf_field_cache_tables: table_sigma, table_palette (, table_clip or NULL) fp32 [n_entries, 2] -> pair = fp16
 *     [n_entries][2 grids][2] (field->table_sigma_palette), clip = fp16 [n_entries, 2] (field->table_clip).
 *   pnerf_field_cache_gather: src [n16 + n32] = device addresses of fp32 scalars (0 stands for the value 0); the first n16
 *     are written to out16 as fp16 (the weight images wpack / wpack_tc), the other n32 to out32 as fp32, clamped to [0, 1]
 *     from element clamp_from on (head_bias, then the palette).
 *   pnerf_field_cache_merge: the two product layers of the tcgen05 weight image (csrc/field_tc.cuh layer table), from the
 *     fp32 parameters sigma_net.1.weight [16,64], diff_net.0.weight [64,15], basis_net.1.weight [15,64],
 *     offsets_radiance_net.weight [13,15], omega_net.0.weight [4,15]: TD0 = diff_net.0 x sigma_net.1[1:16] and
 *     TB1 = [offsets_radiance_net ; omega_net.0] x basis_net.1 (the reference applies no activation between these layers,
 *     nerf/network.py:101-107, palette/network.py:262-268), written into wimage_tc (field->wpack_tc) after the gather. */
PNERF_API int pnerf_field_cache_tables(const float* table_sigma, const float* table_palette, const float* table_clip,
                                       uint32_t n_entries, void* pair, void* clip, void* stream);
PNERF_API int pnerf_field_cache_gather(const uint64_t* src, uint32_t n16, uint32_t n32, uint32_t clamp_from, void* out16,
                                       float* out32, void* stream);
PNERF_API int pnerf_field_cache_merge(const float* sigma1_w, const float* diff0_w, const float* basis1_w, const float* offrad_w,
                                      const float* omega_w, void* wimage_tc, void* stream);

/* Tensor-core (tcgen05 / TMEM) version of pnerf_palette_field_forward: a warpgroup evaluates 128 samples per tile, every
 * dense layer is one tcgen05.mma chain (csrc/field_tc.cuh). Same arguments and results (fp16 operands, fp32 accumulation). */
PNERF_API uint32_t pnerf_palette_tc_weight_bytes(uint32_t pred_clip);
PNERF_API int pnerf_palette_field_forward_tc(const float* xyzs, const float* dirs, uint32_t M,
                                             const pnerf_palette_field* field, float* sigma, float* clip, float* omega,
                                             float* off_rad, float* view_dep, float* diffuse, void* stream);

/* Round-2 renderer: a warp owns ONE ray at a time (csrc/render_rays.cu): warp-cooperative lattice walk of the ray, tiles of
 * 32 consecutive samples (lane = sample) through the fused field with lane-pair hash-grid gathers, warp-scan compositing,
 * one writer per ray. Same arguments as pnerf_palette_render_fused up to the scratch:
 *   queue     [8] u32, zero on entry; on return {ray cursor, samples shaded, rays with samples, 32-sample tiles, candidates}
 *   cand      [N] int32 scratch (rays that can have samples)
 *   t_scratch [pnerf_palette_render_rays_warps() * max_steps] fp32 scratch (per-warp sample lists)
 * Requires field->table_sigma_palette (interleaved tables). Replaces palette/renderer.py:430-523. */
PNERF_API uint32_t pnerf_palette_render_rays_warps(void);
PNERF_API int pnerf_palette_render_rays(const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                                        const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C,
                                        uint32_t Hgrid, uint32_t max_steps, float dt_gamma, float T_thresh,
                                        const pnerf_palette_field* field, float* weights_sum, float* depth, float* image,
                                        float* direct_rgb, float* view_dep_rgb, float* basis_acc, float* basis_rgb,
                                        float* unscaled_basis_rgb, float* clip_feat, uint32_t* queue, int32_t* cand,
                                        float* t_scratch, const float* occ_aabb, void* stream);
/* ------------------------------------------------------------------------------------------------
 * density-grid refresh as kernels (ref: NeRFRenderer.update_extra_state, nerf/renderer.py:467-561) — csrc/density_tc.cu
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_density_tc(const float* xyzs, uint32_t M, const pnerf_palette_field* field, float* sigma, void* stream);
PNERF_API uint32_t pnerf_density_occupied_chunks(uint32_t H);
PNERF_API int pnerf_density_occupied_list(const float* density_grid, uint32_t C, uint32_t H, int32_t* occ_list,
                                          uint32_t* occ_count, uint32_t* chunk_scratch, void* stream);
PNERF_API int pnerf_density_grid_sweep(float* tmp_grid, uint32_t C, uint32_t H, float bound, float density_scale,
                                       uint32_t partial, uint32_t n_random, const int32_t* occ_list, const uint32_t* occ_count,
                                       uint64_t seed, uint32_t rank, uint32_t world, const float* jitter,
                                       const pnerf_palette_field* field, void* stream);
PNERF_API uint32_t pnerf_density_finalize_partials(uint64_t n_cells);
PNERF_API int pnerf_density_grid_finalize(float* density_grid, float* tmp_grid, uint32_t C, uint32_t H, float decay,
                                          float density_thresh, float* partials, uint8_t* bitfield, float* stats, void* stream);

/* GUI-time edit of the palette blend, evaluated per sample INSIDE the persistent renderer (ref: palette/renderer.py:121-147
 * RegionEdit.forward, :166-183 Stylizer.forward, consulted at :474-483). All pointers are DEVICE pointers (the renderer
 * stages them in shared memory; no host synchronisation), NULL = unset.
 *   mode 1 (RegionEdit): every basis colour is recoloured in HSV (hue shift delta_hsv[b][0], saturation / value scales
 *     [b][1], [b][2]), blended with weight exp(-|xyz - mean_xyz|^2 / std_xyz) * exp(-|clip_feat - mean_clip|^2 / std_clip);
 *     weight_mode returns the weight itself as the colour.
 *   mode 2 (Stylizer): rgb = sum_b omega_b * clamp(clamp(softplus(radiance) + dI_b, 0) * (palette_b + dP_b + offsets_b *
 *     ddelta_b), 0, 1) + view_dep; the five debug maps are not produced (as in the reference). */
typedef struct pnerf_palette_edit {
    uint32_t mode, weight_mode;
    const float* delta_hsv;     /* [4,3]   mode 1 */
    const float* mean_xyz;      /* [3] or NULL */
    const float* mean_clip;     /* [clip_dim] or NULL */
    float std_xyz, std_clip;
    const float* dI;            /* [4]     mode 2 */
    const float* dP;            /* [4,3]   */
    const float* ddelta;        /* [4,3,3] */
} pnerf_palette_edit;

/* The same renderer with the field on the 5th-generation tensor cores (tcgen05.mma, activations and accumulators in TMEM;
 * csrc/field_tc.cuh): a warpgroup shades 4 rays x 32 samples per tile. A thread-per-ray pre-pass records each candidate ray's
 * occupied stretches (`runs`), so the persistent kernel never touches the occupancy grid.
 *   runs      [N * pnerf_palette_render_tc_runs_bytes()] bytes scratch
 *   t_scratch [pnerf_palette_render_tc_warps() * max_steps] fp32 scratch
 *   out_index (optional, [N] int32): ray n writes row out_index[n] of the output maps instead of row n; the maps may live in
 *             a peer GPU's memory (one view sharded over several GPUs, each rank storing its rays into the owner's image)
 *   edit      (optional): RegionEdit / Stylizer evaluated in the blend of every sample
 *   flags     0, or PNERF_RENDER_REPRODUCIBLE. By default the free lanes of a ray's last 32-sample window are filled with the
 *             first samples of the next ray in the queue (tile fill 0.90 -> 0.99, ~4 % faster); a ray's partial sums are then
 *             grouped according to its queue neighbour, and since the queue is dealt out with atomics the maps are
 *             reproducible to fp32 rounding of the compositing sums (~1e-7 relative) but not bit for bit. With the flag every
 *             ray's windows start at its own first sample: run-to-run and shard-to-shard bit-identical output.
 * Needs field->wpack_tc and field->table_sigma_palette. Replaces palette/renderer.py:430-523. */
#define PNERF_RENDER_REPRODUCIBLE 1u
PNERF_API uint32_t pnerf_palette_render_tc_warps(void);
PNERF_API uint32_t pnerf_palette_render_tc_runs_bytes(void);
PNERF_API int pnerf_palette_render_tc(const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                                      const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t Hgrid,
                                      uint32_t max_steps, float dt_gamma, float T_thresh, const pnerf_palette_field* field,
                                      float* weights_sum, float* depth, float* image, float* direct_rgb, float* view_dep_rgb,
                                      float* basis_acc, float* basis_rgb, float* unscaled_basis_rgb, float* clip_feat,
                                      uint32_t* queue, int32_t* cand, void* runs, float* t_scratch, const float* occ_aabb,
                                      const int32_t* out_index, const pnerf_palette_edit* edit, uint32_t flags, void* stream);
PNERF_API void pnerf_render_tc_timing(int enable);
PNERF_API float pnerf_render_tc_last_ms(void);
/* bench hook: CUDA-event pair around the persistent kernel of the last pnerf_palette_render_rays call (off by default) */
PNERF_API void pnerf_render_rays_timing(int enable);
PNERF_API float pnerf_render_rays_last_ms(void);

/* measurement hook for bench.py's roofline: with timing enabled, pnerf_palette_render_fused brackets its persistent
 * kernel (k_render_fused, not the pre-pass / ordering launches) with an event pair on the launch stream;
 * pnerf_render_kernel_last_ms() synchronises on it and returns the duration of the last launch (-1 if none). */
PNERF_API void pnerf_render_kernel_timing(int enable);
PNERF_API float pnerf_render_kernel_last_ms(void);

/* ------------------------------------------------------------------------------------------------
 * fused palette field for TRAINING (new entry points; they replace the field evaluation of the training branch,
 * ref: palette/renderer.py:322-359 + palette/network.py:156-280, and its autograd graph)
 *   forward : xyzs, dirs [M,3] -> sigma [M] (= density_scale * exp(logit), a constant of this stage), rgb [M,3]
 *             (= sum_b omega_b softplus(radiance)(palette_b + offsets_b) + view_dep), flex [M, 13+clip_dim+4] =
 *             [omega_sparsity, view_dep_norm, offsets_norm, smooth_norm(=0), view_dep(3), direct_rgb(3), diffuse(3),
 *              clip_feat(clip_dim), omega(4)]; every layer input is saved in xbuf (fp16 mma fragments)
 *   backward: grad_rgb [M,3], grad_flex [M,nflex], flex (forward values) -> ybuf (per-layer pre-activation gradients, fp16
 *             fragments), d_enc [M,32] / d_enc_clip [M,32] fp32 (gradients of the palette / semantic grid features, to be
 *             scattered by pnerf_grid_encode_backward with layout BLC), d_palette [4*3] (+=, may be NULL)
 *   wgrad   : xbuf, ybuf -> dwbuf (+=, fp32, zero-initialised by the caller): per layer [N_pad][K_pad] row-major in the
 *             order D0 64x16, D1 64x64, D2 16x64, V0 64x32, V1 64x64, V2 16x64, B0 64x48, B1 16x64, H 32x16
 *             [, C0 64x32, C1 16x64]; column/row maps to the nn.Linear weights: palettenerf_b200/fused_train.py
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnerf_palette_train {
    const void* table_sigma;    /* fp16 [n_entries,2] : encoder.embeddings                       */
    const void* table_palette;  /* fp16               : encoder_palette.embeddings               */
    const void* table_clip;     /* fp16               : encoder_clip.embeddings (NULL unless pred_clip) */
    const int32_t* offsets;     /* [L+1]                                                         */
    const void* wfwd;           /* fp16 B fragments of the forward layers; head bias folded into column 15 */
    const void* wbwd;           /* fp16 B fragments of the transposed layers (dX = dY W)        */
    const float* palette;       /* [4*3] basis_color clamped to [0,1]                            */
    const int32_t* m_dev;       /* optional: device-side sample count (march counter[0]); the kernels then process
                                   min(M, *m_dev) samples and M is only the capacity of the buffers */
    uint32_t L, H, pred_clip, clip_dim;
    float S, bound, density_scale;
    /* optional (may be NULL): density and palette tables interleaved entry by entry, fp16 [n_entries][2 tables][2]
     * (one 8-byte load per lattice corner serves both grids); when it is given, table_sigma / table_palette may be NULL.
     * The forward pass falls back to the separate tables if a level does not wrap with a mask. */
    const void* table_sigma_palette;
} pnerf_palette_train;

PNERF_API uint64_t pnerf_palette_train_xbuf_bytes(uint32_t M, uint32_t pred_clip);
PNERF_API uint64_t pnerf_palette_train_ybuf_bytes(uint32_t M, uint32_t pred_clip);
PNERF_API uint32_t pnerf_palette_train_dw_floats(uint32_t pred_clip);
PNERF_API uint32_t pnerf_palette_train_wfwd_units(uint32_t pred_clip);   /* uint2 units of the forward blob     */
PNERF_API uint32_t pnerf_palette_train_wbwd_units(uint32_t pred_clip);   /* uint2 units of the transposed blob  */

PNERF_API int pnerf_palette_train_forward(const float* xyzs, const float* dirs, uint32_t M, const pnerf_palette_train* p,
                                          void* xbuf, float* sigma, float* rgb, float* flex, void* stream);
PNERF_API int pnerf_palette_train_backward(uint32_t M, const pnerf_palette_train* p, const void* xbuf, void* ybuf,
                                           const float* grad_rgb, const float* grad_flex, const float* flex, float* d_enc,
                                           float* d_enc_clip, float* d_palette, void* stream);
/* flags: bit 0 = pred_clip (the clip_net layers are present); bit 1 = also compute the weight gradients of basis_net,
 * which the reference's optimizer never steps (palette/network.py:283-308 omits it from get_params) — off = those two
 * slots of dwbuf are left untouched. */
PNERF_API int pnerf_palette_train_wgrad(uint32_t M, uint32_t flags, const void* xbuf, const void* ybuf, float* dwbuf,
                                        const int32_t* m_dev, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Fused TRAINING field of the stage-1 model (csrc/nerf_train.cu). Replaces `sigmas, rgbs = self(xyzs, dirs)` of
 * NeRFRenderer.run_cuda's training branch (ref: nerf/renderer.py:289-298 -> NeRFNetwork.forward, nerf/network.py:78-124:
 * GridEncoder -> sigma_net 32-64-16 -> trunc_exp / geo_feat; SHEncoder(4) ++ geo_feat -> color_net 31-64-64-3 -> sigmoid)
 * and its autograd graph. Architecture: the reference's defaults (L = 16, F = 2 hash grid, 64-wide bias-free MLPs).
 *   forward : sigma [M] = density_scale * exp(h0) (the product nerf/renderer.py:299 forms), rgb [M,3]; layer inputs saved
 *             in xbuf (pnerf_nerf_train_xbuf_bytes)
 *   backward: (grad_sigma [M] w.r.t. the SCALED sigma, grad_rgb [M,3]) -> ybuf (per-layer pre-activation gradients) and
 *             d_enc [M,32] fp32 = gradient of the hash-grid features (scatter it with pnerf_grid_encode_backward*);
 *             trunc_exp's backward clamp (activation.py:14-17) is applied
 *   wgrad   : dwbuf (pnerf_nerf_train_dw_floats, fp32, caller zero-fills) += packed weight gradients, blocks
 *             sigma_net.0 64x32, sigma_net.1 16x64, color_net.0 64x32 (input columns [SH16 | logit(0) | geo15]),
 *             color_net.1 64x64, color_net.2 16x64 (rows 0-2)
 * wfwd / wbwd: fp16 mma B-fragment images built by palettenerf_b200/fused_nerf_train.py (pnerf_nerf_train_w*_units uint2).
 * m_dev: optional device-side sample count (static-capacity step); rows >= min(M, *m_dev) are neither read nor written.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnerf_nerf_train {
    const void* table;          /* fp16 [n_entries,2] : encoder.embeddings */
    const int32_t* offsets;     /* [L+1] */
    const void* wfwd;
    const void* wbwd;
    const int32_t* m_dev;
    uint32_t L, H;
    float S, bound, density_scale;
} pnerf_nerf_train;

PNERF_API uint64_t pnerf_nerf_train_xbuf_bytes(uint32_t M);
PNERF_API uint64_t pnerf_nerf_train_ybuf_bytes(uint32_t M);
PNERF_API uint32_t pnerf_nerf_train_dw_floats(void);
PNERF_API uint32_t pnerf_nerf_train_wfwd_units(void);
PNERF_API uint32_t pnerf_nerf_train_wbwd_units(void);
PNERF_API int pnerf_nerf_train_forward(const float* xyzs, const float* dirs, uint32_t M, const pnerf_nerf_train* p, void* xbuf,
                                       float* sigma, float* rgb, void* stream);
PNERF_API int pnerf_nerf_train_backward(uint32_t M, const pnerf_nerf_train* p, const void* xbuf, void* ybuf,
                                        const float* grad_sigma, const float* grad_rgb, const float* sigma, const float* rgb,
                                        float* d_enc, void* stream);
PNERF_API int pnerf_nerf_train_wgrad(uint32_t M, const void* xbuf, const void* ybuf, float* dwbuf, const int32_t* m_dev,
                                     void* stream);

/* ONE-pass compositor of the stage-1 training step (ref: nerf/renderer.py:301-327 = spread_ray_to_sample + the squared
 * error per sample + composite_rays_train twice, raymarching.cu:504-580, 681-761): rgb, depth, weights_sum and the error
 * channel err_map[ray] = sum_i w_i |gt[ray] - rgb_i|^2 (what the reference returns as rgb_norm) in one warp-per-ray pass.
 * gt: [N,3] target colour per ray (NULL: err_map = 0). backward: gradients w.r.t. sigmas and rgbs from (grad_weights_sum |
 * NULL, grad_image, grad_err | NULL); every sample row of every ray that fits is written (zeros behind the terminating
 * sample), so the gradient buffers need not be zero-initialised. */
PNERF_API int pnerf_nerf_composite_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                                 const float* gt, uint32_t M, uint32_t N, float T_thresh, float* weights_sum,
                                                 float* depth, float* image, float* err_map, void* stream);
PNERF_API int pnerf_nerf_composite_train_backward(const float* grad_weights_sum, const float* grad_image, const float* grad_err,
                                                  const float* sigmas, const float* rgbs, const float* deltas,
                                                  const int32_t* rays, const float* gt, const float* weights_sum,
                                                  const float* image, const float* err_map, uint32_t M, uint32_t N,
                                                  float T_thresh, float* grad_sigmas, float* grad_rgbs, void* stream);

/* ONE-pass compositor of the palette training step: composite_rays_train on (sigma, rgb) and composite_rays_flex_train
 * on the nflex auxiliary channels together (ref: raymarching.cu:504-645, palette/renderer.py:354, 387-397).
 * backward: gradients w.r.t. rgb and flex only (sigma is a constant of the palette stage); EVERY sample row of every ray
 * is written (zeros after termination), so grad_rgbs / grad_flex need not be zero-initialised. nflex must be 33. */
PNERF_API int pnerf_palette_composite_train_forward(const float* sigmas, const float* rgbs, const float* flex,
                                                    const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                                    uint32_t nflex, float T_thresh, float* weights_sum, float* depth,
                                                    float* image, float* maps, void* stream);
PNERF_API int pnerf_palette_composite_train_backward(const float* grad_image, const float* grad_maps, const float* sigmas,
                                                     const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                                     uint32_t nflex, float T_thresh, float* grad_rgbs, float* grad_flex,
                                                     void* stream);

/* pnerf_grid_encode_backward for D = 3, C = 2 with the number of points taken from device memory: processes
 * min(B, *count_dev) rows (static-capacity training step, no host synchronisation). bound > 0: `inputs` are world
 * coordinates in [-bound, bound] and are mapped to [0,1] inside the kernel like GridEncoder.forward does; bound = 0:
 * inputs are already in [0,1]. */
PNERF_API int pnerf_grid_encode_backward_counted(const void* grad, const float* inputs, const int32_t* offsets,
                                                 void* grad_embeddings, uint32_t B, uint32_t L, float S, uint32_t H,
                                                 uint32_t gridtype, int align_corners, int dtype, int grad_layout,
                                                 const int32_t* count_dev, float bound, void* stream);

/* pnerf_grid_encode_backward for an fp16 gradient with D = 3, C = 2, accumulated in an fp32 workspace
 * (replaces kernel_grid_backward<half> with its atomicAdd(__half2*), ref gridencoder/src/gridencoder.cu:226-313):
 * the reductions go to `workspace` ([n_entries, 2] fp32, contents ignored, zeroed here) as red.global.add.v2.f32 --
 * measured faster on B200's L2 than the packed f16x2 reduction, and each table entry is rounded to fp16 once instead of
 * once per contribution -- then grad_embeddings (fp16 [n_entries, 2]) += workspace. n_entries = offsets[L], even. */
PNERF_API int pnerf_grid_encode_backward_ws(const void* grad, const float* inputs, const int32_t* offsets,
                                            void* grad_embeddings, float* workspace, uint64_t n_entries, uint32_t B,
                                            uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                            int grad_layout, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ray generation  (ref: get_rays, nerf/utils.py:52-151, the arithmetic at :132-149)
 * poses [B,4,4] row-major camera-to-world, intrinsics fx fy cx cy, image H x W.
 * inds : int64 pixel indices (row * W + col), element (b, n) at inds[b * inds_batch_stride + n] (stride 0 = one [N]
 *        list shared by the batch, as the reference's `inds.expand([B, N])`); NULL = all H*W pixels in order (N = H*W).
 * Writes rays_o, rays_d [B,N,3] (rays_d normalised and rotated; rays_o materialised, the reference returns a view).
 * nears/fars (both or neither, [B,N]) additionally receive near_far_from_aabb(rays_o, rays_d, aabb, min_near)
 * (ref: raymarching.cu:95-148) computed on the direction in registers.
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                             const int64_t* inds, uint64_t inds_batch_stride, uint32_t N, uint32_t B, float* rays_o,
                             float* rays_d, const float* aabb, float min_near, float* nears, float* fars, void* stream);
/* pnerf_get_rays + the training-pixel gathers of the data loader's collate (ref: palette/provider.py:377-399) in ONE launch:
 * images [B, H*W, c_img] fp32 -> out_images [B, N, c_img], feat_images [B, H*W, c_feat] -> out_feat [B, N, c_feat], both at
 * the pixel indices the rays are generated for; either pair may be NULL. */
PNERF_API int pnerf_get_rays_collate(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                                     const int64_t* inds, uint64_t inds_batch_stride, uint32_t N, uint32_t B, float* rays_o,
                                     float* rays_d, const float* aabb, float min_near, float* nears, float* fars,
                                     const float* images, uint32_t c_img, float* out_images, const float* feat_images,
                                     uint32_t c_feat, float* out_feat, void* stream);

/* ------------------------------------------------------------------------------------------------
 * per-ray losses of the palette training step  (ref: PaletteTrainer.train_step, palette/utils.py:486-567)
 * One pass over the inputs computes the gradient of every input (for an upstream gradient of 1) and per-CTA partial
 * sums; a second tiny launch reduces them in a fixed order to the loss and its terms (deterministic).
 *   image, direct_rgb, gt_rgb [N,3]; maps [N, stride] = the renderer's channel-composite output, whose columns hold the
 *   regulariser maps (col_sparsity / col_offsets / col_view_dep / col_smooth: one column each), the semantic feature
 *   (col_clip .. +clip_dim, with gt_clip [N,clip_dim]) and the blending weights (col_basis .. +num_basis, with
 *   gt_weights [N,num_basis]); a column index of -1 switches the term off. basis_color / basis_color_origin
 *   [num_basis,3] (NULL: no palette term).
 *   terms [10] = total, rgb, direct, clip, sparsity, offsets, view_dep, smooth, weight, palette (lambda-weighted, as
 *   the reference's loss_dict); per_ray [N] (optional) = mean_c (image - gt)^2, the error-map quantity (:590).
 *   g_image, g_direct [N,3], g_maps [N, stride] (every column written; zero where no term reads), g_basis_color.
 * pnerf_scale_buffers: b_i[0..n_i) *= *scale (device scalar) for up to four buffers, one launch (the loss backward).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnerf_palette_loss_args {
    const float* image; const float* direct_rgb; const float* gt_rgb;
    const float* maps; uint32_t stride;
    int32_t col_sparsity, col_offsets, col_view_dep, col_smooth, col_clip, col_basis;
    uint32_t clip_dim, num_basis;
    const float* gt_clip; const float* gt_weights;
    const float* basis_color; const float* basis_color_origin;
    float lambda_sparsity, lambda_offsets, lambda_view_dep, lambda_smooth, lambda_weight, lambda_palette;
    uint32_t N;
    float* terms; float* per_ray;
    float* g_image; float* g_direct; float* g_maps; float* g_basis_color;
    float* partials;   /* scratch, pnerf_palette_loss_partials(N) floats, contents ignored */
} pnerf_palette_loss_args;

PNERF_API uint32_t pnerf_palette_loss_partials(uint32_t N);

PNERF_API int pnerf_palette_loss(const pnerf_palette_loss_args* args, void* stream);
/* Smooth-loss channel of the palette training field (ref: palette/renderer.py:360-381), csrc/loss.cu: from the field's channel
 * rows and the rows of the same field at jittered positions, gate = exp(-|x - x_j|^2 / bound^2 / sigma_xyz - |diffuse -
 * diffuse_j|^2 / sigma_color - |clip - clip_j| / sigma_clip) (a constant for the gradient), smooth = gate * (|omega_j - omega|^2
 * + |clip_j - clip|^2), written into column 3 of `channels` IN PLACE. count (optional): device int32, number of valid rows.
 * Backward: g_channels is updated in place (smooth column folded into the omega / clip columns, column 3 cleared),
 * g_channels_j is written for the valid rows. */
PNERF_API int pnerf_palette_smooth_forward(float* channels, const float* channels_j, const float* xyzs, const float* xyzs_j,
                                           uint32_t M, const int32_t* count, uint32_t nflex, uint32_t clip_dim,
                                           uint32_t num_basis, uint32_t pred_clip, float bound, float sigma_xyz,
                                           float sigma_color, float sigma_clip, float* gate, void* stream);
PNERF_API int pnerf_palette_smooth_backward(float* g_channels, float* g_channels_j, const float* channels,
                                            const float* channels_j, const float* gate, uint32_t M, const int32_t* count,
                                            uint32_t nflex, uint32_t clip_dim, uint32_t num_basis, uint32_t pred_clip,
                                            void* stream);
PNERF_API int pnerf_scale_buffers(float* b0, uint32_t n0, float* b1, uint32_t n1, float* b2, uint32_t n2, float* b3,
                                  uint32_t n3, const float* scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adam step of up to PNERF_ADAM_MAX_TENSORS fp32 tensors in one launch  (ref: the optimizer the
 * trainers build, palette/utils.py:719-724 + main_palette.py: optim.Adam(params, betas=(0.9, 0.99), eps=1e-15), stepped
 * through torch.cuda.amp.GradScaler). torch's fused-optimizer contract: gradients are divided by *grad_scale (device
 * scalar, may be NULL) on the fly; *found_inf != 0 (device scalar, may be NULL) skips the whole step incl. the step
 * counters. `step` (device float per tensor) holds the number of updates done so far and is advanced here.
 * lr_dev (optional device scalar) overrides lr (a learning-rate schedule inside a captured CUDA graph).
 * Non-amsgrad, maximize = false, L2 weight decay added to the gradient (torch.optim.Adam semantics).
 * ---------------------------------------------------------------------------------------------- */
#define PNERF_ADAM_MAX_TENSORS 32
typedef struct pnerf_adam_tensor {
    float* p; const float* g; float* m; float* v;   /* parameter, gradient, exp_avg, exp_avg_sq : [n] fp32 */
    const float* step;                              /* device scalar, advanced by the call               */
    uint64_t n;
    /* optional fp16 mirror of the updated parameter, written in the same pass: elements (2e, 2e+1) go as one half2 to
     * (char*)mirror + e * mirror_stride (n even, mirror 4-byte aligned, stride a multiple of 4). The fused training
     * field reads the trained hash table through such a copy (interleaved with the frozen density table: stride 8), so
     * the separate fp32 -> fp16 refresh pass over the table disappears. NULL: no mirror. */
    void* mirror;
    uint64_t mirror_stride;
} pnerf_adam_tensor;

PNERF_API int pnerf_adam_step(const pnerf_adam_tensor* tensors, uint32_t count, float lr, const float* lr_dev, float beta1,
                              float beta2, float eps, float weight_decay, const float* grad_scale, const float* found_inf,
                              void* stream);

/* GradScaler's check in front of the step (ref: scaler.step(optimizer), palette/utils.py:719-724 -> torch's
 * _amp_foreach_non_finite_check_and_unscale_ with a unit scale for an optimizer that unscales on the fly): *found_inf = 1
 * if any element of the tensors' gradients (`g`, `n` of each entry; the other fields are ignored) is not finite. The caller
 * zeroes found_inf. One streaming pass; at most PNERF_ADAM_MAX_TENSORS tensors per call. */
PNERF_API int pnerf_found_inf(const pnerf_adam_tensor* tensors, uint32_t count, float* found_inf, void* stream);

/* ------------------------------------------------------------------------------------------------
 * per-ray epilogue of run_cuda  (ref: palette/renderer.py:399-429, 525-551; nerf/renderer.py:335-343)
 *   depth_n = clamp(depth - near, 0) / (far - near) (depth_n NULL: skipped); image_out = image + (1 - weights_sum) bg;
 *   direct_out = direct + (1 - weights_sum) bg (direct_out NULL: skipped). `direct` rows are direct_stride floats apart
 *   (3 columns of a wider tensor are read in place). bg: [3] (bg_stride 0) or per ray [N,3] (bg_stride 3).
 * backward: g_weights_sum = -sum_c bg_c (g_image_c + g_direct_c) (either gradient may be NULL = zero);
 *   g_direct_full (optional) [N, stride] = zeros with g_direct in columns [col, col+3). d image = g_image unchanged.
 * ---------------------------------------------------------------------------------------------- */
PNERF_API int pnerf_render_tail_forward(uint32_t N, const float* depth, const float* nears, const float* fars,
                                        const float* image, const float* weights_sum, const float* direct,
                                        uint32_t direct_stride, const float* bg, uint32_t bg_stride, float* depth_n,
                                        float* image_out, float* direct_out, void* stream);
PNERF_API int pnerf_render_tail_backward(uint32_t N, const float* g_image, const float* g_direct, const float* bg,
                                         uint32_t bg_stride, uint32_t stride, uint32_t col, float* g_weights_sum,
                                         float* g_direct_full, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gradient all-reduce over NVLink peer memory (no reference counterpart: the reference is single-GPU)
 * peer_ptrs: HOST array of `world` device addresses, entry p = the bucket of rank p mapped into this process
 * (symmetric memory / CUDA IPC), each holding n floats, n a multiple of 4 * world. Two-shot: this rank sums slice
 * `rank` of every bucket in rank order, multiplies by `scale` (1/world to average) and stores the result into slice
 * `rank` of every bucket. The caller provides a cross-GPU barrier before (all buckets written) and after (all slices
 * delivered) the call; the kernel itself does not synchronise.
 * ---------------------------------------------------------------------------------------------- */
#define PNERF_PEER_MAX 8
PNERF_API int pnerf_peer_allreduce(const uint64_t* peer_ptrs, uint32_t world, uint32_t rank, uint64_t n, float scale,
                                   void* stream);
/* The same all-reduce with the sum formed inside the NVSwitch (NVLS): mc_ptr is the MULTICAST address of the symmetric
 * bucket; the kernel issues multimem.ld_reduce / multimem.st on slice `rank`, so only 1/world of the bucket crosses each
 * GPU's NVLink ports per direction. Same padding contract and the same two barriers around the call. */
PNERF_API int pnerf_peer_allreduce_mc(uint64_t mc_ptr, uint32_t world, uint32_t rank, uint64_t n, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNERF_B200_H_ */
