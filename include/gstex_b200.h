/*
 * gstex_b200.h -- C ABI of libgstex_b200.so, the B200 (sm_100a) textured-2DGS rasteriser.
 *
 * This is the drop-in boundary for the reference's pybind module `gstex_cuda`
 * (victor-rong/GStex_cuda, gstex_cuda/cuda/csrc/ext.cpp:8-25).  Every entry point
 * names the reference interface it replaces.  Rules for all entry points:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in `_host`; float = fp32, ids = int32, keys = int64;
 *   - the callee never allocates, never synchronises and never throws: outputs
 *     and scratch space are provided by the caller (query the *_bytes functions);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; 0 = legacy
 *     default stream) of the CURRENT device;
 *   - returns 0 on success, a negative GSTEX_E_* code otherwise;
 *     gstex_last_error() returns a thread-local message for the last failure.
 * Layout conventions are the reference's: viewmat / c2w 4x4 row-major,
 * quats (w,x,y,z) pre-normalised, scales linear, texture (X, C) row = texel,
 * texture_dims (N,3) int32 {h, w, first_texel}, images (H, W, ...) row-major.
 */
#ifndef GSTEX_B200_H
#define GSTEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSTEX_OK 0
#define GSTEX_E_INVALID (-1)   /* bad argument (size, block width, channel count ...) */
#define GSTEX_E_CUDA (-2)      /* a CUDA runtime call / launch failed */
#define GSTEX_E_WORKSPACE (-3) /* scratch buffer too small */
#define GSTEX_E_UNSUPPORTED (-4) /* settings bits the entry point does not implement (visualisation bits in a backward call) */

typedef void *gstex_stream_t;

const char *gstex_last_error(void);
int gstex_abi_version(void);

/* ---- settings bits, texture.cu:54-65 / :381-386 ---------------------------------------- */
#define GSTEX_SET_NEAREST (1 << 2)       /* nearest texel instead of bilinear */
#define GSTEX_SET_PROPAGATE_UV (1 << 8)  /* UV gradient flows to means / uv maps */
#define GSTEX_SET_BLUR (1 << 9)          /* 2-D screen-space blur floor */
#define GSTEX_SET_NDC (1 << 10)          /* distortion on NDC depth */
#define GSTEX_SET_SUPPORTED (GSTEX_SET_NEAREST | GSTEX_SET_PROPAGATE_UV | GSTEX_SET_BLUR | GSTEX_SET_NDC)
/* visualisation bits, forward only (texture.cu:58-63, :201-241, :269-274); the backward entry points reject them */
#define GSTEX_SET_VIS_NORMALS (1 << 15)        /* normals flipped towards the camera */
#define GSTEX_SET_VIS_ALPHA (1 << 16)          /* hard-edged footprints: alpha = 0.99 inside sigma <= alpha_bound^2 / 2 */
#define GSTEX_SET_VIS_ALPHA_BOUND (0x1f << 17) /* alpha_bound * 8 */
#define GSTEX_SET_VIS_WHITE_OUTLINE (1 << 24)  /* outline pixels add white instead of nothing */
#define GSTEX_SET_VIS_OPACITY_THRESH (1 << 25) /* hide Gaussians of opacity < 0.5 */
#define GSTEX_SET_VIS_OUTLINE_BOUND (0xf << 26) /* outline width * 4 (pixels) */
#define GSTEX_SET_VIS_ALL (GSTEX_SET_VIS_NORMALS | GSTEX_SET_VIS_ALPHA | GSTEX_SET_VIS_ALPHA_BOUND | \
                           GSTEX_SET_VIS_WHITE_OUTLINE | GSTEX_SET_VIS_OPACITY_THRESH | GSTEX_SET_VIS_OUTLINE_BOUND)
#define GSTEX_SET_SUPPORTED_FORWARD (GSTEX_SET_SUPPORTED | GSTEX_SET_VIS_ALL)

/* ======================================================================================== *
 * (1) projection, screen AABB, tile count
 * ======================================================================================== */

/* replaces get_aabb_2d_tensor, get_aabb_2d.cu:91-125 (kernel :11-89).
 * centers, extents: (n,2), fully written. */
int gstex_get_aabb_2d(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                      const float *viewmat, float fx, float fy, float cx, float cy, float *centers,
                      float *extents, gstex_stream_t stream);

/* replaces the torch ops of get_num_tiles_hit_2d, gstex_cuda/get_aabb_2d.py:70-92 (floor based). */
int gstex_num_tiles_hit_2d(int n, const float *centers, const float *extents, int img_height, int img_width,
                           int block_width, int32_t *num_tiles_hit, gstex_stream_t stream);

/* replaces the torch ops of project_points(clip=False), gstex_cuda/get_aabb_2d.py:22-32.
 * pix may be NULL. viewmat: row-major 4x4 (first 12 floats read). */
int gstex_project_points(int n, const float *means, const float *viewmat, float fx, float fy, float cx,
                         float cy, float *pix, float *depths, gstex_stream_t stream);

/* fused replacement of example.py:148-152 (project_points + get_aabb_2d + get_num_tiles_hit_2d):
 * one pass over the Gaussians.  Tile counts use the same truncation rule as the key emitter
 * (helpers.cuh:37-51), which equals the floor rule after clamping for power-of-two block widths.
 * Gaussians whose extents are both <= 1e-4 (clipped, forward.cu:32) count 0 tiles.
 * visible_count (may be NULL): n floats; += 1 for every Gaussian that hits at least one tile (the per-step visibility the
 * visible-only Adam update gates on; the caller zero-fills it once per optimiser step). */
int gstex_project_aabb_count(int n, const float *means, const float *scales, float glob_scale,
                             const float *quats, const float *viewmat, float fx, float fy, float cx, float cy,
                             int img_height, int img_width, int block_width, float *centers, float *extents,
                             float *depths, int32_t *num_tiles_hit, float *visible_count, gstex_stream_t stream);

/* ======================================================================================== *
 * (2) tile binning
 * ======================================================================================== */

/* replaces torch.cumsum(int32) of compute_cumulative_intersects, gstex_cuda/utils.py:40-59.
 * Inclusive scan; out[n-1] is the number of intersections (read it on the host only if you must). */
size_t gstex_scan_temp_bytes(int n);
int gstex_cumsum_i32(int n, const int32_t *in, int32_t *out, void *temp, size_t temp_bytes,
                     gstex_stream_t stream);

/* replaces map_gaussian_to_intersects_tensor, bindings.cu:77-121 (kernel forward.cu:13-71), wrapped=false.
 * isect_ids (m) int64 = (tile_id << 32) | depth bits, gaussian_ids (m) int32.  Slots of Gaussians the
 * emitter skips are NOT written: zero-fill the outputs first to reproduce the reference exactly.
 * num_intersects is the capacity of the two outputs; writes past it are dropped (the reference would
 * write out of bounds when cum_tiles_hit disagrees with the emitter's own tile count). */
int gstex_map_gaussian_to_intersects(int n, int64_t num_intersects, const float *centers, const float *extents,
                                     const float *depths, const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                                     int block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                     gstex_stream_t stream);
/* the same with wrapped=true (torus tile boxes, forward.cu:34-36, 53-62 + helpers.cuh:53-73, 94-111): tile boxes are not
 * clamped and tile indices are taken modulo the grid.  cum_tiles_hit must be the running sum of the wrapped box sizes
 * (the reference leaves that to the caller as well). */
int gstex_map_gaussian_to_intersects_wrapped(int n, int64_t num_intersects, const float *centers,
                                             const float *extents, const float *depths,
                                             const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                                             int block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                             gstex_stream_t stream);

/* replaces torch.sort(int64) + torch.gather, gstex_cuda/utils.py:159-160: stable ascending LSD radix sort
 * of signed 64-bit keys carrying int32 values.  `end_bit` (1..64): keys are known to be non-negative and
 * < 2^end_bit (64 = no assumption).  If d_count is non-NULL the number of valid elements is
 * min(*d_count, m) read on the device (no host sync); elements past it are left untouched. */
size_t gstex_sort_temp_bytes(int64_t m);
int gstex_sort_pairs(int64_t m, const int64_t *keys_in, const int32_t *vals_in, int64_t *keys_out,
                     int32_t *vals_out, int end_bit, const int32_t *d_count, void *temp, size_t temp_bytes,
                     gstex_stream_t stream);

/* replaces get_tile_bin_edges_tensor, bindings.cu:123-140 (kernel forward.cu:76-98).
 * tile_bins (num_tiles,2) int32 must be zero-filled by the caller (as torch::zeros does). */
int gstex_get_tile_bin_edges(int64_t m, const int64_t *isect_ids_sorted, int32_t *tile_bins,
                             const int32_t *d_count, gstex_stream_t stream);

/* Fused tile binning for the no-host-sync pipeline: replaces, in one call and with bit-identical
 * gaussian_ids_sorted / tile_bins, the chain cumsum -> map_gaussian_to_intersects -> sort -> get_tile_bin_edges
 * (gstex_cuda/utils.py:40-162, forward.cu:13-98).  Buckets the intersections by tile with atomics, then one CTA per
 * tile sorts its short list by (depth bits, gaussian id) in shared memory (csrc/binning_tiles.cu).
 * capacity: length of gaussian_ids_sorted / isect_ids_sorted; intersections past it are dropped and tile_bins is
 * clipped to it.  isect_ids_sorted may be NULL.  tile_bins (num_tiles,2) is fully written ((0,0) for empty tiles).
 * num_intersects: one int32 on the device = the true number of intersections (compare with capacity to detect
 * overflow).  max_intersects_seen (may be NULL): one int32 on the device, raised to num_intersects when that is larger -
 * a running maximum over the calls of a no-host-sync loop. */
size_t gstex_bin_tiles_temp_bytes(int num_tiles, int64_t capacity);
int gstex_bin_tiles(int n, const float *centers, const float *extents, const float *depths, int tiles_x, int tiles_y,
                    int block_width, int64_t capacity, int32_t *gaussian_ids_sorted, int64_t *isect_ids_sorted,
                    int32_t *tile_bins, int32_t *num_intersects, int32_t *max_intersects_seen, void *temp,
                    size_t temp_bytes, gstex_stream_t stream);

/* ======================================================================================== *
 * (3) rasterise forward / (4) rasterise backward
 * ======================================================================================== */

/* Scratch needed by gstex_texture_forward / gstex_texture_backward for n Gaussians, x texels with c channels and
 * num_intersects sorted-list entries (the length of gaussian_ids_sorted).  The forward scratch must stay alive and
 * untouched until the matching backward ran: it holds the per-view packed Gaussian records, the float4-padded
 * texture and the forward pass's blend masks (one 32-bit word per list entry and warp: which pixels composited
 * the entry), which is what the backward pass differentiates. */
size_t gstex_texture_forward_temp_bytes(int n, int64_t num_texels, int channels, int64_t num_intersects);
size_t gstex_texture_backward_temp_bytes(int n, int64_t num_texels, int channels);
/* temp size of a gstex_texture_backward call WITHOUT forward scratch (fwd_temp == NULL): the call then rebuilds the
 * records, the padded texture and the blend masks from its own arguments, like the reference's stateless backward. */
size_t gstex_texture_backward_stateless_temp_bytes(int n, int64_t num_texels, int channels, int64_t num_intersects);

/* replaces texture_forward_tensor, texture.cu:766-901 (kernel :11-329).
 * Outputs (all fully written): out_img (H,W,3), out_depth (H,W), out_reg (H,W), out_texture (H,W,C),
 * out_normal (H,W,3), final_Ts (H,W), final_idx (H,W) int32, depth_idx (H,W) int32, out_reg_s (H,W,3).
 * background: 3 floats on the device.  num_intersects = 0 (with the matching temp size) renders without keeping blend
 * masks - for inference-only callers and for callers that will use the stateless backward. */
int gstex_texture_forward(int img_height, int img_width, int block_width, int n, int64_t num_texels,
                          int channels, int64_t num_intersects, const int32_t *texture_dims,
                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
                          const float *opacities, const float *means, const float *scales, float glob_scale,
                          const float *quats, const float *uv0, const float *umap, const float *vmap,
                          const float *texture, const float *viewmat, const float *c2w, float fx, float fy, float cx,
                          float cy, int settings, const float *background, float *out_img, float *out_depth,
                          float *out_reg, float *out_texture, float *out_normal, float *final_Ts,
                          int32_t *final_idx, int32_t *depth_idx, float *out_reg_s, void *temp,
                          size_t temp_bytes, gstex_stream_t stream);

/* replaces texture_backward_tensor, texture.cu:915-1053 (kernel :331-760).
 * fwd_temp: the scratch the matching gstex_texture_forward call (same inputs) filled, or NULL.  The reference's backward
 * is a pure function of its arguments (texture.cuh:120-168); with fwd_temp == NULL this call is too: it re-packs the
 * records, re-pads the texture and re-derives the blend masks from final_Ts / final_idx (a cull + alpha walk, no
 * compositing), and `temp` must then hold gstex_texture_backward_stateless_temp_bytes().  Gradients (n,3) (n,1) (n,3) (n,3) (n,4) (n,1,2)
 * (n,1,3) (n,1,3) (X,C): if accumulate == 0 they are overwritten (no zero-fill needed), otherwise the
 * view's gradient is added to what they hold (multi-view accumulation). */
int gstex_texture_backward(int img_height, int img_width, int block_width, int n, int64_t num_texels,
                           int channels, int64_t num_intersects, const int32_t *texture_dims, const int32_t *gaussian_ids_sorted,
                           const int32_t *tile_bins, const float *colors, const float *opacities,
                           const float *means, const float *scales, float glob_scale, const float *quats,
                           const float *uv0, const float *umap, const float *vmap, const float *texture,
                           const float *viewmat, const float *c2w, float fx, float fy, float cx, float cy,
                           int settings, const float *background, const float *final_Ts,
                           const int32_t *final_idx, const int32_t *depth_idx, const float *final_s,
                           const float *v_out_img, const float *v_out_depth, const float *v_out_reg,
                           const float *v_out_alpha, const float *v_out_texture, const float *v_out_normal,
                           float *v_colors, float *v_opacity, float *v_means, float *v_scales, float *v_quats,
                           float *v_uv0, float *v_umap, float *v_vmap, float *v_texture, int accumulate,
                           const void *fwd_temp, void *temp, size_t temp_bytes, gstex_stream_t stream);

/* replaces texture_edit_tensor, texture_edit.cu:238-354 (kernel :11-236): walks every pixel's
 * depth-sorted list like the forward pass and, for each blended Gaussian whose view depth lies inside the pixel's
 * [depth_lower, depth_upper] window, adds rgb*alpha, alpha and 1 of the edit canvas with the bilinear weights of the
 * intersection's texel into updated_texture (num_texels, channels), which the call zero-fills first.
 * channels = texture_info.z of the reference call (the row pitch of the output), >= 5; channels 5.. stay zero.
 * updated_img (H,W,3), updated_alpha / depth_lower / depth_upper (H,W).
 * settings: bit 0 = blur, bit 1 = ndc (read but unused upstream) -- NOT the rasteriser's bit positions
 * (texture_edit.cu:46-47). */
size_t gstex_texture_edit_temp_bytes(int n);
int gstex_texture_edit(int img_height, int img_width, int block_width, int n, int64_t num_texels, int channels,
                       const int32_t *texture_dims, const float *updated_img, const float *updated_alpha,
                       const float *depth_lower, const float *depth_upper, const int32_t *gaussian_ids_sorted,
                       const int32_t *tile_bins, const float *opacities, const float *means, const float *scales,
                       float glob_scale, const float *quats, const float *uv0, const float *umap, const float *vmap,
                       const float *viewmat, const float *c2w, float fx, float fy, float cx, float cy, int settings,
                       float *updated_texture, void *temp, size_t temp_bytes, gstex_stream_t stream);

/* ======================================================================================== *
 * spherical harmonics, texture sampling
 * ======================================================================================== */

/* replaces compute_sh_forward_tensor / compute_sh_backward_tensor, bindings.cu:18-75 (sh.cuh:212-253).
 * coeffs (n,K,3) with K = num_sh_bases(degree); colors (n,3).  Backward writes all K rows (rows beyond
 * num_sh_bases(degrees_to_use) are zero); with accumulate != 0 it adds instead. */
int gstex_sh_forward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *coeffs,
                     float *colors, gstex_stream_t stream);
int gstex_sh_backward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *v_colors,
                      float *v_coeffs, int accumulate, gstex_stream_t stream);

/* replaces texture_sample_forward_tensor / texture_sample_backward_tensor, texture_sample.cu:72-141.
 * uv is clamped to [0,1] (documented behaviour, texture_sample.py:26-30).  The backward is the intended
 * scatter (the reference kernel reads the wrong buffer, texture_sample.cu:58, and is unreachable):
 * v_texture (X,C) must be zero-filled by the caller. */
int gstex_texture_sample_forward(int num_queries, int channels, const int32_t *texture_dims, const float *uvs,
                                 const float *texture, float *output, gstex_stream_t stream);
int gstex_texture_sample_backward(int num_queries, int channels, const int32_t *texture_dims, const float *uvs,
                                  const float *v_output, float *v_texture, gstex_stream_t stream);

/* ======================================================================================== *
 * fused per-view training step (project -> bin -> sort -> raster fwd -> loss -> raster bwd), no host sync
 * ======================================================================================== */

/* Fused image loss of example.py:189-209 and its gradient w.r.t. the rasteriser outputs:
 *   loss = mean((out_texture - gt)^2) + mean(out_reg) + mean(nx^2 + ny^2 + (1-nz)^2)
 * loss_accum[0] += loss (atomic, one add per block); v_* are fully written.  v_out_img / v_out_depth / v_out_alpha (all
 * zero for this loss) may be NULL: not written. */
int gstex_image_loss(int img_height, int img_width, const float *out_texture, const float *out_reg,
                     const float *out_normal, const float *gt, float *loss_accum, float *v_out_img,
                     float *v_out_depth, float *v_out_reg, float *v_out_alpha, float *v_out_texture,
                     float *v_out_normal, gstex_stream_t stream);

/* ---- staged entry points (host side: gstex_cuda_b200/pipeline.py) ---------------------------------------
 * The stages gstex_texture_forward / gstex_texture_backward run internally, exposed so that a multi-view step
 * pads the texture once, packs records per view and accumulates all views into one gradient arena.
 * recs: n x 32 floats, mean2d: n x 2 floats, acc: n x 32 floats (zero-filled per view: gstex_pack_records does it in
 * the same pass when given acc_to_zero, else the caller),
 * masks: mask_entries x 8 uint32 blend masks written by gstex_raster_forward (which first zeroes the words of the first
 * min(*d_num_intersects, mask_entries) entries; d_num_intersects NULL = all) and read by gstex_raster_backward,
 * tex / vtex: (X,4) padded when channels == 3, else the caller's (X,C) layout; vtex is accumulated into. */
int gstex_pad_texture(int64_t num_texels, const float *texture, float *tex4, gstex_stream_t stream);
int gstex_unpad_texture_grad(int64_t num_texels, const float *g4, float *v_texture, int accumulate,
                             gstex_stream_t stream);
int gstex_pack_records(int n, const int32_t *texture_dims, const float *colors, const float *opacities,
                       const float *means, const float *scales, float glob_scale, const float *quats,
                       const float *uv0, const float *umap, const float *vmap, const float *viewmat,
                       const float *c2w, float fx, float fy, float cx, float cy, float *recs, float *mean2d,
                       float *acc_to_zero, gstex_stream_t stream);
int gstex_raster_forward(int img_height, int img_width, int block_width, int channels, int settings,
                         const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                         const float *mean2d, const float *tex, const float *viewmat, const float *c2w, float fx,
                         float fy, float cx, float cy, const float *background, float *out_img, float *out_depth,
                         float *out_reg, float *out_texture, float *out_normal, float *final_Ts,
                         int32_t *final_idx, int32_t *depth_idx, float *out_reg_s, uint32_t *masks,
                         int64_t mask_entries, const int32_t *d_num_intersects, gstex_stream_t stream);
/* Re-derives the blend masks of a finished forward pass from its saved state: pixel p composited list entry i iff
 * i <= final_idx[p] and the pair passes the skip test (what the reference's backward re-evaluates, texture.cu:484-558).
 * For callers that did not keep the masks gstex_raster_forward wrote. */
int gstex_raster_masks(int img_height, int img_width, int block_width, int settings,
                       const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                       const float *mean2d, const float *viewmat, const float *c2w, float fx, float fy, float cx,
                       float cy, const float *final_Ts, const int32_t *final_idx, uint32_t *masks,
                       int64_t mask_entries, const int32_t *d_num_intersects, gstex_stream_t stream);
/* The upstream gradients v_out_img / v_out_depth / v_out_reg / v_out_alpha / v_out_normal, and v_out_texture when
 * channels == 3, may be NULL: an output the caller's loss does not use (zeros, without the buffer). */
int gstex_raster_backward(int img_height, int img_width, int block_width, int channels, int settings,
                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                          const float *mean2d, const float *tex, const float *viewmat, const float *c2w, float fx,
                          float fy, float cx, float cy, const float *background, const float *final_Ts,
                          const int32_t *final_idx, const int32_t *depth_idx, const float *final_s,
                          const float *v_out_img, const float *v_out_depth, const float *v_out_reg,
                          const float *v_out_alpha, const float *v_out_texture, const float *v_out_normal,
                          const uint32_t *masks, float *acc, float *vtex, gstex_stream_t stream);
int gstex_raster_epilogue(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                          const float *umap, const float *vmap, const float *viewmat, const float *c2w, float fx,
                          float fy, float cx, float cy, const float *acc, const float *recs, float *v_colors,
                          float *v_opacity, float *v_means, float *v_scales, float *v_quats, float *v_uv0,
                          float *v_umap, float *v_vmap, int accumulate, gstex_stream_t stream);
/* recs (may be NULL): the records gstex_pack_records wrote for this view; the scale gradients then use the very form
 * coefficients the rasterisers used (they are computed in double for grazing surfels). */
/* accumulate bit 0: add to the geometry gradients (means ... vmap, opacity) instead of overwriting them; bit 1: the same
 * for v_colors alone (per-view colour gradients that feed an SH backward are overwritten while the rest accumulates). */

/* cudaMemsetAsync(ptr, 0, bytes) on the stream: lets a host layer that owns no CUDA runtime binding (ctypes) clear
 * its buffers in stream order. */
int gstex_fill_zero(void *ptr, size_t bytes, gstex_stream_t stream);

/* Fused view-dependent colour around the SH op: colors = clamp(SH(means - camera origin) + 0.5, 0, 1);
 * mask (n bytes) records which channels were not clamped and gates the backward. */
int gstex_sh_colors_forward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                            const float *coeffs, float *colors, uint8_t *mask, gstex_stream_t stream);
int gstex_sh_colors_backward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                             const float *v_colors, const uint8_t *mask, float *v_coeffs, int accumulate,
                             gstex_stream_t stream);

/* ======================================================================================== *
 * training-step glue around the rasteriser
 * ======================================================================================== */

/* replaces the ~25 torch ops of example.py:126-143 + the sigmoids of :162-163 (one launch):
 *   scales = (exp(raw.x), exp(raw.y), 1e-5 * mean of the two), quats = raw / |raw|,
 *   uv0 = mapping[:, :2], umap = e^m2 (a1 cos m3 + a2 sin m3), vmap = e^m2 (-a1 sin m3 + a2 cos m3) with a1, a2 the
 *   first two columns of R(quats), colors = sigmoid(raw_rgbs) (skipped when raw_rgbs == colors == NULL, e.g. when
 *   colours come from spherical harmonics), opacities = sigmoid(raw_opacities).
 * raw_scales (n,3), raw_quats (n,4), mapping (n,1,4) = (u0, v0, log uv-scale, theta), raw_rgbs (n,3), raw_opacities (n,1). */
int gstex_preprocess_forward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                             const float *raw_rgbs, const float *raw_opacities, float *scales, float *quats,
                             float *uv0, float *umap, float *vmap, float *colors, float *opacities,
                             gstex_stream_t stream);

/* the VJP of gstex_preprocess_forward (what torch autograd does for example.py:126-143): gradients w.r.t. the
 * activated parameters in, gradients w.r.t. the raw parameters out (overwritten).  v_raw_scales[:, 2] = 0: the
 * thickness is detached upstream (example.py:128). */
int gstex_preprocess_backward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                              const float *raw_rgbs, const float *raw_opacities, const float *v_scales,
                              const float *v_quats, const float *v_uv0, const float *v_umap, const float *v_vmap,
                              const float *v_colors, const float *v_opacity, float *v_raw_scales, float *v_raw_quats,
                              float *v_mapping, float *v_raw_rgbs, float *v_raw_opacities, gstex_stream_t stream);

/* torch.sigmoid(texture) (example.py:171) fused into the float4 padding pass, and its VJP fused into the un-padding
 * pass: v_raw (+)= g4 * t * (1 - t) with t read from tex4.  3-channel textures. */
int gstex_sigmoid_pad_texture(int64_t num_texels, const float *raw_texture, float *tex4, gstex_stream_t stream);
int gstex_unpad_texture_grad_sigmoid(int64_t num_texels, const float *g4, const float *tex4, float *v_raw_texture,
                                     int accumulate, gstex_stream_t stream);

/* torch.optim.Adam's update (example.py:223-225, :278; betas / eps as given, no weight decay, no amsgrad) over a
 * contiguous fp32 range: one launch for a whole parameter arena.  step is 1-based; grads are multiplied by
 * grad_scale first (1/world_size or 1/views for a mean).  lr / betas / eps are doubles because torch forms
 * 1 - beta, the bias corrections and lr / bc1 in Python doubles before rounding to fp32. */
int gstex_adam_step(int64_t count, float *params, const float *grads, float *exp_avg, float *exp_avg_sq, double lr,
                    double beta1, double beta2, double eps, int step, float grad_scale, gstex_stream_t stream);

/* The same update with the step counter (and the scalars derived from it) in device memory, for CUDA-graph replay of a
 * whole optimiser step: `state` is gstex_adam_state_bytes() bytes, zero-filled before the first step; every call
 * increments the counter on the device (one extra 1-thread launch), then applies the update with that step's bias
 * corrections - bit-identical to gstex_adam_step(step = 1, 2, 3, ...). */
size_t gstex_adam_state_bytes(void);
int gstex_adam_step_device(int64_t count, float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                           double lr, double beta1, double beta2, double eps, void *state, float grad_scale,
                           gstex_stream_t stream);
/* Visible-only Adam ("sparse / visible-only update"): gstex_adam_prepare_device advances the
 * device-resident step counter once per optimiser step; gstex_adam_apply_rows_device then updates one field of the
 * arena, touching only the elements whose row has visible[row] > 0 (parameters and both moments of unseen rows are
 * left alone).  Element e belongs to unit e / unit_width and the unit to row owner[unit] (owner == NULL: row = unit) -
 * e.g. unit_width 3 and owner = texel -> Gaussian for a jagged texture.  visible: one float per row, e.g. the number of
 * views of the step whose tile count for the Gaussian was non-zero (gstex_project_aabb_count fills it). */
int gstex_adam_prepare_device(void *state, double lr, double beta1, double beta2, double eps, float grad_scale,
                              gstex_stream_t stream);
int gstex_adam_apply_rows_device(int64_t count, float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                                 const void *state, const float *visible, int unit_width, const int32_t *owner,
                                 gstex_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSTEX_B200_H */
