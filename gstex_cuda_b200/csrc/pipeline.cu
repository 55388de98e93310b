// pipeline.cu -- staged entry points for the fused multi-view training step (no host synchronisation).
//
// The reference-shaped calls (gstex_texture_forward / gstex_texture_backward) pack the per-view records and
// pad the texture on every call because the reference API is stateless.  A trainer that renders many views
// per optimiser step pads the texture once, keeps one texel-gradient buffer for the whole step and only
// re-packs the per-view records; these entry points expose the stages for that (host side:
// gstex_cuda_b200/pipeline.py).
#include "raster.cuh"

using namespace gstex;

extern "C" int gstex_fill_zero(void *ptr, size_t bytes, gstex_stream_t stream) {
    GSTEX_REQUIRE(ptr != nullptr || bytes == 0, GSTEX_E_INVALID, "fill_zero: NULL pointer");
    if (bytes) GSTEX_CUDA_OK(cudaMemsetAsync(ptr, 0, bytes, as_stream(stream)));
    return GSTEX_OK;
}

extern "C" int gstex_pad_texture(int64_t num_texels, const float *texture, float *tex4, gstex_stream_t stream) {
    GSTEX_REQUIRE(num_texels >= 0, GSTEX_E_INVALID, "pad_texture: texels = %lld", (long long)num_texels);
    return launch_pad_texture(num_texels, texture, (float4 *)tex4, as_stream(stream));
}

extern "C" int gstex_unpad_texture_grad(int64_t num_texels, const float *g4, float *v_texture, int accumulate,
                                        gstex_stream_t stream) {
    GSTEX_REQUIRE(num_texels >= 0, GSTEX_E_INVALID, "unpad_texture_grad: texels = %lld", (long long)num_texels);
    return launch_unpad_texture_grad(num_texels, (const float4 *)g4, v_texture, accumulate, as_stream(stream));
}

extern "C" int gstex_pack_records(int n, const int32_t *texture_dims, const float *colors, const float *opacities,
                                  const float *means, const float *scales, float glob_scale, const float *quats,
                                  const float *uv0, const float *umap, const float *vmap, const float *viewmat,
                                  const float *c2w, float fx, float fy, float cx, float cy, float *recs,
                                  float *mean2d, float *acc_to_zero, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "pack_records: n = %d", n);
    return launch_pack(n, means, scales, glob_scale, quats, opacities, colors, uv0, umap, vmap, texture_dims, viewmat,
                       c2w, fx, fy, cx, cy, (float4 *)recs, (float2 *)mean2d, as_stream(stream), (float4 *)acc_to_zero);
}

extern "C" int gstex_raster_forward(int img_height, int img_width, int block_width, int channels, int settings,
                                    const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                                    const float *mean2d, const float *tex, const float *viewmat, const float *c2w,
                                    float fx, float fy, float cx, float cy, const float *background, float *out_img,
                                    float *out_depth, float *out_reg, float *out_texture, float *out_normal,
                                    float *final_Ts, int32_t *final_idx, int32_t *depth_idx, float *out_reg_s,
                                    uint32_t *masks, int64_t mask_entries, const int32_t *d_num_intersects,
                                    gstex_stream_t stream) {
    int rc = check_raster_args("raster_forward", img_height, img_width, block_width, 0, 0, channels, settings,
                               GSTEX_SET_SUPPORTED_FORWARD);
    if (rc != GSTEX_OK) return rc;
    const RasterCommon p = make_raster_common(img_height, img_width, block_width, channels, settings,
                                              gaussian_ids_sorted, tile_bins, (const float4 *)recs,
                                              (const float2 *)mean2d, (const float4 *)tex, tex, viewmat, c2w,
                                              background, fx, fy, cx, cy, masks);
    ForwardOut o{out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, out_reg_s, final_idx, depth_idx};
    return launch_raster_forward(p, o, mask_entries, d_num_intersects, as_stream(stream));
}

extern "C" int gstex_raster_masks(int img_height, int img_width, int block_width, int settings,
                                  const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                                  const float *mean2d, const float *viewmat, const float *c2w, float fx, float fy,
                                  float cx, float cy, const float *final_Ts, const int32_t *final_idx, uint32_t *masks,
                                  int64_t mask_entries, const int32_t *d_num_intersects, gstex_stream_t stream) {
    int rc = check_raster_args("raster_masks", img_height, img_width, block_width, 0, 0, 3, settings);
    if (rc != GSTEX_OK) return rc;
    GSTEX_REQUIRE(masks && final_Ts && final_idx, GSTEX_E_INVALID, "raster_masks: NULL masks / final_Ts / final_idx");
    const RasterCommon p = make_raster_common(img_height, img_width, block_width, 3, settings, gaussian_ids_sorted,
                                              tile_bins, (const float4 *)recs, (const float2 *)mean2d, nullptr, nullptr,
                                              viewmat, c2w, nullptr, fx, fy, cx, cy, masks);
    return launch_raster_masks(p, final_Ts, final_idx, mask_entries, d_num_intersects, as_stream(stream));
}

extern "C" int gstex_raster_backward(int img_height, int img_width, int block_width, int channels, int settings,
                                     const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *recs,
                                     const float *mean2d, const float *tex, const float *viewmat, const float *c2w,
                                     float fx, float fy, float cx, float cy, const float *background,
                                     const float *final_Ts, const int32_t *final_idx, const int32_t *depth_idx,
                                     const float *final_s, const float *v_out_img, const float *v_out_depth,
                                     const float *v_out_reg, const float *v_out_alpha, const float *v_out_texture,
                                     const float *v_out_normal, const uint32_t *masks, float *acc, float *vtex,
                                     gstex_stream_t stream) {
    int rc = check_raster_args("raster_backward", img_height, img_width, block_width, 0, 0, channels, settings);
    if (rc != GSTEX_OK) return rc;
    const RasterCommon p = make_raster_common(img_height, img_width, block_width, channels, settings,
                                              gaussian_ids_sorted, tile_bins, (const float4 *)recs,
                                              (const float2 *)mean2d, (const float4 *)tex, tex, viewmat, c2w,
                                              background, fx, fy, cx, cy, const_cast<uint32_t *>(masks));
    GSTEX_REQUIRE(masks != nullptr, GSTEX_E_INVALID, "raster_backward: masks (written by raster_forward) is NULL");
    BackwardIn in{final_Ts, final_s, final_idx, depth_idx, v_out_img, v_out_depth, v_out_reg, v_out_alpha, v_out_texture,
                  v_out_normal};
    BackwardOut o{(float4 *)acc, (float4 *)vtex, vtex};
    return launch_raster_backward(p, in, o, as_stream(stream));
}

extern "C" int gstex_raster_epilogue(int n, const float *means, const float *scales, float glob_scale,
                                     const float *quats, const float *umap, const float *vmap, const float *viewmat,
                                     const float *c2w, float fx, float fy, float cx, float cy, const float *acc,
                                     const float *recs, float *v_colors, float *v_opacity, float *v_means, float *v_scales,
                                     float *v_quats, float *v_uv0, float *v_umap, float *v_vmap, int accumulate,
                                     gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "raster_epilogue: n = %d", n);
    return launch_epilogue(n, means, scales, glob_scale, quats, umap, vmap, viewmat, c2w, fx, fy, cx, cy,
                           (const float4 *)acc, v_colors, v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap,
                           accumulate, as_stream(stream), (const float4 *)recs);
}
