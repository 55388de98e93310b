// texture_edit.cu -- splat an edited canvas back into texel space (reference
// texture_edit.cu:11-236, texture_helpers.cuh:239-250).
//
// The kernel walks each pixel's depth-sorted list exactly like the forward rasteriser (same packed records,
// cp.async staging, warp-level culling, alpha / skip / stop rules) and, for every blended Gaussian whose view
// depth lies inside the pixel's [depth_lower, depth_upper] window, adds five values with the BILINEAR weights
// of the intersection's texel coordinate into updated_texture (X, C): rgb * alpha of the edit canvas, its alpha,
// and the constant 1 (the weight total the caller divides by).  As in the reference the splat is not weighted by
// the blend weight.  Settings bits of THIS entry point (reference texture_edit.cu:46-47): bit 0 = blur, bit 1 = ndc
// (read but unused upstream).
#include "raster.cuh"

namespace gstex {

struct EditIn {
    const float *updated_img, *updated_alpha, *depth_lower, *depth_upper;
    float *out;  // (X, C), zero-filled before the launch
    int C;       // row pitch of `out` (texture_info.z upstream), >= 5
};

template <bool BLUR>
__global__ void __launch_bounds__(RASTER_MAX_THREADS, 4) texture_edit_kernel(const RasterCommon p, const EditIn in) {
    __shared__ float4 stage[2][RASTER_BATCH * REC_PITCH];
    __shared__ uint8_t survivors[RASTER_MAX_THREADS / 32][RASTER_BATCH];

    const int tr = threadIdx.x, lane = tr & 31;
    uint8_t *__restrict__ my_list = survivors[tr >> 5];
    const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
    int lx, ly;
    tile_pixel(p.bw, tr, lx, ly);
    const int col = blockIdx.x * p.bw + lx, row = blockIdx.y * p.bw + ly;
    const bool inside = (tr < p.bw * p.bw) && col < p.img_w && row < p.img_h;
    const PixelConsts pc = make_pixel(col, row, p.c2w, p.viewmat, p.fx, p.fy, p.cx, p.cy);
    const WarpRect wr = make_warp_rect(col, row, inside);
    const int pix = inside ? row * p.img_w + col : 0;
    const float a_upd = in.updated_alpha[pix];
    const float val0 = in.updated_img[3 * pix] * a_upd, val1 = in.updated_img[3 * pix + 1] * a_upd,
                val2 = in.updated_img[3 * pix + 2] * a_upd;
    const float zlo = in.depth_lower[pix], zhi = in.depth_upper[pix];

    const int2 range = p.bins[tile];
    const int total = range.y - range.x;
    const int nbatch = (total + RASTER_BATCH - 1) / RASTER_BATCH;
    float T = 1.f;
    bool done = !inside;

    if (nbatch > 0) stage_records(stage[0], p.recs, p.ids, range.x, min(RASTER_BATCH, total), tr, p.nthreads);
    for (int b = 0; b < nbatch; ++b) {
        const int first = range.x + b * RASTER_BATCH;
        const int cnt = min(RASTER_BATCH, range.y - first);
        if (b + 1 < nbatch) {
            stage_records(stage[(b + 1) & 1], p.recs, p.ids, first + RASTER_BATCH,
                          min(RASTER_BATCH, range.y - first - RASTER_BATCH), tr, p.nthreads);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        if (__syncthreads_count(done) >= p.nthreads) break;
        const float4 *__restrict__ S = stage[b & 1];
        const int nsurv = __all_sync(0xffffffffu, done) ? 0 : build_survivors<BLUR>(S, 0, cnt, wr, p.mean2d, my_list, lane);
        if (!done) {
            for (int si = 0; si < nsurv; ++si) {
                const int i = my_list[si];
                const float4 *__restrict__ R = S + i * REC_PITCH;
                const float4 q0 = R[0], q1 = R[1], q2 = R[2], q3 = R[3];
                PairEval pe;
                eval_pair<BLUR>(q0, q1, q2, q3, pc, p.mean2d, pe);
                const float next_T = __fmul_rn(T, __fsub_rn(1.f, pe.alpha));
                if (next_T <= T_STOP) {  // reference texture_edit.cu:180-184
                    done = true;
                    break;
                }
                if (pair_skipped(pe)) continue;
                const float t_view = pe.t * pc.vdep;
                if (t_view >= zlo && t_view <= zhi) {  // texture_edit.cu:191
                    const float4 q4 = R[4], q5 = R[5], q6 = R[6];
                    const float nu = fmaf(q4.x, pe.ex, fmaf(q4.y, pe.ey, q4.z));
                    const float nv = fmaf(q5.x, pe.ex, fmaf(q5.y, pe.ey, q5.z));
                    const float u = clamp01(fmaf(nu, pe.rD, q4.w)), v = clamp01(fmaf(nv, pe.rD, q5.w));
                    TexFetch tf;
                    texel_setup(__float_as_int(q3.z), __float_as_int(q3.w), __float_as_int(q6.w), u, v, true, tf);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float w = tf.w[k];
                        if (w != 0.f) {  // a zero-weight corner adds 0 upstream
                            float *__restrict__ dst = in.out + (size_t)tf.idx[k] * in.C;
                            atomicAdd(dst + 0, w * val0);
                            atomicAdd(dst + 1, w * val1);
                            atomicAdd(dst + 2, w * val2);
                            atomicAdd(dst + 3, w * a_upd);
                            atomicAdd(dst + 4, w);
                        }
                    }
                }
                T = next_T;
            }
        }
        __syncthreads();
    }
    __pipeline_wait_prior(0);
}

struct EditLayout {
    size_t recs_off, mean2d_off, total;
};

static EditLayout edit_layout(int n) {
    EditLayout L;
    size_t off = 0;
    L.recs_off = off;
    off = align_up(off + sizeof(float) * REC_FLOATS * (size_t)(n > 0 ? n : 1), 256);
    L.mean2d_off = off;
    off = align_up(off + sizeof(float2) * (size_t)(n > 0 ? n : 1), 256);
    L.total = off;
    return L;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_texture_edit_temp_bytes(int n) { return edit_layout(n).total; }

extern "C" int gstex_texture_edit(int img_height, int img_width, int block_width, int n, int64_t num_texels,
                                  int channels, const int32_t *texture_dims, const float *updated_img,
                                  const float *updated_alpha, const float *depth_lower, const float *depth_upper,
                                  const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                  const float *opacities, const float *means, const float *scales, float glob_scale,
                                  const float *quats, const float *uv0, const float *umap, const float *vmap,
                                  const float *viewmat, const float *c2w, float fx, float fy, float cx, float cy,
                                  int settings, float *updated_texture, void *temp, size_t temp_bytes,
                                  gstex_stream_t stream) {
    int rc = check_raster_args("texture_edit", img_height, img_width, block_width, n, num_texels, 5, 0);
    if (rc != GSTEX_OK) return rc;
    GSTEX_REQUIRE(channels >= 5, GSTEX_E_INVALID,
                  "texture_edit: updated_texture needs at least 5 channels (rgb*a, a, weight), texture_info.z = %d",
                  channels);
    GSTEX_REQUIRE((settings & ~3) == 0, GSTEX_E_UNSUPPORTED,
                  "texture_edit: settings 0x%x has bits other than 0 (blur) and 1 (ndc)", settings);
    const EditLayout L = edit_layout(n);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE, "texture_edit: temp too small (%zu < %zu)",
                  temp_bytes, L.total);
    cudaStream_t s = as_stream(stream);
    GSTEX_CUDA_OK(cudaMemsetAsync(updated_texture, 0, sizeof(float) * (size_t)channels * (size_t)num_texels, s));
    if (n == 0) return GSTEX_OK;
    float4 *recs = (float4 *)((char *)temp + L.recs_off);
    float2 *mean2d = (float2 *)((char *)temp + L.mean2d_off);
    rc = launch_pack(n, means, scales, glob_scale, quats, opacities, /*colors=*/nullptr, uv0, umap, vmap, texture_dims,
                     viewmat, c2w, fx, fy, cx, cy, recs, mean2d, s);
    if (rc != GSTEX_OK) return rc;
    const bool blur = (settings & 1) != 0;
    const RasterCommon p = make_raster_common(img_height, img_width, block_width, 5, blur ? GSTEX_SET_BLUR : 0,
                                              gaussian_ids_sorted, tile_bins, recs, mean2d, nullptr, nullptr, viewmat,
                                              c2w, nullptr, fx, fy, cx, cy, nullptr);
    EditIn in{updated_img, updated_alpha, depth_lower, depth_upper, updated_texture, channels};
    const dim3 grid(p.tiles_x, ceil_div(p.img_h, p.bw));
    if (blur) texture_edit_kernel<true><<<grid, p.nthreads, 0, s>>>(p, in);
    else texture_edit_kernel<false><<<grid, p.nthreads, 0, s>>>(p, in);
    GSTEX_LAUNCH_OK("texture_edit_kernel");
    return GSTEX_OK;
}
