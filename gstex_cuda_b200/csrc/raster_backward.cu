// raster_backward.cu -- rasterise backward (reference texture.cu:331-760).
//
// Same CTA / warp / stage organisation as the forward pass, walking each tile's list back to front
// from the last Gaussian any pixel of the CTA blended.  What changes against the reference:
//   * per pair the kernel differentiates the three affine forms of raster.cuh, not the quat->R /
//     ray-plane chain: a lane produces 25 moment values (sum g*(ex, ey, 1) per form, plus colour,
//     normal, opacity, c0, uv0), laid out as 8 float4 quads (common.cuh: AccSlot);
//   * those 32 slots are reduced over the warp with a recursive-halving butterfly (16+8+4 shuffles
//     exchange half of the slots each, then 2x4 to finish a quad): 36 shuffles instead of the
//     reference's 25 x 5 = 125, after which 8 lanes issue ONE 16-byte vector reduction each
//     (REDG.E.ADD.F32x4) into the Gaussian's 128-byte moment line - 8 vector atomics per
//     (warp, Gaussian) instead of 25 scalar ones, and none of the per-pair quat/rotation VJPs;
//   * the chain rule from moments to means / scales / quats / uv maps runs once per Gaussian in
//     pack.cu::epilogue_kernel, without atomics;
//   * texel gradients are one float4 vector reduction per bilinear corner into a padded (X,4) buffer
//     (4 per blended pair instead of 4*C scalar atomics), corners with zero weight are skipped.
#include "raster.cuh"

namespace gstex {

// Sum the 32 per-lane slots over the warp.  Returns, in every lane, the totals of quad (lane >> 2).
__device__ __forceinline__ float4 warp_reduce_slots(float (&a)[32], int lane) {
    const unsigned full = 0xffffffffu;
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float send = hi ? a[i] : a[i + 16];
            const float keep = hi ? a[i + 16] : a[i];
            a[i] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool hi = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? a[i] : a[i + 8];
            const float keep = hi ? a[i + 8] : a[i];
            a[i] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool hi = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = hi ? a[i] : a[i + 4];
            const float keep = hi ? a[i + 4] : a[i];
            a[i] = keep + __shfl_xor_sync(full, send, 4);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] += __shfl_xor_sync(full, a[i], 2);
        a[i] += __shfl_xor_sync(full, a[i], 1);
    }
    return make_float4(a[0], a[1], a[2], a[3]);
}

template <bool C3, bool BLUR>
__global__ void __launch_bounds__(RASTER_MAX_THREADS) raster_backward_kernel(const RasterCommon p, const BackwardIn in,
                                                                            const BackwardOut o) {
    __shared__ float4 stage[2][RASTER_BATCH * 8];
    __shared__ uint8_t survivors[RASTER_MAX_THREADS / 32][RASTER_BATCH];
    __shared__ int block_last;

    const int tr = threadIdx.x, lane = tr & 31;
    const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
    int lx, ly;
    tile_pixel(p.bw, tr, lx, ly);
    const int col = blockIdx.x * p.bw + lx, row = blockIdx.y * p.bw + ly;
    const bool inside = (tr < p.bw * p.bw) && col < p.img_w && row < p.img_h;
    const int pix = inside ? row * p.img_w + col : 0;
    const PixelConsts pc = make_pixel(col, row, p.c2w, p.viewmat, p.fx, p.fy, p.cx, p.cy);
    const WarpRect wr = make_warp_rect(col, row, inside);
    uint8_t *__restrict__ my_list = survivors[tr >> 5];
    const bool use_ndc = (p.settings & GSTEX_SET_NDC) != 0;
    const bool bilinear = !(p.settings & GSTEX_SET_NEAREST);
    const bool prop_uv = (p.settings & GSTEX_SET_PROPAGATE_UV) != 0;
    const int C = C3 ? 3 : p.channels;

    const int2 range = p.bins[tile];
    const int bfinal = inside ? in.final_idx[pix] : -1;
    if (tr == 0) block_last = -1;
    __syncthreads();
    if (bfinal >= 0) atomicMax(&block_last, bfinal);
    __syncthreads();
    const int hi = min(range.y, block_last + 1);
    const int total = hi - range.x;
    if (total <= 0) return;
    const int nbatch = (total + RASTER_BATCH - 1) / RASTER_BATCH;
    int warp_last = bfinal;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, off));

    float T = in.final_Ts[pix];
    const float Sf0 = in.final_s[3 * pix], Sf1 = in.final_s[3 * pix + 1], Sf2 = in.final_s[3 * pix + 2];
    const int dfinal = in.depth_idx[pix];
    const float vi0 = in.v_img[3 * pix], vi1 = in.v_img[3 * pix + 1], vi2 = in.v_img[3 * pix + 2];
    const float vn0 = in.v_normal[3 * pix], vn1 = in.v_normal[3 * pix + 1], vn2 = in.v_normal[3 * pix + 2];
    const float v_dep = in.v_depth[pix], v_reg = in.v_reg[pix];
    float vt0 = 0.f, vt1 = 0.f, vt2 = 0.f;
    if (C3) {
        vt0 = in.v_tex[3 * pix];
        vt1 = in.v_tex[3 * pix + 1];
        vt2 = in.v_tex[3 * pix + 2];
    }
    const float *__restrict__ vtp = in.v_tex + (size_t)C * pix;
    float v_T_run = p.background[0] * vi0 + p.background[1] * vi1 + p.background[2] * vi2 - in.v_alpha[pix];

    // batch j covers [first_j, first_j + cnt_j) counted from the back of [range.x, hi)
    auto batch_first = [&](int j) { return max(range.x, hi - (j + 1) * RASTER_BATCH); };
    auto batch_count = [&](int j) { return (hi - j * RASTER_BATCH) - batch_first(j); };

    stage_records(stage[0], p.recs, p.ids, batch_first(0), batch_count(0), tr, p.nthreads);

    for (int b = 0; b < nbatch; ++b) {
        const int first = batch_first(b), cnt = batch_count(b);
        if (b + 1 < nbatch) {
            stage_records(stage[(b + 1) & 1], p.recs, p.ids, batch_first(b + 1), batch_count(b + 1), tr, p.nthreads);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const float4 *__restrict__ S = stage[b & 1];
        // warp-level culling (raster.cuh): same survivor set as the forward pass, walked back to front
        const int nsurv = build_survivors<BLUR>(S, 0, min(cnt, warp_last - first + 1), wr, p.mean2d, my_list, lane);
        for (int si = nsurv - 1; si >= 0; --si) {
            const int i = my_list[si];
            const int idx = first + i;
            const int sw = i & 7;
            const float4 *__restrict__ R = S + (i << 3);
            const float4 q0 = R[sw], q1 = R[1 ^ sw], q2 = R[2 ^ sw], q3 = R[3 ^ sw];
            bool valid = inside && idx <= bfinal;
            PairEval pe;
            if (valid) {
                eval_pair<BLUR>(q0, q1, q2, q3, pc, p.mean2d, pe);
                valid = !pair_skipped(pe);
            }
            if (!__any_sync(0xffffffffu, valid)) continue;

            float a[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] = 0.f;
            if (valid) {
                const float4 q4 = R[4 ^ sw], q5 = R[5 ^ sw], q6 = R[6 ^ sw], q7 = R[7 ^ sw];
                const float alpha = pe.alpha;
                T *= 1.f / (1.f - alpha);  // reference texture.cu:579-580
                const float vis = alpha * T;
                a[A_CR] = vis * vi0;
                a[A_CG] = vis * vi1;
                a[A_CB] = vis * vi2;
                a[A_NX] = vis * vn0;
                a[A_NY] = vis * vn1;
                a[A_NZ] = vis * vn2;
                float v_vis = q6.x * vi0 + q6.y * vi1 + q6.z * vi2 + q7.x * vn0 + q7.y * vn1 + q7.z * vn2;

                // texture fetch VJP (reference texture.cu:594-642, texture_helpers.cuh:252-300)
                const float nu = fmaf(q4.x, pe.ex, fmaf(q4.y, pe.ey, q4.z));
                const float nv = fmaf(q5.x, pe.ex, fmaf(q5.y, pe.ey, q5.z));
                const float du = nu * pe.rD, dv = nv * pe.rD;
                const float u = clamp01(fmaf(nu, pe.rD, q4.w)), v = clamp01(fmaf(nv, pe.rD, q5.w));
                TexFetch tf;
                texel_setup(__float_as_int(q3.z), __float_as_int(q3.w), __float_as_int(q6.w), u, v, bilinear, tf);
                float v_u = 0.f, v_v = 0.f;
                if (C3) {
                    const float4 t0 = __ldg(p.tex4 + tf.idx[0]), t1 = __ldg(p.tex4 + tf.idx[1]);
                    const float4 t2 = __ldg(p.tex4 + tf.idx[2]), t3 = __ldg(p.tex4 + tf.idx[3]);
                    const float vv0 = vis * vt0, vv1 = vis * vt1, vv2 = vis * vt2;
                    if (tf.w[0] != 0.f) atomicAdd(o.vtex4 + tf.idx[0], make_float4(tf.w[0] * vv0, tf.w[0] * vv1, tf.w[0] * vv2, 0.f));
                    if (tf.w[1] != 0.f) atomicAdd(o.vtex4 + tf.idx[1], make_float4(tf.w[1] * vv0, tf.w[1] * vv1, tf.w[1] * vv2, 0.f));
                    if (tf.w[2] != 0.f) atomicAdd(o.vtex4 + tf.idx[2], make_float4(tf.w[2] * vv0, tf.w[2] * vv1, tf.w[2] * vv2, 0.f));
                    if (tf.w[3] != 0.f) atomicAdd(o.vtex4 + tf.idx[3], make_float4(tf.w[3] * vv0, tf.w[3] * vv1, tf.w[3] * vv2, 0.f));
                    const float val0 = tf.w[0] * t0.x + tf.w[1] * t1.x + tf.w[2] * t2.x + tf.w[3] * t3.x;
                    const float val1 = tf.w[0] * t0.y + tf.w[1] * t1.y + tf.w[2] * t2.y + tf.w[3] * t3.y;
                    const float val2 = tf.w[0] * t0.z + tf.w[1] * t1.z + tf.w[2] * t2.z + tf.w[3] * t3.z;
                    v_vis += val0 * vt0 + val1 * vt1 + val2 * vt2;
                    if (bilinear && prop_uv) {
                        const float ofu = 1.f - tf.fu, ofv = 1.f - tf.fv;
                        const float gu0 = -ofv * t0.x - tf.fv * t1.x + ofv * t2.x + tf.fv * t3.x;
                        const float gu1 = -ofv * t0.y - tf.fv * t1.y + ofv * t2.y + tf.fv * t3.y;
                        const float gu2 = -ofv * t0.z - tf.fv * t1.z + ofv * t2.z + tf.fv * t3.z;
                        const float gv0 = -ofu * t0.x + ofu * t1.x - tf.fu * t2.x + tf.fu * t3.x;
                        const float gv1 = -ofu * t0.y + ofu * t1.y - tf.fu * t2.y + tf.fu * t3.y;
                        const float gv2 = -ofu * t0.z + ofu * t1.z - tf.fu * t2.z + tf.fu * t3.z;
                        v_u = (float)tf.h * (vv0 * gu0 + vv1 * gu1 + vv2 * gu2);
                        v_v = (float)tf.wd * (vv0 * gv0 + vv1 * gv1 + vv2 * gv2);
                    }
                } else {
                    const float *__restrict__ tx = p.tex;
                    for (int c = 0; c < C; ++c) {
                        const float c00 = __ldg(tx + (size_t)tf.idx[0] * C + c), c01 = __ldg(tx + (size_t)tf.idx[1] * C + c);
                        const float c10 = __ldg(tx + (size_t)tf.idx[2] * C + c), c11 = __ldg(tx + (size_t)tf.idx[3] * C + c);
                        const float vtc = vtp[c];
                        const float vv = vis * vtc;
                        if (tf.w[0] != 0.f) atomicAdd(o.vtex + (size_t)tf.idx[0] * C + c, tf.w[0] * vv);
                        if (tf.w[1] != 0.f) atomicAdd(o.vtex + (size_t)tf.idx[1] * C + c, tf.w[1] * vv);
                        if (tf.w[2] != 0.f) atomicAdd(o.vtex + (size_t)tf.idx[2] * C + c, tf.w[2] * vv);
                        if (tf.w[3] != 0.f) atomicAdd(o.vtex + (size_t)tf.idx[3] * C + c, tf.w[3] * vv);
                        v_vis += (tf.w[0] * c00 + tf.w[1] * c01 + tf.w[2] * c10 + tf.w[3] * c11) * vtc;
                        if (bilinear && prop_uv) {
                            v_u += (float)tf.h * (vv * (-(1.f - tf.fv) * c00 - tf.fv * c01 + (1.f - tf.fv) * c10 + tf.fv * c11));
                            v_v += (float)tf.wd * (vv * (-(1.f - tf.fu) * c00 + (1.f - tf.fu) * c01 - tf.fu * c10 + tf.fu * c11));
                        }
                    }
                }

                // alpha / transmittance recurrences and distortion (reference texture.cu:650-670)
                float v_alpha = T * v_vis - T * v_T_run;
                float v_T_cur = alpha * v_vis + (1.f - alpha) * v_T_run;
                const float t_view = pe.t * pc.vdep;
                const float t_ndc = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view);
                const float tv = use_ndc ? t_ndc : pe.t;
                const float v_tv = 2.f * (vis * tv * Sf0 - vis * Sf1) * v_reg;  // FINAL sums, helpers.cuh:266-269
                const float v_w = (tv * tv * Sf0 - 2.f * tv * Sf1 + Sf2) * v_reg;
                v_alpha += v_w * T;
                v_T_cur += v_w * alpha;
                v_T_run = v_T_cur;
                float v_t = use_ndc ? 0.f : v_tv;
                float v_tview = (idx == dfinal && dfinal != -1) ? v_dep : 0.f;  // texture.cu:678-680
                if (use_ndc) v_tview += (T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view * t_view) * v_tv;
                v_t += pc.vdep * v_tview;
                const float v_s = v_t * pc.rn;

                // alpha = min(.99, opac * f): the cap is not masked (reference texture.cu:672, :707)
                a[A_OPAC] = pe.f * v_alpha;
                const float v_q = pe.blur ? 0.f : -LN2_F * q0.w * pe.e * v_alpha;
                const float v_l1 = 2.f * pe.l1 * v_q, v_l2 = 2.f * pe.l2 * v_q;
                const float gN1 = v_l1 * pe.rD, gN2 = v_l2 * pe.rD, gNu = v_u * pe.rD, gNv = v_v * pe.rD;
                const float gD = -(v_l1 * pe.l1 + v_l2 * pe.l2 + v_s * pe.s + v_u * du + v_v * dv) * pe.rD;
                a[A_G1X] = gN1 * pe.ex; a[A_G1Y] = gN1 * pe.ey; a[A_G1C] = gN1;
                a[A_C0] = v_s * pe.rD;
                a[A_G2X] = gN2 * pe.ex; a[A_G2Y] = gN2 * pe.ey; a[A_G2C] = gN2;
                a[A_G3X] = gD * pe.ex; a[A_G3Y] = gD * pe.ey; a[A_G3C] = gD;
                a[A_GUX] = gNu * pe.ex; a[A_GUY] = gNu * pe.ey; a[A_GUC] = gNu;
                a[A_U0] = v_u;
                a[A_GVX] = gNv * pe.ex; a[A_GVY] = gNv * pe.ey; a[A_GVC] = gNv;
                a[A_V0] = v_v;
                if (BLUR) {
                    if (pe.blur) {  // reference texture.cu:683-692
                        const float v_sb = -q0.w * pe.f * v_alpha;
                        a[A_MX] = 2.0f * v_sb * pe.bx;
                        a[A_MY] = 2.0f * v_sb * pe.by;
                    }
                }
            }
            const float4 tot = warp_reduce_slots(a, lane);
            if ((lane & 3) == 0 && (BLUR || lane < 28)) {
                const int g = __float_as_int(q2.w);
                atomicAdd(o.acc + (size_t)g * 8 + (lane >> 2), tot);
            }
        }
        __syncthreads();
    }
}

int launch_raster_backward(const RasterCommon &p, const BackwardIn &in, const BackwardOut &o, cudaStream_t s) {
    const dim3 grid(p.tiles_x, ceil_div(p.img_h, p.bw));
    const bool blur = (p.settings & GSTEX_SET_BLUR) != 0;
    if (p.channels == 3) {
        if (blur) raster_backward_kernel<true, true><<<grid, p.nthreads, 0, s>>>(p, in, o);
        else raster_backward_kernel<true, false><<<grid, p.nthreads, 0, s>>>(p, in, o);
    } else {
        if (blur) raster_backward_kernel<false, true><<<grid, p.nthreads, 0, s>>>(p, in, o);
        else raster_backward_kernel<false, false><<<grid, p.nthreads, 0, s>>>(p, in, o);
    }
    GSTEX_LAUNCH_OK("raster_backward_kernel");
    return GSTEX_OK;
}

struct BwdLayout {
    size_t acc_off, vtex4_off, total;
};

static BwdLayout backward_layout(int n, int64_t num_texels, int channels) {
    BwdLayout L;
    size_t off = 0;
    L.acc_off = off;
    off = align_up(off + sizeof(float) * ACC_FLOATS * (size_t)(n > 0 ? n : 1), 256);
    L.vtex4_off = off;
    if (channels == 3) off = align_up(off + sizeof(float4) * (size_t)(num_texels > 0 ? num_texels : 1), 256);
    L.total = off;
    return L;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_texture_backward_temp_bytes(int n, int64_t num_texels, int channels) {
    return backward_layout(n, num_texels, channels).total;
}

extern "C" int gstex_texture_backward(
    int img_height, int img_width, int block_width, int n, int64_t num_texels, int channels,
    const int32_t *texture_dims, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
    const float *opacities, const float *means, const float *scales, float glob_scale, const float *quats,
    const float *uv0, const float *umap, const float *vmap, const float *texture, const float *viewmat,
    const float *c2w, float fx, float fy, float cx, float cy, int settings, const float *background,
    const float *final_Ts, const int32_t *final_idx, const int32_t *depth_idx, const float *final_s,
    const float *v_out_img, const float *v_out_depth, const float *v_out_reg, const float *v_out_alpha,
    const float *v_out_texture, const float *v_out_normal, float *v_colors, float *v_opacity, float *v_means,
    float *v_scales, float *v_quats, float *v_uv0, float *v_umap, float *v_vmap, float *v_texture, int accumulate,
    const void *fwd_temp, void *temp, size_t temp_bytes, gstex_stream_t stream) {
    (void)texture_dims; (void)colors; (void)opacities; (void)uv0;
    int rc = check_raster_args("texture_backward", img_height, img_width, block_width, n, num_texels, channels, settings);
    if (rc != GSTEX_OK) return rc;
    GSTEX_REQUIRE(fwd_temp != nullptr, GSTEX_E_INVALID, "texture_backward: fwd_temp (forward scratch) is NULL");
    const FwdLayout FL = forward_layout(n, num_texels, channels);
    const BwdLayout L = backward_layout(n, num_texels, channels);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE, "texture_backward: temp too small (%zu < %zu)",
                  temp_bytes, L.total);
    cudaStream_t s = as_stream(stream);
    const char *fbase = (const char *)fwd_temp;
    char *base = (char *)temp;
    float4 *acc = (float4 *)(base + L.acc_off);
    float4 *vtex4 = (float4 *)(base + L.vtex4_off);
    GSTEX_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(float) * ACC_FLOATS * (size_t)n, s));
    if (channels == 3) {
        GSTEX_CUDA_OK(cudaMemsetAsync(vtex4, 0, sizeof(float4) * (size_t)num_texels, s));
    } else if (!accumulate) {
        GSTEX_CUDA_OK(cudaMemsetAsync(v_texture, 0, sizeof(float) * (size_t)channels * (size_t)num_texels, s));
    }

    const RasterCommon p = make_raster_common(
        img_height, img_width, block_width, channels, settings, gaussian_ids_sorted, tile_bins,
        (const float4 *)(fbase + FL.recs_off), (const float2 *)(fbase + FL.mean2d_off),
        (const float4 *)(fbase + FL.tex4_off), texture, viewmat, c2w, background, fx, fy, cx, cy);
    BackwardIn in{final_Ts, final_s, final_idx, depth_idx, v_out_img, v_out_depth, v_out_reg, v_out_alpha, v_out_texture,
                  v_out_normal};
    BackwardOut o{acc, vtex4, v_texture};
    rc = launch_raster_backward(p, in, o, s);
    if (rc != GSTEX_OK) return rc;
    rc = launch_epilogue(n, means, scales, glob_scale, quats, umap, vmap, viewmat, c2w, fx, fy, cx, cy, acc, v_colors,
                         v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap, accumulate, s);
    if (rc != GSTEX_OK) return rc;
    if (channels == 3) {
        rc = launch_unpad_texture_grad(num_texels, vtex4, v_texture, accumulate, s);
        if (rc != GSTEX_OK) return rc;
    }
    return GSTEX_OK;
}
