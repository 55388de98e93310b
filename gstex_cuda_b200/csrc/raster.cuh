// raster.cuh -- device code shared by the forward and backward rasterisers.
//
// Per (pixel, Gaussian) pair the kernels evaluate three affine forms in the pixel offset e = p - p_c
// from the packed record (common.cuh: RecSlot), instead of the reference's per-pixel quat->R,
// ray-plane intersection and delta chain (texture.cu:161-200):
//     N1 = c1 + P1.e   N2 = c2 + P2.e   D = c3 + A3.e
//     l1' = N1/D, l2' = N2/D           (in-plane coordinates, pre-scaled by sqrt(log2(e)/2)/(s*glob))
//     alpha = min(0.99, opac * 2^-(l1'^2 + l2'^2))        == min(0.99, opac * exp(-sigma))
//     t = (c0/D) * |R_w(p)|                                (Euclidean ray distance, texture.cu:170-171)
// That is ~20 flops + 2 MUFU per pair instead of ~100.  tests/formulation.py is the float64 model,
// tests/test_formulation.py proves it equal to the reference algebra (oracle) incl. all gradients.
//
// Forward and backward call the SAME eval_pair() with explicit fmaf / __f*_rn so that both make
// bit-identical skip decisions.
#pragma once
#include <cuda_pipeline.h>

#include "common.cuh"

namespace gstex {

#ifndef GSTEX_RASTER_BATCH
#define GSTEX_RASTER_BATCH 128
#endif
constexpr int RASTER_BATCH = GSTEX_RASTER_BATCH;  // Gaussians per shared-memory stage (<= 256: uint8 survivor indices)
constexpr int RASTER_MAX_THREADS = 256;
constexpr int RASTER_MAX_C = 64;       // generic-channel path (reference texture.cu:102-103)

struct PixelConsts {
    float px, py;   // pixel centre
    float rn;       // |R_w|, norm of the un-normalised world ray c2w_rot * ((px-cx)/fx, (py-cy)/fy, 1)
    float vdep;     // viewmat[2,:3] . normalised ray   (reference texture.cu:74)
    float eps;      // 1e-6 * rn  (plane-denominator clamp, texture_helpers.cuh:304-311)
};

__device__ __forceinline__ PixelConsts make_pixel(int col, int row, const float *__restrict__ c2w,
                                                  const float *__restrict__ viewmat, float fx, float fy,
                                                  float cx, float cy) {
    PixelConsts pc;
    pc.px = (float)col + 0.5f;
    pc.py = (float)row + 0.5f;
    const float u = __fdiv_rn(pc.px - cx, fx), v = __fdiv_rn(pc.py - cy, fy);
    const Vec3 rw = rot_apply(c2w, mk3(u, v, 1.f));
    pc.rn = sqrtf(dot3(rw, rw));
    const float inv = __fdiv_rn(1.f, pc.rn);
    pc.vdep = (viewmat[8] * rw.x + viewmat[9] * rw.y + viewmat[10] * rw.z) * inv;
    pc.eps = 1e-6f * pc.rn;
    return pc;
}

struct PairEval {
    float ex, ey;   // pixel offset from the record's expansion centre
    float rD;       // 1 / D (clamped)
    float l1, l2;   // scaled in-plane coordinates
    float e;        // 2^-(l1^2+l2^2)
    float f;        // d(alpha_raw)/d(opac): e, or exp(-sigma_blur) when the blur branch is taken
    float alpha;    // min(0.99, opac * f)
    float s;        // c0 / D  (depth along the un-normalised ray)
    float t;        // s * rn
    bool blur;      // blur branch selected (BLUR builds only)
    float bx, by;   // projected mean - pixel (BLUR builds only)
};

template <bool BLUR>
__device__ __forceinline__ void eval_pair(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                          const PixelConsts &pc, const float2 *__restrict__ mean2d,
                                          PairEval &o) {
    o.ex = __fsub_rn(pc.px, q0.x);
    o.ey = __fsub_rn(pc.py, q0.y);
    const float n1 = fmaf(q1.x, o.ex, fmaf(q1.y, o.ey, q1.z));
    const float n2 = fmaf(q2.x, o.ex, fmaf(q2.y, o.ey, q2.z));
    float d = fmaf(q3.x, o.ex, fmaf(q3.y, o.ey, q1.w));
    if (fabsf(d) < pc.eps) d = copysignf(pc.eps, d);
    o.rD = fast_rcp(d);
    o.l1 = __fmul_rn(n1, o.rD);
    o.l2 = __fmul_rn(n2, o.rD);
    const float qq = fmaf(o.l1, o.l1, __fmul_rn(o.l2, o.l2));
    o.e = fast_exp2(-qq);
    o.f = o.e;
    o.blur = false;
    o.bx = o.by = 0.f;
    if (BLUR) {
        // reference texture.cu:184-197: sigma_blur = 0.5 * 2 * |xy_mean - p|^2, taken when it is the smaller
        const float2 m2 = mean2d[__float_as_int(q2.w)];
        o.bx = __fsub_rn(m2.x, pc.px);
        o.by = __fsub_rn(m2.y, pc.py);
        const float sb = fmaf(o.bx, o.bx, __fmul_rn(o.by, o.by));
        if (sb < __fmul_rn(qq, LN2_F)) {
            o.blur = true;
            o.f = __expf(-sb);
        }
    }
    o.alpha = fminf(ALPHA_CAP, __fmul_rn(q0.w, o.f));
    o.s = __fmul_rn(q0.z, o.rD);
    o.t = __fmul_rn(o.s, pc.rn);
}

__device__ __forceinline__ bool pair_skipped(const PairEval &pe) {
    return pe.t < T_NEAR || pe.t > T_FAR || pe.alpha < ALPHA_MIN;  // reference texture.cu:213
}

// Bilinear / nearest texel addressing with replicate padding (texture_helpers.cuh:155-215).
// idx[k] are TEXEL indices (not multiplied by the channel count).
struct TexFetch {
    int idx[4];
    float w[4];
    float fu, fv;
    int h, wd;
};

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

__device__ __forceinline__ void texel_setup(int h, int w, int si, float u, float v, bool bilinear, TexFetch &f) {
    const float tu = (float)h * u, tv = (float)w * v;
    int i0 = (int)tu, j0 = (int)tv;
    const int i1 = min(i0 + 1, h - 1), j1 = min(j0 + 1, w - 1);
    const float fu = tu - (float)i0, fv = tv - (float)j0;
    i0 = min(i0, h - 1);
    j0 = min(j0, w - 1);
    const float w00 = (1.f - fu) * (1.f - fv), w01 = (1.f - fu) * fv, w10 = fu * (1.f - fv), w11 = fu * fv;
    f.idx[0] = si + i0 * w + j0;
    f.idx[1] = si + i0 * w + j1;
    f.idx[2] = si + i1 * w + j0;
    f.idx[3] = si + i1 * w + j1;
    if (bilinear) {
        f.w[0] = w00; f.w[1] = w01; f.w[2] = w10; f.w[3] = w11;
    } else {  // largest weight, first wins on ties (texture_helpers.cuh:199-212)
        int pick = 3;
        if (w00 >= w01 && w00 >= w10 && w00 >= w11) pick = 0;
        else if (w01 >= w00 && w01 >= w10 && w01 >= w11) pick = 1;
        else if (w10 >= w00 && w10 >= w01 && w10 >= w11) pick = 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) f.w[k] = (k == pick) ? 1.f : 0.f;
    }
    f.fu = fu; f.fv = fv; f.h = h; f.wd = w;
}

// Thread -> pixel mapping inside a tile.  For 16x16 tiles every warp owns an 8x4 pixel patch
// (compact footprint => fewer (warp, Gaussian) pairs survive the any-valid test in the backward pass).
__device__ __forceinline__ void tile_pixel(int bw, int tr, int &lx, int &ly) {
    if (bw == 16) {
        const int w = tr >> 5, l = tr & 31;
        lx = ((w & 1) << 3) + (l & 7);
        ly = ((w >> 1) << 2) + (l >> 3);
    } else {
        lx = tr % bw;
        ly = tr / bw;
    }
}

// Shared-memory stage layout: quad k of staged record r lives at float4 slot r*9 + k (records padded from 8 to 9
// quads).  The odd pitch makes the per-lane reads of the culling pass (lane <-> record) bank-conflict free - lane r
// reads 16-byte bank group (r + k) mod 8 - while the warp-uniform reads of the compositing loop stay broadcasts, and a
// record's quads sit at compile-time offsets from one base address (no per-quad address arithmetic).
constexpr int REC_PITCH = 9;
__device__ __forceinline__ int quad_slot(int r, int k) { return r * REC_PITCH + k; }

// Stage `cnt` records (ids[first..first+cnt)) into shared memory with 16-byte cp.async copies: eight
// consecutive threads fetch one 128-byte record (one cache line), so both the global reads and the
// shared-memory writes are fully coalesced / conflict-free.
__device__ __forceinline__ void stage_records(float4 *__restrict__ dst, const float4 *__restrict__ recs,
                                              const int32_t *__restrict__ ids, int first, int cnt, int tr,
                                              int nthreads) {
    for (int c = tr; c < cnt * 8; c += nthreads) {
        const int r = c >> 3, k = c & 7;
        const int g = ids[first + r];
        __pipeline_memcpy_async(dst + quad_slot(r, k), recs + (size_t)g * 8 + k, 16);
    }
    __pipeline_commit();
}

// ------------------------------------------------------------------------------------------
// Warp-level culling.  Each warp owns a small pixel rectangle (8x4 for 16x16 tiles).  Before a staged
// batch is composited, lane j tests records j, j+32, ... against the rectangle with an interval bound of
// the three affine forms:  |N1| >= |N1(c)| - r1,  |N2| >= |N2(c)| - r2,  |D| <= |D(c)| + r3  over the
// rectangle (c = its centre, r* = |P.x| hx + |P.y| hy), hence  q = (N1^2+N2^2)/D^2 >= q_min  and
// alpha <= opac * 2^-q_min.  A record whose bound is below 0.0039 (< 1/255, 0.55 % margin for the
// approximate ex2 / rounding) is skipped by every pixel of the warp (reference texture.cu:213), so the
// warp never evaluates it.  Survivor indices are compacted per warp, order preserved.
// The reference tests the stop rule T(1-alpha) <= 1e-4 even for Gaussians it skips; a culled Gaussian (alpha < 1/255) is
// never evaluated here.  That cannot change any output: see the monotonicity argument in raster_forward.cu.
// ------------------------------------------------------------------------------------------
struct WarpRect {
    float cx, cy, hx, hy;  // centre and half extents (pixel centres) of the warp's live pixels
    bool any;              // at least one live pixel
};

__device__ __forceinline__ WarpRect make_warp_rect(int col, int row, bool inside) {
    const unsigned full = 0xffffffffu;
    const int big = 1 << 28;
    const int x0 = __reduce_min_sync(full, inside ? col : big), x1 = __reduce_max_sync(full, inside ? col : -big);
    const int y0 = __reduce_min_sync(full, inside ? row : big), y1 = __reduce_max_sync(full, inside ? row : -big);
    WarpRect r;
    r.any = x1 >= x0;
    r.cx = 0.5f * (float)(x0 + x1) + 0.5f;
    r.cy = 0.5f * (float)(y0 + y1) + 0.5f;
    r.hx = 0.5f * (float)(x1 - x0);
    r.hy = 0.5f * (float)(y1 - y0);
    return r;
}

constexpr float CULL_ALPHA = 0.0039f;  // < 1/255 = 0.0039216

template <bool BLUR>
__device__ __forceinline__ bool record_culled(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                              const WarpRect &wr, const float2 *__restrict__ mean2d) {
    const float ex = wr.cx - q0.x, ey = wr.cy - q0.y;
    const float n1 = fmaf(q1.x, ex, fmaf(q1.y, ey, q1.z));
    const float n2 = fmaf(q2.x, ex, fmaf(q2.y, ey, q2.z));
    const float d = fmaf(q3.x, ex, fmaf(q3.y, ey, q1.w));
    const float r1 = fmaf(fabsf(q1.x), wr.hx, fabsf(q1.y) * wr.hy);
    const float r2 = fmaf(fabsf(q2.x), wr.hx, fabsf(q2.y) * wr.hy);
    const float r3 = fmaf(fabsf(q3.x), wr.hx, fabsf(q3.y) * wr.hy);
    const float m1 = fmaxf(fabsf(n1) - r1, 0.f), m2 = fmaxf(fabsf(n2) - r2, 0.f);
    const float m3 = fmaxf(fabsf(d) + r3, 1e-5f);  // >= the kernels' clamp 1e-6*|R_w|
    const float rd = fast_rcp(m3);
    const float l1 = m1 * rd, l2 = m2 * rd;
    float bound = q0.w * fast_exp2(-fmaf(l1, l1, l2 * l2));
    if (BLUR) {  // the blur branch can only raise alpha: opac * exp(-|mean2d - p|^2)
        const float2 m = mean2d[__float_as_int(q2.w)];
        const float dx = fmaxf(fabsf(m.x - wr.cx) - wr.hx, 0.f), dy = fmaxf(fabsf(m.y - wr.cy) - wr.hy, 0.f);
        bound = fmaxf(bound, q0.w * __expf(-fmaf(dx, dx, dy * dy)));
    }
    return bound < CULL_ALPHA;
}

// Builds the warp's survivor list for staged records [lo, hi) of stage S; returns the survivor count.
template <bool BLUR, bool CULL = true>
__device__ __forceinline__ int build_survivors(const float4 *__restrict__ S, int lo, int hi, const WarpRect &wr,
                                               const float2 *__restrict__ mean2d, uint8_t *__restrict__ list,
                                               int lane) {
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    int n = 0;
    for (int k = lo; k < hi; k += 32) {
        const int r = k + lane;
        bool keep = !CULL && r < hi;  // CULL = false: the alpha-visualisation mode redefines alpha, keep everything
        if (CULL && r < hi) {
            const float4 q0 = S[quad_slot(r, 0)], q1 = S[quad_slot(r, 1)], q2 = S[quad_slot(r, 2)],
                         q3 = S[quad_slot(r, 3)];
            keep = !record_culled<BLUR>(q0, q1, q2, q3, wr, mean2d);
        }
        const unsigned m = __ballot_sync(full, keep);
        if (keep) list[n + __popc(m & lt)] = (uint8_t)r;
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

// Pull the first lines of a Gaussian's texture block towards the SM before the compositing loop needs them: the
// texel fetch is the one long-latency load of a blended pair (4 x 16 B inside a block of h*w*16 B; 256 B for the
// 4x4 textures of the C4 scene).
__device__ __forceinline__ void prefetch_texture_block(const float4 *__restrict__ tex4, int tex0, int h, int w) {
    const char *base = reinterpret_cast<const char *>(tex4 + tex0);
    const int bytes = h * w * 16;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base));
    if (bytes > 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + 128));
    if (bytes > 256) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + bytes - 16));
}

struct RasterCommon {
    int img_w, img_h, tiles_x, bw, nthreads, settings, channels;
    const int32_t *ids;        // gaussian_ids_sorted
    const int2 *bins;          // tile_bins
    const float4 *recs;        // packed records (n x 8 float4)
    const float2 *mean2d;      // projected means (blur only)
    const float4 *tex4;        // padded texture (X x float4), channels == 3
    const float *tex;          // caller's texture (X x C), generic channel count
    const float *viewmat, *c2w, *background;
    float fx, fy, cx, cy;
    // One 32-bit word per (sorted-list entry, warp of the tile's CTA): the lanes (pixels) that BLENDED the entry in the
    // forward pass.  Written by the forward rasteriser (or re-derived from final_Ts / final_idx by raster_masks_kernel),
    // read by the backward one, which therefore differentiates exactly the pairs the forward pass composited (no
    // re-evaluation of the skip / stop rules, no culling pass).
    uint32_t *masks;
};
constexpr int MASK_WARPS = RASTER_MAX_THREADS / 32;  // mask words per list entry

struct ForwardOut {
    float *out_img, *out_depth, *out_reg, *out_texture, *out_normal, *final_Ts, *out_reg_s;
    int32_t *final_idx, *depth_idx;
};

struct BackwardIn {
    const float *final_Ts, *final_s;
    const int32_t *final_idx, *depth_idx;
    const float *v_img, *v_depth, *v_reg, *v_alpha, *v_tex, *v_normal;
};

struct BackwardOut {
    float4 *acc;    // n x 8 float4 moment lines (zero-filled before the launch)
    float4 *vtex4;  // X x float4 texel gradients (channels == 3; accumulated into)
    float *vtex;    // X x C texel gradients (generic channel count; accumulated into)
};

struct FwdLayout {
    size_t recs_off, mean2d_off, tex4_off, masks_off, total;
};

// host-side launchers shared between the reference-shaped entry points and the staged ones (pipeline.cu)
RasterCommon make_raster_common(int img_height, int img_width, int block_width, int channels, int settings,
                                const int32_t *ids, const int32_t *tile_bins, const float4 *recs, const float2 *mean2d,
                                const float4 *tex4, const float *tex, const float *viewmat, const float *c2w,
                                const float *background, float fx, float fy, float cx, float cy, uint32_t *masks);
// zeroes the mask words of the first min(*d_count, mask_entries) list entries (d_count == NULL: all mask_entries),
// then rasterises
int launch_raster_forward(const RasterCommon &p, const ForwardOut &o, int64_t mask_entries, const int32_t *d_count,
                          cudaStream_t s);
// rebuilds the blend masks from a finished forward pass's final_Ts / final_idx (raster_forward.cu: raster_masks_kernel)
int launch_raster_masks(const RasterCommon &p, const float *final_Ts, const int32_t *final_idx, int64_t mask_entries,
                        const int32_t *d_count, cudaStream_t s);
int launch_raster_backward(const RasterCommon &p, const BackwardIn &in, const BackwardOut &o, cudaStream_t s);
int launch_pack(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                const float *opacities, const float *colors, const float *uv0, const float *umap, const float *vmap,
                const int32_t *texture_dims, const float *viewmat, const float *c2w, float fx, float fy, float cx,
                float cy, float4 *recs, float2 *mean2d, cudaStream_t s, float4 *acc_to_zero = nullptr);
int launch_epilogue(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                    const float *umap, const float *vmap, const float *viewmat, const float *c2w, float fx, float fy,
                    float cx, float cy, const float4 *acc, float *v_colors, float *v_opacity, float *v_means,
                    float *v_scales, float *v_quats, float *v_uv0, float *v_umap, float *v_vmap, int accumulate,
                    cudaStream_t s, const float4 *recs = nullptr);  // accumulate: see gstex_raster_epilogue; recs: the view's records
int launch_pad_texture(int64_t num_texels, const float *tex, float4 *tex4, cudaStream_t s);
int launch_unpad_texture_grad(int64_t num_texels, const float4 *g4, float *v_texture, int accumulate, cudaStream_t s);
FwdLayout forward_layout(int n, int64_t num_texels, int channels, int64_t num_intersects);
int check_raster_args(const char *who, int img_height, int img_width, int block_width, int n, int64_t num_texels,
                      int channels, int settings, int supported_settings = GSTEX_SET_SUPPORTED);

}  // namespace gstex
