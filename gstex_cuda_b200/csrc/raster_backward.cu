// raster_backward.cu -- rasterise backward (reference texture.cu:331-760).
//
// Same CTA / warp / stage organisation as the forward pass, walking each tile's list back to front
// from the last Gaussian any pixel of the CTA blended.  What changes against the reference:
//   * per pair the kernel differentiates the three affine forms of raster.cuh, not the quat->R /
//     ray-plane chain: a lane produces 25 moment values (sum g*(ex, ey, 1) per form, plus colour,
//     normal, opacity, c0, uv0), laid out as 8 float4 quads (common.cuh: AccSlot);
//   * those 32 slots are reduced over the warp by recursive halving (16+8+4+2+1 = 31 shuffles, each step exchanges
//     half of the slots a lane still holds) instead of the reference's 25 x 5 = 125, after which lane L holds the total
//     of slot L and the warp issues ONE reduction instruction (28 lanes x 4 bytes, contiguous in the Gaussian's
//     128-byte moment line) - instead of 25 scalar atomics to 9 arrays, and none of the per-pair quat/rotation VJPs;
//   * the chain rule from moments to means / scales / quats / uv maps runs once per Gaussian in
//     pack.cu::epilogue_kernel, without atomics;
//   * texel gradients are one float4 vector reduction per bilinear corner into a padded (X,4) buffer
//     (4 per blended pair instead of 4*C scalar atomics), corners with zero weight are skipped.
#include "raster.cuh"

namespace gstex {

// ------------------------------------------------------------------------------------------
// Dense pair-queue backward.
//
// The forward pass recorded, per (list entry, warp), the mask of pixels that blended the entry.  On C4 a blended
// (warp, Gaussian) pair has on average 11 of 32 pixels set, so a kernel that runs the ~400-instruction gradient
// path "one Gaussian at a time, lane = pixel" wastes two thirds of its issue slots.  Instead each warp
//   1. compacts the set bits of a run of entries (back to front) into a queue of (entry, pixel) PAIRS,
//   2. D1, lane = pair : evaluates alpha and the texel fetch, and y = dL/d(vis-weighted value) of the pair
//                        (colour, normal, texture and distortion terms),
//   3. scan, lane = pixel: walks its own pairs in list order: T_k = T_{k+1}/(1-alpha_k),
//                        v_alpha = T_k (y - R), R <- alpha y + (1-alpha) R   (reference texture.cu:579-670),
//   4. D2, lane = pair : everything that needs v_alpha: texel-gradient reductions, gradients of the five
//                        affine forms, the 28 moment values; rows go to shared memory and 2 x 14 lanes sum the
//                        column pairs over the rows of each Gaussian, one RED.64 per column pair and Gaussian.
// Per-pixel constants (upstream gradients, final distortion sums ...) stay in the registers of the pixel's lane
// and reach the pair lanes by warp shuffles.  Per pixel the pairs are processed in exactly the reference's order.
// ------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = RASTER_MAX_THREADS / 32;
// tuning knobs (overridable with -D for experiments; the defaults are what ships)
#ifndef GSTEX_BWD_QCAP
#define GSTEX_BWD_QCAP 128
#endif
#ifndef GSTEX_BWD_ECAP
#define GSTEX_BWD_ECAP 16
#endif
#ifndef GSTEX_BWD_BATCH
#define GSTEX_BWD_BATCH 64
#endif
#ifndef GSTEX_BWD_MINB
#define GSTEX_BWD_MINB 3
#endif
constexpr int BWD_QCAP = GSTEX_BWD_QCAP;    // pairs per chunk (4 dense iterations)
constexpr int BWD_ECAP = GSTEX_BWD_ECAP;    // list entries (Gaussians) per chunk: their records are staged per warp
constexpr int BWD_BATCH = GSTEX_BWD_BATCH;  // list entries whose mask words a warp fetches at a time
#ifndef GSTEX_BWD_FULL
#define GSTEX_BWD_FULL 14
#endif
constexpr int BWD_FULL = GSTEX_BWD_FULL;  // entries covering at least this many of the 32 pixels run lane = pixel
#ifndef GSTEX_BWD_CTA_WARPS
#define GSTEX_BWD_CTA_WARPS 1
#endif
// Warps per CTA (1, 2, 4 or 8; 8 = one CTA per 16x16 tile).  One-warp CTAs: measured 2.51 -> 2.37 ms on C4 (2 warps: 2.46,
// 4 warps: 2.47) - the warps of a tile finish at very different times, and a CTA's slots are only refilled when its
// slowest warp retires.
constexpr int BWD_CTA_WARPS = GSTEX_BWD_CTA_WARPS;
static_assert(BWD_WARPS % BWD_CTA_WARPS == 0, "a tile's 8 warps are split evenly over its CTAs");

template <bool BLUR>
struct BwdWarpSmem {
    static constexpr int PITCH = BLUR ? 32 : 28;  // floats per moment row (28 used; 30 with BLUR)
    float4 rec[BWD_ECAP * REC_PITCH];          // packed records of the chunk's entries (quad_slot layout)
    float rows[32 * PITCH];            // moment rows of one dense iteration (D2)
    float2 e_ay[BWD_QCAP];             // (alpha, y) (D1) -> (vis = alpha * T_k, v_alpha) (scan): one 8-byte access each way
    float2 e_g[BWD_QCAP];              // d(texture term)/du, /dv per unit vis (D1)
    int row_gid[32];                   // Gaussian id of each row
    uint32_t sv_mask[BWD_BATCH];       // blend mask of each non-empty entry of the batch
#ifndef GSTEX_BWD_NO_SVGID
    int32_t sv_gid[BWD_BATCH];         // its Gaussian id
#endif
    uint16_t q_ent[BWD_QCAP];          // pair -> (chunk-local entry | pixel lane << 8)
    uint16_t ch_off[BWD_ECAP];         // first queue slot of each chunk entry
    uint8_t sv_r[BWD_BATCH];           // its position inside the batch
};

struct PixelShare {  // per-pixel values a pair lane fetches from the pixel's lane
    float vi0, vi1, vi2, vn0, vn1, vn2, vt0, vt1, vt2, Sf0, Sf1, Sf2, v_reg, v_dep;
    int dfinal, pix;
};

struct PairFlags {
    bool use_ndc, bilinear, want_uv;
    int C;
};

// First half of a pair's gradient: everything that does not need the pair's transmittance.
//   y      = dL/d(vis-weighted value) = colour, normal and texture terms + the distortion weight term
//            (reference texture.cu:583-589, :612-634, :661-668 with the FINAL sums, helpers.cuh:266-269)
//   gu, gv = d(texture term)/du, /dv per unit vis (texture_helpers.cuh:252-300)
template <bool C3>
__device__ __forceinline__ void pair_terms(const RasterCommon &p, const BackwardIn &in, const PairFlags &fl,
                                           const float4 q3, const float4 q4, const float4 q5, const float4 q6,
                                           const float4 q7, const PairEval &pe, const PixelConsts &qc,
                                           const PixelShare &px, float &y, float &gu, float &gv) {
    float v_vis = q6.x * px.vi0 + q6.y * px.vi1 + q6.z * px.vi2 + q7.x * px.vn0 + q7.y * px.vn1 + q7.z * px.vn2;
    const float nu = fmaf(q4.x, pe.ex, fmaf(q4.y, pe.ey, q4.z));
    const float nv = fmaf(q5.x, pe.ex, fmaf(q5.y, pe.ey, q5.z));
    const float u = clamp01(fmaf(nu, pe.rD, q4.w)), v = clamp01(fmaf(nv, pe.rD, q5.w));
    TexFetch tf;
    texel_setup(__float_as_int(q3.z), __float_as_int(q3.w), __float_as_int(q6.w), u, v, fl.bilinear, tf);
    gu = 0.f;
    gv = 0.f;
    if (C3) {
        const float4 t0 = __ldg(p.tex4 + tf.idx[0]), t1 = __ldg(p.tex4 + tf.idx[1]);
        const float4 t2 = __ldg(p.tex4 + tf.idx[2]), t3 = __ldg(p.tex4 + tf.idx[3]);
        const float val0 = tf.w[0] * t0.x + tf.w[1] * t1.x + tf.w[2] * t2.x + tf.w[3] * t3.x;
        const float val1 = tf.w[0] * t0.y + tf.w[1] * t1.y + tf.w[2] * t2.y + tf.w[3] * t3.y;
        const float val2 = tf.w[0] * t0.z + tf.w[1] * t1.z + tf.w[2] * t2.z + tf.w[3] * t3.z;
        v_vis += val0 * px.vt0 + val1 * px.vt1 + val2 * px.vt2;
        if (fl.want_uv) {
            const float ofu = 1.f - tf.fu, ofv = 1.f - tf.fv;
            const float gu0 = -ofv * t0.x - tf.fv * t1.x + ofv * t2.x + tf.fv * t3.x;
            const float gu1 = -ofv * t0.y - tf.fv * t1.y + ofv * t2.y + tf.fv * t3.y;
            const float gu2 = -ofv * t0.z - tf.fv * t1.z + ofv * t2.z + tf.fv * t3.z;
            const float gv0 = -ofu * t0.x + ofu * t1.x - tf.fu * t2.x + tf.fu * t3.x;
            const float gv1 = -ofu * t0.y + ofu * t1.y - tf.fu * t2.y + tf.fu * t3.y;
            const float gv2 = -ofu * t0.z + ofu * t1.z - tf.fu * t2.z + tf.fu * t3.z;
            gu = (float)tf.h * (px.vt0 * gu0 + px.vt1 * gu1 + px.vt2 * gu2);
            gv = (float)tf.wd * (px.vt0 * gv0 + px.vt1 * gv1 + px.vt2 * gv2);
        }
    } else {
        const float *__restrict__ tx = p.tex;
        const float *__restrict__ vtp = in.v_tex + (size_t)fl.C * px.pix;
        for (int c = 0; c < fl.C; ++c) {
            const float c00 = __ldg(tx + (size_t)tf.idx[0] * fl.C + c), c01 = __ldg(tx + (size_t)tf.idx[1] * fl.C + c);
            const float c10 = __ldg(tx + (size_t)tf.idx[2] * fl.C + c), c11 = __ldg(tx + (size_t)tf.idx[3] * fl.C + c);
            const float vtc = vtp[c];
            v_vis += (tf.w[0] * c00 + tf.w[1] * c01 + tf.w[2] * c10 + tf.w[3] * c11) * vtc;
            if (fl.want_uv) {
                gu += (float)tf.h * (vtc * (-(1.f - tf.fv) * c00 - tf.fv * c01 + (1.f - tf.fv) * c10 + tf.fv * c11));
                gv += (float)tf.wd * (vtc * (-(1.f - tf.fu) * c00 + (1.f - tf.fu) * c01 - tf.fu * c10 + tf.fu * c11));
            }
        }
    }
    const float t_view = pe.t * qc.vdep;
    const float tv = fl.use_ndc ? (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view) : pe.t;
    y = v_vis + (tv * tv * px.Sf0 - 2.f * tv * px.Sf1 + px.Sf2) * px.v_reg;
}

// red.global.add.v4.f32 of (x, y, z, 0) unless `w` is exactly zero - a predicated instruction, not a branch: the four
// corner reductions of a pair otherwise compile to four divergent regions of 13 instructions each.
__device__ __forceinline__ void red_add_v4_if_nonzero(float4 *addr, float x, float y, float z, float w) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %4, 0f00000000;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %5};\n\t}" ::"l"(addr),
        "f"(x), "f"(y), "f"(z), "f"(w), "f"(0.f)
        : "memory");
}

// Second half: given vis = alpha T_k and v_alpha, the texel-gradient reductions and the moment row (AccSlot order,
// quad by quad; r[7] only with BLUR).
template <bool C3, bool BLUR>
__device__ __forceinline__ void pair_rows(const RasterCommon &p, const BackwardIn &in, const BackwardOut &o,
                                          const PairFlags &fl, const float4 q0, const float4 q3, const float4 q4,
                                          const float4 q5, const float4 q6, const PairEval &pe, const PixelConsts &qc,
                                          const PixelShare &px, float vis, float v_alpha, float gu, float gv,
                                          bool is_median, float4 (&r)[8]) {
    const float v_u = vis * gu, v_v = vis * gv;
    const float nu = fmaf(q4.x, pe.ex, fmaf(q4.y, pe.ey, q4.z));
    const float nv = fmaf(q5.x, pe.ex, fmaf(q5.y, pe.ey, q5.z));
    const float du = nu * pe.rD, dv = nv * pe.rD;
    {  // texel gradients: one vector reduction per bilinear corner (texture_helpers.cuh:257-260)
        const float u = clamp01(fmaf(nu, pe.rD, q4.w)), v = clamp01(fmaf(nv, pe.rD, q5.w));
        TexFetch tf;
        texel_setup(__float_as_int(q3.z), __float_as_int(q3.w), __float_as_int(q6.w), u, v, fl.bilinear, tf);
        if (C3) {
            const float vv0 = vis * px.vt0, vv1 = vis * px.vt1, vv2 = vis * px.vt2;
#pragma unroll
            for (int k = 0; k < 4; ++k)  // texel indices are non-negative: unsigned scaling is one IMAD.WIDE
                red_add_v4_if_nonzero(o.vtex4 + (unsigned)tf.idx[k], tf.w[k] * vv0, tf.w[k] * vv1, tf.w[k] * vv2, tf.w[k]);
        } else {
            const float *__restrict__ vtp = in.v_tex + (size_t)fl.C * px.pix;
            for (int c = 0; c < fl.C; ++c) {
                const float vv = vis * vtp[c];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (tf.w[k] != 0.f) atomicAdd(o.vtex + (size_t)tf.idx[k] * fl.C + c, tf.w[k] * vv);
            }
        }
    }
    // depth / distortion terms through t (reference texture.cu:655-682)
    const float t_view = pe.t * qc.vdep;
    const float tv = fl.use_ndc ? (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view) : pe.t;
    const float v_tv = 2.f * (vis * tv * px.Sf0 - vis * px.Sf1) * px.v_reg;
    float v_t = fl.use_ndc ? 0.f : v_tv;
    float v_tview = is_median ? px.v_dep : 0.f;  // texture.cu:678-680
    if (fl.use_ndc) v_tview += (T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view * t_view) * v_tv;
    v_t += qc.vdep * v_tview;
    const float v_s = v_t * qc.rn;
    // alpha = min(.99, opac * f): the cap is not masked (reference texture.cu:672, :707)
    const float v_q = pe.blur ? 0.f : -LN2_F * q0.w * pe.e * v_alpha;
    const float v_l1 = 2.f * pe.l1 * v_q, v_l2 = 2.f * pe.l2 * v_q;
    const float gN1 = v_l1 * pe.rD, gN2 = v_l2 * pe.rD, gNu = v_u * pe.rD, gNv = v_v * pe.rD;
    const float gD = -(v_l1 * pe.l1 + v_l2 * pe.l2 + v_s * pe.s + v_u * du + v_v * dv) * pe.rD;
    r[0] = make_float4(gN1 * pe.ex, gN1 * pe.ey, gN1, v_s * pe.rD);
    r[1] = make_float4(gN2 * pe.ex, gN2 * pe.ey, gN2, pe.f * v_alpha);
    r[2] = make_float4(gD * pe.ex, gD * pe.ey, gD, 0.f);
    r[3] = make_float4(gNu * pe.ex, gNu * pe.ey, gNu, v_u);
    r[4] = make_float4(gNv * pe.ex, gNv * pe.ey, gNv, v_v);
    r[5] = make_float4(vis * px.vi0, vis * px.vi1, vis * px.vi2, 0.f);
    r[6] = make_float4(vis * px.vn0, vis * px.vn1, vis * px.vn2, 0.f);
    r[7] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BLUR) {
        if (pe.blur) {  // reference texture.cu:683-692
            const float v_sb = -q0.w * pe.f * v_alpha;
            r[7].x = 2.0f * v_sb * pe.bx;
            r[7].y = 2.0f * v_sb * pe.by;
        }
    }
}

// Sum 8 float4 quads (32 slots) held per lane over the warp by recursive halving: at the step with partner lane ^ b
// a lane keeps the upper half of its slots if (lane & b) and sends the other half, so after 16 + 8 + 4 + 2 + 1 = 31
// shuffles lane L holds the warp total of slot L.
__device__ __forceinline__ float warp_reduce_slots(const float4 (&r)[8], int lane) {
    const unsigned full = 0xffffffffu;
    float a[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[4 * k] = r[k].x; a[4 * k + 1] = r[k].y; a[4 * k + 2] = r[k].z; a[4 * k + 3] = r[k].w;
    }
#pragma unroll
    for (int b = 16; b >= 1; b >>= 1) {
        const bool hi = (lane & b) != 0;
#pragma unroll
        for (int i = 0; i < b; ++i) {
            const float send = hi ? a[i] : a[i + b];
            const float keep = hi ? a[i + b] : a[i];
            a[i] = keep + __shfl_xor_sync(full, send, b);
        }
    }
    return a[0];
}

template <bool C3, bool BLUR>
#ifndef GSTEX_BWD_MIN_CTAS
#define GSTEX_BWD_MIN_CTAS (GSTEX_BWD_MINB * (BWD_WARPS / BWD_CTA_WARPS))
#endif
__global__ void __launch_bounds__(BWD_CTA_WARPS * 32, GSTEX_BWD_MIN_CTAS) raster_backward_kernel(const RasterCommon p, const BackwardIn in,
                                                                               const BackwardOut o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using WS = BwdWarpSmem<BLUR>;
    constexpr int PITCH = WS::PITCH;
    const unsigned full = 0xffffffffu;
    // A tile's warps never synchronise with each other, so a tile may be spread over several CTAs of BWD_CTA_WARPS warps:
    // a finished warp's slot is then refilled without waiting for the slowest warp of its tile.
    const int groups = ((p.nthreads >> 5) + BWD_CTA_WARPS - 1) / BWD_CTA_WARPS;  // CTAs per tile
    const int tile_x = BWD_CTA_WARPS == BWD_WARPS ? (int)blockIdx.x : (int)blockIdx.x / groups;
    const int tr = BWD_CTA_WARPS == BWD_WARPS ? (int)threadIdx.x
                                              : ((int)blockIdx.x - tile_x * groups) * (BWD_CTA_WARPS * 32) + (int)threadIdx.x;
    if (tr >= p.nthreads) return;
    const int warp = tr >> 5;  // warp of the TILE: pixel patch and mask word
    // Under the 80-register cap ptxas re-derives `lane` and this warp's shared-memory base from S2R SR_TID.X (a
    // ~20-cycle special-register read at the head of dependent address chains) inside every hot loop.  Passing both
    // through an empty asm makes them opaque, so they are held in registers instead: 2.60 -> 2.51 ms on C4.
    int lane = tr & 31;
    asm volatile("" : "+r"(lane));
    const unsigned lt = (1u << lane) - 1u;
    unsigned wbase = (threadIdx.x >> 5) * (unsigned)sizeof(WS);
    asm volatile("" : "+r"(wbase));
    WS &W = *reinterpret_cast<WS *>(smem_raw + wbase);

    const int tile = blockIdx.y * p.tiles_x + tile_x;
    int lx, ly;
    tile_pixel(p.bw, tr, lx, ly);
    const int col = tile_x * p.bw + lx, row = blockIdx.y * p.bw + ly;
    const bool inside = (tr < p.bw * p.bw) && col < p.img_w && row < p.img_h;
    const PixelConsts pc = make_pixel(col, row, p.c2w, p.viewmat, p.fx, p.fy, p.cx, p.cy);
    PairFlags fl;
    fl.use_ndc = (p.settings & GSTEX_SET_NDC) != 0;
    fl.bilinear = !(p.settings & GSTEX_SET_NEAREST);
    fl.want_uv = fl.bilinear && (p.settings & GSTEX_SET_PROPAGATE_UV) != 0;
    fl.C = C3 ? 3 : p.channels;
    constexpr int NCOL2 = BLUR ? 15 : 14;  // float2 column pairs of a moment row

    PixelShare me;
    me.pix = inside ? row * p.img_w + col : 0;
    const int2 range = p.bins[tile];
    const int bfinal = inside ? in.final_idx[me.pix] : -1;
    // every warp walks the tile's list on its own, from the last entry one of ITS pixels blended back to the front:
    // no block-wide staging, no barriers
    const int warp_last = __reduce_max_sync(full, bfinal);
    const int hi = min(range.y, warp_last + 1);
    if (hi <= range.x) return;
    const int nbatch = (hi - range.x + BWD_BATCH - 1) / BWD_BATCH;

    float T = in.final_Ts[me.pix];
    me.Sf0 = in.final_s[3 * me.pix]; me.Sf1 = in.final_s[3 * me.pix + 1]; me.Sf2 = in.final_s[3 * me.pix + 2];
    me.dfinal = in.depth_idx[me.pix];
    // an upstream gradient the loss does not produce may be passed as NULL (= zeros), except the generic-channel v_tex
    me.vi0 = me.vi1 = me.vi2 = me.vn0 = me.vn1 = me.vn2 = me.vt0 = me.vt1 = me.vt2 = 0.f;
    if (in.v_img) {
        me.vi0 = in.v_img[3 * me.pix]; me.vi1 = in.v_img[3 * me.pix + 1]; me.vi2 = in.v_img[3 * me.pix + 2];
    }
    if (in.v_normal) {
        me.vn0 = in.v_normal[3 * me.pix]; me.vn1 = in.v_normal[3 * me.pix + 1]; me.vn2 = in.v_normal[3 * me.pix + 2];
    }
    me.v_dep = in.v_depth ? in.v_depth[me.pix] : 0.f;
    me.v_reg = in.v_reg ? in.v_reg[me.pix] : 0.f;
    if (C3 && in.v_tex) {
        me.vt0 = in.v_tex[3 * me.pix]; me.vt1 = in.v_tex[3 * me.pix + 1]; me.vt2 = in.v_tex[3 * me.pix + 2];
    }
    float v_T_run = p.background[0] * me.vi0 + p.background[1] * me.vi1 + p.background[2] * me.vi2 -
                    (in.v_alpha ? in.v_alpha[me.pix] : 0.f);

    // values of pixel lane `pl`, fetched by a pair lane (all 32 lanes must call this together)
    auto fetch_pixel = [&](int pl, PixelConsts &qc, PixelShare &px) {
        qc.px = __shfl_sync(full, pc.px, pl); qc.py = __shfl_sync(full, pc.py, pl);
        qc.rn = __shfl_sync(full, pc.rn, pl); qc.vdep = __shfl_sync(full, pc.vdep, pl);
        qc.eps = __shfl_sync(full, pc.eps, pl);
        px.vi0 = __shfl_sync(full, me.vi0, pl); px.vi1 = __shfl_sync(full, me.vi1, pl); px.vi2 = __shfl_sync(full, me.vi2, pl);
        px.vn0 = __shfl_sync(full, me.vn0, pl); px.vn1 = __shfl_sync(full, me.vn1, pl); px.vn2 = __shfl_sync(full, me.vn2, pl);
        px.vt0 = __shfl_sync(full, me.vt0, pl); px.vt1 = __shfl_sync(full, me.vt1, pl); px.vt2 = __shfl_sync(full, me.vt2, pl);
        px.Sf0 = __shfl_sync(full, me.Sf0, pl); px.Sf1 = __shfl_sync(full, me.Sf1, pl); px.Sf2 = __shfl_sync(full, me.Sf2, pl);
        px.v_reg = __shfl_sync(full, me.v_reg, pl); px.v_dep = __shfl_sync(full, me.v_dep, pl);
        px.dfinal = __shfl_sync(full, me.dfinal, pl);
        px.pix = C3 ? 0 : __shfl_sync(full, me.pix, pl);
    };

    static_assert(BWD_BATCH == 64, "the mask prefetch below holds one batch as two words per lane");
    // mask words and ids of batch b+1 are fetched while batch b is processed (two entries per lane)
    uint32_t nx_m0 = 0u, nx_m1 = 0u;
    int32_t nx_g0 = 0, nx_g1 = 0;
    auto fetch_batch = [&](int b) {
        const int first = max(range.x, hi - (b + 1) * BWD_BATCH), cnt = (hi - b * BWD_BATCH) - first;
        const int r0 = lane, r1 = lane + 32;
        nx_m0 = r0 < cnt ? __ldg(p.masks + (size_t)(first + r0) * MASK_WARPS + warp) : 0u;
        nx_g0 = r0 < cnt ? __ldg(p.ids + first + r0) : 0;
        nx_m1 = r1 < cnt ? __ldg(p.masks + (size_t)(first + r1) * MASK_WARPS + warp) : 0u;
        nx_g1 = r1 < cnt ? __ldg(p.ids + first + r1) : 0;
    };
    fetch_batch(0);
    for (int b = 0; b < nbatch; ++b) {
        // batch b covers [first, first + cnt) counted from the back of [range.x, hi)
        const int first = max(range.x, hi - (b + 1) * BWD_BATCH), cnt = (hi - b * BWD_BATCH) - first;
        // this warp's non-empty mask words of the batch (and the Gaussian ids), compacted in list order
        int nsurv = 0;
        const uint32_t cm0 = nx_m0, cm1 = nx_m1;
        const int32_t cg0 = nx_g0, cg1 = nx_g1;
        if (b + 1 < nbatch) fetch_batch(b + 1);
        for (int k = 0; k < cnt; k += 32) {
            const int r = k + lane;
            const uint32_t mw = k ? cm1 : cm0;
            const int32_t g = k ? cg1 : cg0;
            const unsigned m = __ballot_sync(full, mw != 0u);
            if (mw != 0u) {
                const int pos = nsurv + __popc(m & lt);
                W.sv_r[pos] = (uint8_t)r;
                W.sv_mask[pos] = mw;
#ifndef GSTEX_BWD_NO_SVGID
                W.sv_gid[pos] = g;
#endif
            }
            nsurv += __popc(m);
        }
        __syncwarp();

        int si = nsurv - 1;
        while (si >= 0) {
            // ---------------- build one chunk: entries si0, si0-1, ... while they fit ----------------
            // lane j looks at entry si0 - j.  Entries that cover >= BWD_FULL pixels of the patch are handled
            // lane = pixel inside the scan below; the pairs of the sparser ones are queued for the dense stages.
            // (A pixel takes part in an entry iff its bit is set in the forward pass's mask.)
            const int si0 = si;
            int ne, np;
            {
                const bool cand = lane < BWD_ECAP && si0 - lane >= 0;
                const unsigned m = cand ? W.sv_mask[si0 - lane] : 0u;
                const int c = __popc(m);
                const bool sparse = c < BWD_FULL;
                const int cs = sparse ? c : 0;
                int incl = cs;  // inclusive prefix sum of the queued pairs over the entries
#pragma unroll
                for (int o = 1; o < BWD_ECAP; o <<= 1) {
                    const int t = __shfl_up_sync(full, incl, o);
                    if (lane >= o) incl += t;
                }
                ne = __popc(__ballot_sync(full, cand && incl <= BWD_QCAP));  // a prefix: incl is non-decreasing
                np = __shfl_sync(full, incl, ne - 1);
                const int off = incl - cs;
                if (lane < ne) W.ch_off[lane] = (uint16_t)off;
                // stage the records of the chunk's entries (8 lanes fetch one 128-byte record); the copies fly while
                // the pair queue is being written
                for (int t = lane; t < ne * 8; t += 32) {
                    const int j = t >> 3, q = t & 7;
#ifndef GSTEX_BWD_NO_SVGID
                    const int32_t gid = W.sv_gid[si0 - j];
#else
                    const int32_t gid = __ldg(p.ids + first + (int)W.sv_r[si0 - j]);
#endif
                    __pipeline_memcpy_async(W.rec + quad_slot(j, q), p.recs + (size_t)gid * 8 + q, 16);
                }
                __pipeline_commit();
                unsigned mm = (lane < ne && sparse) ? m : 0u;
                int slot = off;
                while (__any_sync(full, mm != 0u)) {
                    if (mm) {
                        const int l = __ffs(mm) - 1;
                        mm &= mm - 1;
                        W.q_ent[slot++] = (uint16_t)(lane | (l << 8));
                    }
                }
                si = si0 - ne;
                __pipeline_wait_prior(0);
            }
            __syncwarp();
            const int niter = (np + 31) >> 5;

            // ---------------- D1: lane = pair ----------------
            for (int it = 0; it < niter; ++it) {
                const int e = it * 32 + lane;
                const bool act = e < np;
                const int ent = act ? (int)W.q_ent[e] : (lane << 8);
                PixelConsts qc;
                PixelShare px;
                fetch_pixel(ent >> 8, qc, px);
                if (act) {
                    const int j = ent & 0xff;
                    const float4 *__restrict__ R = W.rec + j * REC_PITCH;
                    const float4 q0 = R[0], q1 = R[1], q2 = R[2], q3 = R[3];
                    const float4 q4 = R[4], q5 = R[5], q6 = R[6], q7 = R[7];
                    PairEval pe;
                    eval_pair<BLUR>(q0, q1, q2, q3, qc, p.mean2d, pe);
                    float y, gu, gv;
                    pair_terms<C3>(p, in, fl, q3, q4, q5, q6, q7, pe, qc, px, y, gu, gv);
                    W.e_ay[e] = make_float2(pe.alpha, y);
                    W.e_g[e] = make_float2(gu, gv);
                }
            }
            __syncwarp();

            // ---------------- scan: lane = pixel, entries in list order (back to front) ----------------
            {
                for (int jj = 0; jj < ne; ++jj) {
                    const int s = si0 - jj;
                    const unsigned m = W.sv_mask[s];
                    const int c = __popc(m);
                    const bool mine = (m >> lane) & 1u;
                    if (c >= BWD_FULL) {
                        // a Gaussian that covers most of the patch: the whole gradient path right here, lane = pixel,
                        // record read by broadcast, moments reduced with the shuffle butterfly
                        const float4 *__restrict__ R = W.rec + jj * REC_PITCH;
                        const float4 q0 = R[0], q1 = R[1], q2 = R[2], q3 = R[3];
                        float4 r[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (mine) {
                            const float4 q4 = R[4], q5 = R[5], q6 = R[6], q7 = R[7];
                            PairEval pe;
                            eval_pair<BLUR>(q0, q1, q2, q3, pc, p.mean2d, pe);
                            float y, gu, gv;
                            pair_terms<C3>(p, in, fl, q3, q4, q5, q6, q7, pe, pc, me, y, gu, gv);
                            const float alpha = pe.alpha;
                            T *= fast_rcp(1.f - alpha);
                            const float d = y - v_T_run;
                            const float v_alpha = T * d;
                            v_T_run = fmaf(alpha, d, v_T_run);
                            const int idx = first + (int)W.sv_r[s];
                            pair_rows<C3, BLUR>(p, in, o, fl, q0, q3, q4, q5, q6, pe, pc, me, alpha * T, v_alpha, gu, gv,
                                                idx == me.dfinal && me.dfinal != -1, r);
                        }
                        const float tot = warp_reduce_slots(r, lane);
                        if (lane < (BLUR ? 30 : 28))  // one 4-byte reduction per lane, 112 contiguous bytes of the moment line
                            atomicAdd(reinterpret_cast<float *>(o.acc) + (size_t)__float_as_int(q2.w) * 32 + lane, tot);
                    } else {
                        if (mine) {
                            const int e = (int)W.ch_off[jj] + __popc(m & lt);
                            const float2 ay = W.e_ay[e];
                            const float alpha = ay.x, d = ay.y - v_T_run;
                            T *= fast_rcp(1.f - alpha);  // transmittance in front of the entry (texture.cu:579-580)
                            W.e_ay[e] = make_float2(alpha * T, T * d);  // (vis, v_alpha) (texture.cu:650, :667)
                            v_T_run = fmaf(alpha, d, v_T_run);       // alpha y + (1 - alpha) R  (texture.cu:651, :668-670)
                        }
                    }
                }
            }
            __syncwarp();

            // ---------------- D2: lane = pair ----------------
            for (int it = 0; it < niter; ++it) {
                const int e = it * 32 + lane;
                const bool act = e < np;
                const int ent = act ? (int)W.q_ent[e] : (lane << 8);
                PixelConsts qc;
                PixelShare px;
                fetch_pixel(ent >> 8, qc, px);
                const int jloc = act ? (ent & 0xff) : -1 - lane;
                if (act) {
                    const float4 *__restrict__ R = W.rec + jloc * REC_PITCH;
                    const float4 q0 = R[0], q1 = R[1], q2 = R[2], q3 = R[3];
                    const float4 q4 = R[4], q5 = R[5], q6 = R[6];
                    PairEval pe;
                    eval_pair<BLUR>(q0, q1, q2, q3, qc, p.mean2d, pe);
                    const int idx = first + (int)W.sv_r[si0 - jloc];
                    float4 r[8];
                    const float2 va = W.e_ay[e], g2 = W.e_g[e];
                    pair_rows<C3, BLUR>(p, in, o, fl, q0, q3, q4, q5, q6, pe, qc, px, va.x, va.y, g2.x, g2.y,
                                        idx == px.dfinal && px.dfinal != -1, r);
                    float4 *__restrict__ rowp = reinterpret_cast<float4 *>(W.rows + lane * PITCH);
#pragma unroll
                    for (int k = 0; k < (BLUR ? 8 : 7); ++k) rowp[k] = r[k];
                    W.row_gid[lane] = __float_as_int(q2.w);
                }
                // rows of one Gaussian are contiguous: segment starts where the chunk-local entry changes
                const int jprev = __shfl_up_sync(full, jloc, 1);
                const unsigned bnd = __ballot_sync(full, lane == 0 || jloc != jprev);
                __syncwarp();
                const int nact = min(32, np - it * 32);
                if (lane < 2 * NCOL2) {  // 2 x NCOL2 lanes: column pair (lane % NCOL2), rows [16 half, 16 half + 16)
                    const int half = lane >= NCOL2 ? 1 : 0, cp = lane - half * NCOL2;
                    const int r0 = half << 4;
                    const float *__restrict__ colp = W.rows + 2 * cp;
                    float2 acc2 = make_float2(0.f, 0.f);
                    for (int t = 0; t < 16; ++t) {
                        const int r = r0 + t;
                        if (r < nact) {
                            if (t > 0 && ((bnd >> r) & 1u)) {
                                atomicAdd(reinterpret_cast<float2 *>(o.acc) + (size_t)W.row_gid[r - 1] * 16 + cp, acc2);
                                acc2 = make_float2(0.f, 0.f);
                            }
                            const float2 v2 = *reinterpret_cast<const float2 *>(colp + r * PITCH);
                            acc2.x += v2.x;
                            acc2.y += v2.y;
                        }
                    }
                    const int rl = min(nact, r0 + 16) - 1;
                    if (rl >= r0) atomicAdd(reinterpret_cast<float2 *>(o.acc) + (size_t)W.row_gid[rl] * 16 + cp, acc2);
                }
                __syncwarp();
            }
        }
    }
}

template <bool C3, bool BLUR>
static int launch_bwd_variant(const dim3 grid, const RasterCommon &p, const BackwardIn &in, const BackwardOut &o,
                              cudaStream_t s) {
    const size_t smem = sizeof(BwdWarpSmem<BLUR>) * BWD_CTA_WARPS;
    static SmemOnceFlags once;  // one per template instantiation
    const int rc = configure_dynamic_smem((const void *)raster_backward_kernel<C3, BLUR>, smem, true, once);
    if (rc != GSTEX_OK) return rc;
    const int cta_threads = min(p.nthreads, BWD_CTA_WARPS * 32);
    const dim3 g(grid.x * ceil_div(p.nthreads, BWD_CTA_WARPS * 32), grid.y);
    raster_backward_kernel<C3, BLUR><<<g, cta_threads, smem, s>>>(p, in, o);
    return GSTEX_OK;
}

int launch_raster_backward(const RasterCommon &p, const BackwardIn &in, const BackwardOut &o, cudaStream_t s) {
    const dim3 grid(p.tiles_x, ceil_div(p.img_h, p.bw));
    const bool blur = (p.settings & GSTEX_SET_BLUR) != 0;
    int rc;
    if (p.channels == 3) rc = blur ? launch_bwd_variant<true, true>(grid, p, in, o, s) : launch_bwd_variant<true, false>(grid, p, in, o, s);
    else rc = blur ? launch_bwd_variant<false, true>(grid, p, in, o, s) : launch_bwd_variant<false, false>(grid, p, in, o, s);
    if (rc != GSTEX_OK) return rc;
    GSTEX_LAUNCH_OK("raster_backward_kernel");
    return GSTEX_OK;
}

struct BwdLayout {
    size_t acc_off, vtex4_off, fwd_off, total;
};

// `own_forward_state`: the call has no forward scratch to read and rebuilds records / padded texture / masks itself
static BwdLayout backward_layout(int n, int64_t num_texels, int channels, int64_t num_intersects, bool own_forward_state) {
    BwdLayout L;
    size_t off = 0;
    L.acc_off = off;
    off = align_up(off + sizeof(float) * ACC_FLOATS * (size_t)(n > 0 ? n : 1), 256);
    L.vtex4_off = off;
    if (channels == 3) off = align_up(off + sizeof(float4) * (size_t)(num_texels > 0 ? num_texels : 1), 256);
    L.fwd_off = off;
    if (own_forward_state) off += forward_layout(n, num_texels, channels, num_intersects).total;
    L.total = off;
    return L;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_texture_backward_temp_bytes(int n, int64_t num_texels, int channels) {
    return backward_layout(n, num_texels, channels, 0, false).total;
}

extern "C" size_t gstex_texture_backward_stateless_temp_bytes(int n, int64_t num_texels, int channels,
                                                              int64_t num_intersects) {
    return backward_layout(n, num_texels, channels, num_intersects, true).total;
}

extern "C" int gstex_texture_backward(
    int img_height, int img_width, int block_width, int n, int64_t num_texels, int channels, int64_t num_intersects,
    const int32_t *texture_dims, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
    const float *opacities, const float *means, const float *scales, float glob_scale, const float *quats,
    const float *uv0, const float *umap, const float *vmap, const float *texture, const float *viewmat,
    const float *c2w, float fx, float fy, float cx, float cy, int settings, const float *background,
    const float *final_Ts, const int32_t *final_idx, const int32_t *depth_idx, const float *final_s,
    const float *v_out_img, const float *v_out_depth, const float *v_out_reg, const float *v_out_alpha,
    const float *v_out_texture, const float *v_out_normal, float *v_colors, float *v_opacity, float *v_means,
    float *v_scales, float *v_quats, float *v_uv0, float *v_umap, float *v_vmap, float *v_texture, int accumulate,
    const void *fwd_temp, void *temp, size_t temp_bytes, gstex_stream_t stream) {
    GSTEX_REQUIRE(num_intersects >= 0 && num_intersects < ((int64_t)1 << 31), GSTEX_E_INVALID,
                  "texture_backward: num_intersects = %lld", (long long)num_intersects);
    int rc = check_raster_args("texture_backward", img_height, img_width, block_width, n, num_texels, channels, settings);
    if (rc != GSTEX_OK) return rc;
    const bool stateless = fwd_temp == nullptr;
    const FwdLayout FL = forward_layout(n, num_texels, channels, num_intersects);
    const BwdLayout L = backward_layout(n, num_texels, channels, num_intersects, stateless);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE,
                  "texture_backward: temp too small (%zu < %zu; without fwd_temp the call needs "
                  "gstex_texture_backward_stateless_temp_bytes)", temp_bytes, L.total);
    cudaStream_t s = as_stream(stream);
    char *base = (char *)temp;
    float4 *acc = (float4 *)(base + L.acc_off);
    float4 *vtex4 = (float4 *)(base + L.vtex4_off);
    GSTEX_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(float) * ACC_FLOATS * (size_t)n, s));
    if (channels == 3) {
        GSTEX_CUDA_OK(cudaMemsetAsync(vtex4, 0, sizeof(float4) * (size_t)num_texels, s));
    } else if (!accumulate) {
        GSTEX_CUDA_OK(cudaMemsetAsync(v_texture, 0, sizeof(float) * (size_t)channels * (size_t)num_texels, s));
    }

    // Forward state: the caller's scratch when it kept one, else rebuilt here from the call's own arguments - the
    // reference's texture_backward_tensor is a pure function of them (texture.cu:915-1053): records and padded texture
    // are re-packed, and the blend masks are re-derived from final_Ts / final_idx by raster_masks_kernel (no second
    // forward pass: no compositing, no texel fetch, no outputs).
    const char *fbase = stateless ? base + L.fwd_off : (const char *)fwd_temp;
    RasterCommon p = make_raster_common(
        img_height, img_width, block_width, channels, settings, gaussian_ids_sorted, tile_bins,
        (const float4 *)(fbase + FL.recs_off), (const float2 *)(fbase + FL.mean2d_off),
        (const float4 *)(fbase + FL.tex4_off), texture, viewmat, c2w, background, fx, fy, cx, cy,
        (uint32_t *)const_cast<char *>(fbase + FL.masks_off));
    if (stateless && num_intersects > 0) {
        rc = launch_pack(n, means, scales, glob_scale, quats, opacities, colors, uv0, umap, vmap, texture_dims, viewmat,
                         c2w, fx, fy, cx, cy, (float4 *)const_cast<char *>(fbase + FL.recs_off),
                         (float2 *)const_cast<char *>(fbase + FL.mean2d_off), s);
        if (rc != GSTEX_OK) return rc;
        if (channels == 3) {
            rc = launch_pad_texture(num_texels, texture, (float4 *)const_cast<char *>(fbase + FL.tex4_off), s);
            if (rc != GSTEX_OK) return rc;
        }
        rc = launch_raster_masks(p, final_Ts, final_idx, num_intersects, nullptr, s);
        if (rc != GSTEX_OK) return rc;
    }
    BackwardIn in{final_Ts, final_s, final_idx, depth_idx, v_out_img, v_out_depth, v_out_reg, v_out_alpha, v_out_texture,
                  v_out_normal};
    BackwardOut o{acc, vtex4, v_texture};
    if (num_intersects > 0) {
        rc = launch_raster_backward(p, in, o, s);
        if (rc != GSTEX_OK) return rc;
    }
    rc = launch_epilogue(n, means, scales, glob_scale, quats, umap, vmap, viewmat, c2w, fx, fy, cx, cy, acc, v_colors,
                         v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap, accumulate, s, p.recs);
    if (rc != GSTEX_OK) return rc;
    if (channels == 3) {
        rc = launch_unpad_texture_grad(num_texels, vtex4, v_texture, accumulate, s);
        if (rc != GSTEX_OK) return rc;
    }
    return GSTEX_OK;
}
