// raster_forward.cu -- rasterise forward (reference texture.cu:11-329).
//
// One CTA per screen tile, one thread per pixel; for 16x16 tiles each warp owns an 8x4 pixel patch.
// The tile's depth-sorted Gaussian list is streamed through shared memory in 128-record stages with
// cp.async (LDGSTS): while the CTA composites stage b, the 18 KB of stage b+1 are in flight.  Three
// stage buffers (dynamic shared memory) make ONE block barrier per stage enough.  Each record is one 128-byte line gathered through gaussian_ids_sorted; all lanes read the
// same record at the same time, so the shared-memory reads are broadcasts (LDS.128, no conflicts).
// The texture is read as one aligned float4 per corner from the padded copy (4 LDG.128 per blended
// pair instead of 12 scalar loads).
//
// Behavioural quirks reproduced (SURVEY 8a "quirks" 1-12): alpha cap 0.99, skip test t<0.01 | t>1000 |
// alpha<1/255, termination T(1-alpha) <= 1e-4 tested before the skip is honoured, median depth while
// T>0.5, distortion prefix sums, out_texture without background, final_idx = absolute list index.
#include "raster.cuh"

namespace gstex {

// C3 = true : 3-channel texture read through the padded float4 copy, accumulators in registers.
// C3 = false: runtime channel count (<= 64) read from the caller's (X,C) array (slow generic path).
// Visualisation helper (VIS builds only): distance, in pixels, from this pixel to the nearest pixel of its 7x7
// neighbourhood that lies outside the Gaussian's hard footprint sigma <= sigma_thresh (reference local_outline,
// texture_helpers.cuh:390-416).  The neighbours' sigma comes from the same affine forms, evaluated at the shifted
// offset; the plane-denominator clamp uses this pixel's ray norm (the neighbours' differ in the 4th digit of 1e-6).
__device__ __forceinline__ float outline_distance(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                                  const PixelConsts &pc, float sigma_thresh) {
    float min_dis = 1000.0f;
    for (int dx = -3; dx <= 3; ++dx)
        for (int dy = -3; dy <= 3; ++dy) {
            const float ex = __fsub_rn(pc.px + (float)dx, q0.x), ey = __fsub_rn(pc.py + (float)dy, q0.y);
            const float n1 = fmaf(q1.x, ex, fmaf(q1.y, ey, q1.z));
            const float n2 = fmaf(q2.x, ex, fmaf(q2.y, ey, q2.z));
            float d = fmaf(q3.x, ex, fmaf(q3.y, ey, q1.w));
            if (fabsf(d) < pc.eps) d = copysignf(pc.eps, d);
            const float rd = __fdiv_rn(1.f, d);
            const float l1 = n1 * rd, l2 = n2 * rd;
            if (LN2_F * fmaf(l1, l1, l2 * l2) > sigma_thresh) min_dis = fminf(min_dis, sqrtf((float)(dx * dx + dy * dy)));
        }
    return min_dis;
}

// ------------------------------------------------------------------------------------------------------------------
// raster_forward_kernel<C3, BLUR, VIS>  the forward pass.  The warp walks the survivors of a stage in lock step, one
//   record per iteration, every lane evaluating ITS pixel from the broadcast record: the record reads are broadcasts, the
//   four texel loads of the blending lanes fall into the two cache lines of one Gaussian's texture block, and the blend
//   decision of the 32 pixels is one ballot - the mask word the backward pass reads.
//   C3 = false: runtime channel count (<= 64, texels read from the caller's (X,C) array).  VIS = true: the viewer-only
//   settings bits 15-29 (texture.cu:58-63, :201-241, :269-274; no culling - its bound assumes alpha = opac*exp(-sigma) -,
//   no masks, forward only).
// raster_masks_kernel<BLUR>  re-derives the blend masks of a FINISHED forward pass from its saved state, for callers of
//   the stateless reference-shaped backward (texture_backward_tensor, texture.cu:915-1053, is a pure function of its
//   arguments): pixel p composited list entry i iff i <= final_idx[p] and the pair passes the skip test - exactly what the
//   reference's backward re-evaluates (texture.cu:484-558); the stop rule needs no replay because final_idx bounds the
//   walk.  Same staging, culling and pair evaluation as the forward pass, no compositing, no texel fetch, no outputs.
//
// Measured and rejected this round (experiments/README.md): composing each pixel's candidates from a per-pixel queue
// (lanes working on different Gaussians at once, 62-69 % of the lanes busy in the blend instead of 36 %).  The
// instruction count came out the same, but lanes on different Gaussians touch 21 cache lines per texel load instead of 2
// and read their records without broadcast; the L1 tag / LSU pipes went from 56 % to 80 % busy and the kernel from 1.31
// to 1.41-1.61 ms.
//
// The reference tests the stop rule T(1-alpha) <= 1e-4 on EVERY Gaussian, including those it skips for alpha < 1/255
// (texture.cu:213-222); a Gaussian removed by the warp-level cull (alpha < 0.0039 on the whole patch) is never evaluated
// here.  That cannot change any output: if such a Gaussian (alpha_s < 1/255) trips the rule at transmittance T, then for
// the next Gaussian that could blend (alpha_c >= 1/255 > alpha_s) fl(T * fl(1-alpha_c)) <= fl(T * fl(1-alpha_s)) <= 1e-4 by
// monotonicity of rounding, so it trips the rule too: nothing is composited after the point where the reference stopped,
// and T, final_idx and every sum are identical.
// ------------------------------------------------------------------------------------------------------------------
#ifndef GSTEX_FWD_BATCH
#define GSTEX_FWD_BATCH 128
#endif
#ifndef GSTEX_FWD_MINB
#define GSTEX_FWD_MINB 4
#endif
constexpr int FWD_BATCH = GSTEX_FWD_BATCH;  // records per shared-memory stage (<= 256: uint8 indices)
constexpr int FWD_STAGES = 3;
constexpr int FWD_WARPS = RASTER_MAX_THREADS / 32;

// dynamic shared memory: the stage buffers; the per-warp survivor lists are static (their address is a compile-time
// constant: no shared-window base to re-derive inside the survivor walk)
constexpr size_t fwd_smem_bytes() { return sizeof(float4) * FWD_STAGES * FWD_BATCH * REC_PITCH; }
static_assert(fwd_smem_bytes() + (size_t)FWD_WARPS * FWD_BATCH + 1024 <= 227 * 1024 / GSTEX_FWD_MINB, "forward stage buffers exceed the per-CTA shared memory budget");
static_assert(FWD_BATCH % 4 == 0 && FWD_BATCH <= 256, "stage size");

// Shared prologue of one stage iteration: issue stage b+1, wait for stage b, CTA barrier.  Returns false when every
// pixel of the CTA is finished.  Three stage buffers: stage b+1 is filled into the buffer last read two iterations ago,
// which every warp has left by the time it passed this iteration's barrier, so ONE barrier per stage suffices.
__device__ __forceinline__ bool fwd_stage_advance(float4 (*stage)[FWD_BATCH * REC_PITCH], const RasterCommon &p, int2 range,
                                                  int b, int nbatch, int tr, bool done) {
    const int first = range.x + b * FWD_BATCH;
    if (b + 1 < nbatch) {
        stage_records(stage[(b + 1) % FWD_STAGES], p.recs, p.ids, first + FWD_BATCH,
                      min(FWD_BATCH, range.y - first - FWD_BATCH), tr, p.nthreads);
        // the ids of stage b+2 start travelling now: the gather of the next iteration begins with a dependent load of them
        if (tr < (FWD_BATCH + 31) / 32 && first + 2 * FWD_BATCH + 32 * tr < range.y)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.ids + first + 2 * FWD_BATCH + 32 * tr));
        __pipeline_wait_prior(1);
    } else {
        __pipeline_wait_prior(0);
    }
    // stage b is visible to the whole CTA after this barrier
    return __syncthreads_count(done) < p.nthreads;
}


// ------------------------------------------------------------------------------------------------------------------
// Alternative stage pipeline of the forward kernel (GSTEX_FWD_PIPE = 1, NOT the default): bulk asynchronous copies +
// mbarriers instead of cp.async + one CTA barrier per stage.  A record is 128 contiguous bytes gathered through
// gaussian_ids_sorted, so a stage is FWD_BATCH independent `cp.async.bulk.shared.global` copies (one per record, issued
// by the lanes of warp 0) that complete on the stage's FULL mbarrier (expect-tx byte counting); a warp that finished a
// stage arrives on its EMPTY mbarrier, and warp 0 refills a slot - one stage ahead, as before - once every warp has
// released the stage that held it TWO iterations ago.  Warps then wait only for data, never for each other's progress
// through the current stage (the per-stage CTA barrier is 10.7 % of the default build's stall samples).
// Measured on C4 (B200, same box, parity tests green on both): 1.44 ms against 1.31 ms for the default.  The barrier
// stall disappears, but the waiting does not - within a tile the slow warp is the same one stage after stage, so one or
// two stages of slack buy nothing - and waiting on an mbarrier costs instructions (try_wait loop: +80 M warp
// instructions, 1.20 G against 1.12 G) where waiting at a hardware barrier costs none.  Kept selectable for that record.
#ifndef GSTEX_FWD_PIPE
#define GSTEX_FWD_PIPE 0
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Whole-warp wait for the phase, abandoned when `*flag` reaches `target` (every warp of the CTA finished: nothing more
// will arrive).  The verdict is made warp-uniform: if any lane gave up, all do.
__device__ __forceinline__ bool mbar_wait_unless(uint64_t *bar, uint32_t parity, const volatile int *flag, int target) {
    bool ok = true;
    while (!mbar_try_wait(bar, parity))
        if (*flag >= target) {
            ok = false;
            break;
        }
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ bool flag_reached(const volatile int *flag, int target) {  // warp-uniform read
    return __shfl_sync(0xffffffffu, *flag, 0) >= target;
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Issued by ONE warp (warp 0 stages for the whole CTA, so that "is there a next stage" is one warp's decision and no
// thread that left early can leave a phase incomplete): its lanes copy the stage's records, lane 0 posts the byte count.
__device__ __forceinline__ void issue_stage_bulk(float4 *__restrict__ dst, const float4 *__restrict__ recs,
                                                 const int32_t *__restrict__ ids, int first, int cnt, int lane,
                                                 uint64_t *full) {
    for (int r = lane; r < cnt; r += 32)
        bulk_copy_g2s(dst + quad_slot(r, 0), recs + (size_t)ids[first + r] * 8, sizeof(float) * REC_FLOATS, full);
    if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)cnt * (uint32_t)(sizeof(float) * REC_FLOATS));
}

template <bool BLUR>
__global__ void __launch_bounds__(RASTER_MAX_THREADS, GSTEX_FWD_MINB) raster_masks_kernel(const RasterCommon p,
                                                                                          const float *__restrict__ final_Ts,
                                                                                          const int32_t *__restrict__ final_idx) {
    extern __shared__ __align__(16) unsigned char fwd_smem[];
    float4 (*stage)[FWD_BATCH * REC_PITCH] = reinterpret_cast<float4 (*)[FWD_BATCH * REC_PITCH]>(fwd_smem);
    __shared__ uint8_t survivors[FWD_WARPS][FWD_BATCH];
    const int tr = threadIdx.x, lane = tr & 31, warp = tr >> 5;
    uint8_t *__restrict__ my_list = survivors[warp];
    const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
    int lx, ly;
    tile_pixel(p.bw, tr, lx, ly);
    const int col = blockIdx.x * p.bw + lx, row = blockIdx.y * p.bw + ly;
    const bool inside = (tr < p.bw * p.bw) && col < p.img_w && row < p.img_h;
    const PixelConsts pc = make_pixel(col, row, p.c2w, p.viewmat, p.fx, p.fy, p.cx, p.cy);
    const WarpRect wr = make_warp_rect(col, row, inside);
    const int2 range = p.bins[tile];
    const int total = range.y - range.x;
    const int nbatch = (total + FWD_BATCH - 1) / FWD_BATCH;
    // the saved state bounds the walk: a pixel's blends end at the last entry it composited (a pixel that composited
    // nothing has final_T == 1 exactly; its final_idx of 0 would be ambiguous)
    int my_last = -1;
    if (inside) {
        const int pix = row * p.img_w + col;
        my_last = final_Ts[pix] < 1.f ? final_idx[pix] : -1;
    }
    bool done = my_last < range.x;
    if (nbatch > 0) stage_records(stage[0], p.recs, p.ids, range.x, min(FWD_BATCH, total), tr, p.nthreads);
    for (int b = 0; b < nbatch; ++b) {
        const int first = range.x + b * FWD_BATCH;
        const int cnt = min(FWD_BATCH, range.y - first);
        if (!fwd_stage_advance(stage, p, range, b, nbatch, tr, done)) break;
        const float4 *__restrict__ S = stage[b % FWD_STAGES];
        const int nsurv = __all_sync(0xffffffffu, done) ? 0 : build_survivors<BLUR, true>(S, 0, cnt, wr, p.mean2d, my_list, lane);
        for (int si = 0; si < nsurv; ++si) {
            const int i = my_list[si];
            const float4 *__restrict__ R = S + i * REC_PITCH;
            PairEval pe;
            eval_pair<BLUR>(R[0], R[1], R[2], R[3], pc, p.mean2d, pe);
            done = done || first + i > my_last;
            if (__all_sync(0xffffffffu, done)) break;
            const unsigned bm = __ballot_sync(0xffffffffu, !done && !pair_skipped(pe));
            if (bm != 0u && lane == 0) p.masks[(size_t)(first + i) * MASK_WARPS + warp] = bm;
        }
    }
    __pipeline_wait_prior(0);
}

// C3 = true : 3-channel texture read through the padded float4 copy, accumulators in registers.
// C3 = false: runtime channel count (<= 64) read from the caller's (X,C) array (slow generic path).
template <bool C3, bool BLUR, bool VIS>
__global__ void __launch_bounds__(RASTER_MAX_THREADS, GSTEX_FWD_MINB) raster_forward_kernel(const RasterCommon p, const ForwardOut o) {
    extern __shared__ __align__(16) unsigned char fwd_smem[];
    float4 (*stage)[FWD_BATCH * REC_PITCH] = reinterpret_cast<float4 (*)[FWD_BATCH * REC_PITCH]>(fwd_smem);
    __shared__ uint8_t survivors[FWD_WARPS][FWD_BATCH];

    const int tr = threadIdx.x, lane = tr & 31, warp = tr >> 5;
    uint8_t *__restrict__ my_list = survivors[warp];
    const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
    int lx, ly;
    tile_pixel(p.bw, tr, lx, ly);
    const int col = blockIdx.x * p.bw + lx, row = blockIdx.y * p.bw + ly;
    const bool inside = (tr < p.bw * p.bw) && col < p.img_w && row < p.img_h;
    const PixelConsts pc = make_pixel(col, row, p.c2w, p.viewmat, p.fx, p.fy, p.cx, p.cy);
    const WarpRect wr = make_warp_rect(col, row, inside);
    const bool use_ndc = (p.settings & GSTEX_SET_NDC) != 0;
    const bool bilinear = !(p.settings & GSTEX_SET_NEAREST);
    const int C = C3 ? 3 : p.channels;
    const bool vis_normals = VIS && (p.settings & GSTEX_SET_VIS_NORMALS) != 0;
    const bool vis_alpha = VIS && (p.settings & GSTEX_SET_VIS_ALPHA) != 0;
    const bool vis_opac = VIS && (p.settings & GSTEX_SET_VIS_OPACITY_THRESH) != 0;
    const bool vis_white = VIS && (p.settings & GSTEX_SET_VIS_WHITE_OUTLINE) != 0;
    const float alpha_bound = (float)((p.settings & GSTEX_SET_VIS_ALPHA_BOUND) >> 17) / 8.0f;
    const float outline_bound = (float)((p.settings & GSTEX_SET_VIS_OUTLINE_BOUND) >> 26) / 4.0f;
    const float sigma_thresh = 0.5f * alpha_bound * alpha_bound;

    const int2 range = p.bins[tile];
    const int total = range.y - range.x;
    const int nbatch = (total + FWD_BATCH - 1) / FWD_BATCH;

    float T = 1.f;
    float acc_c0 = 0.f, acc_c1 = 0.f, acc_c2 = 0.f;
    float acc_n0 = 0.f, acc_n1 = 0.f, acc_n2 = 0.f;
    float acc_t[C3 ? 3 : RASTER_MAX_C];
#pragma unroll
    for (int c = 0; c < (C3 ? 3 : RASTER_MAX_C); ++c) acc_t[c] = 0.f;
    float depth = 0.f, reg = 0.f, S0 = 0.f, S1 = 0.f, S2 = 0.f;
    int last = 0, dlast = -1;
    bool done = !inside;

#if GSTEX_FWD_PIPE
    __shared__ __align__(8) uint64_t full_bar[FWD_STAGES], empty_bar[FWD_STAGES];
    __shared__ int warps_done;  // warps whose 32 pixels are all finished
    const int nwarps = p.nthreads >> 5;
    if (tr == 0) {
        for (int k = 0; k < FWD_STAGES; ++k) {
            mbar_init(&full_bar[k], 1);
            mbar_init(&empty_bar[k], nwarps);
        }
        warps_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    bool warp_done = false;
    int last_issued = -1;  // warp 0: the last stage it issued
    if (warp == 0 && nbatch > 0) {
        issue_stage_bulk(stage[0], p.recs, p.ids, range.x, min(FWD_BATCH, total), lane, &full_bar[0]);
        last_issued = 0;
    }
#else
    if (nbatch > 0) stage_records(stage[0], p.recs, p.ids, range.x, min(FWD_BATCH, total), tr, p.nthreads);
#endif

    for (int b = 0; b < nbatch; ++b) {
        const int first = range.x + b * FWD_BATCH;
        const int cnt = min(FWD_BATCH, range.y - first);
#if GSTEX_FWD_PIPE
        // warp 0: stage b+1 goes into the slot that held stage b-2, once every warp has released that stage
        if (warp == 0 && b + 1 < nbatch) {
            const int slot = (b + 1) % FWD_STAGES;
            if (b >= 2 && !mbar_wait_unless(&empty_bar[slot], (uint32_t)(((b - 2) / FWD_STAGES) & 1), &warps_done, nwarps)) break;
            if (flag_reached(&warps_done, nwarps)) break;
            issue_stage_bulk(stage[slot], p.recs, p.ids, first + FWD_BATCH, min(FWD_BATCH, range.y - first - FWD_BATCH), lane,
                             &full_bar[slot]);
            last_issued = b + 1;
            // the ids of stage b+2 start travelling now: the next issue begins with a dependent load of them
            if (lane < (FWD_BATCH + 31) / 32 && first + 2 * FWD_BATCH + 32 * lane < range.y)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.ids + first + 2 * FWD_BATCH + 32 * lane));
        }
        if (warp_done) {
            // nothing left to composite for this warp: it only keeps releasing the stages - never ahead of the phase
            // its arrival belongs to (an arrival for stage b may only land once the slot's phase of stage b-3 is over)
            if (b >= FWD_STAGES && !mbar_wait_unless(&empty_bar[b % FWD_STAGES], (uint32_t)(((b - FWD_STAGES) / FWD_STAGES) & 1),
                                                     &warps_done, nwarps))
                break;
            if (flag_reached(&warps_done, nwarps)) break;
            if (lane == 0) mbar_arrive(&empty_bar[b % FWD_STAGES]);
            continue;
        }
        if (!mbar_wait_unless(&full_bar[b % FWD_STAGES], (uint32_t)((b / FWD_STAGES) & 1), &warps_done, nwarps)) break;
#else
        if (!fwd_stage_advance(stage, p, range, b, nbatch, tr, done)) break;
#endif
        const float4 *__restrict__ S = stage[b % FWD_STAGES];
        if (C3 && tr < cnt) {  // one thread per staged record: start fetching its texture block
            const float4 q3 = S[quad_slot(tr, 3)], q6 = S[quad_slot(tr, 6)];
            prefetch_texture_block(p.tex4, __float_as_int(q6.w), __float_as_int(q3.z), __float_as_int(q3.w));
        }
        const int nsurv = __all_sync(0xffffffffu, done) ? 0 : build_survivors<BLUR, !VIS>(S, 0, cnt, wr, p.mean2d, my_list, lane);
        // this warp's mask word of the stage's first entry (only lane 0 stores; a NULL array keeps the pointer NULL)
        uint32_t *__restrict__ mrow = (!VIS && p.masks && lane == 0) ? p.masks + (size_t)first * MASK_WARPS + warp : nullptr;
        // The survivor walk is warp-uniform: finished pixels stay in the loop (predicated off) so that the blend
        // decision of all 32 pixels is one ballot - the mask word the backward pass reads.
        for (int si = 0; si < nsurv; ++si) {
            const int i = my_list[si];
            const float4 *__restrict__ R = S + i * REC_PITCH;
            const float4 q0 = R[0], q1 = R[1], q2 = R[2], q3 = R[3];
            PairEval pe;
            eval_pair<BLUR>(q0, q1, q2, q3, pc, p.mean2d, pe);
            if (VIS && vis_alpha) {  // reference texture.cu:201-211
                const float sigma = LN2_F * fmaf(pe.l1, pe.l1, pe.l2 * pe.l2);
                pe.alpha = (sigma > sigma_thresh || (vis_opac && q0.w < 0.5f)) ? 0.f : ALPHA_CAP;
            }
            const float next_T = __fmul_rn(T, __fsub_rn(1.f, pe.alpha));
            // the stop rule is tested even for skipped Gaussians (reference texture.cu:216-221)
            done = done || next_T <= T_STOP;
            const bool blend = !done && !pair_skipped(pe);
            const unsigned bm = __ballot_sync(0xffffffffu, blend);
            if (bm == 0u) {  // nobody blends: the only place where "every pixel of the warp is finished" can become true
                if (__all_sync(0xffffffffu, done)) break;
                continue;
            }
            if (mrow) mrow[(unsigned)i * MASK_WARPS] = bm;
            if (blend) {
                const float4 q4 = R[4], q5 = R[5], q6 = R[6], q7 = R[7];
                const float vis = pe.alpha * T;
                // VIS: outline pixels of the hard footprint carry no colour (or white), texture.cu:226-235, :269-274
                bool draw = true, white = false;
                if (VIS && vis_alpha) {
                    draw = outline_distance(q0, q1, q2, q3, pc, sigma_thresh) > outline_bound;
                    white = !draw && vis_white;
                }
                const float cvis = draw ? vis : 0.f, wvis = white ? vis : 0.f;  // == vis, 0 outside VIS builds
                acc_c0 = VIS ? fmaf(q6.x, cvis, acc_c0) + wvis : fmaf(q6.x, vis, acc_c0);
                acc_c1 = VIS ? fmaf(q6.y, cvis, acc_c1) + wvis : fmaf(q6.y, vis, acc_c1);
                acc_c2 = VIS ? fmaf(q6.z, cvis, acc_c2) + wvis : fmaf(q6.z, vis, acc_c2);
                // VIS: normals face the camera; D = a3 . R_w(p) has the sign of the reference's dot(ray, ax3) (:236-241)
                const float nvis = (vis_normals && pe.rD > 0.f) ? -vis : vis;
                acc_n0 = fmaf(q7.x, nvis, acc_n0);
                acc_n1 = fmaf(q7.y, nvis, acc_n1);
                acc_n2 = fmaf(q7.z, nvis, acc_n2);
                const float nu = fmaf(q4.x, pe.ex, fmaf(q4.y, pe.ey, q4.z));
                const float nv = fmaf(q5.x, pe.ex, fmaf(q5.y, pe.ey, q5.z));
                const float u = clamp01(fmaf(nu, pe.rD, q4.w)), v = clamp01(fmaf(nv, pe.rD, q5.w));
                TexFetch tf;
                texel_setup(__float_as_int(q3.z), __float_as_int(q3.w), __float_as_int(q6.w), u, v, bilinear, tf);
                if (C3) {
                    const float4 t0 = __ldg(p.tex4 + tf.idx[0]), t1 = __ldg(p.tex4 + tf.idx[1]);
                    const float4 t2 = __ldg(p.tex4 + tf.idx[2]), t3 = __ldg(p.tex4 + tf.idx[3]);
                    const float tvis = VIS ? cvis : vis;
                    const float w0 = tf.w[0] * tvis, w1 = tf.w[1] * tvis, w2 = tf.w[2] * tvis, w3 = tf.w[3] * tvis;
                    acc_t[0] += w0 * t0.x + w1 * t1.x + w2 * t2.x + w3 * t3.x;
                    acc_t[1] += w0 * t0.y + w1 * t1.y + w2 * t2.y + w3 * t3.y;
                    acc_t[2] += w0 * t0.z + w1 * t1.z + w2 * t2.z + w3 * t3.z;
                    if (VIS) {
                        acc_t[0] += wvis; acc_t[1] += wvis; acc_t[2] += wvis;
                    }
                } else {
                    const float *__restrict__ tx = p.tex;
                    for (int c = 0; c < C; ++c) {
                        const float val = tf.w[0] * __ldg(tx + (size_t)tf.idx[0] * C + c) +
                                          tf.w[1] * __ldg(tx + (size_t)tf.idx[1] * C + c) +
                                          tf.w[2] * __ldg(tx + (size_t)tf.idx[2] * C + c) +
                                          tf.w[3] * __ldg(tx + (size_t)tf.idx[3] * C + c);
                        acc_t[c] = VIS ? fmaf(cvis, val, acc_t[c]) + wvis : fmaf(vis, val, acc_t[c]);
                    }
                }
                const float t_view = pe.t * pc.vdep;
                if (T > 0.5f) {  // median depth (reference texture.cu:286-291)
                    depth = t_view;
                    dlast = first + i;
                }
                float tv = pe.t;
                if (use_ndc) tv = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view);
                reg += vis * (tv * tv * S0 + S2 - 2.f * tv * S1);  // helpers.cuh:259-264
                S0 += vis;
                S1 += vis * tv;
                S2 += vis * tv * tv;
                T = next_T;
                last = first + i;
            }
        }
#if GSTEX_FWD_PIPE
        // release the stage; a warp whose pixels are all finished says so once (the CTA stops staging when all have)
        warp_done = __all_sync(0xffffffffu, done);
        __syncwarp();
        if (lane == 0) {
            if (warp_done) atomicAdd(&warps_done, 1);
            mbar_arrive(&empty_bar[b % FWD_STAGES]);
        }
#endif
    }
#if GSTEX_FWD_PIPE
    // no bulk copy may still be in flight towards this CTA's shared memory when it is retired: an early exit (every
    // pixel finished) can leave the last one or two issued stages unread.  Warp 0 issued them; it stays until they landed.
    for (int j = max(0, last_issued - 1); j <= last_issued; ++j)
        while (!mbar_try_wait(&full_bar[j % FWD_STAGES], (uint32_t)((j / FWD_STAGES) & 1))) {
        }
#else
    __pipeline_wait_prior(0);
#endif

    if (inside) {
        const int pix = row * p.img_w + col;
        const float bg0 = p.background[0], bg1 = p.background[1], bg2 = p.background[2];
        o.final_Ts[pix] = T;
        o.final_idx[pix] = last;
        o.depth_idx[pix] = dlast;
        o.out_img[3 * pix + 0] = fmaf(T, bg0, acc_c0);
        o.out_img[3 * pix + 1] = fmaf(T, bg1, acc_c1);
        o.out_img[3 * pix + 2] = fmaf(T, bg2, acc_c2);
        o.out_normal[3 * pix + 0] = acc_n0;
        o.out_normal[3 * pix + 1] = acc_n1;
        o.out_normal[3 * pix + 2] = acc_n2;
        o.out_depth[pix] = depth;
        o.out_reg[pix] = reg;
        o.out_reg_s[3 * pix + 0] = S0;
        o.out_reg_s[3 * pix + 1] = S1;
        o.out_reg_s[3 * pix + 2] = S2;
        if (C3) {
            o.out_texture[3 * pix + 0] = acc_t[0];
            o.out_texture[3 * pix + 1] = acc_t[1];
            o.out_texture[3 * pix + 2] = acc_t[2];
        } else {
            for (int c = 0; c < C; ++c) o.out_texture[(size_t)C * pix + c] = acc_t[c];
        }
    }
}

RasterCommon make_raster_common(int img_height, int img_width, int block_width, int channels, int settings,
                                const int32_t *ids, const int32_t *tile_bins, const float4 *recs, const float2 *mean2d,
                                const float4 *tex4, const float *tex, const float *viewmat, const float *c2w,
                                const float *background, float fx, float fy, float cx, float cy, uint32_t *masks) {
    RasterCommon p;
    p.img_w = img_width;
    p.img_h = img_height;
    p.tiles_x = ceil_div(img_width, block_width);
    p.bw = block_width;
    p.nthreads = ceil_div(block_width * block_width, 32) * 32;
    p.settings = settings & GSTEX_SET_SUPPORTED_FORWARD;  // unknown bits are ignored (see check_raster_args)
    p.channels = channels;
    p.ids = ids;
    p.bins = (const int2 *)tile_bins;
    p.recs = recs;
    p.mean2d = mean2d;
    p.tex4 = tex4;
    p.tex = tex;
    p.viewmat = viewmat;
    p.c2w = c2w;
    p.background = background;
    p.fx = fx; p.fy = fy; p.cx = cx; p.cy = cy;
    p.masks = masks;
    return p;
}

// grid-stride zero fill of the live part of the mask array (the live length may only be known on the device)
__global__ void __launch_bounds__(256) zero_masks_kernel(uint4 *__restrict__ masks, int64_t entries,
                                                         const int32_t *__restrict__ d_count) {
    const int64_t live = d_count ? min((int64_t)max(*d_count, 0), entries) : entries;
    const int64_t quads = live * (MASK_WARPS / 4);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x)
        masks[q] = make_uint4(0u, 0u, 0u, 0u);
}

template <bool C3, bool BLUR, bool VIS>
static int launch_fwd_variant(const dim3 grid, const RasterCommon &p, const ForwardOut &o, cudaStream_t s) {
    static SmemOnceFlags once;  // one per template instantiation
    const int rc = configure_dynamic_smem((const void *)raster_forward_kernel<C3, BLUR, VIS>, fwd_smem_bytes(), true, once);
    if (rc != GSTEX_OK) return rc;
    raster_forward_kernel<C3, BLUR, VIS><<<grid, p.nthreads, fwd_smem_bytes(), s>>>(p, o);
    return GSTEX_OK;
}

static int zero_masks(const RasterCommon &p, int64_t mask_entries, const int32_t *d_count, cudaStream_t s) {
    if (p.masks && mask_entries > 0) {
        const int blocks = (int)min((int64_t)148 * 8, ceil_div64(mask_entries * (MASK_WARPS / 4), 256));
        zero_masks_kernel<<<blocks, 256, 0, s>>>((uint4 *)p.masks, mask_entries, d_count);
        GSTEX_LAUNCH_OK("zero_masks_kernel");
    }
    return GSTEX_OK;
}

int launch_raster_forward(const RasterCommon &p, const ForwardOut &o, int64_t mask_entries, const int32_t *d_count,
                          cudaStream_t s) {
    int rc = zero_masks(p, mask_entries, d_count, s);
    if (rc != GSTEX_OK) return rc;
    const dim3 grid(p.tiles_x, ceil_div(p.img_h, p.bw));
    const bool blur = (p.settings & GSTEX_SET_BLUR) != 0;
    if (p.settings & GSTEX_SET_VIS_ALL) {  // viewer-only modes: one generic build per channel layout, forward only
        if (p.channels == 3) rc = blur ? launch_fwd_variant<true, true, true>(grid, p, o, s) : launch_fwd_variant<true, false, true>(grid, p, o, s);
        else rc = blur ? launch_fwd_variant<false, true, true>(grid, p, o, s) : launch_fwd_variant<false, false, true>(grid, p, o, s);
    } else if (p.channels == 3) {
        rc = blur ? launch_fwd_variant<true, true, false>(grid, p, o, s) : launch_fwd_variant<true, false, false>(grid, p, o, s);
    } else {
        rc = blur ? launch_fwd_variant<false, true, false>(grid, p, o, s) : launch_fwd_variant<false, false, false>(grid, p, o, s);
    }
    if (rc != GSTEX_OK) return rc;
    GSTEX_LAUNCH_OK("raster_forward_kernel");
    return GSTEX_OK;
}

// Blend masks of a finished forward pass from its saved state: see raster_masks_kernel.
int launch_raster_masks(const RasterCommon &p, const float *final_Ts, const int32_t *final_idx, int64_t mask_entries,
                        const int32_t *d_count, cudaStream_t s) {
    int rc = zero_masks(p, mask_entries, d_count, s);
    if (rc != GSTEX_OK) return rc;
    const dim3 grid(p.tiles_x, ceil_div(p.img_h, p.bw));
    static SmemOnceFlags once[2];
    if (p.settings & GSTEX_SET_BLUR) {
        rc = configure_dynamic_smem((const void *)raster_masks_kernel<true>, fwd_smem_bytes(), true, once[1]);
        if (rc != GSTEX_OK) return rc;
        raster_masks_kernel<true><<<grid, p.nthreads, fwd_smem_bytes(), s>>>(p, final_Ts, final_idx);
    } else {
        rc = configure_dynamic_smem((const void *)raster_masks_kernel<false>, fwd_smem_bytes(), true, once[0]);
        if (rc != GSTEX_OK) return rc;
        raster_masks_kernel<false><<<grid, p.nthreads, fwd_smem_bytes(), s>>>(p, final_Ts, final_idx);
    }
    GSTEX_LAUNCH_OK("raster_masks_kernel");
    return GSTEX_OK;
}

FwdLayout forward_layout(int n, int64_t num_texels, int channels, int64_t num_intersects) {
    FwdLayout L;
    size_t off = 0;
    L.recs_off = off;
    off = align_up(off + sizeof(float) * REC_FLOATS * (size_t)(n > 0 ? n : 1), 256);
    L.mean2d_off = off;
    off = align_up(off + sizeof(float2) * (size_t)(n > 0 ? n : 1), 256);
    L.tex4_off = off;
    if (channels == 3) off = align_up(off + sizeof(float4) * (size_t)(num_texels > 0 ? num_texels : 1), 256);
    L.masks_off = off;
    off = align_up(off + sizeof(uint32_t) * MASK_WARPS * (size_t)(num_intersects > 0 ? num_intersects : 1), 256);
    L.total = off;
    return L;
}

int check_raster_args(const char *who, int img_height, int img_width, int block_width, int n, int64_t num_texels,
                      int channels, int settings, int supported_settings) {
    GSTEX_REQUIRE(img_height > 0 && img_width > 0, GSTEX_E_INVALID, "%s: image %dx%d", who, img_height, img_width);
    GSTEX_REQUIRE(block_width > 1 && block_width <= 16, GSTEX_E_INVALID,
                  "%s: block_width must be between 2 and 16 (got %d)", who, block_width);
    GSTEX_REQUIRE(n >= 0 && num_texels >= 0 && num_texels < ((int64_t)1 << 31), GSTEX_E_INVALID,
                  "%s: n = %d, texels = %lld", who, n, (long long)num_texels);
    GSTEX_REQUIRE(channels >= 1 && channels <= RASTER_MAX_C, GSTEX_E_INVALID,
                  "%s: texture channels must be in [1, %d] (got %d)", who, RASTER_MAX_C, channels);
    // Bits no reference kernel reads (0, 1 - texture_edit's blur / ndc when one settings word is shared -, 3-7, 11-14,
    // 30, 31) are ignored, as upstream ignores them.  Only bits this build KNOWS but the entry point does not implement
    // are an error: the visualisation bits 15-29 in a backward call.
    GSTEX_REQUIRE((settings & GSTEX_SET_SUPPORTED_FORWARD & ~supported_settings) == 0, GSTEX_E_UNSUPPORTED,
                  "%s: settings 0x%x has bits this entry point does not implement (supported mask 0x%x); the "
                  "visualisation bits 15-29 are forward-only", who, settings, supported_settings);
    return GSTEX_OK;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_texture_forward_temp_bytes(int n, int64_t num_texels, int channels, int64_t num_intersects) {
    return forward_layout(n, num_texels, channels, num_intersects).total;
}

extern "C" int gstex_texture_forward(int img_height, int img_width, int block_width, int n, int64_t num_texels,
                                     int channels, int64_t num_intersects, const int32_t *texture_dims,
                                     const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
                                     const float *opacities, const float *means, const float *scales, float glob_scale,
                                     const float *quats, const float *uv0, const float *umap, const float *vmap,
                                     const float *texture, const float *viewmat, const float *c2w, float fx, float fy,
                                     float cx, float cy, int settings, const float *background, float *out_img,
                                     float *out_depth, float *out_reg, float *out_texture, float *out_normal,
                                     float *final_Ts, int32_t *final_idx, int32_t *depth_idx, float *out_reg_s,
                                     void *temp, size_t temp_bytes, gstex_stream_t stream) {
    int rc = check_raster_args("texture_forward", img_height, img_width, block_width, n, num_texels, channels, settings,
                               GSTEX_SET_SUPPORTED_FORWARD);
    if (rc != GSTEX_OK) return rc;
    GSTEX_REQUIRE(num_intersects >= 0 && num_intersects < ((int64_t)1 << 31), GSTEX_E_INVALID,
                  "texture_forward: num_intersects = %lld", (long long)num_intersects);
    const FwdLayout L = forward_layout(n, num_texels, channels, num_intersects);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE, "texture_forward: temp too small (%zu < %zu)",
                  temp_bytes, L.total);
    cudaStream_t s = as_stream(stream);
    char *base = (char *)temp;
    float4 *recs = (float4 *)(base + L.recs_off);
    float2 *mean2d = (float2 *)(base + L.mean2d_off);
    float4 *tex4 = (float4 *)(base + L.tex4_off);
    // num_intersects == 0 also serves inference-only callers: no blend masks are kept (and none zero-filled)
    uint32_t *masks = num_intersects > 0 ? (uint32_t *)(base + L.masks_off) : nullptr;
    rc = launch_pack(n, means, scales, glob_scale, quats, opacities, colors, uv0, umap, vmap, texture_dims, viewmat, c2w,
                     fx, fy, cx, cy, recs, mean2d, s);
    if (rc != GSTEX_OK) return rc;
    if (channels == 3) {
        rc = launch_pad_texture(num_texels, texture, tex4, s);
        if (rc != GSTEX_OK) return rc;
    }
    const RasterCommon p = make_raster_common(img_height, img_width, block_width, channels, settings,
                                              gaussian_ids_sorted, tile_bins, recs, mean2d, tex4, texture, viewmat,
                                              c2w, background, fx, fy, cx, cy, masks);
    ForwardOut o{out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, out_reg_s, final_idx, depth_idx};
    return launch_raster_forward(p, o, num_intersects, nullptr, s);
}
