/*
 * gstex_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the algorithms on the reference's hot path
 * (victor-rong/GStex_cuda @ abdc217), written per pixel / per Gaussian as
 * sequential loops.  It exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py can check and time the
 * reference behaviour without a GPU.  The product path (gstex_cuda_b200/)
 * never links, imports or calls anything in this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   (a) vectors produced by importing the reference's own pure-PyTorch twin
 *       (gstex_cuda/_torch_impl.py) in the build container
 *       (tests/golden/make_golden_torch_impl.py), and
 *   (b) vectors produced by the unmodified reference CUDA extension
 *       (oracle/_ref, built by oracle/build_ref.py) run on a B200
 *       (tests/golden/make_golden_ref_cuda.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * gstex_cuda/cuda/csrc/ unless stated otherwise).
 *
 * Conventions: viewmat / c2w are 4x4 row-major, quats
 * are (w,x,y,z) and already normalised, surfel axes are the columns of R(q),
 * scales are linear, texture_dims[g] = (h, w, first_texel), u indexes rows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3_load(const float *p) { return v3_make(p[0], p[1], p[2]); }

/* helpers.cuh:124-131  (R p + t, row-major 4x4, first three rows) */
static inline v3 xform_point(const float *m, v3 p) {
    return v3_make(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
                   m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                   m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}

/* helpers.cuh:114-121  (R^T p) */
static inline v3 xform_rot_t(const float *m, v3 p) {
    return v3_make(m[0] * p.x + m[4] * p.y + m[8] * p.z,
                   m[1] * p.x + m[5] * p.y + m[9] * p.z,
                   m[2] * p.x + m[6] * p.y + m[10] * p.z);
}

/* helpers.cuh:145-152 */
static inline void pinhole(float fx, float fy, float cx, float cy, v3 pv, float *ox, float *oy) {
    float rw = 1.f / (pv.z + 1e-6f);
    *ox = (pv.x * rw) * fx + cx;
    *oy = (pv.y * rw) * fy + cy;
}

/* helpers.cuh:155-164 */
static inline v3 pinhole_vjp(float fx, float fy, v3 pv, float vx, float vy) {
    float rw = 1.f / (pv.z + 1e-6f);
    float gx = fx * vx, gy = fy * vy;
    return v3_make(gx * rw, gy * rw, -(gx * pv.x + gy * pv.y) * rw * rw);
}

/* helpers.cuh:166-185: columns of the rotation matrix of a unit quaternion (w,x,y,z) */
static inline void surfel_axes(const float *q, v3 *a1, v3 *a2, v3 *a3) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    *a1 = v3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
    *a2 = v3_make(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
    *a3 = v3_make(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
}

/* helpers.cuh:187-228: VJP of surfel_axes; g1,g2,g3 are the gradients of the three columns */
static inline void surfel_axes_vjp(const float *q, v3 g1, v3 g2, v3 g3, float *vq) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    vq[0] = 2.f * (x * (g2.z - g3.y) + y * (g3.x - g1.z) + z * (g1.y - g2.x));
    vq[1] = 2.f * (-2.f * x * (g2.y + g3.z) + y * (g1.y + g2.x) + z * (g1.z + g3.x) + w * (g2.z - g3.y));
    vq[2] = 2.f * (x * (g1.y + g2.x) - 2.f * y * (g1.x + g3.z) + z * (g2.z + g3.y) + w * (g3.x - g1.z));
    vq[3] = 2.f * (x * (g1.z + g3.x) + y * (g2.z + g3.y) - 2.f * z * (g1.x + g2.y) + w * (g1.y - g2.x));
}

/* ------------------------------------------------------------------------------------------
 * project_points: gstex_cuda/get_aabb_2d.py:22-32 with clip=False (+ _torch_impl.py:140-147)
 * viewmat: first 12 floats of the row-major 4x4.  depths are NOT clipped.
 * ---------------------------------------------------------------------------------------- */
void orc_project_points(int n, const float *means, const float *viewmat, float fx, float fy, float cx,
                        float cy, float *pix, float *depths) {
    for (int i = 0; i < n; ++i) {
        v3 pv = xform_point(viewmat, v3_load(means + 3 * i));
        if (pix) pinhole(fx, fy, cx, cy, pv, &pix[2 * i], &pix[2 * i + 1]);
        depths[i] = pv.z;
    }
}

/* ------------------------------------------------------------------------------------------
 * get_aabb_2d_kernel: get_aabb_2d.cu:11-89
 * ---------------------------------------------------------------------------------------- */
static inline void corner_pix(v3 c, const float *viewmat, float near_z, float fx, float fy, float cx,
                              float cy, float *ox, float *oy) {
    /* helpers.cuh:250-257 */
    v3 pv = xform_point(viewmat, c);
    pv.z = pv.z > near_z ? pv.z : near_z;
    pinhole(fx, fy, cx, cy, pv, ox, oy);
}

void orc_get_aabb_2d(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                     const float *viewmat, float fx, float fy, float cx, float cy, float *centers,
                     float *extents) {
    const float near_z = 0.01f;
    for (int i = 0; i < n; ++i) {
        v3 m = v3_load(means + 3 * i);
        v3 pv = xform_point(viewmat, m);
        int clipped = pv.z <= near_z; /* helpers.cuh:240-248 */
        float mx, my;
        pinhole(fx, fy, cx, cy, pv, &mx, &my);
        v3 a1, a2, a3;
        surfel_axes(quats + 4 * i, &a1, &a2, &a3);
        float ell = 3.0f * glob_scale;
        float s1 = scales[3 * i], s2 = scales[3 * i + 1];
        float lo_x = 0, lo_y = 0, hi_x = 0, hi_y = 0;
        for (int k = 0; k < 4; ++k) {
            float sa = (k & 2) ? -1.f : 1.f, sb = (k & 1) ? -1.f : 1.f;
            v3 c = v3_make(m.x + sa * (ell * s1 * a1.x) + sb * (ell * s2 * a2.x),
                           m.y + sa * (ell * s1 * a1.y) + sb * (ell * s2 * a2.y),
                           m.z + sa * (ell * s1 * a1.z) + sb * (ell * s2 * a2.z));
            float px, py;
            corner_pix(c, viewmat, near_z, fx, fy, cx, cy, &px, &py);
            if (k == 0) { lo_x = hi_x = px; lo_y = hi_y = py; }
            else {
                lo_x = fminf(lo_x, px); hi_x = fmaxf(hi_x, px);
                lo_y = fminf(lo_y, py); hi_y = fmaxf(hi_y, py);
            }
        }
        if (clipped) { /* get_aabb_2d.cu:81-84 */
            centers[2 * i] = mx; centers[2 * i + 1] = my;
            extents[2 * i] = 0.f; extents[2 * i + 1] = 0.f;
        } else {
            centers[2 * i] = 0.5f * (hi_x + lo_x); centers[2 * i + 1] = 0.5f * (hi_y + lo_y);
            extents[2 * i] = 0.5f * (hi_x - lo_x); extents[2 * i + 1] = 0.5f * (hi_y - lo_y);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * get_num_tiles_hit_2d: gstex_cuda/get_aabb_2d.py:70-92  (floor-based, fp32 torch ops)
 * ---------------------------------------------------------------------------------------- */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void orc_num_tiles_hit(int n, const float *centers, const float *extents, int img_h, int img_w, int bw,
                       int32_t *num_tiles_hit) {
    int tx = (img_w + bw - 1) / bw, ty = (img_h + bw - 1) / bw;
    float fbw = (float)bw;
    for (int i = 0; i < n; ++i) {
        float cx = centers[2 * i], cy = centers[2 * i + 1], ex = extents[2 * i], ey = extents[2 * i + 1];
        int l = clampi((int)floorf((cx - ex) / fbw), 0, tx);
        int t = clampi((int)floorf((cy - ey) / fbw), 0, ty);
        int r = clampi((int)floorf((cx + ex) / fbw + 1.f), 0, tx);
        int b = clampi((int)floorf((cy + ey) / fbw + 1.f), 0, ty);
        num_tiles_hit[i] = (r - l) * (b - t);
    }
}

/* utils.py:40-59: inclusive int32 cumsum; returns the total */
int orc_cumsum(int n, const int32_t *v, int32_t *out) {
    int32_t acc = 0;
    for (int i = 0; i < n; ++i) { acc += v[i]; out[i] = acc; }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * map_gaussian_to_intersects: forward.cu:13-71 (+ helpers.cuh:37-51, 75-92); wrapped (torus) tile boxes:
 * helpers.cuh:53-73, 94-111 and the modulo of forward.cu:53-62.
 * Outputs must be zero-initialised by the caller (bindings.cu:96-99).
 * ---------------------------------------------------------------------------------------- */
static inline void tile_bbox(float cx, float cy, float ex, float ey, int tx, int ty, int bw, int *x0,
                             int *y0, int *x1, int *y1) {
    float fbw = (float)bw;
    float tcx = cx / fbw, tcy = cy / fbw, tex = ex / fbw, tey = ey / fbw;
    *x0 = clampi((int)(tcx - tex), 0, tx);       /* C truncation, helpers.cuh:47-50 */
    *x1 = clampi((int)(tcx + tex + 1), 0, tx);
    *y0 = clampi((int)(tcy - tey), 0, ty);
    *y1 = clampi((int)(tcy + tey + 1), 0, ty);
}

/* helpers.cuh:53-73: no clamping; a box that starts at or left of / above tile 0 grows by one tile */
static inline void tile_bbox_wrapped(float cx, float cy, float ex, float ey, int bw, int *x0, int *y0, int *x1,
                                     int *y1) {
    float fbw = (float)bw;
    float tcx = cx / fbw, tcy = cy / fbw, tex = ex / fbw, tey = ey / fbw;
    *x0 = (int)(tcx - tex);
    if (*x0 <= 0) *x0 -= 1;
    *x1 = (int)(tcx + tex + 1);
    *y0 = (int)(tcy - tey);
    if (*y0 <= 0) *y0 -= 1;
    *y1 = (int)(tcy + tey + 1);
}

/* tiles a Gaussian writes under `wrapped` (what the caller's cum_tiles_hit must be the running sum of) */
void orc_num_tiles_hit_wrapped(int n, const float *centers, const float *extents, int bw, int32_t *num_tiles_hit) {
    for (int i = 0; i < n; ++i) {
        float ex = extents[2 * i], ey = extents[2 * i + 1];
        int x0, y0, x1, y1;
        tile_bbox_wrapped(centers[2 * i], centers[2 * i + 1], ex, ey, bw, &x0, &y0, &x1, &y1);
        int c = (x1 > x0 ? x1 - x0 : 0) * (y1 > y0 ? y1 - y0 : 0);
        if (ex <= 1e-4 && ey <= 1e-4) c = 0;
        num_tiles_hit[i] = c;
    }
}

void orc_map_gaussian_to_intersects(int n, const float *centers, const float *extents, const float *depths,
                                    const int32_t *cum_tiles_hit, int tiles_x, int tiles_y, int bw, int wrapped,
                                    int64_t *isect_ids, int32_t *gaussian_ids) {
    for (int i = 0; i < n; ++i) {
        float ex = extents[2 * i], ey = extents[2 * i + 1];
        if (ex <= 1e-4 && ey <= 1e-4) continue; /* forward.cu:32 (double literal on purpose) */
        int x0, y0, x1, y1;
        if (wrapped) tile_bbox_wrapped(centers[2 * i], centers[2 * i + 1], ex, ey, bw, &x0, &y0, &x1, &y1);
        else tile_bbox(centers[2 * i], centers[2 * i + 1], ex, ey, tiles_x, tiles_y, bw, &x0, &y0, &x1, &y1);
        int32_t cur = i == 0 ? 0 : cum_tiles_hit[i - 1];
        int32_t dbits;
        memcpy(&dbits, depths + i, 4);
        int64_t depth_id = (int64_t)dbits; /* sign-extends, forward.cu:48 */
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                int wy = ty, wx = tx;
                if (wrapped) {
                    /* forward.cu:53-62.  tile_bounds is a dim3 (unsigned), so `ti % tile_bounds.y` converts a
                     * negative ti to unsigned first and the reference's `if (ti < 0)` fix-up never fires: a tile
                     * index of -1 lands on (2^32 - 1) mod tiles, which is the torus neighbour only when the
                     * tile count divides 2^32.  Reproduced as is. */
                    wy = (int)((unsigned)wy % (unsigned)tiles_y);
                    wx = (int)((unsigned)wx % (unsigned)tiles_x);
                }
                int64_t tile = (int64_t)wy * tiles_x + wx;
                isect_ids[cur] = (tile << 32) | depth_id;
                gaussian_ids[cur] = i;
                ++cur;
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * torch.sort(int64) + torch.gather(ids, perm): utils.py:159-160.
 * Third-party arithmetic (PyTorch / cub radix sort, version = "whatever torch is installed",
 * here 2.11.0+cu128); restated as a stable LSD radix sort on the signed 64-bit key.
 * ---------------------------------------------------------------------------------------- */
void orc_sort_pairs(int64_t m, const int64_t *keys_in, const int32_t *vals_in, int64_t *keys_out,
                    int32_t *vals_out) {
    if (m <= 0) return;
    uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * m), *kb = (uint64_t *)malloc(sizeof(uint64_t) * m);
    int32_t *va = (int32_t *)malloc(sizeof(int32_t) * m), *vb = (int32_t *)malloc(sizeof(int32_t) * m);
    for (int64_t i = 0; i < m; ++i) { ka[i] = (uint64_t)keys_in[i] ^ 0x8000000000000000ull; va[i] = vals_in[i]; }
    for (int pass = 0; pass < 8; ++pass) {
        int64_t hist[257];
        memset(hist, 0, sizeof(hist));
        int sh = 8 * pass;
        for (int64_t i = 0; i < m; ++i) hist[((ka[i] >> sh) & 255) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (int64_t i = 0; i < m; ++i) {
            int64_t dst = hist[(ka[i] >> sh) & 255]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        int32_t *tv = va; va = vb; vb = tv;
    }
    for (int64_t i = 0; i < m; ++i) { keys_out[i] = (int64_t)(ka[i] ^ 0x8000000000000000ull); vals_out[i] = va[i]; }
    free(ka); free(kb); free(va); free(vb);
}

/* get_tile_bin_edges: forward.cu:76-98.  tile_bins (num_tiles,2) must be zero-initialised. */
void orc_get_tile_bin_edges(int64_t m, const int64_t *keys_sorted, int32_t *tile_bins) {
    for (int64_t i = 0; i < m; ++i) {
        int32_t cur = (int32_t)(keys_sorted[i] >> 32);
        if (i == 0) tile_bins[2 * cur] = 0;
        if (i == m - 1) tile_bins[2 * cur + 1] = (int32_t)m;
        if (i == 0) continue;
        int32_t prev = (int32_t)(keys_sorted[i - 1] >> 32);
        if (prev != cur) { tile_bins[2 * prev + 1] = (int32_t)i; tile_bins[2 * cur] = (int32_t)i; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Jagged bilinear texel addressing: texture_helpers.cuh:155-215 (replicate pad only).
 * Returns the 4 flat texel indices (without the channel term) and the 4 weights.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int idx[4]; float w[4]; float fu, fv; int h, wd; } texfetch;

static inline void texel_setup(const int32_t *dims, float u, float v, int bilinear, int C, texfetch *f) {
    int h = dims[0], w = dims[1], si = dims[2];
    float tu = h * u, tv = w * v;
    int i0 = (int)tu, j0 = (int)tv;
    int i1 = i0 + 1 < h - 1 ? i0 + 1 : h - 1; /* min(i0+1, h-1) */
    int j1 = j0 + 1 < w - 1 ? j0 + 1 : w - 1;
    float fu = tu - (float)i0, fv = tv - (float)j0;
    if (i0 > h - 1) i0 = h - 1;
    if (j0 > w - 1) j0 = w - 1;
    float w00 = (1.f - fu) * (1.f - fv), w01 = (1.f - fu) * fv, w10 = fu * (1.f - fv), w11 = fu * fv;
    f->idx[0] = (si + i0 * w + j0) * C; f->idx[1] = (si + i0 * w + j1) * C;
    f->idx[2] = (si + i1 * w + j0) * C; f->idx[3] = (si + i1 * w + j1) * C;
    if (bilinear) { f->w[0] = w00; f->w[1] = w01; f->w[2] = w10; f->w[3] = w11; }
    else { /* texture_helpers.cuh:199-212: largest weight, first wins */
        int pick = 3;
        if (w00 >= w01 && w00 >= w10 && w00 >= w11) pick = 0;
        else if (w01 >= w00 && w01 >= w10 && w01 >= w11) pick = 1;
        else if (w10 >= w00 && w10 >= w01 && w10 >= w11) pick = 2;
        for (int k = 0; k < 4; ++k) f->w[k] = (k == pick) ? 1.f : 0.f;
    }
    f->fu = fu; f->fv = fv; f->h = h; f->wd = w;
}

static inline float clamp01(float x) { /* texture_helpers.cuh:37-51 with eps = 0 */
    if (x <= 0.f) x = 0.f;
    if (x >= 1.f) x = 1.f;
    return x;
}

/* texture_sample_forward: texture_sample.cu:11-36.
 * The CUDA kernel does not clamp uv (out-of-range queries index out of bounds there); the documented
 * behaviour (texture_sample.py:26-30) and the reference's torch twin (_torch_impl.py:151-153) clamp to
 * [0,1], which is what is restated here. */
void orc_texture_sample_forward(int nq, int C, const int32_t *dims, const float *uvs, const float *texture,
                                float *out) {
    for (int q = 0; q < nq; ++q) {
        texfetch f;
        texel_setup(dims + 3 * q, clamp01(uvs[2 * q]), clamp01(uvs[2 * q + 1]), 1, C, &f);
        for (int c = 0; c < C; ++c)
            out[q * C + c] = f.w[0] * texture[f.idx[0] + c] + f.w[1] * texture[f.idx[1] + c] +
                             f.w[2] * texture[f.idx[2] + c] + f.w[3] * texture[f.idx[3] + c];
    }
}

/* The reference's texture_sample_backward (texture_sample.cu:38-70) is unreachable from Python
 * and reads v_texture where it means v_output (:58).  This is the intended scatter:
 * v_texture[corner] += w_corner * v_output[q].   v_texture must be zero-initialised. */
void orc_texture_sample_backward(int nq, int C, const int32_t *dims, const float *uvs, const float *v_out,
                                 float *v_texture) {
    for (int q = 0; q < nq; ++q) {
        texfetch f;
        texel_setup(dims + 3 * q, clamp01(uvs[2 * q]), clamp01(uvs[2 * q + 1]), 1, C, &f);
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < 4; ++k) v_texture[f.idx[k] + c] += f.w[k] * v_out[q * C + c];
    }
}

/* ------------------------------------------------------------------------------------------
 * Spherical harmonics: sh.cuh:46-118 (forward), :120-210 (VJP), kernels :212-253.
 * ---------------------------------------------------------------------------------------- */
static int sh_num_bases(int degree) { /* sh.cuh:34-44 */
    return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;
}

static void sh_basis(int deg, const float *dir, float *Y) {
    Y[0] = 0.28209479177387814f;
    if (deg < 1) return;
    float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float x = dir[0] / nrm, y = dir[1] / nrm, z = dir[2] / nrm;
    const float c1 = 0.4886025119029199f;
    Y[1] = -c1 * y; Y[2] = c1 * z; Y[3] = -c1 * x;
    if (deg < 2) return;
    float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    Y[4] = 1.0925484305920792f * xy;
    Y[5] = -1.0925484305920792f * yz;
    Y[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
    Y[7] = -1.0925484305920792f * xz;
    Y[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    Y[9] = -0.5900435899266435f * y * (3.f * xx - yy);
    Y[10] = 2.890611442640554f * xy * z;
    Y[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
    Y[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Y[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
    Y[14] = 1.445305721320277f * z * (xx - yy);
    Y[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
    if (deg < 4) return;
    Y[16] = 2.5033429417967046f * xy * (xx - yy);
    Y[17] = -1.7701307697799304f * yz * (3.f * xx - yy); /* sh.cuh:25 lacks the f suffix, but the array is float */
    Y[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f);
    Y[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
    Y[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
    Y[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f);
    Y[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
    Y[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
    Y[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

void orc_sh_forward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *coeffs,
                    float *colors) {
    int K = sh_num_bases(degree), Ku = sh_num_bases(degrees_to_use);
    for (int i = 0; i < n; ++i) {
        float Y[25];
        sh_basis(degrees_to_use, viewdirs + 3 * i, Y);
        for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int b = 0; b < Ku; ++b) acc += Y[b] * coeffs[(i * K + b) * 3 + c];
            colors[3 * i + c] = acc;
        }
    }
}

/* v_coeffs (n,K,3) must be zero-initialised (bindings.cu:62-63); rows >= Ku stay zero. */
void orc_sh_backward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *v_colors,
                     float *v_coeffs) {
    int K = sh_num_bases(degree), Ku = sh_num_bases(degrees_to_use);
    for (int i = 0; i < n; ++i) {
        float Y[25];
        sh_basis(degrees_to_use, viewdirs + 3 * i, Y);
        for (int b = 0; b < Ku; ++b)
            for (int c = 0; c < 3; ++c) v_coeffs[(i * K + b) * 3 + c] = Y[b] * v_colors[3 * i + c];
    }
}

/* ------------------------------------------------------------------------------------------
 * Rasteriser.  Per-pixel state shared by forward and backward.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    v3 origin, ray;
    float view_depth, px, py;
} pixray;

/* texture_helpers.cuh:336-352 and texture.cu:67-74 */
static inline pixray make_ray_at(const float *c2w, const float *viewmat, float fx, float fy, float cx, float cy,
                                 float px, float py) {
    pixray r;
    r.px = px; r.py = py;
    r.origin = v3_make(c2w[3], c2w[7], c2w[11]);
    float u = (r.px - cx) / fx, v = (r.py - cy) / fy;
    v3 d = xform_point(c2w, v3_make(u, v, 1.f));
    d.x -= r.origin.x; d.y -= r.origin.y; d.z -= r.origin.z;
    float nrm = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    r.ray = v3_make(d.x / nrm, d.y / nrm, d.z / nrm);
    r.view_depth = viewmat[8] * r.ray.x + viewmat[9] * r.ray.y + viewmat[10] * r.ray.z;
    return r;
}

static inline pixray make_ray(const float *c2w, const float *viewmat, float fx, float fy, float cx, float cy,
                              int col, int row) {
    return make_ray_at(c2w, viewmat, fx, fy, cx, cy, (float)col + 0.5f, (float)row + 0.5f);
}

/* texture_helpers.cuh:302-313 */
static inline float plane_denominator(v3 n, v3 ray) {
    float den = v3_dot(n, ray);
    const float eps = 1e-6f;
    if (0.f <= den && den < eps) den = eps;
    else if (-eps < den && den <= 0.f) den = -eps;
    return den;
}

/* Everything the blend of one (pixel, Gaussian) pair needs.  texture.cu:161-200 / :518-555 */
typedef struct {
    v3 a1, a2, a3, mean, delta, pview;
    float s1, s2, opac, t, l1, l2, sigma, sigma_blur, e_sig, e_blur, alpha, nb, bl, mx, my;
} pairgeom;

static inline void pair_geometry(const pixray *r, const float *mean, const float *scale, const float *quat,
                                 float opac, float glob_scale, const float *viewmat, float fx, float fy,
                                 float cx, float cy, int use_blur, pairgeom *g) {
    surfel_axes(quat, &g->a1, &g->a2, &g->a3);
    g->mean = v3_load(mean);
    g->s1 = scale[0]; g->s2 = scale[1]; g->opac = opac;
    v3 diff = v3_make(g->mean.x - r->origin.x, g->mean.y - r->origin.y, g->mean.z - r->origin.z);
    g->t = v3_dot(g->a3, diff) / plane_denominator(g->a3, r->ray);
    v3 pos = v3_make(r->origin.x + g->t * r->ray.x, r->origin.y + g->t * r->ray.y, r->origin.z + g->t * r->ray.z);
    g->delta = v3_make(pos.x - g->mean.x, pos.y - g->mean.y, pos.z - g->mean.z);
    g->l1 = v3_dot(g->delta, g->a1);
    g->l2 = v3_dot(g->delta, g->a2);
    float is1 = 1.f / (g->s1 * glob_scale), is2 = 1.f / (g->s2 * glob_scale);
    g->sigma = 0.5f * (is1 * is1 * g->l1 * g->l1 + is2 * is2 * g->l2 * g->l2);
    g->pview = xform_point(viewmat, g->mean);
    pinhole(fx, fy, cx, cy, g->pview, &g->mx, &g->my);
    float dx = g->mx - r->px, dy = g->my - r->py;
    g->sigma_blur = 0.5f * 2.0f * (dx * dx + dy * dy);
    g->nb = 1.f; g->bl = 0.f;
    if (use_blur && g->sigma_blur < g->sigma) { g->nb = 0.f; g->bl = 1.f; }
    g->e_sig = expf(-g->sigma);          /* __expf on the GPU */
    g->e_blur = expf(-g->sigma_blur);
    g->alpha = fminf(0.99f, opac * (g->nb * g->e_sig + g->bl * g->e_blur));
}

#define SET_PROPAGATE_UV (1 << 8)
#define SET_NEAREST (1 << 2)
#define SET_BLUR (1 << 9)
#define SET_NDC (1 << 10)
#define SET_VIS (1 << 15)        /* texture.cu:58  normals flipped towards the camera */
#define SET_ALPHA_VIS (1 << 16)  /* texture.cu:59  hard-edged footprints + outlines */
#define ORC_MAX_C 64

static const float T_NEAR = 0.01f, T_FAR = 1000.0f;

/* ------------------------------------------------------------------------------------------
 * texture_forward: texture.cu:11-329.  All outputs are written for every in-image pixel.
 * Visualisation bits (texture.cu:58-63, :201-241, :269-274): 15 flips normals towards the camera, 16 draws hard-edged
 * footprints (alpha = 0.99 inside sigma <= alpha_bound^2/2, bits 17-21 = alpha_bound*8) with an outline of width
 * outline_bound (bits 26-29 = *4) that is black or, with bit 24, white; bit 25 hides Gaussians of opacity < 0.5.
 * ---------------------------------------------------------------------------------------- */
/* texture_helpers.cuh:390-416 (local_outline) with :355-366 (get_pixel_to_texel) inlined */
static float local_outline(const float *c2w, const float *viewmat, float fx, float fy, float cx, float cy, float px,
                           float py, const pairgeom *g, float glob_scale, int width, float sigma_thresh) {
    float min_dis = 1000.0f;
    const float scale1 = g->s1 * glob_scale, scale2 = g->s2 * glob_scale;
    for (int dx = -width; dx <= width; ++dx)
        for (int dy = -width; dy <= width; ++dy) {
            pixray r = make_ray_at(c2w, viewmat, fx, fy, cx, cy, px + (float)dx, py + (float)dy);
            v3 diff = v3_make(g->mean.x - r.origin.x, g->mean.y - r.origin.y, g->mean.z - r.origin.z);
            float t = v3_dot(g->a3, diff) / plane_denominator(g->a3, r.ray);
            v3 pos = v3_make(r.origin.x + t * r.ray.x, r.origin.y + t * r.ray.y, r.origin.z + t * r.ray.z);
            v3 delta = v3_make(pos.x - g->mean.x, pos.y - g->mean.y, pos.z - g->mean.z);
            float as1 = v3_dot(delta, g->a1) / scale1, as2 = v3_dot(delta, g->a2) / scale2;
            float sigma = 0.5f * (as1 * as1 + as2 * as2);
            if (sigma > sigma_thresh) {
                float cur = sqrtf((float)(dx * dx + dy * dy));
                if (min_dis > cur) min_dis = cur;
            }
        }
    return min_dis;
}

void orc_texture_forward(int img_w, int img_h, int bw, int C, const int32_t *texture_dims,
                         const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
                         const float *opacities, const float *means, const float *scales, float glob_scale,
                         const float *quats, const float *uv0, const float *umap, const float *vmap,
                         const float *texture, const float *viewmat, const float *c2w, float fx, float fy,
                         float cx, float cy, int settings, const float *background, float *out_img,
                         float *out_depth, float *out_reg, float *out_texture, float *out_normal,
                         float *final_Ts, int32_t *final_idx, int32_t *depth_idx, float *out_reg_s) {
    const int tiles_x = (img_w + bw - 1) / bw;
    const int use_blur = (settings & SET_BLUR) != 0, use_ndc = (settings & SET_NDC) != 0;
    const int bilinear = !(settings & SET_NEAREST);
    const int vis_mode = (settings & SET_VIS) != 0, alpha_vis = (settings & SET_ALPHA_VIS) != 0;
    const float alpha_bound = (float)((settings & (0x1f << 17)) >> 17) / 8.0f;
    const float outline_bound = (float)((settings & (0xf << 26)) >> 26) / 4.0f;
    const int accept_opacity_thresh = (settings & (1 << 25)) != 0, white_outline = (settings & (1 << 24)) != 0;
    const float sigma_thresh = 0.5f * alpha_bound * alpha_bound;
#pragma omp parallel for schedule(dynamic, 4)
    for (int pix = 0; pix < img_w * img_h; ++pix) {
        int row = pix / img_w, col = pix % img_w;
        int tile = (row / bw) * tiles_x + (col / bw);
        int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
        pixray r = make_ray(c2w, viewmat, fx, fy, cx, cy, col, row);
        float T = 1.f, acc_c[3] = {0, 0, 0}, acc_n[3] = {0, 0, 0}, acc_t[ORC_MAX_C];
        for (int c = 0; c < C; ++c) acc_t[c] = 0.f;
        float depth = 0.f, reg = 0.f, S0 = 0.f, S1 = 0.f, S2 = 0.f;
        int last = 0, dlast = -1;
        for (int idx = lo; idx < hi; ++idx) {
            int g = gaussian_ids_sorted[idx];
            pairgeom pg;
            pair_geometry(&r, means + 3 * g, scales + 3 * g, quats + 4 * g, opacities[g], glob_scale, viewmat,
                          fx, fy, cx, cy, use_blur, &pg);
            float pixel_dis = 1000.0f;
            if (alpha_vis) { /* texture.cu:201-211 */
                pg.alpha = 0.99f;
                if (pg.sigma > sigma_thresh) pg.alpha = 0.0f;
                if (accept_opacity_thresh && pg.opac < 0.5f) pg.alpha = 0.0f;
                pixel_dis = local_outline(c2w, viewmat, fx, fy, cx, cy, r.px, r.py, &pg, glob_scale, 3, sigma_thresh);
            }
            int skip = (pg.t < T_NEAR || pg.t > T_FAR || pg.alpha < 1.f / 255.f);
            float next_T = T * (1.f - pg.alpha);
            if (next_T <= 1e-4f) break; /* texture.cu:216-221: tested before the skip is honoured */
            if (skip) continue;
            float vis = pg.alpha * T;
            const int draw = !alpha_vis || pixel_dis > outline_bound;   /* texture.cu:226-235 */
            const int white = !draw && white_outline;
            for (int c = 0; c < 3; ++c) {
                if (draw) acc_c[c] += colors[3 * g + c] * vis;
                else if (white) acc_c[c] += vis;
            }
            {
                float sgn = (vis_mode && v3_dot(r.ray, pg.a3) > 0.f) ? -1.f : 1.f; /* texture.cu:236-241 */
                acc_n[0] += vis * (sgn * pg.a3.x); acc_n[1] += vis * (sgn * pg.a3.y); acc_n[2] += vis * (sgn * pg.a3.z);
            }
            float t_view = pg.t * r.view_depth;
            float u = clamp01(uv0[2 * g] + v3_dot(v3_load(umap + 3 * g), pg.delta));
            float v = clamp01(uv0[2 * g + 1] + v3_dot(v3_load(vmap + 3 * g), pg.delta));
            texfetch f;
            texel_setup(texture_dims + 3 * g, u, v, bilinear, C, &f);
            for (int c = 0; c < C; ++c) {
                float val = f.w[0] * texture[f.idx[0] + c] + f.w[1] * texture[f.idx[1] + c] +
                            f.w[2] * texture[f.idx[2] + c] + f.w[3] * texture[f.idx[3] + c];
                if (draw) acc_t[c] += vis * val;   /* texture.cu:269-274 */
                else if (white) acc_t[c] += vis;
            }
            if (T > 0.5f) { depth = t_view; dlast = idx; } /* median depth, texture.cu:286-291 */
            float tv = pg.t;
            if (use_ndc) tv = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view);
            reg += vis * (tv * tv * S0 + S2 - 2.f * tv * S1); /* helpers.cuh:259-264 */
            S0 += vis; S1 += vis * tv; S2 += vis * tv * tv;
            T = next_T;
            last = idx;
        }
        final_Ts[pix] = T; final_idx[pix] = last; depth_idx[pix] = dlast;
        for (int c = 0; c < 3; ++c) {
            out_img[3 * pix + c] = acc_c[c] + T * background[c];
            out_normal[3 * pix + c] = acc_n[c];
        }
        out_depth[pix] = depth; out_reg[pix] = reg;
        out_reg_s[3 * pix] = S0; out_reg_s[3 * pix + 1] = S1; out_reg_s[3 * pix + 2] = S2;
        for (int c = 0; c < C; ++c) out_texture[C * pix + c] = acc_t[c];
    }
}

static inline void atomic_addf(float *p, float v) {
#pragma omp atomic
    *p += v;
}

/* ------------------------------------------------------------------------------------------
 * texture_backward: texture.cu:331-760.  All v_* outputs must be zero-initialised.
 * ---------------------------------------------------------------------------------------- */
void orc_texture_backward(int img_w, int img_h, int bw, int C, const int32_t *texture_dims,
                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *colors,
                          const float *opacities, const float *means, const float *scales, float glob_scale,
                          const float *quats, const float *uv0, const float *umap, const float *vmap,
                          const float *texture, const float *viewmat, const float *c2w, float fx, float fy,
                          float cx, float cy, int settings, const float *background, const float *final_Ts,
                          const int32_t *final_idx, const int32_t *depth_idx, const float *final_s,
                          const float *v_out_img, const float *v_out_depth, const float *v_out_reg,
                          const float *v_out_alpha, const float *v_out_texture, const float *v_out_normal,
                          float *v_colors, float *v_opacity, float *v_means, float *v_scales, float *v_quats,
                          float *v_uv0, float *v_umap, float *v_vmap, float *v_texture) {
    const int tiles_x = (img_w + bw - 1) / bw;
    const int use_blur = (settings & SET_BLUR) != 0, use_ndc = (settings & SET_NDC) != 0;
    const int bilinear = !(settings & SET_NEAREST), prop_uv = (settings & SET_PROPAGATE_UV) != 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (int pix = 0; pix < img_w * img_h; ++pix) {
        int row = pix / img_w, col = pix % img_w;
        int tile = (row / bw) * tiles_x + (col / bw);
        int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
        pixray r = make_ray(c2w, viewmat, fx, fy, cx, cy, col, row);
        float T = final_Ts[pix];
        const float S0 = final_s[3 * pix], S1 = final_s[3 * pix + 1], S2 = final_s[3 * pix + 2];
        const int dfinal = depth_idx[pix], bfinal = final_idx[pix];
        v3 vo = v3_load(v_out_img + 3 * pix), vn = v3_load(v_out_normal + 3 * pix);
        const float vod = v_out_depth[pix], vor = v_out_reg[pix];
        const float *vot = v_out_texture + C * pix;
        float v_T_run = v3_dot(v3_load(background), vo) - v_out_alpha[pix]; /* texture.cu:449 */
        for (int idx = (bfinal < hi - 1 ? bfinal : hi - 1); idx >= lo; --idx) {
            int g = gaussian_ids_sorted[idx];
            pairgeom pg;
            pair_geometry(&r, means + 3 * g, scales + 3 * g, quats + 4 * g, opacities[g], glob_scale, viewmat,
                          fx, fy, cx, cy, use_blur, &pg);
            if (pg.t < T_NEAR || pg.t > T_FAR || pg.alpha < 1.f / 255.f) continue;
            T *= 1.f / (1.f - pg.alpha);
            float vis = pg.alpha * T;
            for (int c = 0; c < 3; ++c) atomic_addf(v_colors + 3 * g + c, vis * (&vo.x)[c]);
            v3 g_n = v3_make(vis * vn.x, vis * vn.y, vis * vn.z); /* direct normal term */
            float v_vis = v3_dot(v3_load(colors + 3 * g), vo) + v3_dot(pg.a3, vn);

            /* texture fetch VJP: texture.cu:594-642, texture_helpers.cuh:252-300 */
            v3 um = v3_load(umap + 3 * g), vm = v3_load(vmap + 3 * g);
            float u = clamp01(uv0[2 * g] + v3_dot(um, pg.delta));
            float v = clamp01(uv0[2 * g + 1] + v3_dot(vm, pg.delta));
            texfetch f;
            texel_setup(texture_dims + 3 * g, u, v, bilinear, C, &f);
            float v_u = 0.f, v_v = 0.f;
            for (int c = 0; c < C; ++c) {
                float c00 = texture[f.idx[0] + c], c01 = texture[f.idx[1] + c], c10 = texture[f.idx[2] + c],
                      c11 = texture[f.idx[3] + c];
                float val = f.w[0] * c00 + f.w[1] * c01 + f.w[2] * c10 + f.w[3] * c11;
                float v_val = vis * vot[c];
                for (int k = 0; k < 4; ++k) atomic_addf(v_texture + f.idx[k] + c, f.w[k] * v_val);
                if (bilinear && prop_uv) {
                    v_u += f.h * (v_val * (-(1.f - f.fv) * c00 - f.fv * c01 + (1.f - f.fv) * c10 + f.fv * c11));
                    v_v += f.wd * (v_val * (-(1.f - f.fu) * c00 + (1.f - f.fu) * c01 - f.fu * c10 + f.fu * c11));
                }
                v_vis += val * vot[c];
            }
            /* get_uv_vjp: texture_helpers.cuh:62-80 (always called; v_u,v_v are zero unless bit 8) */
            atomic_addf(v_uv0 + 2 * g, v_u); atomic_addf(v_uv0 + 2 * g + 1, v_v);
            v3 v_delta = v3_make(0, 0, 0);
            for (int c = 0; c < 3; ++c) {
                atomic_addf(v_umap + 3 * g + c, (&pg.delta.x)[c] * v_u);
                atomic_addf(v_vmap + 3 * g + c, (&pg.delta.x)[c] * v_v);
            }
            if (prop_uv) v_delta = v3_make(um.x * v_u + vm.x * v_v, um.y * v_u + vm.y * v_v, um.z * v_u + vm.z * v_v);

            /* alpha / transmittance recurrences: texture.cu:650-670 */
            float v_alpha = T * v_vis - T * v_T_run;
            float v_T_cur = pg.alpha * v_vis + (1.f - pg.alpha) * v_T_run;
            float t_view = pg.t * r.view_depth;
            float t_ndc = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view);
            float tv = use_ndc ? t_ndc : pg.t;
            float v_tv = 2.f * (vis * tv * S0 - vis * S1) * vor;       /* helpers.cuh:266-269, FINAL sums */
            float v_w = (tv * tv * S0 - 2.f * tv * S1 + S2) * vor;
            v_alpha += v_w * T;
            v_T_cur += v_w * pg.alpha;
            v_T_run = v_T_cur;
            float v_t = use_ndc ? 0.f : v_tv;
            float v_tndc = use_ndc ? v_tv : 0.f;

            /* sigma -> scales, l1, l2 : texture.cu:535-538, 672-675 */
            float sf = 0.5f / (glob_scale * glob_scale);
            float q1 = pg.l1 * pg.l1 / (pg.s1 * pg.s1), q2 = pg.l2 * pg.l2 / (pg.s2 * pg.s2);
            float v_sigma = -pg.nb * pg.opac * pg.e_sig * v_alpha;
            atomic_addf(v_scales + 3 * g, -2.f * sf * q1 * v_sigma / pg.s1);
            atomic_addf(v_scales + 3 * g + 1, -2.f * sf * q2 * v_sigma / pg.s2);
            float v_l1 = 2.f * sf * pg.l1 * v_sigma / (pg.s1 * pg.s1);
            float v_l2 = 2.f * sf * pg.l2 * v_sigma / (pg.s2 * pg.s2);
            v_t += v_l1 * v3_dot(r.ray, pg.a1) + v_l2 * v3_dot(r.ray, pg.a2);
            v_t += v3_dot(r.ray, v_delta);
            float v_tview = 0.f;
            if (idx == dfinal && dfinal != -1) v_tview += vod; /* texture.cu:678-680 */
            v_tview += (T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view * t_view) * v_tndc;
            v_t += r.view_depth * v_tview;

            /* blur branch sends its gradient to the mean through the pinhole: texture.cu:683-692 */
            float v_sblur = -pg.bl * pg.opac * pg.e_blur * v_alpha;
            v3 v_pv = pinhole_vjp(fx, fy, pg.pview, 2.0f * v_sblur * (pg.mx - r.px), 2.0f * v_sblur * (pg.my - r.py));
            v3 v_mean_blur = xform_rot_t(viewmat, v_pv);

            /* ray-plane t VJP: texture_helpers.cuh:315-334 */
            float den = plane_denominator(pg.a3, r.ray);
            v3 diff = v3_make(pg.mean.x - r.origin.x, pg.mean.y - r.origin.y, pg.mean.z - r.origin.z);
            v3 v_mean_t = v3_make(pg.a3.x * v_t / den, pg.a3.y * v_t / den, pg.a3.z * v_t / den);
            v3 v_n_t = v3_make((diff.x - r.ray.x * pg.t) * v_t / den, (diff.y - r.ray.y * pg.t) * v_t / den,
                               (diff.z - r.ray.z * pg.t) * v_t / den);

            atomic_addf(v_means + 3 * g, -(pg.a1.x * v_l1 + pg.a2.x * v_l2) + v_mean_t.x + v_mean_blur.x - v_delta.x);
            atomic_addf(v_means + 3 * g + 1, -(pg.a1.y * v_l1 + pg.a2.y * v_l2) + v_mean_t.y + v_mean_blur.y - v_delta.y);
            atomic_addf(v_means + 3 * g + 2, -(pg.a1.z * v_l1 + pg.a2.z * v_l2) + v_mean_t.z + v_mean_blur.z - v_delta.z);
            v3 g1 = v3_make(pg.delta.x * v_l1, pg.delta.y * v_l1, pg.delta.z * v_l1);
            v3 g2 = v3_make(pg.delta.x * v_l2, pg.delta.y * v_l2, pg.delta.z * v_l2);
            v3 g3 = v3_make(v_n_t.x + g_n.x, v_n_t.y + g_n.y, v_n_t.z + g_n.z);
            float vq[4];
            surfel_axes_vjp(quats + 4 * g, g1, g2, g3, vq);
            for (int c = 0; c < 4; ++c) atomic_addf(v_quats + 4 * g + c, vq[c]);
            atomic_addf(v_opacity + g, (pg.nb * pg.e_sig + pg.bl * pg.e_blur) * v_alpha);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * texture_edit: texture_edit.cu:11-236.  Walks every pixel's list exactly like the forward
 * pass (same alpha, skip and stop rules, :161-186) and, for every blended Gaussian whose view depth lies inside
 * the pixel's [depth_lower, depth_upper] window (:191), splats five values with the BILINEAR weights of the
 * intersection's texel coordinate into `updated_texture` (X, C) -- channel 0-2 rgb*alpha of the edit canvas,
 * 3 its alpha, 4 the constant 1 (:204-227; texture_helpers.cuh:239-250).  The splat is NOT weighted by the blend
 * weight `vis` (computed at :189 but unused).  C = texture_info.z is the row pitch of the output and must be >= 5.
 * NB the settings bits differ from the rasteriser's: bit 0 = blur, bit 1 = ndc (unused) (:46-47).
 * ---------------------------------------------------------------------------------------- */
void orc_texture_edit(int img_w, int img_h, int bw, int C, const int32_t *texture_dims, const float *updated_img,
                      const float *updated_alpha, const float *depth_lower, const float *depth_upper,
                      const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const float *opacities,
                      const float *means, const float *scales, float glob_scale, const float *quats,
                      const float *uv0, const float *umap, const float *vmap, const float *viewmat,
                      const float *c2w, float fx, float fy, float cx, float cy, int settings,
                      float *updated_texture) {
    const int tiles_x = (img_w + bw - 1) / bw;
    const int use_blur = (settings & 1) != 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (int pix = 0; pix < img_w * img_h; ++pix) {
        int row = pix / img_w, col = pix % img_w;
        int tile = (row / bw) * tiles_x + (col / bw);
        int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
        pixray r = make_ray(c2w, viewmat, fx, fy, cx, cy, col, row);
        const float a_upd = updated_alpha[pix];
        const float vals[5] = {updated_img[3 * pix] * a_upd, updated_img[3 * pix + 1] * a_upd,
                               updated_img[3 * pix + 2] * a_upd, a_upd, 1.f};
        const float zlo = depth_lower[pix], zhi = depth_upper[pix];
        float T = 1.f;
        for (int idx = lo; idx < hi; ++idx) {
            int g = gaussian_ids_sorted[idx];
            pairgeom pg;
            pair_geometry(&r, means + 3 * g, scales + 3 * g, quats + 4 * g, opacities[g], glob_scale, viewmat,
                          fx, fy, cx, cy, use_blur, &pg);
            int skip = (pg.t < T_NEAR || pg.t > T_FAR || pg.alpha < 1.f / 255.f);
            float next_T = T * (1.f - pg.alpha);
            if (next_T <= 1e-4f) break; /* texture_edit.cu:180-184 */
            if (skip) continue;
            float t_view = pg.t * r.view_depth;
            if (t_view >= zlo && t_view <= zhi) {
                float u = clamp01(uv0[2 * g] + v3_dot(v3_load(umap + 3 * g), pg.delta));
                float v = clamp01(uv0[2 * g + 1] + v3_dot(v3_load(vmap + 3 * g), pg.delta));
                texfetch f;
                texel_setup(texture_dims + 3 * g, u, v, 1, C, &f);
                for (int k = 0; k < 5; ++k)
                    for (int c4 = 0; c4 < 4; ++c4) atomic_addf(updated_texture + f.idx[c4] + k, f.w[c4] * vals[k]);
            }
            T = next_T;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Training-step glue.
 *
 * orc_preprocess_forward: example.py:126-143 (+ the sigmoids of :162-163).  raw_rgbs may be NULL.
 * ---------------------------------------------------------------------------------------- */
static inline float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

void orc_preprocess_forward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                            const float *raw_rgbs, const float *raw_opac, float *scales, float *quats, float *uv0,
                            float *umap, float *vmap, float *colors, float *opacities) {
    for (int g = 0; g < n; ++g) {
        float s1 = expf(raw_scales[3 * g]), s2 = expf(raw_scales[3 * g + 1]);       /* :126-127 */
        scales[3 * g] = s1; scales[3 * g + 1] = s2;
        scales[3 * g + 2] = 1e-5f * (0.5f * (s1 + s2));                              /* :128 */
        const float *q = raw_quats + 4 * g;
        float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);    /* :129 */
        float qn[4] = {q[0] / nrm, q[1] / nrm, q[2] / nrm, q[3] / nrm};
        for (int c = 0; c < 4; ++c) quats[4 * g + c] = qn[c];
        v3 a1, a2, a3;
        surfel_axes(qn, &a1, &a2, &a3);                                              /* :130, Rs[:, :, 0], Rs[:, :, 1] */
        const float *m = mapping + 4 * g;
        float us = expf(m[2]), c = cosf(m[3]), s = sinf(m[3]);                       /* :132-133 */
        uv0[2 * g] = m[0]; uv0[2 * g + 1] = m[1];                                    /* :131 */
        umap[3 * g] = us * (a1.x * c + a2.x * s); umap[3 * g + 1] = us * (a1.y * c + a2.y * s);
        umap[3 * g + 2] = us * (a1.z * c + a2.z * s);                                /* :136 */
        vmap[3 * g] = us * (-a1.x * s + a2.x * c); vmap[3 * g + 1] = us * (-a1.y * s + a2.y * c);
        vmap[3 * g + 2] = us * (-a1.z * s + a2.z * c);                               /* :137 */
        if (raw_rgbs) for (int k = 0; k < 3; ++k) colors[3 * g + k] = sigmoidf_(raw_rgbs[3 * g + k]);
        opacities[g] = sigmoidf_(raw_opac[g]);
    }
}

/* The VJP torch autograd computes for the block above (no reference source: it is autograd's; pinned by the
 * golden vectors of tests/golden/make_golden_train_ops.py, which run torch autograd over the reference's own
 * normalized_quat_to_rotmat). */
void orc_preprocess_backward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                             const float *raw_rgbs, const float *raw_opac, const float *v_scales,
                             const float *v_quats, const float *v_uv0, const float *v_umap, const float *v_vmap,
                             const float *v_colors, const float *v_opacity, float *v_raw_scales, float *v_raw_quats,
                             float *v_mapping, float *v_raw_rgbs, float *v_raw_opac) {
    for (int g = 0; g < n; ++g) {
        v_raw_scales[3 * g] = v_scales[3 * g] * expf(raw_scales[3 * g]);
        v_raw_scales[3 * g + 1] = v_scales[3 * g + 1] * expf(raw_scales[3 * g + 1]);
        v_raw_scales[3 * g + 2] = 0.f; /* detached thickness, example.py:128 */
        const float *q = raw_quats + 4 * g;
        float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        float qn[4] = {q[0] / nrm, q[1] / nrm, q[2] / nrm, q[3] / nrm};
        v3 a1, a2, a3;
        surfel_axes(qn, &a1, &a2, &a3);
        const float *m = mapping + 4 * g;
        float us = expf(m[2]), c = cosf(m[3]), s = sinf(m[3]);
        v3 um = v3_make(us * (a1.x * c + a2.x * s), us * (a1.y * c + a2.y * s), us * (a1.z * c + a2.z * s));
        v3 vm = v3_make(us * (-a1.x * s + a2.x * c), us * (-a1.y * s + a2.y * c), us * (-a1.z * s + a2.z * c));
        v3 gu = v3_load(v_umap + 3 * g), gv = v3_load(v_vmap + 3 * g);
        v_mapping[4 * g] = v_uv0[2 * g]; v_mapping[4 * g + 1] = v_uv0[2 * g + 1];
        v_mapping[4 * g + 2] = v3_dot(um, gu) + v3_dot(vm, gv);
        v_mapping[4 * g + 3] = v3_dot(vm, gu) - v3_dot(um, gv);
        v3 g1 = v3_make(us * (c * gu.x - s * gv.x), us * (c * gu.y - s * gv.y), us * (c * gu.z - s * gv.z));
        v3 g2 = v3_make(us * (s * gu.x + c * gv.x), us * (s * gu.y + c * gv.y), us * (s * gu.z + c * gv.z));
        float va[4];
        surfel_axes_vjp(qn, g1, g2, v3_make(0.f, 0.f, 0.f), va);
        float vq[4], d = 0.f;
        for (int k = 0; k < 4; ++k) { vq[k] = v_quats[4 * g + k] + va[k]; d += qn[k] * vq[k]; }
        for (int k = 0; k < 4; ++k) v_raw_quats[4 * g + k] = (vq[k] - qn[k] * d) / nrm;
        if (raw_rgbs)
            for (int k = 0; k < 3; ++k) {
                float cv = sigmoidf_(raw_rgbs[3 * g + k]);
                v_raw_rgbs[3 * g + k] = v_colors[3 * g + k] * cv * (1.f - cv);
            }
        float o = sigmoidf_(raw_opac[g]);
        v_raw_opac[g] = v_opacity[g] * o * (1.f - o);
    }
}

/* torch.optim.Adam, single-tensor form (torch/optim/adam.py, pinned version torch 2.11.0; third-party to the
 * reference, call site example.py:223-225, :278): defaults, no weight decay, no amsgrad.  step is 1-based. */
void orc_adam_step(int64_t count, float *p, const float *g, float *m, float *v, double lr, double beta1,
                   double beta2, double eps_d, int step, float grad_scale) {
    /* torch forms these in Python doubles before the fp32 kernels see them */
    double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    float step_size = (float)(lr / bc1), sq2 = (float)sqrt(bc2), eps = (float)eps_d;
    float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), b2 = (float)beta2;
    for (int64_t i = 0; i < count; ++i) {
        float gi = g[i] * grad_scale;
        m[i] = m[i] + (gi - m[i]) * omb1;
        v[i] = b2 * v[i] + omb2 * gi * gi;
        float denom = sqrtf(v[i]) / sq2 + eps;
        p[i] = p[i] - step_size * (m[i] / denom);
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
