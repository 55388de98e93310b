// sh.cu -- spherical-harmonics colour evaluation, degrees 0-4, 3 channels (SURVEY 8a row a-11;
// reference sh.cuh:46-253, bindings.cu:18-75).  HBM bound: 12 + 12K bytes in, 12 out per Gaussian.
// One thread per Gaussian; the basis is evaluated once and shared by the three channels.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace gstex {

__host__ __device__ inline int sh_num_bases(int degree) {
    return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;
}

// Y[0..num_bases(deg)) for the normalised direction (reference normalises inside, sh.cuh:61-66)
__device__ __forceinline__ void sh_basis(int deg, Vec3 dir, float *Y) {
    Y[0] = 0.28209479177387814f;
    if (deg < 1) return;
    const float nrm = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    const float x = dir.x / nrm, y = dir.y / nrm, z = dir.z / nrm;
    const float c1 = 0.4886025119029199f;
    Y[1] = -c1 * y;
    Y[2] = c1 * z;
    Y[3] = -c1 * x;
    if (deg < 2) return;
    const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
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
    Y[17] = -1.7701307697799304f * yz * (3.f * xx - yy);
    Y[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f);
    Y[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
    Y[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
    Y[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f);
    Y[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
    Y[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
    Y[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

// ------------------------------------------------------------------------------------------
// Tiled kernels.  One CTA = SH_ROWS Gaussians = one CONTIGUOUS block of SH_ROWS * L floats (L = 3K) of the
// coefficient array.  The block is moved between global and shared memory with fully coalesced 16-byte
// accesses and each thread then works on its own row of the tile (row pitch L|1 floats: odd, so the 32 lanes
// of a warp hit 32 different banks).  A thread-per-Gaussian kernel reading its 192-byte row directly runs at
// ~1/6 of HBM bandwidth (measured: 0.30 ms for the 192 MB backward write at N = 1M, degree 3).
// ------------------------------------------------------------------------------------------
constexpr int SH_ROWS = 128;

__host__ __device__ inline int sh_pitch(int L) { return L | 1; }
// Row pitch (floats) of the forward tile when rows are whole float4s (L % 4 == 0): a multiple of 4 floats, so that rows
// can be filled with 16-byte cp.async copies, and an ODD number of quads, so that the 8 lanes of a quarter warp reading
// quad q of 8 consecutive rows (LDS.128) touch 8 different 16-byte bank groups.
__host__ __device__ inline int sh_pitch4(int L) { return ((L >> 2) | 1) << 2; }

__device__ __forceinline__ void sh_tile_load(float *__restrict__ tile, int pitch, const float *__restrict__ g,
                                             int rows, int L, bool vec4) {
    const int total = rows * L;
    if (vec4) {
        const float4 *__restrict__ g4 = reinterpret_cast<const float4 *>(g);
        for (int q = threadIdx.x; q < (total >> 2); q += blockDim.x) {
            const float4 v = g4[q];
            const int e = q << 2, r = e / L, c = e - r * L;  // L % 4 == 0: the four floats share a row
            float *__restrict__ d = tile + r * pitch + c;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int r = e / L, c = e - r * L;
            tile[r * pitch + c] = g[e];
        }
    }
}

__device__ __forceinline__ void sh_tile_store(const float *__restrict__ tile, int pitch, float *__restrict__ g,
                                              int rows, int L, bool vec4, int accumulate) {
    const int total = rows * L;
    if (vec4) {
        float4 *__restrict__ g4 = reinterpret_cast<float4 *>(g);
        for (int q = threadIdx.x; q < (total >> 2); q += blockDim.x) {
            const int e = q << 2, r = e / L, c = e - r * L;
            const float *__restrict__ d = tile + r * pitch + c;
            float4 v = make_float4(d[0], d[1], d[2], d[3]);
            if (accumulate) {
                const float4 old = g4[q];
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            g4[q] = v;
        }
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int r = e / L, c = e - r * L;
            const float v = tile[r * pitch + c];
            g[e] = accumulate ? g[e] + v : v;
        }
    }
}

// FUSED = false: colours = SH(viewdirs) . coeffs                       (reference sh.cuh:212-232)
// FUSED = true : dirs = means - camera origin; colours = clamp(SH + 0.5, 0, 1); mask bit c = channel c not clamped
//                (the torch glue around the SH op in GStex-style trainers)
template <int DEG, bool FUSED>
__global__ void __launch_bounds__(SH_ROWS) sh_forward_tiled(int n, int K, const float *__restrict__ dirs_or_means,
                                                            const float *__restrict__ c2w,
                                                            const float *__restrict__ coeffs,
                                                            float *__restrict__ colors, uint8_t *__restrict__ mask,
                                                            int vec4) {
    extern __shared__ __align__(16) float sh_tile[];
    constexpr int KU = (DEG + 1) * (DEG + 1);
    const int L = 3 * K, pitch = vec4 ? sh_pitch4(L) : sh_pitch(L);
    const int row0 = blockIdx.x * SH_ROWS, rows = min(SH_ROWS, n - row0);
    if (vec4) {  // the block's rows * L floats go to shared memory as 16-byte asynchronous copies (no register staging)
        const float4 *__restrict__ g4 = reinterpret_cast<const float4 *>(coeffs + (size_t)row0 * L);
        const int qpr = L >> 2;
        for (int q = threadIdx.x; q < rows * qpr; q += blockDim.x) {
            const int r = q / qpr, c = q - r * qpr;
            __pipeline_memcpy_async(sh_tile + r * pitch + 4 * c, g4 + q, 16);
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
    } else {
        sh_tile_load(sh_tile, pitch, coeffs + (size_t)row0 * L, rows, L, false);
    }
    __syncthreads();
    const int t = threadIdx.x, i = row0 + t;
    if (t >= rows) return;
    Vec3 dir = ld3(dirs_or_means + 3 * (size_t)i);
    if (FUSED) dir = mk3(dir.x - c2w[3], dir.y - c2w[7], dir.z - c2w[11]);
    float Y[KU];
    sh_basis(DEG, dir, Y);
    const float *__restrict__ c = sh_tile + t * pitch;
    float v[3];
    v[0] = v[1] = v[2] = FUSED ? 0.5f : 0.f;
    if (vec4) {  // the same sums in the same order, the row read as float4s
        const float4 *__restrict__ c4 = reinterpret_cast<const float4 *>(c);
#pragma unroll
        for (int q = 0; q < (3 * KU + 3) / 4; ++q) {
            const float4 w4 = c4[q];
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * q + e;
                if (j < 3 * KU) v[j % 3] = fmaf(Y[j / 3], w[e], v[j % 3]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            v[0] = fmaf(Y[k], c[3 * k], v[0]);
            v[1] = fmaf(Y[k], c[3 * k + 1], v[1]);
            v[2] = fmaf(Y[k], c[3 * k + 2], v[2]);
        }
    }
    if (FUSED) {
        unsigned m = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            if (v[ch] > 0.f && v[ch] < 1.f) m |= 1u << ch;
            v[ch] = fminf(fmaxf(v[ch], 0.f), 1.f);
        }
        mask[i] = (uint8_t)m;
    }
    colors[3 * (size_t)i] = v[0];
    colors[3 * (size_t)i + 1] = v[1];
    colors[3 * (size_t)i + 2] = v[2];
}

// v_coeffs[b, c] = Y_b * v_colors[c] for b < num_bases(degrees_to_use), 0 beyond (reference sh.cuh:234-253; no
// gradient to the directions).  FUSED gates v_colors with the clamp mask of the forward pass.
template <int DEG, bool FUSED>
__global__ void __launch_bounds__(SH_ROWS) sh_backward_tiled(int n, int K, const float *__restrict__ dirs_or_means,
                                                             const float *__restrict__ c2w,
                                                             const float *__restrict__ v_colors,
                                                             const uint8_t *__restrict__ mask,
                                                             float *__restrict__ v_coeffs, int accumulate, int vec4) {
    extern __shared__ __align__(16) float sh_tile[];
    constexpr int KU = (DEG + 1) * (DEG + 1);
    const int L = 3 * K, pitch = sh_pitch(L);
    const int row0 = blockIdx.x * SH_ROWS, rows = min(SH_ROWS, n - row0);
    const int t = threadIdx.x, i = row0 + t;
    if (t < rows) {
        Vec3 dir = ld3(dirs_or_means + 3 * (size_t)i);
        if (FUSED) dir = mk3(dir.x - c2w[3], dir.y - c2w[7], dir.z - c2w[11]);
        float Y[KU];
        sh_basis(DEG, dir, Y);
        float vr = v_colors[3 * (size_t)i], vg = v_colors[3 * (size_t)i + 1], vb = v_colors[3 * (size_t)i + 2];
        if (FUSED) {
            const unsigned m = mask[i];
            vr = (m & 1u) ? vr : 0.f;
            vg = (m & 2u) ? vg : 0.f;
            vb = (m & 4u) ? vb : 0.f;
        }
        float *__restrict__ c = sh_tile + t * pitch;
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            c[3 * k] = Y[k] * vr;
            c[3 * k + 1] = Y[k] * vg;
            c[3 * k + 2] = Y[k] * vb;
        }
        for (int k = KU * 3; k < L; ++k) c[k] = 0.f;  // rows beyond degrees_to_use (torch::zeros upstream)
    }
    __syncthreads();
    sh_tile_store(sh_tile, pitch, v_coeffs + (size_t)row0 * L, rows, L, vec4 != 0, accumulate);
}

template <bool FUSED>
static int launch_sh_forward(int n, int degree, int degrees_to_use, const float *dirs_or_means, const float *c2w,
                             const float *coeffs, float *colors, uint8_t *mask, cudaStream_t s) {
    const int K = sh_num_bases(degree), L = 3 * K, grid = ceil_div(n, SH_ROWS);
    const int vec4 = (L % 4 == 0) && ((uintptr_t)coeffs % 16 == 0);
    const size_t smem = sizeof(float) * SH_ROWS * (vec4 ? sh_pitch4(L) : sh_pitch(L));
    switch (degrees_to_use) {
        case 0: sh_forward_tiled<0, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, coeffs, colors, mask, vec4); break;
        case 1: sh_forward_tiled<1, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, coeffs, colors, mask, vec4); break;
        case 2: sh_forward_tiled<2, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, coeffs, colors, mask, vec4); break;
        case 3: sh_forward_tiled<3, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, coeffs, colors, mask, vec4); break;
        default: sh_forward_tiled<4, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, coeffs, colors, mask, vec4); break;
    }
    GSTEX_LAUNCH_OK("sh_forward_tiled");
    return GSTEX_OK;
}

template <bool FUSED>
static int launch_sh_backward(int n, int degree, int degrees_to_use, const float *dirs_or_means, const float *c2w,
                              const float *v_colors, const uint8_t *mask, float *v_coeffs, int accumulate,
                              cudaStream_t s) {
    const int K = sh_num_bases(degree), L = 3 * K, grid = ceil_div(n, SH_ROWS);
    const size_t smem = sizeof(float) * SH_ROWS * sh_pitch(L);
    const int vec4 = (L % 4 == 0) && ((uintptr_t)v_coeffs % 16 == 0);
    switch (degrees_to_use) {
        case 0: sh_backward_tiled<0, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, v_colors, mask, v_coeffs, accumulate, vec4); break;
        case 1: sh_backward_tiled<1, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, v_colors, mask, v_coeffs, accumulate, vec4); break;
        case 2: sh_backward_tiled<2, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, v_colors, mask, v_coeffs, accumulate, vec4); break;
        case 3: sh_backward_tiled<3, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, v_colors, mask, v_coeffs, accumulate, vec4); break;
        default: sh_backward_tiled<4, FUSED><<<grid, SH_ROWS, smem, s>>>(n, K, dirs_or_means, c2w, v_colors, mask, v_coeffs, accumulate, vec4); break;
    }
    GSTEX_LAUNCH_OK("sh_backward_tiled");
    return GSTEX_OK;
}

}  // namespace gstex

using namespace gstex;

static int check_sh(const char *who, int n, int degree, int degrees_to_use) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "%s: n = %d", who, n);
    GSTEX_REQUIRE(degree >= 0 && degree <= 4, GSTEX_E_INVALID, "%s: degree %d not in [0,4]", who, degree);
    GSTEX_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= degree, GSTEX_E_INVALID,
                  "%s: degrees_to_use %d not in [0, degree=%d]", who, degrees_to_use, degree);
    return GSTEX_OK;
}

extern "C" int gstex_sh_forward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *coeffs,
                                float *colors, gstex_stream_t stream) {
    int rc = check_sh("sh_forward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    return launch_sh_forward<false>(n, degree, degrees_to_use, viewdirs, nullptr, coeffs, colors, nullptr, as_stream(stream));
}

extern "C" int gstex_sh_backward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *v_colors,
                                 float *v_coeffs, int accumulate, gstex_stream_t stream) {
    int rc = check_sh("sh_backward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    return launch_sh_backward<false>(n, degree, degrees_to_use, viewdirs, nullptr, v_colors, nullptr, v_coeffs, accumulate, as_stream(stream));
}

extern "C" int gstex_sh_colors_forward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                                       const float *coeffs, float *colors, uint8_t *mask, gstex_stream_t stream) {
    int rc = check_sh("sh_colors_forward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    return launch_sh_forward<true>(n, degree, degrees_to_use, means, c2w, coeffs, colors, mask, as_stream(stream));
}

extern "C" int gstex_sh_colors_backward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                                        const float *v_colors, const uint8_t *mask, float *v_coeffs, int accumulate,
                                        gstex_stream_t stream) {
    int rc = check_sh("sh_colors_backward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    return launch_sh_backward<true>(n, degree, degrees_to_use, means, c2w, v_colors, mask, v_coeffs, accumulate, as_stream(stream));
}
