// sh.cu -- spherical-harmonics colour evaluation, degrees 0-4, 3 channels (SURVEY 8a row a-11;
// reference sh.cuh:46-253, bindings.cu:18-75).  HBM bound: 12 + 12K bytes in, 12 out per Gaussian.
// One thread per Gaussian; the basis is evaluated once and shared by the three channels.
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

template <int DEG>
__global__ void __launch_bounds__(256) sh_forward_kernel(int n, int K, const float *__restrict__ viewdirs,
                                                         const float *__restrict__ coeffs,
                                                         float *__restrict__ colors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int KU = (DEG + 1) * (DEG + 1);
    float Y[KU];
    sh_basis(DEG, ld3(viewdirs + 3 * i), Y);
    const float *__restrict__ c = coeffs + (size_t)i * K * 3;
    float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        r = fmaf(Y[k], c[3 * k], r);
        g = fmaf(Y[k], c[3 * k + 1], g);
        b = fmaf(Y[k], c[3 * k + 2], b);
    }
    colors[3 * i] = r;
    colors[3 * i + 1] = g;
    colors[3 * i + 2] = b;
}

template <int DEG>
__global__ void __launch_bounds__(256) sh_backward_kernel(int n, int K, const float *__restrict__ viewdirs,
                                                          const float *__restrict__ v_colors,
                                                          float *__restrict__ v_coeffs, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int KU = (DEG + 1) * (DEG + 1);
    float Y[KU];
    sh_basis(DEG, ld3(viewdirs + 3 * i), Y);
    const float vr = v_colors[3 * i], vg = v_colors[3 * i + 1], vb = v_colors[3 * i + 2];
    float *__restrict__ o = v_coeffs + (size_t)i * K * 3;
    if (accumulate) {
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            o[3 * k] += Y[k] * vr;
            o[3 * k + 1] += Y[k] * vg;
            o[3 * k + 2] += Y[k] * vb;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            o[3 * k] = Y[k] * vr;
            o[3 * k + 1] = Y[k] * vg;
            o[3 * k + 2] = Y[k] * vb;
        }
        for (int k = KU * 3; k < K * 3; ++k) o[k] = 0.f;  // rows beyond degrees_to_use (torch::zeros upstream)
    }
}

// Fused view-dependent colour: colours = clamp(SH(mean - camera origin) + 0.5, 0, 1)  (the torch glue around
// the SH op in GStex-style trainers).  mask bit c = channel c was not clamped.
template <int DEG>
__global__ void __launch_bounds__(256) sh_colors_forward_kernel(int n, int K, const float *__restrict__ means,
                                                                const float *__restrict__ c2w,
                                                                const float *__restrict__ coeffs,
                                                                float *__restrict__ colors,
                                                                uint8_t *__restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int KU = (DEG + 1) * (DEG + 1);
    float Y[KU];
    const Vec3 dir = mk3(means[3 * i] - c2w[3], means[3 * i + 1] - c2w[7], means[3 * i + 2] - c2w[11]);
    sh_basis(DEG, dir, Y);
    const float *__restrict__ c = coeffs + (size_t)i * K * 3;
    float v[3] = {0.5f, 0.5f, 0.5f};
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        v[0] = fmaf(Y[k], c[3 * k], v[0]);
        v[1] = fmaf(Y[k], c[3 * k + 1], v[1]);
        v[2] = fmaf(Y[k], c[3 * k + 2], v[2]);
    }
    unsigned m = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (v[ch] > 0.f && v[ch] < 1.f) m |= 1u << ch;
        colors[3 * i + ch] = fminf(fmaxf(v[ch], 0.f), 1.f);
    }
    mask[i] = (uint8_t)m;
}

template <int DEG>
__global__ void __launch_bounds__(256) sh_colors_backward_kernel(int n, int K, const float *__restrict__ means,
                                                                 const float *__restrict__ c2w,
                                                                 const float *__restrict__ v_colors,
                                                                 const uint8_t *__restrict__ mask,
                                                                 float *__restrict__ v_coeffs, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int KU = (DEG + 1) * (DEG + 1);
    float Y[KU];
    const Vec3 dir = mk3(means[3 * i] - c2w[3], means[3 * i + 1] - c2w[7], means[3 * i + 2] - c2w[11]);
    sh_basis(DEG, dir, Y);
    const unsigned m = mask[i];
    const float vr = (m & 1u) ? v_colors[3 * i] : 0.f, vg = (m & 2u) ? v_colors[3 * i + 1] : 0.f;
    const float vb = (m & 4u) ? v_colors[3 * i + 2] : 0.f;
    float *__restrict__ o = v_coeffs + (size_t)i * K * 3;
    if (accumulate) {
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            o[3 * k] += Y[k] * vr;
            o[3 * k + 1] += Y[k] * vg;
            o[3 * k + 2] += Y[k] * vb;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            o[3 * k] = Y[k] * vr;
            o[3 * k + 1] = Y[k] * vg;
            o[3 * k + 2] = Y[k] * vb;
        }
        for (int k = KU * 3; k < K * 3; ++k) o[k] = 0.f;
    }
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
    const int K = sh_num_bases(degree), grid = ceil_div(n, 256);
    cudaStream_t s = as_stream(stream);
    switch (degrees_to_use) {
        case 0: sh_forward_kernel<0><<<grid, 256, 0, s>>>(n, K, viewdirs, coeffs, colors); break;
        case 1: sh_forward_kernel<1><<<grid, 256, 0, s>>>(n, K, viewdirs, coeffs, colors); break;
        case 2: sh_forward_kernel<2><<<grid, 256, 0, s>>>(n, K, viewdirs, coeffs, colors); break;
        case 3: sh_forward_kernel<3><<<grid, 256, 0, s>>>(n, K, viewdirs, coeffs, colors); break;
        default: sh_forward_kernel<4><<<grid, 256, 0, s>>>(n, K, viewdirs, coeffs, colors); break;
    }
    GSTEX_LAUNCH_OK("sh_forward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_sh_backward(int n, int degree, int degrees_to_use, const float *viewdirs, const float *v_colors,
                                 float *v_coeffs, int accumulate, gstex_stream_t stream) {
    int rc = check_sh("sh_backward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    const int K = sh_num_bases(degree), grid = ceil_div(n, 256);
    cudaStream_t s = as_stream(stream);
    switch (degrees_to_use) {
        case 0: sh_backward_kernel<0><<<grid, 256, 0, s>>>(n, K, viewdirs, v_colors, v_coeffs, accumulate); break;
        case 1: sh_backward_kernel<1><<<grid, 256, 0, s>>>(n, K, viewdirs, v_colors, v_coeffs, accumulate); break;
        case 2: sh_backward_kernel<2><<<grid, 256, 0, s>>>(n, K, viewdirs, v_colors, v_coeffs, accumulate); break;
        case 3: sh_backward_kernel<3><<<grid, 256, 0, s>>>(n, K, viewdirs, v_colors, v_coeffs, accumulate); break;
        default: sh_backward_kernel<4><<<grid, 256, 0, s>>>(n, K, viewdirs, v_colors, v_coeffs, accumulate); break;
    }
    GSTEX_LAUNCH_OK("sh_backward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_sh_colors_forward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                                       const float *coeffs, float *colors, uint8_t *mask, gstex_stream_t stream) {
    int rc = check_sh("sh_colors_forward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    const int K = sh_num_bases(degree), grid = ceil_div(n, 256);
    cudaStream_t s = as_stream(stream);
    switch (degrees_to_use) {
        case 0: sh_colors_forward_kernel<0><<<grid, 256, 0, s>>>(n, K, means, c2w, coeffs, colors, mask); break;
        case 1: sh_colors_forward_kernel<1><<<grid, 256, 0, s>>>(n, K, means, c2w, coeffs, colors, mask); break;
        case 2: sh_colors_forward_kernel<2><<<grid, 256, 0, s>>>(n, K, means, c2w, coeffs, colors, mask); break;
        case 3: sh_colors_forward_kernel<3><<<grid, 256, 0, s>>>(n, K, means, c2w, coeffs, colors, mask); break;
        default: sh_colors_forward_kernel<4><<<grid, 256, 0, s>>>(n, K, means, c2w, coeffs, colors, mask); break;
    }
    GSTEX_LAUNCH_OK("sh_colors_forward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_sh_colors_backward(int n, int degree, int degrees_to_use, const float *means, const float *c2w,
                                        const float *v_colors, const uint8_t *mask, float *v_coeffs, int accumulate,
                                        gstex_stream_t stream) {
    int rc = check_sh("sh_colors_backward", n, degree, degrees_to_use);
    if (rc != GSTEX_OK) return rc;
    if (n == 0) return GSTEX_OK;
    const int K = sh_num_bases(degree), grid = ceil_div(n, 256);
    cudaStream_t s = as_stream(stream);
    switch (degrees_to_use) {
        case 0: sh_colors_backward_kernel<0><<<grid, 256, 0, s>>>(n, K, means, c2w, v_colors, mask, v_coeffs, accumulate); break;
        case 1: sh_colors_backward_kernel<1><<<grid, 256, 0, s>>>(n, K, means, c2w, v_colors, mask, v_coeffs, accumulate); break;
        case 2: sh_colors_backward_kernel<2><<<grid, 256, 0, s>>>(n, K, means, c2w, v_colors, mask, v_coeffs, accumulate); break;
        case 3: sh_colors_backward_kernel<3><<<grid, 256, 0, s>>>(n, K, means, c2w, v_colors, mask, v_coeffs, accumulate); break;
        default: sh_colors_backward_kernel<4><<<grid, 256, 0, s>>>(n, K, means, c2w, v_colors, mask, v_coeffs, accumulate); break;
    }
    GSTEX_LAUNCH_OK("sh_colors_backward_kernel");
    return GSTEX_OK;
}
