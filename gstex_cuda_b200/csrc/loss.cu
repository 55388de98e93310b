// loss.cu -- fused image loss of the reference trainer (example.py:189-209) and its gradient with respect
// to the six rasteriser outputs, one pass over the image (needed by the fused step):
//   loss = mean((out_texture - gt)^2) + mean(out_reg) + mean(nx^2 + ny^2 + (1 - nz)^2)
#include "common.cuh"

namespace gstex {

__global__ void __launch_bounds__(256) image_loss_kernel(int npix, const float *__restrict__ out_texture,
                                                         const float *__restrict__ out_reg,
                                                         const float *__restrict__ out_normal,
                                                         const float *__restrict__ gt, float *__restrict__ loss_accum,
                                                         float *__restrict__ v_img, float *__restrict__ v_depth,
                                                         float *__restrict__ v_reg, float *__restrict__ v_alpha,
                                                         float *__restrict__ v_tex, float *__restrict__ v_normal) {
    __shared__ float warp_part[8];
    const float inv_p = 1.f / (float)npix, inv_3p = 1.f / (3.f * (float)npix);
    float part = 0.f;
    // grid-stride: a bounded number of CTAs, so that the single loss accumulator receives ~1 k atomics, not one per 256 pixels
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        float mse = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = out_texture[3 * i + c] - gt[3 * i + c];
            mse += d * d;
            v_tex[3 * i + c] = 2.f * d * inv_3p;
            if (v_img) v_img[3 * i + c] = 0.f;
        }
        const float nx = out_normal[3 * i], ny = out_normal[3 * i + 1], nz = out_normal[3 * i + 2];
        part += mse * inv_3p + out_reg[i] * inv_p + (nx * nx + ny * ny + (1.f - nz) * (1.f - nz)) * inv_p;
        v_normal[3 * i] = 2.f * nx * inv_p;
        v_normal[3 * i + 1] = 2.f * ny * inv_p;
        v_normal[3 * i + 2] = -2.f * (1.f - nz) * inv_p;
        v_reg[i] = inv_p;
        if (v_depth) v_depth[i] = 0.f;  // the three outputs this loss does not use: NULL = not wanted (the backward
        if (v_alpha) v_alpha[i] = 0.f;  // rasteriser takes NULL for a zero gradient)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += warp_part[w];
        atomicAdd(loss_accum, s);
    }
}

}  // namespace gstex

using namespace gstex;

extern "C" int gstex_image_loss(int img_height, int img_width, const float *out_texture, const float *out_reg,
                                const float *out_normal, const float *gt, float *loss_accum, float *v_out_img,
                                float *v_out_depth, float *v_out_reg, float *v_out_alpha, float *v_out_texture,
                                float *v_out_normal, gstex_stream_t stream) {
    GSTEX_REQUIRE(img_height > 0 && img_width > 0, GSTEX_E_INVALID, "image_loss: image %dx%d", img_height, img_width);
    const int npix = img_height * img_width;
    const int blocks = min(ceil_div(npix, 256), 148 * 8);
    image_loss_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
        npix, out_texture, out_reg, out_normal, gt, loss_accum, v_out_img, v_out_depth, v_out_reg, v_out_alpha,
        v_out_texture, v_out_normal);
    GSTEX_LAUNCH_OK("image_loss_kernel");
    return GSTEX_OK;
}
