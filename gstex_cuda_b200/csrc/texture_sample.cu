// texture_sample.cu -- per-query bilinear fetch from the jagged texture (reference
// texture_sample.cu:11-141) and the scatter that is its transpose.  One thread per (query, 4-channel group).
#include "raster.cuh"

namespace gstex {

__global__ void __launch_bounds__(256) sample_forward_kernel(int nq, int C, const int32_t *__restrict__ dims,
                                                             const float2 *__restrict__ uvs,
                                                             const float *__restrict__ tex, float *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 uv = uvs[q];
    TexFetch f;
    texel_setup(dims[3 * q], dims[3 * q + 1], dims[3 * q + 2], clamp01(uv.x), clamp01(uv.y), true, f);
    for (int c = 0; c < C; ++c)
        out[(size_t)q * C + c] = f.w[0] * tex[(size_t)f.idx[0] * C + c] + f.w[1] * tex[(size_t)f.idx[1] * C + c] +
                                 f.w[2] * tex[(size_t)f.idx[2] * C + c] + f.w[3] * tex[(size_t)f.idx[3] * C + c];
}

__global__ void __launch_bounds__(256) sample_backward_kernel(int nq, int C, const int32_t *__restrict__ dims,
                                                              const float2 *__restrict__ uvs,
                                                              const float *__restrict__ v_out,
                                                              float *__restrict__ v_tex) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 uv = uvs[q];
    TexFetch f;
    texel_setup(dims[3 * q], dims[3 * q + 1], dims[3 * q + 2], clamp01(uv.x), clamp01(uv.y), true, f);
    for (int c = 0; c < C; ++c) {
        const float g = v_out[(size_t)q * C + c];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (f.w[k] != 0.f) atomicAdd(v_tex + (size_t)f.idx[k] * C + c, f.w[k] * g);
    }
}

}  // namespace gstex

using namespace gstex;

extern "C" int gstex_texture_sample_forward(int num_queries, int channels, const int32_t *texture_dims,
                                            const float *uvs, const float *texture, float *output,
                                            gstex_stream_t stream) {
    GSTEX_REQUIRE(num_queries >= 0 && channels >= 1, GSTEX_E_INVALID, "texture_sample_forward: q = %d, c = %d",
                  num_queries, channels);
    if (num_queries == 0) return GSTEX_OK;
    sample_forward_kernel<<<ceil_div(num_queries, 256), 256, 0, as_stream(stream)>>>(
        num_queries, channels, texture_dims, (const float2 *)uvs, texture, output);
    GSTEX_LAUNCH_OK("sample_forward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_texture_sample_backward(int num_queries, int channels, const int32_t *texture_dims,
                                             const float *uvs, const float *v_output, float *v_texture,
                                             gstex_stream_t stream) {
    GSTEX_REQUIRE(num_queries >= 0 && channels >= 1, GSTEX_E_INVALID, "texture_sample_backward: q = %d, c = %d",
                  num_queries, channels);
    if (num_queries == 0) return GSTEX_OK;
    sample_backward_kernel<<<ceil_div(num_queries, 256), 256, 0, as_stream(stream)>>>(
        num_queries, channels, texture_dims, (const float2 *)uvs, v_output, v_texture);
    GSTEX_LAUNCH_OK("sample_backward_kernel");
    return GSTEX_OK;
}
