// train_ops.cu -- the per-Gaussian glue around the rasteriser in a GStex training step:
//
//   preprocess_forward / _backward : raw parameters -> activated parameters and the VJP back
//        (reference example.py:126-143, :162-163, :171: exp on the in-plane scales with the 1e-5 * mean
//        thickness, quaternion normalisation, the uv maps umap/vmap = e^m2 (+-a1 cos/sin m3 + a2 sin/cos m3) built
//        from the surfel axes, sigmoid on colours / opacities).  ~25 torch kernels each way upstream, one launch here.
//   sigmoid_pad_texture / unpad_texture_grad_sigmoid : torch.sigmoid(texture) (example.py:171) and its VJP fused
//        into the float4 padding / un-padding passes the rasteriser needs anyway (no extra pass over the texels).
//   adam_step : torch.optim.Adam's update (example.py:223-225, :278; defaults: no weight decay, no amsgrad) on a
//        contiguous fp32 arena -- one launch for all parameters, natural epilogue of the gradient all-reduce.
//
// All kernels are one pass over their arrays: HBM bound.
#include "raster.cuh"

namespace gstex {

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

struct PreIn {
    const float *raw_scales, *raw_quats, *mapping, *raw_rgbs, *raw_opac;
};

__global__ void __launch_bounds__(256) preprocess_forward_kernel(int n, const PreIn in, float *__restrict__ scales,
                                                                 float4 *__restrict__ quats, float2 *__restrict__ uv0,
                                                                 float *__restrict__ umap, float *__restrict__ vmap,
                                                                 float *__restrict__ colors,
                                                                 float *__restrict__ opacities) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    // example.py:126-128
    const float s1 = expf(in.raw_scales[3 * g]), s2 = expf(in.raw_scales[3 * g + 1]);
    scales[3 * g] = s1;
    scales[3 * g + 1] = s2;
    scales[3 * g + 2] = 1e-5f * (0.5f * (s1 + s2));
    // example.py:129
    float4 q = reinterpret_cast<const float4 *>(in.raw_quats)[g];
    const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
    quats[g] = q;
    // example.py:130-137
    Vec3 a1, a2, a3;
    surfel_axes(q, a1, a2, a3);
    const float4 m = reinterpret_cast<const float4 *>(in.mapping)[g];
    const float us = expf(m.z), c = cosf(m.w), s = sinf(m.w);
    uv0[g] = make_float2(m.x, m.y);
    umap[3 * g] = us * (a1.x * c + a2.x * s);
    umap[3 * g + 1] = us * (a1.y * c + a2.y * s);
    umap[3 * g + 2] = us * (a1.z * c + a2.z * s);
    vmap[3 * g] = us * (-a1.x * s + a2.x * c);
    vmap[3 * g + 1] = us * (-a1.y * s + a2.y * c);
    vmap[3 * g + 2] = us * (-a1.z * s + a2.z * c);
    // example.py:162-163
    if (in.raw_rgbs) {
        colors[3 * g] = sigmoidf(in.raw_rgbs[3 * g]);
        colors[3 * g + 1] = sigmoidf(in.raw_rgbs[3 * g + 1]);
        colors[3 * g + 2] = sigmoidf(in.raw_rgbs[3 * g + 2]);
    }
    opacities[g] = sigmoidf(in.raw_opac[g]);
}

struct PreGrad {
    const float *v_scales, *v_quats, *v_uv0, *v_umap, *v_vmap, *v_colors, *v_opacity;
};

__global__ void __launch_bounds__(256) preprocess_backward_kernel(int n, const PreIn in, const PreGrad gr,
                                                                  float *__restrict__ v_raw_scales,
                                                                  float4 *__restrict__ v_raw_quats,
                                                                  float4 *__restrict__ v_mapping,
                                                                  float *__restrict__ v_raw_rgbs,
                                                                  float *__restrict__ v_raw_opac) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    // scales: d exp; the thickness is detached upstream (example.py:128) and the rasteriser never reads it
    v_raw_scales[3 * g] = gr.v_scales[3 * g] * expf(in.raw_scales[3 * g]);
    v_raw_scales[3 * g + 1] = gr.v_scales[3 * g + 1] * expf(in.raw_scales[3 * g + 1]);
    v_raw_scales[3 * g + 2] = 0.f;
    // uv maps
    const float4 qr = reinterpret_cast<const float4 *>(in.raw_quats)[g];
    const float nrm = sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
    const float4 q = make_float4(qr.x / nrm, qr.y / nrm, qr.z / nrm, qr.w / nrm);
    Vec3 a1, a2, a3;
    surfel_axes(q, a1, a2, a3);
    const float4 m = reinterpret_cast<const float4 *>(in.mapping)[g];
    const float us = expf(m.z), c = cosf(m.w), s = sinf(m.w);
    const Vec3 um = mk3(us * (a1.x * c + a2.x * s), us * (a1.y * c + a2.y * s), us * (a1.z * c + a2.z * s));
    const Vec3 vm = mk3(us * (-a1.x * s + a2.x * c), us * (-a1.y * s + a2.y * c), us * (-a1.z * s + a2.z * c));
    const Vec3 gu = ld3(gr.v_umap + 3 * g), gv = ld3(gr.v_vmap + 3 * g);
    const float2 g0 = reinterpret_cast<const float2 *>(gr.v_uv0)[g];
    v_mapping[g] = make_float4(g0.x, g0.y, dot3(um, gu) + dot3(vm, gv), dot3(vm, gu) - dot3(um, gv));
    const Vec3 v_a1 = mk3(us * (c * gu.x - s * gv.x), us * (c * gu.y - s * gv.y), us * (c * gu.z - s * gv.z));
    const Vec3 v_a2 = mk3(us * (s * gu.x + c * gv.x), us * (s * gu.y + c * gv.y), us * (s * gu.z + c * gv.z));
    // quaternion: rasteriser gradient + the axes' share, then through q / |q|
    const float4 va = surfel_axes_vjp(q, v_a1, v_a2, mk3(0.f, 0.f, 0.f));
    const float4 vq0 = reinterpret_cast<const float4 *>(gr.v_quats)[g];
    const float4 vq = make_float4(vq0.x + va.x, vq0.y + va.y, vq0.z + va.z, vq0.w + va.w);
    const float d = q.x * vq.x + q.y * vq.y + q.z * vq.z + q.w * vq.w;
    v_raw_quats[g] = make_float4((vq.x - q.x * d) / nrm, (vq.y - q.y * d) / nrm, (vq.z - q.z * d) / nrm,
                                 (vq.w - q.w * d) / nrm);
    if (in.raw_rgbs) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float cval = sigmoidf(in.raw_rgbs[3 * g + k]);
            v_raw_rgbs[3 * g + k] = gr.v_colors[3 * g + k] * cval * (1.f - cval);
        }
    }
    const float o = sigmoidf(in.raw_opac[g]);
    v_raw_opac[g] = gr.v_opacity[g] * o * (1.f - o);
}

// tex4[i] = (sigmoid(raw[i, 0..2]), 0)
__global__ void __launch_bounds__(256) sigmoid_pad_texture_kernel(int64_t num_texels, const float *__restrict__ raw,
                                                                  float4 *__restrict__ tex4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_texels) return;
    tex4[i] = make_float4(sigmoidf(raw[3 * i]), sigmoidf(raw[3 * i + 1]), sigmoidf(raw[3 * i + 2]), 0.f);
}

// v_raw[i, c] (+)= g4[i].c * t (1 - t) with t = tex4[i].c, the activated texel the forward pass used
__global__ void __launch_bounds__(256) unpad_texture_grad_sigmoid_kernel(int64_t num_texels,
                                                                         const float4 *__restrict__ g4,
                                                                         const float4 *__restrict__ tex4,
                                                                         float *__restrict__ v_raw, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_texels) return;
    const float4 g = g4[i], t = tex4[i];
    const float r0 = g.x * t.x * (1.f - t.x), r1 = g.y * t.y * (1.f - t.y), r2 = g.z * t.z * (1.f - t.z);
    v_raw[3 * i] = accumulate ? v_raw[3 * i] + r0 : r0;
    v_raw[3 * i + 1] = accumulate ? v_raw[3 * i + 1] + r1 : r1;
    v_raw[3 * i + 2] = accumulate ? v_raw[3 * i + 2] + r2 : r2;
}

// torch.optim.Adam (torch/optim/adam.py _single_tensor_adam, defaults): exp_avg.lerp_(g, 1-b1);
// exp_avg_sq = b2 * exp_avg_sq + (1-b2) g^2; p -= (lr / bc1) * exp_avg / (sqrt(exp_avg_sq) / sqrt(bc2) + eps)
struct AdamArgs {
    float lr_over_bc1, sqrt_bc2, beta1, beta2, omb1, omb2, eps, grad_scale;  // omb = 1 - beta, formed in double
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamArgs &a) {
    g *= a.grad_scale;
    m = m + (g - m) * a.omb1;
    v = a.beta2 * v + a.omb2 * g * g;
    const float denom = sqrtf(v) / a.sqrt_bc2 + a.eps;
    p = p - a.lr_over_bc1 * (m / denom);
}

// Device-resident optimiser state for CUDA-graph replay: the step counter and the scalars derived from it live in
// memory, so that the same captured launch computes a different bias correction every replay.
struct AdamDeviceState {
    int32_t step;
    int32_t pad[3];
    AdamArgs args;
};

// one thread: ++step, then the scalars exactly as the host path forms them (double arithmetic, rounded to fp32 once)
__global__ void adam_prepare_kernel(AdamDeviceState *st, double lr, double beta1, double beta2, double eps,
                                    float grad_scale) {
    const int step = ++st->step;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamArgs a;
    a.lr_over_bc1 = (float)(lr / bc1);
    a.sqrt_bc2 = (float)sqrt(bc2);
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.grad_scale = grad_scale;
    a.omb1 = (float)(1.0 - beta1);
    a.omb2 = (float)(1.0 - beta2);
    st->args = a;
}

__global__ void __launch_bounds__(256) adam_step_kernel(int64_t count, float *__restrict__ p,
                                                        const float *__restrict__ g, float *__restrict__ m,
                                                        float *__restrict__ v, const AdamArgs a_host,
                                                        const AdamDeviceState *__restrict__ st, int vec4) {
    const AdamArgs a = st ? st->args : a_host;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec4) {
        const int64_t quads = count >> 2;
        float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        for (int64_t q = i; q < quads; q += stride) {
            float4 pp = p4[q], mm = m4[q], vv = v4[q];
            const float4 gg = g4[q];
            adam_one(pp.x, gg.x, mm.x, vv.x, a);
            adam_one(pp.y, gg.y, mm.y, vv.y, a);
            adam_one(pp.z, gg.z, mm.z, vv.z, a);
            adam_one(pp.w, gg.w, mm.w, vv.w, a);
            p4[q] = pp; m4[q] = mm; v4[q] = vv;
        }
        for (int64_t e = (quads << 2) + i; e < count; e += stride) adam_one(p[e], g[e], m[e], v[e], a);
    } else {
        for (; i < count; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
    }
}

}  // namespace gstex

using namespace gstex;

extern "C" int gstex_preprocess_forward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                                        const float *raw_rgbs, const float *raw_opacities, float *scales,
                                        float *quats, float *uv0, float *umap, float *vmap, float *colors,
                                        float *opacities, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "preprocess_forward: n = %d", n);
    GSTEX_REQUIRE((raw_rgbs == nullptr) == (colors == nullptr), GSTEX_E_INVALID,
                  "preprocess_forward: raw_rgbs and colors must both be given or both be NULL");
    if (n == 0) return GSTEX_OK;
    const PreIn in{raw_scales, raw_quats, mapping, raw_rgbs, raw_opacities};
    preprocess_forward_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        n, in, scales, (float4 *)quats, (float2 *)uv0, umap, vmap, colors, opacities);
    GSTEX_LAUNCH_OK("preprocess_forward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_preprocess_backward(int n, const float *raw_scales, const float *raw_quats, const float *mapping,
                                         const float *raw_rgbs, const float *raw_opacities, const float *v_scales,
                                         const float *v_quats, const float *v_uv0, const float *v_umap,
                                         const float *v_vmap, const float *v_colors, const float *v_opacity,
                                         float *v_raw_scales, float *v_raw_quats, float *v_mapping, float *v_raw_rgbs,
                                         float *v_raw_opacities, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "preprocess_backward: n = %d", n);
    GSTEX_REQUIRE((raw_rgbs == nullptr) == (v_raw_rgbs == nullptr) && (raw_rgbs == nullptr) == (v_colors == nullptr),
                  GSTEX_E_INVALID, "preprocess_backward: raw_rgbs, v_colors and v_raw_rgbs go together");
    if (n == 0) return GSTEX_OK;
    const PreIn in{raw_scales, raw_quats, mapping, raw_rgbs, raw_opacities};
    const PreGrad gr{v_scales, v_quats, v_uv0, v_umap, v_vmap, v_colors, v_opacity};
    preprocess_backward_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        n, in, gr, v_raw_scales, (float4 *)v_raw_quats, (float4 *)v_mapping, v_raw_rgbs, v_raw_opacities);
    GSTEX_LAUNCH_OK("preprocess_backward_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_sigmoid_pad_texture(int64_t num_texels, const float *raw_texture, float *tex4,
                                         gstex_stream_t stream) {
    GSTEX_REQUIRE(num_texels >= 0, GSTEX_E_INVALID, "sigmoid_pad_texture: texels = %lld", (long long)num_texels);
    if (num_texels == 0) return GSTEX_OK;
    sigmoid_pad_texture_kernel<<<(unsigned)ceil_div64(num_texels, 256), 256, 0, as_stream(stream)>>>(
        num_texels, raw_texture, (float4 *)tex4);
    GSTEX_LAUNCH_OK("sigmoid_pad_texture_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_unpad_texture_grad_sigmoid(int64_t num_texels, const float *g4, const float *tex4,
                                                float *v_raw_texture, int accumulate, gstex_stream_t stream) {
    GSTEX_REQUIRE(num_texels >= 0, GSTEX_E_INVALID, "unpad_texture_grad_sigmoid: texels = %lld", (long long)num_texels);
    if (num_texels == 0) return GSTEX_OK;
    unpad_texture_grad_sigmoid_kernel<<<(unsigned)ceil_div64(num_texels, 256), 256, 0, as_stream(stream)>>>(
        num_texels, (const float4 *)g4, (const float4 *)tex4, v_raw_texture, accumulate);
    GSTEX_LAUNCH_OK("unpad_texture_grad_sigmoid_kernel");
    return GSTEX_OK;
}

// Visible-only ("selective") Adam over one field of the arena: element e belongs to unit e / unit_width, the unit to row
// owner[unit] (or to row `unit` itself), and only rows with visible[row] > 0 are touched - parameters AND both moments of
// an unseen Gaussian stay as they are, so a Gaussian outside every view of the step neither drifts on stale momentum nor
// has its second moment decayed.  The bias corrections use the global step counter, as torch.optim.SparseAdam's dense
// twin and gsplat's SelectiveAdam do.
__global__ void __launch_bounds__(256) adam_rows_kernel(int64_t count, float *__restrict__ p, const float *__restrict__ g,
                                                        float *__restrict__ m, float *__restrict__ v,
                                                        const AdamDeviceState *__restrict__ st,
                                                        const float *__restrict__ visible, int unit_width,
                                                        const int32_t *__restrict__ owner) {
    const AdamArgs a = st->args;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) {
        const int64_t unit = e / unit_width;
        const int64_t row = owner ? (int64_t)owner[unit] : unit;
        if (visible[row] > 0.f) adam_one(p[e], g[e], m[e], v[e], a);
    }
}

extern "C" int gstex_adam_prepare_device(void *state, double lr, double beta1, double beta2, double eps, float grad_scale,
                                         gstex_stream_t stream) {
    GSTEX_REQUIRE(state != nullptr, GSTEX_E_INVALID, "adam_prepare_device: state is NULL");
    adam_prepare_kernel<<<1, 1, 0, as_stream(stream)>>>((AdamDeviceState *)state, lr, beta1, beta2, eps, grad_scale);
    GSTEX_LAUNCH_OK("adam_prepare_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_adam_apply_rows_device(int64_t count, float *params, const float *grads, float *exp_avg,
                                            float *exp_avg_sq, const void *state, const float *visible, int unit_width,
                                            const int32_t *owner, gstex_stream_t stream) {
    GSTEX_REQUIRE(count >= 0 && state != nullptr && visible != nullptr && unit_width >= 1, GSTEX_E_INVALID,
                  "adam_apply_rows_device: count = %lld, unit_width = %d", (long long)count, unit_width);
    if (count == 0) return GSTEX_OK;
    const int blocks = (int)min((int64_t)148 * 8, ceil_div64(count, 256));
    adam_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(count, params, grads, exp_avg, exp_avg_sq,
                                                            (const AdamDeviceState *)state, visible, unit_width, owner);
    GSTEX_LAUNCH_OK("adam_rows_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_adam_step(int64_t count, float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                               double lr, double beta1, double beta2, double eps, int step, float grad_scale,
                               gstex_stream_t stream) {
    GSTEX_REQUIRE(count >= 0 && step >= 1, GSTEX_E_INVALID, "adam_step: count = %lld, step = %d (1-based)",
                  (long long)count, step);
    if (count == 0) return GSTEX_OK;
    // scalars are formed in double on the host, as torch does with its Python floats, then rounded to fp32
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamArgs a;
    a.lr_over_bc1 = (float)(lr / bc1);
    a.sqrt_bc2 = (float)sqrt(bc2);
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.grad_scale = grad_scale;
    a.omb1 = (float)(1.0 - beta1);
    a.omb2 = (float)(1.0 - beta2);
    const int vec4 = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0;
    const int blocks = (int)min((int64_t)148 * 8, ceil_div64(vec4 ? ceil_div64(count, 4) : count, 256));
    adam_step_kernel<<<blocks, 256, 0, as_stream(stream)>>>(count, params, grads, exp_avg, exp_avg_sq, a, nullptr, vec4);
    GSTEX_LAUNCH_OK("adam_step_kernel");
    return GSTEX_OK;
}

extern "C" size_t gstex_adam_state_bytes(void) { return sizeof(AdamDeviceState); }

extern "C" int gstex_adam_step_device(int64_t count, float *params, const float *grads, float *exp_avg,
                                      float *exp_avg_sq, double lr, double beta1, double beta2, double eps,
                                      void *state, float grad_scale, gstex_stream_t stream) {
    GSTEX_REQUIRE(count >= 0 && state != nullptr, GSTEX_E_INVALID, "adam_step_device: count = %lld, state = %p",
                  (long long)count, state);
    AdamDeviceState *st = (AdamDeviceState *)state;
    adam_prepare_kernel<<<1, 1, 0, as_stream(stream)>>>(st, lr, beta1, beta2, eps, grad_scale);
    GSTEX_LAUNCH_OK("adam_prepare_kernel");
    if (count == 0) return GSTEX_OK;
    const int vec4 = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0;
    const int blocks = (int)min((int64_t)148 * 8, ceil_div64(vec4 ? ceil_div64(count, 4) : count, 256));
    adam_step_kernel<<<blocks, 256, 0, as_stream(stream)>>>(count, params, grads, exp_avg, exp_avg_sq, AdamArgs{}, st, vec4);
    GSTEX_LAUNCH_OK("adam_step_kernel");
    return GSTEX_OK;
}
