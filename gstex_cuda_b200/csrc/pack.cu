// pack.cu -- per-view Gaussian record packing, its VJP (the backward "epilogue"), and the
// float4-padded texture / texel-gradient staging buffers.
//
// pack_kernel    : (mean, scale, quat, opacity, colour, uv map, texture dims; camera) -> 128-byte record
// epilogue_kernel: per-Gaussian gradient moments (written by the backward rasteriser) -> gradients of
//                  means / scales / quats / opacity / colour / uv0 / umap / vmap.  No atomics, no zero-fill.
// Both mirror tests/formulation.py::pack_record / ::epilogue line by line (same names).
#include "raster.cuh"

namespace gstex {

struct PackCamera {
    Vec3 o;          // camera origin (c2w[:3,3])
    float Rc[12];    // c2w rows 0..2 (rotation in [0,1,2],[4,5,6],[8,9,10])
    float vm[12];    // viewmat rows 0..2
};

__device__ __forceinline__ void load_camera(const float *__restrict__ c2w, const float *__restrict__ viewmat,
                                            PackCamera &cam) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        cam.Rc[i] = c2w[i];
        cam.vm[i] = viewmat[i];
    }
    cam.o = mk3(c2w[3], c2w[7], c2w[11]);
}

struct SurfelFrame {
    Vec3 a1, a2, a3, d, um, vm_;
    float c0, b1, b2, bu, bv, k1, k2;
    Vec3 rc;       // expansion direction (uc, vc, 1)
    bool exact;    // expansion centre is the exact projection of the mean (constants c1,c2,cu,cv vanish)
};

__device__ __forceinline__ void make_frame(Vec3 mean, float s1, float s2, float4 quat, Vec3 umap, Vec3 vmap,
                                           float glob_scale, const PackCamera &cam, SurfelFrame &f) {
    surfel_axes(quat, f.a1, f.a2, f.a3);
    f.d = sub3(mean, cam.o);
    f.um = umap;
    f.vm_ = vmap;
    f.c0 = dot3(f.a3, f.d);
    f.b1 = dot3(f.a1, f.d);
    f.b2 = dot3(f.a2, f.d);
    f.bu = dot3(umap, f.d);
    f.bv = dot3(vmap, f.d);
    f.k1 = K_SIGMA / (s1 * glob_scale);
    f.k2 = K_SIGMA / (s2 * glob_scale);
    const Vec3 mc = rot_apply_t(cam.Rc, f.d);
    f.exact = false;
    f.rc = mk3(0.f, 0.f, 1.f);
    if (mc.z > 1e-4f) {
        const float uc = mc.x / mc.z, vc = mc.y / mc.z;
        if (fabsf(uc) < 1e3f && fabsf(vc) < 1e3f) {
            f.exact = true;
            f.rc = mk3(uc, vc, 1.f);
        }
    }
}

// w = c0 * a - (a.d) * a3 ; h = Rc^T w   (numerator of a.(delta) as a linear form in the ray)
__device__ __forceinline__ Vec3 form_vector(const SurfelFrame &f, Vec3 a, float b, const PackCamera &cam) {
    const Vec3 w = mk3(fmaf(f.c0, a.x, -b * f.a3.x), fmaf(f.c0, a.y, -b * f.a3.y), fmaf(f.c0, a.z, -b * f.a3.z));
    return rot_apply_t(cam.Rc, w);
}

#ifndef GSTEX_PACK_MINB
#define GSTEX_PACK_MINB 1
#endif
// |c3| below which a surfel counts as grazing (its normal within ~3 degrees of perpendicular to the centre ray)
constexpr float GRAZING_C3 = 0.05f;
__global__ void __launch_bounds__(256, GSTEX_PACK_MINB) pack_kernel(int n, const float *__restrict__ means,
                                                   const float *__restrict__ scales, float glob_scale,
                                                   const float4 *__restrict__ quats,
                                                   const float *__restrict__ opacities,
                                                   const float *__restrict__ colors, const float2 *__restrict__ uv0,
                                                   const float *__restrict__ umap, const float *__restrict__ vmap,
                                                   const int32_t *__restrict__ texture_dims,
                                                   const float *__restrict__ viewmat, const float *__restrict__ c2w,
                                                   float fx, float fy, float cx, float cy,
                                                   float4 *__restrict__ recs, float2 *__restrict__ mean2d,
                                                   float4 *__restrict__ acc_to_zero) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    if (acc_to_zero) {  // the view's moment line starts from zero: cleared here instead of by a separate fill pass
        float4 *a = acc_to_zero + (size_t)g * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    PackCamera cam;
    load_camera(c2w, viewmat, cam);
    const Vec3 mean = ld3(means + 3 * g);
    SurfelFrame f;
    make_frame(mean, scales[3 * g], scales[3 * g + 1], quats[g], ld3(umap + 3 * g), ld3(vmap + 3 * g), glob_scale,
               cam, f);
    const Vec3 h1 = form_vector(f, f.a1, f.b1, cam), h2 = form_vector(f, f.a2, f.b2, cam);
    const Vec3 hu = form_vector(f, f.um, f.bu, cam), hv = form_vector(f, f.vm_, f.bv, cam);
    const Vec3 h3 = rot_apply_t(cam.Rc, f.a3);
    const float ifx = 1.f / fx, ify = 1.f / fy;
    const float c1 = f.exact ? 0.f : f.k1 * dot3(h1, f.rc), c2 = f.exact ? 0.f : f.k2 * dot3(h2, f.rc);
    const float cu = f.exact ? 0.f : dot3(hu, f.rc), cv = f.exact ? 0.f : dot3(hv, f.rc);
    const float c3 = dot3(h3, f.rc);
    const float2 t0 = uv0[g];
    float4 *r = recs + (size_t)g * 8;
    const float xc = fmaf(fx, f.rc.x, cx), yc = fmaf(fy, f.rc.y, cy);
    float4 r0 = make_float4(xc, yc, f.c0, opacities[g]);
    float4 r1 = make_float4(f.k1 * h1.x * ifx, f.k1 * h1.y * ify, c1, c3);
    float4 r2 = make_float4(f.k2 * h2.x * ifx, f.k2 * h2.y * ify, c2, __int_as_float(g));
    float4 r3 = make_float4(h3.x * ifx, h3.y * ify, __int_as_float(texture_dims[3 * g]),
                            __int_as_float(texture_dims[3 * g + 1]));
    float4 r4 = make_float4(hu.x * ifx, hu.y * ify, cu, t0.x);
    float4 r5 = make_float4(hv.x * ifx, hv.y * ify, cv, t0.y);
    // Grazing surfels.  c3 = a3 . R(centre ray) is the cosine-like plane denominator at the Gaussian's centre: a sum of
    // O(1) products, so its fp32 value carries ~1e-7 ABSOLUTE error - a relative error of 1e-7 / |c3| that every 1/D and
    // 1/D^2 factor of the pair evaluation and its gradients inherits (the reference's dot(ray, ax3), texture_helpers.cuh:
    // 302-313, has the same conditioning; at |c3| ~ 1e-3 either evaluation is only good to 1e-4).  For the few surfels seen
    // nearly edge-on the form coefficients are therefore recomputed in double, with the constants evaluated at the
    // ROUNDED expansion centre (xc, yc) the kernels subtract from the pixel - so that N(p) = c + P.(p - centre) holds to
    // fp32 rounding of the result instead of to fp32 rounding of its O(1) terms.
    if (fabsf(c3) < GRAZING_C3) {
        const double qw = quats[g].x, qx = quats[g].y, qy = quats[g].z, qz = quats[g].w;
        const double a1[3] = {1.0 - 2.0 * (qy * qy + qz * qz), 2.0 * (qx * qy + qw * qz), 2.0 * (qx * qz - qw * qy)};
        const double a2[3] = {2.0 * (qx * qy - qw * qz), 1.0 - 2.0 * (qx * qx + qz * qz), 2.0 * (qy * qz + qw * qx)};
        const double a3[3] = {2.0 * (qx * qz + qw * qy), 2.0 * (qy * qz - qw * qx), 1.0 - 2.0 * (qx * qx + qy * qy)};
        const double um[3] = {(double)umap[3 * g], (double)umap[3 * g + 1], (double)umap[3 * g + 2]};
        const double vm[3] = {(double)vmap[3 * g], (double)vmap[3 * g + 1], (double)vmap[3 * g + 2]};
        const double d[3] = {(double)mean.x - (double)cam.o.x, (double)mean.y - (double)cam.o.y, (double)mean.z - (double)cam.o.z};
        auto dotd = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
        const double c0d = dotd(a3, d);
        // centre ray through the rounded expansion centre, camera frame
        const double rr[3] = {((double)xc - (double)cx) / (double)fx, ((double)yc - (double)cy) / (double)fy, 1.0};
        // h = Rc^T (c0 a - (a.d) a3): coefficients of the linear form in the camera-frame ray; h3 = Rc^T a3
        auto form = [&](const double *a, double kappa, bool plane, float &px, float &py, float &c) {
            double w[3];
            const double b = plane ? 0.0 : dotd(a, d);
#pragma unroll
            for (int k = 0; k < 3; ++k) w[k] = plane ? a[k] : c0d * a[k] - b * a3[k];
            double h[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                h[k] = (double)cam.Rc[k] * w[0] + (double)cam.Rc[4 + k] * w[1] + (double)cam.Rc[8 + k] * w[2];
            px = (float)(kappa * h[0] / (double)fx);
            py = (float)(kappa * h[1] / (double)fy);
            c = (float)(kappa * dotd(h, rr));
        };
        const double k1d = (double)K_SIGMA / ((double)scales[3 * g] * (double)glob_scale);
        const double k2d = (double)K_SIGMA / ((double)scales[3 * g + 1] * (double)glob_scale);
        r0.z = (float)c0d;
        form(a1, k1d, false, r1.x, r1.y, r1.z);
        form(a2, k2d, false, r2.x, r2.y, r2.z);
        form(a3, 1.0, true, r3.x, r3.y, r1.w);
        form(um, 1.0, false, r4.x, r4.y, r4.z);
        form(vm, 1.0, false, r5.x, r5.y, r5.z);
    }
    r[0] = r0;
    r[1] = r1;
    r[2] = r2;
    r[3] = r3;
    r[4] = r4;
    r[5] = r5;
    // colours are optional: texture_edit walks the same records without them
    const Vec3 col = colors ? ld3(colors + 3 * g) : mk3(0.f, 0.f, 0.f);
    r[6] = make_float4(col.x, col.y, col.z, __int_as_float(texture_dims[3 * g + 2]));
    r[7] = make_float4(f.a3.x, f.a3.y, f.a3.z, 0.f);
    mean2d[g] = pinhole(fx, fy, cx, cy, xform_point(cam.vm, mean));
}

// *p = v, or *p += v when accumulating: the old value is only read in the accumulating case
__device__ __forceinline__ void put(float *__restrict__ p, float v, int accumulate) { *p = accumulate ? *p + v : v; }

#ifndef GSTEX_EPI_MINB
#define GSTEX_EPI_MINB 1
#endif
__global__ void __launch_bounds__(256, GSTEX_EPI_MINB) epilogue_kernel(
    int n, const float *__restrict__ means, const float *__restrict__ scales, float glob_scale,
    const float4 *__restrict__ quats, const float *__restrict__ umap, const float *__restrict__ vmap,
    const float *__restrict__ viewmat, const float *__restrict__ c2w, float fx, float fy, float cx, float cy,
    const float4 *__restrict__ acc, const float4 *__restrict__ recs, float *__restrict__ v_colors,
    float *__restrict__ v_opacity,
    float *__restrict__ v_means, float *__restrict__ v_scales, float4 *__restrict__ v_quats,
    float2 *__restrict__ v_uv0, float *__restrict__ v_umap, float *__restrict__ v_vmap, int accumulate_flags) {
    const int accumulate = accumulate_flags & 1, accumulate_colors = (accumulate_flags >> 1) & 1;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    PackCamera cam;
    load_camera(c2w, viewmat, cam);
    const Vec3 mean = ld3(means + 3 * g);
    const float s1 = scales[3 * g], s2 = scales[3 * g + 1];
    const float4 quat = quats[g];
    SurfelFrame f;
    make_frame(mean, s1, s2, quat, ld3(umap + 3 * g), ld3(vmap + 3 * g), glob_scale, cam, f);
    const float4 *A = acc + (size_t)g * 8;
    const float4 q0 = A[0], q1 = A[1], q2 = A[2], q3 = A[3], q4 = A[4], q5 = A[5], q6 = A[6], q7 = A[7];
    const float ifx = 1.f / fx, ify = 1.f / fy;
    // dL/dh from the moments of dL/dN:  kappa * (Gx/fx + Gc*uc, Gy/fy + Gc*vc, Gc)
    auto v_h = [&](float gx, float gy, float gc, float kappa) {
        return mk3(kappa * fmaf(gc, f.rc.x, gx * ifx), kappa * fmaf(gc, f.rc.y, gy * ify), kappa * gc);
    };
    const Vec3 v_w1 = rot_apply(cam.Rc, v_h(q0.x, q0.y, q0.z, f.k1));
    const Vec3 v_w2 = rot_apply(cam.Rc, v_h(q1.x, q1.y, q1.z, f.k2));
    const Vec3 v_wu = rot_apply(cam.Rc, v_h(q3.x, q3.y, q3.z, 1.f));
    const Vec3 v_wv = rot_apply(cam.Rc, v_h(q4.x, q4.y, q4.z, 1.f));
    Vec3 v_a3 = add3(rot_apply(cam.Rc, v_h(q2.x, q2.y, q2.z, 1.f)), mk3(q6.x, q6.y, q6.z));
    const float v_c0 = q0.w + dot3(v_w1, f.a1) + dot3(v_w2, f.a2) + dot3(v_wu, f.um) + dot3(v_wv, f.vm_);
    const float v_b1 = -dot3(v_w1, f.a3), v_b2 = -dot3(v_w2, f.a3);
    const float v_bu = -dot3(v_wu, f.a3), v_bv = -dot3(v_wv, f.a3);
    const Vec3 v_a1 = axpy3(f.c0, v_w1, scale3(v_b1, f.d));
    const Vec3 v_a2 = axpy3(f.c0, v_w2, scale3(v_b2, f.d));
    const Vec3 v_um = axpy3(f.c0, v_wu, scale3(v_bu, f.d));
    const Vec3 v_vm = axpy3(f.c0, v_wv, scale3(v_bv, f.d));
    v_a3 = axpy3(-f.b1, v_w1, v_a3);
    v_a3 = axpy3(-f.b2, v_w2, v_a3);
    v_a3 = axpy3(-f.bu, v_wu, v_a3);
    v_a3 = axpy3(-f.bv, v_wv, v_a3);
    v_a3 = axpy3(v_c0, f.d, v_a3);
    Vec3 v_d = scale3(v_b1, f.a1);
    v_d = axpy3(v_b2, f.a2, v_d);
    v_d = axpy3(v_bu, f.um, v_d);
    v_d = axpy3(v_bv, f.vm_, v_d);
    v_d = axpy3(v_c0, f.a3, v_d);
    // blur branch: gradient of the projected mean (reference helpers.cuh:155-164, texture.cu:685-692)
    {
        const Vec3 pv = xform_point(cam.vm, mean);
        const float rw = 1.f / (pv.z + 1e-6f);
        const float gx = fx * q7.x, gy = fy * q7.y;
        const Vec3 v_pv = mk3(gx * rw, gy * rw, -(gx * pv.x + gy * pv.y) * rw * rw);
        v_d = add3(v_d, rot_apply_t(cam.vm, v_pv));
    }
    // scales: kappa_i = K/(s_i*glob)  =>  dL/ds_i = -(F_i . G_i)/s_i with F_i the record's form coefficients - read back
    // from the record the rasterisers used when the caller has it (for grazing surfels pack_kernel computes them in
    // double; F . G = sum over pairs of g * N is a sum of small N's there, and must use the same coefficients)
    float fg1, fg2;
    if (recs) {
        const float4 F1 = recs[(size_t)g * 8 + 1], F2 = recs[(size_t)g * 8 + 2];
        fg1 = fmaf(F1.x, q0.x, fmaf(F1.y, q0.y, F1.z * q0.z));
        fg2 = fmaf(F2.x, q1.x, fmaf(F2.y, q1.y, F2.z * q1.z));
    } else {
        const Vec3 h1 = form_vector(f, f.a1, f.b1, cam), h2 = form_vector(f, f.a2, f.b2, cam);
        const float c1 = f.exact ? 0.f : f.k1 * dot3(h1, f.rc), c2 = f.exact ? 0.f : f.k2 * dot3(h2, f.rc);
        fg1 = fmaf(f.k1 * h1.x * ifx, q0.x, fmaf(f.k1 * h1.y * ify, q0.y, c1 * q0.z));
        fg2 = fmaf(f.k2 * h2.x * ifx, q1.x, fmaf(f.k2 * h2.y * ify, q1.y, c2 * q1.z));
    }
    const float4 vq = surfel_axes_vjp(quat, v_a1, v_a2, v_a3);

    put(&v_means[3 * g + 0], v_d.x, accumulate);
    put(&v_means[3 * g + 1], v_d.y, accumulate);
    put(&v_means[3 * g + 2], v_d.z, accumulate);
    put(&v_scales[3 * g + 0], -fg1 / s1, accumulate);
    put(&v_scales[3 * g + 1], -fg2 / s2, accumulate);
    put(&v_scales[3 * g + 2], 0.f, accumulate);
    float4 oq = accumulate ? v_quats[g] : make_float4(0.f, 0.f, 0.f, 0.f);
    v_quats[g] = make_float4(oq.x + vq.x, oq.y + vq.y, oq.z + vq.z, oq.w + vq.w);
    float2 ou = accumulate ? v_uv0[g] : make_float2(0.f, 0.f);
    v_uv0[g] = make_float2(ou.x + q3.w, ou.y + q4.w);
    put(&v_umap[3 * g + 0], v_um.x, accumulate);
    put(&v_umap[3 * g + 1], v_um.y, accumulate);
    put(&v_umap[3 * g + 2], v_um.z, accumulate);
    put(&v_vmap[3 * g + 0], v_vm.x, accumulate);
    put(&v_vmap[3 * g + 1], v_vm.y, accumulate);
    put(&v_vmap[3 * g + 2], v_vm.z, accumulate);
    put(&v_colors[3 * g + 0], q5.x, accumulate_colors);
    put(&v_colors[3 * g + 1], q5.y, accumulate_colors);
    put(&v_colors[3 * g + 2], q5.z, accumulate_colors);
    put(&v_opacity[g], q1.w, accumulate);
}

// (X,3) -> (X,float4) so that a texel is one aligned 16-byte load / one vector atomic
__global__ void __launch_bounds__(256) pad_texture_kernel(int64_t num_texels, const float *__restrict__ tex,
                                                          float4 *__restrict__ tex4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_texels) return;
    tex4[i] = make_float4(tex[3 * i], tex[3 * i + 1], tex[3 * i + 2], 0.f);
}

__global__ void __launch_bounds__(256) unpad_texture_grad_kernel(int64_t num_texels, const float4 *__restrict__ g4,
                                                                 float *__restrict__ v_texture, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_texels) return;
    const float4 g = g4[i];
    put(&v_texture[3 * i + 0], g.x, accumulate);
    put(&v_texture[3 * i + 1], g.y, accumulate);
    put(&v_texture[3 * i + 2], g.z, accumulate);
}

// ---- host launchers used by raster_forward.cu / raster_backward.cu ----------------------------------
int launch_pack(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                const float *opacities, const float *colors, const float *uv0, const float *umap, const float *vmap,
                const int32_t *texture_dims, const float *viewmat, const float *c2w, float fx, float fy, float cx,
                float cy, float4 *recs, float2 *mean2d, cudaStream_t s, float4 *acc_to_zero) {
    if (n == 0) return GSTEX_OK;
    pack_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, means, scales, glob_scale, (const float4 *)quats, opacities,
                                                 colors, (const float2 *)uv0, umap, vmap, texture_dims, viewmat, c2w,
                                                 fx, fy, cx, cy, recs, mean2d, acc_to_zero);
    GSTEX_LAUNCH_OK("pack_kernel");
    return GSTEX_OK;
}

int launch_epilogue(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                    const float *umap, const float *vmap, const float *viewmat, const float *c2w, float fx, float fy,
                    float cx, float cy, const float4 *acc, float *v_colors, float *v_opacity, float *v_means,
                    float *v_scales, float *v_quats, float *v_uv0, float *v_umap, float *v_vmap, int accumulate,
                    cudaStream_t s, const float4 *recs) {
    if (n == 0) return GSTEX_OK;
    epilogue_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, means, scales, glob_scale, (const float4 *)quats, umap, vmap,
                                                     viewmat, c2w, fx, fy, cx, cy, acc, recs, v_colors, v_opacity, v_means,
                                                     v_scales, (float4 *)v_quats, (float2 *)v_uv0, v_umap, v_vmap,
                                                     accumulate);
    GSTEX_LAUNCH_OK("epilogue_kernel");
    return GSTEX_OK;
}

int launch_pad_texture(int64_t num_texels, const float *tex, float4 *tex4, cudaStream_t s) {
    if (num_texels == 0) return GSTEX_OK;
    pad_texture_kernel<<<(unsigned)ceil_div64(num_texels, 256), 256, 0, s>>>(num_texels, tex, tex4);
    GSTEX_LAUNCH_OK("pad_texture_kernel");
    return GSTEX_OK;
}

int launch_unpad_texture_grad(int64_t num_texels, const float4 *g4, float *v_texture, int accumulate,
                              cudaStream_t s) {
    if (num_texels == 0) return GSTEX_OK;
    unpad_texture_grad_kernel<<<(unsigned)ceil_div64(num_texels, 256), 256, 0, s>>>(num_texels, g4, v_texture,
                                                                                   accumulate);
    GSTEX_LAUNCH_OK("unpad_texture_grad_kernel");
    return GSTEX_OK;
}

}  // namespace gstex
