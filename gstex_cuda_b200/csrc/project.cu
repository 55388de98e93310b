// project.cu -- surfel projection, screen-space AABB and tile counts.
//
// One thread per Gaussian, 256-thread blocks.  The kernels are trivially HBM bound
// (40 B in, <= 24 B out per Gaussian); the fused variant makes one pass instead of the
// reference's ~17 launches (a matmul, the AABB kernel and ~10 torch elementwise ops).
#include "common.cuh"

namespace gstex {

// Bit-exactness against the reference's compiled kernel.  The tile ranges and the sorted ids downstream are integer
// functions of (centre, extent, depth): one ulp at a tile edge changes M and every index after it.  The helpers below
// therefore fix the rounding points to the ones the reference's own build makes (nvcc -O3, default -fmad=true; read
// from the SASS of get_aabb_2d_kernel in oracle/_ref/gstex_ref_C.so) instead of leaving FMA contraction to the compiler:
//   transform_4x3 (helpers.cuh:124-131)   m0*x + m1*y + m2*z + m3  ->  (fma(z, m2, fma(x, m0, rn(y*m1)))) + m3
//   quat_to_rotmat (helpers.cuh:166-185)  a*b + c*d                ->  fma(a, b, rn(c*d)); y*z + w*x -> fma(w, x, rn(y*z))
//   project_pix (helpers.cuh:145-152)     rw = rcp.rn(z + 1e-6);  rn(x*rw) ; fma(., fx, cx)
//   corners (get_aabb_2d.cu:46-65)        fma(ax2, +-rn(ell*s2), fma(ax1, +-rn(ell*s1), p)),  ell = rn(3*glob_scale)
__device__ __forceinline__ Vec3 xform_point_ref(const float *__restrict__ m, Vec3 p) {
    return Vec3{__fadd_rn(fmaf(p.z, m[2], fmaf(p.x, m[0], __fmul_rn(p.y, m[1]))), m[3]),
                __fadd_rn(fmaf(p.z, m[6], fmaf(p.x, m[4], __fmul_rn(p.y, m[5]))), m[7]),
                __fadd_rn(fmaf(p.z, m[10], fmaf(p.x, m[8], __fmul_rn(p.y, m[9]))), m[11])};
}
__device__ __forceinline__ float2 pinhole_ref(float fx, float fy, float cx, float cy, Vec3 pv) {
    const float rw = __frcp_rn(__fadd_rn(pv.z, 1e-6f));
    return float2{fmaf(__fmul_rn(pv.x, rw), fx, cx), fmaf(__fmul_rn(pv.y, rw), fy, cy)};
}
// view-space depth as the reference's Python computes it (get_aabb_2d.py:22-32): `points @ viewmat.T[:3,:3]` is an
// fp32 GEMM that accumulates k = 0, 1, 2 in order from zero, the translation is a separate torch add.  The depth bits
// are the low half of the sort key (forward.cu:63-66), so near-ties in depth order by them.
__device__ __forceinline__ float view_depth_ref(const float *__restrict__ m, Vec3 p) {
    return __fadd_rn(fmaf(p.z, m[10], fmaf(p.y, m[9], __fmul_rn(p.x, m[8]))), m[11]);
}

// Screen AABB of the 3-sigma rectangle of a surfel.  Follows get_aabb_2d_kernel
// (reference get_aabb_2d.cu:11-89): near plane 0.01, each corner's z clamped to the near plane,
// clipped means (z <= 0.01) get extent 0 and the projected mean as centre.
__device__ __forceinline__ void surfel_aabb(Vec3 m, float s1, float s2, float glob_scale, float4 q,
                                            const float *__restrict__ vm, float fx, float fy, float cx,
                                            float cy, float2 &center, float2 &extent, float &depth) {
    const Vec3 pv = xform_point_ref(vm, m);
    depth = view_depth_ref(vm, m);
    const bool clipped = pv.z <= T_NEAR;
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    const float zz = __fmul_rn(z, z), wz = __fmul_rn(w, z), wy = __fmul_rn(w, y), yz = __fmul_rn(y, z);
    const float s_yz = fmaf(y, y, zz), s_xz = fmaf(x, x, zz);
    const Vec3 a1 = Vec3{__fsub_rn(1.f, __fadd_rn(s_yz, s_yz)), __fmul_rn(2.f, fmaf(x, y, wz)),
                         __fmul_rn(2.f, fmaf(x, z, -wy))};
    const Vec3 a2 = Vec3{__fmul_rn(2.f, fmaf(x, y, -wz)), __fsub_rn(1.f, __fadd_rn(s_xz, s_xz)),
                         __fmul_rn(2.f, fmaf(w, x, yz))};
    const float ell = __fmul_rn(3.0f, glob_scale);
    const float r1 = __fmul_rn(ell, s1), r2 = __fmul_rn(ell, s2);
    float lo_x = 0.f, lo_y = 0.f, hi_x = 0.f, hi_y = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float sa = (k & 2) ? -r1 : r1, sb = (k & 1) ? -r2 : r2;
        const Vec3 c = Vec3{fmaf(a2.x, sb, fmaf(a1.x, sa, m.x)), fmaf(a2.y, sb, fmaf(a1.y, sa, m.y)),
                            fmaf(a2.z, sb, fmaf(a1.z, sa, m.z))};
        Vec3 cv = xform_point_ref(vm, c);
        cv.z = fmaxf(cv.z, T_NEAR);
        const float2 p = pinhole_ref(fx, fy, cx, cy, cv);
        if (k == 0) {
            lo_x = hi_x = p.x;
            lo_y = hi_y = p.y;
        } else {
            lo_x = fminf(lo_x, p.x);
            hi_x = fmaxf(hi_x, p.x);
            lo_y = fminf(lo_y, p.y);
            hi_y = fmaxf(hi_y, p.y);
        }
    }
    if (clipped) {
        center = pinhole_ref(fx, fy, cx, cy, pv);
        extent = float2{0.f, 0.f};
    } else {
        center = float2{__fmul_rn(0.5f, __fadd_rn(hi_x, lo_x)), __fmul_rn(0.5f, __fadd_rn(hi_y, lo_y))};
        extent = float2{__fmul_rn(0.5f, __fsub_rn(hi_x, lo_x)), __fmul_rn(0.5f, __fsub_rn(hi_y, lo_y))};
    }
}

__global__ void __launch_bounds__(256) aabb_kernel(int n, const float *__restrict__ means,
                                                   const float *__restrict__ scales, float glob_scale,
                                                   const float4 *__restrict__ quats,
                                                   const float *__restrict__ viewmat, float fx, float fy,
                                                   float cx, float cy, float2 *__restrict__ centers,
                                                   float2 *__restrict__ extents) {
    __shared__ float vm[12];
    if (threadIdx.x < 12) vm[threadIdx.x] = viewmat[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 c, e;
    float depth;
    surfel_aabb(ld3(means + 3 * i), scales[3 * i], scales[3 * i + 1], glob_scale, quats[i], vm, fx, fy, cx, cy,
                c, e, depth);
    centers[i] = c;
    extents[i] = e;
}

// get_num_tiles_hit_2d, gstex_cuda/get_aabb_2d.py:70-92: floor((c -/+ e)/bw [+1]) clamped, fp32 ops in
// the order torch evaluates them.
__global__ void __launch_bounds__(256) tiles_hit_kernel(int n, const float2 *__restrict__ centers,
                                                        const float2 *__restrict__ extents, int tiles_x,
                                                        int tiles_y, float fbw,
                                                        int32_t *__restrict__ num_tiles_hit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 c = centers[i], e = extents[i];
    const int l = min(max((int)floorf(__fdiv_rn(__fsub_rn(c.x, e.x), fbw)), 0), tiles_x);
    const int t = min(max((int)floorf(__fdiv_rn(__fsub_rn(c.y, e.y), fbw)), 0), tiles_y);
    const int r = min(max((int)floorf(__fadd_rn(__fdiv_rn(__fadd_rn(c.x, e.x), fbw), 1.f)), 0), tiles_x);
    const int b = min(max((int)floorf(__fadd_rn(__fdiv_rn(__fadd_rn(c.y, e.y), fbw), 1.f)), 0), tiles_y);
    num_tiles_hit[i] = (r - l) * (b - t);
}

__global__ void __launch_bounds__(256) project_points_kernel(int n, const float *__restrict__ means,
                                                             const float *__restrict__ viewmat, float fx,
                                                             float fy, float cx, float cy,
                                                             float2 *__restrict__ pix,
                                                             float *__restrict__ depths) {
    __shared__ float vm[12];
    if (threadIdx.x < 12) vm[threadIdx.x] = viewmat[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec3 p = ld3(means + 3 * i);
    Vec3 pv = xform_point(vm, p);
    pv.z = view_depth_ref(vm, p);
    if (pix) pix[i] = pinhole(fx, fy, cx, cy, pv);
    depths[i] = pv.z;
}

__global__ void __launch_bounds__(256) project_aabb_count_kernel(
    int n, const float *__restrict__ means, const float *__restrict__ scales, float glob_scale,
    const float4 *__restrict__ quats, const float *__restrict__ viewmat, float fx, float fy, float cx, float cy,
    int tiles_x, int tiles_y, float fbw, float2 *__restrict__ centers, float2 *__restrict__ extents,
    float *__restrict__ depths, int32_t *__restrict__ num_tiles_hit, float *__restrict__ visible_count) {
    __shared__ float vm[12];
    if (threadIdx.x < 12) vm[threadIdx.x] = viewmat[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 c, e;
    float depth;
    surfel_aabb(ld3(means + 3 * i), scales[3 * i], scales[3 * i + 1], glob_scale, quats[i], vm, fx, fy, cx, cy,
                c, e, depth);
    centers[i] = c;
    extents[i] = e;
    depths[i] = depth;
    int cnt = 0;
    if (!(e.x <= 1e-4 && e.y <= 1e-4)) {  // same predicate as the key emitter (reference forward.cu:32)
        int x0, y0, x1, y1;
        tile_bbox(c.x, c.y, e.x, e.y, tiles_x, tiles_y, fbw, x0, y0, x1, y1);
        cnt = (x1 - x0) * (y1 - y0);
    }
    num_tiles_hit[i] = cnt;
    if (visible_count && cnt > 0) visible_count[i] += 1.f;  // one thread per Gaussian, views in stream order: no atomic
}

}  // namespace gstex

using namespace gstex;

extern "C" int gstex_get_aabb_2d(int n, const float *means, const float *scales, float glob_scale,
                                 const float *quats, const float *viewmat, float fx, float fy, float cx,
                                 float cy, float *centers, float *extents, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "get_aabb_2d: n = %d", n);
    if (n == 0) return GSTEX_OK;
    aabb_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, means, scales, glob_scale,
                                                                 (const float4 *)quats, viewmat, fx, fy, cx, cy,
                                                                 (float2 *)centers, (float2 *)extents);
    GSTEX_LAUNCH_OK("aabb_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_num_tiles_hit_2d(int n, const float *centers, const float *extents, int img_height,
                                      int img_width, int block_width, int32_t *num_tiles_hit,
                                      gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0 && block_width > 0, GSTEX_E_INVALID, "num_tiles_hit_2d: n = %d, bw = %d", n, block_width);
    if (n == 0) return GSTEX_OK;
    const int tx = ceil_div(img_width, block_width), ty = ceil_div(img_height, block_width);
    tiles_hit_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        n, (const float2 *)centers, (const float2 *)extents, tx, ty, (float)block_width, num_tiles_hit);
    GSTEX_LAUNCH_OK("tiles_hit_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_project_points(int n, const float *means, const float *viewmat, float fx, float fy,
                                    float cx, float cy, float *pix, float *depths, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "project_points: n = %d", n);
    if (n == 0) return GSTEX_OK;
    project_points_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, means, viewmat, fx, fy, cx, cy,
                                                                           (float2 *)pix, depths);
    GSTEX_LAUNCH_OK("project_points_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_project_aabb_count(int n, const float *means, const float *scales, float glob_scale,
                                        const float *quats, const float *viewmat, float fx, float fy, float cx,
                                        float cy, int img_height, int img_width, int block_width,
                                        float *centers, float *extents, float *depths,
                                        int32_t *num_tiles_hit, float *visible_count, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0 && block_width > 0, GSTEX_E_INVALID, "project_aabb_count: n = %d, bw = %d", n,
                  block_width);
    if (n == 0) return GSTEX_OK;
    const int tx = ceil_div(img_width, block_width), ty = ceil_div(img_height, block_width);
    project_aabb_count_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        n, means, scales, glob_scale, (const float4 *)quats, viewmat, fx, fy, cx, cy, tx, ty, (float)block_width,
        (float2 *)centers, (float2 *)extents, depths, num_tiles_hit, visible_count);
    GSTEX_LAUNCH_OK("project_aabb_count_kernel");
    return GSTEX_OK;
}
