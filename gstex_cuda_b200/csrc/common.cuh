// common.cuh -- shared device helpers, error plumbing and the packed per-view record layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/gstex_b200.h"

namespace gstex {

// ------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, never throws)
// ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define GSTEX_REQUIRE(cond, code, ...)      \
    do {                                    \
        if (!(cond)) {                      \
            ::gstex::set_error(__VA_ARGS__); \
            return (code);                  \
        }                                   \
    } while (0)

#define GSTEX_CUDA_OK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ::gstex::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return GSTEX_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define GSTEX_LAUNCH_OK(name)                                                                       \
    do {                                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess) {                                                                   \
            ::gstex::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return GSTEX_E_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// per-device "dynamic shared memory configured" flags of one kernel (see configure_dynamic_smem in util.cu)
struct SmemOnceFlags {
    static constexpr int MAX_DEVICES = 64;
    std::atomic<bool> done[MAX_DEVICES];
};
int configure_dynamic_smem(const void *fn, size_t bytes, bool max_carveout, SmemOnceFlags &flags);

static inline cudaStream_t as_stream(gstex_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------
struct Vec3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ Vec3 mk3(float x, float y, float z) { return Vec3{x, y, z}; }
__device__ __forceinline__ Vec3 ld3(const float *p) { return Vec3{p[0], p[1], p[2]}; }
__device__ __forceinline__ float dot3(Vec3 a, Vec3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ Vec3 axpy3(float s, Vec3 a, Vec3 b) {  // s*a + b
    return Vec3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)};
}
__device__ __forceinline__ Vec3 add3(Vec3 a, Vec3 b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3 sub3(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 scale3(float s, Vec3 a) { return Vec3{s * a.x, s * a.y, s * a.z}; }

// rows 0..2 of a row-major 4x4: R p + t  (reference helpers.cuh:124-131)
__device__ __forceinline__ Vec3 xform_point(const float *m, Vec3 p) {
    return Vec3{m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
// R p (rotation block only)
__device__ __forceinline__ Vec3 rot_apply(const float *m, Vec3 p) {
    return Vec3{fmaf(m[0], p.x, fmaf(m[1], p.y, m[2] * p.z)), fmaf(m[4], p.x, fmaf(m[5], p.y, m[6] * p.z)),
                fmaf(m[8], p.x, fmaf(m[9], p.y, m[10] * p.z))};
}
// R^T p (reference helpers.cuh:114-121)
__device__ __forceinline__ Vec3 rot_apply_t(const float *m, Vec3 p) {
    return Vec3{fmaf(m[0], p.x, fmaf(m[4], p.y, m[8] * p.z)), fmaf(m[1], p.x, fmaf(m[5], p.y, m[9] * p.z)),
                fmaf(m[2], p.x, fmaf(m[6], p.y, m[10] * p.z))};
}

// columns of R(q), q = (w,x,y,z) unit quaternion (reference helpers.cuh:166-185)
__device__ __forceinline__ void surfel_axes(const float4 q, Vec3 &a1, Vec3 &a2, Vec3 &a3) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    a1 = Vec3{1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y)};
    a2 = Vec3{2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x)};
    a3 = Vec3{2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y)};
}

// VJP of surfel_axes (reference helpers.cuh:187-228); g1..g3 = gradients of the three columns
__device__ __forceinline__ float4 surfel_axes_vjp(const float4 q, Vec3 g1, Vec3 g2, Vec3 g3) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    float4 v;
    v.x = 2.f * (x * (g2.z - g3.y) + y * (g3.x - g1.z) + z * (g1.y - g2.x));
    v.y = 2.f * (-2.f * x * (g2.y + g3.z) + y * (g1.y + g2.x) + z * (g1.z + g3.x) + w * (g2.z - g3.y));
    v.z = 2.f * (x * (g1.y + g2.x) - 2.f * y * (g1.x + g3.z) + z * (g2.z + g3.y) + w * (g3.x - g1.z));
    v.w = 2.f * (x * (g1.z + g3.x) + y * (g2.z + g3.y) - 2.f * z * (g1.x + g2.y) + w * (g1.y - g2.x));
    return v;
}

// pinhole projection with the reference's +1e-6 on z (helpers.cuh:145-152)
__device__ __forceinline__ float2 pinhole(float fx, float fy, float cx, float cy, Vec3 pv) {
    const float rw = 1.f / (pv.z + 1e-6f);
    return float2{(pv.x * rw) * fx + cx, (pv.y * rw) * fy + cy};
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tile bbox of a screen AABB, C truncation + clamp (reference helpers.cuh:37-51, 75-92)
__device__ __forceinline__ void tile_bbox(float cx, float cy, float ex, float ey, int tiles_x, int tiles_y,
                                          float fbw, int &x0, int &y0, int &x1, int &y1) {
    const float tcx = __fdiv_rn(cx, fbw), tcy = __fdiv_rn(cy, fbw);
    const float tex = __fdiv_rn(ex, fbw), tey = __fdiv_rn(ey, fbw);
    x0 = min(max(0, (int)(__fsub_rn(tcx, tex))), tiles_x);
    x1 = min(max(0, (int)(__fadd_rn(__fadd_rn(tcx, tex), 1.f))), tiles_x);
    y0 = min(max(0, (int)(__fsub_rn(tcy, tey))), tiles_y);
    y1 = min(max(0, (int)(__fadd_rn(__fadd_rn(tcy, tey), 1.f))), tiles_y);
}

// torus variant (reference helpers.cuh:53-73, 94-111): no clamping, a box that starts at or before tile 0 grows by one
__device__ __forceinline__ void tile_bbox_wrapped(float cx, float cy, float ex, float ey, float fbw, int &x0, int &y0,
                                                  int &x1, int &y1) {
    const float tcx = __fdiv_rn(cx, fbw), tcy = __fdiv_rn(cy, fbw);
    const float tex = __fdiv_rn(ex, fbw), tey = __fdiv_rn(ey, fbw);
    x0 = (int)(__fsub_rn(tcx, tex));
    if (x0 <= 0) x0 -= 1;
    x1 = (int)(__fadd_rn(__fadd_rn(tcx, tex), 1.f));
    y0 = (int)(__fsub_rn(tcy, tey));
    if (y0 <= 0) y0 -= 1;
    y1 = (int)(__fadd_rn(__fadd_rn(tcy, tey), 1.f));
}

// ------------------------------------------------------------------------------------------
// Packed per-view Gaussian record: 32 floats = 128 B = one cache line, eight 16-byte quads.
// Quads 0-3 feed the alpha test of every (pixel, Gaussian) pair, quads 4-7 only the blend.
// See DESIGN.md section 3 for the algebra; tests/formulation.py is the float64 model.
// ------------------------------------------------------------------------------------------
constexpr int REC_FLOATS = 32;
enum RecSlot : int {
    R_XC = 0, R_YC = 1, R_C0 = 2, R_OPAC = 3,      // expansion centre (pixels), n.(mean-o), opacity
    R_P1X = 4, R_P1Y = 5, R_C1 = 6, R_C3 = 7,      // N1 = c1 + P1.e ; c3 = constant of D
    R_P2X = 8, R_P2Y = 9, R_C2 = 10, R_GID = 11,   // N2 = c2 + P2.e ; Gaussian id (int bits)
    R_A3X = 12, R_A3Y = 13, R_TEXH = 14, R_TEXW = 15,  // D = c3 + A3.e ; texture height / width (int bits)
    R_PUX = 16, R_PUY = 17, R_CU = 18, R_U0 = 19,  // u = u0 + (cu + PU.e)/D
    R_PVX = 20, R_PVY = 21, R_CV = 22, R_V0 = 23,
    R_CR = 24, R_CG = 25, R_CB = 26, R_TEX0 = 27,  // colour, first texel (int bits)
    R_NX = 28, R_NY = 29, R_NZ = 30, R_PAD = 31,   // world normal
};

// Per-view, per-Gaussian gradient moments written by the backward rasteriser (32 floats, 8 quads).
enum AccSlot : int {
    A_G1X = 0, A_G1Y = 1, A_G1C = 2, A_C0 = 3,
    A_G2X = 4, A_G2Y = 5, A_G2C = 6, A_OPAC = 7,
    A_G3X = 8, A_G3Y = 9, A_G3C = 10, A_PAD0 = 11,
    A_GUX = 12, A_GUY = 13, A_GUC = 14, A_U0 = 15,
    A_GVX = 16, A_GVY = 17, A_GVC = 18, A_V0 = 19,
    A_CR = 20, A_CG = 21, A_CB = 22, A_PAD1 = 23,
    A_NX = 24, A_NY = 25, A_NZ = 26, A_PAD2 = 27,
    A_MX = 28, A_MY = 29, A_PAD3 = 30, A_PAD4 = 31,
};
constexpr int ACC_FLOATS = 32;

constexpr float K_SIGMA = 0.84932180028801904f;  // sqrt(0.5 * log2(e)): alpha = opac * 2^-(l1'^2 + l2'^2)
constexpr float LN2_F = 0.69314718055994531f;
constexpr float T_NEAR = 0.01f;
constexpr float T_FAR = 1000.0f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float ALPHA_CAP = 0.99f;
constexpr float T_STOP = 1e-4f;

}  // namespace gstex
