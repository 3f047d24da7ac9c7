// ygl_math.cuh — fp32 vector algebra with a FIXED evaluation order.
//
// Parity contract (SURVEY.md §7 hard part 1): the reference CPU renderer is compiled
// without FMA contraction, so every product/sum below is a separately rounded IEEE
// operation, written in exactly the association order of the reference expression
// it stands for (cited per function, paths under libs/yocto/). This translation unit is
// compiled with -fmad=false (device) and -ffp-contract=off (host); division and sqrt
// are IEEE-exact (-prec-div/-prec-sqrt defaults). Transcendentals restate glibc's float
// routines bit for bit on device (ygl_glibm.cuh, DESIGN.md "libm").
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define YGL_HD __host__ __device__ __forceinline__
#define YGL_D __device__ __forceinline__
// fp64 libm bodies are large: one out-of-line copy per kernel keeps the shading kernels inside the
// instruction cache (the first profile showed k_shade stalled on instruction fetch).
#define YGL_HD_NOINLINE __host__ __device__ __noinline__ inline
// YGL_OUTLINE (build option): the large shared pieces of the shading code - IEEE division and square root
// expansions, texture fetch, material evaluation, the BSDF dispatchers - get one out-of-line copy per kernel instead
// of one per call site (k_shade<path>: 30.7 K -> 19.8 K SASS instructions; its profile is dominated by instruction
// fetch stalls). Pure code layout: every operation and its rounding are unchanged.
#ifdef YGL_OUTLINE
#define YGL_D_BIG static __device__ __noinline__
#define YGL_HD_BIG __host__ __device__ __noinline__ inline
#else
#define YGL_D_BIG __device__ __forceinline__
#define YGL_HD_BIG __host__ __device__ __forceinline__
#endif
#else
#define YGL_HD_BIG inline
#define YGL_HD inline
#define YGL_HD_NOINLINE inline
#endif
#ifdef __CUDACC__
#include "ygl_glibm.cuh"
#endif

namespace ygl {

constexpr float kPi     = 3.14159265358979323846f;  // pif, yocto_math.h:73
constexpr float kFltMax = 3.402823466e+38f;         // flt_max, yocto_math.h:77
constexpr float kRayEps = 1e-4f;                    // ray_eps, yocto_geometry.h:125

struct f2 {
  float x, y;
};
struct f3 {
  float x, y, z;
};
struct f4 {
  float x, y, z, w;
};
struct frame3 {
  f3 x, y, z, o;
};
struct mat3 {
  f3 x, y, z;
};

// ---- scalar helpers with yocto's NaN behaviour (yocto_math.h:1045-1050) ----
YGL_HD float yabs(float a) { return a < 0 ? -a : a; }
YGL_HD float ymin(float a, float b) { return (a < b) ? a : b; }
YGL_HD float ymax(float a, float b) { return (a > b) ? a : b; }
YGL_HD float yclamp(float a, float lo, float hi) { return ymin(ymax(a, lo), hi); }
YGL_HD int   imin(int a, int b) { return (a < b) ? a : b; }
YGL_HD int   imax(int a, int b) { return (a > b) ? a : b; }
YGL_HD int   iclamp(int a, int lo, int hi) { return imin(imax(a, lo), hi); }
YGL_HD bool  yfinite(float a) { return isfinite(a); }

// ---- transcendental functions ----
// Device: the glibc 2.39 x86-64 float routines restated bit for bit (ygl_glibm.cuh): sinf/cosf/
// expf/logf/powf as their FMA builds (what glibc's ifunc selects on FMA-capable hosts), atanf/acosf/
// atan2f as the plain fdlibm kernels — verified exhaustively against the host libm
// (tools/libm_check_host.cpp: 0 mismatches over every float input in the covered ranges).
// expf/logf/powf cover every finite operand, including negative bases with integer exponents,
// subnormals and the overflow/underflow results; what is left (|x| >= 120 for sin/cos, zero/inf/nan
// operands, where fp64 is exact) falls back to fp64 evaluation rounded once. Host: the host libm itself.
#if defined(__CUDA_ARCH__) && defined(YGL_OUTLINE)
static __device__ __noinline__ float ysqrt(float a) { return sqrtf(a); }  // IEEE exact, one copy per kernel
#else
YGL_HD float ysqrt(float a) { return sqrtf(a); }  // IEEE exact
#endif
#ifdef __CUDA_ARCH__
#define YGL_LIBM1(name, impl, fallback)            \
  YGL_HD_NOINLINE float name(float a) {            \
    float r;                                       \
    if (glibm::impl(a, &r)) return r;              \
    return (float)fallback((double)a);             \
  }
YGL_LIBM1(ysin, sinf_<true>, sin)
YGL_LIBM1(ycos, cosf_<true>, cos)
YGL_LIBM1(yexp, expf_<true>, exp)
YGL_LIBM1(ylog, logf_<true>, log)
YGL_LIBM1(yatan, atanf_<false>, atan)
YGL_LIBM1(yacos, acosf_<false>, acos)
#undef YGL_LIBM1
YGL_HD_NOINLINE float yatan2(float a, float b) {
  float r;
  if (glibm::atan2f_<false>(a, b, &r)) return r;
  return (float)atan2((double)a, (double)b);
}
YGL_HD_NOINLINE float ypow(float a, float b) {
  float r;
  if (glibm::powf_<true>(a, b, &r)) return r;
  return (float)pow((double)a, (double)b);
}
#else
YGL_HD float ysin(float a) { return sinf(a); }
YGL_HD float ycos(float a) { return cosf(a); }
YGL_HD float yatan(float a) { return atanf(a); }
YGL_HD float yacos(float a) { return acosf(a); }
YGL_HD float yatan2(float a, float b) { return atan2f(a, b); }
YGL_HD float ylog(float a) { return logf(a); }
YGL_HD float yexp(float a) { return expf(a); }
YGL_HD float ypow(float a, float b) { return powf(a, b); }
#endif
YGL_HD float yfmod(float a, float b) { return fmodf(a, b); }  // exact by definition

// ---- f3 operators (yocto_math.h:1254-1292) ----
YGL_HD f3 make3(float x, float y, float z) { return f3{x, y, z}; }
YGL_HD f3 operator-(const f3& a) { return {-a.x, -a.y, -a.z}; }
YGL_HD f3 operator+(const f3& a, const f3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
YGL_HD f3 operator+(const f3& a, float b) { return {a.x + b, a.y + b, a.z + b}; }
YGL_HD f3 operator+(float a, const f3& b) { return {a + b.x, a + b.y, a + b.z}; }
YGL_HD f3 operator-(const f3& a, const f3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
YGL_HD f3 operator-(const f3& a, float b) { return {a.x - b, a.y - b, a.z - b}; }
YGL_HD f3 operator-(float a, const f3& b) { return {a - b.x, a - b.y, a - b.z}; }
YGL_HD f3 operator*(const f3& a, const f3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
YGL_HD f3 operator*(const f3& a, float b) { return {a.x * b, a.y * b, a.z * b}; }
YGL_HD f3 operator*(float a, const f3& b) { return {a * b.x, a * b.y, a * b.z}; }
#if defined(__CUDA_ARCH__) && defined(YGL_OUTLINE)
// IEEE division expands to ~13 instructions per quotient and the shading kernels hold hundreds of them
static __device__ __noinline__ f3 ydiv3(float ax, float ay, float az, float b) { return {ax / b, ay / b, az / b}; }
static __device__ __noinline__ float ydiv1(float a, float b) { return a / b; }
YGL_HD f3 operator/(const f3& a, const f3& b) { return {ydiv1(a.x, b.x), ydiv1(a.y, b.y), ydiv1(a.z, b.z)}; }
YGL_HD f3 operator/(const f3& a, float b) { return ydiv3(a.x, a.y, a.z, b); }
YGL_HD f3 operator/(float a, const f3& b) { return {ydiv1(a, b.x), ydiv1(a, b.y), ydiv1(a, b.z)}; }
#else
YGL_HD f3 operator/(const f3& a, const f3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
YGL_HD f3 operator/(const f3& a, float b) { return {a.x / b, a.y / b, a.z / b}; }
YGL_HD f3 operator/(float a, const f3& b) { return {a / b.x, a / b.y, a / b.z}; }
#endif
YGL_HD bool operator==(const f3& a, const f3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
YGL_HD bool is_zero(const f3& a) { return a.x == 0 && a.y == 0 && a.z == 0; }

YGL_HD f2 operator+(const f2& a, const f2& b) { return {a.x + b.x, a.y + b.y}; }
YGL_HD f2 operator*(const f2& a, float b) { return {a.x * b, a.y * b}; }
YGL_HD f2 operator-(float a, const f2& b) { return {a - b.x, a - b.y}; }

YGL_HD f4 operator+(const f4& a, const f4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
YGL_HD f4 operator*(const f4& a, float b) { return {a.x * b, a.y * b, a.z * b, a.w * b}; }
YGL_HD f3 xyz(const f4& a) { return {a.x, a.y, a.z}; }

// yocto_math.h:1302-1312
YGL_HD float dot(const f3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
YGL_HD f3    cross(const f3& a, const f3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
YGL_HD float length(const f3& a) { return ysqrt(dot(a, a)); }
YGL_HD f3    normalize(const f3& a) {  // yocto_math.h:1314-1317
  auto l = length(a);
  return (l != 0) ? a / l : a;
}
YGL_HD float distance_squared(const f3& a, const f3& b) { return dot(a - b, a - b); }
YGL_HD f3 orthonormalize(const f3& a, const f3& b) { return normalize(a - b * dot(a, b)); }  // :1331
YGL_HD f3 reflect(const f3& w, const f3& n) { return -w + 2 * dot(n, w) * n; }               // :1336
YGL_HD f3 refract(const f3& w, const f3& n, float inv_eta) {                                 // :1339
  auto cosine = dot(n, w);
  auto k      = 1 + inv_eta * inv_eta * (cosine * cosine - 1);
  if (k < 0) return {0, 0, 0};
  return -w * inv_eta + (inv_eta * cosine - ysqrt(k)) * n;
}

YGL_HD float max3(const f3& a) { return ymax(ymax(a.x, a.y), a.z); }  // yocto_math.h:1369
YGL_HD float min3(const f3& a) { return ymin(ymin(a.x, a.y), a.z); }
YGL_HD float sum3(const f3& a) { return a.x + a.y + a.z; }
YGL_HD float mean3(const f3& a) { return sum3(a) / 3; }
YGL_HD f3    vmin(const f3& a, const f3& b) { return {ymin(a.x, b.x), ymin(a.y, b.y), ymin(a.z, b.z)}; }
YGL_HD f3    vmax(const f3& a, const f3& b) { return {ymax(a.x, b.x), ymax(a.y, b.y), ymax(a.z, b.z)}; }
YGL_HD f3    vclamp(const f3& a, float lo, float hi) {
  return {yclamp(a.x, lo, hi), yclamp(a.y, lo, hi), yclamp(a.z, lo, hi)};
}
YGL_HD f3   vsqrt(const f3& a) { return {ysqrt(a.x), ysqrt(a.y), ysqrt(a.z)}; }
YGL_HD f3   vexp(const f3& a) { return {yexp(a.x), yexp(a.y), yexp(a.z)}; }
YGL_HD f3   vlog(const f3& a) { return {ylog(a.x), ylog(a.y), ylog(a.z)}; }
YGL_HD bool vfinite(const f3& a) { return yfinite(a.x) && yfinite(a.y) && yfinite(a.z); }
YGL_HD f3   lerp3(const f3& a, const f3& b, float u) { return a * (1 - u) + b * u; }  // :1362
YGL_HD f4   lerp4(const f4& a, const f4& b, float u) { return a * (1 - u) + b * u; }  // :1512
YGL_HD float comp(const f3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---- frames and 3x3 matrices ----
// transform_point / transform_vector / transform_direction, yocto_math.h:2263-2271
YGL_HD f3 transform_point(const frame3& a, const f3& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.o;
}
YGL_HD f3 transform_vector(const frame3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
YGL_HD f3 transform_direction(const frame3& a, const f3& b) { return normalize(transform_vector(a, b)); }
// transform_normal(frame, n, non_rigid = false), yocto_math.h:2272-2279
YGL_HD f3 transform_normal(const frame3& a, const f3& b) { return normalize(transform_vector(a, b)); }
YGL_HD f3 mat_mul(const mat3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :1942

// inverse(frame, non_rigid): yocto_math.h:2114-2122 with adjoint/determinant :1965-1972
YGL_HD frame3 frame_inverse(const frame3& a, bool non_rigid) {
  mat3 minv;
  if (non_rigid) {
    auto cyz = cross(a.y, a.z), czx = cross(a.z, a.x), cxy = cross(a.x, a.y);
    auto det = dot(a.x, cross(a.y, a.z));
    auto idt = 1 / det;
    // adjoint = transpose({cyz, czx, cxy}); inverse = adjoint * (1/det)
    minv.x = f3{cyz.x, czx.x, cxy.x} * idt;
    minv.y = f3{cyz.y, czx.y, cxy.y} * idt;
    minv.z = f3{cyz.z, czx.z, cxy.z} * idt;
  } else {
    minv.x = {a.x.x, a.y.x, a.z.x};
    minv.y = {a.x.y, a.y.y, a.z.y};
    minv.z = {a.x.z, a.y.z, a.z.z};
  }
  auto o = -mat_mul(minv, a.o);
  return {minv.x, minv.y, minv.z, o};
}

// basis_fromz, yocto_math.h:1977-1986 (Duff et al. orthonormal basis)
YGL_HD mat3 basis_fromz(const f3& v) {
  auto z    = normalize(v);
  auto sign = copysignf(1.0f, z.z);
  auto a    = -1.0f / (sign + z.z);
  auto b    = z.x * z.y * a;
  auto x    = f3{1.0f + sign * z.x * z.x * a, sign * b, -sign * z.x};
  auto y    = f3{b, sign + z.y * z.y * a, -z.y};
  return {x, y, z};
}
// transform_direction(mat3, v), yocto_math.h:2236
YGL_HD f3 transform_direction(const mat3& a, const f3& b) { return normalize(mat_mul(a, b)); }

}  // namespace ygl
