// ygl_sampling.cuh — PCG32 streams and the Monte-Carlo warps of the path integrator.
// Behavioural contract: libs/yocto/yocto_sampling.h:81-232 (rng) and :252-398 (warps).
#pragma once

#include "ygl_math.cuh"

namespace ygl {

// rng_state, yocto_sampling.h:81-87 (16 B)
struct rng_t {
  uint64_t state, inc;
};

// _advance_rng, yocto_sampling.h:187-194: PCG32 XSH-RR
YGL_HD uint32_t rng_next(rng_t& rng) {
  uint64_t old = rng.state;
  rng.state    = old * 6364136223846793005ULL + rng.inc;
  uint32_t xs  = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}

// make_rng, yocto_sampling.h:197-205
YGL_HD rng_t rng_make(uint64_t seed, uint64_t seq) {
  rng_t rng{0, (seq << 1u) | 1u};
  rng_next(rng);
  rng.state += seed;
  rng_next(rng);
  return rng;
}

// rand1i / rand1f / rand2f, yocto_sampling.h:208-226
YGL_HD int rand1i(rng_t& rng, int n) { return (int)(rng_next(rng) % (uint32_t)n); }
YGL_HD float rand1f(rng_t& rng) {
  uint32_t u = (rng_next(rng) >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u) - 1.0f;
#else
  float f;
  memcpy(&f, &u, 4);
  return f - 1.0f;
#endif
}
YGL_HD f2 rand2f(rng_t& rng) {
  auto x = rand1f(rng);
  auto y = rand1f(rng);
  return {x, y};
}
YGL_HD f3 rand3f(rng_t& rng) {
  auto x = rand1f(rng);
  auto y = rand1f(rng);
  auto z = rand1f(rng);
  return {x, y, z};
}

// sample_hemisphere_cos(normal, ruv) + pdf, yocto_sampling.h:296-306
YGL_HD f3 sample_hemisphere_cos(const f3& normal, const f2& ruv) {
  auto z     = ysqrt(ruv.y);
  auto r     = ysqrt(1 - z * z);
  auto phi   = 2 * kPi * ruv.x;
  auto local = f3{r * ycos(phi), r * ysin(phi), z};
  return transform_direction(basis_fromz(normal), local);
}
YGL_HD float sample_hemisphere_cos_pdf(const f3& normal, const f3& direction) {
  auto cosw = dot(normal, direction);
  return (cosw <= 0) ? 0 : cosw / kPi;
}
// sample_sphere, yocto_sampling.h:277-282
YGL_HD f3 sample_sphere(const f2& ruv) {
  auto z   = 2 * ruv.y - 1;
  auto r   = ysqrt(yclamp(1 - z * z, 0.0f, 1.0f));
  auto phi = 2 * kPi * ruv.x;
  return {r * ycos(phi), r * ysin(phi), z};
}
// sample_disk, yocto_sampling.h:336-340
YGL_HD f2 sample_disk(const f2& ruv) {
  auto r   = ysqrt(ruv.y);
  auto phi = 2 * kPi * ruv.x;
  return {ycos(phi) * r, ysin(phi) * r};
}
// sample_triangle (barycentric), yocto_sampling.h:351-353
YGL_HD f2 sample_triangle(const f2& ruv) { return {1 - ysqrt(ruv.x), ruv.y * ysqrt(ruv.x)}; }
// sample_uniform / pdf, yocto_sampling.h:371-374
YGL_HD int   sample_uniform(int size, float r) { return iclamp((int)(r * size), 0, size - 1); }
YGL_HD float sample_uniform_pdf(int size) { return (float)1 / (float)size; }

// sample_discrete over a cdf, yocto_sampling.h:388-393 (std::upper_bound = first element > r)
YGL_HD int sample_discrete(const float* cdf, int n, float r) {
  float last = cdf[n - 1];
  r          = yclamp(r * last, (float)0, last - (float)0.00001);
  int lo = 0, count = n;
  while (count > 0) {
    int step = count / 2, mid = lo + step;
    if (!(r < cdf[mid])) {
      lo = mid + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return iclamp(lo, 0, n - 1);
}
// sample_discrete_pdf, yocto_sampling.h:395-398
YGL_HD float sample_discrete_pdf(const float* cdf, int idx) {
  if (idx == 0) return cdf[0];
  return cdf[idx] - cdf[idx - 1];
}

}  // namespace ygl
