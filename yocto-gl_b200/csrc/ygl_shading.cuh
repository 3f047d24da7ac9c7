// ygl_shading.cuh — Fresnel, GGX microfacet terms, BSDF lobes, transmittance and phase function.
// Behavioural contract: libs/yocto/yocto_shading.h:303-1111 and the material dispatch of
// libs/yocto/yocto_trace.cpp:166-335. Operation order follows the reference expressions.
#pragma once

#include "ygl_sampling.cuh"

namespace ygl {

enum : int {
  kMatte = 0, kGlossy, kReflective, kTransparent, kRefractive, kSubsurface, kVolumetric, kGltfPbr
};

// material_point, yocto_scene.h:258-270
struct mpoint {
  int   type;
  f3    emission, color;
  float opacity, roughness, metallic, ior;
  f3    density, scattering;
  float scanisotropy, trdepth;
};

YGL_HD bool same_hemisphere(const f3& n, const f3& o, const f3& i) {  // shading.h:303
  return dot(n, o) * dot(n, i) >= 0;
}
YGL_HD f3 up_normal_of(const f3& n, const f3& o) { return dot(n, o) <= 0 ? -n : n; }

// fresnel_schlick, shading.h:309-315
// pow(x, 2.0f) in the reference (yocto_shading.h:907,952): g++ -O2 expands a constant exponent of 2 to
// x * x, so the reference binary never calls powf there (glibc's powf(x, 2) differs from x * x by one ulp on
// 0.07 % of inputs). Other exponents (5, 2.4, ...) stay libm calls.
YGL_HD float ysqr(float a) { return a * a; }
YGL_HD f3 fresnel_schlick(const f3& specular, const f3& normal, const f3& outgoing) {
  if (is_zero(specular)) return {0, 0, 0};
  auto cosine = dot(normal, outgoing);
  return specular + (1 - specular) * ypow(yclamp(1 - yabs(cosine), 0.0f, 1.0f), 5.0f);
}
// fresnel_dielectric, shading.h:318-338
YGL_HD float fresnel_dielectric(float eta, const f3& normal, const f3& outgoing) {
  auto cosw  = yabs(dot(normal, outgoing));
  auto sin2  = 1 - cosw * cosw;
  auto eta2  = eta * eta;
  auto cos2t = 1 - sin2 / eta2;
  if (cos2t < 0) return 1;
  auto t0 = ysqrt(cos2t);
  auto t1 = eta * t0;
  auto t2 = eta * cosw;
  auto rs = (cosw - t1) / (cosw + t1);
  auto rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2;
}
// fresnel_conductor, shading.h:341-366
YGL_HD f3 fresnel_conductor(const f3& eta, const f3& etak, const f3& normal, const f3& outgoing) {
  auto cosw = dot(normal, outgoing);
  if (cosw <= 0) return {0, 0, 0};
  cosw          = yclamp(cosw, (float)-1, (float)1);
  auto cos2     = cosw * cosw;
  auto sin2     = yclamp(1 - cos2, (float)0, (float)1);
  auto eta2     = eta * eta;
  auto etak2    = etak * etak;
  auto t0       = eta2 - etak2 - sin2;
  auto a2plusb2 = vsqrt(t0 * t0 + 4 * eta2 * etak2);
  auto t1       = a2plusb2 + cos2;
  auto a        = vsqrt((a2plusb2 + t0) / 2);
  auto t2       = 2 * a * cosw;
  auto rs       = (t1 - t2) / (t1 + t2);
  auto t3       = cos2 * a2plusb2 + sin2 * sin2;
  auto t4       = t2 * sin2;
  auto rp       = rs * (t3 - t4) / (t3 + t4);
  return (rp + rs) / 2;
}
// eta_to_reflectivity / reflectivity_to_eta, shading.h:369-376
YGL_HD f3 eta_to_reflectivity(const f3& eta) { return ((eta - 1) * (eta - 1)) / ((eta + 1) * (eta + 1)); }
YGL_HD f3 reflectivity_to_eta(const f3& reflectivity_) {
  auto reflectivity = vclamp(reflectivity_, 0.0f, 0.99f);
  return (1 + vsqrt(reflectivity)) / (1 - vsqrt(reflectivity));
}

// microfacet_distribution (GGX branch), shading.h:409-425
YGL_HD float microfacet_distribution(float roughness, const f3& normal, const f3& halfway) {
  auto cosine = dot(normal, halfway);
  if (cosine <= 0) return 0;
  auto roughness2 = roughness * roughness;
  auto cosine2    = cosine * cosine;
  return roughness2 /
         (kPi * (cosine2 * roughness2 + 1 - cosine2) * (cosine2 * roughness2 + 1 - cosine2));
}
// microfacet_shadowing1 / shadowing (GGX), shading.h:428-456
YGL_HD float microfacet_shadowing1(float roughness, const f3& normal, const f3& halfway, const f3& direction) {
  auto cosine  = dot(normal, direction);
  auto cosineh = dot(halfway, direction);
  if (cosine * cosineh <= 0) return 0;
  auto roughness2 = roughness * roughness;
  auto cosine2    = cosine * cosine;
  return 2 * yabs(cosine) / (yabs(cosine) + ysqrt(cosine2 - roughness2 * cosine2 + roughness2));
}
YGL_HD float microfacet_shadowing(float roughness, const f3& normal, const f3& halfway, const f3& outgoing,
    const f3& incoming) {
  return microfacet_shadowing1(roughness, normal, halfway, outgoing) *
         microfacet_shadowing1(roughness, normal, halfway, incoming);
}
// sample_microfacet(roughness, normal, rn) (GGX), shading.h:459-472
YGL_HD f3 sample_microfacet(float roughness, const f3& normal, const f2& rn) {
  auto phi   = 2 * kPi * rn.x;
  auto theta = yatan(roughness * ysqrt(rn.y / (1 - rn.y)));
  auto local = f3{ycos(phi) * ysin(theta), ysin(phi) * ysin(theta), ycos(theta)};
  return transform_direction(basis_fromz(normal), local);
}
// sample_microfacet_pdf, shading.h:475-480
YGL_HD float sample_microfacet_pdf(float roughness, const f3& normal, const f3& halfway) {
  auto cosine = dot(normal, halfway);
  if (cosine < 0) return 0;
  return microfacet_distribution(roughness, normal, halfway) * cosine;
}

// ---- matte, shading.h:554-574 ----
YGL_HD f3 eval_matte(const f3& color, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return {0, 0, 0};
  return color / kPi * yabs(dot(n, i));
}
YGL_HD f3 sample_matte(const f3& n, const f3& o, const f2& rn) {
  return sample_hemisphere_cos(up_normal_of(n, o), rn);
}
YGL_HD float sample_matte_pdf(const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  return sample_hemisphere_cos_pdf(up_normal_of(n, o), i);
}

// ---- glossy, shading.h:577-619 ----
YGL_HD f3 eval_glossy(const f3& color, float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return {0, 0, 0};
  auto up      = up_normal_of(n, o);
  auto F1      = fresnel_dielectric(ior, up, o);
  auto halfway = normalize(i + o);
  auto F       = fresnel_dielectric(ior, halfway, i);
  auto D       = microfacet_distribution(roughness, up, halfway);
  auto G       = microfacet_shadowing(roughness, up, halfway, o, i);
  return color * (1 - F1) / kPi * yabs(dot(up, i)) +
         f3{1, 1, 1} * F * D * G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
}
YGL_HD f3 sample_glossy(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto up = up_normal_of(n, o);
  if (rnl < fresnel_dielectric(ior, up, o)) {
    auto halfway  = sample_microfacet(roughness, up, rn);
    auto incoming = reflect(o, halfway);
    if (!same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  } else {
    return sample_hemisphere_cos(up, rn);
  }
}
YGL_HD float sample_glossy_pdf(float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  auto up      = up_normal_of(n, o);
  auto halfway = normalize(o + i);
  auto F       = fresnel_dielectric(ior, up, o);
  return F * sample_microfacet_pdf(roughness, up, halfway) / (4 * yabs(dot(o, halfway))) +
         (1 - F) * sample_hemisphere_cos_pdf(up, i);
}

// ---- reflective (rough), shading.h:622-654 ----
YGL_HD f3 eval_reflective(const f3& color, float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return {0, 0, 0};
  auto up      = up_normal_of(n, o);
  auto halfway = normalize(i + o);
  auto F       = fresnel_conductor(reflectivity_to_eta(color), {0, 0, 0}, halfway, i);
  auto D       = microfacet_distribution(roughness, up, halfway);
  auto G       = microfacet_shadowing(roughness, up, halfway, o, i);
  return F * D * G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
}
YGL_HD f3 sample_reflective(float roughness, const f3& n, const f3& o, const f2& rn) {
  auto up       = up_normal_of(n, o);
  auto halfway  = sample_microfacet(roughness, up, rn);
  auto incoming = reflect(o, halfway);
  if (!same_hemisphere(up, o, incoming)) return {0, 0, 0};
  return incoming;
}
YGL_HD float sample_reflective_pdf(float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  auto up      = up_normal_of(n, o);
  auto halfway = normalize(o + i);
  return sample_microfacet_pdf(roughness, up, halfway) / (4 * yabs(dot(o, halfway)));
}
// ---- reflective (delta), shading.h:693-712 ----
YGL_HD f3 eval_reflective_delta(const f3& color, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return {0, 0, 0};
  auto up = up_normal_of(n, o);
  return fresnel_conductor(reflectivity_to_eta(color), {0, 0, 0}, up, o);
}
YGL_HD f3 sample_reflective_delta(const f3& n, const f3& o) { return reflect(o, up_normal_of(n, o)); }
YGL_HD float sample_reflective_delta_pdf(const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  return 1;
}

// ---- gltfpbr, shading.h:736-788 ----
YGL_HD f3 eval_gltfpbr(const f3& color, float ior, float roughness, float metallic, const f3& n, const f3& o,
    const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return {0, 0, 0};
  auto reflectivity = lerp3(eta_to_reflectivity(f3{ior, ior, ior}), color, metallic);
  auto up           = up_normal_of(n, o);
  auto F1           = fresnel_schlick(reflectivity, up, o);
  auto halfway      = normalize(i + o);
  auto F            = fresnel_schlick(reflectivity, halfway, i);
  auto D            = microfacet_distribution(roughness, up, halfway);
  auto G            = microfacet_shadowing(roughness, up, halfway, o, i);
  return color * (1 - metallic) * (1 - F1) / kPi * yabs(dot(up, i)) +
         F * D * G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
}
YGL_HD f3 sample_gltfpbr(const f3& color, float ior, float roughness, float metallic, const f3& n, const f3& o,
    float rnl, const f2& rn) {
  auto up           = up_normal_of(n, o);
  auto reflectivity = lerp3(eta_to_reflectivity(f3{ior, ior, ior}), color, metallic);
  if (rnl < mean3(fresnel_schlick(reflectivity, up, o))) {
    auto halfway  = sample_microfacet(roughness, up, rn);
    auto incoming = reflect(o, halfway);
    if (!same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  } else {
    return sample_hemisphere_cos(up, rn);
  }
}
YGL_HD float sample_gltfpbr_pdf(const f3& color, float ior, float roughness, float metallic, const f3& n,
    const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  auto up           = up_normal_of(n, o);
  auto halfway      = normalize(o + i);
  auto reflectivity = lerp3(eta_to_reflectivity(f3{ior, ior, ior}), color, metallic);
  auto F            = mean3(fresnel_schlick(reflectivity, up, o));
  return F * sample_microfacet_pdf(roughness, up, halfway) / (4 * yabs(dot(o, halfway))) +
         (1 - F) * sample_hemisphere_cos_pdf(up, i);
}

// ---- transparent (rough), shading.h:791-849 ----
YGL_HD f3 eval_transparent(const f3& color, float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto up = up_normal_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    auto F       = fresnel_dielectric(ior, halfway, o);
    auto D       = microfacet_distribution(roughness, up, halfway);
    auto G       = microfacet_shadowing(roughness, up, halfway, o, i);
    return f3{1, 1, 1} * F * D * G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
  } else {
    auto reflected = reflect(-i, up);
    auto halfway   = normalize(reflected + o);
    auto F         = fresnel_dielectric(ior, halfway, o);
    auto D         = microfacet_distribution(roughness, up, halfway);
    auto G         = microfacet_shadowing(roughness, up, halfway, o, reflected);
    return color * (1 - F) * D * G / (4 * dot(up, o) * dot(up, reflected)) * (yabs(dot(up, reflected)));
  }
}
YGL_HD f3 sample_transparent(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto up      = up_normal_of(n, o);
  auto halfway = sample_microfacet(roughness, up, rn);
  if (rnl < fresnel_dielectric(ior, halfway, o)) {
    auto incoming = reflect(o, halfway);
    if (!same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  } else {
    auto reflected = reflect(o, halfway);
    auto incoming  = -reflect(reflected, up);
    if (same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  }
}
YGL_HD float sample_transparent_pdf(float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto up = up_normal_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    return fresnel_dielectric(ior, halfway, o) * sample_microfacet_pdf(roughness, up, halfway) /
           (4 * yabs(dot(o, halfway)));
  } else {
    auto reflected = reflect(-i, up);
    auto halfway   = normalize(reflected + o);
    auto d = (1 - fresnel_dielectric(ior, halfway, o)) * sample_microfacet_pdf(roughness, up, halfway);
    return d / (4 * yabs(dot(o, halfway)));
  }
}
// ---- transparent (delta), shading.h:852-881 ----
YGL_HD f3 eval_transparent_delta(const f3& color, float ior, const f3& n, const f3& o, const f3& i) {
  auto up = up_normal_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    return f3{1, 1, 1} * fresnel_dielectric(ior, up, o);
  } else {
    return color * (1 - fresnel_dielectric(ior, up, o));
  }
}
YGL_HD f3 sample_transparent_delta(float ior, const f3& n, const f3& o, float rnl) {
  auto up = up_normal_of(n, o);
  if (rnl < fresnel_dielectric(ior, up, o)) {
    return reflect(o, up);
  } else {
    return -o;
  }
}
YGL_HD float sample_transparent_delta_pdf(float ior, const f3& n, const f3& o, const f3& i) {
  auto up = up_normal_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    return fresnel_dielectric(ior, up, o);
  } else {
    return 1 - fresnel_dielectric(ior, up, o);
  }
}

// ---- refractive (rough), shading.h:884-957 ----
YGL_HD f3 eval_refractive(float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    auto F       = fresnel_dielectric(rel_ior, halfway, o);
    auto D       = microfacet_distribution(roughness, up, halfway);
    auto G       = microfacet_shadowing(roughness, up, halfway, o, i);
    return f3{1, 1, 1} * F * D * G / yabs(4 * dot(n, o) * dot(n, i)) * yabs(dot(n, i));
  } else {
    auto halfway = -normalize(rel_ior * i + o) * (entering ? 1.0f : -1.0f);
    auto F       = fresnel_dielectric(rel_ior, halfway, o);
    auto D       = microfacet_distribution(roughness, up, halfway);
    auto G       = microfacet_shadowing(roughness, up, halfway, o, i);
    return f3{1, 1, 1} * yabs((dot(o, halfway) * dot(i, halfway)) / (dot(o, n) * dot(i, n))) * (1 - F) * D *
           G / ysqr(rel_ior * dot(halfway, i) + dot(halfway, o)) * yabs(dot(n, i));
  }
}
YGL_HD f3 sample_refractive(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto halfway  = sample_microfacet(roughness, up, rn);
  if (rnl < fresnel_dielectric(entering ? ior : (1 / ior), halfway, o)) {
    auto incoming = reflect(o, halfway);
    if (!same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  } else {
    auto incoming = refract(o, halfway, entering ? (1 / ior) : ior);
    if (same_hemisphere(up, o, incoming)) return {0, 0, 0};
    return incoming;
  }
}
YGL_HD float sample_refractive_pdf(float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    return fresnel_dielectric(rel_ior, halfway, o) * sample_microfacet_pdf(roughness, up, halfway) /
           (4 * yabs(dot(o, halfway)));
  } else {
    auto halfway = -normalize(rel_ior * i + o) * (entering ? 1.0f : -1.0f);
    return (1 - fresnel_dielectric(rel_ior, halfway, o)) * sample_microfacet_pdf(roughness, up, halfway) *
           yabs(dot(halfway, i)) / ysqr(rel_ior * dot(halfway, i) + dot(halfway, o));
  }
}
// ---- refractive (delta), shading.h:960-1005; `abs(ior - 1) < 1e-3` compares in double ----
YGL_HD bool ior_is_one(float ior) { return (double)yabs(ior - 1) < 1e-3; }
YGL_HD f3 eval_refractive_delta(float ior, const f3& n, const f3& o, const f3& i) {
  if (ior_is_one(ior)) return dot(n, i) * dot(n, o) <= 0 ? f3{1, 1, 1} : f3{0, 0, 0};
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    return f3{1, 1, 1} * fresnel_dielectric(rel_ior, up, o);
  } else {
    return f3{1, 1, 1} * (1 / (rel_ior * rel_ior)) * (1 - fresnel_dielectric(rel_ior, up, o));
  }
}
YGL_HD f3 sample_refractive_delta(float ior, const f3& n, const f3& o, float rnl) {
  if (ior_is_one(ior)) return -o;
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (rnl < fresnel_dielectric(rel_ior, up, o)) {
    return reflect(o, up);
  } else {
    return refract(o, up, 1 / rel_ior);
  }
}
YGL_HD float sample_refractive_delta_pdf(float ior, const f3& n, const f3& o, const f3& i) {
  if (ior_is_one(ior)) return dot(n, i) * dot(n, o) < 0 ? 1.0f : 0.0f;
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    return fresnel_dielectric(rel_ior, up, o);
  } else {
    return (1 - fresnel_dielectric(rel_ior, up, o));
  }
}
// ---- passthrough, shading.h:1028-1048 ----
YGL_HD f3 eval_passthrough(const f3& n, const f3& o, const f3& i) {
  return (dot(n, i) * dot(n, o) >= 0) ? f3{0, 0, 0} : f3{1, 1, 1};
}
YGL_HD float sample_passthrough_pdf(const f3& n, const f3& o, const f3& i) {
  return (dot(n, i) * dot(n, o) >= 0) ? 0.0f : 1.0f;
}

// ---- volumes, shading.h:1056-1111 ----
YGL_HD f3 eval_transmittance(const f3& density, float distance) { return vexp(-density * distance); }
YGL_HD float sample_transmittance(const f3& density, float max_distance, float rl, float rd) {
  auto channel  = iclamp((int)(rl * 3), 0, 2);
  auto dc       = comp(density, channel);
  auto distance = (dc == 0) ? kFltMax : -ylog(1 - rd) / dc;
  return ymin(distance, max_distance);
}
YGL_HD float sample_transmittance_pdf(const f3& density, float distance, float max_distance) {
  if (distance < max_distance) {
    return sum3(density * vexp(-density * distance)) / 3;
  } else {
    return sum3(vexp(-density * max_distance)) / 3;
  }
}
YGL_HD float eval_phasefunction(float anisotropy, const f3& o, const f3& i) {
  auto cosine = -dot(o, i);
  auto denom  = 1 + anisotropy * anisotropy - 2 * anisotropy * cosine;
  return (1 - anisotropy * anisotropy) / (4 * kPi * denom * ysqrt(denom));
}
YGL_HD f3 sample_phasefunction(float anisotropy, const f3& o, const f2& rn) {
  auto cos_theta = 0.0f;
  if (yabs(anisotropy) < 1e-3f) {
    cos_theta = 1 - 2 * rn.y;
  } else {
    auto square = (1 - anisotropy * anisotropy) / (1 + anisotropy - 2 * anisotropy * rn.y);
    cos_theta   = (1 + anisotropy * anisotropy - square * square) / (2 * anisotropy);
  }
  auto sin_theta = ysqrt(ymax(0.0f, 1 - cos_theta * cos_theta));
  auto phi       = 2 * kPi * rn.x;
  auto local     = f3{sin_theta * ycos(phi), sin_theta * ysin(phi), cos_theta};
  return mat_mul(basis_fromz(-o), local);
}

// ---- material dispatch, yocto_trace.cpp:166-335 ----
YGL_HD bool is_delta(const mpoint& m) {  // yocto_scene.cpp:263-271
  return (m.type == kReflective && m.roughness == 0) || (m.type == kRefractive && m.roughness == 0) ||
         (m.type == kTransparent && m.roughness == 0) || (m.type == kVolumetric);
}
YGL_HD f3 eval_emission(const mpoint& m, const f3& n, const f3& o) {
  return dot(n, o) >= 0 ? m.emission : f3{0, 0, 0};
}
YGL_HD_BIG f3 eval_bsdfcos(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness == 0) return {0, 0, 0};
  switch (m.type) {
    case kMatte: return eval_matte(m.color, n, o, i);
    case kGlossy: return eval_glossy(m.color, m.ior, m.roughness, n, o, i);
    case kReflective: return eval_reflective(m.color, m.roughness, n, o, i);
    case kTransparent: return eval_transparent(m.color, m.ior, m.roughness, n, o, i);
    case kRefractive:
    case kSubsurface: return eval_refractive(m.ior, m.roughness, n, o, i);
    case kGltfPbr: return eval_gltfpbr(m.color, m.ior, m.roughness, m.metallic, n, o, i);
    default: return {0, 0, 0};
  }
}
YGL_HD f3 eval_delta(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness != 0) return {0, 0, 0};
  switch (m.type) {
    case kReflective: return eval_reflective_delta(m.color, n, o, i);
    case kTransparent: return eval_transparent_delta(m.color, m.ior, n, o, i);
    case kRefractive: return eval_refractive_delta(m.ior, n, o, i);
    case kVolumetric: return eval_passthrough(n, o, i);
    default: return {0, 0, 0};
  }
}
YGL_HD_BIG f3 sample_bsdfcos(const mpoint& m, const f3& n, const f3& o, float rnl, const f2& rn) {
  if (m.roughness == 0) return {0, 0, 0};
  switch (m.type) {
    case kMatte: return sample_matte(n, o, rn);
    case kGlossy: return sample_glossy(m.ior, m.roughness, n, o, rnl, rn);
    case kReflective: return sample_reflective(m.roughness, n, o, rn);
    case kTransparent: return sample_transparent(m.ior, m.roughness, n, o, rnl, rn);
    case kRefractive:
    case kSubsurface: return sample_refractive(m.ior, m.roughness, n, o, rnl, rn);
    case kGltfPbr: return sample_gltfpbr(m.color, m.ior, m.roughness, m.metallic, n, o, rnl, rn);
    default: return {0, 0, 0};
  }
}
YGL_HD f3 sample_delta(const mpoint& m, const f3& n, const f3& o, float rnl) {
  if (m.roughness != 0) return {0, 0, 0};
  switch (m.type) {
    case kReflective: return sample_reflective_delta(n, o);
    case kTransparent: return sample_transparent_delta(m.ior, n, o, rnl);
    case kRefractive: return sample_refractive_delta(m.ior, n, o, rnl);
    case kVolumetric: return -o;  // sample_passthrough
    default: return {0, 0, 0};
  }
}
YGL_HD_BIG float sample_bsdfcos_pdf(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness == 0) return 0;
  switch (m.type) {
    case kMatte: return sample_matte_pdf(n, o, i);
    case kGlossy: return sample_glossy_pdf(m.ior, m.roughness, n, o, i);
    case kReflective: return sample_reflective_pdf(m.roughness, n, o, i);
    case kTransparent: return sample_transparent_pdf(m.ior, m.roughness, n, o, i);
    case kRefractive:
    case kSubsurface: return sample_refractive_pdf(m.ior, m.roughness, n, o, i);
    case kGltfPbr: return sample_gltfpbr_pdf(m.color, m.ior, m.roughness, m.metallic, n, o, i);
    default: return 0;
  }
}
YGL_HD float sample_delta_pdf(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness != 0) return 0;
  switch (m.type) {
    case kReflective: return sample_reflective_delta_pdf(n, o, i);
    case kTransparent: return sample_transparent_delta_pdf(m.ior, n, o, i);
    case kRefractive: return sample_refractive_delta_pdf(m.ior, n, o, i);
    case kVolumetric: return sample_passthrough_pdf(n, o, i);
    default: return 0;
  }
}
// volume scattering dispatch, yocto_trace.cpp:316-335 (vsdf = density, scattering, scanisotropy)
struct vsdf_t {
  f3    density, scattering;
  float scanisotropy;
};
YGL_HD f3 eval_scattering(const vsdf_t& v, const f3& o, const f3& i) {
  if (is_zero(v.density)) return {0, 0, 0};
  return v.scattering * v.density * eval_phasefunction(v.scanisotropy, o, i);
}
YGL_HD f3 sample_scattering(const vsdf_t& v, const f3& o, const f2& rn) {
  if (is_zero(v.density)) return {0, 0, 0};
  return sample_phasefunction(v.scanisotropy, o, rn);
}
YGL_HD float sample_scattering_pdf(const vsdf_t& v, const f3& o, const f3& i) {
  if (is_zero(v.density)) return 0;
  return eval_phasefunction(v.scanisotropy, o, i);
}

}  // namespace ygl
