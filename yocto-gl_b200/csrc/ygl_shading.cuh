// ygl_shading.cuh — surface and volume scattering of the path tracer: one FUSED routine per material that returns
// the cosine-weighted BSDF value AND the sampling pdf of a direction pair together (the integrator always wants both:
// yocto_trace.cpp:528-529, :700-706, :892), plus the matching direction samplers.
//
// Behavioural contract: libs/yocto/yocto_shading.h:303-1111 (Fresnel, GGX distribution / masking, the lobes of the
// eight material types, transmittance, Henyey-Greenstein phase function) and the dispatch of
// libs/yocto/yocto_trace.cpp:166-335. Tolerance 0: every returned float equals the reference's bit for bit
// (tests/test_lobes.py: 3.2 M random and edge-case direction pairs per run against the reference's own dispatchers).
// That contract fixes the arithmetic of each formula — operand order, where a product is rounded, which branch
// returns +0 — but not the program around it. What is shared here and computed once per call (the reference
// evaluates it separately in eval_* and sample_*_pdf): the upward normal, the half vector, the Fresnel terms, the
// GGX distribution D; `lobe_t` carries value and pdf out together. Identical subexpressions are pure functions of
// the same inputs, so sharing them cannot change a bit.
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

// value of bsdf * |cos| and the pdf with which the samplers below produce the incoming direction
struct lobe_t {
  f3    bsdfcos;
  float pdf;
};
YGL_HD lobe_t no_lobe() { return {{0, 0, 0}, 0}; }

// pow(x, 2.0f) in the reference (yocto_shading.h:907,952): g++ -O2 expands a constant exponent of 2 to x * x, so the
// reference binary never calls powf there (glibc's powf(x, 2) differs from x * x by one ulp on 0.07 % of inputs).
YGL_HD float ysqr(float a) { return a * a; }

// ---------------------------------------------------------------------------------------------------------------
// Fresnel (yocto_shading.h:309-376)
// ---------------------------------------------------------------------------------------------------------------
YGL_HD f3 fresnel_schlick(const f3& specular, const f3& normal, const f3& outgoing) {
  if (is_zero(specular)) return {0, 0, 0};
  auto cosine = dot(normal, outgoing);
  return specular + (1 - specular) * ypow(yclamp(1 - yabs(cosine), 0.0f, 1.0f), 5.0f);
}
YGL_HD float fresnel_dielectric(float eta, const f3& normal, const f3& outgoing) {
  auto cosw  = yabs(dot(normal, outgoing));
  auto sin2  = 1 - cosw * cosw;
  auto eta2  = eta * eta;
  auto cos2t = 1 - sin2 / eta2;
  if (cos2t < 0) return 1;  // total internal reflection
  auto t0 = ysqrt(cos2t);
  auto t1 = eta * t0;
  auto t2 = eta * cosw;
  auto rs = (cosw - t1) / (cosw + t1);
  auto rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2;
}
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
YGL_HD f3 eta_to_reflectivity(const f3& eta) { return ((eta - 1) * (eta - 1)) / ((eta + 1) * (eta + 1)); }
YGL_HD f3 reflectivity_to_eta(const f3& reflectivity_) {
  auto reflectivity = vclamp(reflectivity_, 0.0f, 0.99f);
  return (1 + vsqrt(reflectivity)) / (1 - vsqrt(reflectivity));
}
// a metal's Fresnel term from its colour (the reflective lobes)
YGL_HD f3 metal_fresnel(const f3& color, const f3& normal, const f3& direction) {
  return fresnel_conductor(reflectivity_to_eta(color), {0, 0, 0}, normal, direction);
}

// ---------------------------------------------------------------------------------------------------------------
// GGX microfacets (the only distribution the lobes use; yocto_shading.h:409-480)
// ---------------------------------------------------------------------------------------------------------------
YGL_HD float ggx_distribution(float roughness, const f3& normal, const f3& halfway) {
  auto cosine = dot(normal, halfway);
  if (cosine <= 0) return 0;
  auto roughness2 = roughness * roughness;
  auto cosine2    = cosine * cosine;
  return roughness2 / (kPi * (cosine2 * roughness2 + 1 - cosine2) * (cosine2 * roughness2 + 1 - cosine2));
}
YGL_HD float ggx_masking1(float roughness, const f3& normal, const f3& halfway, const f3& direction) {
  auto cosine  = dot(normal, direction);
  auto cosineh = dot(halfway, direction);
  if (cosine * cosineh <= 0) return 0;
  auto roughness2 = roughness * roughness;
  auto cosine2    = cosine * cosine;
  return 2 * yabs(cosine) / (yabs(cosine) + ysqrt(cosine2 - roughness2 * cosine2 + roughness2));
}
// the terms of one microfacet interaction, shared by the value and the pdf of a lobe
struct ggx_t {
  float D;      // distribution of normals at the half vector
  float G;      // masking-shadowing of the pair
  float hpdf;   // pdf of sampling that half vector (sample_microfacet_pdf: D * cos, +0 below the horizon)
};
YGL_HD ggx_t ggx_terms(float roughness, const f3& normal, const f3& halfway, const f3& outgoing, const f3& incoming) {
  ggx_t t;
  t.D = ggx_distribution(roughness, normal, halfway);
  t.G = ggx_masking1(roughness, normal, halfway, outgoing) * ggx_masking1(roughness, normal, halfway, incoming);
  auto cosine = dot(normal, halfway);
  t.hpdf      = cosine < 0 ? 0.0f : t.D * cosine;
  return t;
}
// a half vector drawn from the GGX distribution around `normal` (sample_microfacet, yocto_shading.h:459-472)
YGL_HD f3 ggx_sample(float roughness, const f3& normal, const f2& rn) {
  auto phi   = 2 * kPi * rn.x;
  auto theta = yatan(roughness * ysqrt(rn.y / (1 - rn.y)));
  auto local = f3{ycos(phi) * ysin(theta), ysin(phi) * ysin(theta), ycos(theta)};
  return transform_direction(basis_fromz(normal), local);
}

YGL_HD bool same_side(const f3& n, const f3& o, const f3& i) { return dot(n, o) * dot(n, i) >= 0; }
YGL_HD f3   facing(const f3& n, const f3& o) { return dot(n, o) <= 0 ? -n : n; }  // the normal on the side of `o`

// a reflected direction sampled through a GGX half vector; zero when it leaves the hemisphere of `up`
YGL_HD f3 ggx_reflect(float roughness, const f3& up, const f3& o, const f2& rn) {
  auto halfway  = ggx_sample(roughness, up, rn);
  auto incoming = reflect(o, halfway);
  if (!same_side(up, o, incoming)) return {0, 0, 0};
  return incoming;
}

// ---------------------------------------------------------------------------------------------------------------
// rough lobes: value and pdf together
// ---------------------------------------------------------------------------------------------------------------
// matte: Lambert (yocto_shading.h:554-574)
YGL_HD lobe_t matte_lobe(const f3& color, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return no_lobe();
  return {color / kPi * yabs(dot(n, i)), sample_hemisphere_cos_pdf(facing(n, o), i)};
}
// glossy: dielectric coat over Lambert (yocto_shading.h:577-619)
YGL_HD lobe_t glossy_lobe(const f3& color, float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return no_lobe();
  auto up      = facing(n, o);
  auto F_view  = fresnel_dielectric(ior, up, o);  // also the probability of choosing the coat
  auto halfway = normalize(i + o);
  auto F       = fresnel_dielectric(ior, halfway, i);
  auto ggx     = ggx_terms(roughness, up, halfway, o, i);
  lobe_t lobe;
  lobe.bsdfcos = color * (1 - F_view) / kPi * yabs(dot(up, i)) +
                 f3{1, 1, 1} * F * ggx.D * ggx.G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
  lobe.pdf = F_view * ggx.hpdf / (4 * yabs(dot(o, halfway))) + (1 - F_view) * sample_hemisphere_cos_pdf(up, i);
  return lobe;
}
// reflective: rough metal (yocto_shading.h:622-654)
YGL_HD lobe_t reflective_lobe(const f3& color, float roughness, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return no_lobe();
  auto up      = facing(n, o);
  auto halfway = normalize(i + o);
  auto F       = metal_fresnel(color, halfway, i);
  auto ggx     = ggx_terms(roughness, up, halfway, o, i);
  return {F * ggx.D * ggx.G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i)), ggx.hpdf / (4 * yabs(dot(o, halfway)))};
}
// gltfpbr: metallic-roughness (yocto_shading.h:736-788)
YGL_HD lobe_t gltfpbr_lobe(const f3& color, float ior, float roughness, float metallic, const f3& n, const f3& o,
    const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return no_lobe();
  auto reflectivity = lerp3(eta_to_reflectivity(f3{ior, ior, ior}), color, metallic);
  auto up           = facing(n, o);
  auto F_view       = fresnel_schlick(reflectivity, up, o);
  auto halfway      = normalize(i + o);
  auto F            = fresnel_schlick(reflectivity, halfway, i);
  auto ggx          = ggx_terms(roughness, up, halfway, o, i);
  auto pick         = mean3(F_view);  // probability of the specular choice
  lobe_t lobe;
  lobe.bsdfcos = color * (1 - metallic) * (1 - F_view) / kPi * yabs(dot(up, i)) +
                 F * ggx.D * ggx.G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i));
  lobe.pdf = pick * ggx.hpdf / (4 * yabs(dot(o, halfway))) + (1 - pick) * sample_hemisphere_cos_pdf(up, i);
  return lobe;
}
// transparent: thin dielectric sheet, transmission mirrored about the surface (yocto_shading.h:791-849)
YGL_HD lobe_t transparent_lobe(const f3& color, float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto up = facing(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    auto F       = fresnel_dielectric(ior, halfway, o);
    auto ggx     = ggx_terms(roughness, up, halfway, o, i);
    return {f3{1, 1, 1} * F * ggx.D * ggx.G / (4 * dot(up, o) * dot(up, i)) * yabs(dot(up, i)),
        F * ggx.hpdf / (4 * yabs(dot(o, halfway)))};
  }
  auto mirrored = reflect(-i, up);  // the transmitted direction folded back to the side of `o`
  auto halfway  = normalize(mirrored + o);
  auto F        = fresnel_dielectric(ior, halfway, o);
  auto ggx      = ggx_terms(roughness, up, halfway, o, mirrored);
  auto d        = (1 - F) * ggx.hpdf;
  return {color * (1 - F) * ggx.D * ggx.G / (4 * dot(up, o) * dot(up, mirrored)) * (yabs(dot(up, mirrored))),
      d / (4 * yabs(dot(o, halfway)))};
}
// refractive (and subsurface): rough dielectric interface (yocto_shading.h:884-957)
YGL_HD lobe_t refractive_lobe(float ior, float roughness, const f3& n, const f3& o, const f3& i) {
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    auto halfway = normalize(i + o);
    auto F       = fresnel_dielectric(rel_ior, halfway, o);
    auto ggx     = ggx_terms(roughness, up, halfway, o, i);
    return {f3{1, 1, 1} * F * ggx.D * ggx.G / yabs(4 * dot(n, o) * dot(n, i)) * yabs(dot(n, i)),
        F * ggx.hpdf / (4 * yabs(dot(o, halfway)))};
  }
  auto halfway = -normalize(rel_ior * i + o) * (entering ? 1.0f : -1.0f);
  auto F       = fresnel_dielectric(rel_ior, halfway, o);
  auto ggx     = ggx_terms(roughness, up, halfway, o, i);
  auto jacobi  = ysqr(rel_ior * dot(halfway, i) + dot(halfway, o));  // change of variables half vector -> direction
  return {f3{1, 1, 1} * yabs((dot(o, halfway) * dot(i, halfway)) / (dot(o, n) * dot(i, n))) * (1 - F) * ggx.D * ggx.G /
              jacobi * yabs(dot(n, i)),
      (1 - F) * ggx.hpdf * yabs(dot(halfway, i)) / jacobi};
}

// ---------------------------------------------------------------------------------------------------------------
// rough lobes: direction samplers (rnl picks the layer, rn the direction)
// ---------------------------------------------------------------------------------------------------------------
YGL_HD f3 matte_sample(const f3& n, const f3& o, const f2& rn) { return sample_hemisphere_cos(facing(n, o), rn); }
YGL_HD f3 glossy_sample(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto up = facing(n, o);
  if (rnl < fresnel_dielectric(ior, up, o)) return ggx_reflect(roughness, up, o, rn);
  return sample_hemisphere_cos(up, rn);
}
YGL_HD f3 reflective_sample(float roughness, const f3& n, const f3& o, const f2& rn) {
  return ggx_reflect(roughness, facing(n, o), o, rn);
}
YGL_HD f3 gltfpbr_sample(const f3& color, float ior, float roughness, float metallic, const f3& n, const f3& o, float rnl,
    const f2& rn) {
  auto up           = facing(n, o);
  auto reflectivity = lerp3(eta_to_reflectivity(f3{ior, ior, ior}), color, metallic);
  if (rnl < mean3(fresnel_schlick(reflectivity, up, o))) return ggx_reflect(roughness, up, o, rn);
  return sample_hemisphere_cos(up, rn);
}
YGL_HD f3 transparent_sample(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto up        = facing(n, o);
  auto halfway   = ggx_sample(roughness, up, rn);
  auto reflected = reflect(o, halfway);
  if (rnl < fresnel_dielectric(ior, halfway, o)) {
    if (!same_side(up, o, reflected)) return {0, 0, 0};
    return reflected;
  }
  auto through = -reflect(reflected, up);
  if (same_side(up, o, through)) return {0, 0, 0};
  return through;
}
YGL_HD f3 refractive_sample(float ior, float roughness, const f3& n, const f3& o, float rnl, const f2& rn) {
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto halfway  = ggx_sample(roughness, up, rn);
  if (rnl < fresnel_dielectric(entering ? ior : (1 / ior), halfway, o)) {
    auto reflected = reflect(o, halfway);
    if (!same_side(up, o, reflected)) return {0, 0, 0};
    return reflected;
  }
  auto refracted = refract(o, halfway, entering ? (1 / ior) : ior);
  if (same_side(up, o, refracted)) return {0, 0, 0};
  return refracted;
}

// ---------------------------------------------------------------------------------------------------------------
// delta lobes (roughness 0): value and discrete probability together, and their samplers
// (yocto_shading.h:693-712, :852-881, :960-1005, :1028-1048)
// ---------------------------------------------------------------------------------------------------------------
YGL_HD lobe_t mirror_lobe(const f3& color, const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) <= 0) return no_lobe();
  return {metal_fresnel(color, facing(n, o), o), 1};
}
YGL_HD lobe_t thin_glass_lobe(const f3& color, float ior, const f3& n, const f3& o, const f3& i) {
  auto F = fresnel_dielectric(ior, facing(n, o), o);
  if (dot(n, i) * dot(n, o) >= 0) return {f3{1, 1, 1} * F, F};
  return {color * (1 - F), 1 - F};
}
// `abs(ior - 1) < 1e-3` compares in double in the reference (yocto_shading.h:963)
YGL_HD bool index_matched(float ior) { return (double)yabs(ior - 1) < 1e-3; }
YGL_HD lobe_t glass_lobe(float ior, const f3& n, const f3& o, const f3& i) {
  if (index_matched(ior)) {
    // the value tests "<= 0", the probability "< 0" (yocto_shading.h:964, :998)
    return {dot(n, i) * dot(n, o) <= 0 ? f3{1, 1, 1} : f3{0, 0, 0}, dot(n, i) * dot(n, o) < 0 ? 1.0f : 0.0f};
  }
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  auto F        = fresnel_dielectric(rel_ior, up, o);
  if (dot(n, i) * dot(n, o) >= 0) return {f3{1, 1, 1} * F, F};
  return {f3{1, 1, 1} * (1 / (rel_ior * rel_ior)) * (1 - F), (1 - F)};
}
YGL_HD lobe_t passthrough_lobe(const f3& n, const f3& o, const f3& i) {
  if (dot(n, i) * dot(n, o) >= 0) return no_lobe();
  return {{1, 1, 1}, 1};
}
YGL_HD f3 thin_glass_sample(float ior, const f3& n, const f3& o, float rnl) {
  auto up = facing(n, o);
  if (rnl < fresnel_dielectric(ior, up, o)) return reflect(o, up);
  return -o;
}
YGL_HD f3 glass_sample(float ior, const f3& n, const f3& o, float rnl) {
  if (index_matched(ior)) return -o;
  auto entering = dot(n, o) >= 0;
  auto up       = entering ? n : -n;
  auto rel_ior  = entering ? ior : (1 / ior);
  if (rnl < fresnel_dielectric(rel_ior, up, o)) return reflect(o, up);
  return refract(o, up, 1 / rel_ior);
}

// ---------------------------------------------------------------------------------------------------------------
// participating media (yocto_shading.h:1056-1111)
// ---------------------------------------------------------------------------------------------------------------
YGL_HD f3 eval_transmittance(const f3& density, float distance) { return vexp(-density * distance); }
YGL_HD float sample_transmittance(const f3& density, float max_distance, float rl, float rd) {
  auto channel  = iclamp((int)(rl * 3), 0, 2);
  auto dc       = comp(density, channel);
  auto distance = (dc == 0) ? kFltMax : -ylog(1 - rd) / dc;
  return ymin(distance, max_distance);
}
YGL_HD float sample_transmittance_pdf(const f3& density, float distance, float max_distance) {
  if (distance < max_distance) return sum3(density * vexp(-density * distance)) / 3;
  return sum3(vexp(-density * max_distance)) / 3;
}
YGL_HD float henyey_greenstein(float anisotropy, const f3& o, const f3& i) {
  auto cosine = -dot(o, i);
  auto denom  = 1 + anisotropy * anisotropy - 2 * anisotropy * cosine;
  return (1 - anisotropy * anisotropy) / (4 * kPi * denom * ysqrt(denom));
}
YGL_HD f3 henyey_greenstein_sample(float anisotropy, const f3& o, const f2& rn) {
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

// ---------------------------------------------------------------------------------------------------------------
// material dispatch (yocto_trace.cpp:166-335). A material is `delta` when its roughness is zero (or it is a pure
// volume boundary): rough and delta lobes are disjoint, each returns nothing for the other kind.
// ---------------------------------------------------------------------------------------------------------------
YGL_HD bool is_delta(const mpoint& m) {  // yocto_scene.cpp:263-271
  return (m.type == kReflective && m.roughness == 0) || (m.type == kRefractive && m.roughness == 0) ||
         (m.type == kTransparent && m.roughness == 0) || (m.type == kVolumetric);
}
YGL_HD f3 eval_emission(const mpoint& m, const f3& n, const f3& o) { return dot(n, o) >= 0 ? m.emission : f3{0, 0, 0}; }

// eval_bsdfcos + sample_bsdfcos_pdf in one pass
YGL_HD_BIG lobe_t bsdf_lobe(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness == 0) return no_lobe();
  switch (m.type) {
    case kMatte: return matte_lobe(m.color, n, o, i);
    case kGlossy: return glossy_lobe(m.color, m.ior, m.roughness, n, o, i);
    case kReflective: return reflective_lobe(m.color, m.roughness, n, o, i);
    case kTransparent: return transparent_lobe(m.color, m.ior, m.roughness, n, o, i);
    case kRefractive:
    case kSubsurface: return refractive_lobe(m.ior, m.roughness, n, o, i);
    case kGltfPbr: return gltfpbr_lobe(m.color, m.ior, m.roughness, m.metallic, n, o, i);
    default: return no_lobe();
  }
}
YGL_HD_BIG f3 sample_bsdfcos(const mpoint& m, const f3& n, const f3& o, float rnl, const f2& rn) {
  if (m.roughness == 0) return {0, 0, 0};
  switch (m.type) {
    case kMatte: return matte_sample(n, o, rn);
    case kGlossy: return glossy_sample(m.ior, m.roughness, n, o, rnl, rn);
    case kReflective: return reflective_sample(m.roughness, n, o, rn);
    case kTransparent: return transparent_sample(m.ior, m.roughness, n, o, rnl, rn);
    case kRefractive:
    case kSubsurface: return refractive_sample(m.ior, m.roughness, n, o, rnl, rn);
    case kGltfPbr: return gltfpbr_sample(m.color, m.ior, m.roughness, m.metallic, n, o, rnl, rn);
    default: return {0, 0, 0};
  }
}
// eval_delta + sample_delta_pdf in one pass
YGL_HD lobe_t delta_lobe(const mpoint& m, const f3& n, const f3& o, const f3& i) {
  if (m.roughness != 0) return no_lobe();
  switch (m.type) {
    case kReflective: return mirror_lobe(m.color, n, o, i);
    case kTransparent: return thin_glass_lobe(m.color, m.ior, n, o, i);
    case kRefractive: return glass_lobe(m.ior, n, o, i);
    case kVolumetric: return passthrough_lobe(n, o, i);
    default: return no_lobe();
  }
}
YGL_HD f3 sample_delta(const mpoint& m, const f3& n, const f3& o, float rnl) {
  if (m.roughness != 0) return {0, 0, 0};
  switch (m.type) {
    case kReflective: return reflect(o, facing(n, o));
    case kTransparent: return thin_glass_sample(m.ior, n, o, rnl);
    case kRefractive: return glass_sample(m.ior, n, o, rnl);
    case kVolumetric: return -o;  // sample_passthrough
    default: return {0, 0, 0};
  }
}
// the reference's one-quantity entry points, for callers that want only one of the two
YGL_HD f3    eval_bsdfcos(const mpoint& m, const f3& n, const f3& o, const f3& i) { return bsdf_lobe(m, n, o, i).bsdfcos; }
YGL_HD float sample_bsdfcos_pdf(const mpoint& m, const f3& n, const f3& o, const f3& i) { return bsdf_lobe(m, n, o, i).pdf; }
YGL_HD f3    eval_delta(const mpoint& m, const f3& n, const f3& o, const f3& i) { return delta_lobe(m, n, o, i).bsdfcos; }
YGL_HD float sample_delta_pdf(const mpoint& m, const f3& n, const f3& o, const f3& i) { return delta_lobe(m, n, o, i).pdf; }

// volume scattering, yocto_trace.cpp:316-335 (vsdf = density, scattering, scanisotropy)
struct vsdf_t {
  f3    density, scattering;
  float scanisotropy;
};
YGL_HD lobe_t scattering_lobe(const vsdf_t& v, const f3& o, const f3& i) {
  if (is_zero(v.density)) return no_lobe();
  auto phase = henyey_greenstein(v.scanisotropy, o, i);
  return {v.scattering * v.density * phase, phase};
}
YGL_HD f3 sample_scattering(const vsdf_t& v, const f3& o, const f2& rn) {
  if (is_zero(v.density)) return {0, 0, 0};
  return henyey_greenstein_sample(v.scanisotropy, o, rn);
}
YGL_HD f3    eval_scattering(const vsdf_t& v, const f3& o, const f3& i) { return scattering_lobe(v, o, i).bsdfcos; }
YGL_HD float sample_scattering_pdf(const vsdf_t& v, const f3& o, const f3& i) { return scattering_lobe(v, o, i).pdf; }

}  // namespace ygl
