// oracle/ref_lobes.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Function-level oracle of the shading code: the reference's own BSDF dispatchers are `static` in yocto_trace.cpp
// (libs/yocto/yocto_trace.cpp:166-335), so this translation unit includes that file where it lies and exports batch
// forms of them. Built by oracle/Makefile into oracle/_ref/libyocto_ref_lobes.so (which therefore contains its own
// copy of yocto_trace.cpp instead of yocto_trace.o). One record per call:
//   in : type, color[3], roughness, metallic, ior, density[3], scattering[3], scanisotropy, normal[3], outgoing[3],
//        incoming[3], rnl, rn[2]                                                           (26 floats, type as float)
//   out: eval_bsdfcos[3], sample_bsdfcos[3], sample_bsdfcos_pdf, eval_delta[3], sample_delta[3], sample_delta_pdf,
//        eval_scattering[3], sample_scattering[3], sample_scattering_pdf, bsdfcos/pdf at the SAMPLED direction [4]   (25)
#include <yocto/yocto_trace.cpp>

extern "C" void ref_lobes(const float* in, long long n, float* out) {
  using namespace yocto;
  for (long long k = 0; k < n; k++) {
    const float* a = in + 26 * k;
    float*       o = out + 25 * k;
    auto m         = material_point{};
    m.type         = (material_type)(int)a[0];
    m.color        = {a[1], a[2], a[3]};
    m.roughness    = a[4];
    m.metallic     = a[5];
    m.ior          = a[6];
    m.density      = {a[7], a[8], a[9]};
    m.scattering   = {a[10], a[11], a[12]};
    m.scanisotropy = a[13];
    auto normal = vec3f{a[14], a[15], a[16]}, outgoing = vec3f{a[17], a[18], a[19]}, incoming = vec3f{a[20], a[21], a[22]};
    auto rnl = a[23];
    auto rn  = vec2f{a[24], a[25]};
    auto put3 = [&](int at, const vec3f& v) { o[at] = v.x, o[at + 1] = v.y, o[at + 2] = v.z; };
    put3(0, eval_bsdfcos(m, normal, outgoing, incoming));
    auto sampled = sample_bsdfcos(m, normal, outgoing, rnl, rn);
    put3(3, sampled);
    o[6] = sample_bsdfcos_pdf(m, normal, outgoing, incoming);
    put3(7, eval_delta(m, normal, outgoing, incoming));
    auto dsampled = sample_delta(m, normal, outgoing, rnl);
    put3(10, dsampled);
    o[13] = sample_delta_pdf(m, normal, outgoing, incoming);
    put3(14, eval_scattering(m, outgoing, incoming));
    put3(17, sample_scattering(m, outgoing, rnl, rn));
    o[20] = sample_scattering_pdf(m, outgoing, incoming);
    // what the integrator computes next for the sampled direction (yocto_trace.cpp:528-529 / :540-541)
    if (m.roughness != 0) {
      put3(21, eval_bsdfcos(m, normal, outgoing, sampled));
      o[24] = sample_bsdfcos_pdf(m, normal, outgoing, sampled);
    } else {
      put3(21, eval_delta(m, normal, outgoing, dsampled));
      o[24] = sample_delta_pdf(m, normal, outgoing, dsampled);
    }
  }
}
