// tests/cpp/lobes_host.cpp — TEST INFRASTRUCTURE. The product's shading header (yocto-gl_b200/csrc/ygl_shading.cuh)
// compiled for the HOST (its functions are __host__ __device__; on the host the transcendental calls are glibc's,
// which the device restates bit for bit: tests/test_gpu_parity.py::test_device_libm_matches_host_glibc), exporting the
// same batch record as oracle/ref_lobes.cpp. tests/test_lobes.py compares the two bit for bit on millions of random
// and edge-case inputs. Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "../../yocto-gl_b200/csrc/ygl_shading.cuh"

extern "C" void ygl_host_lobes(const float* in, long long n, float* out) {
  using namespace ygl;
  for (long long k = 0; k < n; k++) {
    const float* a = in + 26 * k;
    float*       o = out + 25 * k;
    mpoint m       = {};
    m.type         = (int)a[0];
    m.color        = {a[1], a[2], a[3]};
    m.opacity      = 1;
    m.roughness    = a[4];
    m.metallic     = a[5];
    m.ior          = a[6];
    m.density      = {a[7], a[8], a[9]};
    m.scattering   = {a[10], a[11], a[12]};
    m.scanisotropy = a[13];
    vsdf_t v       = {m.density, m.scattering, m.scanisotropy};
    f3 normal = {a[14], a[15], a[16]}, outgoing = {a[17], a[18], a[19]}, incoming = {a[20], a[21], a[22]};
    float rnl = a[23];
    f2    rn  = {a[24], a[25]};
    auto put3 = [&](int at, const f3& x) { o[at] = x.x, o[at + 1] = x.y, o[at + 2] = x.z; };
    put3(0, eval_bsdfcos(m, normal, outgoing, incoming));
    f3 sampled = sample_bsdfcos(m, normal, outgoing, rnl, rn);
    put3(3, sampled);
    o[6] = sample_bsdfcos_pdf(m, normal, outgoing, incoming);
    put3(7, eval_delta(m, normal, outgoing, incoming));
    f3 dsampled = sample_delta(m, normal, outgoing, rnl);
    put3(10, dsampled);
    o[13] = sample_delta_pdf(m, normal, outgoing, incoming);
    put3(14, eval_scattering(v, outgoing, incoming));
    put3(17, sample_scattering(v, outgoing, rn));
    o[20] = sample_scattering_pdf(v, outgoing, incoming);
    if (m.roughness != 0) {
      put3(21, eval_bsdfcos(m, normal, outgoing, sampled));
      o[24] = sample_bsdfcos_pdf(m, normal, outgoing, sampled);
    } else {
      put3(21, eval_delta(m, normal, outgoing, dsampled));
      o[24] = sample_delta_pdf(m, normal, outgoing, dsampled);
    }
  }
}
