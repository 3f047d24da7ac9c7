// Development tool: exhaustive comparison of the restated libm routines (ygl_glibm.cuh, host
// instantiation = same IEEE arithmetic as the device) against this host's glibc, for both the
// FMA-contracted and the plain variant. usage: libm_check_host [threads]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../yocto-gl_b200/csrc/ygl_glibm.cuh"

using namespace ygl::glibm;
typedef bool (*fn_t)(float, float*);
typedef float (*ref_t)(float);

static void run(const char* name, fn_t fn, ref_t ref, int nthreads) {
  std::atomic<uint64_t> bad{0}, covered{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      uint64_t b = 0, c = 0;
      for (uint64_t u = t; u < (1ull << 32); u += nthreads) {
        uint32_t bits = (uint32_t)u;
        float    x, y;
        memcpy(&x, &bits, 4);
        if (!fn(x, &y)) continue;
        c++;
        float    r = ref(x);
        uint32_t a, e;
        memcpy(&a, &y, 4);
        memcpy(&e, &r, 4);
        if (a != e && !(y != y && r != r)) b++;
      }
      bad += b;
      covered += c;
    });
  for (auto& t : th) t.join();
  printf("%-12s covered %llu inputs, mismatches %llu\n", name, (unsigned long long)covered.load(), (unsigned long long)bad.load());
  fflush(stdout);
}

int main(int argc, char** argv) {
  int nt = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
  run("sinf  fma", sinf_<true>, sinf, nt);
  run("sinf  plain", sinf_<false>, sinf, nt);
  run("cosf  fma", cosf_<true>, cosf, nt);
  run("cosf  plain", cosf_<false>, cosf, nt);
  run("expf  fma", expf_<true>, expf, nt);
  run("expf  plain", expf_<false>, expf, nt);
  run("logf  fma", logf_<true>, logf, nt);
  run("logf  plain", logf_<false>, logf, nt);
  run("atanf fma", atanf_<true>, atanf, nt);
  run("atanf plain", atanf_<false>, atanf, nt);
  run("acosf fma", acosf_<true>, acosf, nt);
  run("acosf plain", acosf_<false>, acosf, nt);
  for (float y : {5.0f, 6.0f, 2.2f, 2.4f, 0.75f, 2.0f}) {
    static float yy;
    yy = y;
    char name[32];
    snprintf(name, sizeof(name), "powf(x,%g) fma", y);
    run(name, [](float x, float* o) { return powf_<true>(x, yy, o); }, [](float x) { return powf(x, yy); }, nt);
  }
  // two-argument routines cannot be enumerated: random pairs, half of them arbitrary bit patterns, half in the
  // renderer's ranges (unit-range operands for atan2f; bases in [0, 4), exponents in [0, 8) for powf)
  {
    std::atomic<uint64_t> bad_atan2{0}, bad_pow{0}, covered{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t]() {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1), b0 = 0, b1 = 0, c = 0;
        for (uint64_t i = 0; i < 30000000ull; i++) {
          s ^= s << 13, s ^= s >> 7, s ^= s << 17;
          float x, y, o;
          if (i & 1) {
            uint32_t a = (uint32_t)s, b = (uint32_t)(s >> 32);
            memcpy(&x, &a, 4), memcpy(&y, &b, 4);
          } else {
            x = ((int32_t)(uint32_t)s) / 2147483648.0f, y = ((int32_t)(uint32_t)(s >> 32)) / 2147483648.0f;
          }
          if (i & 2) {
            if (!atan2f_<false>(y, x, &o)) continue;
            float e = atan2f(y, x);
            c++;
            if (memcmp(&o, &e, 4) && !(o != o && e != e)) b0++;
          } else {
            float bx = (i & 4) ? x * 4 : fabsf(x) * 4, ey = fabsf(y) * 8;
            if (!powf_<true>(bx, ey, &o)) continue;
            float e = powf(bx, ey);
            c++;
            if (memcmp(&o, &e, 4) && !(o != o && e != e)) b1++;
          }
        }
        bad_atan2 += b0, bad_pow += b1, covered += c;
      });
    for (auto& t : th) t.join();
    printf("atan2f plain / powf(x,y) fma: covered %llu random pairs, mismatches %llu / %llu\n",
        (unsigned long long)covered.load(), (unsigned long long)bad_atan2.load(), (unsigned long long)bad_pow.load());
  }
  return 0;
}
