// oracle/shim_demo.cpp — TEST INFRASTRUCTURE. A reference-side C++ program using the drop-in shim
// (yocto-gl_b200/host/yocto_b200trace.h): renders make_cornellbox() with yocto::trace_image (CPU
// reference) and yocto::b200::trace_image (libygl_b200.so) and prints the comparison. Built by
// oracle/Makefile into oracle/_ref/shim_demo where the reference headers exist; run on the GPU box
// by tests/test_gpu_parity.py::test_reference_side_shim_runs.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../yocto-gl_b200/host/yocto_b200trace.h"

int main(int argc, char** argv) {
  auto scene        = yocto::make_cornellbox();
  auto params       = yocto::trace_params{};
  params.resolution = argc > 1 ? atoi(argv[1]) : 64;
  params.samples    = argc > 2 ? atoi(argv[2]) : 4;
  params.bounces    = 4;
  try {
    auto a = yocto::trace_image(scene, params);
    auto b = yocto::b200::trace_image(scene, params);
    if (a.width != b.width || a.height != b.height) return printf("size mismatch\n"), 1;
    double se = 0;
    size_t exact = 0, flipped = 0;
    for (size_t i = 0; i < a.pixels.size(); i++) {
      auto &p = a.pixels[i], &q = b.pixels[i];
      double d[3] = {(double)p.x - q.x, (double)p.y - q.y, (double)p.z - q.z};
      double m    = std::fmax(std::fabs(d[0]), std::fmax(std::fabs(d[1]), std::fabs(d[2])));
      if (m > 1e-4) flipped++;
      else se += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (memcmp(&p, &q, sizeof(p)) == 0) exact++;
    }
    double rmse = std::sqrt(se / (3.0 * a.pixels.size()));
    printf("shim_demo %dx%d spp=%d rmse_same_paths=%.3e exact=%.4f flipped=%zu\n", a.width, a.height, params.samples,
        rmse, (double)exact / a.pixels.size(), flipped);
    return (rmse < 1e-5 && flipped * 500 < a.pixels.size()) ? 0 : 3;
  } catch (std::exception& e) {
    printf("shim_demo error: %s\n", e.what());
    return 2;
  }
}
