// oracle/shim_demo.cpp — TEST INFRASTRUCTURE. A reference-side C++ program using the drop-in shim
// (yocto-gl_b200/host/yocto_b200trace.h): renders make_cornellbox() with yocto::trace_image (CPU
// reference) and yocto::b200::trace_image (libygl_b200.so) and prints the comparison. Built by
// oracle/Makefile into oracle/_ref/shim_demo where the reference headers exist; run on the GPU box
// by tests/test_gpu_parity.py::test_reference_side_shim_runs.
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <thread>

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
    // the low-level sequence (yocto_trace.h:160-190) next to the reference's own, incl. the reference's trees adopted
    // verbatim, the denoise guides, a per-pixel trace_sample and the progressive calls
    {
      using namespace yocto;
      auto lp    = params;
      lp.batch   = 2;
      lp.samples = 4;
      auto rbvh = make_trace_bvh(scene, lp);
      auto rlights = make_trace_lights(scene, lp);
      auto rstate = make_trace_state(scene, lp);
      while (rstate.samples < lp.samples) trace_samples(rstate, scene, rbvh, rlights, lp);
      auto ctx    = b200::b200_context{0};
      auto dscene = b200::make_b200_scene(ctx, scene);
      auto dbvh   = b200::make_trace_bvh(scene, rbvh);  // the reference's own bvh_tree arrays
      auto dlights = b200::make_trace_lights(scene, lp);
      auto dstate = b200::make_trace_state(ctx, scene, lp);
      while (true) {
        int before = 0;
        ygl_state_size(dstate->handle, nullptr, nullptr, &before);
        if (before >= lp.samples) break;
        b200::trace_samples(ctx, *dstate, *dscene, *dbvh, *dlights, lp);
      }
      auto same = [](const image_data& x, const image_data& y) {
        return x.pixels.size() == y.pixels.size() && memcmp(x.pixels.data(), y.pixels.data(), x.pixels.size() * sizeof(vec4f)) == 0;
      };
      bool ok = same(get_image(rstate), b200::get_image(*dstate)) &&
                same(get_albedo_image(rstate), b200::get_albedo_image(*dstate)) &&
                same(get_normal_image(rstate), b200::get_normal_image(*dstate));
      // one more sample of one pixel on both sides
      trace_sample(rstate, scene, rbvh, rlights, 3, 5, lp.samples, lp);
      b200::trace_sample(ctx, *dstate, *dscene, *dbvh, *dlights, 3, 5, lp.samples, lp);
      ok = ok && same(get_image(rstate), b200::get_image(*dstate));
      // progressive: a started batch completes and reports done
      b200::reset_trace_state(*dstate, lp);
      b200::trace_start(ctx, *dstate, *dscene, *dbvh, *dlights, lp);
      for (int spin = 0; !b200::trace_done(ctx) && spin < 200000; spin++) std::this_thread::sleep_for(std::chrono::microseconds(50));
      ok = ok && b200::trace_done(ctx);
      b200::trace_cancel(ctx);  // joins the finished worker
      printf("shim_demo low-level sequence (adopted bvh, guides, trace_sample, trace_start): %s\n", ok ? "bit-exact" : "MISMATCH");
      if (!ok) return 4;
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
