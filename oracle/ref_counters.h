// oracle/ref_counters.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Traversal counters of the instrumented oracle ("means per traced ray as counted by the
// instrumented oracle"). oracle/make_counted_bvh.py writes a copy of the reference's yocto_bvh.cpp into
// oracle/_ref/gen/ (a build product, never committed) with one REF_COUNT(...) statement added at each counter
// site (yocto_bvh.cpp:466,487,506-545,560,581,600,621); that copy is compiled INSTEAD of yocto_bvh.cpp into
// oracle/_ref/libyocto_ref_count.so. The arithmetic of the walk is untouched.
//
// Counters are thread-local (the reference renders on a thread pool) and are folded into process-wide totals when
// a worker thread exits or when ref_counters_read() is called on the calling thread.
#pragma once
#include <atomic>
#include <cstdint>

namespace refcount {

enum : int {
  kTopNodes = 0,   // instance-tree nodes popped            (yocto_bvh.cpp:581)
  kBotNodes,       // shape-tree nodes popped               (yocto_bvh.cpp:487)
  kInstVisits,     // intersect_shape_bvh invocations       (yocto_bvh.cpp:466)
  kPoints,         // primitives tested, by type            (yocto_bvh.cpp:506-545)
  kLines,
  kTriangles,
  kQuads,
  kRays,           // intersect_scene_bvh / intersect_instance_bvh calls (yocto_bvh.cpp:560 / :621)
  kPerMode
};
enum : int { kModeScene = 0, kModeInstance = 1, kNumModes = 2 };

extern std::atomic<uint64_t> g_totals[kNumModes * kPerMode];

struct tls_counters {
  uint64_t v[kNumModes * kPerMode] = {};
  int      mode                    = kModeScene;
  void     flush() {
    for (int k = 0; k < kNumModes * kPerMode; k++) {
      if (v[k]) g_totals[k].fetch_add(v[k], std::memory_order_relaxed);
      v[k] = 0;
    }
  }
  ~tls_counters() { flush(); }
};
extern thread_local tls_counters t_counters;

}  // namespace refcount

#define REF_COUNT(what) (++refcount::t_counters.v[refcount::t_counters.mode * refcount::kPerMode + refcount::what])
#define REF_MODE(m) (refcount::t_counters.mode = refcount::m)
