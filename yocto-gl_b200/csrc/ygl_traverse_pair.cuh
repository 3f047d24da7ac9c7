// ygl_traverse_pair.cuh — the PAIR-VISIT variant of the closest-hit stream (build with -DYGL_PAIR_VISIT).
// Measured on B200 (C3, profiles/r02_extend_pair_visit_ab.md): correct (every parity test passes) but SLOWER than the
// single-visit stream in ygl_traverse.cuh - 150 vs 118 ms of extend time per 16 spp of the 1080p frame: with half as
// many node steps per ray the lanes of a warp spread more evenly over the three code paths (node / primitive /
// instance entry) and the majority vote runs ~10 of 32 lanes per instruction instead of ~14. Kept as the committed
// side of that A/B, not compiled by default.
#pragma once

namespace ygl {

// ==========================================================================================================
// Persistent closest-hit stream (intersect_scene_bvh with the default ray range tmin = ray_eps, tmax = flt_max):
// a warp keeps refilling finished lanes from `src` (src.fetch hands out rays, src.commit_finished stores a
// finished lane's hit). Must be called by full, converged warps.
//
// PAIR VISITS. The reference pops a node, tests its box, and pushes both children (yocto_bvh.cpp:485-503). The two
// children of a node are adjacent in memory, so this walk tests BOTH child boxes when it expands a node (one
// 64-byte fetch, no second dependent round trip for the sibling) and stacks the far child only if the ray hits it,
// together with its entry distance t0. When the far child is popped later, `tmax` may have shrunk; the reference's
// test at that moment is t0 <= fl(min(t1, tmax) * 1.00000024f). Rounding is monotone, so that equals
// t0 <= fl(t1 * k) (known true: the child passed with a larger tmax) AND t0 <= fl(tmax * k) (one compare at pop).
// Hence every node the reference would accept is expanded, every node it would reject is dropped, in the same
// order: the sequence of primitive tests and tmax updates - and so every hit bit - is the reference's.
// Rays with an infinite 1/d component (NaN-capable slabs) cannot use the split test; their children are stacked
// untested and take the reference's pop-test-push route (kWorkUntested).
// COUNT: a node counts when the reference would pop it: both children of every expanded node, and tree roots.
// ==========================================================================================================
constexpr int kRefillThreshold = 8;  // refill once this many lanes are idle
constexpr int kPollInterval    = 6;  // scheduling rounds between two polls of an empty ring queue ...
constexpr int kPollMaxInterval = 96;  // ... doubling up to this while it stays empty
constexpr int kStreamThreads   = 128;  // block size of kernels that call trace_stream
constexpr int kSharedStack     = 24;   // stack entries (t0, word) per lane kept in shared memory by k_extend (24 KB per 128 threads)
constexpr int kSuspendMinRounds = 96;  // a warp walks at least this many rounds per launch before suspending: progress guarantee
constexpr int kSuspendBelow    = 8;    // suspend a drained warp's stragglers once this few lanes are busy (0 = never)

// work words of a lane (`cur`) and of stack entries. Node words (word_*) are < 0xC0000000 as unsigned.
constexpr int kWorkDone     = (int)0xC0000000;  // lane idle / bottom of the stack
constexpr int kWorkExit     = (int)0xC0000001;  // stack only: leave the instance (back to the world-space ray)
constexpr int kWorkUntested = (int)0xD0000000;  // | node index (28 bits): fetch the node and test its own box
constexpr int kWorkEnter    = (int)0xE0000000;  // | (count - 1) << 26 | first packet (26 bits): instances of a top-level leaf
YGL_D bool work_is_leaf(int w) { return w >= 0; }
YGL_D bool work_is_node(int w) { return (w & (int)0xC0000000) == (int)0x80000000; }
YGL_D bool work_is_untested(int w) { return (w & (int)0xF0000000) == kWorkUntested; }
YGL_D bool work_is_enter(int w) { return (unsigned)w >= 0xE0000000u; }

// slab test that also returns the entry distance t0 (fast variant: no operand can be NaN)
YGL_D bool slab_test_t0(const f3& o, const f3& dinv, float tmin, float tmax, const float4& n0, const float4& n1, float& t0) {
  float ax = (n0.x - o.x) * dinv.x, bx = (n0.w - o.x) * dinv.x;
  float ay = (n0.y - o.y) * dinv.y, by = (n1.x - o.y) * dinv.y;
  float az = (n0.z - o.z) * dinv.z, bz = (n1.y - o.z) * dinv.z;
  t0       = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tmin);
  float t1 = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}

template <bool COUNT, int THREADS, int SHARED, bool DEEP, class Source, bool TOP = false>  // (DEEP, TOP: not used here)
YGL_D void trace_stream(const DScene& scene, Source& src, trav_counters& cnt) {
  constexpr int kSharedStack = SHARED;  // entries of this instantiation (static shared memory is limited to 48 KB)
  // Traversal stack of (t0, work word) entries: the first kSharedStack entries of every lane live in shared memory,
  // laid out [entry][thread] so a lane always hits its own bank (conflict-free whatever the lanes' depths are);
  // deeper entries spill to per-thread local arrays (rare).
  __shared__ float s_t0[kSharedStack][THREADS];
  __shared__ int   s_wd[kSharedStack][THREADS];
  float     l_t0[2 * kStackSize + 8 - kSharedStack];
  int       l_wd[2 * kStackSize + 8 - kSharedStack];
  const int tix  = threadIdx.x;
  int       sp   = 1;
  auto      push = [&](float t0, int w) {
    if (sp < kSharedStack) s_t0[sp][tix] = t0, s_wd[sp][tix] = w;
    else l_t0[sp - kSharedStack] = t0, l_wd[sp - kSharedStack] = w;
    sp++;
  };
  s_t0[0][tix] = -kFltMax, s_wd[0][tix] = kWorkDone;  // sentinel: popping an empty stack yields "done"
  hit_t res = {-1, -1, {0, 0}, 0, false};

  const float tmin = kRayEps;
  float       tmax = kFltMax;
  ray_setup   world = make_ray(f3{0, 0, 0}, f3{0, 0, 1});
  ray_setup   ray   = world;

  const float4* __restrict__ nodes = scene.top_nodes;
  const float4* packets            = nullptr;
  const int*    prims              = nullptr;
  int  kind = kElemNone, cur_instance = -1, cur_packet = -1;
  bool bottom = false, have = false, more = true;

  // pop until an entry survives the reference's box test with the current tmax (see the header comment);
  // markers, ENTER runs and untested nodes carry t0 = -flt_max and always survive
  auto advance = [&]() {
    while (true) {
      --sp;
      const float t0 = sp < kSharedStack ? s_t0[sp][tix] : l_t0[sp - kSharedStack];
      const int   w  = sp < kSharedStack ? s_wd[sp][tix] : l_wd[sp - kSharedStack];
      if (w == kWorkExit) {  // leaving an instance: back to the world-space ray
        ray        = world;
        nodes      = scene.top_nodes;
        bottom     = false;
        cur_packet = -1;
        continue;
      }
      if (w == kWorkDone) {
        sp = 1;
        return w;
      }
      if (t0 <= tmax * 1.00000024f) return w;
    }
  };
  // the instance-level part of entering packet `first`: transform_ray(inverse(frame, true), ray), yocto_bvh.cpp:601-604
  // Returns the work word of the shape's root (its box is tested here: the packet carries the root node).
  auto load_instance = [&](int first, bool test_root) {
    cur_packet = first;
    const DInstancePacket* pk = scene.top_packets + first;
    float4 a = __ldg(&pk->q[0]), b = __ldg(&pk->q[1]), c = __ldg(&pk->q[2]), e = __ldg(&pk->q[3]);
    float4 p0 = __ldg(&pk->q[4]), p1 = __ldg(&pk->q[5]);
    frame3 inv = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
    ray          = make_ray(transform_point(inv, world.o), transform_vector(inv, world.d));
    cur_instance = __float_as_int(e.y);
    kind         = __float_as_int(e.z);
    nodes   = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.y) << 32) | (unsigned)__float_as_int(p0.x));
    packets = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.w) << 32) | (unsigned)__float_as_int(p0.z));
    prims   = (const int*)(((unsigned long long)(unsigned)__float_as_int(p1.y) << 32) | (unsigned)__float_as_int(p1.x));
    bottom  = true;
    if (!test_root) return kWorkDone;
    float4 r0 = __ldg(&pk->q[6]), r1 = __ldg(&pk->q[7]);
    if (COUNT) cnt.instances++, cnt.bot_nodes++;
    const bool inside = ray.exact ? slab_test_exact(ray.o, ray.dinv, tmin, tmax, r0, r1)
                                  : slab_test_fast(ray.o, ray.dinv, tmin, tmax, r0, r1);
    return inside ? __float_as_int(r1.w) : kWorkDone;
  };
  // ENTER work: the next instance of a top-level leaf's run (instances of a leaf in order, yocto_bvh.cpp:599-610)
  auto enter = [&](int run) {
    const int first = run & 0x03ffffff, left = (run >> 26) & 7;
    if (left > 0) push(-kFltMax, kWorkEnter | ((left - 1) << 26) | (first + 1));
    push(-kFltMax, kWorkExit);
    const int w = load_instance(first, true);
    return w == kWorkDone ? advance() : w;  // root missed: straight to EXIT
  };

  int cur = kWorkDone, leaf_next = 0, leaf_end = 0, rounds = 0, poll_wait = 0, poll_gap = kPollInterval;
  // one unit of node work: expand an accepted internal node (test both children), or test an untested node
  auto visit_node = [&]() {
    if (work_is_untested(cur)) {
      const int n  = cur & 0x0fffffff;
      float4    n0 = __ldg(nodes + 2 * n), n1 = __ldg(nodes + 2 * n + 1);
      if (COUNT) {
        if (bottom) cnt.bot_nodes++;
        else cnt.top_nodes++;
      }
      const bool inside = ray.exact ? slab_test_exact(ray.o, ray.dinv, tmin, tmax, n0, n1)
                                    : slab_test_fast(ray.o, ray.dinv, tmin, tmax, n0, n1);
      cur = inside ? __float_as_int(n1.w) : advance();
      return;
    }
    const int first = word_first_child(cur);
    const int neg   = (ray.sgn >> word_axis(cur)) & 1;  // visit child `neg` first (ray_dsign[axis], yocto_bvh.cpp:592-598)
    if (ray.exact) {
      push(-kFltMax, kWorkUntested | (first + 1 - neg));
      cur = kWorkUntested | (first + neg);
      return;
    }
    const float4* pair = nodes + 2 * (size_t)first;  // node i = float4s 2i, 2i+1: the pair is 64 contiguous bytes
    float4 a0 = __ldg(pair), a1 = __ldg(pair + 1), b0 = __ldg(pair + 2), b1 = __ldg(pair + 3);
    if (COUNT) {
      if (bottom) cnt.bot_nodes += 2;
      else cnt.top_nodes += 2;
    }
    float      ta, tb;
    const bool ha = slab_test_t0(ray.o, ray.dinv, tmin, tmax, a0, a1, ta);
    const bool hb = slab_test_t0(ray.o, ray.dinv, tmin, tmax, b0, b1, tb);
    const int  wa = __float_as_int(a1.w), wb = __float_as_int(b1.w);
    // near = the child the reference visits first, far = the one it stacks
    const bool  h_near = neg ? hb : ha, h_far = neg ? ha : hb;
    const int   w_near = neg ? wb : wa, w_far = neg ? wa : wb;
    const float t_far  = neg ? ta : tb;
    if (h_far) push(t_far, w_far);
    cur = h_near ? w_near : advance();
  };
  // one primitive of the current leaf
  auto test_prim = [&](int idx) {
    if (COUNT) cnt.prims_by_kind[kind]++;
    f2    puv = {0, 0};
    float pd  = 0;
    bool  h;
    if (kind == kElemTriangles) {
      float4 a = __ldg(packets + 3 * idx), b = __ldg(packets + 3 * idx + 1), c = __ldg(packets + 3 * idx + 2);
      h = hit_triangle(ray.o, ray.d, tmin, tmax, f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y}, f3{b.z, b.w, c.x}, puv, pd);
    } else if (kind == kElemQuads) {
      float4 a = __ldg(packets + 4 * idx), b = __ldg(packets + 4 * idx + 1), c = __ldg(packets + 4 * idx + 2),
             e = __ldg(packets + 4 * idx + 3);
      h = hit_quad(ray.o, ray.d, tmin, tmax, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, f3{c.x, c.y, c.z},
          f3{e.x, e.y, e.z}, puv, pd);
    } else if (kind == kElemLines) {
      float4 a = __ldg(packets + 2 * idx), b = __ldg(packets + 2 * idx + 1);
      h = hit_line(ray.o, ray.d, tmin, tmax, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, a.w, b.w, puv, pd);
    } else {
      float4 a = __ldg(packets + idx);
      h = hit_point(ray.o, ray.d, tmin, tmax, f3{a.x, a.y, a.z}, a.w, puv, pd);
    }
    if (h) {
      res  = {cur_instance, __ldg(prims + idx), puv, pd, true};
      tmax = pd;
    }
  };
  // a lane holding a leaf word of the instance tree turns it into an ENTER run
  auto top_leaf_to_run = [&](int w) { return kWorkEnter | ((word_num(w) - 1) << 26) | word_first_prim(w); };
  // park the lane's unfinished walk in this thread's save slot; it resumes in the next launch exactly where it stopped
  auto suspend_lane = [&]() {
    int* sv = src.save_slot();
    sv[0] = cur, sv[1] = sp, sv[2] = leaf_next, sv[3] = leaf_end, sv[4] = cur_packet;
    sv[5] = __float_as_int(tmax);
    sv[6] = res.instance, sv[7] = res.element, sv[8] = __float_as_int(res.uv.x), sv[9] = __float_as_int(res.uv.y);
    sv[10] = __float_as_int(res.distance), sv[11] = res.hit ? 1 : 0;
    for (int k = 1; k < sp; k++) {
      sv[12 + 2 * k]     = __float_as_int(k < kSharedStack ? s_t0[k][tix] : l_t0[k - kSharedStack]);
      sv[12 + 2 * k + 1] = k < kSharedStack ? s_wd[k][tix] : l_wd[k - kSharedStack];
    }
    src.commit_suspended();
    cur = kWorkDone, sp = 1, have = false;
  };
  while (true) {
    const bool     is_leaf    = work_is_leaf(cur);
    const unsigned want_node  = __ballot_sync(kFullWarp, work_is_node(cur) || work_is_untested(cur));
    const unsigned want_prim  = __ballot_sync(kFullWarp, is_leaf && bottom);
    const unsigned want_enter = __ballot_sync(kFullWarp, work_is_enter(cur) || (is_leaf && !bottom));
    const unsigned busy       = want_node | want_prim | want_enter;
    const int      n_idle     = 32 - __popc(busy);

    if (more && n_idle >= src.refill_thr && (!Source::kPolling || !busy || --poll_wait < 0)) {
      // ---- refill: finished lanes hand over their hit and take the next queued ray ----
      const bool idle = cur == kWorkDone;
      src.commit_finished(idle && have, res);
      f3   o, d;
      bool resume = false;
      const bool got = src.fetch(idle, o, d, more, resume);
      if (idle) have = got;
      if (got) {
        world = make_ray(o, d);
        ray   = world;
        tmax  = kFltMax;
        res   = {-1, -1, {0, 0}, 0, false};
        sp = 1, bottom = false, cur_packet = -1;
        nodes = scene.top_nodes;
        cur   = scene.top_num_nodes > 0 ? (kWorkUntested | 0) : kWorkDone;  // the root of the instance tree
        if (resume) {
          // a ray suspended by the previous launch: restore its walk exactly where it stopped
          const int* sv = src.load_slot();
          cur = sv[0], sp = sv[1], leaf_next = sv[2], leaf_end = sv[3];
          const int pkt = sv[4];
          tmax = __int_as_float(sv[5]);
          res  = {sv[6], sv[7], {__int_as_float(sv[8]), __int_as_float(sv[9])}, __int_as_float(sv[10]), sv[11] != 0};
          for (int k = 1; k < sp; k++) {
            if (k < kSharedStack) s_t0[k][tix] = __int_as_float(sv[12 + 2 * k]), s_wd[k][tix] = sv[12 + 2 * k + 1];
            else l_t0[k - kSharedStack] = __int_as_float(sv[12 + 2 * k]), l_wd[k - kSharedStack] = sv[12 + 2 * k + 1];
          }
          if (pkt >= 0) load_instance(pkt, false);
        }
      }
      if (!Source::kPolling) continue;
      // a polling source (ring queue fed by other warps while this one runs) may have had nothing to hand out:
      // keep walking the busy lanes and ask again a few rounds later; a fully idle warp returns to its caller
      if (__any_sync(kFullWarp, got)) {
        poll_wait = 0, poll_gap = kPollInterval;
        continue;
      }
      poll_wait = poll_gap;  // empty ring: exponential back-off keeps idle polls off the L2 hot line
      poll_gap  = min(2 * poll_gap, kPollMaxInterval);
      if (!busy) break;
    }
    if (!busy) break;
    if (Source::kPark && !more && src.suspend_below > 0 && ++rounds >= (src.suspend_below >> 8) &&
        __popc(busy) <= (src.suspend_below & 0xff)) {
      // ---- suspend: the queue is exhausted and this warp is running nearly empty. Instead of draining
      // the stragglers at 1-2 lanes per instruction, save their traversal state; they resume in the next
      // launch (one wavefront iteration later — a per-lane delay only, the walk itself is unchanged). ----
      if (cur != kWorkDone && sp <= kSuspendStack) suspend_lane();
      if (!__any_sync(kFullWarp, cur != kWorkDone)) break;
      // lanes with a deeper stack than the save area keep walking
    }
    if (Source::kLone && !more && src.lone_below > 0 && __popc(busy) <= src.lone_below) {
      // ---- tail: the queue is exhausted and only a few rays of this warp are left. Their remaining walk is a
      // chain of dependent steps; without the per-round votes each step is ~20 % shorter. ----
      // lone_steps > 0 caps the tail: a ray that still walks after that many more steps is parked for the next
      // launch (only the extreme stragglers get there; a parked ray always advances lone_steps per launch)
      const int limit = src.lone_steps > 0 ? src.lone_steps : 0x7fffffff;
      int       steps = 0;
      while (cur != kWorkDone) {
        if (++steps > limit && sp <= kSuspendStack) {
          suspend_lane();
          break;
        }
        if (work_is_node(cur) || work_is_untested(cur)) {
          visit_node();
        } else if (work_is_leaf(cur) && bottom) {
          for (int idx = word_first_prim(cur), end = idx + word_num(cur); idx < end; idx++) test_prim(idx);
          cur = advance();
        } else {
          cur = work_is_leaf(cur) ? (word_num(cur) > 0 ? enter(top_leaf_to_run(cur)) : advance()) : enter(cur);
        }
      }
      break;
    }
    const int n_node = __popc(want_node), n_prim = __popc(want_prim), n_enter = __popc(want_enter);

    if (n_node >= n_prim && n_node >= n_enter) {
      // several units of node work per scheduling round: cuts the vote overhead on the most frequent path
#pragma unroll 1
      for (int rep = 0; rep < src.node_reps; rep++)
        if (work_is_node(cur) || work_is_untested(cur)) visit_node();
    } else if (n_prim >= n_enter) {
      // all primitives of the lane's leaf (<= 4, bvh_max_prims) in one go, warp-uniform trip count
      const bool in_leaf = is_leaf && bottom;
      if (in_leaf) leaf_next = word_first_prim(cur), leaf_end = leaf_next + word_num(cur);
      for (int k = 0; k < 4; k++) {
        const bool test = in_leaf && leaf_next < leaf_end;
        if (!__any_sync(kFullWarp, test)) break;
        if (test) test_prim(leaf_next++);
      }
      if (in_leaf) cur = advance();
    } else {
      if (is_leaf && !bottom) cur = word_num(cur) > 0 ? enter(top_leaf_to_run(cur)) : advance();
      else if (work_is_enter(cur)) cur = enter(cur);
    }
  }
  // lanes still holding an uncommitted result (more == false path)
  src.commit_finished(have && cur == kWorkDone, res);
}

}  // namespace ygl
