// ygl_traverse.cuh — two-level BVH closest-hit / any-hit traversal and ray-primitive tests.
// Behavioural contract: libs/yocto/yocto_bvh.cpp:460-628 (intersect_shape_bvh,
// intersect_scene_bvh, intersect_instance_bvh) and libs/yocto/yocto_geometry.h:697-864
// (intersect_point/line/triangle/quad/bbox). Node visit order, tie-breaking ("last tested
// t <= tmax wins") and every rounding step are the reference's, so hit ids are bit-exact.
//
// Shape of the kernel-side walk (DESIGN.md "traversal"): the reference nests the shape walk inside
// the instance walk; on a GPU that leaves the lanes of a warp scattered over two loops plus the
// primitive tests (measured: 2.4 of 32 lanes active per issued instruction). Here ONE loop serves
// both levels: a lane's per-ray sequence of node visits, primitive tests and tmax updates is
// exactly the reference's, but instance entry/exit are stack events (ENTER / EXIT markers) and
// the walk is organised "while-while": all lanes pop and slab-test nodes together until each has
// found a leaf (or a marker), then leaves are processed together.
#pragma once

#include "ygl_scene.cuh"

namespace ygl {

constexpr int kShallowStack = 72;  // stack capacity of the !DEEP kernel variants (both levels + markers)
constexpr int kStackSize = 128;  // per level, the reference's own stack size (yocto_bvh.cpp:469); the host build rejects
                                 // deeper trees (the reference's array is unchecked there)

struct hit_t {
  int   instance, element;
  f2    uv;
  float distance;
  bool  hit;
};

struct trav_counters {  // per-thread traversal statistics (SURVEY.md §8d algorithmic bytes)
  unsigned top_nodes, bot_nodes, instances, prims;
  unsigned prims_by_kind[5];  // indexed by kElem* (trace_stream only)
};

// intersect_bbox(ray, ray_dinv, bbox), yocto_geometry.h:854-864, with yocto's min/max
// ((a<b)?a:b / (a>b)?a:b), used when a slab product can be NaN (a zero or denormal direction
// component makes dinv infinite).
YGL_D bool slab_test_exact(const f3& o, const f3& dinv, float tmin, float tmax, const float4& n0, const float4& n1) {
  auto bmin   = f3{n0.x, n0.y, n0.z};
  auto bmax   = f3{n0.w, n1.x, n1.y};
  auto it_min = (bmin - o) * dinv;
  auto it_max = (bmax - o) * dinv;
  auto lo     = vmin(it_min, it_max);
  auto hi     = vmax(it_min, it_max);
  auto t0     = ymax(max3(lo), tmin);
  auto t1     = ymin(min3(hi), tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}
// Same test with hardware min/max (FMNMX). Valid when no operand is NaN: then IEEE min/max and
// yocto's differ only in the sign of a zero result, which cannot change `t0 <= t1`.
YGL_D bool slab_test_fast(const f3& o, const f3& dinv, float tmin, float tmax, const float4& n0, const float4& n1) {
  float ax = (n0.x - o.x) * dinv.x, bx = (n0.w - o.x) * dinv.x;
  float ay = (n0.y - o.y) * dinv.y, by = (n1.x - o.y) * dinv.y;
  float az = (n0.z - o.z) * dinv.z, bz = (n1.y - o.z) * dinv.z;
  float t0 = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tmin);
  float t1 = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}

// intersect_triangle, yocto_geometry.h:794-825 (edges precomputed: e1 = p1 - p0, e2 = p2 - p0)
YGL_D bool hit_triangle(const f3& o, const f3& d, float tmin, float tmax, const f3& p0, const f3& edge1,
    const f3& edge2, f2& uv, float& dist) {
  auto pvec = cross(d, edge2);
  auto det  = dot(edge1, pvec);
  if (det == 0) return false;
  auto inv_det = 1.0f / det;
  auto tvec    = o - p0;
  auto u       = dot(tvec, pvec) * inv_det;
  if (u < 0 || u > 1) return false;
  auto qvec = cross(tvec, edge1);
  auto v    = dot(d, qvec) * inv_det;
  if (v < 0 || u + v > 1) return false;
  auto t = dot(edge2, qvec) * inv_det;
  if (t < tmin || t > tmax) return false;
  uv   = {u, v};
  dist = t;
  return true;
}

// intersect_quad, yocto_geometry.h:828-835
YGL_D bool hit_quad(const f3& o, const f3& d, float tmin, float tmax, const f3& p0, const f3& p1, const f3& p2,
    const f3& p3, f2& uv, float& dist) {
  if (p2 == p3) return hit_triangle(o, d, tmin, tmax, p0, p1 - p0, p3 - p0, uv, dist);
  f2    uv1 = {0, 0}, uv2 = {0, 0};
  float d1 = kFltMax, d2 = kFltMax;  // prim_intersection{}.distance == flt_max
  bool  h1 = hit_triangle(o, d, tmin, tmax, p0, p1 - p0, p3 - p0, uv1, d1);
  bool  h2 = hit_triangle(o, d, tmin, tmax, p2, p3 - p2, p1 - p2, uv2, d2);
  if (h2) uv2 = f2{1 - uv2.x, 1 - uv2.y};
  if (d1 < d2) {
    uv   = uv1;
    dist = d1;
    return h1;
  }
  uv   = uv2;
  dist = d2;
  return h2;
}

// intersect_line, yocto_geometry.h:716-757
YGL_D bool hit_line(const f3& o, const f3& dir, float tmin, float tmax, const f3& p0, const f3& p1, float r0,
    float r1, f2& uv, float& dist) {
  auto u   = dir;
  auto v   = p1 - p0;
  auto w   = o - p0;
  auto a   = dot(u, u);
  auto b   = dot(u, v);
  auto c   = dot(v, v);
  auto d   = dot(u, w);
  auto e   = dot(v, w);
  auto det = a * c - b * b;
  if (det == 0) return false;
  auto t = (b * e - c * d) / det;
  auto s = (a * e - b * d) / det;
  if (t < tmin || t > tmax) return false;
  s        = yclamp(s, (float)0, (float)1);
  auto pr  = o + dir * t;
  auto pl  = p0 + (p1 - p0) * s;
  auto prl = pr - pl;
  auto d2  = dot(prl, prl);
  auto r   = r0 * (1 - s) + r1 * s;
  if (d2 > r * r) return false;
  uv   = {s, ysqrt(d2) / r};
  dist = t;
  return true;
}

// intersect_point, yocto_geometry.h:697-713
YGL_D bool hit_point(const f3& o, const f3& d, float tmin, float tmax, const f3& p, float r, f2& uv, float& dist) {
  auto w = p - o;
  auto t = dot(w, d) / dot(d, d);
  if (t < tmin || t > tmax) return false;
  auto rp  = o + d * t;
  auto prp = p - rp;
  if (dot(prp, prp) > r * r) return false;
  uv   = {0, 0};
  dist = t;
  return true;
}

// node words (second float4 of a node, .w; written by pack_nodes in ygl_build.cpp):
//   internal: 0x80000000 | axis << 28 | first child index (28 bits)     leaf: num << 26 | first primitive (26 bits)
YGL_D bool word_internal(int w) { return w < 0; }
YGL_D int  word_first_child(int w) { return w & 0x0fffffff; }
YGL_D int  word_axis(int w) { return (w >> 28) & 3; }
YGL_D int  word_first_prim(int w) { return w & 0x03ffffff; }
YGL_D int  word_num(int w) { return (w >> 26) & 7; }

// trace_ray's stack markers (node indices are >= 0; an ENTER entry is ~(index into the leaf-ordered packets))
constexpr int kMarkDone    = (int)0x80000000;
constexpr int kMarkExit    = (int)0x80000001;
constexpr int kMarkLeafEnd = (int)0x80000002;
constexpr int kMarkLeaf    = (int)0x80000003;  // lane holds a leaf (leaf_start, leaf_num) to process
constexpr unsigned kFullWarp = 0xffffffffu;

struct ray_setup {
  f3       o, d, dinv;
  unsigned sgn;    // bit a set <=> dinv[a] < 0 (ray_dsign, yocto_bvh.cpp:476-478)
  bool     exact;  // some dinv component is infinite: NaN-capable slabs, use yocto min/max
};
YGL_D ray_setup make_ray(const f3& o, const f3& d) {
  ray_setup r;
  r.o = o, r.d = d;
  r.dinv  = f3{1 / d.x, 1 / d.y, 1 / d.z};
  r.sgn   = ((r.dinv.x < 0) ? 1u : 0u) | ((r.dinv.y < 0) ? 2u : 0u) | ((r.dinv.z < 0) ? 4u : 0u);
  // finite non-zero dinv and finite origin/bounds => no slab product can be NaN
  r.exact = !(yfinite(r.dinv.x) && yfinite(r.dinv.y) && yfinite(r.dinv.z) && r.dinv.x != 0 && r.dinv.y != 0 &&
              r.dinv.z != 0 && yfinite(o.x) && yfinite(o.y) && yfinite(o.z));
  return r;
}

// intersect_scene_bvh (start_instance < 0) / intersect_instance_bvh (start_instance >= 0).
// ANY = find_any. COUNT = gather traversal statistics.
//
// WARP-COOPERATIVE: all 32 lanes of a warp must call this together (converged); lanes without a
// ray pass active = false. The votes (__any_sync / __all_sync) are the reconvergence points: sm_100
// schedules diverged lanes independently and would otherwise never bring them back in step.
// DEEP: stack for trees up to the reference's 128 levels per level of the hierarchy; !DEEP: kShallowStack entries, for
// scenes whose trees fit (DScene::stack_mode, chosen by the host) - a 1 KB local array per thread otherwise.
template <bool ANY, bool COUNT, bool DEEP = true>
YGL_D hit_t trace_ray(const DScene& scene, bool active, const f3& ray_o, const f3& ray_d, float tmin, float tmax,
    int start_instance, trav_counters& cnt) {
  int   stack[DEEP ? 2 * kStackSize + 8 : kShallowStack];
  int   sp  = 0;
  hit_t res = {-1, -1, {0, 0}, 0, false};

  const bool range_nan = !(tmin == tmin && tmax == tmax);
  ray_setup  world     = make_ray(ray_o, ray_d);
  world.exact |= range_nan;
  ray_setup ray = world;

  const float4* __restrict__ nodes = scene.top_nodes;
  const float4* packets            = nullptr;
  const int*    prims              = nullptr;
  int  kind = kElemNone, cur_instance = -1, inst_sp = 0;
  bool bottom = false, shape_hit = false;

  auto pop = [&]() { return sp > 0 ? stack[--sp] : kMarkDone; };
  // one instance visit: transform_ray(inverse(frame, true), ray), yocto_bvh.cpp:601-604 / :621-624
  auto enter = [&](const DInstancePacket* pk) {
    float4 a = __ldg(&pk->q[0]), b = __ldg(&pk->q[1]), c = __ldg(&pk->q[2]), e = __ldg(&pk->q[3]);
    float4 p0 = __ldg(&pk->q[4]), p1 = __ldg(&pk->q[5]);
    frame3 inv = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
    ray        = make_ray(transform_point(inv, world.o), transform_vector(inv, world.d));
    ray.exact |= range_nan;
    if (COUNT) cnt.instances++;
    cur_instance = __float_as_int(e.y);
    kind         = __float_as_int(e.z);
    nodes   = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.y) << 32) | (unsigned)__float_as_int(p0.x));
    packets = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.w) << 32) | (unsigned)__float_as_int(p0.z));
    prims   = (const int*)(((unsigned long long)(unsigned)__float_as_int(p1.y) << 32) | (unsigned)__float_as_int(p1.x));
    bottom       = true;
    shape_hit    = false;
    stack[sp++]  = kMarkExit;
    inst_sp      = sp;
    return 0;  // shape root (make_bvh always emits one)
  };

  // pop + resolve the cheap markers (instance exit, end of an instance leaf) until the lane holds a node,
  // an ENTER entry or is done
  auto advance = [&]() {
    while (true) {
      int v = pop();
      if (v == kMarkExit) {
        ray    = world;
        nodes  = scene.top_nodes;
        bottom = false;
        continue;
      }
      if (ANY && v == kMarkLeafEnd) {
        if (res.hit) {  // yocto_bvh.cpp:613
          sp = 0;
          return kMarkDone;
        }
        continue;
      }
      return v;
    }
  };

  int cur = kMarkDone;  // >= 0: node to visit; kMarkLeaf: primitive leaf in progress; < 0 else: ENTER ~idx / done
  if (active) cur = start_instance >= 0 ? enter(scene.inst_packets + start_instance) : (scene.top_num_nodes > 0 ? 0 : kMarkDone);

  int leaf_next = 0, leaf_end = 0;
  // Warp-level greedy scheduling of the three code paths: each round the path wanted by most lanes runs
  // for those lanes (node visit / one primitive test / instance entry); the others keep their state.
  while (true) {
    const unsigned want_node  = __ballot_sync(kFullWarp, cur >= 0);
    const unsigned want_prim  = __ballot_sync(kFullWarp, cur == kMarkLeaf);
    const unsigned want_enter = __ballot_sync(kFullWarp, cur < 0 && cur > kMarkLeaf);
    if (!(want_node | want_prim | want_enter)) break;
    const int n_node = __popc(want_node), n_prim = __popc(want_prim), n_enter = __popc(want_enter);

    if (n_node >= n_prim && n_node >= n_enter) {
      // ---- node visit: pop + slab test, yocto_bvh.cpp:485-503 / :579-598 ----
      if (cur >= 0) {
        float4 n0 = __ldg(nodes + 2 * cur), n1 = __ldg(nodes + 2 * cur + 1);
        if (COUNT) {
          if (bottom) cnt.bot_nodes++;
          else cnt.top_nodes++;
        }
        bool inside = ray.exact ? slab_test_exact(ray.o, ray.dinv, tmin, tmax, n0, n1)
                                : slab_test_fast(ray.o, ray.dinv, tmin, tmax, n0, n1);
        const int word = __float_as_int(n1.w);
        if (!inside) {
          cur = advance();
        } else if (word_internal(word)) {
          // internal: the reference pushes both children ordered by ray_dsign[axis] and pops the last
          // pushed; visiting that one directly and stacking the other is the same sequence.
          const int start = word_first_child(word);
          int       neg   = (ray.sgn >> word_axis(word)) & 1;
          stack[sp++]     = start + 1 - neg;
          cur             = start + neg;
        } else if (bottom) {
          leaf_next = word_first_prim(word);
          leaf_end  = leaf_next + word_num(word);
          cur       = leaf_next < leaf_end ? kMarkLeaf : advance();
        } else {
          // leaf of the instance tree: instances are visited in order, each seeing the tmax left by the
          // previous one (yocto_bvh.cpp:599-610) -> stack them in reverse
          const int start = word_first_prim(word);
          if (ANY) stack[sp++] = kMarkLeafEnd;
          for (int idx = start + word_num(word) - 1; idx >= start; idx--) stack[sp++] = ~idx;
          cur = advance();
        }
      }
    } else if (n_prim >= n_enter) {
      // ---- one primitive test per round, yocto_bvh.cpp:505-545 ----
      if (cur == kMarkLeaf) {
        const int idx = leaf_next++;
        if (COUNT) cnt.prims++;
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
          res       = {cur_instance, __ldg(prims + idx), puv, pd, true};
          tmax      = pd;
          shape_hit = true;
        }
        if (leaf_next == leaf_end) {
          if (ANY && shape_hit) sp = inst_sp;  // intersect_shape_bvh returns at once: unwind to EXIT
          cur = advance();
        }
      }
    } else {
      // ---- instance entry ----
      if (cur < 0 && cur > kMarkLeaf) cur = enter(scene.top_packets + ~cur);
    }
  }
  return res;
}

#ifdef YGL_PAIR_VISIT
}  // namespace ygl
#include "ygl_traverse_pair.cuh"
namespace ygl {
#else
// Persistent closest-hit stream (intersect_scene_bvh with the default ray range tmin = ray_eps,
// tmax = flt_max): the same per-ray walk as trace_ray, but a warp keeps refilling finished lanes from
// `src` (src.fetch hands out rays, src.commit stores a finished lane's hit). Must be called by full,
// converged warps.
constexpr int kRefillThreshold = 8;  // refill once this many lanes are idle
constexpr int kPollInterval    = 6;  // scheduling rounds between two polls of an empty ring queue ...
constexpr int kPollMaxInterval = 96;  // ... doubling up to this while it stays empty
constexpr int kStreamThreads   = 128;  // block size of kernels that call trace_stream
constexpr int kStackShared = 0, kStackShallow = 1, kStackDeep = 2;  // trace_stream<STACK>
constexpr int kSharedStack     = 28;   // stack entries per lane kept in shared memory by k_extend (14 KB per block)
constexpr int kSuspendEntries  = kSuspendWords - 12;  // stack entries a parked ray can carry (12 header words)
constexpr int kSuspendMinRounds = 96;  // a warp walks at least this many rounds per launch before suspending: progress guarantee
constexpr int kSuspendBelow    = 8;    // suspend a drained warp's stragglers once this few lanes are busy (0 = never)

// STACK: kStackDeep = capacity for trees up to the reference's 128 levels per level of the hierarchy (a 1 KB local array
// per thread: its mere presence in the frame costs the kernel ~8 % through L1 pressure); kStackShallow = kShallowStack
// entries, the ones beyond the shared part in a small local array; kStackShared = the shared part only (SHARED entries,
// no local array, no bounds branches on push / pop). The host picks by the depth of the bound trees (DScene::stack_need).
// TOP: the instance-level tree is staged once per block into (dynamic) shared memory with one bulk asynchronous copy
// (cp.async.bulk + mbarrier, the TMA engine's 1-D path) and its nodes are then read with LDS instead of through L1.
// BASELINE.json's north_star names this staging; measured on B200 it is an A/B, not the default.
template <bool COUNT, int THREADS, int SHARED, int STACK, class Source, bool TOP = false>
YGL_D void trace_stream(const DScene& scene, Source& src, trav_counters& cnt) {
  constexpr bool kSpill = STACK != kStackShared;  // entries beyond the shared part exist (a per-thread local array)
  constexpr int kSharedStack = SHARED;
  extern __shared__ __align__(16) float4 s_top[];  // TOP only: top_num_nodes x 2 float4 (dynamic shared memory)
  if (TOP) {
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar   = (unsigned)__cvta_generic_to_shared(&s_bar);
    const unsigned dst   = (unsigned)__cvta_generic_to_shared(s_top);
    const unsigned bytes = (unsigned)scene.top_num_nodes * 32u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(scene.top_nodes), "r"(bytes), "r"(bar)
                   : "memory");
    }
    unsigned ready = 0;
    while (!ready)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ready)
          : "r"(bar), "r"(0u)
          : "memory");
  }
  // Traversal stack: the first kSharedStack entries of every lane live in shared memory, laid out
  // [entry][thread] so a lane always hits its own bank (conflict-free whatever the lanes' depths are);
  // deeper entries spill to a per-thread local array (rare: kSharedStack covers a 1000-instance tree
  // plus a 64K-primitive shape tree).
  __shared__ int s_stack[kSharedStack][THREADS];
#ifdef YGL_COOP_PRIMS
  __shared__ unsigned char s_tasks[THREADS / 32][128];  // per warp: (owner lane | primitive of its leaf << 5) per numbered test
#endif
  int            l_stack[kSpill ? (STACK == kStackDeep ? 2 * kStackSize + 8 : kShallowStack) - kSharedStack : 1];
  // The thread's column of the shared stack as a 32-bit shared-space address, computed once: through the generic
  // `s_stack[sp][tix]` form every push and pop recomputed the shared window base (S2UR CgaCtaId / ULEA / UMOV: 11 % of
  // the kernel's stall samples in the round-2 profile, profiles/r02_c3_kernels_full.txt).
  const unsigned s_col = (unsigned)__cvta_generic_to_shared(&s_stack[0][threadIdx.x]);
  auto           s_put = [&](int entry, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(s_col + (unsigned)entry * (unsigned)(THREADS * sizeof(int))), "r"(v) : "memory");
  };
  auto s_get = [&](int entry) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(s_col + (unsigned)entry * (unsigned)(THREADS * sizeof(int))) : "memory");
    return v;
  };
  const int      tix  = threadIdx.x;
  int            sp   = 1;
  auto           push = [&](int v) {
    if (!kSpill || sp < kSharedStack) s_put(sp, v);
    else l_stack[sp - kSharedStack] = v;
    sp++;
  };
  auto peek = [&]() { return !kSpill || sp <= kSharedStack ? s_get(sp - 1) : l_stack[sp - 1 - kSharedStack]; };
  s_put(0, kMarkDone);  // sentinel: popping an empty stack yields "done"
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

  // ENTER entries carry a run of instances of one top-level leaf: ~(first | (count - 1) << 28)
  auto enter = [&](int entry) {
    const int run = ~entry, first = run & 0x0fffffff, left = run >> 28;
    if (left > 0) push(~((first + 1) | ((left - 1) << 28)));  // next instance of the leaf, same order
    cur_packet = first;
    const DInstancePacket* pk = scene.top_packets + first;
    float4 a = __ldg(&pk->q[0]), b = __ldg(&pk->q[1]), c = __ldg(&pk->q[2]), e = __ldg(&pk->q[3]);
    float4 p0 = __ldg(&pk->q[4]), p1 = __ldg(&pk->q[5]);
    frame3 inv = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
    ray          = make_ray(transform_point(inv, world.o), transform_vector(inv, world.d));
    if (COUNT) cnt.instances++;
    cur_instance = __float_as_int(e.y);
    kind         = __float_as_int(e.z);
    nodes   = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.y) << 32) | (unsigned)__float_as_int(p0.x));
    packets = (const float4*)(((unsigned long long)(unsigned)__float_as_int(p0.w) << 32) | (unsigned)__float_as_int(p0.z));
    prims   = (const int*)(((unsigned long long)(unsigned)__float_as_int(p1.y) << 32) | (unsigned)__float_as_int(p1.x));
    bottom      = true;
    push(kMarkExit);
    return 0;
  };
  // the part of `enter` that only depends on the packet (used to resume a suspended ray)
  auto reload_instance = [&](int first) {
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
    bottom       = true;
    cur_packet   = first;
  };
  // The top entry comes off the stack; the sentinel at entry 0 stays (an empty stack keeps yielding "done"). An EXIT
  // marker is handed back as it is: the lane leaves its instance at the start of its next node round (leave_instance).
  auto pop = [&]() {
    const int v = peek();
    sp -= v == kMarkDone ? 0 : 1;
    return v;
  };
  // leaving an instance: back to the world-space ray. Only what the instance-level walk reads is restored (its slabs
  // need origin, 1/direction and the signs; the direction itself is set again by the next `enter`).
  auto leave_instance = [&]() {
    ray.o = world.o, ray.dinv = world.dinv, ray.sgn = world.sgn, ray.exact = world.exact;
    nodes      = scene.top_nodes;
    bottom     = false;
    cur_packet = -1;
  };

  int cur = kMarkDone, leaf_next = 0, leaf_end = 0, rounds = 0, poll_wait = 0, poll_gap = kPollInterval;
  // one node visit of the lane's walk (cur >= 0, or the EXIT marker: leave the instance, then visit what follows).
  // After the slab test the three outcomes (inner node hit: stack the far child, walk into the near one; leaf hit: hand
  // the leaf to the primitive / instance path; miss or empty leaf: pop) are folded into selects around one speculative
  // read of the stack top, so the lanes of a warp stay together whatever each of them found.
  auto visit_node = [&]() {
    if (cur == kMarkExit) {  // never two EXITs in a row
      leave_instance();
      cur = pop();
    }
    if (cur < 0) return;
    float4 n0, n1;
    if (TOP && !bottom) n0 = s_top[2 * cur], n1 = s_top[2 * cur + 1];
    else n0 = __ldg(nodes + 2 * cur), n1 = __ldg(nodes + 2 * cur + 1);
    if (COUNT) {
      if (bottom) cnt.bot_nodes++;
      else cnt.top_nodes++;
    }
    const int  top_entry = peek();
    const bool inside    = ray.exact ? slab_test_exact(ray.o, ray.dinv, tmin, tmax, n0, n1)
                                     : slab_test_fast(ray.o, ray.dinv, tmin, tmax, n0, n1);
    const int  word  = __float_as_int(n1.w);
    const bool inner = inside && word_internal(word);
    const int  num   = inside ? word_num(word) : 0;  // (read as a leaf word; unused for inner nodes)
    const int  first = word_first_prim(word);
    const int  start = word_first_child(word);
    const int  neg   = (ray.sgn >> word_axis(word)) & 1;
    if (inner) push(start + 1 - neg);
    const bool leaf = !inner && num > 0;
    if (leaf && bottom) leaf_next = first, leaf_end = first + num;
    // leaf of the instance tree (<= 4 instances, visited in order): one ENTER entry for the run
    const int leaf_cur = bottom ? kMarkLeaf : ~(first | (int)((unsigned)(num - 1) << 28));
    if (!inner && !leaf) sp -= top_entry == kMarkDone ? 0 : 1;
    cur = inner ? start + neg : leaf ? leaf_cur : top_entry;
  };
  // one primitive test: element `idx` of a shape of type `kd` whose leaf packets start at `pk`, against the ray (o, d)
  // with the range [tmin, tmx]
  auto test_element = [&](const f3& o, const f3& d, float tmx, const float4* pk, int kd, int idx, f2& puv, float& pd) {
    if (kd == kElemTriangles) {
      float4 a = __ldg(pk + 3 * idx), b = __ldg(pk + 3 * idx + 1), c = __ldg(pk + 3 * idx + 2);
      return hit_triangle(o, d, tmin, tmx, f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y}, f3{b.z, b.w, c.x}, puv, pd);
    } else if (kd == kElemQuads) {
      float4 a = __ldg(pk + 4 * idx), b = __ldg(pk + 4 * idx + 1), c = __ldg(pk + 4 * idx + 2), e = __ldg(pk + 4 * idx + 3);
      return hit_quad(o, d, tmin, tmx, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, f3{c.x, c.y, c.z}, f3{e.x, e.y, e.z}, puv, pd);
    } else if (kd == kElemLines) {
      float4 a = __ldg(pk + 2 * idx), b = __ldg(pk + 2 * idx + 1);
      return hit_line(o, d, tmin, tmx, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, a.w, b.w, puv, pd);
    }
    float4 a = __ldg(pk + idx);
    return hit_point(o, d, tmin, tmx, f3{a.x, a.y, a.z}, a.w, puv, pd);
  };
  // one primitive of the current leaf
  auto test_prim = [&](int idx) {
    if (COUNT) cnt.prims_by_kind[kind]++;
    f2    puv = {0, 0};
    float pd  = 0;
    if (test_element(ray.o, ray.d, tmax, packets, kind, idx, puv, pd)) {
      res  = {cur_instance, __ldg(prims + idx), puv, pd, true};
      tmax = pd;
    }
  };
  // park the lane's unfinished walk in its save area; it resumes in the next launch exactly where it stopped
  auto suspend_lane = [&]() {
    int* sv = src.save_slot();
    sv[0] = cur, sv[1] = sp, sv[2] = leaf_next, sv[3] = leaf_end, sv[4] = cur_packet;
    sv[5] = __float_as_int(tmax);
    sv[6] = res.instance, sv[7] = res.element, sv[8] = __float_as_int(res.uv.x), sv[9] = __float_as_int(res.uv.y);
    sv[10] = __float_as_int(res.distance), sv[11] = res.hit ? 1 : 0;
    for (int k = 1; k < sp; k++) sv[12 + k] = k < kSharedStack ? s_get(k) : l_stack[k - kSharedStack];
    src.commit_suspended();
    cur = kMarkDone, sp = 1, have = false;
  };
  while (true) {
    const unsigned want_node  = __ballot_sync(kFullWarp, cur >= 0 || cur == kMarkExit);
    const unsigned want_prim  = __ballot_sync(kFullWarp, cur == kMarkLeaf);
    const unsigned want_enter = __ballot_sync(kFullWarp, cur < 0 && cur > kMarkLeaf);
    const unsigned busy       = want_node | want_prim | want_enter;
    const int      n_idle     = 32 - __popc(busy);

    if (more && n_idle >= src.refill_thr && (!Source::kPolling || !busy || --poll_wait < 0)) {
      // ---- refill: finished lanes hand over their hit and take the next queued ray ----
      const bool idle = cur == kMarkDone;
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
        cur   = scene.top_num_nodes > 0 ? 0 : kMarkDone;
        if (resume) {
          // a ray suspended by the previous launch: restore its walk exactly where it stopped
          const int* sv = src.load_slot();
          cur = sv[0], sp = sv[1], leaf_next = sv[2], leaf_end = sv[3];
          const int pkt = sv[4];
          tmax = __int_as_float(sv[5]);
          res  = {sv[6], sv[7], {__int_as_float(sv[8]), __int_as_float(sv[9])}, __int_as_float(sv[10]), sv[11] != 0};
          for (int k = 1; k < sp; k++) {
            if (k < kSharedStack) s_put(k, sv[12 + k]);
            else l_stack[k - kSharedStack] = sv[12 + k];
          }
          if (pkt >= 0) reload_instance(pkt);
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
      if (cur != kMarkDone && sp <= kSuspendEntries) suspend_lane();
      if (!__any_sync(kFullWarp, cur != kMarkDone)) break;
      // lanes with a deeper stack than the save area keep walking
    }
    if (Source::kLone && !more && src.lone_below > 0 && __popc(busy) <= src.lone_below) {
      // ---- tail: the queue is exhausted and only a few rays of this warp are left. Their remaining walk is a
      // chain of dependent steps; without the per-round votes each step is ~20 % shorter. ----
      // lone_steps > 0 caps the tail: a ray that still walks after that many more steps is parked for the next
      // launch (only the extreme stragglers get there; a parked ray always advances lone_steps per launch)
      const int limit = src.lone_steps > 0 ? src.lone_steps : 0x7fffffff;
      int       steps = 0;
      while (cur != kMarkDone) {
        if (++steps > limit && sp <= kSuspendEntries) {
          suspend_lane();
          break;
        }
        if (cur >= 0 || cur == kMarkExit) {
          visit_node();
        } else if (cur == kMarkLeaf) {
          while (leaf_next < leaf_end) test_prim(leaf_next++);
          cur = pop();
        } else {
          cur = enter(cur);
        }
      }
      break;
    }
    const int n_node = __popc(want_node), n_prim = __popc(want_prim), n_enter = __popc(want_enter);

    // the path most lanes wait for runs; prim_weight / enter_weight (eighths) can tilt the vote towards the two
    // minority paths, whose lanes would otherwise sit out until they outnumber the node walkers
    const int v_node = 8 * n_node, v_prim = src.prim_weight * n_prim, v_enter = src.enter_weight * n_enter;
    if (v_node >= v_prim && v_node >= v_enter) {
      // several node visits per scheduling round: cuts the vote overhead on the most frequent path
#pragma unroll 1
      for (int rep = 0; rep < src.node_reps; rep++)
        if (cur >= 0 || cur == kMarkExit) visit_node();
#ifdef YGL_COOP_PRIMS
    } else if (v_prim >= v_enter) {
      // ---- primitive tests, spread over the whole warp: the leaves of the waiting lanes hold 1-4 primitives each; all
      // of them are numbered (owner lanes in lane order, a leaf's primitives in leaf order) and lane w tests number w,
      // w + 32, ... on behalf of its owner, whose ray it fetches with shuffles. Owners then take their results in leaf
      // order: a primitive counts if the test passed with the range the owner had when the round began AND its distance
      // is not beyond the owner's current tmax - the same decisions as testing one after the other (a hit only ever
      // shrinks tmax; should tmax become NaN, the rest of that leaf is tested one by one). ----
      const int      wl      = tix & 31;
      const bool     in_leaf = cur == kMarkLeaf;
      const int      mine    = in_leaf ? leaf_end - leaf_next : 0;
      const unsigned lt      = (1u << wl) - 1u;
      const unsigned b0 = __ballot_sync(kFullWarp, mine & 1), b1 = __ballot_sync(kFullWarp, mine & 2),
                     b2 = __ballot_sync(kFullWarp, mine & 4);
      const int my_first = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
      const int total    = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
      unsigned char* tasks = s_tasks[tix >> 5];
      for (int k = 0; k < 4; k++)
        if (k < mine) tasks[my_first + k] = (unsigned char)(wl | k << 5);
      __syncwarp();
      if (COUNT && in_leaf) cnt.prims_by_kind[kind] += mine;
      int                      taken = -1, redo_from = 4;
      const unsigned long long pk64  = (unsigned long long)packets;
      for (int base = 0; base < total; base += 32) {
        const int  t    = base + wl;
        const bool work = t < total;
        const int  task = tasks[work ? t : 0];
        const int  from = task & 31, k = task >> 5;
        const f3   o    = {__shfl_sync(kFullWarp, ray.o.x, from), __shfl_sync(kFullWarp, ray.o.y, from), __shfl_sync(kFullWarp, ray.o.z, from)};
        const f3   d    = {__shfl_sync(kFullWarp, ray.d.x, from), __shfl_sync(kFullWarp, ray.d.y, from), __shfl_sync(kFullWarp, ray.d.z, from)};
        const float tmx = __shfl_sync(kFullWarp, tmax, from);
        const int   idx = __shfl_sync(kFullWarp, leaf_next, from) + k;
        const int   kd  = __shfl_sync(kFullWarp, kind, from);
        const float4* pk = (const float4*)(((unsigned long long)__shfl_sync(kFullWarp, (unsigned)(pk64 >> 32), from) << 32) |
                                           __shfl_sync(kFullWarp, (unsigned)pk64, from));
        f2    puv = {0, 0};
        float pd  = 0;
        const bool     h    = work && test_element(o, d, tmx, pk, kd, idx, puv, pd);
        const unsigned hits = __ballot_sync(kFullWarp, h);
        for (int k2 = 0; k2 < 4; k2++) {
          const int   tt = my_first + k2, lane_of = tt & 31;
          const float ru = __shfl_sync(kFullWarp, puv.x, lane_of), rv = __shfl_sync(kFullWarp, puv.y, lane_of),
                      rt = __shfl_sync(kFullWarp, pd, lane_of);
          if (k2 < mine && k2 < redo_from && tt >= base && tt < base + 32 && (hits >> lane_of & 1) && !(rt > tmax)) {
            res.instance = cur_instance, res.uv = {ru, rv}, res.distance = rt, res.hit = true;
            tmax  = rt;
            taken = k2;
            if (rt != rt) redo_from = k2 + 1;
          }
        }
      }
      if (in_leaf) {
        if (taken >= 0) res.element = __ldg(prims + leaf_next + taken);
        for (int k = redo_from; k < mine; k++) test_prim(leaf_next + k);  // (tmax is NaN: practically never)
        leaf_next = leaf_end;
        cur       = pop();
      }
#else
    } else if (v_prim >= v_enter) {
      // all primitives of the lane's leaf (<= 4, bvh_max_prims) in one go, warp-uniform trip count
      const bool in_leaf = cur == kMarkLeaf;
      for (int k = 0; k < 4; k++) {
        const bool test = in_leaf && leaf_next < leaf_end;
        if (!__any_sync(kFullWarp, test)) break;
        if (test) test_prim(leaf_next++);
      }
      if (in_leaf) cur = pop();
#endif
    } else {
      if (cur < 0 && cur > kMarkLeaf) cur = enter(cur);
    }
  }
  // lanes still holding an uncommitted result (more == false path)
  src.commit_finished(have && cur == kMarkDone, res);
}

#endif  // YGL_PAIR_VISIT

}  // namespace ygl
