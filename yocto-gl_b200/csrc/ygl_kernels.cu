// ygl_kernels.cu — the path tracer's stages as device functions, and the two schedulers that drive them (sm_100a).
//
//   generate_lane    (ray generation)        sample_camera                       yocto_trace.cpp:338, 1461-1468
//   trace_stream     (closest-hit traversal) intersect_scene_bvh                 yocto_bvh.cpp:554  (ygl_traverse.cuh)
//   shade_lane       (eval + sample)         trace_path / pathtest / naive / eyelight / diagram / furnace / falsecolor
//   shade_multi                              trace_pathdirect / trace_pathmis    yocto_trace.cpp:599-950
//   lightpdf_lane    (instance traversal)    sample_lights_pdf + MIS weight + RR yocto_trace.cpp:391, 532-590
//   accumulate_lane                          trace_sample tail                   yocto_trace.cpp:1469-1491
//
// Wavefront scheduler (default): k_generate / k_extend / k_shade_path<class> (or k_shade<sampler>) / k_lightpdf /
// k_finish, linked by compacted lane queues. Every kernel is persistent-style: a fixed grid (multiple of the SM count) strides over the queue whose
// length it reads from device memory, so the host never synchronises inside the sample loop. Survivors are appended
// to the next queue with one atomicAdd per warp (__ballot_sync + __popc + __shfl_sync).
// (A second scheduler - one resident kernel with SM-specialised stages linked by ring queues - existed in round 1; it
// was slower than the wavefront on every tile size and its results turned out to depend on timing, so it was
// removed: DESIGN.md §3.3.) Compiled with -fmad=false: see ygl_math.cuh.
#include <algorithm>
#include <cstdlib>

#include "ygl_eval.cuh"
#include "ygl_kernels.cuh"
#include "ygl_traverse.cuh"

namespace ygl {

enum : int { kDestNone = 0, kDestExt = 1, kDestLpdf = 2, kDestAcc = 3 };
enum : int { kFlagHit = 1, kFlagVolume = 2, kFlagNoEmission = 4 };  // NoEmission: next_emission == false
// extend-queue entry = lane id | flags
enum : int { kEntryResume = (int)0x80000000, kEntryShadow = 0x40000000, kEntryPass = 0x20000000, kEntryLane = 0x1fffffff };
// phases of a pathdirect / pathmis bounce (PathState::aux_bsdf.w)
enum : int { kPhaseMain = 0, kPhaseDirectPdf = 1, kPhaseShadowHit = 2, kPhaseShadowSkip = 3, kPhaseNextPdf = 4,
  kPhaseBsdfPdf = 5, kPhaseBsdfHit = 6, kPhaseBsdfSkip = 7 };
enum : int {
  kSamplerPath = 0, kSamplerPathDirect, kSamplerPathMis, kSamplerPathTest, kSamplerNaive, kSamplerEyelight,
  kSamplerDiagram, kSamplerFurnace, kSamplerFalsecolor
};

YGL_D void queue_push(int* __restrict__ q, int* counter, bool pred, int value) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (!m) return;
  int lane   = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  int base   = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = value;
}

YGL_D float4 pack(const f3& v, float w) { return make_float4(v.x, v.y, v.z, w); }
YGL_D float4 pack(const f3& v, int w) { return make_float4(v.x, v.y, v.z, __int_as_float(w)); }
YGL_D f3     unpack3(const float4& v) { return f3{v.x, v.y, v.z}; }

template <class PS>
YGL_D rng_t load_rng(const PS& st, int lane) {
  ulonglong2 r = st.rngs[lane];
  return rng_t{r.x, r.y};
}
template <class PS>
YGL_D void store_rng(const PS& st, int lane, const rng_t& rng) {
  st.rngs[lane] = make_ulonglong2(rng.state, rng.inc);
}

// ------------------------------------------------------------------------------------------
__global__ void k_begin_iteration(Counters* c, int parity) {
  c->n_ext[1 - parity] = 0;
  c->n_gen[1 - parity] = 0;
  c->n_lpdf            = 0;
  c->n_acc             = 0;
  c->ext_head          = 0;
  for (int k = 0; k < kNumClasses; k++) c->n_shade[k] = 0;
}

__global__ void k_seed_lanes(PathState st, Queues q, int parity, int sample_begin, int lane_lo, int lane_hi) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = lane_lo + tid; l < lane_hi; l += gridDim.x * blockDim.x) {
    st.sample[l]               = sample_begin;
    q.gen[parity][l - lane_lo] = l;
  }
  if (tid == 0) {
    Counters* c          = q.counters;
    c->n_gen[parity]     = lane_hi - lane_lo;
    c->n_gen[1 - parity] = 0;
    c->n_ext[0] = c->n_ext[1] = 0;
    c->n_lpdf = c->n_acc = 0;
    for (int k = 0; k < kNumClasses; k++) c->n_shade[k] = 0;
    c->done_lanes        = 0;
    c->ext_head          = 0;
    c->camera_samples = c->scene_rays = c->instance_rays = c->shade_calls = 0;
  }
}

// ---- ray generation: trace_sample head, yocto_trace.cpp:1461-1468. g++ evaluates the two rand2f
// arguments right-to-left, so the lens sample `luv` is drawn BEFORE the pixel sample `puv`. ----
template <class PS>
YGL_D bool accumulate_lane(const DScene& scene, const PS& st, const KParams& p, int lane);

template <class PS>
YGL_D void generate_lane(const DScene& scene, const PS& st, const KParams& p, int lane) {
  const DCamera& camera = scene.cameras[p.camera];
  int px = lane % st.width, py = st.row_begin + (lane / st.width) * st.row_step;
  rng_t rng = load_rng(st, lane);
  f2 luv    = rand2f(rng);
  f2 puv    = rand2f(rng);
  f3 o, d;
  sample_camera(camera, px, py, st.width, st.height, puv, luv, p.tentfilter != 0, o, d);
  store_rng(st, lane, rng);
  st.ray_o[lane]    = pack(o, 0);
  st.ray_d[lane]    = pack(d, 0);
  st.radiance[lane] = pack(f3{0, 0, 0}, 0);
  st.weight[lane]   = pack(f3{1, 1, 1}, 0.0f);
  st.albedo0[lane]  = pack(f3{0, 0, 0}, 0.0f);
  st.normal0[lane]  = pack(-d, 0.0f);
  if (p.sampler == kSamplerPathDirect || p.sampler == kSamplerPathMis) {
    st.aux_bsdf[lane] = pack(f3{0, 0, 0}, kPhaseMain);
    st.next_uvd[lane] = make_float4(0, 0, 0, __int_as_float(0));  // next_intersection = {}
    st.next_ids[lane] = make_int2(-1, -1);
  }
}
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_gen[parity];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int  i     = i0 + wl;
    bool valid = i < n;
    int  lane  = 0;
    if (valid) {
      lane = q.gen[parity][i];
      generate_lane(scene, st, p, lane);
    }
    queue_push(q.ext[parity], &c->n_ext[parity], valid, lane);
  }
  if (tid == 0) atomicAdd(&c->camera_samples, (unsigned long long)n);
}

// ---- extend: closest hit over the instance BVH for every queued ray ----
// Persistent warps: each warp keeps 32 rays in flight and refills finished lanes from the queue
// (extend_stream in ygl_traverse.cuh), so short rays (sky misses) do not idle lanes while long ones walk.
// TAIL: how a drained warp ends a launch - 0 = park the stragglers for the next launch, 1 = vote-free walk (see
// launch_extend). A compile-time choice so that each kernel variant carries only its own tail code.
// Append to the shade queue of class `cls` (WARP-UNIFORM call; lanes with `pred` take part): the lanes of a warp
// are grouped by class with __match_any_sync, one atomicAdd per class present in the warp.
YGL_D void shade_push(const Queues& q, bool pred, int cls, int value) {
  const unsigned active = __ballot_sync(kFullWarp, pred);
  if (!active || !pred) return;
  const unsigned peers  = __match_any_sync(active, cls);
  const int      wl     = threadIdx.x & 31, leader = __ffs(peers) - 1;
  int            base   = 0;
  if (wl == leader) base = atomicAdd(&q.counters->n_shade[cls], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  q.shade[cls][base + __popc(peers & ((1u << wl) - 1u))] = value;
}

template <int TAIL>
struct ExtendSource {
  // st and q are REFERENCES to the kernel's __grid_constant__ parameters: a by-value copy put the whole source
  // (~500 bytes of pointers) into local memory, and every queue or state access in the traversal loop went through it
  const int* __restrict__ queue;
  int        n;
  int*       head;
  const PathState& st;
  const Queues&    q;
  int        parity;
  const unsigned char* __restrict__ inst_class;  // shading class per instance, or null: every lane goes to class 0
  int        has_volumes;
  int        lane;
  int        refill_thr, node_reps, suspend_below, lone_below, lone_steps;  // tuning knobs (see launch_extend)
  int        prim_weight, enter_weight;                                     //   (path vote weights, in eighths)
  unsigned   finished;                              // rays completed by this thread (scene_rays counter)
  bool       shadow;                                // current ray is a shadow ray (pathdirect / pathmis)
  // Queue entries: lane id; bit 31 set = the ray was suspended by the previous launch (resume it).
  YGL_D bool fetch(bool idle, f3& o, f3& d, bool& more, bool& resume) {
    const unsigned m    = __ballot_sync(kFullWarp, idle);
    const int      wl   = threadIdx.x & 31;
    int            base = 0;
    if (wl == __ffs(m) - 1) base = atomicAdd(head, __popc(m));
    base         = __shfl_sync(kFullWarp, base, __ffs(m) - 1);
    const int my = base + __popc(m & ((1u << wl) - 1u));
    more         = base + __popc(m) < n;
    const bool valid = idle && my < n;
    const int  entry = valid ? queue[my] : 0;
    // the lane skips this extend (its hit record is already in place): straight on to shading
    const bool pass = valid && (entry & kEntryPass) != 0;
    shade_push(q, pass, kClsGeneric, entry);
    if (!valid || pass) return false;
    lane     = entry & kEntryLane;
    resume   = entry < 0;
    shadow   = (entry & kEntryShadow) != 0;
    float4 a = shadow ? st.aux_o[lane] : st.ray_o[lane], b = shadow ? st.aux_dir[lane] : st.ray_d[lane];
    o = unpack3(a), d = unpack3(b);
    return true;
  }
  static constexpr bool kPolling = false;  // the queue is complete when the kernel starts
  static constexpr bool kPark = TAIL == 0, kLone = TAIL == 1;
  // WARP-UNIFORM: lanes with `flag` store their hit and join the shade queue of their class
  YGL_D void commit_finished(bool flag, const hit_t& h) {
    int cls = kClsGeneric;
    if (flag) {
      (shadow ? st.aux_uvd : st.hit_uvd)[lane] = make_float4(h.uv.x, h.uv.y, h.distance, __int_as_float(h.hit ? 1 : 0));
      if (shadow) st.aux_ids[lane] = make_int2(h.instance, h.element);  // (the two arrays differ in stride: no ?: between them)
      else st.hit_ids[lane] = make_int2(h.instance, h.element);
      finished++;
      if (inst_class) {
        cls = h.hit ? (int)__ldg(inst_class + h.instance) : kClsMiss;
        // a path inside a participating medium takes the unspecialised kernel (transmittance, scattering events)
        if (has_volumes && (__float_as_int(float4(st.radiance[lane]).w) & kFlagVolume)) cls = kClsGeneric;
      }
    }
    shade_push(q, flag, cls, lane | (shadow ? kEntryShadow : 0));
  }
  // Parked rays: a thread parks at most one ray per launch, into ITS slot of this launch's half of the pool; the ray
  // resumes in the next launch (other parity), which reads that half while parking into the other one.
  YGL_D int* save_slot() { return q.park[parity] + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * kSuspendWords; }
  YGL_D const int* load_slot() { return q.park[1 - parity] + (size_t)(int)st.susp[lane] * kSuspendWords; }
  // per-lane (divergent) call: the parked ray goes straight back into the next extend queue, flagged for resumption
  YGL_D void commit_suspended() {
    st.susp[lane]           = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int pos           = atomicAdd(&q.counters->n_ext[1 - parity], 1);
    q.ext[1 - parity][pos]  = lane | kEntryResume | (shadow ? kEntryShadow : 0);
  }
};

#ifndef YGL_EXT_MINBLOCKS
#define YGL_EXT_MINBLOCKS 7  // measured on B200 (C3): 4/5/6/7/8 blocks/SM -> 249/249/233/224/227 ms per 32 spp
#endif
template <bool COUNT, int TAIL, int STACK, bool TOP = false>
__global__ void __launch_bounds__(128, YGL_EXT_MINBLOCKS) k_extend(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, int parity,
    unsigned long long* trav, int refill_thr, int node_reps, int suspend_below, int lone_below, int lone_steps) {
  Counters*     c = q.counters;
  const int     n = c->n_ext[parity];
  ExtendSource<TAIL> src{q.ext[parity], n, &c->ext_head, st, q, parity, scene.inst_class, scene.has_volumes, 0, refill_thr,
      node_reps & 0xff, suspend_below, lone_below, lone_steps, (node_reps >> 8) & 0xff, (node_reps >> 16) & 0xff, 0, false};
  trav_counters tc = {};
  trace_stream<COUNT, kStreamThreads, kSharedStack, STACK, ExtendSource<TAIL>, TOP>(scene, src, tc);
  {
    unsigned v = src.finished;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(kFullWarp, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&c->scene_rays, (unsigned long long)v);
  }
  if (COUNT) {
    // traversal statistics for the algorithmic-bytes formula; counting runs are not timed
    unsigned vals[7] = {tc.top_nodes, tc.bot_nodes, tc.instances, tc.prims_by_kind[kElemTriangles],
        tc.prims_by_kind[kElemQuads], tc.prims_by_kind[kElemLines], tc.prims_by_kind[kElemPoints]};
    for (int k = 0; k < 7; k++) {
      unsigned v = vals[k];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(kFullWarp, v, off);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(trav + k, (unsigned long long)v);
    }
  }
}

// ---- light sampling, yocto_trace.cpp:361-388 ----
YGL_D f3 sample_lights(const DScene& scene, const f3& position, float rl, float rel, const f2& ruv) {
  if (scene.num_lights == 0) return {0, 0, 0};  // the reference indexes lights[-1] here (UB); we stop the path
  int           light_id = sample_uniform(scene.num_lights, rl);
  const DLight& light    = scene.lights[light_id];
  if (light.instance >= 0) {
    const DInstance& inst  = scene.instances[light.instance];
    const DShape&    shape = scene.shapes[inst.shape];
    int  element   = sample_discrete(light.cdf, light.cdf_size, rel);
    f2   uv        = (shape.num_triangles > 0) ? sample_triangle(ruv) : ruv;
    auto lposition = eval_position(scene, inst, element, uv);
    return normalize(lposition - position);
  } else if (light.environment >= 0) {
    const DEnvironment& env = scene.environments[light.environment];
    if (env.emission_tex >= 0) {
      const DTexture& tex = scene.textures[env.emission_tex];
      int idx = sample_discrete(light.cdf, light.cdf_size, rel);
      f2  uv  = f2{((idx % tex.width) + 0.5f) / tex.width, ((idx / tex.width) + 0.5f) / tex.height};
      return transform_direction(env.frame, f3{ycos(uv.x * 2 * kPi) * ysin(uv.y * kPi), ycos(uv.y * kPi),
                                                ysin(uv.x * 2 * kPi) * ysin(uv.y * kPi)});
    } else {
      return sample_sphere(ruv);
    }
  }
  return {0, 0, 0};
}

// ---- sample_lights_pdf, yocto_trace.cpp:391-443. WARP-COOPERATIVE (calls trace_ray): the per-light
// chain of up to 100 instance rays is unrolled into one warp-uniform loop in which every lane that
// still owes a ray traces its next one; lanes add their terms in the reference's order. ----
template <bool DEEP>
YGL_D float sample_lights_pdf(const DScene& scene, bool valid, const f3& position, const f3& direction,
    unsigned& rays) {
  float pdf = 0.0f, lpdf = 0.0f;
  int   li = 0, bounce = 0;
  f3    next_position = position;
  bool  done          = !valid;
  trav_counters tc;
  while (true) {
    // walk this lane's light list up to the next area light (environment terms need no ray)
    int instance = -1;
    while (!done) {
      if (li >= scene.num_lights) {
        done = true;
        break;
      }
      const DLight& light = scene.lights[li];
      if (light.instance >= 0) {
        instance = light.instance;
        break;
      }
      if (light.environment >= 0) {
        const DEnvironment& env = scene.environments[light.environment];
        if (env.emission_tex >= 0) {
          const DTexture& tex = scene.textures[env.emission_tex];
          auto wl       = transform_direction(env.inv_frame, direction);
          auto texcoord = f2{yatan2(wl.z, wl.x) / (2 * kPi), yacos(yclamp(wl.y, -1.0f, 1.0f)) / kPi};
          if (texcoord.x < 0) texcoord.x += 1;
          int  i    = iclamp((int)(texcoord.x * tex.width), 0, tex.width - 1);
          int  j    = iclamp((int)(texcoord.y * tex.height), 0, tex.height - 1);
          auto prob = sample_discrete_pdf(light.cdf, j * tex.width + i) / __ldg(light.cdf + light.cdf_size - 1);
          auto angle = (2 * kPi / tex.width) * (kPi / tex.height) * ysin(kPi * (j + 0.5f) / tex.height);
          pdf += prob / angle;
        } else {
          pdf += 1 / (4 * kPi);
        }
      }
      li++;
    }
    const bool want = !done;
    if (!__any_sync(kFullWarp, want)) break;
    hit_t h = trace_ray<false, false, DEEP>(scene, want, next_position, direction, kRayEps, kFltMax, want ? instance : -1, tc);
    if (want) {
      rays++;
      bool next_light = !h.hit;
      if (h.hit) {
        const DLight&    light = scene.lights[li];
        const DInstance& inst  = scene.instances[instance];
        auto lposition = eval_position(scene, inst, h.element, h.uv);
        auto lnormal   = eval_element_normal(scene, inst, h.element);
        auto area      = __ldg(light.cdf + light.cdf_size - 1);
        lpdf += distance_squared(lposition, position) / (yabs(dot(lnormal, direction)) * area);
        next_position = lposition + direction * 1e-3f;
        next_light    = ++bounce >= 100;
      }
      if (next_light) {
        pdf += lpdf;
        lpdf = 0.0f, bounce = 0, next_position = position;
        li++;
      }
    }
  }
  pdf *= sample_uniform_pdf(scene.num_lights);
  return pdf;
}

// throughput update of a sampled direction: value / pdf of its lobe (one evaluation gives both)
YGL_D f3 lobe_weight(const lobe_t& lobe) { return lobe.bsdfcos / lobe.pdf; }

// tail of one trace_path iteration, yocto_trace.cpp:581-591 + the for-loop increment
YGL_D int finish_bounce(f3& weight, int& bounce, rng_t& rng, const KParams& p) {
  if (is_zero(weight) || !vfinite(weight)) return kDestAcc;
  if (bounce > 3) {
    auto rr_prob = ymin((float)0.99, max3(weight));
    if (rand1f(rng) >= rr_prob) return kDestAcc;
    weight = weight * (1 / rr_prob);
  }
  bounce += 1;
  return bounce < p.bounces ? kDestExt : kDestAcc;
}

// hashed_color of trace_falsecolor, yocto_trace.cpp:1357-1361 (std::hash<int> is the identity)
YGL_D f3 hashed_color(int id) {
  rng_t rng = rng_make(961748941ull, (uint64_t)(int64_t)id);
  f3    r   = rand3f(rng);
  f3    v   = 0.5f + 0.5f * r;
  return {ypow(v.x, 2.2f), ypow(v.y, 2.2f), ypow(v.z, 2.2f)};
}

// ---- shade one lane. Returns the destination queue. ----
// CLS: shading class the lane was queued under (ygl_kernels.cuh). kClsGeneric = any lane. 1 + t = the lane hit a
// surface whose material has type t and is not inside a medium: the material switches fold to that type's lobes and
// the medium code drops out. A specialised instantiation computes exactly what the generic one would for such a
// lane (it only omits code the lane cannot reach). Misses of the path sampler have their own function (miss_lane).
template <int SAMPLER, int CLS = kClsGeneric, class PS>
YGL_D int shade_lane(const DScene& scene, const PS& st, const KParams& p, int lane, unsigned& inst_rays) {
  constexpr bool kSurface = CLS >= 1 && CLS <= 8;  // a surface hit of material type CLS - 1, outside any medium
  float4 ro = st.ray_o[lane], rd = st.ray_d[lane], rad4 = st.radiance[lane], w4 = st.weight[lane];
  float4 huvd = st.hit_uvd[lane];
  f3  o = unpack3(ro), d = unpack3(rd), radiance = unpack3(rad4), weight = unpack3(w4);
  int bounce = __float_as_int(ro.w), opbounce = __float_as_int(rd.w), flags = __float_as_int(rad4.w);
  float max_roughness = w4.w;
  bool  hit           = kSurface ? true : __float_as_int(huvd.w) != 0;

  // ---- falsecolor: single intersection, yocto_trace.cpp:1341-1419 ----
  if (SAMPLER == kSamplerFalsecolor) {
    if (!hit) return kDestAcc;  // trace_result{}: radiance 0, hit false
    int2 ids = st.hit_ids[lane];
    const DInstance& inst = scene.instances[ids.x];
    f2   uv       = {huvd.x, huvd.y};
    auto outgoing = -d;
    auto position = eval_shading_position(scene, inst, ids.y, uv);
    auto normal   = eval_shading_normal(scene, inst, ids.y, uv, outgoing);
    auto gnormal  = eval_element_normal(scene, inst, ids.y);
    auto texcoord = eval_texcoord(scene, inst, ids.y, uv);
    auto material = eval_material(scene, inst, ids.y, uv);
    auto delta    = is_delta(material) ? 1.0f : 0.0f;
    f3   result   = {0, 0, 0};
    switch (p.falsecolor) {
      case 0: result = position * 0.5f + 0.5f; break;
      case 1: result = normal * 0.5f + 0.5f; break;
      case 2: result = dot(normal, -d) > 0 ? f3{0, 1, 0} : f3{1, 0, 0}; break;
      case 3: result = gnormal * 0.5f + 0.5f; break;
      case 4: result = dot(gnormal, -d) > 0 ? f3{0, 1, 0} : f3{1, 0, 0}; break;
      case 5: result = {yfmod(texcoord.x, 1.0f), yfmod(texcoord.y, 1.0f), 0}; break;
      case 6: result = hashed_color(material.type); break;
      case 7: result = material.color; break;
      case 8: result = material.emission; break;
      case 9: result = {material.roughness, material.roughness, material.roughness}; break;
      case 10: result = {material.opacity, material.opacity, material.opacity}; break;
      case 11: result = {material.metallic, material.metallic, material.metallic}; break;
      case 12: result = {delta, delta, delta}; break;
      case 13: result = hashed_color(ids.x); break;
      case 14: result = hashed_color(inst.shape); break;
      case 15: result = hashed_color(inst.material); break;
      case 16: result = hashed_color(ids.y); break;
      case 17: {
        if (is_zero(material.emission)) material.emission = {0.2f, 0.2f, 0.2f};
        result = material.emission * yabs(dot(-d, normal));
      } break;
      default: result = {0, 0, 0};
    }
    radiance          = {srgb_to_rgb(result.x), srgb_to_rgb(result.y), srgb_to_rgb(result.z)};
    st.radiance[lane] = pack(radiance, kFlagHit);
    st.albedo0[lane]  = pack(material.color, 0.0f);
    st.normal0[lane]  = pack(normal, 0.0f);
    return kDestAcc;
  }

  // ---- furnace: a path that left the object stops at the next loop head, yocto_trace.cpp:1262-1266
  // (the ray was traced but its hit is ignored, which is equivalent: tracing has no side effect) ----
  if (SAMPLER == kSamplerFurnace && bounce > 0 && !(flags & kFlagVolume)) {
    radiance          = radiance + weight * eval_environment(scene, d);
    st.radiance[lane] = pack(radiance, flags);
    return kDestAcc;
  }

  // ---- miss: environment, yocto_trace.cpp:469-473 / :1127-1131 ----
  if (!hit) {
    if (SAMPLER == kSamplerDiagram) {
      // trace_diagram, yocto_trace.cpp:1194-1199: a miss is white and counts as a hit (albedo / normal stay
      // at their bounce-0 values, or zero if there was no surface before)
      radiance = radiance + weight * f3{1, 1, 1};
      if (!(flags & kFlagHit)) st.normal0[lane] = pack(f3{0, 0, 0}, 0.0f);
      flags |= kFlagHit;
    } else if (bounce > 0 || !p.envhidden) {
      radiance = radiance + weight * eval_environment(scene, d);
    }
    st.radiance[lane] = pack(radiance, flags);
    return kDestAcc;
  }

  int2 ids = st.hit_ids[lane];
  const DInstance& inst = scene.instances[ids.x];
  f2    uv       = {huvd.x, huvd.y};
  float distance = huvd.z;
  rng_t rng      = load_rng(st, lane);
  int   dest     = kDestAcc;

  // ---- participating medium, yocto_trace.cpp:476-488 (path only). Draw order: rd, then rl. ----
  bool   in_volume = false;
  vsdf_t vsdf      = {};
  if (SAMPLER == kSamplerPath && !kSurface && (flags & kFlagVolume)) {
    float4 va = st.vol_a[lane], vb = st.vol_b[lane];
    vsdf     = {unpack3(va), unpack3(vb), va.w};
    float r_d = rand1f(rng);
    float r_l = rand1f(rng);
    float dist = sample_transmittance(vsdf.density, distance, r_l, r_d);
    weight     = weight * (eval_transmittance(vsdf.density, dist) / sample_transmittance_pdf(vsdf.density, dist, distance));
    in_volume  = dist < distance;
    distance   = dist;
  }

  if (!in_volume) {
    auto outgoing = -d;
    auto position = SAMPLER == kSamplerFurnace ? eval_position(scene, inst, ids.y, uv)  // yocto_trace.cpp:1282
                                               : eval_shading_position(scene, inst, ids.y, uv);
    auto normal   = eval_shading_normal(scene, inst, ids.y, uv, outgoing);
    auto material = eval_material(scene, inst, ids.y, uv);
    if (kSurface) material.type = CLS - 1;  // what eval_material returned: now a compile-time constant

    if (SAMPLER == kSamplerPath && p.nocaustics) {
      max_roughness      = ymax(material.roughness, max_roughness);
      material.roughness = max_roughness;
    }

    if (SAMPLER == kSamplerPathTest) material.type = kMatte;  // yocto_trace.cpp:981 (after eval_material)

    // opacity pass-through, yocto_trace.cpp:505-510 (rng drawn only when opacity < 1); trace_pathtest has none
    if (SAMPLER != kSamplerPathTest && material.opacity < 1 && rand1f(rng) >= material.opacity) {
      store_rng(st, lane, rng);
      if (opbounce++ > 128) {
        st.radiance[lane] = pack(radiance, flags);
        return kDestAcc;
      }
      st.ray_o[lane]  = pack(position + d * 1e-2f, bounce);  // bounce -= 1; continue; bounce++
      st.ray_d[lane]  = pack(d, opbounce);
      st.weight[lane] = pack(weight, max_roughness);
      return kDestExt;
    }

    if (bounce == 0) {
      flags |= kFlagHit;
      st.albedo0[lane] = pack(material.color, 0.0f);
      st.normal0[lane] = pack(normal, 0.0f);
    }

    radiance = radiance + weight * eval_emission(material, normal, outgoing);

    if (SAMPLER == kSamplerEyelight || SAMPLER == kSamplerDiagram) {
      // yocto_trace.cpp:1155-1172 / :1224-1240 (trace_diagram differs from trace_eyelight only on a miss)
      auto incoming = outgoing;
      radiance      = radiance + weight * kPi * eval_bsdfcos(material, normal, outgoing, incoming);
      dest          = kDestAcc;
      if (is_delta(material)) {
        incoming = sample_delta(material, normal, outgoing, rand1f(rng));
        if (!is_zero(incoming)) {
          weight = weight * lobe_weight(delta_lobe(material, normal, outgoing, incoming));
          if (!(is_zero(weight) || !vfinite(weight))) {
            bounce += 1;
            if (bounce < imax(p.bounces, 4)) {
              dest           = kDestExt;
              st.ray_o[lane] = pack(position, bounce);
              st.ray_d[lane] = pack(incoming, opbounce);
            }
          }
        }
      }
      store_rng(st, lane, rng);
      st.radiance[lane] = pack(radiance, flags);
      st.weight[lane]   = pack(weight, max_roughness);
      return dest;
    }

    if (SAMPLER == kSamplerNaive || SAMPLER == kSamplerFurnace) {
      // trace_naive / trace_furnace, yocto_trace.cpp:1078-1100 / :1305-1336: bsdf sampling only
      f3 incoming = {0, 0, 0};
      if (material.roughness != 0) {
        f2    rn  = rand2f(rng);  // g++: rand2f evaluated before rand1f
        float rnl = rand1f(rng);
        incoming  = sample_bsdfcos(material, normal, outgoing, rnl, rn);
        if (!is_zero(incoming))
          weight = weight * lobe_weight(bsdf_lobe(material, normal, outgoing, incoming));
      } else {
        incoming = sample_delta(material, normal, outgoing, rand1f(rng));
        if (!is_zero(incoming))
          weight = weight * lobe_weight(delta_lobe(material, normal, outgoing, incoming));
      }
      dest = is_zero(incoming) ? kDestAcc : finish_bounce(weight, bounce, rng, p);
      if (SAMPLER == kSamplerFurnace && dot(normal, outgoing) * dot(normal, incoming) < 0) flags ^= kFlagVolume;
      store_rng(st, lane, rng);
      st.ray_o[lane]    = pack(position, bounce);
      st.ray_d[lane]    = pack(incoming, opbounce);
      st.radiance[lane] = pack(radiance, flags);
      st.weight[lane]   = pack(weight, max_roughness);
      return dest;
    }

    // ---- next direction, yocto_trace.cpp:522-542 ----
    f3 incoming = {0, 0, 0};
    if (!is_delta(material)) {
      if (rand1f(rng) < 0.5f) {
        f2    rn  = rand2f(rng);  // g++: rand2f evaluated before rand1f
        float rnl = rand1f(rng);
        incoming  = sample_bsdfcos(material, normal, outgoing, rnl, rn);
      } else {
        f2    ruv = rand2f(rng);  // g++ order: ruv, rel, rl
        float rel = rand1f(rng);
        float rl  = rand1f(rng);
        incoming  = sample_lights(scene, position, rl, rel, ruv);
      }
      store_rng(st, lane, rng);
      if (is_zero(incoming)) {
        st.radiance[lane] = pack(radiance, flags);
        return kDestAcc;
      }
      auto lobe      = bsdf_lobe(material, normal, outgoing, incoming);
      st.pend[lane]  = pack(lobe.bsdfcos, lobe.pdf);
      dest           = kDestLpdf;
    } else {
      incoming = sample_delta(material, normal, outgoing, rand1f(rng));
      weight   = weight * lobe_weight(delta_lobe(material, normal, outgoing, incoming));
    }

    // ---- volume slot update, yocto_trace.cpp:545-553 (independent of the pending weight) ----
    const int mat_type = kSurface ? CLS - 1 : scene.materials[inst.material].type;
    if (SAMPLER != kSamplerPathTest && is_volumetric_type(mat_type) &&
        dot(normal, outgoing) * dot(normal, incoming) < 0) {
      if (!(flags & kFlagVolume)) {
        auto vm        = eval_material(scene, inst, ids.y, uv);
        st.vol_a[lane] = pack(vm.density, vm.scanisotropy);
        st.vol_b[lane] = pack(vm.scattering, 0.0f);
        flags |= kFlagVolume;
      } else {
        flags &= ~kFlagVolume;
      }
    }

    if (dest != kDestLpdf) {  // delta: finish the bounce here
      dest = finish_bounce(weight, bounce, rng, p);
      store_rng(st, lane, rng);
    }
    st.ray_o[lane]    = pack(position, bounce);
    st.ray_d[lane]    = pack(incoming, opbounce);
    st.radiance[lane] = pack(radiance, flags);
    st.weight[lane]   = pack(weight, max_roughness);
    return dest;
  } else {
    // ---- scattering event inside the medium, yocto_trace.cpp:557-579 ----
    auto outgoing = -d;
    auto position = o + d * distance;
    f3   incoming = {0, 0, 0};
    if (rand1f(rng) < 0.5f) {
      f2    rn  = rand2f(rng);
      float rnl = rand1f(rng);  // drawn and unused, as in the reference
      (void)rnl;
      incoming = sample_scattering(vsdf, outgoing, rn);
    } else {
      f2    ruv = rand2f(rng);
      float rel = rand1f(rng);
      float rl  = rand1f(rng);
      incoming  = sample_lights(scene, position, rl, rel, ruv);
    }
    store_rng(st, lane, rng);
    st.radiance[lane] = pack(radiance, flags);
    if (is_zero(incoming)) return kDestAcc;
    auto phase      = scattering_lobe(vsdf, outgoing, incoming);
    st.pend[lane]   = pack(phase.bsdfcos, phase.pdf);
    st.ray_o[lane]  = pack(position, bounce);
    st.ray_d[lane]  = pack(incoming, opbounce);
    st.weight[lane] = pack(weight, max_roughness);
    return kDestLpdf;
  }
}


// trace_path for a ray that left the scene (class kClsMiss), yocto_trace.cpp:469-473: the environment seen along the
// ray ends the path. Always returns kDestAcc.
template <class PS>
YGL_D int miss_lane(const DScene& scene, const PS& st, const KParams& p, int lane) {
  float4 ro = st.ray_o[lane], rd = st.ray_d[lane], rad4 = st.radiance[lane], w4 = st.weight[lane];
  const int bounce = __float_as_int(ro.w);
  if (bounce > 0 || !p.envhidden) {
    f3 radiance       = unpack3(rad4) + unpack3(w4) * eval_environment(scene, unpack3(rd));
    st.radiance[lane] = pack(radiance, __float_as_int(rad4.w));
  }
  return kDestAcc;
}

// ---- pathdirect / pathmis: trace_pathdirect (yocto_trace.cpp:599-767) and trace_pathmis (:770-950). A non-delta
// bounce needs extra light-pdf and shadow-ray stages, so a lane passes through shade several times per bounce;
// the phase (PathState::aux_bsdf.w) says where it is. Returns dest | extend-entry flags. ----
YGL_D float mis_heuristic(float this_pdf, float other_pdf) {  // yocto_trace.cpp:785-788
  return (this_pdf * this_pdf) / (this_pdf * this_pdf + other_pdf * other_pdf);
}
// emission seen along a shadow ray, yocto_trace.cpp:667-677 / :874-886
template <class PS>
YGL_D f3 shadow_emission(const DScene& scene, const PS& st, int lane, const f3& incoming) {
  float4 h = st.aux_uvd[lane];
  if (__float_as_int(h.w) == 0) return eval_environment(scene, incoming);
  int2             ids  = st.aux_ids[lane];
  const DInstance& inst = scene.instances[ids.x];
  f2               uv   = {h.x, h.y};
  auto material = eval_material(scene, inst, ids.y, uv);
  auto normal   = eval_shading_normal(scene, inst, ids.y, uv, -incoming);
  return eval_emission(material, normal, -incoming);
}

template <int SAMPLER, class PS>
YGL_D int shade_multi(const DScene& scene, const PS& st, const KParams& p, int lane) {
  constexpr bool MIS = SAMPLER == kSamplerPathMis;
  float4 ro = st.ray_o[lane], rd = st.ray_d[lane], rad4 = st.radiance[lane], w4 = st.weight[lane];
  float4 huvd = st.hit_uvd[lane], ab = st.aux_bsdf[lane];
  f3  o = unpack3(ro), d = unpack3(rd), radiance = unpack3(rad4), weight = unpack3(w4);
  int bounce = __float_as_int(ro.w), opbounce = __float_as_int(rd.w), flags = __float_as_int(rad4.w);
  float max_roughness = w4.w;
  bool  hit   = __float_as_int(huvd.w) != 0;
  int   phase = __float_as_int(ab.w);
  const bool next_emission = !(flags & kFlagNoEmission);

  if (!hit) {  // only reachable in kPhaseMain (a lane in a later phase has a surface hit)
    if ((bounce > 0 || !p.envhidden) && next_emission) radiance = radiance + weight * eval_environment(scene, d);
    st.radiance[lane] = pack(radiance, flags);
    return kDestAcc;
  }
  int2 ids = st.hit_ids[lane];
  const DInstance& inst = scene.instances[ids.x];
  f2    uv       = {huvd.x, huvd.y};
  float distance = huvd.z;
  rng_t rng      = load_rng(st, lane);

  auto store_common = [&]() {
    store_rng(st, lane, rng);
    st.radiance[lane] = pack(radiance, flags);
    st.weight[lane]   = pack(weight, max_roughness);
  };

  // ---- participating medium (first pass of the bounce only) ----
  if (phase == kPhaseMain && (flags & kFlagVolume)) {
    float4 va = st.vol_a[lane], vb = st.vol_b[lane];
    vsdf_t vsdf = {unpack3(va), unpack3(vb), va.w};
    float  r_d = rand1f(rng), r_l = rand1f(rng);
    float  dist = sample_transmittance(vsdf.density, distance, r_l, r_d);
    weight = weight * (eval_transmittance(vsdf.density, dist) / sample_transmittance_pdf(vsdf.density, dist, distance));
    if (dist < distance) {
      // scattering event, yocto_trace.cpp:738-757 / :921-941 (pathmis has no zero-direction check)
      auto outgoing = -d;
      auto position = o + d * dist;
      f3   incoming = {0, 0, 0};
      if (rand1f(rng) < 0.5f) {
        f2 rn = rand2f(rng);
        (void)rand1f(rng);
        incoming = sample_scattering(vsdf, outgoing, rn);
      } else {
        f2    ruv = rand2f(rng);
        float rel = rand1f(rng), rl = rand1f(rng);
        incoming  = sample_lights(scene, position, rl, rel, ruv);
      }
      if (MIS) flags &= ~kFlagNoEmission;  // next_emission = true
      store_common();
      if (!MIS && is_zero(incoming)) return kDestAcc;
      auto phase        = scattering_lobe(vsdf, outgoing, incoming);
      st.pend[lane]     = pack(phase.bsdfcos, phase.pdf);
      st.ray_o[lane]    = pack(position, bounce);
      st.ray_d[lane]    = pack(incoming, opbounce);
      st.aux_bsdf[lane] = pack(f3{0, 0, 0}, kPhaseNextPdf);
      return kDestLpdf;
    }
  }

  // ---- the surface point (recomputed identically on every pass: it is a pure function of the hit) ----
  auto outgoing = -d;
  auto position = eval_shading_position(scene, inst, ids.y, uv);
  auto normal   = eval_shading_normal(scene, inst, ids.y, uv, outgoing);
  auto material = eval_material(scene, inst, ids.y, uv);
  if (p.nocaustics) {
    max_roughness      = ymax(material.roughness, max_roughness);
    material.roughness = max_roughness;
  }
  const bool delta = is_delta(material);
  f3         incoming = {0, 0, 0};
  bool       have_next = false;  // `incoming` is the path's next direction and its weight update is done/pending

  if (phase == kPhaseMain) {
    if (material.opacity < 1 && rand1f(rng) >= material.opacity) {
      store_rng(st, lane, rng);
      if (opbounce++ > 128) {
        st.radiance[lane] = pack(radiance, flags);
        return kDestAcc;
      }
      st.ray_o[lane]  = pack(position + d * 1e-2f, bounce);
      st.ray_d[lane]  = pack(d, opbounce);
      st.weight[lane] = pack(weight, max_roughness);
      // pathmis after a non-delta bounce re-uses next_intersection instead of tracing (yocto_trace.cpp:793-794):
      // the lane's hit record stays as it is and the extend is skipped
      return kDestExt | ((MIS && !next_emission) ? kEntryPass : 0);
    }
    if (bounce == 0) {
      flags |= kFlagHit;
      st.albedo0[lane] = pack(material.color, 0.0f);
      st.normal0[lane] = pack(normal, 0.0f);
    }
    if (next_emission) radiance = radiance + weight * eval_emission(material, normal, outgoing);
    if (!delta) {
      // first direct sample: the lights (draw order ruv, rel, rl)
      f2    ruv = rand2f(rng);
      float rel = rand1f(rng), rl = rand1f(rng);
      f3    dl  = sample_lights(scene, position, rl, rel, ruv);
      if (MIS && is_zero(dl)) {
        // `break` out of the sample loop: the indirect update then runs with incoming = 0 (yocto_trace.cpp:857, :892)
        weight = weight * lobe_weight(bsdf_lobe(material, normal, outgoing, dl));
        flags |= kFlagNoEmission;
        incoming = dl, have_next = true;
      } else {
        auto light_lobe   = bsdf_lobe(material, normal, outgoing, dl);
        auto bsdfcos      = light_lobe.bsdfcos;
        st.aux_o[lane]    = pack(position, MIS ? light_lobe.pdf : 0.0f);
        st.aux_dir[lane]  = pack(dl, 0.0f);
        st.aux_bsdf[lane] = pack(bsdfcos, kPhaseDirectPdf);
        store_common();
        return kDestLpdf;
      }
    } else if (!MIS) {
      flags &= ~kFlagNoEmission;  // pathdirect: next_emission = true after a delta surface
    }
  }

  if (phase == kPhaseShadowHit || phase == kPhaseShadowSkip) {
    // contribution of the light sample
    float4 ad = st.aux_dir[lane];
    if (phase == kPhaseShadowHit) {
      f3 emission = shadow_emission(scene, st, lane, unpack3(ad));
      if (MIS) radiance = radiance + weight * unpack3(ab) * emission * ad.w;
      else radiance = radiance + weight * unpack3(ab) * emission / ad.w;
    }
    if (!MIS) {
      flags |= kFlagNoEmission;  // pathdirect: next_emission = false
    } else {
      // second direct sample: the bsdf (draw order rn, rnl)
      f2    rn  = rand2f(rng);
      float rnl = rand1f(rng);
      f3    db  = sample_bsdfcos(material, normal, outgoing, rnl, rn);
      if (is_zero(db)) {
        weight = weight * lobe_weight(bsdf_lobe(material, normal, outgoing, db));
        flags |= kFlagNoEmission;
        incoming = db, have_next = true;
      } else {
        auto bsdf_side    = bsdf_lobe(material, normal, outgoing, db);
        auto bsdfcos      = bsdf_side.bsdfcos;
        st.aux_o[lane]    = pack(position, bsdf_side.pdf);
        st.aux_dir[lane]  = pack(db, 0.0f);
        st.aux_bsdf[lane] = pack(bsdfcos, kPhaseBsdfPdf);
        store_common();
        return kDestLpdf;
      }
    }
  }

  if (MIS && (phase == kPhaseBsdfHit || phase == kPhaseBsdfSkip)) {
    float4 ad = st.aux_dir[lane], ao = st.aux_o[lane];
    incoming  = unpack3(ad);
    if (phase == kPhaseBsdfHit) {
      st.next_uvd[lane] = st.aux_uvd[lane];  // next_intersection = intersection (yocto_trace.cpp:873)
      st.next_ids[lane] = st.aux_ids[lane];
      f3 emission       = shadow_emission(scene, st, lane, incoming);
      radiance          = radiance + weight * unpack3(ab) * emission * ad.w;
    }
    weight = weight * (unpack3(ab) / ao.w);  // eval_bsdfcos / sample_bsdfcos_pdf of the bsdf sample (same values)
    flags |= kFlagNoEmission;
    have_next = true;
  }

  int dest;
  if (!have_next) {
    if (!delta) {
      // pathdirect: next direction with one-sample MIS exactly like trace_path (yocto_trace.cpp:691-707)
      if (rand1f(rng) < 0.5f) {
        f2    rn  = rand2f(rng);
        float rnl = rand1f(rng);
        incoming  = sample_bsdfcos(material, normal, outgoing, rnl, rn);
      } else {
        f2    ruv = rand2f(rng);
        float rel = rand1f(rng), rl = rand1f(rng);
        incoming  = sample_lights(scene, position, rl, rel, ruv);
      }
      if (is_zero(incoming)) {
        store_common();
        return kDestAcc;
      }
      auto lobe         = bsdf_lobe(material, normal, outgoing, incoming);
      st.pend[lane]     = pack(lobe.bsdfcos, lobe.pdf);
      st.aux_bsdf[lane] = pack(f3{0, 0, 0}, kPhaseNextPdf);
      dest              = kDestLpdf;
    } else {
      incoming = sample_delta(material, normal, outgoing, rand1f(rng));
      if (!MIS && is_zero(incoming)) {  // pathdirect only (yocto_trace.cpp:710)
        store_common();
        return kDestAcc;
      }
      weight = weight * lobe_weight(delta_lobe(material, normal, outgoing, incoming));
      if (MIS) flags &= ~kFlagNoEmission;  // next_emission = true
      dest = kDestNone;
    }
  } else {
    dest = kDestNone;
  }

  // ---- volume slot update, yocto_trace.cpp:716-724 / :903-911 ----
  const DMaterial& mat = scene.materials[inst.material];
  if (is_volumetric_type(mat.type) && dot(normal, outgoing) * dot(normal, incoming) < 0) {
    if (!(flags & kFlagVolume)) {
      auto vm        = eval_material(scene, inst, ids.y, uv);
      st.vol_a[lane] = pack(vm.density, vm.scanisotropy);
      st.vol_b[lane] = pack(vm.scattering, 0.0f);
      flags |= kFlagVolume;
    } else {
      flags &= ~kFlagVolume;
    }
  }
  int entry_flags = 0;
  if (dest != kDestLpdf) {
    dest = finish_bounce(weight, bounce, rng, p);
    st.aux_bsdf[lane] = pack(f3{0, 0, 0}, kPhaseMain);
    if (MIS && dest == kDestExt && (flags & kFlagNoEmission)) {
      // the next loop iteration uses next_intersection instead of tracing
      st.hit_uvd[lane] = st.next_uvd[lane];
      st.hit_ids[lane] = st.next_ids[lane];
      entry_flags      = kEntryPass;
    }
  }
  st.ray_o[lane] = pack(position, bounce);
  st.ray_d[lane] = pack(incoming, opbounce);
  store_common();
  return dest | entry_flags;
}


// End of a path inside a shading-side kernel: accumulate the sample and, if the lane has more samples to do, start
// its next camera ray at once (the lane goes straight back to the extend queue - no accumulate / generate launches).
// WARP-UNIFORM. `counts`: x = samples started, y = lanes finished (per thread).
template <class PS>
YGL_D void end_of_path(const DScene& scene, const PS& st, const KParams& p, int lane, int& dest, int& entry,
    int2& counts) {
  if (dest == kDestAcc && p.fuse) {
    if (accumulate_lane(scene, st, p, lane)) {
      generate_lane(scene, st, p, lane);
      dest = kDestExt, entry = lane;
      counts.x++;
    } else {
      dest = kDestNone;
      counts.y++;
    }
  }
}
YGL_D void flush_counts(Counters* c, int2 counts) {
  for (int off = 16; off > 0; off >>= 1) {
    counts.x += __shfl_down_sync(0xffffffffu, counts.x, off);
    counts.y += __shfl_down_sync(0xffffffffu, counts.y, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (counts.x) atomicAdd(&c->camera_samples, (unsigned long long)counts.x);
    if (counts.y) atomicAdd(&c->done_lanes, counts.y);
  }
}

#ifndef YGL_SHADE_THREADS
#define YGL_SHADE_THREADS 256
#endif
#ifndef YGL_SHADE_MINBLOCKS
#define YGL_SHADE_MINBLOCKS 4  // generic kernel, measured on B200 (C3, ms per 32 spp): 256x1 361, 256x3 350, 256x4 338, 128x6 352
#endif
#ifndef YGL_SHADE_CLS_MINBLOCKS
#define YGL_SHADE_CLS_MINBLOCKS 3  // class kernels: 256 x 3 leaves 80 registers per thread (no spills, see the ptxas log)
#endif
// Common tail of the shading kernels: finish the path or pass the lane on (WARP-UNIFORM).
// FUSE (must equal p.fuse != 0): a finished path is accumulated and the lane's next camera sample generated right here;
// otherwise the lane goes to the acc queue for k_finish. A template parameter so that the kernels of a non-fused run
// do not carry the accumulate + camera code (~1.5 K instructions) at all.
template <bool FUSE>
YGL_D void route_shaded(const DScene& scene, const PathState& st, const Queues& q, const KParams& p, int parity, int lane,
    int dest, int entry, int2& counts) {
  Counters* c = q.counters;
  if (FUSE) end_of_path(scene, st, p, lane, dest, entry, counts);
  queue_push(q.ext[1 - parity], &c->n_ext[1 - parity], dest == kDestExt, entry);
  queue_push(q.lpdf, &c->n_lpdf, dest == kDestLpdf, lane);
  if (!FUSE) queue_push(q.acc, &c->n_acc, dest == kDestAcc, lane);
}

// The unspecialised shading kernel: every sampler, every kind of lane (queue of class 0).
template <int SAMPLER>
__global__ void __launch_bounds__(YGL_SHADE_THREADS, YGL_SHADE_MINBLOCKS) k_shade(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_shade[kClsGeneric];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  unsigned inst_rays = 0;
  int2     counts    = make_int2(0, 0);
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int i    = i0 + wl;
    int lane = 0, dest = kDestNone;
    int entry = 0;
    if (i < n) {
      lane  = q.shade[kClsGeneric][i] & kEntryLane;
      int r = (SAMPLER == kSamplerPathDirect || SAMPLER == kSamplerPathMis)
                  ? shade_multi<SAMPLER>(scene, st, p, lane)
                  : shade_lane<SAMPLER>(scene, st, p, lane, inst_rays);
      dest  = r & 3;
      entry = lane | (r & (kEntryShadow | kEntryPass));
    }
    if (p.fuse) route_shaded<true>(scene, st, q, p, parity, lane, dest, entry, counts);
    else route_shaded<false>(scene, st, q, p, parity, lane, dest, entry, counts);
  }
  flush_counts(c, counts);
  if (tid == 0) atomicAdd(&c->shade_calls, (unsigned long long)n);
}

// One kernel per shading class of the path sampler (trace_path): its warps hold lanes of one kind only, and it
// contains that kind's code only (k_shade<path> carried ~25 K instructions for every lane and ran 8 of 32 lanes per
// instruction; see DESIGN.md "shading classes").
template <int CLS, bool FUSE>
__global__ void __launch_bounds__(YGL_SHADE_THREADS, YGL_SHADE_CLS_MINBLOCKS) k_shade_path(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_shade[CLS];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  unsigned inst_rays = 0;
  int2     counts    = make_int2(0, 0);
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int i    = i0 + wl;
    int lane = 0, dest = kDestNone;
    if (i < n) {
      lane = q.shade[CLS][i] & kEntryLane;
      dest = shade_lane<kSamplerPath, CLS>(scene, st, p, lane, inst_rays) & 3;
    }
    route_shaded<FUSE>(scene, st, q, p, parity, lane, dest, lane, counts);
  }
  flush_counts(c, counts);
  if (tid == 0 && n) atomicAdd(&c->shade_calls, (unsigned long long)n);
}

// ---- lightpdf: sample_lights_pdf for the pending direction, then the MIS weight, weight checks,
// russian roulette and the loop increment (yocto_trace.cpp:532-536 / :573-576, :581-591) ----
#ifndef YGL_LPDF_MINBLOCKS
#define YGL_LPDF_MINBLOCKS 3
#endif
// One warp-wide step (WARP-COOPERATIVE: every lane of the warp must call it; `valid` marks lanes that hold a queued
// lane id). Outputs the lane's destination queue and its extend-queue entry.
template <bool DEEP, class PS>
YGL_D void lightpdf_lane(const DScene& scene, const PS& st, const KParams& p, bool valid, int lane,
    unsigned& rays, int& dest, int& entry) {
  float4 ro = make_float4(0, 0, 0, 0), rd = make_float4(0, 0, 1, 0);
  const bool multi = p.sampler == kSamplerPathDirect || p.sampler == kSamplerPathMis;
  int        phase = kPhaseMain;
  dest = kDestNone, entry = 0;
  if (valid) {
    if (multi) phase = __float_as_int(float4(st.aux_bsdf[lane]).w);
    const bool direct = phase == kPhaseDirectPdf || phase == kPhaseBsdfPdf;  // pdf of a direct sample at aux_o
    ro = direct ? st.aux_o[lane] : st.ray_o[lane], rd = direct ? st.aux_dir[lane] : st.ray_d[lane];
  }
  float lpdf = sample_lights_pdf<DEEP>(scene, valid, unpack3(ro), unpack3(rd), rays);
  if (valid && (phase == kPhaseDirectPdf || phase == kPhaseBsdfPdf)) {
    // direct lighting: decide whether the shadow ray is needed (yocto_trace.cpp:662-665 / :860-869)
    float4 ab      = st.aux_bsdf[lane];
    f3     bsdfcos = unpack3(ab);
    bool   trace;
    if (p.sampler == kSamplerPathDirect) {
      st.aux_dir[lane] = pack(unpack3(rd), lpdf);
      trace            = !is_zero(bsdfcos) && lpdf > 0;
    } else {
      float bsdf_pdf   = ro.w;
      float mis_weight = phase == kPhaseDirectPdf ? mis_heuristic(lpdf, bsdf_pdf) / lpdf
                                                  : mis_heuristic(bsdf_pdf, lpdf) / bsdf_pdf;
      st.aux_dir[lane] = pack(unpack3(rd), mis_weight);
      trace            = !is_zero(bsdfcos) && mis_weight != 0;
    }
    int next_phase    = phase == kPhaseDirectPdf ? (trace ? kPhaseShadowHit : kPhaseShadowSkip)
                                                 : (trace ? kPhaseBsdfHit : kPhaseBsdfSkip);
    st.aux_bsdf[lane] = pack(bsdfcos, next_phase);
    dest              = kDestExt;
    entry             = lane | (trace ? kEntryShadow : kEntryPass);
  } else if (valid) {
    float4 w4 = st.weight[lane], pd = st.pend[lane];
    f3     position = unpack3(ro), weight = unpack3(w4);
    int    bounce   = __float_as_int(ro.w);
    weight          = weight * (unpack3(pd) / (0.5f * pd.w + 0.5f * lpdf));
    rng_t rng      = load_rng(st, lane);
    dest           = finish_bounce(weight, bounce, rng, p);
    store_rng(st, lane, rng);
    st.ray_o[lane]  = pack(position, bounce);
    st.weight[lane] = pack(weight, w4.w);
    if (multi) st.aux_bsdf[lane] = pack(f3{0, 0, 0}, kPhaseMain);
    entry = lane;
  }
}
template <bool FUSE, bool DEEP>
__global__ void __launch_bounds__(256, YGL_LPDF_MINBLOCKS) k_lightpdf(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_lpdf;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  unsigned rays   = 0;
  int2     counts = make_int2(0, 0);
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int i    = i0 + wl;
    int lane = i < n ? q.lpdf[i] : 0, dest, entry;
    lightpdf_lane<DEEP>(scene, st, p, i < n, lane, rays, dest, entry);
    if (FUSE) end_of_path(scene, st, p, lane, dest, entry, counts);
    queue_push(q.ext[1 - parity], &c->n_ext[1 - parity], dest == kDestExt, entry);
    if (!FUSE) queue_push(q.acc, &c->n_acc, dest == kDestAcc, lane);
  }
  flush_counts(c, counts);
  // one atomic per warp for the instance-ray count
  for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, off);
  if (wl == 0 && rays) atomicAdd(&c->instance_rays, (unsigned long long)rays);
}

// ---- accumulate: trace_sample tail, yocto_trace.cpp:1469-1491. Returns whether the lane has more samples to do. ----
template <class PS>
YGL_D bool accumulate_lane(const DScene& scene, const PS& st, const KParams& p, int lane) {
  float4 rad4   = st.radiance[lane];
  f3     radiance = unpack3(rad4);
  bool   hit    = (__float_as_int(rad4.w) & kFlagHit) != 0;
  int    sample = st.sample[lane];
  if (!vfinite(radiance)) radiance = {0, 0, 0};
  if (max3(radiance) > p.clamp) radiance = radiance * (p.clamp / max3(radiance));
  float  w   = 1.0f / (sample + 1);
  float4 im4 = st.image[lane];
  f4     image  = {im4.x, im4.y, im4.z, im4.w};
  f3     albedo = {st.albedo[3 * lane + 0], st.albedo[3 * lane + 1], st.albedo[3 * lane + 2]};
  f3     normal = {st.normal[3 * lane + 0], st.normal[3 * lane + 1], st.normal[3 * lane + 2]};
  f3     n0     = unpack3(st.normal0[lane]);
  if (hit) {
    image  = lerp4(image, f4{radiance.x, radiance.y, radiance.z, 1}, w);
    albedo = lerp3(albedo, unpack3(st.albedo0[lane]), w);
    normal = lerp3(normal, n0, w);
    st.hits[lane] += 1;
  } else if (!p.envhidden && scene.num_environments > 0) {
    image  = lerp4(image, f4{radiance.x, radiance.y, radiance.z, 1}, w);
    albedo = lerp3(albedo, f3{1, 1, 1}, w);
    normal = lerp3(normal, n0, w);
    st.hits[lane] += 1;
  } else {
    image  = lerp4(image, f4{0, 0, 0, 0}, w);
    albedo = lerp3(albedo, f3{0, 0, 0}, w);
    normal = lerp3(normal, n0, w);
  }
  st.image[lane]          = make_float4(image.x, image.y, image.z, image.w);
  st.albedo[3 * lane + 0] = albedo.x, st.albedo[3 * lane + 1] = albedo.y, st.albedo[3 * lane + 2] = albedo.z;
  st.normal[3 * lane + 0] = normal.x, st.normal[3 * lane + 1] = normal.y, st.normal[3 * lane + 2] = normal.z;
  sample += 1;
  st.sample[lane] = sample;
  return sample < p.sample_end;
}
// Class kClsMiss of the path sampler: every lane of the warp ends its path on the environment, so the kernel
// accumulates the sample and starts the lane's next camera ray at once, whatever p.fuse says (the code is coherent,
// and it saves the lane a trip through k_finish).
__global__ void __launch_bounds__(256, 4) k_shade_miss(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_shade[kClsMiss];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  int2      counts = make_int2(0, 0);
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int  i    = i0 + wl;
    int  lane = 0;
    bool more = false;
    if (i < n) {
      lane = q.shade[kClsMiss][i] & kEntryLane;
      miss_lane(scene, st, p, lane);
      more = accumulate_lane(scene, st, p, lane);
      if (more) {
        generate_lane(scene, st, p, lane);
        counts.x++;
      } else {
        counts.y++;
      }
    }
    queue_push(q.ext[1 - parity], &c->n_ext[1 - parity], more, lane);
  }
  flush_counts(c, counts);
  if (tid == 0 && n) atomicAdd(&c->shade_calls, (unsigned long long)n);
}

// The finished paths of an iteration (p.fuse == 0): accumulate, then the lane's next camera sample goes straight into
// the next extend queue.
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ DScene scene, const __grid_constant__ PathState st,
    const __grid_constant__ Queues q, const __grid_constant__ KParams p, int parity) {
  Counters* c   = q.counters;
  const int n   = c->n_acc;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  int2      counts = make_int2(0, 0);
  for (int i0 = tid - wl; i0 < n; i0 += stride) {
    int  i    = i0 + wl;
    int  lane = 0;
    bool more = false;
    if (i < n) {
      lane = q.acc[i];
      more = accumulate_lane(scene, st, p, lane);
      if (more) {
        generate_lane(scene, st, p, lane);
        counts.x++;
      } else {
        counts.y++;
      }
    }
    queue_push(q.ext[1 - parity], &c->n_ext[1 - parity], more, lane);
  }
  flush_counts(c, counts);
}


// ---- batch form of intersect_scene_bvh / intersect_instance_bvh ----
template <bool COUNT>
__global__ void __launch_bounds__(256) k_intersect_rays(DScene scene, const float4* __restrict__ rays, long long n,
    int instance, int find_any, int2* __restrict__ out, unsigned long long* counters) {
  long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const int wl  = threadIdx.x & 31;
  trav_counters tc = {0, 0, 0, 0};
  unsigned      nhits = 0;
  for (long long i0 = tid - wl; i0 < n; i0 += stride) {
    const long long i     = i0 + wl;
    const bool      valid = i < n;
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(1, 0, 0, 0);
    if (valid) a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
    f3    o = {a.x, a.y, a.z}, d = {a.w, b.x, b.y};
    hit_t h = find_any ? trace_ray<true, COUNT>(scene, valid, o, d, b.z, b.w, instance, tc)
                       : trace_ray<false, COUNT>(scene, valid, o, d, b.z, b.w, instance, tc);
    if (valid) {
      out[3 * i + 0] = make_int2(h.instance, h.element);
      out[3 * i + 1] = make_int2(__float_as_int(h.uv.x), __float_as_int(h.uv.y));
      out[3 * i + 2] = make_int2(__float_as_int(h.distance), h.hit ? 1 : 0);
      nhits += h.hit ? 1 : 0;
    }
  }
  if (COUNT) {
    atomicAdd(counters + 0, (unsigned long long)tc.top_nodes);
    atomicAdd(counters + 1, (unsigned long long)tc.bot_nodes);
    atomicAdd(counters + 2, (unsigned long long)tc.instances);
    atomicAdd(counters + 3, (unsigned long long)tc.prims);
    atomicAdd(counters + 4, (unsigned long long)nhits);
  }
}

// ---- test hook: the device libm (ygl_math.cuh) on arrays, compared against the host libm by the tests ----
__global__ void k_debug_libm(int fn, const float* __restrict__ x, const float* __restrict__ y, long long n,
    float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = x[i], b = y ? y[i] : 0.0f, r = 0;
    switch (fn) {
      case 0: r = ysin(a); break;
      case 1: r = ycos(a); break;
      case 2: r = yexp(a); break;
      case 3: r = ylog(a); break;
      case 4: r = yatan(a); break;
      case 5: r = yacos(a); break;
      case 6: r = yatan2(a, b); break;
      case 7: r = ypow(a, b); break;
      case 8: r = ysqrt(a); break;
      case 9: r = yfmod(a, b); break;
    }
    out[i] = r;
  }
}
// ---- tonemap_image, yocto_image.cpp:911-922 (tonemap: yocto_color.h:356-366; tonemap_filmic :287-296; rgb_to_srgb
// :239-242; float_to_byte :218-222). exp2(exposure) is evaluated once, on the host, by the same libm the reference
// calls; the per-pixel arithmetic (ACES fit, one powf per channel) runs here, bit for bit. ----
YGL_D float tonemap_channel(float v, float scale, bool scaled, bool filmic, bool srgb) {
  if (scaled) v = v * scale;
  if (filmic) {
    float h = v * 0.6f;
    float l = (h * h * 2.51f + h * 0.03f) / (h * h * 2.43f + h * 0.59f + 0.14f);
    v       = (0.0f > l) ? 0.0f : l;  // max({0,0,0}, ldr): yocto's max keeps a NaN
  }
  if (srgb) v = (v <= 0.0031308f) ? 12.92f * v : (1 + 0.055f) * ypow(v, 1 / 2.4f) - 0.055f;
  return v;
}
// int(a * 256) as the reference's x86-64 build computes it (cvttss2si: NaN and out-of-range give INT_MIN), then clamp
YGL_D unsigned char tonemap_byte(float a) {
  float v = a * 256;
  int   i = (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000;
  return (unsigned char)(i < 0 ? 0 : i > 255 ? 255 : i);
}
__global__ void __launch_bounds__(256) k_tonemap(const float4* __restrict__ hdr, long long n, float scale, int scaled, int filmic,
    int srgb, float4* __restrict__ ldr, uchar4* __restrict__ ldr_bytes) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 p = hdr[i];
    float4 o = make_float4(tonemap_channel(p.x, scale, scaled, filmic, srgb), tonemap_channel(p.y, scale, scaled, filmic, srgb),
        tonemap_channel(p.z, scale, scaled, filmic, srgb), p.w);
    if (ldr) ldr[i] = o;
    if (ldr_bytes) ldr_bytes[i] = make_uchar4(tonemap_byte(o.x), tonemap_byte(o.y), tonemap_byte(o.z), tonemap_byte(o.w));
  }
}
void launch_tonemap(cudaStream_t s, int num_sms, const float4* hdr, long long n, float scale, bool scaled, bool filmic, bool srgb,
    float4* ldr, uchar4* ldr_bytes) {
  if (n <= 0) return;
  const long long blocks = std::min<long long>((n + 255) / 256, (long long)num_sms * 8);
  k_tonemap<<<(unsigned)blocks, 256, 0, s>>>(hdr, n, scale, scaled ? 1 : 0, filmic ? 1 : 0, srgb ? 1 : 0, ldr, ldr_bytes);
}
void launch_debug_libm(cudaStream_t s, int fn, const float* x, const float* y, long long n, float* out) {
  k_debug_libm<<<1184, 256, 0, s>>>(fn, x, y, n, out);
}

// ------------------------------------------------------------------------------------------
void launch_begin_iteration(cudaStream_t s, Queues q, int parity) { k_begin_iteration<<<1, 1, 0, s>>>(q.counters, parity); }
void launch_seed_lanes(cudaStream_t s, LaunchCfg cfg, PathState st, Queues q, int parity, int sample_begin,
    int lane_lo, int lane_hi) {
  k_seed_lanes<<<cfg.blocks, cfg.threads, 0, s>>>(st, q, parity, sample_begin, lane_lo, lane_hi);
}
void launch_generate(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity) {
  k_generate<<<cfg.blocks, cfg.threads, 0, s>>>(scene, st, q, p, parity);
}
static int extend_blocks_per_sm() {
  static int per_sm = 0;
  if (!per_sm) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_extend<false, 0, kStackShallow>), 128, 0);
    if (per_sm < 1) per_sm = 1;
  }
  return per_sm;
}
int extend_grid_threads(int num_sms) { return num_sms * extend_blocks_per_sm() * 128; }
void launch_extend(cudaStream_t s, int num_sms, const Tuning& tune, DScene scene, PathState st, Queues q, int parity,
    unsigned long long* trav) {
  // persistent kernel: exactly the resident capacity (SMs x blocks/SM from the occupancy API), capped
  // by the lanes that can exist
  const int per_sm = extend_blocks_per_sm();
  int blocks = num_sms * (tune.ext_blocks_per_sm > 0 ? std::min(per_sm, tune.ext_blocks_per_sm) : per_sm);
  int needed = (st.num_lanes + 127) / 128;
  if (blocks > needed) blocks = needed;
  if (blocks < 1) blocks = 1;
  // Tail knobs (the queue is exhausted, the warp drains). Two alternatives, chosen by the size of the tile
  // (measured on B200, C3): on a full 1080p frame (2.07 M lanes, ~9 rays per slot and launch) parking the last
  // <= 8 busy lanes of a warp for the next launch once it has walked 96 more rounds wins 2 %; on small tiles
  // (one ray per slot: the launch time IS the tail) parking only adds iterations, and finishing the last
  // <= 12 lanes in a vote-free per-lane loop wins 8 % (1/8 tile: 78.8 -> 72.4 ms per 32 spp).
  const bool big_tile = st.num_lanes > 800000;
  const int  sb       = tune.suspend >= 0 ? tune.suspend : (big_tile ? kSuspendBelow : 0);
  const int  suspend  = sb > 0 ? (std::min(sb, 31) | std::max(1, tune.suspend_rounds) << 8) : 0;
  const int  lone     = tune.lone >= 0 ? tune.lone : (big_tile ? 0 : 12);
  // node visits per round and the two vote weights travel in one kernel argument
  const int reps_and_weights = (std::max(1, tune.node_reps) & 0xff) | (std::min(255, std::max(1, tune.prim_weight)) << 8) |
                               (std::min(255, std::max(1, tune.enter_weight)) << 16);
  // the stack variant the bound trees need (DScene::stack_mode): shared memory only, + a small local array, or the
  // full 128 levels per tree
  using kernel_t = void (*)(const DScene, const PathState, const Queues, int, unsigned long long*, int, int, int, int, int);
  static const kernel_t kernels[3][2][2] = {
      {{k_extend<false, 0, kStackShared>, k_extend<true, 0, kStackShared>},
          {k_extend<false, 1, kStackShared>, k_extend<true, 1, kStackShared>}},
      {{k_extend<false, 0, kStackShallow>, k_extend<true, 0, kStackShallow>},
          {k_extend<false, 1, kStackShallow>, k_extend<true, 1, kStackShallow>}},
      {{k_extend<false, 0, kStackDeep>, k_extend<true, 0, kStackDeep>},
          {k_extend<false, 1, kStackDeep>, k_extend<true, 1, kStackDeep>}}};
  kernel_t kernel = kernels[scene.stack_mode][lone > 0 ? 1 : 0][trav ? 1 : 0];
#ifndef YGL_PAIR_VISIT
  // A/B (tune.top_smem = 1): the instance-level tree staged in shared memory by a bulk asynchronous copy; only the
  // variant the headline workload runs (timed, parked tail, shallow stack) has this instantiation
  const size_t top_bytes = (size_t)scene.top_num_nodes * 32;
  if (tune.top_smem > 0 && !trav && lone == 0 && scene.stack_mode == kStackShared && top_bytes > 0 && top_bytes <= 24 * 1024) {
    auto* top_kernel = k_extend<false, 0, kStackShared, true>;
    int   top_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&top_per_sm, top_kernel, 128, top_bytes);
    int top_blocks = num_sms * std::max(1, tune.ext_blocks_per_sm > 0 ? std::min(top_per_sm, tune.ext_blocks_per_sm) : top_per_sm);
    top_blocks     = std::max(1, std::min(top_blocks, needed));
    top_kernel<<<top_blocks, 128, top_bytes, s>>>(scene, st, q, parity, trav, tune.refill, reps_and_weights, suspend, lone,
        tune.lone_steps);
    return;
  }
#endif
  // how the SM's 256 KB split between L1 and shared memory is a per-kernel, per-device preference
  if (tune.carveout >= 0)
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, tune.carveout));
  kernel<<<blocks, 128, 0, s>>>(scene, st, q, parity, trav, tune.refill, reps_and_weights, suspend, lone, tune.lone_steps);
}

template <int CLS>
static void launch_shade_class(cudaStream_t s, int blocks, DScene scene, PathState st, Queues q, KParams p, int parity,
    unsigned class_mask) {
  if (!(class_mask >> CLS & 1)) return;
  if (p.fuse) k_shade_path<CLS, true><<<blocks, YGL_SHADE_THREADS, 0, s>>>(scene, st, q, p, parity);
  else k_shade_path<CLS, false><<<blocks, YGL_SHADE_THREADS, 0, s>>>(scene, st, q, p, parity);
}
void launch_shade(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity,
    unsigned class_mask) {
  const int threads = YGL_SHADE_THREADS;
  const int blocks  = std::max(1, cfg.blocks * cfg.threads / threads);
  if (p.sampler == kSamplerPath && scene.inst_class) {
    // binned queues: misses first (each of them starts a new camera ray for the next extend), then one kernel per
    // material type present in the scene, then the lanes inside media
    if (class_mask >> kClsMiss & 1) k_shade_miss<<<blocks, threads, 0, s>>>(scene, st, q, p, parity);
    launch_shade_class<1>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<2>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<3>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<4>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<5>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<6>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<7>(s, blocks, scene, st, q, p, parity, class_mask);
    launch_shade_class<8>(s, blocks, scene, st, q, p, parity, class_mask);
    if (class_mask & 1u) k_shade<kSamplerPath><<<blocks, threads, 0, s>>>(scene, st, q, p, parity);
    return;
  }
  switch (p.sampler) {
    case kSamplerFalsecolor: k_shade<kSamplerFalsecolor><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerEyelight: k_shade<kSamplerEyelight><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerNaive: k_shade<kSamplerNaive><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerFurnace: k_shade<kSamplerFurnace><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerPathTest: k_shade<kSamplerPathTest><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerDiagram: k_shade<kSamplerDiagram><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerPathDirect: k_shade<kSamplerPathDirect><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    case kSamplerPathMis: k_shade<kSamplerPathMis><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
    default: k_shade<kSamplerPath><<<blocks, threads, 0, s>>>(scene, st, q, p, parity); break;
  }
}
void launch_lightpdf(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity) {
  auto* kernel = scene.stack_mode == kStackDeep ? (p.fuse ? k_lightpdf<true, true> : k_lightpdf<false, true>)
                                  : (p.fuse ? k_lightpdf<true, false> : k_lightpdf<false, false>);
  kernel<<<cfg.blocks, cfg.threads, 0, s>>>(scene, st, q, p, parity);
}
void launch_finish(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity) {
  k_finish<<<cfg.blocks, cfg.threads, 0, s>>>(scene, st, q, p, parity);
}
void launch_intersect_rays(cudaStream_t s, LaunchCfg cfg, DScene scene, const float4* rays, long long n, int instance,
    int find_any, void* out, unsigned long long* counters) {
  if (counters)
    k_intersect_rays<true><<<cfg.blocks, cfg.threads, 0, s>>>(scene, rays, n, instance, find_any, (int2*)out, counters);
  else
    k_intersect_rays<false><<<cfg.blocks, cfg.threads, 0, s>>>(scene, rays, n, instance, find_any, (int2*)out, nullptr);
}

}  // namespace ygl
