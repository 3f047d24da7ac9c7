// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C ABI over the UNMODIFIED reference CPU renderer (Yocto/GL, sources
// compiled where they lie under /root/reference by oracle/Makefile; outputs go
// to oracle/_ref/ only). It lets tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg run the reference's own
// make_scene_bvh / make_trace_lights / make_trace_state / trace_samples /
// intersect_scene_bvh on the same flat scene views (include/ygl_b200.h) that
// the CUDA path consumes. Nothing under yocto-gl_b200/ may link or load this.
//
// Reference entry points used (libs/yocto/): yocto_trace.h:116-190,
// yocto_bvh.h:83-112, yocto_scene.h:83-213, yocto_scene.cpp:970 (make_cornellbox).

#include <yocto/yocto_bvh.h>
#include <yocto/yocto_geometry.h>
#include <yocto/yocto_scene.h>
#include <yocto/yocto_sceneio.h>
#include <yocto/yocto_shape.h>
#include <yocto/yocto_image.h>
#include <yocto/yocto_trace.h>

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/ygl_b200.h"
#ifdef REF_COUNTERS
#include "ref_counters.h"
namespace refcount {
std::atomic<uint64_t>     g_totals[kNumModes * kPerMode];
thread_local tls_counters t_counters;
}  // namespace refcount
#endif

#ifdef REF_DOUBLE_LIBM
// Variant used only to localise arithmetic differences: every float libm call of the reference
// is evaluated in double and rounded once (what the CUDA path does on device). Bound with
// -Wl,-Bsymbolic-functions so that only this .so sees these definitions.
#include <math.h>
extern "C" {
float sinf(float x) { return (float)sin((double)x); }
float cosf(float x) { return (float)cos((double)x); }
float tanf(float x) { return (float)tan((double)x); }
float asinf(float x) { return (float)asin((double)x); }
float acosf(float x) { return (float)acos((double)x); }
float atanf(float x) { return (float)atan((double)x); }
float atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
float expf(float x) { return (float)exp((double)x); }
float logf(float x) { return (float)log((double)x); }
float powf(float x, float y) { return (float)pow((double)x, (double)y); }
void  sincosf(float x, float* s, float* c) {
  *s = (float)sin((double)x);
  *c = (float)cos((double)x);
}
}
#endif

using namespace yocto;

namespace {

struct ref_scene {
  scene_data             scene;
  std::vector<ygl_shape> shape_views;  // storage for ref_scene_describe
  std::vector<ygl_texture> texture_views;
  std::vector<ygl_camera> cameras;
  std::vector<ygl_instance> instances;
  std::vector<ygl_environment> environments;
  std::vector<ygl_material> materials;
};

frame3f to_frame(const ygl_frame3f& f) {
  return frame3f{{f.x[0], f.x[1], f.x[2]}, {f.y[0], f.y[1], f.y[2]}, {f.z[0], f.z[1], f.z[2]},
      {f.o[0], f.o[1], f.o[2]}};
}
ygl_frame3f from_frame(const frame3f& f) {
  return ygl_frame3f{{f.x.x, f.x.y, f.x.z}, {f.y.x, f.y.y, f.y.z}, {f.z.x, f.z.y, f.z.z},
      {f.o.x, f.o.y, f.o.z}};
}

trace_params to_params(const ygl_trace_params& p) {
  auto params           = trace_params{};
  params.camera         = p.camera;
  params.resolution     = p.resolution;
  params.sampler        = (trace_sampler_type)p.sampler;
  params.falsecolor     = (trace_falsecolor_type)p.falsecolor;
  params.samples        = p.samples;
  params.bounces        = p.bounces;
  params.clamp          = p.clamp;
  params.nocaustics     = p.nocaustics != 0;
  params.envhidden      = p.envhidden != 0;
  params.tentfilter     = p.tentfilter != 0;
  params.seed           = p.seed;
  params.embreebvh      = false;
  params.highqualitybvh = p.highqualitybvh != 0;
  params.noparallel     = p.noparallel != 0;
  params.pratio         = p.pratio;
  params.denoise        = false;
  params.batch          = p.batch;
  return params;
}

}  // namespace

extern "C" {

int ref_sizeof(const char* name) {
  auto n = std::string{name};
  if (n == "bvh_node") return (int)sizeof(bvh_node);
  if (n == "ray3f") return (int)sizeof(ray3f);
  if (n == "frame3f") return (int)sizeof(frame3f);
  if (n == "instance_data") return (int)sizeof(instance_data);
  if (n == "material_data") return (int)sizeof(material_data);
  if (n == "camera_data") return (int)sizeof(camera_data);
  if (n == "environment_data") return (int)sizeof(environment_data);
  if (n == "rng_state") return (int)sizeof(rng_state);
  if (n == "scene_intersection") return (int)sizeof(scene_intersection);
  if (n == "trace_params") return (int)sizeof(trace_params);
  return -1;
}

void* ref_scene_create(const ygl_scene_desc* desc) {
  auto  rs    = new ref_scene{};
  auto& scene = rs->scene;
  for (int i = 0; i < desc->num_cameras; i++) {
    auto& c        = desc->cameras[i];
    auto& camera   = scene.cameras.emplace_back();
    camera.frame   = to_frame(c.frame);
    camera.orthographic = c.orthographic != 0;
    camera.lens    = c.lens;
    camera.film    = c.film;
    camera.aspect  = c.aspect;
    camera.focus   = c.focus;
    camera.aperture = c.aperture;
  }
  for (int i = 0; i < desc->num_instances; i++) {
    auto& instance    = scene.instances.emplace_back();
    instance.frame    = to_frame(desc->instances[i].frame);
    instance.shape    = desc->instances[i].shape;
    instance.material = desc->instances[i].material;
  }
  for (int i = 0; i < desc->num_environments; i++) {
    auto& e      = desc->environments[i];
    auto& env    = scene.environments.emplace_back();
    env.frame    = to_frame(e.frame);
    env.emission = {e.emission[0], e.emission[1], e.emission[2]};
    env.emission_tex = e.emission_tex;
  }
  for (int i = 0; i < desc->num_materials; i++) {
    auto& m        = desc->materials[i];
    auto& material = scene.materials.emplace_back();
    material.type  = (material_type)m.type;
    material.emission   = {m.emission[0], m.emission[1], m.emission[2]};
    material.color      = {m.color[0], m.color[1], m.color[2]};
    material.roughness  = m.roughness;
    material.metallic   = m.metallic;
    material.ior        = m.ior;
    material.scattering = {m.scattering[0], m.scattering[1], m.scattering[2]};
    material.scanisotropy = m.scanisotropy;
    material.trdepth    = m.trdepth;
    material.opacity    = m.opacity;
    material.emission_tex   = m.emission_tex;
    material.color_tex      = m.color_tex;
    material.roughness_tex  = m.roughness_tex;
    material.scattering_tex = m.scattering_tex;
    material.normal_tex     = m.normal_tex;
  }
  for (int i = 0; i < desc->num_textures; i++) {
    auto& t       = desc->textures[i];
    auto& texture = scene.textures.emplace_back();
    texture.width = t.width;
    texture.height = t.height;
    texture.linear = t.linear != 0;
    texture.nearest = t.nearest != 0;
    texture.clamp = t.clamp != 0;
    auto n = (size_t)t.width * t.height;
    if (t.pixelsf) {
      texture.pixelsf.resize(n);
      memcpy(texture.pixelsf.data(), t.pixelsf, n * sizeof(vec4f));
    } else if (t.pixelsb) {
      texture.pixelsb.resize(n);
      memcpy(texture.pixelsb.data(), t.pixelsb, n * sizeof(vec4b));
    }
  }
  for (int i = 0; i < desc->num_shapes; i++) {
    auto& s     = desc->shapes[i];
    auto& shape = scene.shapes.emplace_back();
    auto copy = [](auto& dst, const void* src, size_t n) {
      dst.resize(n);
      if (n) memcpy(dst.data(), src, n * sizeof(dst[0]));
    };
    copy(shape.points, s.points, s.num_points);
    copy(shape.lines, s.lines, s.num_lines);
    copy(shape.triangles, s.triangles, s.num_triangles);
    copy(shape.quads, s.quads, s.num_quads);
    copy(shape.positions, s.positions, s.num_positions);
    copy(shape.normals, s.normals, s.num_normals);
    copy(shape.texcoords, s.texcoords, s.num_texcoords);
    copy(shape.colors, s.colors, s.num_colors);
    copy(shape.radius, s.radius, s.num_radius);
  }
  return rs;
}

// The reference's procedural Cornell box (yocto_scene.cpp:970) as a scene handle.
void* ref_scene_cornellbox() {
  auto rs   = new ref_scene{};
  rs->scene = make_cornellbox();
  return rs;
}

// The reference's own loader (load_scene, yocto_sceneio.cpp:2761): JSON + PLY + textures. Returns null on error
// (message in ref_last_error). Subdivs are tesselated like ytrace does (apps/ytrace.cpp:110-113).
static std::string g_ref_error;
const char* ref_last_error() { return g_ref_error.c_str(); }
void* ref_scene_load(const char* filename) {
  auto rs    = new ref_scene{};
  auto error = std::string{};
  if (!load_scene(filename, rs->scene, error)) {
    g_ref_error = error;
    delete rs;
    return nullptr;
  }
  if (!rs->scene.subdivs.empty()) tesselate_subdivs(rs->scene);
  return rs;
}

void ref_scene_destroy(void* scene) { delete (ref_scene*)scene; }

// Flat views of a reference scene (pointers stay valid until ref_scene_destroy).
void ref_scene_describe(void* scene_, ygl_scene_desc* desc) {
  auto  rs    = (ref_scene*)scene_;
  auto& scene = rs->scene;
  rs->cameras.clear();
  for (auto& camera : scene.cameras) {
    rs->cameras.push_back({from_frame(camera.frame), camera.orthographic ? 1 : 0, camera.lens,
        camera.film, camera.aspect, camera.focus, camera.aperture});
  }
  rs->instances.clear();
  for (auto& instance : scene.instances)
    rs->instances.push_back({from_frame(instance.frame), instance.shape, instance.material});
  rs->environments.clear();
  for (auto& env : scene.environments)
    rs->environments.push_back({from_frame(env.frame),
        {env.emission.x, env.emission.y, env.emission.z}, env.emission_tex});
  rs->materials.clear();
  for (auto& m : scene.materials) {
    rs->materials.push_back({(int)m.type, {m.emission.x, m.emission.y, m.emission.z},
        {m.color.x, m.color.y, m.color.z}, m.roughness, m.metallic, m.ior,
        {m.scattering.x, m.scattering.y, m.scattering.z}, m.scanisotropy, m.trdepth, m.opacity,
        m.emission_tex, m.color_tex, m.roughness_tex, m.scattering_tex, m.normal_tex});
  }
  rs->texture_views.clear();
  for (auto& t : scene.textures) {
    rs->texture_views.push_back({t.width, t.height, t.linear ? 1 : 0, t.nearest ? 1 : 0,
        t.clamp ? 1 : 0, t.pixelsf.empty() ? nullptr : (const float*)t.pixelsf.data(),
        t.pixelsb.empty() ? nullptr : (const uint8_t*)t.pixelsb.data()});
  }
  rs->shape_views.clear();
  for (auto& s : scene.shapes) {
    rs->shape_views.push_back({(int)s.points.size(), (int)s.lines.size(),
        (int)s.triangles.size(), (int)s.quads.size(), (const int32_t*)s.points.data(),
        (const int32_t*)s.lines.data(), (const int32_t*)s.triangles.data(),
        (const int32_t*)s.quads.data(), (int)s.positions.size(), (int)s.normals.size(),
        (int)s.texcoords.size(), (int)s.colors.size(), (int)s.radius.size(),
        (const float*)s.positions.data(), (const float*)s.normals.data(),
        (const float*)s.texcoords.data(), (const float*)s.colors.data(),
        (const float*)s.radius.data()});
  }
  desc->num_cameras      = (int)rs->cameras.size();
  desc->num_instances    = (int)rs->instances.size();
  desc->num_environments = (int)rs->environments.size();
  desc->num_shapes       = (int)rs->shape_views.size();
  desc->num_textures     = (int)rs->texture_views.size();
  desc->num_materials    = (int)rs->materials.size();
  desc->cameras          = rs->cameras.data();
  desc->instances        = rs->instances.data();
  desc->environments     = rs->environments.data();
  desc->shapes           = rs->shape_views.data();
  desc->textures         = rs->texture_views.data();
  desc->materials        = rs->materials.data();
}

// ---- BVH (make_scene_bvh, yocto_bvh.cpp:364) ----
void* ref_bvh_build(void* scene, int highquality) {
  auto bvh = new scene_bvh{};
  *bvh     = make_scene_bvh(((ref_scene*)scene)->scene, highquality != 0, false);
  return bvh;
}
// update_scene_bvh, yocto_bvh.cpp:434: refit `bvh` to the (edited) scene
void ref_bvh_update(void* bvh, void* scene, const int* shapes, int num_shapes) {
  update_scene_bvh(*(scene_bvh*)bvh, ((ref_scene*)scene)->scene, {}, vector<int>(shapes, shapes + num_shapes));
}
void ref_bvh_destroy(void* bvh) { delete (scene_bvh*)bvh; }
static const bvh_tree& pick_tree(void* bvh_, int shape) {
  auto bvh = (scene_bvh*)bvh_;
  return shape < 0 ? bvh->bvh : bvh->shapes[shape].bvh;
}
void ref_bvh_tree_size(void* bvh, int shape, int* num_nodes, int* num_primitives) {
  auto& tree      = pick_tree(bvh, shape);
  *num_nodes      = (int)tree.nodes.size();
  *num_primitives = (int)tree.primitives.size();
}
void ref_bvh_tree_get(void* bvh, int shape, ygl_bvh_node* nodes, int32_t* primitives) {
  auto& tree = pick_tree(bvh, shape);
  static_assert(sizeof(bvh_node) == sizeof(ygl_bvh_node));
  for (size_t i = 0; i < tree.nodes.size(); i++) {
    auto& n = tree.nodes[i];
    nodes[i] = {{n.bbox.min.x, n.bbox.min.y, n.bbox.min.z}, {n.bbox.max.x, n.bbox.max.y, n.bbox.max.z},
        n.start, n.num, n.axis, (uint8_t)(n.internal ? 1 : 0)};
  }
  memcpy(primitives, tree.primitives.data(), tree.primitives.size() * sizeof(int));
}

// ---- intersect_scene_bvh / intersect_instance_bvh (yocto_bvh.cpp:554,619), batch form ----
void ref_intersect_rays(void* scene_, void* bvh_, const ygl_ray* rays, int64_t n, int instance,
    int find_any, ygl_intersection* out, int nthreads) {
  auto& scene = ((ref_scene*)scene_)->scene;
  auto& bvh   = *(scene_bvh*)bvh_;
  auto  work  = [&](int64_t begin, int64_t end) {
    for (auto i = begin; i < end; i++) {
      auto& r   = rays[i];
      auto  ray = ray3f{{r.o[0], r.o[1], r.o[2]}, {r.d[0], r.d[1], r.d[2]}, r.tmin, r.tmax};
      auto  isec = instance < 0 ? intersect_scene_bvh(bvh, scene, ray, find_any != 0)
                                : intersect_instance_bvh(bvh, scene, instance, ray, find_any != 0);
      out[i] = {isec.instance, isec.element, {isec.uv.x, isec.uv.y}, isec.distance,
          isec.hit ? 1 : 0};
    }
  };
  if (nthreads <= 1) {
    work(0, n);
  } else {
    auto threads = std::vector<std::thread>{};
    for (int t = 0; t < nthreads; t++)
      threads.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    for (auto& t : threads) t.join();
  }
}

// ---- make_trace_lights (yocto_trace.cpp:1528) ----
void* ref_lights_create(void* scene, const ygl_trace_params* params) {
  auto lights = new trace_lights{};
  *lights     = make_trace_lights(((ref_scene*)scene)->scene, to_params(*params));
  return lights;
}
void ref_lights_destroy(void* lights) { delete (trace_lights*)lights; }
int  ref_lights_count(void* lights) { return (int)((trace_lights*)lights)->lights.size(); }
void ref_lights_get(void* lights, int i, int* instance, int* environment, int* cdf_size, float* cdf) {
  auto& light  = ((trace_lights*)lights)->lights[i];
  *instance    = light.instance;
  *environment = light.environment;
  *cdf_size    = (int)light.elements_cdf.size();
  if (cdf) memcpy(cdf, light.elements_cdf.data(), light.elements_cdf.size() * sizeof(float));
}

// ---- make_trace_state (yocto_trace.cpp:1495): size + per-pixel rng table ----
void ref_state_rngs(void* scene, const ygl_trace_params* params, int* width, int* height,
    uint64_t* rngs) {
  auto state = make_trace_state(((ref_scene*)scene)->scene, to_params(*params));
  *width     = state.width;
  *height    = state.height;
  if (rngs) memcpy(rngs, state.rngs.data(), state.rngs.size() * sizeof(rng_state));
}

// ---- the render itself: make_trace_bvh + lights + state + trace_samples loop ----
// Returns the seconds spent in the trace_samples loop only (the span ytrace times,
// apps/ytrace.cpp:141-154). Any output pointer may be null. `samples_override` > 0 renders
// that many samples instead of params.samples (bounded CPU-baseline sample).
double ref_trace_image(void* scene_, const ygl_trace_params* params_, int samples_override,
    int* width, int* height, float* image, float* albedo, float* normal, int32_t* hits,
    uint64_t* rngs) {
  auto& scene  = ((ref_scene*)scene_)->scene;
  auto  params = to_params(*params_);
  if (samples_override > 0) params.samples = samples_override;
  auto bvh    = make_trace_bvh(scene, params);
  auto lights = make_trace_lights(scene, params);
  auto state  = make_trace_state(scene, params);
  auto start  = std::chrono::steady_clock::now();
  while (state.samples < params.samples) trace_samples(state, scene, bvh, lights, params);
  auto elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  if (width) *width = state.width;
  if (height) *height = state.height;
  auto n = (size_t)state.width * state.height;
  if (image) memcpy(image, state.image.data(), n * sizeof(vec4f));
  if (albedo) memcpy(albedo, state.albedo.data(), n * sizeof(vec3f));
  if (normal) memcpy(normal, state.normal.data(), n * sizeof(vec3f));
  if (hits) memcpy(hits, state.hits.data(), n * sizeof(int));
  if (rngs) memcpy(rngs, state.rngs.data(), n * sizeof(rng_state));
  return elapsed;
}

// tonemap_image, yocto_image.cpp:911-922 (both outputs)
void ref_tonemap_image(const float* hdr, int64_t n, float exposure, int filmic, int srgb, float* ldr, uint8_t* ldr_bytes) {
  auto in = vector<vec4f>((const vec4f*)hdr, (const vec4f*)hdr + n);
  auto f  = vector<vec4f>(n);
  auto b  = vector<vec4b>(n);
  tonemap_image(f, in, exposure, filmic != 0, srgb != 0);
  tonemap_image(b, in, exposure, filmic != 0, srgb != 0);
  if (ldr) memcpy(ldr, f.data(), n * sizeof(vec4f));
  if (ldr_bytes) memcpy(ldr_bytes, b.data(), n * sizeof(vec4b));
}

int ref_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

// ---- instrumented oracle (libyocto_ref_count.so only): traversal counters of SURVEY.md §8d ----
// out[mode * 8 + k], mode 0 = intersect_scene_bvh queries, 1 = intersect_instance_bvh queries;
// k = top nodes, bottom nodes, instance visits, point / line / triangle / quad tests, rays (ref_counters.h).
#ifdef REF_COUNTERS
int  ref_counters_available() { return 1; }
void ref_counters_reset() {
  refcount::t_counters.flush();
  for (auto& total : refcount::g_totals) total.store(0);
}
void ref_counters_read(uint64_t* out) {
  refcount::t_counters.flush();
  for (int k = 0; k < refcount::kNumModes * refcount::kPerMode; k++) out[k] = refcount::g_totals[k].load();
}
#else
int  ref_counters_available() { return 0; }
void ref_counters_reset() {}
void ref_counters_read(uint64_t* out) { memset(out, 0, 16 * sizeof(uint64_t)); }
#endif

// PCG32 known-answer helper (yocto_sampling.h:197-214): n floats of make_rng(seed, seq)
void ref_rng_floats(uint64_t seed, uint64_t seq, int n, float* out, uint64_t* state_inc) {
  auto rng = make_rng(seed, seq);
  if (state_inc) {
    state_inc[0] = rng.state;
    state_inc[1] = rng.inc;
  }
  for (int i = 0; i < n; i++) out[i] = rand1f(rng);
}

}  // extern "C"
