// ygl_api.cpp — the C ABI of include/ygl_b200.h: device arenas, the wavefront driver loop,
// batch intersection, and the multi-GPU tile gather. Host C++ only; all device work goes through
// the launch_* entry points of ygl_kernels.cu. There is no CPU fallback anywhere in this file:
// without a usable CUDA device every rendering call fails with YGL_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <memory>
#include <exception>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/ygl_b200.h"
#include "ygl_build.h"
#include "ygl_kernels.cuh"

using namespace ygl;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t err__ = (expr);                                                                     \
    if (err__ != cudaSuccess)                                                                       \
      return fail(YGL_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorName(err__) + " (" +      \
                                    cudaGetErrorString(err__) + ") at " #expr);                     \
  } while (0)

// A host-built byte arena uploaded with one cudaMemcpy; sub-allocations are 256-B aligned so every
// float4 / packet array starts on a sector boundary.
struct Arena {
  std::vector<uint8_t> host;
  uint8_t*             dev = nullptr;
  size_t add(const void* data, size_t bytes) {
    size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + bytes);
    if (bytes && data) memcpy(host.data() + off, data, bytes);
    return off;
  }
  template <typename T>
  const T* ptr(size_t off) const {
    return (const T*)(dev + off);
  }
  cudaError_t upload() {
    if (host.empty()) host.resize(256);
    cudaError_t e = cudaMalloc((void**)&dev, host.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice);
  }
  void release() {
    if (dev) cudaFree(dev);
    dev = nullptr;
  }
};

}  // namespace

struct ygl_context {
  int          device = 0;
  cudaStream_t stream = nullptr;
  int          num_sms = 0;
  // Wavefront pipelines: lanes of a state are split over up to kMaxPipes independent queue sets that run
  // on their own streams, so one pipeline's kernel ramp-up/drain overlaps the other's work (matters when a
  // tile is small, i.e. multi-GPU). Pipe 0 runs on `stream`.
  static constexpr int kMaxPipes = 2;
  struct Pipe {
    cudaStream_t stream      = nullptr;
    int          queue_lanes = 0;
    unsigned     queue_classes = 0;  // shade queues allocated (bit c = class c)
    int*         queue_mem   = nullptr;
    int*         park_mem    = nullptr;  // Queues::park (two halves)
    size_t       park_words  = 0;        // words per half
    Counters*    counters    = nullptr;
    int*         h_done      = nullptr;  // pinned, 2 slots
    cudaEvent_t  ev[2]       = {nullptr, nullptr};
    cudaEvent_t  join        = nullptr;
  } pipes[kMaxPipes];
  uint64_t  stats[16]   = {0};
  double    timings[4]  = {0};
  bool      time_kernels = false, count_traversal = false;
  int       mode = YGL_MODE_WAVEFRONT;
  Tuning    tune;  // scheduling knobs (ygl_context_set_option)
  unsigned long long*      d_trav = nullptr;  // 7 traversal counters
  std::vector<cudaEvent_t> ev_pool;           // event pairs around extend launches
  cudaEvent_t              ev_loop[2] = {nullptr, nullptr};
  // binding cache: DShape table for a (scene, bvh, lights) triple
  const ygl_scene*  bound_scene  = nullptr;
  const ygl_bvh*    bound_bvh    = nullptr;
  const ygl_lights* bound_lights = nullptr;
  uint64_t          bound_epoch  = 0, bound_bvh_epoch = 0, bound_lights_epoch = 0;
  DShape*           d_shapes     = nullptr;
  DLight*           d_lights     = nullptr;
  unsigned char*    d_inst_class = nullptr;  // shading class per instance (1 + material type)
  unsigned          class_mask   = 0;        // classes present in the bound scene (bit c = class c)
  DScene            dscene       = {};
  // progressive rendering (ygl_trace_start / cancel / done): one worker thread per context at a time
  std::thread      worker;
  std::atomic<int> stop{0}, done{0};
  int              worker_rc = YGL_OK;
  std::string      worker_error;
  // nccl
  void* nccl_lib  = nullptr;
  void* nccl_comm = nullptr;
  int   rank = 0, nranks = 1;
};

struct ygl_scene {
  Arena arena;
  int   device = 0;
  // offsets
  size_t off_cameras = 0, off_instances = 0, off_materials = 0, off_environments = 0, off_textures = 0;
  int    num_cameras = 0, num_instances = 0, num_materials = 0, num_environments = 0, num_textures = 0,
      num_shapes = 0;
  struct ShapeOff {
    size_t points, lines, triangles, quads, positions, normals, texcoords, colors, radius;
    int    np, nl, nt, nq, nnormals, ntexcoords, ncolors, nradius;
  };
  std::vector<ShapeOff> shapes;
  std::vector<ygl_camera> cameras_host;  // image sizes are derived from the camera aspect (make_trace_state)
  std::vector<uint8_t>  inst_class;   // shading class per instance: 1 + material type (kClsGeneric if out of range)
  unsigned              class_mask = 0;  // classes that can occur (bit c), incl. kClsMiss and, with media, kClsGeneric
  bool                  has_volumes = false;
  uint64_t              epoch = 0;
};

struct ygl_bvh {
  HostBvh  host;
  uint64_t epoch = 0;  // unique per object: a new bvh at a recycled address never matches a cached binding
  // device copy (made on first use, on the device of the context that first binds it)
  mutable Arena arena;
  mutable bool  uploaded = false;
  mutable int   device   = -1;
  mutable std::vector<size_t> off_nodes, off_packets, off_prims;
  mutable size_t off_top_nodes = 0, off_top_packets = 0, off_top_prims = 0, off_inst_packets = 0;
};

struct ygl_lights {
  std::vector<HostLight> host;
  uint64_t               epoch = 0;
  mutable Arena          arena;
  mutable bool           uploaded = false;
  mutable int            device   = -1;
  mutable std::vector<size_t> off_cdf;
};

struct ygl_state {
  ygl_context* ctx   = nullptr;
  int          width = 0, height = 0, samples = 0;
  int          row_begin = 0, row_end = 0;  // rows [row_begin, row_end) with stride row_step
  int          row_step = 1, num_rows = 0;
  PathState    st    = {};
  uint8_t*     mem   = nullptr;
};

static std::atomic<uint64_t> g_epoch{1};
constexpr int kSharedStackHost  = 28;  // = kSharedStack of ygl_traverse.cuh (the extend kernel's shared-memory stack)
constexpr int kShallowStackHost = 72;  // = kShallowStack of ygl_traverse.cuh (stack entries of the shallow extend kernel)

extern "C" {

const char* ygl_last_error(void) { return g_error.c_str(); }
// for the other translation units of the library (ygl_sceneio.cpp)
void ygl_internal_set_error(const char* message) { g_error = message ? message : ""; }
const char* ygl_version(void) { return "ygl_b200 0.2 (sm_100a wavefront path tracer; float libm = glibc 2.39 x86-64 FMA builds)"; }

void ygl_trace_params_default(ygl_trace_params* p) {
  *p            = ygl_trace_params{};
  p->camera     = 0;
  p->resolution = 1280;
  p->sampler    = YGL_SAMPLER_PATH;
  p->falsecolor = YGL_FALSECOLOR_COLOR;
  p->samples    = 512;
  p->bounces    = 8;
  p->clamp      = 10;
  p->seed       = YGL_DEFAULT_SEED;
  p->pratio     = 8;
  p->batch      = 1;
}

// ------------------------------------------------------------------------------------------
int ygl_context_create(int device, ygl_context** out) try {
  if (!out) return fail(YGL_ERR_INVALID, "null output");
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(YGL_ERR_CUDA, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  auto ctx    = std::make_unique<ygl_context>();
  ctx->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (int k = 0; k < ygl_context::kMaxPipes; k++) {
    auto& pipe = ctx->pipes[k];
    if (k == 0) pipe.stream = ctx->stream;
    else CUDA_TRY(cudaStreamCreateWithFlags(&pipe.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc((void**)&pipe.counters, sizeof(Counters)));
    CUDA_TRY(cudaMemset(pipe.counters, 0, sizeof(Counters)));
    CUDA_TRY(cudaHostAlloc((void**)&pipe.h_done, 2 * sizeof(int), cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&pipe.ev[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&pipe.ev[1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&pipe.join, cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreate(&ctx->ev_loop[0]));
  CUDA_TRY(cudaEventCreate(&ctx->ev_loop[1]));
  CUDA_TRY(cudaMalloc((void**)&ctx->d_trav, 8 * sizeof(unsigned long long)));
  *out = ctx.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_context_create: ") + e.what());
}

void ygl_comm_destroy(ygl_context* ctx);

void ygl_context_destroy(ygl_context* ctx) {
  if (!ctx) return;
  ctx->stop = 1;
  if (ctx->worker.joinable()) ctx->worker.join();
  cudaSetDevice(ctx->device);
  ygl_comm_destroy(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < ygl_context::kMaxPipes; k++) {
    auto& pipe = ctx->pipes[k];
    if (pipe.stream) cudaStreamSynchronize(pipe.stream);
    if (pipe.queue_mem) cudaFree(pipe.queue_mem);
    if (pipe.park_mem) cudaFree(pipe.park_mem);
    if (pipe.counters) cudaFree(pipe.counters);
    if (pipe.h_done) cudaFreeHost(pipe.h_done);
    for (auto& e : pipe.ev)
      if (e) cudaEventDestroy(e);
    if (pipe.join) cudaEventDestroy(pipe.join);
    if (k > 0 && pipe.stream) cudaStreamDestroy(pipe.stream);
  }
  if (ctx->d_shapes) cudaFree(ctx->d_shapes);
  if (ctx->d_lights) cudaFree(ctx->d_lights);
  if (ctx->d_inst_class) cudaFree(ctx->d_inst_class);
  if (ctx->d_trav) cudaFree(ctx->d_trav);
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto& e : ctx->ev_loop)
    if (e) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int ygl_context_synchronize(ygl_context* ctx) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return YGL_OK;
}
void* ygl_context_stream(ygl_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ------------------------------------------------------------------------------------------
static int check_desc(const ygl_scene_desc* d) {
  if (!d) return fail(YGL_ERR_INVALID, "null scene description");
  for (int i = 0; i < d->num_instances; i++) {
    if (d->instances[i].shape < 0 || d->instances[i].shape >= d->num_shapes)
      return fail(YGL_ERR_INVALID, "instance shape id out of range");
    if (d->instances[i].material < 0 || d->instances[i].material >= d->num_materials)
      return fail(YGL_ERR_INVALID, "instance material id out of range");
  }
  auto texok = [&](int t) { return t == YGL_INVALID_ID || (t >= 0 && t < d->num_textures); };
  for (int i = 0; i < d->num_materials; i++) {
    auto& m = d->materials[i];
    if (!texok(m.emission_tex) || !texok(m.color_tex) || !texok(m.roughness_tex) || !texok(m.scattering_tex) ||
        !texok(m.normal_tex))
      return fail(YGL_ERR_INVALID, "material texture id out of range");
  }
  for (int i = 0; i < d->num_environments; i++)
    if (!texok(d->environments[i].emission_tex)) return fail(YGL_ERR_INVALID, "environment texture id out of range");
  return YGL_OK;
}

static DCamera to_dcamera(const ygl_camera& c) {
  DCamera d;
  memcpy(&d.frame, &c.frame, sizeof(frame3));
  d.orthographic = c.orthographic;
  d.lens = c.lens, d.film = c.film, d.aspect = c.aspect, d.focus = c.focus, d.aperture = c.aperture;
  return d;
}

int ygl_scene_create(ygl_context* ctx, const ygl_scene_desc* desc, ygl_scene** out) try {
  if (!ctx || !out) return fail(YGL_ERR_INVALID, "null argument");
  if (int rc = check_desc(desc)) return rc;
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto scene    = std::make_unique<ygl_scene>();
  scene->device = ctx->device;
  scene->epoch  = g_epoch++;
  Arena& A      = scene->arena;

  std::vector<DCamera> cams(desc->num_cameras);
  for (int i = 0; i < desc->num_cameras; i++) cams[i] = to_dcamera(desc->cameras[i]);
  scene->off_cameras = A.add(cams.data(), cams.size() * sizeof(DCamera));
  scene->num_cameras = desc->num_cameras;
  scene->cameras_host.assign(desc->cameras, desc->cameras + desc->num_cameras);

  std::vector<DInstance> insts(desc->num_instances);
  for (int i = 0; i < desc->num_instances; i++) {
    memcpy(&insts[i].frame, &desc->instances[i].frame, sizeof(frame3));
    insts[i].shape = desc->instances[i].shape, insts[i].material = desc->instances[i].material;
    insts[i].pad0 = insts[i].pad1 = 0;
  }
  scene->off_instances = A.add(insts.data(), insts.size() * sizeof(DInstance));
  scene->num_instances = desc->num_instances;
  // shading classes (ygl_kernels.cuh): what kind of shading work a hit on each instance is
  scene->inst_class.resize(desc->num_instances);
  scene->class_mask = 1u << kClsMiss;
  for (int i = 0; i < desc->num_instances; i++) {
    const int type = desc->materials[desc->instances[i].material].type;
    const int cls  = type >= 0 && type <= 7 ? 1 + type : kClsGeneric;
    scene->inst_class[i] = (uint8_t)cls;
    scene->class_mask |= 1u << cls;
    // refractive, subsurface and volumetric surfaces open a participating medium (yocto_scene.cpp:257-261): the
    // lanes inside it are shaded by the unspecialised kernel
    if (type == YGL_MATERIAL_REFRACTIVE || type == YGL_MATERIAL_SUBSURFACE || type == YGL_MATERIAL_VOLUMETRIC)
      scene->has_volumes = true;
  }
  if (scene->has_volumes) scene->class_mask |= 1u << kClsGeneric;

  static_assert(sizeof(DMaterial) == sizeof(ygl_material), "material layout");
  scene->off_materials = A.add(desc->materials, (size_t)desc->num_materials * sizeof(DMaterial));
  scene->num_materials = desc->num_materials;

  std::vector<DEnvironment> envs(desc->num_environments);
  for (int i = 0; i < desc->num_environments; i++) {
    auto& e = desc->environments[i];
    memcpy(&envs[i].frame, &e.frame, sizeof(frame3));
    envs[i].inv_frame    = frame_inverse(envs[i].frame, false);  // inverse(environment.frame), yocto_scene.cpp:598
    envs[i].emission     = {e.emission[0], e.emission[1], e.emission[2]};
    envs[i].emission_tex = e.emission_tex;
  }
  scene->off_environments = A.add(envs.data(), envs.size() * sizeof(DEnvironment));
  scene->num_environments = desc->num_environments;

  // textures: pixel arrays first, table patched after upload
  std::vector<size_t> tex_off(desc->num_textures);
  for (int i = 0; i < desc->num_textures; i++) {
    auto&  t = desc->textures[i];
    size_t n = (size_t)t.width * t.height;
    if (t.pixelsf)
      tex_off[i] = A.add(t.pixelsf, n * 16);
    else if (t.pixelsb)
      tex_off[i] = A.add(t.pixelsb, n * 4);
    else if (n)
      return fail(YGL_ERR_INVALID, "texture without pixels");
  }
  scene->off_textures = A.add(nullptr, (size_t)desc->num_textures * sizeof(DTexture));
  scene->num_textures = desc->num_textures;

  scene->num_shapes = desc->num_shapes;
  scene->shapes.resize(desc->num_shapes);
  for (int i = 0; i < desc->num_shapes; i++) {
    auto& s = desc->shapes[i];
    auto& o = scene->shapes[i];
    o.np = s.num_points, o.nl = s.num_lines, o.nt = s.num_triangles, o.nq = s.num_quads;
    o.nnormals = s.num_normals, o.ntexcoords = s.num_texcoords, o.ncolors = s.num_colors, o.nradius = s.num_radius;
    o.points    = A.add(s.points, (size_t)s.num_points * 4);
    o.lines     = A.add(s.lines, (size_t)s.num_lines * 8);
    o.triangles = A.add(s.triangles, (size_t)s.num_triangles * 12);
    o.quads     = A.add(s.quads, (size_t)s.num_quads * 16);
    o.positions = A.add(s.positions, (size_t)s.num_positions * 12);
    o.normals   = A.add(s.normals, (size_t)s.num_normals * 12);
    o.texcoords = A.add(s.texcoords, (size_t)s.num_texcoords * 8);
    o.colors    = A.add(s.colors, (size_t)s.num_colors * 16);
    o.radius    = A.add(s.radius, (size_t)s.num_radius * 4);
  }
  // patch the texture table with device pointers: allocate first to learn the base address
  if (A.host.empty()) A.host.resize(256);
  CUDA_TRY(cudaMalloc((void**)&A.dev, A.host.size()));
  for (int i = 0; i < desc->num_textures; i++) {
    auto&    t = desc->textures[i];
    DTexture d = {t.width, t.height, t.linear, t.nearest, t.clamp, 0, nullptr, nullptr};
    if (t.pixelsf) d.pixelsf = A.ptr<float4>(tex_off[i]);
    else if (t.pixelsb) d.pixelsb = A.ptr<uchar4>(tex_off[i]);
    memcpy(A.host.data() + scene->off_textures + i * sizeof(DTexture), &d, sizeof(DTexture));
  }
  CUDA_TRY(cudaMemcpy(A.dev, A.host.data(), A.host.size(), cudaMemcpyHostToDevice));
  A.host.clear();
  A.host.shrink_to_fit();
  *out = scene.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_scene_create: ") + e.what());
}

int ygl_scene_update_cameras(ygl_scene* scene, const ygl_camera* cameras, int num_cameras) {
  if (!scene || !cameras) return fail(YGL_ERR_INVALID, "null argument");
  if (num_cameras != scene->num_cameras) return fail(YGL_ERR_INVALID, "camera count mismatch");
  CUDA_TRY(cudaSetDevice(scene->device));
  std::vector<DCamera> cams(num_cameras);
  for (int i = 0; i < num_cameras; i++) cams[i] = to_dcamera(cameras[i]);
  CUDA_TRY(cudaMemcpy(scene->arena.dev + scene->off_cameras, cams.data(), cams.size() * sizeof(DCamera),
      cudaMemcpyHostToDevice));
  scene->cameras_host.assign(cameras, cameras + num_cameras);
  return YGL_OK;
}

void ygl_scene_destroy(ygl_scene* scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  scene->arena.release();
  delete scene;
}

// ------------------------------------------------------------------------------------------
int ygl_bvh_build(const ygl_scene_desc* desc, int highquality, ygl_bvh** out) try {
  if (!out) return fail(YGL_ERR_INVALID, "null output");
  if (int rc = check_desc(desc)) return rc;
  auto        bvh = std::make_unique<ygl_bvh>();
  std::string error;
  if (!build_scene_bvh(*desc, highquality != 0, bvh->host, error)) return fail(YGL_ERR_INVALID, error);
  bvh->epoch = g_epoch++;
  *out       = bvh.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_bvh_build: ") + e.what());
}
int ygl_bvh_build_device(ygl_context* ctx, const ygl_scene_desc* desc, int highquality, ygl_bvh** out) try {
  if (!ctx || !out) return fail(YGL_ERR_INVALID, "null argument");
  if (int rc = check_desc(desc)) return rc;
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto        bvh = std::make_unique<ygl_bvh>();
  std::string error;
  if (!build_scene_bvh(*desc, highquality != 0, bvh->host, error, (void*)ctx->stream, true))
    return fail(YGL_ERR_INVALID, error);
  bvh->epoch = g_epoch++;
  *out       = bvh.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_bvh_build_device: ") + e.what());
}
int ygl_bvh_create_from_host(const ygl_scene_desc* desc, const ygl_bvh_node* top_nodes, int num_top_nodes,
    const int32_t* top_primitives, int num_top_primitives, const ygl_bvh_node* const* shape_nodes,
    const int* shape_num_nodes, const int32_t* const* shape_primitives, const int* shape_num_primitives, ygl_bvh** out) try {
  if (!out) return fail(YGL_ERR_INVALID, "null output");
  if (int rc = check_desc(desc)) return rc;
  if (!top_nodes || num_top_nodes < 1 || (num_top_primitives > 0 && !top_primitives))
    return fail(YGL_ERR_INVALID, "instance tree missing");
  if (desc->num_shapes > 0 && (!shape_nodes || !shape_num_nodes || !shape_primitives || !shape_num_primitives))
    return fail(YGL_ERR_INVALID, "shape trees missing");
  for (int i = 0; i < desc->num_shapes; i++)
    if (!shape_nodes[i] || shape_num_nodes[i] < 1 || shape_num_primitives[i] < 0 ||
        (shape_num_primitives[i] > 0 && !shape_primitives[i]))
      return fail(YGL_ERR_INVALID, "shape tree " + std::to_string(i) + " missing");
  auto        bvh = std::make_unique<ygl_bvh>();
  std::string error;
  if (!adopt_scene_bvh(*desc, top_nodes, num_top_nodes, top_primitives, num_top_primitives, shape_nodes, shape_num_nodes,
          shape_primitives, shape_num_primitives, bvh->host, error))
    return fail(YGL_ERR_INVALID, error);
  bvh->epoch = g_epoch++;
  *out       = bvh.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_bvh_create_from_host: ") + e.what());
}
int ygl_bvh_update(ygl_bvh* bvh, const ygl_scene_desc* desc, const int* updated_instances, int num_updated_instances,
    const int* updated_shapes, int num_updated_shapes) try {
  if (!bvh) return fail(YGL_ERR_INVALID, "null bvh");
  if (int rc = check_desc(desc)) return rc;
  if (num_updated_instances < 0 || num_updated_shapes < 0 || (num_updated_shapes > 0 && !updated_shapes))
    return fail(YGL_ERR_INVALID, "updated id lists missing");
  (void)updated_instances;  // all instance boxes are refreshed, as in the reference
  std::string error;
  if (!update_scene_bvh(*desc, updated_shapes, num_updated_shapes, bvh->host, error)) return fail(YGL_ERR_INVALID, error);
  // a new epoch drops every cached binding; the device copy is rebuilt on the next use
  if (bvh->uploaded) {
    cudaSetDevice(bvh->device);
    cudaDeviceSynchronize();
    bvh->arena.release();
    bvh->arena.host.clear();
    bvh->uploaded = false;
  }
  bvh->epoch = g_epoch++;
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_bvh_update: ") + e.what());
}
static const HostTree* pick_tree(const ygl_bvh* bvh, int shape) {
  if (shape < 0) return &bvh->host.top;
  if (shape >= (int)bvh->host.shapes.size()) return nullptr;
  return &bvh->host.shapes[shape];
}
int ygl_bvh_tree_size(const ygl_bvh* bvh, int shape, int* num_nodes, int* num_primitives) {
  if (!bvh) return fail(YGL_ERR_INVALID, "null bvh");
  auto tree = pick_tree(bvh, shape);
  if (!tree) return fail(YGL_ERR_INVALID, "shape id out of range");
  if (num_nodes) *num_nodes = (int)tree->nodes.size();
  if (num_primitives) *num_primitives = (int)tree->prims.size();
  return YGL_OK;
}
int ygl_bvh_tree_get(const ygl_bvh* bvh, int shape, ygl_bvh_node* nodes, int32_t* primitives) {
  if (!bvh) return fail(YGL_ERR_INVALID, "null bvh");
  auto tree = pick_tree(bvh, shape);
  if (!tree) return fail(YGL_ERR_INVALID, "shape id out of range");
  if (nodes) memcpy(nodes, tree->nodes.data(), tree->nodes.size() * sizeof(ygl_bvh_node));
  if (primitives) memcpy(primitives, tree->prims.data(), tree->prims.size() * sizeof(int32_t));
  return YGL_OK;
}
void ygl_bvh_destroy(ygl_bvh* bvh) {
  if (!bvh) return;
  if (bvh->uploaded) cudaSetDevice(bvh->device);
  bvh->arena.release();
  delete bvh;
}

static int bvh_upload(const ygl_bvh* bvh, int device) {
  if (bvh->uploaded) {
    // the instance packets carry device pointers: one device copy per bvh object
    if (bvh->device != device) return fail(YGL_ERR_INVALID, "bvh was uploaded to another device: build one per GPU");
    return YGL_OK;
  }
  Arena& A  = bvh->arena;
  auto&  H  = bvh->host;
  size_t ns = H.shapes.size();
  bvh->off_nodes.resize(ns), bvh->off_packets.resize(ns), bvh->off_prims.resize(ns);
  for (size_t i = 0; i < ns; i++) {
    bvh->off_nodes[i]   = A.add(H.shape_nodes[i].data(), H.shape_nodes[i].size() * 16);
    bvh->off_packets[i] = A.add(H.shape_packets[i].data(), H.shape_packets[i].size() * 16);
    bvh->off_prims[i]   = A.add(H.shapes[i].prims.data(), H.shapes[i].prims.size() * 4);
  }
  bvh->off_top_nodes    = A.add(H.top_nodes.data(), H.top_nodes.size() * 16);
  bvh->off_top_packets  = A.add(H.top_packets.data(), H.top_packets.size() * 16);
  bvh->off_top_prims    = A.add(H.top.prims.data(), H.top.prims.size() * 4);
  bvh->off_inst_packets = A.add(H.inst_packets.data(), H.inst_packets.size() * 16);
  // allocate first so the instance packets can carry the device pointers of their shape's tree
  CUDA_TRY(cudaMalloc((void**)&A.dev, A.host.size()));
  auto patch = [&](size_t off, size_t count) {
    for (size_t k = 0; k < count; k++) {
      float4h* q     = (float4h*)(A.host.data() + off) + k * kInstancePacketQuads;
      int      shape = 0;
      memcpy(&shape, &q[3].x, 4);
      const void* ptrs[3] = {A.ptr<uint8_t>(bvh->off_nodes[shape]), A.ptr<uint8_t>(bvh->off_packets[shape]),
          A.ptr<uint8_t>(bvh->off_prims[shape])};
      memcpy(&q[4].x, &ptrs[0], 8);
      memcpy(&q[4].z, &ptrs[1], 8);
      memcpy(&q[5].x, &ptrs[2], 8);
    }
  };
  patch(bvh->off_top_packets, H.top.prims.size());
  patch(bvh->off_inst_packets, H.inst_packets.size() / kInstancePacketQuads);
  CUDA_TRY(cudaMemcpy(A.dev, A.host.data(), A.host.size(), cudaMemcpyHostToDevice));
  A.host.clear();
  A.host.shrink_to_fit();
  bvh->uploaded = true;
  bvh->device   = device;
  return YGL_OK;
}

// ------------------------------------------------------------------------------------------
int ygl_lights_create(const ygl_scene_desc* desc, ygl_lights** out) try {
  if (!out) return fail(YGL_ERR_INVALID, "null output");
  if (int rc = check_desc(desc)) return rc;
  auto lights = std::make_unique<ygl_lights>();
  build_lights(*desc, lights->host);
  lights->epoch = g_epoch++;
  *out          = lights.release();
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_lights_create: ") + e.what());
}
int ygl_lights_count(const ygl_lights* lights) { return lights ? (int)lights->host.size() : 0; }
int ygl_lights_get(const ygl_lights* lights, int i, int* instance, int* environment, int* cdf_size, float* cdf) {
  if (!lights || i < 0 || i >= (int)lights->host.size()) return fail(YGL_ERR_INVALID, "light id out of range");
  auto& l = lights->host[i];
  if (instance) *instance = l.instance;
  if (environment) *environment = l.environment;
  if (cdf_size) *cdf_size = (int)l.cdf.size();
  if (cdf) memcpy(cdf, l.cdf.data(), l.cdf.size() * sizeof(float));
  return YGL_OK;
}
void ygl_lights_destroy(ygl_lights* lights) {
  if (!lights) return;
  if (lights->uploaded) cudaSetDevice(lights->device);
  lights->arena.release();
  delete lights;
}
static int lights_upload(const ygl_lights* lights, int device) {
  if (lights->uploaded) {
    if (lights->device != device) return fail(YGL_ERR_INVALID, "lights were uploaded to another device: create one per GPU");
    return YGL_OK;
  }
  Arena& A = lights->arena;
  lights->off_cdf.resize(lights->host.size());
  for (size_t i = 0; i < lights->host.size(); i++)
    lights->off_cdf[i] = A.add(lights->host[i].cdf.data(), lights->host[i].cdf.size() * 4);
  CUDA_TRY(A.upload());
  A.host.clear();
  A.host.shrink_to_fit();
  lights->uploaded = true;
  lights->device   = device;
  return YGL_OK;
}

// Assemble the DScene for a (scene, bvh, lights) triple; cached on the context.
// any_lights: the caller does not read lights (batch intersection): whatever light table is bound may stay.
static int bind_scene(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh, const ygl_lights* lights,
    bool any_lights = false) {
  if (!scene || !bvh) return fail(YGL_ERR_INVALID, "null scene or bvh");
  if (scene->device != ctx->device) return fail(YGL_ERR_INVALID, "scene lives on another device");
  if ((int)bvh->host.shapes.size() != scene->num_shapes ||
      (int)(bvh->host.inst_packets.size() / kInstancePacketQuads) != scene->num_instances)
    return fail(YGL_ERR_INVALID, "bvh was built for a different scene");
  // the cache key is the identity of the three OBJECTS (their epochs), not their addresses: a bvh or lights object
  // destroyed and re-created at the same heap address must not hit
  if (ctx->bound_scene == scene && ctx->bound_bvh == bvh && ctx->bound_epoch == scene->epoch &&
      ctx->bound_bvh_epoch == bvh->epoch &&
      (any_lights || (ctx->bound_lights == lights && ctx->bound_lights_epoch == (lights ? lights->epoch : 0))))
    return YGL_OK;
  ctx->bound_scene = nullptr;  // a failure below must not leave a half-built binding cached
  if (int rc = bvh_upload(bvh, ctx->device)) return rc;
  if (lights)
    if (int rc = lights_upload(lights, ctx->device)) return rc;
  const Arena& S = scene->arena;
  const Arena& B = bvh->arena;
  std::vector<DShape> shapes(scene->num_shapes);
  for (int i = 0; i < scene->num_shapes; i++) {
    auto&   o = scene->shapes[i];
    DShape& d = shapes[i];
    d.nodes     = B.ptr<float4>(bvh->off_nodes[i]);
    d.packets   = B.ptr<float4>(bvh->off_packets[i]);
    d.prims     = B.ptr<int>(bvh->off_prims[i]);
    d.num_nodes = (int)bvh->host.shapes[i].nodes.size();
    d.bvh_kind  = bvh->host.shape_kind[i];
    d.eval_kind = o.nt > 0 ? kElemTriangles : o.nq > 0 ? kElemQuads : o.nl > 0 ? kElemLines : o.np > 0 ? kElemPoints : kElemNone;
    d.pad0      = 0;
    d.points    = S.ptr<int>(o.points);
    d.lines     = S.ptr<int>(o.lines);
    d.triangles = S.ptr<int>(o.triangles);
    d.quads     = S.ptr<int>(o.quads);
    d.positions = S.ptr<float>(o.positions);
    d.normals   = o.nnormals ? S.ptr<float>(o.normals) : nullptr;
    d.texcoords = o.ntexcoords ? S.ptr<float>(o.texcoords) : nullptr;
    d.colors    = o.ncolors ? S.ptr<float>(o.colors) : nullptr;
    d.radius    = o.nradius ? S.ptr<float>(o.radius) : nullptr;
    d.num_points = o.np, d.num_lines = o.nl, d.num_triangles = o.nt, d.num_quads = o.nq;
  }
  if (ctx->d_shapes) cudaFree(ctx->d_shapes), ctx->d_shapes = nullptr;
  if (ctx->d_lights) cudaFree(ctx->d_lights), ctx->d_lights = nullptr;
  if (ctx->d_inst_class) cudaFree(ctx->d_inst_class), ctx->d_inst_class = nullptr;
  CUDA_TRY(cudaMalloc((void**)&ctx->d_inst_class, std::max<size_t>(1, scene->inst_class.size())));
  CUDA_TRY(cudaMemcpy(ctx->d_inst_class, scene->inst_class.data(), scene->inst_class.size(), cudaMemcpyHostToDevice));
  ctx->class_mask = scene->class_mask;
  CUDA_TRY(cudaMalloc((void**)&ctx->d_shapes, std::max<size_t>(1, shapes.size()) * sizeof(DShape)));
  CUDA_TRY(cudaMemcpy(ctx->d_shapes, shapes.data(), shapes.size() * sizeof(DShape), cudaMemcpyHostToDevice));
  int nlights = lights ? (int)lights->host.size() : 0;
  if (nlights) {
    std::vector<DLight> dl(nlights);
    for (int i = 0; i < nlights; i++) {
      auto& l = lights->host[i];
      dl[i]   = {l.instance, l.environment, lights->arena.ptr<float>(lights->off_cdf[i]), (int)l.cdf.size(), 0};
    }
    CUDA_TRY(cudaMalloc((void**)&ctx->d_lights, dl.size() * sizeof(DLight)));
    CUDA_TRY(cudaMemcpy(ctx->d_lights, dl.data(), dl.size() * sizeof(DLight), cudaMemcpyHostToDevice));
  }
  DScene& D          = ctx->dscene;
  D.cameras          = S.ptr<DCamera>(scene->off_cameras);
  D.instances        = S.ptr<DInstance>(scene->off_instances);
  D.inst_packets     = B.ptr<DInstancePacket>(bvh->off_inst_packets);
  D.materials        = S.ptr<DMaterial>(scene->off_materials);
  D.environments     = S.ptr<DEnvironment>(scene->off_environments);
  D.textures         = S.ptr<DTexture>(scene->off_textures);
  D.shapes           = ctx->d_shapes;
  D.num_cameras      = scene->num_cameras;
  D.num_instances    = scene->num_instances;
  D.num_materials    = scene->num_materials;
  D.num_environments = scene->num_environments;
  D.num_textures     = scene->num_textures;
  D.num_shapes       = scene->num_shapes;
  D.top_nodes        = B.ptr<float4>(bvh->off_top_nodes);
  D.top_packets      = B.ptr<DInstancePacket>(bvh->off_top_packets);
  D.top_prims        = B.ptr<int>(bvh->off_top_prims);
  D.top_num_nodes    = (int)bvh->host.top.nodes.size();
  D.lights           = ctx->d_lights;
  D.num_lights       = nlights;
  D.inst_class       = nullptr;  // set per run (binned shade queues: path sampler only)
  D.has_volumes      = scene->has_volumes ? 1 : 0;
  int shape_depth = 0;
  for (auto& tree : bvh->host.shapes) shape_depth = std::max(shape_depth, tree.max_stack);
  // entries a walk can hold: one far sibling per level of both trees + the instance run, the EXIT marker, the sentinel
  const int stack_need = bvh->host.top.max_stack + shape_depth + 4;
  D.stack_mode = stack_need > kShallowStackHost ? 2 : stack_need > kSharedStackHost ? 1 : 0;
  ctx->bound_scene = scene, ctx->bound_bvh = bvh, ctx->bound_lights = lights, ctx->bound_epoch = scene->epoch;
  ctx->bound_bvh_epoch = bvh->epoch, ctx->bound_lights_epoch = lights ? lights->epoch : 0;
  return YGL_OK;
}

// ------------------------------------------------------------------------------------------
int ygl_make_state_rngs(const ygl_scene_desc* desc, const ygl_trace_params* params, int* width, int* height,
    uint64_t* rngs) try {
  if (!desc || !params) return fail(YGL_ERR_INVALID, "null argument");
  int         w, h;
  std::string error;
  if (!state_size(*desc, *params, w, h, error)) return fail(YGL_ERR_INVALID, error);
  if (width) *width = w;
  if (height) *height = h;
  if (rngs) state_rngs(*params, w, h, rngs);
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_make_state_rngs: ") + e.what());
}

// Rows row_begin, row_begin + row_step, ... below row_end of the full image.
static int state_create_sized(ygl_context* ctx, int w, int h, const ygl_trace_params* params, int row_begin,
    int row_end, int row_step, ygl_state** out);
static int state_create_rows(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params,
    int row_begin, int row_end, int row_step, ygl_state** out) {
  if (!ctx || !desc || !params || !out) return fail(YGL_ERR_INVALID, "null argument");
  int         w, h;
  std::string error;
  if (!state_size(*desc, *params, w, h, error)) return fail(YGL_ERR_INVALID, error);
  return state_create_sized(ctx, w, h, params, row_begin, row_end, row_step, out);
}
static int state_create_sized(ygl_context* ctx, int w, int h, const ygl_trace_params* params, int row_begin,
    int row_end, int row_step, ygl_state** out) {
  if (row_end < 0) row_end = h;
  if (row_begin < 0 || row_begin > row_end || row_end > h || row_step < 1) return fail(YGL_ERR_INVALID, "bad row range");
  const int num_rows = (row_end - row_begin + row_step - 1) / row_step;
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto state       = std::make_unique<ygl_state>();
  state->ctx       = ctx;
  state->width     = w;
  state->height    = h;
  state->row_begin = row_begin;
  state->row_end   = row_end;
  state->row_step  = row_step;
  state->num_rows  = num_rows;
  size_t n         = (size_t)w * num_rows;
  size_t lanes     = std::max<size_t>(n, 1);
  // one allocation, 256-B aligned slices
  size_t off = 0;
  auto   take = [&](size_t bytes) {
    size_t o = off;
    off      = (off + bytes + 255) & ~size_t(255);
    return o;
  };
  // reference-layout accumulators (downloaded as they are), per-lane scalars, then the path state as interleaved
  // groups of 16-byte records (PathStateT)
  size_t o_image = take(lanes * 16), o_albedo = take(lanes * 12), o_normal = take(lanes * 12), o_hits = take(lanes * 4),
         o_rngs = take(lanes * 16), o_sample = take(lanes * 4), o_hid = take(lanes * 8), o_susp = take(lanes * 4),
         o_ai = take(lanes * 8), o_ni = take(lanes * 8);
  size_t o_groups = take(lanes * 16 * 16);  // 16 float4 records per lane, interleaved in groups (see below)
  CUDA_TRY(cudaMalloc((void**)&state->mem, off));
  CUDA_TRY(cudaMemset(state->mem, 0, off));
  PathState& st = state->st;
  st.num_lanes  = (int)n;
  st.width = w, st.height = h, st.row_begin = row_begin, st.row_step = row_step;
  uint8_t* m  = state->mem;
  st.image    = (float4*)(m + o_image);
  st.albedo   = (float*)(m + o_albedo);
  st.normal   = (float*)(m + o_normal);
  st.hits     = (int*)(m + o_hits);
  st.rngs     = (ulonglong2*)(m + o_rngs);
  st.sample   = (int*)(m + o_sample);
  st.hit_ids  = (int2*)(m + o_hid);
  st.susp     = (int*)(m + o_susp);
  st.aux_ids  = (int2*)(m + o_ai);
  st.next_ids = (int2*)(m + o_ni);
  SRec<false> spare;
  SRec<false>* members[16] = {&st.ray_o, &st.ray_d,  // read together by extend, shade, light pdf
      &st.radiance, &st.weight,                          // shade, accumulate
      &st.hit_uvd, &st.pend,                             // shade reads the hit and writes the pending MIS numerator
      &st.albedo0, &st.normal0,                          // bounce-0 guides
      &st.vol_a, &st.vol_b,                              // the volume slot
      &st.aux_o, &st.aux_dir,                            // shadow-ray origin and direction (pathdirect / pathmis)
      &st.aux_uvd, &st.aux_bsdf, &st.next_uvd, &spare};
  constexpr int kGroup = YGL_STATE_GROUP;  // consecutive members share one 16 * kGroup byte record per lane
  for (int k = 0; k < 16; k++) *members[k] = (float4*)(m + o_groups + (size_t)(k / kGroup) * lanes * 16 * kGroup) + k % kGroup;
  // rng table: seeded sequentially over the FULL image (yocto_trace.cpp:1512-1515), tile slice uploaded
  std::vector<uint64_t> rngs((size_t)w * h * 2);
  state_rngs(*params, w, h, rngs.data());
  if (n)
    CUDA_TRY(cudaMemcpy2D(st.rngs, (size_t)w * 16, rngs.data() + (size_t)row_begin * w * 2, (size_t)row_step * w * 16,
        (size_t)w * 16, num_rows, cudaMemcpyHostToDevice));
  *out = state.release();
  return YGL_OK;
}

int ygl_state_create_tile(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params,
    int row_begin, int row_end, ygl_state** out) {
  return state_create_rows(ctx, desc, params, row_begin, row_end, 1, out);
}
int ygl_state_create_interleaved(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params,
    int rank, int nranks, ygl_state** out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(YGL_ERR_INVALID, "bad rank");
  int         w, h;
  std::string error;
  if (!desc || !params) return fail(YGL_ERR_INVALID, "null argument");
  if (!state_size(*desc, *params, w, h, error)) return fail(YGL_ERR_INVALID, error);
  return state_create_rows(ctx, desc, params, std::min(rank, h), h, nranks, out);
}

int ygl_state_create(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params, ygl_state** out) {
  return ygl_state_create_tile(ctx, desc, params, 0, -1, out);
}

int ygl_state_size(const ygl_state* state, int* width, int* height, int* samples) {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  if (width) *width = state->width;
  if (height) *height = state->height;
  if (samples) *samples = state->samples;
  return YGL_OK;
}
int ygl_state_rows(const ygl_state* state, int* row_begin, int* row_end) {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  if (row_begin) *row_begin = state->row_begin;
  if (row_end) *row_end = state->row_end;
  return YGL_OK;
}
int ygl_state_layout(const ygl_state* state, int* row_first, int* row_step, int* num_rows) {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  if (row_first) *row_first = state->row_begin;
  if (row_step) *row_step = state->row_step;
  if (num_rows) *num_rows = state->num_rows;
  return YGL_OK;
}

int ygl_state_download(ygl_state* state, float* image, float* albedo, float* normal, int32_t* hits, uint64_t* rngs) try {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  CUDA_TRY(cudaSetDevice(state->ctx->device));
  cudaStream_t s = state->ctx->stream;
  size_t       n = state->st.num_lanes;
  if (image) CUDA_TRY(cudaMemcpyAsync(image, state->st.image, n * 16, cudaMemcpyDeviceToHost, s));
  if (albedo) CUDA_TRY(cudaMemcpyAsync(albedo, state->st.albedo, n * 12, cudaMemcpyDeviceToHost, s));
  if (normal) CUDA_TRY(cudaMemcpyAsync(normal, state->st.normal, n * 12, cudaMemcpyDeviceToHost, s));
  if (hits) CUDA_TRY(cudaMemcpyAsync(hits, state->st.hits, n * 4, cudaMemcpyDeviceToHost, s));
  if (rngs) CUDA_TRY(cudaMemcpyAsync(rngs, state->st.rngs, n * 16, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_state_download: ") + e.what());
}

int ygl_state_upload(ygl_state* state, int samples, const float* image, const float* albedo, const float* normal,
    const int32_t* hits, const uint64_t* rngs) try {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  CUDA_TRY(cudaSetDevice(state->ctx->device));
  cudaStream_t s = state->ctx->stream;
  size_t       n = state->st.num_lanes;
  if (image) CUDA_TRY(cudaMemcpyAsync(state->st.image, image, n * 16, cudaMemcpyHostToDevice, s));
  if (albedo) CUDA_TRY(cudaMemcpyAsync(state->st.albedo, albedo, n * 12, cudaMemcpyHostToDevice, s));
  if (normal) CUDA_TRY(cudaMemcpyAsync(state->st.normal, normal, n * 12, cudaMemcpyHostToDevice, s));
  if (hits) CUDA_TRY(cudaMemcpyAsync(state->st.hits, hits, n * 4, cudaMemcpyHostToDevice, s));
  if (rngs) CUDA_TRY(cudaMemcpyAsync(state->st.rngs, rngs, n * 16, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  state->samples = samples;
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_state_upload: ") + e.what());
}

int ygl_state_reset(ygl_state* state, const ygl_trace_params* params) {
  if (!state || !params) return fail(YGL_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(state->ctx->device));
  cudaStream_t s  = state->ctx->stream;
  PathState&   st = state->st;
  const size_t n  = (size_t)st.num_lanes;
  CUDA_TRY(cudaMemsetAsync((float4*)st.image, 0, n * 16, s));
  CUDA_TRY(cudaMemsetAsync((float*)st.albedo, 0, n * 12, s));
  CUDA_TRY(cudaMemsetAsync((float*)st.normal, 0, n * 12, s));
  CUDA_TRY(cudaMemsetAsync((int*)st.hits, 0, n * 4, s));
  std::vector<uint64_t> rngs((size_t)state->width * state->height * 2);
  state_rngs(*params, state->width, state->height, rngs.data());
  if (n)
    CUDA_TRY(cudaMemcpy2DAsync((ulonglong2*)st.rngs, (size_t)state->width * 16,
        rngs.data() + (size_t)state->row_begin * state->width * 2, (size_t)state->row_step * state->width * 16,
        (size_t)state->width * 16, state->num_rows, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));  // `rngs` is pageable host memory: keep it alive until the copy is done
  state->samples = 0;
  return YGL_OK;
}

void ygl_state_destroy(ygl_state* state) {
  if (!state) return;
  cudaSetDevice(state->ctx->device);
  if (state->mem) cudaFree(state->mem);
  delete state;
}

// ------------------------------------------------------------------------------------------
// Queue memory of a pipeline: gen[2], ext[2], lpdf, acc and one shade queue per class in `classes` (class 0 always).
static int ensure_queues(ygl_context::Pipe& pipe, int lanes, unsigned classes, int num_sms) {
  classes |= 1u;
  if (!pipe.park_mem) {
    pipe.park_words = (size_t)extend_grid_threads(num_sms) * kSuspendWords;
    CUDA_TRY(cudaMalloc((void**)&pipe.park_mem, 2 * pipe.park_words * sizeof(int)));
  }
  if (lanes <= pipe.queue_lanes && (classes & ~pipe.queue_classes) == 0) return YGL_OK;
  if (pipe.queue_mem) cudaFree(pipe.queue_mem), pipe.queue_mem = nullptr;
  lanes   = std::max(lanes, pipe.queue_lanes);
  classes |= pipe.queue_classes;
  size_t per = ((size_t)lanes + 63) & ~size_t(63);
  int    nq  = 6;
  for (int c = 0; c < kNumClasses; c++) nq += (classes >> c) & 1;
  CUDA_TRY(cudaMalloc((void**)&pipe.queue_mem, per * nq * sizeof(int)));
  pipe.queue_lanes   = lanes;
  pipe.queue_classes = classes;
  return YGL_OK;
}
static Queues make_queues(ygl_context::Pipe& pipe) {
  size_t per = ((size_t)pipe.queue_lanes + 63) & ~size_t(63);
  Queues q;
  q.gen[0]   = pipe.queue_mem + 0 * per;
  q.gen[1]   = pipe.queue_mem + 1 * per;
  q.ext[0]   = pipe.queue_mem + 2 * per;
  q.ext[1]   = pipe.queue_mem + 3 * per;
  q.lpdf     = pipe.queue_mem + 4 * per;
  q.acc      = pipe.queue_mem + 5 * per;
  int next   = 6;
  for (int c = 0; c < kNumClasses; c++) q.shade[c] = (pipe.queue_classes >> c) & 1 ? pipe.queue_mem + (next++) * per : nullptr;
  q.park[0]  = pipe.park_mem;
  q.park[1]  = pipe.park_mem + pipe.park_words;
  q.counters = pipe.counters;
  return q;
}

static KParams make_kparams(const ygl_trace_params& params, int sample_end) {
  KParams kp;
  kp.camera = params.camera, kp.sampler = params.sampler, kp.falsecolor = params.falsecolor;
  kp.bounces = params.bounces, kp.clamp = params.clamp;
  kp.nocaustics = params.nocaustics, kp.envhidden = params.envhidden, kp.tentfilter = params.tentfilter;
  kp.sample_end = sample_end;
  kp.fuse       = 1;
  return kp;
}

// The wavefront driver: runs `nsamples` more samples on every lane of the state.
// `only_lane` >= 0: that lane alone renders sample `only_sample` (trace_sample); state->samples then stays as it is.
static int run_wavefront(ygl_context* ctx, ygl_state* state, const ygl_trace_params& params, int nsamples,
    int only_lane = -1, int only_sample = 0) {
  PathState& st = state->st;
  if (st.num_lanes == 0 || nsamples <= 0) return YGL_OK;
  const Tuning& tune       = ctx->tune;
  const int  lane_begin    = only_lane >= 0 ? only_lane : 0;
  const int  lane_end      = only_lane >= 0 ? only_lane + 1 : st.num_lanes;
  const int  sample_begin  = only_lane >= 0 ? only_sample : state->samples;
  KParams    kp            = make_kparams(params, sample_begin + nsamples);
  // Binned shade queues + one kernel per shading class: the path sampler (trace_path). The other samplers run the
  // unspecialised kernel on one queue.
  // Measured on B200 (C3, per 32 spp, final state layout): full 1080p frame binned 142 ms per 16 spp against 160; half
  // frame (1.04 M lanes) 159.7 against 184.4 ms; quarter (518 K lanes) 100.4 against 108.0; on a 1/8 tile (259 K lanes)
  // the extra launches and their tails cost what the coherence gains (70.3 vs 70.0). Hence: by lane count unless the
  // option says otherwise.
  const bool binned     = params.sampler == YGL_SAMPLER_PATH && (tune.bin >= 0 ? tune.bin != 0 : st.num_lanes > 400000);
  const unsigned classes = binned ? ctx->class_mask : 1u;
  DScene     dscene     = ctx->dscene;
  dscene.inst_class     = binned ? ctx->d_inst_class : nullptr;
  // The shading kernels can finish a path themselves (accumulate + next camera sample: no k_finish launch). Measured on
  // B200 (C3): with the unspecialised kernel that is 4-6 % faster on small tiles and 7 % slower on the full frame (the
  // extra divergence costs more than the launch); with per-class kernels it wins on the full frame too (176.6 vs 181.3 ms
  // per 16 spp).
  const int fuse = tune.fuse >= 0 ? tune.fuse : (binned || st.num_lanes <= 800000 ? 1 : 0);
  kp.fuse        = fuse;

  // Pipelines: lanes can be split over independent queue sets on separate streams (ramp-up/drain of one
  // overlapping the other). Measured on B200 (C3, 1/8 tile): 2 pipelines were 7 % slower than 1 -> default 1.
  const int npipes = std::max(1, std::min(ygl_context::kMaxPipes, tune.pipes));
  struct Run {
    ygl_context::Pipe* pipe;
    Queues             q;
    int                lo, hi, parity = 0, slot = 0;
    bool               pending[2] = {false, false}, done = false, first = true;
    LaunchCfg          light, heavy;
    cudaGraphExec_t    round_graph = nullptr;  // check_every iterations as one graph launch (see use_graph)
  } runs[ygl_context::kMaxPipes];
  for (int j = 0; j < npipes; j++) {
    Run& r = runs[j];
    r.pipe = &ctx->pipes[j];
    r.lo   = lane_begin + (int)((long long)(lane_end - lane_begin) * j / npipes);
    r.hi   = lane_begin + (int)((long long)(lane_end - lane_begin) * (j + 1) / npipes);
    if (int rc = ensure_queues(*r.pipe, r.hi - r.lo, classes, ctx->num_sms)) return rc;
    r.q = make_queues(*r.pipe);
    // persistent-style grids: a multiple of the SM count, capped by the work available
    int lanes  = r.hi - r.lo;
    int blocks = std::max(1, std::min(ctx->num_sms * 8, (lanes + 255) / 256));
    r.light = r.heavy = LaunchCfg{blocks, 256};
    r.pipe->h_done[0] = r.pipe->h_done[1] = 0;
  }
  cudaStream_t s0 = ctx->stream;

  const bool path_like = params.sampler == YGL_SAMPLER_PATH || params.sampler == YGL_SAMPLER_PATHDIRECT ||
                         params.sampler == YGL_SAMPLER_PATHMIS ||
                         params.sampler == YGL_SAMPLER_PATHTEST;  // samplers with a light-pdf stage
  int shade_launches = 1;
  if (binned) {
    shade_launches = 0;
    for (int c = 0; c < kNumClasses; c++) shade_launches += (classes >> c) & 1;
  }
  const int  per_iteration = 2 + shade_launches + (path_like ? 1 : 0) + (fuse ? 0 : 1);
  uint64_t   iterations = 0, launches = 0;
  const int  check_every = 4;
  const bool timing = ctx->time_kernels;
  unsigned long long* trav = ctx->count_traversal ? ctx->d_trav : nullptr;
  if (trav) CUDA_TRY(cudaMemsetAsync(trav, 0, 8 * sizeof(unsigned long long), s0));
  size_t ev_used = 0;
  if (timing) CUDA_TRY(cudaEventRecord(ctx->ev_loop[0], s0));
  // the other pipelines start after everything already queued on the context stream
  for (int j = 1; j < npipes; j++) {
    CUDA_TRY(cudaEventRecord(runs[j].pipe->join, s0));
    CUDA_TRY(cudaStreamWaitEvent(runs[j].pipe->stream, runs[j].pipe->join, 0));
  }
  for (int j = 0; j < npipes; j++) {
    launch_seed_lanes(runs[j].pipe->stream, runs[j].light, st, runs[j].q, 0, sample_begin, runs[j].lo, runs[j].hi);
    launches++;
  }
  // one wavefront iteration of a pipeline: extend -> shade (per class) -> light pdf -> finish
  auto enqueue_iteration = [&](Run& r, cudaStream_t s, bool with_events) -> int {
    launch_begin_iteration(s, r.q, r.parity);
    if (with_events) CUDA_TRY(cudaEventRecord(ctx->ev_pool[ev_used], s));
    launch_extend(s, ctx->num_sms, tune, dscene, st, r.q, r.parity, trav);
    if (with_events) {
      CUDA_TRY(cudaEventRecord(ctx->ev_pool[ev_used + 1], s));
      ev_used += 2;
    }
    launch_shade(s, r.heavy, dscene, st, r.q, kp, r.parity, classes);
    if (path_like) launch_lightpdf(s, r.heavy, dscene, st, r.q, kp, r.parity);
    if (!fuse) launch_finish(s, r.light, dscene, st, r.q, kp, r.parity);
    r.parity = 1 - r.parity;
    launches += per_iteration;
    return YGL_OK;
  };
  // A round of check_every iterations is the same launch sequence every time (the queue parity is back where it
  // started), so it can be captured once into a CUDA graph and re-launched as a unit (option "graph"). Not while
  // timing or counting (those need the individual launches).
  // Measured on B200 (C3, full frame and 1/8 tile): no difference either way (179.1 vs 179.0 ms, 80.8 vs 81.1 ms) - the
  // host already enqueues two rounds ahead, so the launches never were the gap. Off unless the option asks for it.
  const bool use_graph = !timing && !trav && tune.graph > 0;
  (void)extend_grid_threads(ctx->num_sms);  // occupancy query of the extend kernel: done before any capture
  auto destroy_graphs = [&]() {
    for (int j = 0; j < npipes; j++)
      if (runs[j].round_graph) cudaGraphExecDestroy(runs[j].round_graph), runs[j].round_graph = nullptr;
  };
  int remaining = npipes;
  while (remaining > 0) {
    if (ctx->stop.load(std::memory_order_relaxed)) break;  // trace_cancel: polled between rounds of iterations
    for (int j = 0; j < npipes; j++) {
      Run& r = runs[j];
      if (r.done) continue;
      cudaStream_t s = r.pipe->stream;
      if (r.first) {
        // the lanes' first camera samples; afterwards a finished path's lane starts its next sample in the
        // kernel that ends the path (end_of_path / k_finish)
        launch_generate(s, r.light, dscene, st, r.q, kp, r.parity), launches++;
        r.first = false;
      }
      if (use_graph) {
        if (!r.round_graph) {
          cudaGraph_t graph = nullptr;
          CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
          const uint64_t before = launches;
          for (int k = 0; k < check_every; k++) enqueue_iteration(r, s, false);
          launches = before;
          cudaError_t e = cudaStreamEndCapture(s, &graph);
          if (e == cudaSuccess) e = cudaGraphInstantiate(&r.round_graph, graph, 0);
          if (graph) cudaGraphDestroy(graph);
          if (e != cudaSuccess) {
            destroy_graphs();
            return fail(YGL_ERR_CUDA, std::string("CUDA graph of a wavefront round: ") + cudaGetErrorString(e));
          }
        }
        CUDA_TRY(cudaGraphLaunch(r.round_graph, s));
        launches += (uint64_t)per_iteration * check_every;
      } else {
        for (int k = 0; k < check_every; k++) {
          if (timing) {
            while (ctx->ev_pool.size() < ev_used + 2) {
              cudaEvent_t e;
              CUDA_TRY(cudaEventCreate(&e));
              ctx->ev_pool.push_back(e);
            }
          }
          if (int rc = enqueue_iteration(r, s, timing)) return rc;
        }
      }
    }
    iterations += check_every;
    for (int j = 0; j < npipes; j++) {
      Run& r = runs[j];
      if (r.done) continue;
      // wait for the check issued one round ago (keeps <= 2 rounds of launches in flight per pipeline)
      int prev = 1 - r.slot;
      if (r.pending[prev]) {
        CUDA_TRY(cudaEventSynchronize(r.pipe->ev[prev]));
        r.pending[prev] = false;
        if (r.pipe->h_done[prev] >= r.hi - r.lo) {
          r.done = true;
          remaining--;
          continue;
        }
      }
      CUDA_TRY(cudaMemcpyAsync(&r.pipe->h_done[r.slot], &r.pipe->counters->done_lanes, sizeof(int),
          cudaMemcpyDeviceToHost, r.pipe->stream));
      CUDA_TRY(cudaEventRecord(r.pipe->ev[r.slot], r.pipe->stream));
      r.pending[r.slot] = true;
      r.slot            = 1 - r.slot;
    }
  }
  // join: later work on the context stream (downloads, the tile gather) is ordered after every pipeline
  for (int j = 1; j < npipes; j++) {
    CUDA_TRY(cudaEventRecord(runs[j].pipe->join, runs[j].pipe->stream));
    CUDA_TRY(cudaStreamWaitEvent(s0, runs[j].pipe->join, 0));
  }
  if (timing) CUDA_TRY(cudaEventRecord(ctx->ev_loop[1], s0));
  for (int j = 0; j < npipes; j++) CUDA_TRY(cudaStreamSynchronize(runs[j].pipe->stream));
  destroy_graphs();
  CUDA_TRY(cudaGetLastError());
  for (int j = 0; j < npipes; j++) {
    Counters c;
    CUDA_TRY(cudaMemcpy(&c, runs[j].pipe->counters, sizeof(Counters), cudaMemcpyDeviceToHost));
    ctx->stats[0] += c.camera_samples;
    ctx->stats[1] += c.scene_rays;
    ctx->stats[2] += c.instance_rays;
  }
  ctx->stats[3] += iterations;
  ctx->stats[4] += launches;
  ctx->stats[5] += ev_used / 2;
  if (trav) {
    unsigned long long t[8];
    CUDA_TRY(cudaMemcpy(t, trav, sizeof(t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 7; k++) ctx->stats[6 + k] += t[k];
  }
  if (timing) {
    double ext_ms = 0;
    for (size_t k = 0; k + 1 < ev_used; k += 2) {
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_pool[k], ctx->ev_pool[k + 1]));
      ext_ms += ms;
    }
    float loop_ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&loop_ms, ctx->ev_loop[0], ctx->ev_loop[1]));
    ctx->timings[0] += ext_ms;
    ctx->timings[1] += loop_ms;
    ctx->timings[2] += (double)(ev_used / 2);
    ctx->timings[3] = npipes;
  }
  if (only_lane < 0) state->samples += nsamples;
  return YGL_OK;
}


static int run_render(ygl_context* ctx, ygl_state* state, const ygl_trace_params& params, int nsamples) {
  return run_wavefront(ctx, state, params, nsamples);
}

static int check_sampler(const ygl_trace_params& p) {
  switch (p.sampler) {
    case YGL_SAMPLER_PATH:
    case YGL_SAMPLER_EYELIGHT:
    case YGL_SAMPLER_NAIVE:
    case YGL_SAMPLER_FURNACE:
    case YGL_SAMPLER_PATHDIRECT:
    case YGL_SAMPLER_PATHMIS:
    case YGL_SAMPLER_PATHTEST:
    case YGL_SAMPLER_DIAGRAM:
    case YGL_SAMPLER_FALSECOLOR: return YGL_OK;
    default: return fail(YGL_ERR_RUNTIME, "sampler unknown");  // yocto_trace.cpp:1437
  }
}

int ygl_trace_samples(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, const ygl_trace_params* params) try {
  if (!ctx || !state || !params) return fail(YGL_ERR_INVALID, "null argument");
  if (state->ctx != ctx) return fail(YGL_ERR_INVALID, "state belongs to another context");
  if (int rc = check_sampler(*params)) return rc;
  if (params->camera < 0 || params->camera >= (scene ? scene->num_cameras : 0))
    return fail(YGL_ERR_INVALID, "camera id out of range");
  if (state->samples >= params->samples) return YGL_OK;  // yocto_trace.cpp:1598
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (int rc = bind_scene(ctx, scene, bvh, lights)) return rc;
  memset(ctx->stats, 0, sizeof(ctx->stats));
  memset(ctx->timings, 0, sizeof(ctx->timings));
  return run_render(ctx, state, *params, params->batch);
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_trace_samples: ") + e.what());
}

int ygl_trace_sample(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, int i, int j, int sample, const ygl_trace_params* params) try {
  if (!ctx || !state || !params) return fail(YGL_ERR_INVALID, "null argument");
  if (state->ctx != ctx) return fail(YGL_ERR_INVALID, "state belongs to another context");
  if (int rc = check_sampler(*params)) return rc;
  if (params->camera < 0 || params->camera >= (scene ? scene->num_cameras : 0))
    return fail(YGL_ERR_INVALID, "camera id out of range");
  if (i < 0 || i >= state->width || j < 0 || j >= state->height || sample < 0)
    return fail(YGL_ERR_INVALID, "pixel or sample index out of range");
  if (j < state->row_begin || j >= state->row_end || (j - state->row_begin) % state->row_step != 0)
    return fail(YGL_ERR_INVALID, "pixel row is not part of this state's tile");
  const int lane = ((j - state->row_begin) / state->row_step) * state->width + i;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (int rc = bind_scene(ctx, scene, bvh, lights)) return rc;
  memset(ctx->stats, 0, sizeof(ctx->stats));
  memset(ctx->timings, 0, sizeof(ctx->timings));
  return run_wavefront(ctx, state, *params, 1, lane, sample);
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_trace_sample: ") + e.what());
}

// ---- progressive rendering: trace_start / trace_cancel / trace_done / trace_preview (yocto_trace.cpp:1627-1676) ----
int ygl_trace_start(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, const ygl_trace_params* params) try {
  if (!ctx || !state || !params) return fail(YGL_ERR_INVALID, "null argument");
  if (state->ctx != ctx) return fail(YGL_ERR_INVALID, "state belongs to another context");
  if (ctx->worker.joinable()) {
    if (!ctx->done.load() && !ctx->stop.load()) return fail(YGL_ERR_INVALID, "a batch is already running: cancel or wait first");
    ctx->worker.join();
  }
  if (int rc = check_sampler(*params)) return rc;
  if (params->camera < 0 || params->camera >= (scene ? scene->num_cameras : 0))
    return fail(YGL_ERR_INVALID, "camera id out of range");
  if (state->samples >= params->samples) return YGL_OK;  // yocto_trace.cpp:1630
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (int rc = bind_scene(ctx, scene, bvh, lights)) return rc;
  memset(ctx->stats, 0, sizeof(ctx->stats));
  memset(ctx->timings, 0, sizeof(ctx->timings));
  ctx->stop = 0, ctx->done = 0;
  ctx->worker_rc = YGL_OK;
  ctx->worker_error.clear();
  const ygl_trace_params p = *params;
  ctx->worker = std::thread([ctx, state, p]() {
    // like the reference's worker (yocto_trace.cpp:1633-1648): one batch, abandoned early when stop is raised
    int rc;
    try {
      rc = cudaSetDevice(ctx->device) == cudaSuccess ? run_render(ctx, state, p, p.batch) : fail(YGL_ERR_CUDA, "cudaSetDevice");
    } catch (const std::exception& e) {
      rc = fail(YGL_ERR_RUNTIME, std::string("ygl_trace_start: ") + e.what());
    }
    ctx->worker_rc = rc;
    if (rc) ctx->worker_error = g_error;  // g_error is per thread: hand the message to the caller's thread
    if (!ctx->stop.load() && rc == YGL_OK) ctx->done = 1;
  });
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_trace_start: ") + e.what());
}
static int worker_join(ygl_context* ctx) {
  if (ctx->worker.joinable()) ctx->worker.join();
  if (ctx->worker_rc) return fail(ctx->worker_rc, ctx->worker_error);
  return YGL_OK;
}
int ygl_trace_cancel(ygl_context* ctx) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  ctx->stop = 1;  // polled by the host between rounds of wavefront iterations
  int rc    = worker_join(ctx);
  ctx->stop = 0;
  return rc;
}
int ygl_trace_wait(ygl_context* ctx) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  return worker_join(ctx);
}
int ygl_trace_done(ygl_context* ctx) { return ctx && ctx->done.load() ? 1 : 0; }

int ygl_trace_preview(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh, const ygl_lights* lights,
    const ygl_trace_params* params, int width, int height, float* image) try {
  if (!ctx || !scene || !params || !image) return fail(YGL_ERR_INVALID, "null argument");
  if (ctx->worker.joinable() && !ctx->done.load()) return fail(YGL_ERR_INVALID, "a batch is running: cancel or wait first");
  if (params->camera < 0 || params->camera >= scene->num_cameras) return fail(YGL_ERR_INVALID, "camera id out of range");
  if (params->pratio < 1) return fail(YGL_ERR_INVALID, "pratio must be >= 1");
  // yocto_trace.cpp:1661-1675: 1 sample at resolution / pratio, then nearest-neighbour replication to the full frame
  ygl_trace_params pp = *params;
  pp.resolution /= params->pratio;
  pp.samples = 1, pp.batch = 1;
  if (pp.resolution < 1) return fail(YGL_ERR_INVALID, "resolution / pratio must be >= 1");
  ygl_scene_desc cam_only = {};
  cam_only.num_cameras    = scene->num_cameras;
  cam_only.cameras        = scene->cameras_host.data();
  int         pw, ph;
  std::string error;
  if (!state_size(cam_only, pp, pw, ph, error)) return fail(YGL_ERR_INVALID, error);
  ygl_state* pstate = nullptr;
  if (int rc = state_create_sized(ctx, pw, ph, &pp, 0, -1, 1, &pstate)) return rc;
  int rc = ygl_trace_samples(ctx, pstate, scene, bvh, lights, &pp);
  std::vector<float> preview((size_t)pw * ph * 4);
  if (!rc) rc = ygl_state_download(pstate, preview.data(), nullptr, nullptr, nullptr, nullptr);
  std::string keep = g_error;
  ygl_state_destroy(pstate);
  if (rc) return fail(rc, keep);
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++) {
      const int pi = std::max(0, std::min(i / params->pratio, pw - 1)), pj = std::max(0, std::min(j / params->pratio, ph - 1));
      memcpy(image + ((size_t)j * width + i) * 4, preview.data() + ((size_t)pj * pw + pi) * 4, 16);
    }
  return YGL_OK;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_trace_preview: ") + e.what());
}

int ygl_trace_counters(ygl_context* ctx, uint64_t counters[16]) {
  if (!ctx || !counters) return fail(YGL_ERR_INVALID, "null argument");
  memcpy(counters, ctx->stats, sizeof(ctx->stats));
  return YGL_OK;
}
int ygl_context_set_mode(ygl_context* ctx, int mode) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  // YGL_MODE_PERSISTENT named round 1's single-kernel scheduler, which no longer exists; the value is still accepted
  // (results never depended on the mode) and selects the wavefront scheduler
  if (mode != YGL_MODE_WAVEFRONT && mode != YGL_MODE_PERSISTENT) return fail(YGL_ERR_INVALID, "unknown mode");
  ctx->mode = YGL_MODE_WAVEFRONT;
  return YGL_OK;
}

namespace {
struct OptionRef {
  const char* name;
  int Tuning::*field;
};
const OptionRef kOptions[] = {
    {"ext_blocks_per_sm", &Tuning::ext_blocks_per_sm}, {"refill", &Tuning::refill}, {"node_reps", &Tuning::node_reps},
    {"prim_weight", &Tuning::prim_weight}, {"enter_weight", &Tuning::enter_weight},
    {"suspend", &Tuning::suspend}, {"suspend_rounds", &Tuning::suspend_rounds}, {"lone", &Tuning::lone},
    {"lone_steps", &Tuning::lone_steps}, {"fuse", &Tuning::fuse}, {"bin", &Tuning::bin}, {"pipes", &Tuning::pipes},
    {"graph", &Tuning::graph}, {"top_smem", &Tuning::top_smem}, {"carveout", &Tuning::carveout},
};
}  // namespace
int ygl_context_set_option(ygl_context* ctx, const char* name, double value) {
  if (!ctx || !name) return fail(YGL_ERR_INVALID, "null argument");
  for (auto& o : kOptions)
    if (!strcmp(o.name, name)) return ctx->tune.*(o.field) = (int)value, YGL_OK;
  return fail(YGL_ERR_INVALID, std::string("unknown option: ") + name);
}
int ygl_context_get_option(ygl_context* ctx, const char* name, double* value) {
  if (!ctx || !name || !value) return fail(YGL_ERR_INVALID, "null argument");
  for (auto& o : kOptions)
    if (!strcmp(o.name, name)) return *value = (double)(ctx->tune.*(o.field)), YGL_OK;
  return fail(YGL_ERR_INVALID, std::string("unknown option: ") + name);
}

int ygl_context_set_profiling(ygl_context* ctx, int time_kernels, int count_traversal) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  ctx->time_kernels    = time_kernels != 0;
  ctx->count_traversal = count_traversal != 0;
  return YGL_OK;
}
int ygl_trace_timings(ygl_context* ctx, double ms[4]) {
  if (!ctx || !ms) return fail(YGL_ERR_INVALID, "null argument");
  memcpy(ms, ctx->timings, sizeof(ctx->timings));
  return YGL_OK;
}

int ygl_trace_image(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params, int* width,
    int* height, float* image) try {
  if (!ctx || !desc || !params) return fail(YGL_ERR_INVALID, "null argument");
  int         w, h;
  std::string error;
  if (!state_size(*desc, *params, w, h, error)) return fail(YGL_ERR_INVALID, error);
  if (width) *width = w;
  if (height) *height = h;
  if (!image) return YGL_OK;
  if (int rc = check_sampler(*params)) return rc;
  ygl_scene*  scene  = nullptr;
  ygl_bvh*    bvh    = nullptr;
  ygl_lights* lights = nullptr;
  ygl_state*  state  = nullptr;
  int         rc     = YGL_OK;
  do {
    if ((rc = ygl_bvh_build_device(ctx, desc, params->highqualitybvh, &bvh))) break;
    if ((rc = ygl_lights_create(desc, &lights))) break;
    if ((rc = ygl_scene_create(ctx, desc, &scene))) break;
    if ((rc = ygl_state_create(ctx, desc, params, &state))) break;
    if ((rc = bind_scene(ctx, scene, bvh, lights))) break;
    memset(ctx->stats, 0, sizeof(ctx->stats));
    memset(ctx->timings, 0, sizeof(ctx->timings));
    // trace_image calls trace_samples `samples` times with `batch` samples each until
    // state.samples >= params.samples (yocto_trace.cpp:1588-1590); results are batch-invariant,
    // so the same total runs as one wavefront launch sequence.
    int batch = std::max(1, params->batch);
    int total = ((params->samples + batch - 1) / batch) * batch;
    if (params->samples <= 0) total = 0;
    if ((rc = run_render(ctx, state, *params, total))) break;
    rc = ygl_state_download(state, image, nullptr, nullptr, nullptr, nullptr);
  } while (false);
  std::string keep = g_error;
  ygl_state_destroy(state);
  ygl_scene_destroy(scene);
  ygl_lights_destroy(lights);
  ygl_bvh_destroy(bvh);
  g_error = keep;
  return rc;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_trace_image: ") + e.what());
}

// ------------------------------------------------------------------------------------------
int ygl_intersect_rays_device(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh, const void* d_rays,
    int64_t n, int instance, int find_any, void* d_out, void* d_counters) {
  if (!ctx) return fail(YGL_ERR_INVALID, "null context");
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (int rc = bind_scene(ctx, scene, bvh, nullptr, /*any_lights=*/true)) return rc;
  if (instance >= scene->num_instances) return fail(YGL_ERR_INVALID, "instance id out of range");
  if (n <= 0) return YGL_OK;
  int       threads = 128;
  long long blocks  = std::min<long long>((long long)ctx->num_sms * 16, (n + threads - 1) / threads);
  launch_intersect_rays(ctx->stream, LaunchCfg{(int)std::max<long long>(blocks, 1), threads}, ctx->dscene,
      (const float4*)d_rays, n, instance, find_any, d_out, (unsigned long long*)d_counters);
  CUDA_TRY(cudaGetLastError());
  return YGL_OK;
}

int ygl_intersect_rays(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh, const ygl_ray* rays, int64_t n,
    int instance, int find_any, ygl_intersection* out) try {
  if (!ctx || !rays || !out) return fail(YGL_ERR_INVALID, "null argument");
  if (n <= 0) return YGL_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  void *d_rays = nullptr, *d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_rays, n * sizeof(ygl_ray)));
  cudaError_t e = cudaMalloc(&d_out, n * sizeof(ygl_intersection));
  if (e != cudaSuccess) {
    cudaFree(d_rays);
    return fail(YGL_ERR_CUDA, "out of device memory");
  }
  int rc = YGL_OK;
  do {
    if (cudaMemcpyAsync(d_rays, rays, n * sizeof(ygl_ray), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      rc = fail(YGL_ERR_CUDA, "H2D copy failed");
      break;
    }
    if ((rc = ygl_intersect_rays_device(ctx, scene, bvh, d_rays, n, instance, find_any, d_out, nullptr))) break;
    if (cudaMemcpyAsync(out, d_out, n * sizeof(ygl_intersection), cudaMemcpyDeviceToHost, ctx->stream) !=
            cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      rc = fail(YGL_ERR_CUDA, std::string("intersect failed: ") + cudaGetErrorString(cudaGetLastError()));
      break;
    }
  } while (false);
  cudaFree(d_rays);
  cudaFree(d_out);
  return rc;
} catch (const std::exception& e) {
  return fail(YGL_ERR_RUNTIME, std::string("ygl_intersect_rays: ") + e.what());
}

// Test hook: evaluates the device libm function `fn` (0 sin, 1 cos, 2 exp, 3 log, 4 atan, 5 acos,
// 6 atan2(x,y) with x as first argument, 7 pow(x,y), 8 sqrt, 9 fmod(x,y)) on host arrays.
// tonemap_image (yocto_image.h:242-245, yocto_image.cpp:911-922)
struct DeviceBuffer {  // cudaMalloc'd scratch that is freed on every path out of a function
  void* p = nullptr;
  ~DeviceBuffer() {
    if (p) cudaFree(p);
  }
};
static int tonemap_on_device(ygl_context* ctx, const float4* d_hdr, int64_t n, float exposure, int filmic, int srgb,
    float* ldr, uint8_t* ldr_bytes) {
  DeviceBuffer d_ldr, d_bytes;
  if (ldr) CUDA_TRY(cudaMalloc(&d_ldr.p, n * 16));
  if (ldr_bytes) CUDA_TRY(cudaMalloc(&d_bytes.p, n * 4));
  // `if (exposure != 0) rgb *= exp2(exposure)`: one libm call per image, made where the reference makes it
  const bool  scaled = exposure != 0;
  const float scale  = scaled ? std::exp2(exposure) : 1.0f;
  launch_tonemap(ctx->stream, ctx->num_sms, d_hdr, n, scale, scaled, filmic != 0, srgb != 0, (float4*)d_ldr.p, (uchar4*)d_bytes.p);
  if (ldr) CUDA_TRY(cudaMemcpyAsync(ldr, d_ldr.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (ldr_bytes) CUDA_TRY(cudaMemcpyAsync(ldr_bytes, d_bytes.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return YGL_OK;
}
int ygl_tonemap_image(ygl_context* ctx, const float* hdr, int64_t num_pixels, float exposure, int filmic, int srgb, float* ldr,
    uint8_t* ldr_bytes) {
  if (!ctx || num_pixels < 0 || (num_pixels > 0 && !hdr)) return fail(YGL_ERR_INVALID, "bad argument");
  if (!ldr && !ldr_bytes) return fail(YGL_ERR_INVALID, "no output buffer");
  if (num_pixels == 0) return YGL_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  DeviceBuffer d_hdr;
  CUDA_TRY(cudaMalloc(&d_hdr.p, num_pixels * 16));
  CUDA_TRY(cudaMemcpyAsync(d_hdr.p, hdr, num_pixels * 16, cudaMemcpyHostToDevice, ctx->stream));
  return tonemap_on_device(ctx, (const float4*)d_hdr.p, num_pixels, exposure, filmic, srgb, ldr, ldr_bytes);
}
int ygl_state_tonemap(ygl_state* state, float exposure, int filmic, int srgb, float* ldr, uint8_t* ldr_bytes) {
  if (!state) return fail(YGL_ERR_INVALID, "null state");
  if (!ldr && !ldr_bytes) return fail(YGL_ERR_INVALID, "no output buffer");
  ygl_context* ctx = state->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)state->st.num_lanes;
  if (n == 0) return YGL_OK;
  return tonemap_on_device(ctx, (const float4*)state->st.image, n, exposure, filmic, srgb, ldr, ldr_bytes);
}
int ygl_debug_libm(ygl_context* ctx, int fn, const float* x, const float* y, int64_t n, float* out) {
  if (!ctx || !x || !out || n < 0) return fail(YGL_ERR_INVALID, "bad argument");
  if (n == 0) return YGL_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  float *dx = nullptr, *dy = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc((void**)&dx, n * 4));
  CUDA_TRY(cudaMalloc((void**)&dout, n * 4));
  if (y) CUDA_TRY(cudaMalloc((void**)&dy, n * 4));
  cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (y) cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  launch_debug_libm(ctx->stream, fn, dx, dy, n, dout);
  cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dx), cudaFree(dout);
  if (dy) cudaFree(dy);
  if (e != cudaSuccess) return fail(YGL_ERR_CUDA, cudaGetErrorString(e));
  return YGL_OK;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU: one process per device; NCCL is loaded lazily (dlopen) so single-GPU users and the
// CPU-only symbol check need no NCCL at all.
struct ncclUniqueIdBlob {  // ncclUniqueId: 128 opaque bytes (nccl.h NCCL_UNIQUE_ID_BYTES)
  char internal[128];
};
typedef int (*pfn_ncclGetUniqueId)(ncclUniqueIdBlob*);
typedef int (*pfn_ncclCommInitRank)(void**, int, ncclUniqueIdBlob, int);
typedef int (*pfn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*pfn_ncclCommDestroy)(void*);
typedef const char* (*pfn_ncclGetErrorString)(int);

static void* nccl_open(std::string& error) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (auto name : names) {
    if (void* lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL)) return lib;
  }
  error = std::string("cannot load NCCL: ") + dlerror();
  return nullptr;
}

int ygl_comm_id_size(void) { return (int)sizeof(ncclUniqueIdBlob); }

int ygl_comm_create_id(void* id_blob) {
  std::string error;
  void*       lib = nccl_open(error);
  if (!lib) return fail(YGL_ERR_NCCL, error);
  auto fn = (pfn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
  if (!fn) return fail(YGL_ERR_NCCL, "ncclGetUniqueId missing");
  if (int rc = fn((ncclUniqueIdBlob*)id_blob)) return fail(YGL_ERR_NCCL, "ncclGetUniqueId failed: " + std::to_string(rc));
  return YGL_OK;
}

int ygl_comm_init(ygl_context* ctx, const void* id_blob, int rank, int nranks) {
  if (!ctx || !id_blob) return fail(YGL_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::string error;
  if (!ctx->nccl_lib) ctx->nccl_lib = nccl_open(error);
  if (!ctx->nccl_lib) return fail(YGL_ERR_NCCL, error);
  auto fn = (pfn_ncclCommInitRank)dlsym(ctx->nccl_lib, "ncclCommInitRank");
  if (!fn) return fail(YGL_ERR_NCCL, "ncclCommInitRank missing");
  ncclUniqueIdBlob id;
  memcpy(&id, id_blob, sizeof(id));
  if (int rc = fn(&ctx->nccl_comm, nranks, id, rank))
    return fail(YGL_ERR_NCCL, "ncclCommInitRank failed: " + std::to_string(rc));
  ctx->rank = rank, ctx->nranks = nranks;
  return YGL_OK;
}

void ygl_tile_rows(int height, int rank, int nranks, int* row_begin, int* row_end) {
  int per = (height + nranks - 1) / nranks;
  int b = std::min(height, rank * per), e = std::min(height, (rank + 1) * per);
  if (row_begin) *row_begin = b;
  if (row_end) *row_end = e;
}

int ygl_gather_image(ygl_context* ctx, ygl_state* state, float* image) {
  if (!ctx || !state) return fail(YGL_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int w = state->width, h = state->height;
  if (ctx->nranks == 1) {
    if (state->row_begin != 0 || state->row_end != h || state->row_step != 1)
      return fail(YGL_ERR_INVALID, "state is a partial tile but no communicator");
    return image ? ygl_state_download(state, image, nullptr, nullptr, nullptr, nullptr) : YGL_OK;
  }
  if (!ctx->nccl_comm) return fail(YGL_ERR_NCCL, "communicator not initialised");
  // two tilings: contiguous blocks of ceil(h/n) rows (ygl_tile_rows) or interleaved rows j % n == rank
  const int  n           = ctx->nranks;
  const bool interleaved = state->row_step == n && state->row_begin == std::min(ctx->rank, h) && n > 1;
  int        rb, re;
  ygl_tile_rows(h, ctx->rank, n, &rb, &re);
  if (!interleaved && !(state->row_step == 1 && rb == state->row_begin && re == state->row_end))
    return fail(YGL_ERR_INVALID, "state rows do not match this rank's tile");
  auto allgather = (pfn_ncclAllGather)dlsym(ctx->nccl_lib, "ncclAllGather");
  if (!allgather) return fail(YGL_ERR_NCCL, "ncclAllGather missing");
  // equal-sized send buffers of ceil(h / nranks) rows; short tiles are padded
  int     per   = (h + n - 1) / n;
  size_t  count = (size_t)per * w * 4;  // floats per rank
  float * d_send = nullptr, *d_recv = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_send, count * sizeof(float)));
  cudaError_t e = cudaMalloc((void**)&d_recv, count * sizeof(float) * n);
  if (e != cudaSuccess) {
    cudaFree(d_send);
    return fail(YGL_ERR_CUDA, "out of device memory");
  }
  int rc = YGL_OK;
  do {
    cudaMemsetAsync(d_send, 0, count * sizeof(float), ctx->stream);
    cudaMemcpyAsync(d_send, state->st.image, (size_t)state->st.num_lanes * 16, cudaMemcpyDeviceToDevice, ctx->stream);
    if (int nrc = allgather(d_send, d_recv, count, /*ncclFloat*/ 7, ctx->nccl_comm, ctx->stream)) {
      rc = fail(YGL_ERR_NCCL, "ncclAllGather failed: " + std::to_string(nrc));
      break;
    }
    if (image) {
      if (!interleaved) {
        cudaMemcpyAsync(image, d_recv, (size_t)w * h * 16, cudaMemcpyDeviceToHost, ctx->stream);
      } else {
        // chunk r holds rows r, r+n, r+2n, ...: de-interleave with one strided copy per rank
        const size_t rowb = (size_t)w * 16;
        for (int r = 0; r < n && r < h; r++) {
          int nrows = (h - r + n - 1) / n;
          cudaMemcpy2DAsync((char*)image + r * rowb, n * rowb, (const char*)d_recv + (size_t)r * per * rowb, rowb, rowb,
              nrows, cudaMemcpyDeviceToHost, ctx->stream);
        }
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(YGL_ERR_CUDA, "gather failed");
  } while (false);
  cudaFree(d_send);
  cudaFree(d_recv);
  return rc;
}

void ygl_comm_destroy(ygl_context* ctx) {
  if (!ctx || !ctx->nccl_comm) return;
  auto fn = (pfn_ncclCommDestroy)dlsym(ctx->nccl_lib, "ncclCommDestroy");
  if (fn) fn(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks    = 1;
  ctx->rank      = 0;
}

}  // extern "C"
