// ygl_kernels.cuh — launch-side declarations shared by ygl_kernels.cu (device) and ygl_api.cpp (host).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ygl_scene.cuh"

namespace ygl {

// Lane state is written by one kernel and read by the next: plain loads and stores. (COH = true turns every state
// load into ld.global.cg; it served round 1's single-kernel scheduler, where reader and writer could sit on different
// SMs inside one kernel, and is kept only as a debugging aid.)
template <class T, bool COH>
struct SRef {
  T* p;
#ifdef __CUDACC__
  __device__ __forceinline__ T load() const { return COH ? __ldcg(p) : *p; }
  __device__ __forceinline__ operator T() const { return load(); }
  __device__ __forceinline__ const SRef& operator=(const T& v) const {
    *p = v;
    return *this;
  }
  __device__ __forceinline__ const SRef& operator=(const SRef& o) const {
    *p = o.load();
    return *this;
  }
  __device__ __forceinline__ const SRef& operator+=(const T& v) const {
    *p = load() + v;
    return *this;
  }
#endif
};
// STRIDE (in elements): > 1 for the path-state records that are stored interleaved per lane (see PathStateT).
// Strided arrays do NOT decay to their pointer: two arrays of different stride in a ?: would both decay and then be
// indexed as contiguous (that happened once). Contiguous ones (the specialisation below) do: memcpy, memset.
template <class T, bool COH, int STRIDE = 1>
struct SArr {
  T* p;
#ifdef __CUDACC__
  __device__ __forceinline__ SRef<T, COH> operator[](long long i) const { return SRef<T, COH>{p + i * STRIDE}; }
#endif
  __host__ __device__ explicit operator T*() const { return p; }
  __host__ __device__ SArr& operator=(T* q) {
    p = q;
    return *this;
  }
};
template <class T, bool COH>
struct SArr<T, COH, 1> {
  T* p;
#ifdef __CUDACC__
  __device__ __forceinline__ SRef<T, COH> operator[](long long i) const { return SRef<T, COH>{p + i}; }
#endif
  __host__ __device__ operator T*() const { return p; }
  __host__ __device__ SArr& operator=(T* q) {
    p = q;
    return *this;
  }
};

// Per-lane wavefront state: 16-byte records, moved as whole 128-bit words. Lanes reach a stage in queue order, i.e.
// scattered, so every 16-byte access costs a 32-byte DRAM sector; records that a stage touches together are therefore
// stored interleaved, YGL_STATE_GROUP of them per lane: neighbours share a 32-byte sector (ray_o|ray_d, radiance|weight,
// hit_uvd|pend, ...) and the hot eight share one 128-byte line. (Packing hit_ids / sample / susp into a ninth record
// next to hit_uvd was measured too: +3.8 % on the C3 frame, dropped.)
// One lane = one pixel of the tile (the only legal parallel axis: each pixel's samples form a
// sequential chain through its rng stream and running mean, yocto_trace.cpp:1461-1492).
#ifndef YGL_STATE_GROUP
#define YGL_STATE_GROUP 8  // float4 records interleaved per lane. Measured on B200 (C3 frame): pairs -4 % against separate arrays,
                           // quads another -2.8 %, eight (one 128-byte line per lane) another -1.1 %
#endif
template <bool COH>
using SRec = SArr<float4, COH, YGL_STATE_GROUP>;  // one member of an interleaved group of float4 records
template <bool COH>
struct PathStateT {
  int num_lanes;  // pixels in this tile
  int width, height, row_begin, row_step;  // lane l <-> pixel (l % width, row_begin + (l / width) * row_step)
  // ---- trace_state accumulators (reference layout, yocto_trace.h:147-157) ----
  SArr<float4, COH>     image;   // vec4f
  SArr<float, COH>      albedo;  // vec3f packed
  SArr<float, COH>      normal;  // vec3f packed
  SArr<int, COH>        hits;
  SArr<ulonglong2, COH> rngs;  // rng_state {state, inc}
  // ---- per-lane progress ----
  SArr<int, COH> sample;  // index of the sample in flight
  // ---- path in flight ----
  SRec<COH>        ray_o;     // origin.xyz, w: bounce (int bits)
  SRec<COH>        ray_d;     // direction.xyz, w: opbounce (int bits)
  SRec<COH>        radiance;  // rgb, w: flags (int bits): 1 hit, 2 in-volume slot occupied
  SRec<COH>        weight;    // rgb, w: max_roughness
  SRec<COH>        hit_uvd;   // uv.x uv.y distance hit(int bits)
  SArr<int2, COH>   hit_ids;   // instance, element
  SRec<COH>        albedo0;   // bounce-0 albedo rgb, w unused
  SRec<COH>        normal0;   // bounce-0 normal (or -camera dir on miss) xyz
  SRec<COH>        vol_a;     // volume slot: density.xyz, scanisotropy
  SRec<COH>        vol_b;     // volume slot: scattering.xyz
  SRec<COH>        pend;      // pending MIS numerator: bsdfcos.rgb, w: bsdf/phase pdf
  SArr<int, COH>    susp;      // save slot of the lane's ray while it is parked by k_extend (see Queues::park)
  // ---- pathdirect / pathmis only: the extra shadow-ray stage of a bounce ----
  SRec<COH>        aux_o;     // shading position (shadow-ray origin), w: bsdf pdf of the pending direct sample
  SRec<COH>        aux_dir;   // direct-sample direction, w: light pdf (pathdirect) / mis weight (pathmis)
  SRec<COH>        aux_bsdf;  // bsdfcos of the direct sample, w: phase (int bits)
  SRec<COH>        aux_uvd;   // shadow-ray hit: uv.x uv.y distance hit
  SArr<int2, COH>   aux_ids;   // shadow-ray hit: instance, element
  SRec<COH>        next_uvd;  // pathmis next_intersection (persists across bounces)
  SArr<int2, COH>   next_ids;
};
using PathState = PathStateT<false>;  // a kernel boundary separates every writer of lane state from its readers

// Shading classes: the extend kernel appends every finished ray to the shade queue of its class, so that a shading
// warp holds ONE kind of work (one material type, or misses) and each class gets a kernel that contains only its
// own code. Class 0 is the unspecialised kernel: every lane when binning is off (all samplers but `path`), and,
// with binning on, the lanes that travel inside a participating medium.
constexpr int kClsGeneric = 0;  // 1..8 = 1 + material_type (yocto_scene.h:111-120)
constexpr int kClsMiss    = 9;
constexpr int kNumClasses = 10;

// Work queues: lane ids compacted with warp-ballot / one atomic per warp.
struct Queues {
  int* gen[2];              // lanes that start a new sample
  int* ext[2];              // lanes with a ray to trace
  int* shade[kNumClasses];  // lanes whose ray has been traced, by shading class (null: class absent from the scene)
  int* lpdf;                // lanes waiting for sample_lights_pdf
  int* acc;                 // lanes whose sample finished
  int* park[2];             // save slots of rays parked by k_extend: [launch parity][thread of its grid][kSuspendWords]
  // counters (device): see QueueCounters
  struct Counters* counters;
};

struct Counters {
  int n_gen[2], n_ext[2], n_lpdf, n_acc;
  int done_lanes;  // lanes that finished all their samples
  int ext_head;    // work cursor of the persistent extend kernel (reset every iteration)
  int n_shade[kNumClasses];
  unsigned long long camera_samples, scene_rays, instance_rays, shade_calls;
};

struct KParams {  // trace_params subset used on device (yocto_trace.h:95-113)
  int   camera, sampler, falsecolor, bounces;
  float clamp;
  int   nocaustics, envhidden, tentfilter;
  int   sample_end;  // lanes stop when sample index reaches this
  int   fuse;        // shading kernels finish a path themselves (accumulate + next camera sample); 0 = separate kernels
};

struct LaunchCfg {
  int blocks, threads;
};

// Scheduling knobs of a context (ygl_context_set_option). None of them can change a result bit: they choose grid
// sizes, the tail strategy of the extend kernel, which kernels finish a path, and the scheduler's own parameters.
struct Tuning {
  int ext_blocks_per_sm = 0;  // extend grid: blocks per SM (0 = the occupancy maximum)
  int refill            = 8;  // extend: idle lanes of a warp that trigger a refill
  int node_reps         = 3;  // extend: node visits per scheduling round
  int prim_weight       = 12;  // extend: weight (in eighths) of the lanes waiting for primitive tests in the path vote
  int enter_weight      = 12; //   ... and of the lanes waiting to enter an instance (8 = plain majority; measured on B200: 12 / 12 wins 2 %)
  int suspend           = -1; // extend tail: park a drained warp's stragglers once <= this many lanes are busy (-1 = by tile size, 0 = never)
  int suspend_rounds    = 96; //   ... after at least this many rounds
  int lone              = -1; // extend tail: vote-free walk of the last <= this many lanes (-1 = by tile size, 0 = never)
  int lone_steps        = 0;  //   ... parked after this many steps (0 = never)
  int fuse              = -1; // shading kernels finish paths themselves (1) or hand them to k_finish (0); -1 = by tile size
  int bin               = -1; // class-binned shade queues + per-class kernels for the path sampler (1/0; -1 = by tile size)
  int pipes             = 1;  // independent wavefront pipelines (streams) per state
  int graph             = -1; // submit rounds of iterations as a CUDA graph (1 = on; measured: no gain, so -1 = off)
  int carveout          = -1; // extend: preferred shared-memory carve-out in percent (-1 = leave the driver's choice)
  int top_smem          = -1; // extend: stage the instance-level tree in shared memory with a bulk async copy (1 = on; measured slower, so -1 = off)
};

void launch_begin_iteration(cudaStream_t s, Queues q, int parity);
void launch_seed_lanes(cudaStream_t s, LaunchCfg cfg, PathState st, Queues q, int parity, int sample_begin,
    int lane_lo, int lane_hi);
void launch_generate(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity);
// trav: device array of 7 counters {top nodes, bottom nodes, instance visits, tri, quad, line, point tests} or null.
// Finished rays go to q.shade[class] (class 0 for every lane unless scene.inst_class is set).
int  extend_grid_threads(int num_sms);  // threads of the largest extend grid (sizes Queues::park)
void launch_extend(cudaStream_t s, int num_sms, const Tuning& tune, DScene scene, PathState st, Queues q, int parity,
    unsigned long long* trav);
// one launch per shading class present (class_mask bit c), or the single generic kernel when binning is off
void launch_shade(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity,
    unsigned class_mask);
void launch_lightpdf(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity);
// accumulate the finished paths of q.acc and start each lane's next camera sample (used when p.fuse == 0)
void launch_finish(cudaStream_t s, LaunchCfg cfg, DScene scene, PathState st, Queues q, KParams p, int parity);

// batch intersect (test hook + traversal micro-benchmark). counters may be null.
void launch_intersect_rays(cudaStream_t s, LaunchCfg cfg, DScene scene, const float4* rays, long long n, int instance,
    int find_any, void* out, unsigned long long* counters);

void launch_debug_libm(cudaStream_t s, int fn, const float* x, const float* y, long long n, float* out);
// tonemap_image (yocto_image.cpp:911-922): scale = exp2(exposure) from the host; ldr and / or ldr_bytes may be null
void launch_tonemap(cudaStream_t s, int num_sms, const float4* hdr, long long n, float scale, bool scaled, bool filmic, bool srgb,
    float4* ldr, uchar4* ldr_bytes);

}  // namespace ygl
