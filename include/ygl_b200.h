/*
 * ygl_b200.h — C ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary for Yocto/GL's renderer entry points. Every
 * function below names the reference interface it stands in for
 * (paths relative to the reference checkout, libs/yocto/...):
 *
 *   trace_image            yocto_trace.h:116   / yocto_trace.cpp:1584
 *   make_trace_bvh         yocto_trace.h:171   / yocto_trace.cpp:88   (make_scene_bvh, yocto_bvh.cpp:364)
 *   make_trace_lights      yocto_trace.h:167   / yocto_trace.cpp:1528
 *   make_trace_state       yocto_trace.h:163   / yocto_trace.cpp:1495
 *   trace_samples          yocto_trace.h:174   / yocto_trace.cpp:1595
 *   get_image & friends    yocto_trace.h:182+  / yocto_trace.cpp:1694
 *   intersect_scene_bvh    yocto_bvh.h:108     / yocto_bvh.cpp:554
 *   intersect_instance_bvh yocto_bvh.h:110     / yocto_bvh.cpp:619
 *
 * The mirrored-backend precedent in the reference is yocto_cutrace.h:71-146.
 *
 * Conventions: plain C, POD structs, pointers + counts, no torch/C++ types.
 * Every call returns YGL_OK (0) or a negative status; ygl_last_error() gives
 * the message of the last failure on the calling thread. Handles are opaque.
 * All host arrays passed in are only read during the call (value semantics,
 * like the reference's const& inputs). There is no CPU fallback: calls that
 * need the device fail with YGL_ERR_CUDA when no B200 is usable.
 */
#ifndef YGL_B200_H
#define YGL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YGL_OK 0
#define YGL_ERR_INVALID (-1) /* bad argument (std::invalid_argument in the reference) */
#define YGL_ERR_CUDA (-2)    /* CUDA runtime failure / no device */
#define YGL_ERR_RUNTIME (-3) /* std::runtime_error in the reference ("sampler unknown") */
#define YGL_ERR_NCCL (-4)

#define YGL_INVALID_ID (-1)

/* ---- scene description: flat views of yocto::scene_data (yocto_scene.h:83-213) ---- */

/* frame3f = {x, y, z, o} column vectors, 12 floats (yocto_math.h frame3f) */
typedef struct ygl_frame3f {
  float x[3], y[3], z[3], o[3];
} ygl_frame3f;

/* camera_data, yocto_scene.h:83-91 (72 B) */
typedef struct ygl_camera {
  ygl_frame3f frame;
  int32_t     orthographic;
  float       lens, film, aspect, focus, aperture;
} ygl_camera;

/* material_type, yocto_scene.h:107-112 */
enum {
  YGL_MATERIAL_MATTE = 0,
  YGL_MATERIAL_GLOSSY,
  YGL_MATERIAL_REFLECTIVE,
  YGL_MATERIAL_TRANSPARENT,
  YGL_MATERIAL_REFRACTIVE,
  YGL_MATERIAL_SUBSURFACE,
  YGL_MATERIAL_VOLUMETRIC,
  YGL_MATERIAL_GLTFPBR
};

/* material_data, yocto_scene.h:123-142 (84 B) */
typedef struct ygl_material {
  int32_t type;
  float   emission[3];
  float   color[3];
  float   roughness, metallic, ior;
  float   scattering[3];
  float   scanisotropy, trdepth, opacity;
  int32_t emission_tex, color_tex, roughness_tex, scattering_tex, normal_tex;
} ygl_material;

/* instance_data, yocto_scene.h:145-150 (56 B) */
typedef struct ygl_instance {
  ygl_frame3f frame;
  int32_t     shape, material;
} ygl_instance;

/* environment_data, yocto_scene.h:153-158 (64 B) */
typedef struct ygl_environment {
  ygl_frame3f frame;
  float       emission[3];
  int32_t     emission_tex;
} ygl_environment;

/* texture_data, yocto_scene.h:95-103. Exactly one of pixelsf / pixelsb is non-null. */
typedef struct ygl_texture {
  int32_t        width, height;
  int32_t        linear, nearest, clamp;
  const float*   pixelsf; /* width*height vec4f, or NULL */
  const uint8_t* pixelsb; /* width*height vec4b, or NULL */
} ygl_texture;

/* shape_data, yocto_shape.h:74-88 (tangents are not read on the hot path) */
typedef struct ygl_shape {
  int32_t        num_points, num_lines, num_triangles, num_quads;
  const int32_t* points;    /* num_points   x 1 */
  const int32_t* lines;     /* num_lines    x 2 */
  const int32_t* triangles; /* num_triangles x 3 */
  const int32_t* quads;     /* num_quads    x 4 */
  int32_t        num_positions, num_normals, num_texcoords, num_colors, num_radius;
  const float*   positions; /* x3 */
  const float*   normals;   /* x3 */
  const float*   texcoords; /* x2 */
  const float*   colors;    /* x4 */
  const float*   radius;    /* x1 */
} ygl_shape;

typedef struct ygl_scene_desc {
  int32_t                num_cameras, num_instances, num_environments, num_shapes, num_textures,
      num_materials;
  const ygl_camera*      cameras;
  const ygl_instance*    instances;
  const ygl_environment* environments;
  const ygl_shape*       shapes;
  const ygl_texture*     textures;
  const ygl_material*    materials;
} ygl_scene_desc;

/* trace_sampler_type, yocto_trace.h:71-81 */
enum {
  YGL_SAMPLER_PATH = 0,
  YGL_SAMPLER_PATHDIRECT,
  YGL_SAMPLER_PATHMIS,
  YGL_SAMPLER_PATHTEST,
  YGL_SAMPLER_NAIVE,
  YGL_SAMPLER_EYELIGHT,
  YGL_SAMPLER_DIAGRAM,
  YGL_SAMPLER_FURNACE,
  YGL_SAMPLER_FALSECOLOR
};

/* trace_falsecolor_type, yocto_trace.h:83-89 */
enum {
  YGL_FALSECOLOR_POSITION = 0, YGL_FALSECOLOR_NORMAL, YGL_FALSECOLOR_FRONTFACING,
  YGL_FALSECOLOR_GNORMAL, YGL_FALSECOLOR_GFRONTFACING, YGL_FALSECOLOR_TEXCOORD,
  YGL_FALSECOLOR_MTYPE, YGL_FALSECOLOR_COLOR, YGL_FALSECOLOR_EMISSION,
  YGL_FALSECOLOR_ROUGHNESS, YGL_FALSECOLOR_OPACITY, YGL_FALSECOLOR_METALLIC,
  YGL_FALSECOLOR_DELTA, YGL_FALSECOLOR_INSTANCE, YGL_FALSECOLOR_SHAPE,
  YGL_FALSECOLOR_MATERIAL, YGL_FALSECOLOR_ELEMENT, YGL_FALSECOLOR_HIGHLIGHT
};

#define YGL_DEFAULT_SEED 961748941ull /* trace_default_seed, yocto_trace.h:92 */

/* trace_params, yocto_trace.h:95-113. embreebvh / noparallel / pratio / denoise are accepted
 * and ignored (no Embree, the grid replaces the thread pool, no GUI preview, no OIDN). */
typedef struct ygl_trace_params {
  int32_t  camera;
  int32_t  resolution;
  int32_t  sampler;
  int32_t  falsecolor;
  int32_t  samples;
  int32_t  bounces;
  float    clamp;
  int32_t  nocaustics, envhidden, tentfilter;
  uint64_t seed;
  int32_t  embreebvh, highqualitybvh, noparallel;
  int32_t  pratio;
  int32_t  denoise;
  int32_t  batch;
} ygl_trace_params;

/* Fills the defaults of yocto_trace.h:95-113. */
void ygl_trace_params_default(ygl_trace_params* params);

/* ray3f, yocto_geometry.h:135-140 (32 B) */
typedef struct ygl_ray {
  float o[3], d[3], tmin, tmax;
} ygl_ray;

/* scene_intersection, yocto_bvh.h:97-103 (24 B; hit widened to int32) */
typedef struct ygl_intersection {
  int32_t instance, element;
  float   uv[2];
  float   distance;
  int32_t hit;
} ygl_intersection;

/* bvh_node, yocto_shape.h:474-480 (32 B, same field order and widths) */
typedef struct ygl_bvh_node {
  float   bbox_min[3], bbox_max[3];
  int32_t start;
  int16_t num;
  int8_t  axis;
  uint8_t internal;
} ygl_bvh_node;

typedef struct ygl_context ygl_context; /* one CUDA device + stream + scratch queues */
typedef struct ygl_scene   ygl_scene;   /* device-resident scene arena */
typedef struct ygl_bvh     ygl_bvh;     /* two-level BVH (host trees + device packets) */
typedef struct ygl_lights  ygl_lights;  /* trace_lights */
typedef struct ygl_state   ygl_state;   /* trace_state */

const char* ygl_last_error(void);
const char* ygl_version(void);

/* ---- context ---- */
int  ygl_context_create(int device, ygl_context** out);
void ygl_context_destroy(ygl_context* ctx);
int  ygl_context_synchronize(ygl_context* ctx);
/* cudaStream_t of the context (for CUDA-event timing by the caller), as void* */
void* ygl_context_stream(ygl_context* ctx);

/* ---- scene: replaces make_cutrace_scene-style upload; input = scene_data views ---- */
int  ygl_scene_create(ygl_context* ctx, const ygl_scene_desc* desc, ygl_scene** out);
int  ygl_scene_update_cameras(ygl_scene* scene, const ygl_camera* cameras, int num_cameras);
void ygl_scene_destroy(ygl_scene* scene);

/* ---- scene ingestion: load_scene (yocto_sceneio.h:93, yocto_sceneio.cpp:2761-2782), read into flat host arrays.
 * The extension picks the format: ".json" = Yocto/GL scenes of asset version 4.2 / 5.0 (load_json_scene, :3618-3860)
 * and of the older format 4.0 that carries no version (load_json_scene_version40, :3025-3372: groups keyed by name,
 * shapes / textures created by mention and found by trying extensions, "objects" with instance lists); ".ply" = one
 * shape as a scene with the default material, a framing camera and add_sky's procedural sky (load_ply_scene,
 * :4364-4381); ".gltf" / ".glb" = glTF 2.0 (load_gltf_scene, :4430-4767 over cgltf: meshes, node transforms, cameras,
 * pbr / transmission materials, external, base64 and GLB buffers; the sky is added as for .ply); ".obj" = Wavefront OBJ
 * with its .mtl libraries and the reference's .obx side file (load_obj_scene, :4111-4243). Shapes are .ply, .obj (load_shape's de-duplicated reading) or binary .stl, textures .png, .jpg (baseline
 * and progressive, decoded exactly as stb_image does), Radiance .hdr or OpenEXR (scanline; none / RLE / ZIPS / ZIP / PIZ).
 * The returned object owns the arrays its ygl_scene_desc views; hand that desc to ygl_scene_create / ygl_bvh_build /
 * ygl_lights_create / ygl_state_create. Defaults, lookat frames, texcoord flip, polygon fans, missing camera and
 * missing radius follow the reference (:1008-1035, :1796-1837, :2119-2148). Subdivs (.obj control meshes) are loaded
 * and tesselated right away (tesselate_subdivs, yocto_scene.cpp:739-813: what every reference app does after
 * load_scene). Format 4.1 (the reference's loader for it always fails, :3614), pbrt / mitsuba / stl
 * scenes, ascii stl (the reference cannot read it either) and tiled / multi-part EXR textures are refused with an error. ---- */
typedef struct ygl_loaded_scene ygl_loaded_scene;
int  ygl_scene_load(const char* filename, ygl_loaded_scene** out);
const ygl_scene_desc* ygl_loaded_scene_desc(const ygl_loaded_scene* scene);
/* names kept from the file; kind: 0 camera, 1 texture, 2 material, 3 shape, 4 instance, 5 environment */
const char* ygl_loaded_scene_name(const ygl_loaded_scene* scene, int kind, int index);
void ygl_loaded_scene_destroy(ygl_loaded_scene* scene);

/* ---- bvh: make_trace_bvh / make_scene_bvh (host build in the reference's node order) ---- */
/* Pure host work; the device copy is made on first use with a context. */
int  ygl_bvh_build(const ygl_scene_desc* desc, int highquality, ygl_bvh** out);
/* The same trees, bit for bit, with the large ones (>= 4096 primitives, default split) built on the context's device:
 * level-parallel split_middle that reproduces the reference's node indices and std::partition's permutation
 * (ygl_bvh_device.cu). highquality (SAH) trees and small trees are built on the host cores. ygl_trace_image uses it. */
int  ygl_bvh_build_device(ygl_context* ctx, const ygl_scene_desc* desc, int highquality, ygl_bvh** out);
/* Number of nodes / primitives of tree `shape` (>=0) or of the instance tree (shape == -1). */
int  ygl_bvh_tree_size(const ygl_bvh* bvh, int shape, int* num_nodes, int* num_primitives);
/* Copies the tree out (reference bvh_tree layout) — used by the parity tests. */
int  ygl_bvh_tree_get(const ygl_bvh* bvh, int shape, ygl_bvh_node* nodes, int32_t* primitives);
/* Parity mode: adopt trees built elsewhere — e.g. the reference's own make_scene_bvh output (scene_bvh,
 * yocto_bvh.h:70-79) — verbatim: top_nodes / top_primitives = scene_bvh::bvh, shape_nodes[s] / shape_primitives[s] =
 * scene_bvh::shapes[s].bvh (bvh_node layout of yocto_shape.h:474-480, which ygl_bvh_node mirrors). The trees are
 * checked (index ranges, leaves of <= 4 primitives, depth <= 128) and only the traversal packets are derived. */
int  ygl_bvh_create_from_host(const ygl_scene_desc* desc, const ygl_bvh_node* top_nodes, int num_top_nodes,
     const int32_t* top_primitives, int num_top_primitives, const ygl_bvh_node* const* shape_nodes,
     const int* shape_num_nodes, const int32_t* const* shape_primitives, const int* shape_num_primitives,
     ygl_bvh** out);
/* update_scene_bvh (yocto_bvh.h:94-96, yocto_bvh.cpp:434-451): refit after vertex positions / radii of the listed shapes
 * or any instance frame changed - same topology, boxes recomputed bottom-up in the reference's merge order, so the
 * trees equal the reference's refit bit for bit. `desc` is the edited scene (same element counts as at build time).
 * Like the reference, every instance box is recomputed, whatever updated_instances lists. The device copy is replaced
 * on the next call that uses the bvh; not to be called while a trace_start batch is running on it. */
int  ygl_bvh_update(ygl_bvh* bvh, const ygl_scene_desc* desc, const int* updated_instances, int num_updated_instances,
     const int* updated_shapes, int num_updated_shapes);
void ygl_bvh_destroy(ygl_bvh* bvh);

/* ---- lights: make_trace_lights (host; CDFs in the reference's summation order) ---- */
int  ygl_lights_create(const ygl_scene_desc* desc, ygl_lights** out);
int  ygl_lights_count(const ygl_lights* lights);
/* light i: instance id, environment id, cdf length; cdf copied if non-null */
int  ygl_lights_get(const ygl_lights* lights, int i, int* instance, int* environment, int* cdf_size,
     float* cdf);
void ygl_lights_destroy(ygl_lights* lights);

/* ---- state: make_trace_state; resumable accumulator ---- */
int  ygl_state_create(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params,
     ygl_state** out);
/* Restrict the state to image rows [row_begin,row_end) (tile partition for multi-GPU). The rng
 * table is still seeded for the full image so values equal the single-device run. */
int  ygl_state_create_tile(ygl_context* ctx, const ygl_scene_desc* desc,
     const ygl_trace_params* params, int row_begin, int row_end, ygl_state** out);
/* Interleaved tile for multi-GPU load balance: image rows rank, rank + nranks, rank + 2 nranks, ...
 * (sky and geometry rows are spread evenly over the ranks; values are unaffected). */
int  ygl_state_create_interleaved(ygl_context* ctx, const ygl_scene_desc* desc,
     const ygl_trace_params* params, int rank, int nranks, ygl_state** out);
int  ygl_state_size(const ygl_state* state, int* width, int* height, int* samples);
/* first row, row stride and number of rows held by this state */
int  ygl_state_layout(const ygl_state* state, int* row_first, int* row_step, int* num_rows);
int  ygl_state_rows(const ygl_state* state, int* row_begin, int* row_end);
/* Any pointer may be NULL. image: w*h*4 floats, albedo/normal: w*h*3, hits: w*h ints,
 * rngs: w*h*2 uint64 {state, inc} (rows of this tile only). */
int  ygl_state_download(ygl_state* state, float* image, float* albedo, float* normal, int32_t* hits,
     uint64_t* rngs);
int  ygl_state_upload(ygl_state* state, int samples, const float* image, const float* albedo,
     const float* normal, const int32_t* hits, const uint64_t* rngs);
/* reset_cutrace_state-style (yocto_cutrace.h:119): zero the accumulators, samples = 0, re-seed the per-pixel rng
 * streams from params->seed exactly as make_trace_state does (yocto_trace.cpp:1512-1515). */
int  ygl_state_reset(ygl_state* state, const ygl_trace_params* params);
void ygl_state_destroy(ygl_state* state);
/* Host-only helper (no device): image size and the per-pixel rng table of make_trace_state. */
int  ygl_make_state_rngs(const ygl_scene_desc* desc, const ygl_trace_params* params, int* width,
     int* height, uint64_t* rngs /* may be NULL; else w*h*2 */);

/* ---- rendering ---- */
/* trace_samples: advances every pixel of the state by params->batch samples (no-op once
 * state.samples >= params->samples). Like the reference it returns when the batch is complete;
 * results stay on the device until ygl_state_download / ygl_gather_image. */
int ygl_trace_samples(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, const ygl_trace_params* params);
/* trace_sample (yocto_trace.h:173-175): ONE sample `sample` of pixel (i, j) of the full image — camera sample,
 * sampler, running-mean accumulation with weight 1 / (sample + 1) — on the pixel's own rng stream. state.samples is
 * not changed (the reference's trace_sample does not change it either). The pixel must belong to the state's tile. */
int ygl_trace_sample(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, int i, int j, int sample, const ygl_trace_params* params);
/* ---- progressive rendering: trace_start / trace_cancel / trace_done / trace_preview (yocto_trace.h:202-223,
 * yocto_trace.cpp:1627-1676) ----
 * ygl_trace_start returns at once; ONE batch (params->batch samples per pixel) renders on a worker thread of the
 * context. Until ygl_trace_done() reports 1 (or ygl_trace_cancel / ygl_trace_wait returned) the context, the state and
 * the scene objects belong to the worker — the reference's rule (exclusive access to `state`). ygl_trace_cancel raises
 * the stop flag — polled by the host between rounds of wavefront iterations — and joins the worker; as in the reference, a cancelled batch leaves the pixels at mixed sample counts
 * (state.samples is advanced by the batch all the same, yocto_trace.cpp:1641) and the caller resets the state.
 * ygl_trace_wait joins without cancelling and returns the batch's status. */
int ygl_trace_start(ygl_context* ctx, ygl_state* state, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_lights* lights, const ygl_trace_params* params);
int ygl_trace_cancel(ygl_context* ctx);
int ygl_trace_wait(ygl_context* ctx);
int ygl_trace_done(ygl_context* ctx); /* 1 once the started batch has completed, else 0 */
/* trace_preview: 1 sample at params->resolution / params->pratio, replicated (nearest) into the full-size rgba image
 * of width x height pixels, exactly like yocto_trace.cpp:1657-1676. */
int ygl_trace_preview(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh, const ygl_lights* lights,
    const ygl_trace_params* params, int width, int height, float* image);
/* trace_image: bvh + lights + state + all samples + get_image, from host scene views to a host
 * rgba float image (image may be NULL to query the size). */
int ygl_trace_image(ygl_context* ctx, const ygl_scene_desc* desc, const ygl_trace_params* params,
    int* width, int* height, float* image);
/* Work counters of the last ygl_trace_samples / ygl_trace_image on this context:
 * [0] camera samples, [1] scene rays (intersect_scene calls), [2] instance rays (intersect_instance
 * calls), [3] wavefront iterations, [4] kernel launches, [5] extend-kernel launches; with traversal
 * counting enabled (ygl_context_set_profiling) also, for the scene rays of the extend kernel:
 * [6] instance-tree nodes popped, [7] shape-tree nodes popped, [8] instance visits,
 * [9] triangle, [10] quad, [11] line, [12] point tests (SURVEY.md 8d algorithmic-bytes inputs). */
int ygl_trace_counters(ygl_context* ctx, uint64_t counters[16]);
/* Measurement hooks. time_kernels != 0: bracket every extend-kernel launch and the whole sample loop
 * with CUDA events on the context stream. count_traversal != 0: run the counting variant of the
 * extend kernel (slower; never combine with timing you report). */
int ygl_context_set_profiling(ygl_context* ctx, int time_kernels, int count_traversal);
/* Round 1 had a second scheduler (one resident kernel, YGL_MODE_PERSISTENT). It was slower than the wavefront on every
 * tile size and timing-dependent, and has been removed; the call and both values remain for source compatibility and
 * always select the wavefront scheduler (results never depended on the mode). */
enum { YGL_MODE_WAVEFRONT = 0, YGL_MODE_PERSISTENT = 1 };
int ygl_context_set_mode(ygl_context* ctx, int mode);
/* Scheduling knobs of a context, by name (the library reads no environment variables). None can change a result
 * bit. Names: "ext_blocks_per_sm", "refill", "node_reps", "prim_weight", "enter_weight", "suspend", "suspend_rounds", "lone", "lone_steps", "fuse",
 * "bin", "pipes", "graph", "top_smem", "carveout" (struct Tuning in ygl_kernels.cuh documents each). -1 restores the automatic choice where one exists. */
int ygl_context_set_option(ygl_context* ctx, const char* name, double value);
int ygl_context_get_option(ygl_context* ctx, const char* name, double* value);
/* Timings of the last ygl_trace_samples / ygl_trace_image with time_kernels on:
 * [0] sum of extend-kernel launch durations (ms), [1] whole sample loop (ms), [2] extend launches. */
int ygl_trace_timings(ygl_context* ctx, double ms[4]);

/* Batch form of intersect_scene_bvh (instance < 0) / intersect_instance_bvh (instance >= 0).
 * rays/out are HOST arrays of n elements. */
int ygl_intersect_rays(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh,
    const ygl_ray* rays, int64_t n, int instance, int find_any, ygl_intersection* out);
/* Same with rays/out already resident on the device (used by bench.py's kernel-only timing);
 * asynchronous on the context stream. counters (device, may be NULL): 8 x uint64 traversal
 * statistics {top nodes, bottom nodes, instances, primitives, hits, ...}. */
int ygl_intersect_rays_device(ygl_context* ctx, const ygl_scene* scene, const ygl_bvh* bvh,
    const void* d_rays, int64_t n, int instance, int find_any, void* d_out, void* d_counters);

/* ---- post-process: tonemap_image (yocto_image.h:242-245, yocto_image.cpp:911-922; tonemap yocto_color.h:356-366) ----
 * HDR -> LDR of `num_pixels` vec4f pixels: rgb *= exp2(exposure) if exposure != 0, the ACES fit if filmic, the sRGB
 * curve if srgb, alpha copied; results are bit-identical to the reference's (its float libm is restated on the device).
 * `ldr` (vec4f per pixel) and / or `ldr_bytes` (vec4b per pixel = float_to_byte of it) are written; either may be null. */
int ygl_tonemap_image(ygl_context* ctx, const float* hdr, int64_t num_pixels, float exposure, int filmic, int srgb,
    float* ldr, uint8_t* ldr_bytes);
/* The same on the image a state holds on its device (the tile's pixels, row-major): what an interactive viewer shows
 * after every batch (apps/ytrace.cpp:219-226) without downloading the float image first. */
int ygl_state_tonemap(ygl_state* state, float exposure, int filmic, int srgb, float* ldr, uint8_t* ldr_bytes);

/* Test hook: the device-side libm (glibc's float routines restated for the GPU) applied to host arrays.
 * fn: 0 sin, 1 cos, 2 exp, 3 log, 4 atan, 5 acos, 6 atan2(x[i], y[i]), 7 pow(x[i], y[i]), 8 sqrt, 9 fmod. */
int ygl_debug_libm(ygl_context* ctx, int fn, const float* x, const float* y, int64_t n, float* out);

/* ---- multi-GPU: one process per device, row tiles, one all-gather at the end ---- */
/* Size in bytes of the opaque NCCL unique id blob. */
int ygl_comm_id_size(void);
/* Rank 0 creates the id; the caller broadcasts it out of band (torch.distributed / file). */
int ygl_comm_create_id(void* id_blob);
int ygl_comm_init(ygl_context* ctx, const void* id_blob, int rank, int nranks);
/* Row range of `rank` for an image of `height` rows: contiguous blocks of ceil(height/nranks). */
void ygl_tile_rows(int height, int rank, int nranks, int* row_begin, int* row_end);
/* ncclAllGather of every rank's tile of state.image into a full host image (w*h*4 floats,
 * may be NULL on ranks that do not want it) — the single collective of the path. Accepts the
 * contiguous tiling of ygl_tile_rows and the interleaved tiling of ygl_state_create_interleaved. */
int  ygl_gather_image(ygl_context* ctx, ygl_state* state, float* image);
void ygl_comm_destroy(ygl_context* ctx);

#ifdef __cplusplus
}
#endif

#endif /* YGL_B200_H */
