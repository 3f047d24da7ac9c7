//
// yocto_b200trace.h — the reference-side binding of libygl_b200.so.
//
// This is the file a Yocto/GL maintainer adds next to libs/yocto/yocto_cutrace.h: it mirrors the
// CPU renderer's API (yocto_trace.h:116-190) one-to-one on top of the C ABI in include/ygl_b200.h,
// exactly as yocto_cutrace.h:71-146 mirrors it for OptiX. It needs the reference headers
// (<yocto/yocto_scene.h>, <yocto/yocto_trace.h>) and is therefore compiled only where they exist
// (tests/test_host_parity.py::test_reference_side_shim_compiles does so in the dev container).
//
//   yocto::trace_image(scene, params)            ->  yocto::b200::trace_image(scene, params)
//   make_trace_bvh / update_trace_bvh / make_trace_lights / make_trace_state / trace_samples / trace_sample / get_image /
//   get_albedo_image / get_normal_image / trace_start / trace_cancel / trace_done / trace_preview likewise;
//   yocto::load_scene(filename)                  ->  yocto::b200::load_scene(filename)  (+ tesselate_subdivs)
//
// Errors of the C ABI come back as the exceptions the reference uses (yocto_trace.cpp:1437,
// :1679-1691): YGL_ERR_INVALID -> std::invalid_argument, everything else -> std::runtime_error.
//
#ifndef YOCTO_B200TRACE_H
#define YOCTO_B200TRACE_H

#include <yocto/yocto_scene.h>
#include <yocto/yocto_trace.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ygl_b200.h"

namespace yocto::b200 {

inline void check(int status) {
  if (status == YGL_OK) return;
  if (status == YGL_ERR_INVALID) throw std::invalid_argument{ygl_last_error()};
  throw std::runtime_error{ygl_last_error()};
}

// Flat views of scene_data for the C ABI (storage for the converted PODs lives here).
struct scene_views {
  std::vector<ygl_camera>      cameras;
  std::vector<ygl_instance>    instances;
  std::vector<ygl_environment> environments;
  std::vector<ygl_material>    materials;
  std::vector<ygl_texture>     textures;
  std::vector<ygl_shape>       shapes;
  ygl_scene_desc               desc = {};
};

inline ygl_frame3f to_abi(const frame3f& f) {
  return {{f.x.x, f.x.y, f.x.z}, {f.y.x, f.y.y, f.y.z}, {f.z.x, f.z.y, f.z.z}, {f.o.x, f.o.y, f.o.z}};
}

inline scene_views make_views(const scene_data& scene) {
  auto v = scene_views{};
  for (auto& c : scene.cameras)
    v.cameras.push_back({to_abi(c.frame), c.orthographic ? 1 : 0, c.lens, c.film, c.aspect, c.focus, c.aperture});
  for (auto& i : scene.instances) v.instances.push_back({to_abi(i.frame), i.shape, i.material});
  for (auto& e : scene.environments)
    v.environments.push_back({to_abi(e.frame), {e.emission.x, e.emission.y, e.emission.z}, e.emission_tex});
  for (auto& m : scene.materials)
    v.materials.push_back({(int)m.type, {m.emission.x, m.emission.y, m.emission.z}, {m.color.x, m.color.y, m.color.z},
        m.roughness, m.metallic, m.ior, {m.scattering.x, m.scattering.y, m.scattering.z}, m.scanisotropy, m.trdepth,
        m.opacity, m.emission_tex, m.color_tex, m.roughness_tex, m.scattering_tex, m.normal_tex});
  for (auto& t : scene.textures)
    v.textures.push_back({t.width, t.height, t.linear ? 1 : 0, t.nearest ? 1 : 0, t.clamp ? 1 : 0,
        t.pixelsf.empty() ? nullptr : (const float*)t.pixelsf.data(),
        t.pixelsb.empty() ? nullptr : (const uint8_t*)t.pixelsb.data()});
  for (auto& s : scene.shapes)
    v.shapes.push_back({(int)s.points.size(), (int)s.lines.size(), (int)s.triangles.size(), (int)s.quads.size(),
        (const int32_t*)s.points.data(), (const int32_t*)s.lines.data(), (const int32_t*)s.triangles.data(),
        (const int32_t*)s.quads.data(), (int)s.positions.size(), (int)s.normals.size(), (int)s.texcoords.size(),
        (int)s.colors.size(), (int)s.radius.size(), (const float*)s.positions.data(), (const float*)s.normals.data(),
        (const float*)s.texcoords.data(), (const float*)s.colors.data(), (const float*)s.radius.data()});
  v.desc = {(int)v.cameras.size(), (int)v.instances.size(), (int)v.environments.size(), (int)v.shapes.size(),
      (int)v.textures.size(), (int)v.materials.size(), v.cameras.data(), v.instances.data(), v.environments.data(),
      v.shapes.data(), v.textures.data(), v.materials.data()};
  return v;
}

inline ygl_trace_params to_abi(const trace_params& p) {
  auto a = ygl_trace_params{};
  a.camera = p.camera, a.resolution = p.resolution, a.sampler = (int)p.sampler, a.falsecolor = (int)p.falsecolor;
  a.samples = p.samples, a.bounces = p.bounces, a.clamp = p.clamp;
  a.nocaustics = p.nocaustics, a.envhidden = p.envhidden, a.tentfilter = p.tentfilter, a.seed = p.seed;
  a.embreebvh = p.embreebvh, a.highqualitybvh = p.highqualitybvh, a.noparallel = p.noparallel;
  a.pratio = p.pratio, a.denoise = p.denoise, a.batch = p.batch;
  return a;
}

// RAII owners, like the cutrace_* types (yocto_cutrace.h:270-397)
struct b200_context {
  ygl_context* handle = nullptr;
  explicit b200_context(int device = 0) { check(ygl_context_create(device, &handle)); }
  b200_context(const b200_context&)            = delete;
  b200_context& operator=(const b200_context&) = delete;
  ~b200_context() { ygl_context_destroy(handle); }
  // scheduling knobs by name (never change a result): ygl_context_set_option
  void set_option(const char* name, double value) { check(ygl_context_set_option(handle, name, value)); }
};
struct b200_scene {
  ygl_scene* handle = nullptr;
  ~b200_scene() { ygl_scene_destroy(handle); }
};
struct b200_bvh {
  ygl_bvh* handle = nullptr;
  ~b200_bvh() { ygl_bvh_destroy(handle); }
};
struct b200_lights {
  ygl_lights* handle = nullptr;
  ~b200_lights() { ygl_lights_destroy(handle); }
};
struct b200_state {
  ygl_state* handle = nullptr;
  int        width = 0, height = 0;
  ~b200_state() { ygl_state_destroy(handle); }
};

// make_cutrace_scene-style upload of scene_data
inline std::unique_ptr<b200_scene> make_b200_scene(b200_context& ctx, const scene_data& scene) {
  auto views = make_views(scene);
  auto out   = std::make_unique<b200_scene>();
  check(ygl_scene_create(ctx.handle, &views.desc, &out->handle));
  return out;
}
// make_trace_bvh (yocto_trace.cpp:88)
inline std::unique_ptr<b200_bvh> make_trace_bvh(const scene_data& scene, const trace_params& params) {
  auto views = make_views(scene);
  auto out   = std::make_unique<b200_bvh>();
  check(ygl_bvh_build(&views.desc, params.highqualitybvh ? 1 : 0, &out->handle));
  return out;
}
// tonemap_image (yocto_image.h:242-245): HDR -> LDR on the device, bit-identical to the CPU version
inline void tonemap_image(b200_context& ctx, std::vector<vec4f>& ldr, const std::vector<vec4f>& hdr, float exposure,
    bool filmic = false, bool srgb = true) {
  ldr.resize(hdr.size());
  check(ygl_tonemap_image(ctx.handle, (const float*)hdr.data(), (int64_t)hdr.size(), exposure, filmic ? 1 : 0, srgb ? 1 : 0,
      (float*)ldr.data(), nullptr));
}
inline void tonemap_image(b200_context& ctx, std::vector<vec4b>& ldr, const std::vector<vec4f>& hdr, float exposure,
    bool filmic = false, bool srgb = true) {
  ldr.resize(hdr.size());
  check(ygl_tonemap_image(ctx.handle, (const float*)hdr.data(), (int64_t)hdr.size(), exposure, filmic ? 1 : 0, srgb ? 1 : 0,
      nullptr, (uint8_t*)ldr.data()));
}
// ... and of the image a state holds, without downloading the floats first (the interactive loop, apps/ytrace.cpp:219-226)
inline void tonemap_image(std::vector<vec4b>& ldr, const b200_state& state, float exposure, bool filmic = false,
    bool srgb = true) {
  ldr.resize((size_t)state.width * state.height);
  check(ygl_state_tonemap(state.handle, exposure, filmic ? 1 : 0, srgb ? 1 : 0, nullptr, (uint8_t*)ldr.data()));
}
// update_scene_bvh (yocto_bvh.h:88, yocto_bvh.cpp:434): refit to the edited scene, same topology
inline void update_trace_bvh(b200_bvh& bvh, const scene_data& scene, const std::vector<int>& updated_instances,
    const std::vector<int>& updated_shapes) {
  auto views = make_views(scene);
  check(ygl_bvh_update(bvh.handle, &views.desc, updated_instances.data(), (int)updated_instances.size(),
      updated_shapes.data(), (int)updated_shapes.size()));
}
// make_trace_lights (yocto_trace.cpp:1528)
inline std::unique_ptr<b200_lights> make_trace_lights(const scene_data& scene, const trace_params&) {
  auto views = make_views(scene);
  auto out   = std::make_unique<b200_lights>();
  check(ygl_lights_create(&views.desc, &out->handle));
  return out;
}
// make_trace_state (yocto_trace.cpp:1495)
inline std::unique_ptr<b200_state> make_trace_state(
    b200_context& ctx, const scene_data& scene, const trace_params& params) {
  auto views = make_views(scene);
  auto abi   = to_abi(params);
  auto out   = std::make_unique<b200_state>();
  check(ygl_state_create(ctx.handle, &views.desc, &abi, &out->handle));
  check(ygl_state_size(out->handle, &out->width, &out->height, nullptr));
  return out;
}
// trace_samples (yocto_trace.cpp:1595)
inline void trace_samples(b200_context& ctx, b200_state& state, const b200_scene& scene, const b200_bvh& bvh,
    const b200_lights& lights, const trace_params& params) {
  auto abi = to_abi(params);
  check(ygl_trace_samples(ctx.handle, state.handle, scene.handle, bvh.handle, lights.handle, &abi));
}
// make_trace_bvh from trees the reference has already built (trace_bvh::bvh, yocto_trace.h:138): adopted verbatim
inline std::unique_ptr<b200_bvh> make_trace_bvh(const scene_data& scene, const trace_bvh& built) {
  static_assert(sizeof(bvh_node) == sizeof(ygl_bvh_node), "bvh_node layout");
  auto views = make_views(scene);
  auto nodes = std::vector<const ygl_bvh_node*>{};
  auto prims = std::vector<const int32_t*>{};
  auto nn = std::vector<int>{}, np = std::vector<int>{};
  for (auto& shape : built.bvh.shapes) {
    nodes.push_back((const ygl_bvh_node*)shape.bvh.nodes.data());
    prims.push_back((const int32_t*)shape.bvh.primitives.data());
    nn.push_back((int)shape.bvh.nodes.size());
    np.push_back((int)shape.bvh.primitives.size());
  }
  auto out = std::make_unique<b200_bvh>();
  check(ygl_bvh_create_from_host(&views.desc, (const ygl_bvh_node*)built.bvh.bvh.nodes.data(),
      (int)built.bvh.bvh.nodes.size(), (const int32_t*)built.bvh.bvh.primitives.data(),
      (int)built.bvh.bvh.primitives.size(), nodes.data(), nn.data(), prims.data(), np.data(), &out->handle));
  return out;
}
// trace_sample (yocto_trace.cpp:1461)
inline void trace_sample(b200_context& ctx, b200_state& state, const b200_scene& scene, const b200_bvh& bvh,
    const b200_lights& lights, int i, int j, int sample, const trace_params& params) {
  auto abi = to_abi(params);
  check(ygl_trace_sample(ctx.handle, state.handle, scene.handle, bvh.handle, lights.handle, i, j, sample, &abi));
}
// reset_cutrace_state-style reset (yocto_cutrace.h:119)
inline void reset_trace_state(b200_state& state, const trace_params& params) {
  auto abi = to_abi(params);
  check(ygl_state_reset(state.handle, &abi));
}
// check_image, yocto_trace.cpp:1679-1686
inline void check_image(const image_data& image, int width, int height, bool linear) {
  if (image.width != width || image.height != height) throw std::invalid_argument{"image should have the same size"};
  if (image.linear != linear) throw std::invalid_argument{linear ? "expected linear image" : "expected srgb image"};
}
// get_image (yocto_trace.cpp:1694-1708)
inline void get_image(image_data& image, b200_state& state) {
  image.width = state.width, image.height = state.height, image.linear = true;
  image.pixels.resize((size_t)state.width * state.height);
  check(ygl_state_download(state.handle, (float*)image.pixels.data(), nullptr, nullptr, nullptr, nullptr));
}
inline image_data get_image(b200_state& state) {
  auto image = make_image(state.width, state.height, true);
  get_image(image, state);
  return image;
}
// get_albedo_image / get_normal_image (yocto_trace.cpp:1767-1790): the denoise guides as rgba images, alpha 1
inline void get_guide_image(image_data& image, b200_state& state, bool normal) {
  check_image(image, state.width, state.height, true);
  auto rgb = std::vector<vec3f>((size_t)state.width * state.height);
  check(ygl_state_download(state.handle, nullptr, normal ? nullptr : (float*)rgb.data(),
      normal ? (float*)rgb.data() : nullptr, nullptr, nullptr));
  for (auto idx = (size_t)0; idx < rgb.size(); idx++) image.pixels[idx] = {rgb[idx].x, rgb[idx].y, rgb[idx].z, 1.0f};
}
inline void       get_albedo_image(image_data& image, b200_state& state) { get_guide_image(image, state, false); }
inline void       get_normal_image(image_data& image, b200_state& state) { get_guide_image(image, state, true); }
inline image_data get_albedo_image(b200_state& state) {
  auto image = make_image(state.width, state.height, true);
  get_albedo_image(image, state);
  return image;
}
inline image_data get_normal_image(b200_state& state) {
  auto image = make_image(state.width, state.height, true);
  get_normal_image(image, state);
  return image;
}
// trace_start / trace_cancel / trace_done / trace_preview (yocto_trace.h:209-223): the context plays trace_context
inline void trace_start(b200_context& ctx, b200_state& state, const b200_scene& scene, const b200_bvh& bvh,
    const b200_lights& lights, const trace_params& params) {
  auto abi = to_abi(params);
  check(ygl_trace_start(ctx.handle, state.handle, scene.handle, bvh.handle, lights.handle, &abi));
}
inline void trace_cancel(b200_context& ctx) { check(ygl_trace_cancel(ctx.handle)); }
inline bool trace_done(const b200_context& ctx) { return ygl_trace_done(ctx.handle) != 0; }
inline void trace_preview(image_data& image, b200_context& ctx, b200_state& state, const b200_scene& scene,
    const b200_bvh& bvh, const b200_lights& lights, const trace_params& params) {
  check_image(image, state.width, state.height, true);
  auto abi = to_abi(params);
  check(ygl_trace_preview(ctx.handle, scene.handle, bvh.handle, lights.handle, &abi, state.width, state.height,
      (float*)image.pixels.data()));
}
// trace_image (yocto_trace.cpp:1584): the drop-in
inline image_data trace_image(const scene_data& scene, const trace_params& params, int device = 0) {
  auto ctx   = b200_context{device};
  auto views = make_views(scene);
  auto abi   = to_abi(params);
  int  w = 0, h = 0;
  check(ygl_trace_image(ctx.handle, &views.desc, &abi, &w, &h, nullptr));
  auto image = make_image(w, h, true);
  check(ygl_trace_image(ctx.handle, &views.desc, &abi, &w, &h, (float*)image.pixels.data()));
  return image;
}

// load_scene (yocto_sceneio.h:93-96): the library's own reader (JSON 4.0 / 4.2 / 5.0, .ply scenes, PLY / OBJ shapes,
// PNG / HDR textures) filling the reference's scene_data. Subdivs arrive tesselated (scene.subdivs stays empty), which is
// the state every reference app reaches with tesselate_subdivs right after load_scene (apps/ytrace.cpp:110-113).
inline bool load_scene(const std::string& filename, scene_data& scene, std::string& error) {
  ygl_loaded_scene* loaded = nullptr;
  if (ygl_scene_load(filename.c_str(), &loaded) != YGL_OK) {
    error = ygl_last_error();
    return false;
  }
  auto guard = std::unique_ptr<ygl_loaded_scene, void (*)(ygl_loaded_scene*)>{loaded, ygl_loaded_scene_destroy};
  auto desc  = ygl_loaded_scene_desc(loaded);
  auto frame = [](const ygl_frame3f& f) { return *(const frame3f*)&f; };
  static_assert(sizeof(ygl_frame3f) == sizeof(frame3f), "frame layout");
  scene = scene_data{};
  for (auto k = 0; k < desc->num_cameras; k++) {
    auto& c = desc->cameras[k];
    auto& camera = scene.cameras.emplace_back();
    camera.frame = frame(c.frame), camera.orthographic = c.orthographic != 0, camera.lens = c.lens, camera.film = c.film;
    camera.aspect = c.aspect, camera.focus = c.focus, camera.aperture = c.aperture;
  }
  for (auto k = 0; k < desc->num_instances; k++)
    scene.instances.push_back({frame(desc->instances[k].frame), desc->instances[k].shape, desc->instances[k].material});
  for (auto k = 0; k < desc->num_environments; k++) {
    auto& e = desc->environments[k];
    scene.environments.push_back({frame(e.frame), {e.emission[0], e.emission[1], e.emission[2]}, e.emission_tex});
  }
  for (auto k = 0; k < desc->num_materials; k++) {
    auto& m        = desc->materials[k];
    auto& material = scene.materials.emplace_back();
    material.type = (material_type)m.type, material.emission = {m.emission[0], m.emission[1], m.emission[2]};
    material.color = {m.color[0], m.color[1], m.color[2]}, material.roughness = m.roughness, material.metallic = m.metallic;
    material.ior = m.ior, material.scattering = {m.scattering[0], m.scattering[1], m.scattering[2]};
    material.scanisotropy = m.scanisotropy, material.trdepth = m.trdepth, material.opacity = m.opacity;
    material.emission_tex = m.emission_tex, material.color_tex = m.color_tex, material.roughness_tex = m.roughness_tex;
    material.scattering_tex = m.scattering_tex, material.normal_tex = m.normal_tex;
  }
  for (auto k = 0; k < desc->num_textures; k++) {
    auto& t       = desc->textures[k];
    auto& texture = scene.textures.emplace_back();
    texture.width = t.width, texture.height = t.height, texture.linear = t.linear != 0, texture.nearest = t.nearest != 0;
    texture.clamp = t.clamp != 0;
    auto count    = (size_t)t.width * (size_t)t.height;
    if (t.pixelsf) texture.pixelsf.assign((const vec4f*)t.pixelsf, (const vec4f*)t.pixelsf + count);
    if (t.pixelsb) texture.pixelsb.assign((const vec4b*)t.pixelsb, (const vec4b*)t.pixelsb + count);
  }
  for (auto k = 0; k < desc->num_shapes; k++) {
    auto& s     = desc->shapes[k];
    auto& shape = scene.shapes.emplace_back();
    shape.points.assign(s.points, s.points + s.num_points);
    shape.lines.assign((const vec2i*)s.lines, (const vec2i*)s.lines + s.num_lines);
    shape.triangles.assign((const vec3i*)s.triangles, (const vec3i*)s.triangles + s.num_triangles);
    shape.quads.assign((const vec4i*)s.quads, (const vec4i*)s.quads + s.num_quads);
    shape.positions.assign((const vec3f*)s.positions, (const vec3f*)s.positions + s.num_positions);
    shape.normals.assign((const vec3f*)s.normals, (const vec3f*)s.normals + s.num_normals);
    shape.texcoords.assign((const vec2f*)s.texcoords, (const vec2f*)s.texcoords + s.num_texcoords);
    shape.colors.assign((const vec4f*)s.colors, (const vec4f*)s.colors + s.num_colors);
    shape.radius.assign(s.radius, s.radius + s.num_radius);
  }
  std::vector<std::string>* names[6] = {&scene.camera_names, &scene.texture_names, &scene.material_names, &scene.shape_names,
      &scene.instance_names, &scene.environment_names};
  for (auto kind = 0; kind < 6; kind++)
    for (auto k = 0; auto name = ygl_loaded_scene_name(loaded, kind, k); k++) names[kind]->push_back(name);
  return true;
}
inline scene_data load_scene(const std::string& filename) {
  auto scene = scene_data{};
  auto error = std::string{};
  if (!b200::load_scene(filename, scene, error)) throw std::runtime_error{error};
  return scene;
}

}  // namespace yocto::b200

#endif
