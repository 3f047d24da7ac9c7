// ygl_eval.cuh — hit-point evaluation: camera rays, positions, normals, materials, textures,
// environments. Behavioural contract: libs/yocto/yocto_scene.cpp:66-178 (camera, textures),
// :203-613 (materials, instance properties, environments), yocto_geometry.h:508-654,
// yocto_color.h:223-249, yocto_shape.cpp:63-82.
#pragma once

#include "ygl_scene.cuh"

namespace ygl {

// ---- eval_camera, yocto_scene.cpp:66-101 ----
YGL_D void eval_camera(const DCamera& camera, const f2& image_uv, const f2& lens_uv, f3& ray_o, f3& ray_d) {
  f2 film = camera.aspect >= 1 ? f2{camera.film, camera.film / camera.aspect}
                               : f2{camera.film * camera.aspect, camera.film};
  if (!camera.orthographic) {
    auto q  = f3{film.x * (0.5f - image_uv.x), film.y * (image_uv.y - 0.5f), camera.lens};
    auto dc = -normalize(q);
    auto e  = f3{lens_uv.x * camera.aperture / 2, lens_uv.y * camera.aperture / 2, 0};
    auto p  = dc * camera.focus / yabs(dc.z);
    auto d  = normalize(p - e);
    ray_o   = transform_point(camera.frame, e);
    ray_d   = transform_direction(camera.frame, d);
  } else {
    auto scale = 1 / camera.lens;
    auto q     = f3{film.x * (0.5f - image_uv.x) * scale, film.y * (image_uv.y - 0.5f) * scale, camera.lens};
    auto e     = f3{-q.x, -q.y, 0} + f3{lens_uv.x * camera.aperture / 2, lens_uv.y * camera.aperture / 2, 0};
    auto p     = f3{-q.x, -q.y, -camera.focus};
    auto d     = normalize(p - e);
    ray_o      = transform_point(camera.frame, e);
    ray_d      = transform_direction(camera.frame, d);
  }
}

// sample_camera, yocto_trace.cpp:338-358
YGL_D void sample_camera(const DCamera& camera, int i, int j, int width, int height, const f2& puv,
    const f2& luv, bool tent, f3& ray_o, f3& ray_d) {
  if (!tent) {
    auto uv = f2{(i + puv.x) / width, (j + puv.y) / height};
    eval_camera(camera, uv, sample_disk(luv), ray_o, ray_d);
  } else {
    const auto width_  = 2.0f;
    const auto offset  = 0.5f;
    auto       fuv     = f2{puv.x < 0.5f ? ysqrt(2 * puv.x) - 1 : 1 - ysqrt(2 - 2 * puv.x),
                  puv.y < 0.5f ? ysqrt(2 * puv.y) - 1 : 1 - ysqrt(2 - 2 * puv.y)};
    fuv                = f2{width_ * fuv.x + offset, width_ * fuv.y + offset};
    auto uv            = f2{(i + fuv.x) / width, (j + fuv.y) / height};
    eval_camera(camera, uv, sample_disk(luv), ray_o, ray_d);
  }
}

// ---- textures ----
// srgb_to_rgb, yocto_color.h:235-238 (the threshold is a double literal)
YGL_D float srgb_to_rgb(float srgb) {
  return ((double)srgb <= 0.04045) ? srgb / 12.92f : ypow((srgb + 0.055f) / (1.0f + 0.055f), 2.4f);
}
// lookup_texture, yocto_scene.cpp:111-124
YGL_D f4 lookup_texture(const DTexture& tex, int i, int j, bool as_linear) {
  f4 color;
  if (tex.pixelsf) {
    float4 v = __ldg(tex.pixelsf + (size_t)j * tex.width + i);
    color    = {v.x, v.y, v.z, v.w};
  } else {
    uchar4 b = __ldg(tex.pixelsb + (size_t)j * tex.width + i);
    color    = {b.x / 255.0f, b.y / 255.0f, b.z / 255.0f, b.w / 255.0f};
  }
  if (as_linear && !tex.linear) {
    return {srgb_to_rgb(color.x), srgb_to_rgb(color.y), srgb_to_rgb(color.z), color.w};
  }
  return color;
}
// eval_texture, yocto_scene.cpp:127-160. The fetch (wrap, four taps, sRGB decode with one powf per channel) is ~700
// SASS instructions and used to be inlined at six call sites of the shading kernels, whose profile is dominated by
// instruction-fetch stalls; untextured materials (texture id < 0) never reach it. One out-of-line copy per kernel:
// pure code layout, the arithmetic is unchanged.
// RAW: the copy for callers that never ask for the sRGB decode (environment maps): texel values as stored, no powf.
template <bool RAW>
static __device__ __noinline__ f4 eval_texture_fetch(const DTexture* texp, float uvx, float uvy, bool as_linear_,
    bool no_interpolation, bool clamp_to_edge) {
  const bool      as_linear = RAW ? false : as_linear_;
  const DTexture& tex       = *texp;
  if (tex.width == 0 || tex.height == 0) return {0, 0, 0, 0};
  int   sx = tex.width, sy = tex.height;
  float s = 0.0f, t = 0.0f;
  if (clamp_to_edge) {
    s = yclamp(uvx, 0.0f, 1.0f) * sx;
    t = yclamp(uvy, 0.0f, 1.0f) * sy;
  } else {
    s = yfmod(uvx, 1.0f) * sx;
    if (s < 0) s += sx;
    t = yfmod(uvy, 1.0f) * sy;
    if (t < 0) t += sy;
  }
  int   i = iclamp((int)s, 0, sx - 1), j = iclamp((int)t, 0, sy - 1);
  int   ii = (i + 1) % sx, jj = (j + 1) % sy;
  float u = s - i, v = t - j;
  if (no_interpolation) return lookup_texture(tex, i, j, as_linear);
  return lookup_texture(tex, i, j, as_linear) * (1 - u) * (1 - v) +
         lookup_texture(tex, i, jj, as_linear) * (1 - u) * v +
         lookup_texture(tex, ii, j, as_linear) * u * (1 - v) +
         lookup_texture(tex, ii, jj, as_linear) * u * v;
}
YGL_D f4 eval_texture(const DTexture& tex, const f2& uv, bool as_linear, bool no_interpolation,
    bool clamp_to_edge) {
  return eval_texture_fetch<false>(&tex, uv.x, uv.y, as_linear, no_interpolation, clamp_to_edge);
}
// eval_texture(scene, id, uv, as_linear), yocto_scene.cpp:167-171
YGL_D_BIG f4 eval_texture(const DScene& scene, int texture, const f2& uv, bool as_linear) {
  if (texture < 0) return {1, 1, 1, 1};
  const DTexture& tex = scene.textures[texture];
  return eval_texture(tex, uv, as_linear, tex.nearest != 0, tex.clamp != 0);
}
// eval_texture(scene, id, uv, false): the values as stored (no sRGB decode)
YGL_D_BIG f4 eval_texture_raw(const DScene& scene, int texture, const f2& uv) {
  if (texture < 0) return {1, 1, 1, 1};
  const DTexture& tex = scene.textures[texture];
  return eval_texture_fetch<true>(&tex, uv.x, uv.y, false, tex.nearest != 0, tex.clamp != 0);
}

// ---- interpolation, yocto_geometry.h:536-556 ----
template <typename T>
YGL_D T interp_line(const T& p0, const T& p1, float u) {
  return p0 * (1 - u) + p1 * u;
}
template <typename T>
YGL_D T interp_triangle(const T& p0, const T& p1, const T& p2, const f2& uv) {
  return p0 * (1 - uv.x - uv.y) + p1 * uv.x + p2 * uv.y;
}
template <typename T>
YGL_D T interp_quad(const T& p0, const T& p1, const T& p2, const T& p3, const f2& uv) {
  if (uv.x + uv.y <= 1) return interp_triangle(p0, p1, p3, uv);
  return interp_triangle(p2, p3, p1, f2{1 - uv.x, 1 - uv.y});
}
YGL_D f3 triangle_normal(const f3& p0, const f3& p1, const f3& p2) { return normalize(cross(p1 - p0, p2 - p0)); }
YGL_D f3 quad_normal(const f3& p0, const f3& p1, const f3& p2, const f3& p3) {
  return normalize(triangle_normal(p0, p1, p3) + triangle_normal(p2, p3, p1));
}

struct elem_ids {
  int x, y, z, w;
};
YGL_D elem_ids load_triangle(const DShape& s, int e) {
  return {__ldg(s.triangles + 3 * e), __ldg(s.triangles + 3 * e + 1), __ldg(s.triangles + 3 * e + 2), 0};
}
YGL_D elem_ids load_quad(const DShape& s, int e) {
  int4 q = __ldg((const int4*)s.quads + e);
  return {q.x, q.y, q.z, q.w};
}
YGL_D elem_ids load_line(const DShape& s, int e) { return {__ldg(s.lines + 2 * e), __ldg(s.lines + 2 * e + 1), 0, 0}; }

// eval_position (object space part), yocto_scene.cpp:288-312 order: triangles, quads, lines, points
YGL_D f3 eval_position_local(const DShape& s, int element, const f2& uv) {
  if (s.eval_kind == kElemTriangles) {
    auto t = load_triangle(s, element);
    return interp_triangle(ld3(s.positions, t.x), ld3(s.positions, t.y), ld3(s.positions, t.z), uv);
  } else if (s.eval_kind == kElemQuads) {
    auto q = load_quad(s, element);
    return interp_quad(ld3(s.positions, q.x), ld3(s.positions, q.y), ld3(s.positions, q.z), ld3(s.positions, q.w), uv);
  } else if (s.eval_kind == kElemLines) {
    auto l = load_line(s, element);
    return interp_line(ld3(s.positions, l.x), ld3(s.positions, l.y), uv.x);
  } else if (s.eval_kind == kElemPoints) {
    return ld3(s.positions, __ldg(s.points + element));
  }
  return {0, 0, 0};
}
YGL_D f3 eval_position(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape& s = scene.shapes[inst.shape];
  if (s.eval_kind == kElemNone) return {0, 0, 0};
  return transform_point(inst.frame, eval_position_local(s, element, uv));
}
// eval_position(shape, element, uv), yocto_shape.cpp:63-82 order: points, lines, triangles, quads
YGL_D f3 eval_position_shape(const DShape& s, int element, const f2& uv) {
  if (s.num_points > 0) {
    return ld3(s.positions, __ldg(s.points + element));
  } else if (s.num_lines > 0) {
    auto l = load_line(s, element);
    return interp_line(ld3(s.positions, l.x), ld3(s.positions, l.y), uv.x);
  } else if (s.num_triangles > 0) {
    auto t = load_triangle(s, element);
    return interp_triangle(ld3(s.positions, t.x), ld3(s.positions, t.y), ld3(s.positions, t.z), uv);
  } else if (s.num_quads > 0) {
    auto q = load_quad(s, element);
    return interp_quad(ld3(s.positions, q.x), ld3(s.positions, q.y), ld3(s.positions, q.z), ld3(s.positions, q.w), uv);
  }
  return {0, 0, 0};
}

// eval_element_normal, yocto_scene.cpp:315-337
YGL_D f3 eval_element_normal(const DScene& scene, const DInstance& inst, int element) {
  const DShape& s = scene.shapes[inst.shape];
  if (s.eval_kind == kElemTriangles) {
    auto t = load_triangle(s, element);
    return transform_normal(inst.frame, triangle_normal(ld3(s.positions, t.x), ld3(s.positions, t.y), ld3(s.positions, t.z)));
  } else if (s.eval_kind == kElemQuads) {
    auto q = load_quad(s, element);
    return transform_normal(inst.frame,
        quad_normal(ld3(s.positions, q.x), ld3(s.positions, q.y), ld3(s.positions, q.z), ld3(s.positions, q.w)));
  } else if (s.eval_kind == kElemLines) {
    auto l = load_line(s, element);
    return transform_normal(inst.frame, normalize(ld3(s.positions, l.y) - ld3(s.positions, l.x)));
  } else if (s.eval_kind == kElemPoints) {
    return {0, 0, 1};
  }
  return {0, 0, 0};
}

// eval_normal, yocto_scene.cpp:340-366
YGL_D f3 eval_normal(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape& s = scene.shapes[inst.shape];
  if (!s.normals) return eval_element_normal(scene, inst, element);
  if (s.eval_kind == kElemTriangles) {
    auto t = load_triangle(s, element);
    return transform_normal(inst.frame,
        normalize(interp_triangle(ld3(s.normals, t.x), ld3(s.normals, t.y), ld3(s.normals, t.z), uv)));
  } else if (s.eval_kind == kElemQuads) {
    auto q = load_quad(s, element);
    return transform_normal(inst.frame,
        normalize(interp_quad(ld3(s.normals, q.x), ld3(s.normals, q.y), ld3(s.normals, q.z), ld3(s.normals, q.w), uv)));
  } else if (s.eval_kind == kElemLines) {
    auto l = load_line(s, element);
    return transform_normal(inst.frame, normalize(interp_line(ld3(s.normals, l.x), ld3(s.normals, l.y), uv.x)));
  } else if (s.eval_kind == kElemPoints) {
    return transform_normal(inst.frame, normalize(ld3(s.normals, __ldg(s.points + element))));
  }
  return {0, 0, 0};
}

// eval_texcoord, yocto_scene.cpp:369-389
YGL_D f2 eval_texcoord(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape& s = scene.shapes[inst.shape];
  if (!s.texcoords) return uv;
  if (s.eval_kind == kElemTriangles) {
    auto t = load_triangle(s, element);
    return interp_triangle(ld2(s.texcoords, t.x), ld2(s.texcoords, t.y), ld2(s.texcoords, t.z), uv);
  } else if (s.eval_kind == kElemQuads) {
    auto q = load_quad(s, element);
    return interp_quad(ld2(s.texcoords, q.x), ld2(s.texcoords, q.y), ld2(s.texcoords, q.z), ld2(s.texcoords, q.w), uv);
  } else if (s.eval_kind == kElemLines) {
    auto l = load_line(s, element);
    return interp_line(ld2(s.texcoords, l.x), ld2(s.texcoords, l.y), uv.x);
  } else if (s.eval_kind == kElemPoints) {
    return ld2(s.texcoords, __ldg(s.points + element));
  }
  return {0, 0};
}

// eval_color, yocto_scene.cpp:507-527
YGL_D f4 eval_color(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape& s = scene.shapes[inst.shape];
  if (!s.colors) return {1, 1, 1, 1};
  if (s.eval_kind == kElemTriangles) {
    auto t = load_triangle(s, element);
    return interp_triangle(ld4(s.colors, t.x), ld4(s.colors, t.y), ld4(s.colors, t.z), uv);
  } else if (s.eval_kind == kElemQuads) {
    auto q = load_quad(s, element);
    return interp_quad(ld4(s.colors, q.x), ld4(s.colors, q.y), ld4(s.colors, q.z), ld4(s.colors, q.w), uv);
  } else if (s.eval_kind == kElemLines) {
    auto l = load_line(s, element);
    return interp_line(ld4(s.colors, l.x), ld4(s.colors, l.y), uv.x);
  } else if (s.eval_kind == kElemPoints) {
    return ld4(s.colors, __ldg(s.points + element));
  }
  return {0, 0, 0, 0};
}

// triangle_tangents_fromuv, yocto_geometry.h:620-643
YGL_D void triangle_tangents_fromuv(const f3& p0, const f3& p1, const f3& p2, const f2& uv0, const f2& uv1,
    const f2& uv2, f3& tu, f3& tv) {
  auto p   = p1 - p0;
  auto q   = p2 - p0;
  auto s   = f2{uv1.x - uv0.x, uv2.x - uv0.x};
  auto t   = f2{uv1.y - uv0.y, uv2.y - uv0.y};
  auto div = s.x * t.y - s.y * t.x;
  if (div != 0) {
    tu = f3{t.y * p.x - t.x * q.x, t.y * p.y - t.x * q.y, t.y * p.z - t.x * q.z} / div;
    tv = f3{s.x * q.x - s.y * p.x, s.x * q.y - s.y * p.y, s.x * q.z - s.y * p.z} / div;
  } else {
    tu = {1, 0, 0};
    tv = {0, 1, 0};
  }
}
// eval_element_tangents, yocto_scene.cpp:425-446 (quads: current_uv = {0,0} -> first triangle)
YGL_D void eval_element_tangents(const DScene& scene, const DInstance& inst, int element, f3& tu, f3& tv) {
  const DShape& s = scene.shapes[inst.shape];
  if (s.num_triangles > 0 && s.texcoords) {
    auto t = load_triangle(s, element);
    f3   a, b;
    triangle_tangents_fromuv(ld3(s.positions, t.x), ld3(s.positions, t.y), ld3(s.positions, t.z),
        ld2(s.texcoords, t.x), ld2(s.texcoords, t.y), ld2(s.texcoords, t.z), a, b);
    tu = transform_direction(inst.frame, a);
    tv = transform_direction(inst.frame, b);
  } else if (s.num_quads > 0 && s.texcoords) {
    auto q = load_quad(s, element);
    f3   a, b;
    triangle_tangents_fromuv(ld3(s.positions, q.x), ld3(s.positions, q.y), ld3(s.positions, q.w),
        ld2(s.texcoords, q.x), ld2(s.texcoords, q.y), ld2(s.texcoords, q.w), a, b);
    tu = transform_direction(inst.frame, a);
    tv = transform_direction(inst.frame, b);
  } else {
    tu = {0, 0, 0};
    tv = {0, 0, 0};
  }
}

// eval_normalmap, yocto_scene.cpp:448-468
YGL_D f3 eval_normalmap(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape&    s   = scene.shapes[inst.shape];
  const DMaterial& mat = scene.materials[inst.material];
  auto normal   = eval_normal(scene, inst, element, uv);
  auto texcoord = eval_texcoord(scene, inst, element, uv);
  if (mat.normal_tex >= 0 && (s.num_triangles > 0 || s.num_quads > 0)) {
    const DTexture& tex = scene.textures[mat.normal_tex];
    auto normalmap = -1 + 2 * xyz(eval_texture(tex, texcoord, false, tex.nearest != 0, tex.clamp != 0));
    f3   tu, tv;
    eval_element_tangents(scene, inst, element, tu, tv);
    auto fx     = orthonormalize(tu, normal);
    auto fy     = normalize(cross(normal, fx));
    auto flip_v = dot(fy, tv) < 0;
    normalmap.y *= flip_v ? 1 : -1;
    // transform_normal(frame{fx, fy, normal, 0}, normalmap) with non_rigid = false
    normal = normalize(fx * normalmap.x + fy * normalmap.y + normal * normalmap.z);
  }
  return normal;
}

// eval_shading_position, yocto_scene.cpp:471-483 (points: object-space position, reference quirk)
YGL_D f3 eval_shading_position(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DShape& s = scene.shapes[inst.shape];
  if (s.num_triangles > 0 || s.num_quads > 0) return eval_position(scene, inst, element, uv);
  if (s.num_lines > 0) return eval_position(scene, inst, element, uv);
  if (s.num_points > 0) return eval_position_shape(s, element, uv);
  return {0, 0, 0};
}
// eval_shading_normal, yocto_scene.cpp:486-505
YGL_D_BIG f3 eval_shading_normal(const DScene& scene, const DInstance& inst, int element, const f2& uv, const f3& outgoing) {
  const DShape&    s   = scene.shapes[inst.shape];
  const DMaterial& mat = scene.materials[inst.material];
  if (s.num_triangles > 0 || s.num_quads > 0) {
    auto normal = eval_normal(scene, inst, element, uv);
    if (mat.normal_tex >= 0) normal = eval_normalmap(scene, inst, element, uv);
    if (mat.type == kRefractive) return normal;
    return dot(normal, outgoing) >= 0 ? normal : -normal;
  } else if (s.num_lines > 0) {
    auto normal = eval_normal(scene, inst, element, uv);
    return orthonormalize(outgoing, normal);
  } else if (s.num_points > 0) {
    return outgoing;
  }
  return {0, 0, 0};
}

// eval_material, yocto_scene.cpp:531-581
YGL_D_BIG mpoint eval_material(const DScene& scene, const DInstance& inst, int element, const f2& uv) {
  const DMaterial& mat = scene.materials[inst.material];
  auto texcoord        = eval_texcoord(scene, inst, element, uv);
  auto emission_tex    = eval_texture(scene, mat.emission_tex, texcoord, true);
  auto color_shp       = eval_color(scene, inst, element, uv);
  auto color_tex       = eval_texture(scene, mat.color_tex, texcoord, true);
  auto roughness_tex   = eval_texture(scene, mat.roughness_tex, texcoord, false);
  auto scattering_tex  = eval_texture(scene, mat.scattering_tex, texcoord, true);

  mpoint p;
  p.type         = mat.type;
  p.emission     = mat.emission * xyz(emission_tex) * xyz(color_shp);
  p.color        = mat.color * xyz(color_tex) * xyz(color_shp);
  p.opacity      = mat.opacity * color_tex.w * color_shp.w;
  p.metallic     = mat.metallic * roughness_tex.z;
  p.roughness    = mat.roughness * roughness_tex.y;
  p.roughness    = p.roughness * p.roughness;
  p.ior          = mat.ior;
  p.scattering   = mat.scattering * xyz(scattering_tex);
  p.scanisotropy = mat.scanisotropy;
  p.trdepth      = mat.trdepth;
  if (mat.type == kRefractive || mat.type == kVolumetric || mat.type == kSubsurface) {
    p.density = -vlog(vclamp(p.color, 0.0001f, 1.0f)) / p.trdepth;
  } else {
    p.density = {0, 0, 0};
  }
  const float min_roughness = 0.03f * 0.03f;
  if (p.type == kMatte || p.type == kGltfPbr || p.type == kGlossy) {
    p.roughness = yclamp(p.roughness, min_roughness, 1.0f);
  } else if (mat.type == kVolumetric) {
    p.roughness = 0;
  } else {
    if (p.roughness < min_roughness) p.roughness = 0;
  }
  return p;
}
YGL_D bool is_volumetric_type(int type) {  // yocto_scene.cpp:257-261
  return type == kRefractive || type == kVolumetric || type == kSubsurface;
}

// eval_environment, yocto_scene.cpp:596-613
YGL_D_BIG f3 eval_environment(const DScene& scene, const f3& direction) {
  f3 emission = {0, 0, 0};
  for (int e = 0; e < scene.num_environments; e++) {
    const DEnvironment& env = scene.environments[e];
    auto wl       = transform_direction(env.inv_frame, direction);
    auto texcoord = f2{yatan2(wl.z, wl.x) / (2 * kPi), yacos(yclamp(wl.y, -1.0f, 1.0f)) / kPi};
    if (texcoord.x < 0) texcoord.x += 1;
    emission = emission + env.emission * xyz(eval_texture_raw(scene, env.emission_tex, texcoord));
  }
  return emission;
}

}  // namespace ygl
