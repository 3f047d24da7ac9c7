// TEST INFRASTRUCTURE (reference-side): what a Yocto/GL application does with its scene file, done twice in one
// program - yocto::load_scene + tesselate_subdivs (the reference, libs/yocto/yocto_sceneio.cpp:2761,
// yocto_scene.cpp:807) and yocto::b200::load_scene (yocto-gl_b200/host/yocto_b200trace.h over ygl_scene_load) - and
// the two scene_data objects compared member by member, bit for bit. Host only: no device is touched.
// usage: shim_load_demo <scene file>...      exit code 0 = every file identical
#include <yocto/yocto_scene.h>
#include <yocto/yocto_sceneio.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../yocto-gl_b200/host/yocto_b200trace.h"

using namespace yocto;

template <typename T>
static bool same_bits(const std::vector<T>& a, const std::vector<T>& b) {
  return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0);
}
template <typename T>
static bool same_pod(const T& a, const T& b) {
  return memcmp(&a, &b, sizeof(T)) == 0;
}

static int compare(const scene_data& a, const scene_data& b) {
  int  bad  = 0;
  auto fail = [&](const char* what, size_t k) { bad++, printf("  differs: %s[%zu]\n", what, k); };
  if (a.cameras.size() != b.cameras.size() || a.instances.size() != b.instances.size() ||
      a.materials.size() != b.materials.size() || a.environments.size() != b.environments.size() ||
      a.shapes.size() != b.shapes.size() || a.textures.size() != b.textures.size())
    return printf("  differs: element counts\n"), 1;
  for (size_t k = 0; k < a.cameras.size(); k++) {
    auto &x = a.cameras[k], &y = b.cameras[k];
    if (!same_pod(x.frame, y.frame) || x.orthographic != y.orthographic || !same_pod(x.lens, y.lens) ||
        !same_pod(x.film, y.film) || !same_pod(x.aspect, y.aspect) || !same_pod(x.focus, y.focus) ||
        !same_pod(x.aperture, y.aperture))
      fail("cameras", k);
  }
  for (size_t k = 0; k < a.instances.size(); k++)
    if (!same_pod(a.instances[k].frame, b.instances[k].frame) || a.instances[k].shape != b.instances[k].shape ||
        a.instances[k].material != b.instances[k].material)
      fail("instances", k);
  for (size_t k = 0; k < a.environments.size(); k++)
    if (!same_pod(a.environments[k].frame, b.environments[k].frame) ||
        !same_pod(a.environments[k].emission, b.environments[k].emission) ||
        a.environments[k].emission_tex != b.environments[k].emission_tex)
      fail("environments", k);
  for (size_t k = 0; k < a.materials.size(); k++) {
    auto &x = a.materials[k], &y = b.materials[k];
    if (x.type != y.type || !same_pod(x.emission, y.emission) || !same_pod(x.color, y.color) ||
        !same_pod(x.roughness, y.roughness) || !same_pod(x.metallic, y.metallic) || !same_pod(x.ior, y.ior) ||
        !same_pod(x.scattering, y.scattering) || !same_pod(x.scanisotropy, y.scanisotropy) ||
        !same_pod(x.trdepth, y.trdepth) || !same_pod(x.opacity, y.opacity) || x.emission_tex != y.emission_tex ||
        x.color_tex != y.color_tex || x.roughness_tex != y.roughness_tex || x.scattering_tex != y.scattering_tex ||
        x.normal_tex != y.normal_tex)
      fail("materials", k);
  }
  for (size_t k = 0; k < a.textures.size(); k++) {
    auto &x = a.textures[k], &y = b.textures[k];
    if (x.width != y.width || x.height != y.height || x.linear != y.linear || x.nearest != y.nearest ||
        x.clamp != y.clamp || !same_bits(x.pixelsf, y.pixelsf) || !same_bits(x.pixelsb, y.pixelsb))
      fail("textures", k);
  }
  for (size_t k = 0; k < a.shapes.size(); k++) {
    auto &x = a.shapes[k], &y = b.shapes[k];
    if (!same_bits(x.points, y.points) || !same_bits(x.lines, y.lines) || !same_bits(x.triangles, y.triangles) ||
        !same_bits(x.quads, y.quads) || !same_bits(x.positions, y.positions) || !same_bits(x.normals, y.normals) ||
        !same_bits(x.texcoords, y.texcoords) || !same_bits(x.colors, y.colors) || !same_bits(x.radius, y.radius))
      fail("shapes", k);
  }
  if (a.camera_names != b.camera_names) fail("camera_names", 0);
  if (a.instance_names != b.instance_names) fail("instance_names", 0);
  if (a.material_names != b.material_names) fail("material_names", 0);
  if (a.environment_names != b.environment_names) fail("environment_names", 0);
  if (a.texture_names != b.texture_names) fail("texture_names", 0);
  if (a.shape_names != b.shape_names) fail("shape_names", 0);
  return bad;
}

int main(int argc, const char** argv) {
  int bad = 0;
  for (int k = 1; k < argc; k++) {
    auto theirs = scene_data{};
    auto error  = std::string{};
    if (!yocto::load_scene(argv[k], theirs, error, false)) return printf("reference cannot load %s: %s\n", argv[k], error.c_str()), 2;
    if (!theirs.subdivs.empty()) tesselate_subdivs(theirs);
    auto ours = scene_data{};
    if (!b200::load_scene(argv[k], ours, error)) return printf("drop-in cannot load %s: %s\n", argv[k], error.c_str()), 2;
    int differ = compare(ours, theirs);
    printf("%s: %zu shapes, %zu instances, %zu textures: %s\n", argv[k], ours.shapes.size(), ours.instances.size(),
        ours.textures.size(), differ ? "DIFFERENT" : "identical");
    bad += differ;
  }
  return bad ? 1 : 0;
}
