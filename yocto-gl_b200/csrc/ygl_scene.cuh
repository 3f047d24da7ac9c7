// ygl_scene.cuh — device-resident scene arena (HBM layout) shared by all kernels.
//
// Everything the hot path reads lives in one arena per scene (see DESIGN.md "Data layout"):
//   * BVH nodes as 2 x float4 (32 B, 16-B aligned -> two LDG.128; the two children of a node are adjacent,
//     yocto_bvh.cpp:277-279, so one 64-byte fetch brings both):
//       n0 = {min.x, min.y, min.z, max.x}   n1 = {max.y, max.z, start(bits), word(bits)}
//       word = internal ? 0x80000000 | axis << 28 | first child : num << 26 | first primitive
//                                                            (bvh_node, yocto_shape.h:474-480)
//   * leaf primitives pre-gathered into float4 packets IN bvh.primitives ORDER, so a leaf's
//     range is one contiguous run of 128-bit loads (vertex indices + positions are not chased):
//       triangle 3 x float4 {p0.xyz,e1.x}{e1.yz,e2.xy}{e2.z,-,-,-}  with e1=p1-p0, e2=p2-p0
//       quad     4 x float4 {p0,-}{p1,-}{p2,-}{p3,-}
//       line     2 x float4 {p0,r0}{p1,r1}
//       point    1 x float4 {p,r}
//     (e1/e2 are pure functions of the inputs evaluated with the reference's own rounding, so
//      precomputing them cannot change a result bit.)
//   * instances as 6 x float4 packets {inverse frame (12 floats), shape, instance, kind, #nodes,
//     node/packet/primitive pointers of the shape's tree} (8 with the tree's root node in the pair-visit build): the
//     reference recomputes inverse(frame, true) per instance visit (yocto_bvh.cpp:602); it is a
//     pure function of the frame, so it is evaluated once on the host with the same arithmetic.
//   * original element/vertex arrays (reference layout) for the shading-side eval_* functions.
#pragma once

#include "ygl_shading.cuh"

namespace ygl {

// a ray parked by the extend kernel carries its traversal state to the next launch: 12 header words + its stack
constexpr int kSuspendStack = 24;                      // stack entries (t0, work word) a parked ray can carry
constexpr int kSuspendWords = 12 + 2 * kSuspendStack;  // words per save slot

enum : int { kElemNone = 0, kElemPoints = 1, kElemLines = 2, kElemTriangles = 3, kElemQuads = 4 };

struct DShape {
  // traversal side
  const float4* nodes;    // 2 float4 per node
  const float4* packets;  // leaf packets, primitive order
  const int*    prims;    // bvh.primitives
  int           num_nodes;
  int           bvh_kind;   // element type the tree was built over (points > lines > triangles > quads)
  int           eval_kind;  // element type eval_* uses (triangles > quads > lines > points)
  int           pad0;
  // shading side (reference arrays)
  const int*   points;
  const int*   lines;      // x2
  const int*   triangles;  // x3
  const int*   quads;      // x4
  const float* positions;  // x3
  const float* normals;    // x3 or null
  const float* texcoords;  // x2 or null
  const float* colors;     // x4 or null
  const float* radius;     // or null
  int          num_points, num_lines, num_triangles, num_quads;
};

#ifdef YGL_PAIR_VISIT
struct __align__(128) DInstancePacket {  // 128 B = one line; q6-q7: the root node of the shape's tree
  float4 q[8];
};
#else
struct DInstancePacket {  // 96 B = 6 x LDG.128, no dependent load to reach the shape's tree
  float4 q[6];  // q0-q2: inverse frame x,y,z,o (12 floats); q3: shape, instance, bvh kind, num nodes (int bits);
                // q4: nodes pointer, leaf-packets pointer; q5: primitives pointer, -, -
};
#endif

struct DInstance {  // shading side: forward frame + ids (instance_data, yocto_scene.h:145)
  frame3 frame;
  int    shape, material;
  int    pad0, pad1;
};

struct DMaterial {  // material_data, yocto_scene.h:123
  int   type;
  f3    emission, color;
  float roughness, metallic, ior;
  f3    scattering;
  float scanisotropy, trdepth, opacity;
  int   emission_tex, color_tex, roughness_tex, scattering_tex, normal_tex;
};

struct DTexture {  // texture_data, yocto_scene.h:95
  int           width, height;
  int           linear, nearest, clamp;
  int           pad;
  const float4* pixelsf;
  const uchar4* pixelsb;
};

struct DEnvironment {  // environment_data + inverse(frame) precomputed (rigid inverse)
  frame3 frame, inv_frame;
  f3     emission;
  int    emission_tex;
};

struct DCamera {  // camera_data, yocto_scene.h:83
  frame3 frame;
  int    orthographic;
  float  lens, film, aspect, focus, aperture;
};

struct DLight {  // trace_light, yocto_trace.h:126
  int          instance, environment;
  const float* cdf;
  int          cdf_size;
  int          pad;
};

struct DScene {
  const DCamera*         cameras;
  const DInstance*       instances;
  const DInstancePacket* inst_packets;  // by instance id (intersect_instance)
  const DMaterial*       materials;
  const DEnvironment*    environments;
  const DTexture*        textures;
  const DShape*          shapes;
  int num_cameras, num_instances, num_materials, num_environments, num_textures, num_shapes;
  // instance-level tree
  const float4*          top_nodes;
  const DInstancePacket* top_packets;  // leaf order
  const int*             top_prims;
  int                    top_num_nodes;
  // lights
  const DLight* lights;
  int           num_lights;
  // shading class of a hit on instance i = 1 + material type (ygl_kernels.cuh); null = shade queues are not binned
  const unsigned char* inst_class;
  int                  stack_mode;   // traversal stack the trees need: kStackShared / kStackShallow / kStackDeep (ygl_traverse.cuh)
  int                  has_volumes;  // some material can start a participating medium (refractive / subsurface / volumetric)
};

YGL_D f3 ld3(const float* p, int i) { return f3{__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)}; }
YGL_D f2 ld2(const float* p, int i) { return f2{__ldg(p + 2 * i), __ldg(p + 2 * i + 1)}; }
YGL_D f4 ld4(const float* p, int i) {
  float4 v = __ldg((const float4*)p + i);
  return f4{v.x, v.y, v.z, v.w};
}

}  // namespace ygl
