// ygl_build.h — host-side preparation that precedes the hot path: BVH build in the reference's
// node/primitive order, leaf/instance packets, light CDFs and the per-pixel rng table.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ygl_b200.h"

namespace ygl {

struct HostTree {
  std::vector<ygl_bvh_node> nodes;
  std::vector<int32_t>      prims;
  int                       max_stack = 0;  // stack entries a traversal can need (depth + 1)
};

struct float4h {
  float x, y, z, w;
};
constexpr int kMaxTreeDepth        = 128;  // levels per tree the device walk accepts = the reference's stack (yocto_bvh.cpp:469)
#ifdef YGL_PAIR_VISIT
constexpr int kInstancePacketQuads = 8;  // float4 per DInstancePacket (ygl_scene.cuh): + the shape's root node
#else
constexpr int kInstancePacketQuads = 6;  // float4 per DInstancePacket (ygl_scene.cuh)
#endif
constexpr int kMaxTreeNodes        = 1 << 28, kMaxTreePrims = 1 << 26;  // ranges of the device node words

struct HostBvh {
  HostTree              top;
  std::vector<HostTree> shapes;
  std::vector<int>      shape_kind;  // element type each shape tree is built over (kElem*)
  // device-ready data (see ygl_scene.cuh)
  std::vector<std::vector<float4h>> shape_nodes;    // 2 float4 per node
  std::vector<std::vector<float4h>> shape_packets;  // leaf packets, primitive order
  std::vector<float4h>              top_nodes;
  std::vector<float4h>              top_packets;   // 8 float4 per instance, leaf order
  std::vector<float4h>              inst_packets;  // 8 float4 per instance, id order
};

struct HostLight {
  int                instance = -1, environment = -1;
  std::vector<float> cdf;
};

// make_scene_bvh, yocto_bvh.cpp:364-396 (+ make_shape_bvh :321-362, make_bvh :238-302).
// device_stream != null: trees over at least kDeviceBuildMin primitives are built on the current CUDA device
// (ygl_bvh_device.cu; split_middle only - highquality trees stay on the host), the others on the host cores.
constexpr int kDeviceBuildMin = 4096;
bool build_scene_bvh(const ygl_scene_desc& desc, bool highquality, HostBvh& out, std::string& error,
    void* device_stream = nullptr, bool use_device = false);
// one tree on the device, bit-identical to make_tree(bboxes, false) (boxes: n x {min.xyz, max.xyz})
bool build_tree_device(void* stream, const float* boxes, int n, HostTree& tree, std::string& error);
// the same device-ready data from trees built elsewhere (ygl_bvh_create_from_host): validated, then packed
bool adopt_scene_bvh(const ygl_scene_desc& desc, const ygl_bvh_node* top_nodes, int num_top_nodes,
    const int32_t* top_prims, int num_top_prims, const ygl_bvh_node* const* shape_nodes, const int* shape_num_nodes,
    const int32_t* const* shape_prims, const int* shape_num_prims, HostBvh& out, std::string& error);
// update_scene_bvh, yocto_bvh.cpp:434-451: refit after positions / radii / instance frames moved (topology kept)
bool update_scene_bvh(const ygl_scene_desc& desc, const int* updated_shapes, int num_updated_shapes, HostBvh& bvh,
    std::string& error);
// make_trace_lights, yocto_trace.cpp:1528-1581
void build_lights(const ygl_scene_desc& desc, std::vector<HostLight>& lights);
// image size + rng table of make_trace_state, yocto_trace.cpp:1495-1520
bool state_size(const ygl_scene_desc& desc, const ygl_trace_params& params, int& width, int& height,
    std::string& error);
void state_rngs(const ygl_trace_params& params, int width, int height, uint64_t* rngs);

}  // namespace ygl
