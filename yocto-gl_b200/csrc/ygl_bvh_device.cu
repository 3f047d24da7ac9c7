// ygl_bvh_device.cu — device build of one BVH tree that reproduces the reference's make_bvh with split_middle
// (libs/yocto/yocto_bvh.cpp:202-302) BIT FOR BIT: same nodes at the same indices, same `primitives` permutation.
// the step before the hot path; on scenes with large unique meshes (C2: 144 K triangles,
// C5: 262 K line segments per hairball) it is the time to the first pixel.
//
// What has to be reproduced, and how it is made parallel:
//  * The TOPOLOGY (which primitives end in which node) does not depend on the order in which nodes are processed:
//    the tree is grown level by level, all segments of a level at once.
//  * std::partition is not stable, so `primitives` depends on its algorithm. libstdc++'s bidirectional partition swaps
//    the k-th misplaced element from the left (predicate false, left of the final boundary) with the k-th misplaced
//    element from the right (predicate true, right of the boundary) and touches nothing else. Ranks come from one
//    prefix sum of the predicate; every pair is swapped by one thread.
//  * NODE INDICES. The reference pops work items from a LIFO and allocates a node's two children when it visits
//    the node: the child pair of an internal node sits at 1 + 2 * (number of internal nodes visited before it), and
//    the visit order is a pre-order that takes the RIGHT child first. With c(n) = internal nodes in n's subtree:
//    rank(right child) = rank(parent) + 1, rank(left child) = rank(parent) + 1 + c(right sibling). c() is a bottom-up
//    pass over the levels, rank() a top-down one.
//  * NODE BOXES. The reference folds merge(bbox, prim_bbox) over a node's primitives in order, with yocto's
//    min/max ((a<b)?a:b): among equal values (+0 / -0) the LAST one wins. Leaves fold their <= 4 primitives
//    exactly so; an internal node is merge(left, right) with the same min/max, which picks the right (= later)
//    operand on ties: by induction the bits are those of the sequential fold.
// The host build (ygl_build.cpp, verified bit-identical to the reference on the CPU) is the oracle of this file:
// tests/test_gpu_parity.py::test_device_bvh_build_matches_host_build.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ygl_build.h"

namespace ygl {

namespace {

constexpr int kMaxPrimsPerLeaf = 4;  // bvh_max_prims, yocto_bvh.cpp:236

struct TempNode {  // one record per node, in creation (level) order
  int start, end;    // primitive range
  int left;          // temp id of the left child (right = left + 1), -1 for a leaf
  int axis;
  int icount, rank;  // internal nodes in the subtree; rank among internal nodes in the reference's visit order
  int final_index;
};

struct Segment {  // an open node of the current level
  int   start, end, node;
  int   axis, mid, partition;  // filled by the split kernels
  float split;
  unsigned cmin[3], cmax[3];   // centroid box, order-preserving encoding
};

__device__ __forceinline__ unsigned encode_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float decode_ordered(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}
__device__ __forceinline__ float ymin_(float a, float b) { return (a < b) ? a : b; }  // yocto_math.h:1046-1047
__device__ __forceinline__ float ymax_(float a, float b) { return (a > b) ? a : b; }

#define YGL_BVH_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t err__ = (expr);                                                              \
    if (err__ != cudaSuccess) {                                                              \
      error = std::string("device bvh build: ") + cudaGetErrorString(err__) + " at " #expr; \
      return false;                                                                          \
    }                                                                                        \
  } while (0)

// centers[i] = (min + max) / 2, yocto_bvh.cpp:245 (center(bbox), yocto_geometry.h)
__global__ void k_centers(const float* __restrict__ boxes, int n, float* __restrict__ centers, int* __restrict__ prims,
    int* __restrict__ elem_seg, int* __restrict__ bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool finite = true;
  for (int c = 0; c < 3; c++) {
    float lo = boxes[6 * i + c], hi = boxes[6 * i + 3 + c];
    centers[3 * i + c] = __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
    finite             = finite && isfinite(lo) && isfinite(hi);
  }
  if (!finite) *bad = 1;
  prims[i]    = i;
  elem_seg[i] = 0;
}

// centroid box of every open segment (only its value matters, not the sign of a zero: see the file comment)
__global__ void k_centroid_bounds(const float* __restrict__ centers, const int* __restrict__ prims,
    const int* __restrict__ elem_seg, Segment* __restrict__ segs, int n) {
  int  i      = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid  = i < n;
  int  seg    = valid ? elem_seg[i] : -1;
  valid       = valid && seg >= 0;
  unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  unsigned lo[3], hi[3];
  int      p = prims[i];
  for (int c = 0; c < 3; c++) lo[c] = hi[c] = encode_ordered(centers[3 * p + c]);
  unsigned peers = __match_any_sync(active, seg);
  if (peers == active && __popc(active) > 1) {  // the whole warp sits in one segment: one atomic per warp and component
    for (int c = 0; c < 3; c++) {
      lo[c] = __reduce_min_sync(active, lo[c]);
      hi[c] = __reduce_max_sync(active, hi[c]);
    }
    if ((threadIdx.x & 31) != __ffs(active) - 1) return;
  }
  for (int c = 0; c < 3; c++) {
    atomicMin(&segs[seg].cmin[c], lo[c]);
    atomicMax(&segs[seg].cmax[c], hi[c]);
  }
}

// split_middle, yocto_bvh.cpp:202-232: axis and split position of every open segment
__global__ void k_choose_split(Segment* __restrict__ segs, int nsegs) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs) return;
  Segment& g = segs[s];
  float    mn[3], mx[3], size[3];
  for (int c = 0; c < 3; c++) {
    mn[c] = decode_ordered(g.cmin[c]), mx[c] = decode_ordered(g.cmax[c]);
    size[c] = __fsub_rn(mx[c], mn[c]);
  }
  g.partition = 1;
  if (size[0] == 0 && size[1] == 0 && size[2] == 0) {  // csize == zero3f: split in the middle, axis 0
    g.axis = 0, g.partition = 0, g.mid = (g.start + g.end) / 2, g.split = 0;
    return;
  }
  int axis = 0;
  if (size[0] >= size[1] && size[0] >= size[2]) axis = 0;
  if (size[1] >= size[0] && size[1] >= size[2]) axis = 1;
  if (size[2] >= size[0] && size[2] >= size[1]) axis = 2;
  g.axis  = axis;
  g.split = __fdiv_rn(__fadd_rn(mn[axis], mx[axis]), 2.0f);  // center(cbbox)[axis]
}

// predicate of the partition for every element of an open segment
__global__ void k_flags(const float* __restrict__ centers, const int* __restrict__ prims, const int* __restrict__ elem_seg,
    const Segment* __restrict__ segs, int n, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = elem_seg[i], f = 0;
  if (seg >= 0 && segs[seg].partition) f = centers[3 * prims[i] + segs[seg].axis] < segs[seg].split ? 1 : 0;
  flags[i] = f;
}

// ---- exclusive prefix sum over n ints (three small kernels; n is a few million at most) ----
constexpr int kScanBlock = 1024;
__global__ void k_scan_blocks(const int* __restrict__ in, int n, int* __restrict__ out, int* __restrict__ block_sums) {
  __shared__ int warp_sums[32];
  int  i = blockIdx.x * kScanBlock + threadIdx.x;
  int  v = i < n ? in[i] : 0, x = v;
  int  lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 1; off < 32; off <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int incl = x + (warp ? warp_sums[warp - 1] : 0);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = incl;
}
__global__ void k_scan_sums(int* __restrict__ block_sums, int nblocks) {  // single block, serial over chunks
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? block_sums[i] : 0, x = v;
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
      for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int incl = x + (warp ? warp_sums[warp - 1] : 0) + carry;
    if (i < nblocks) block_sums[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ block_sums, int* __restrict__ total) {
  int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) out[i] += block_sums[blockIdx.x];
  if (i == n - 1 && total) *total = 0;  // (unused: the scan array has n + 1 entries, the last is written by k_scan_last)
}
__global__ void k_scan_last(const int* __restrict__ in, int* __restrict__ out, int n) { out[n] = out[n - 1] + in[n - 1]; }

// boundary of every open segment; falls back to the midpoint when the predicate did not split it (yocto_bvh.cpp:229)
__global__ void k_boundaries(Segment* __restrict__ segs, int nsegs, const int* __restrict__ scan) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs) return;
  Segment& g = segs[s];
  if (!g.partition) return;
  int trues = scan[g.end] - scan[g.start];
  g.mid     = g.start + trues;
  if (g.mid == g.start || g.mid == g.end) g.mid = (g.start + g.end) / 2, g.partition = 0;  // std::partition moved nothing
}

// the k-th misplaced element from the right of every partitioned segment (predicate true, right of the boundary)
__global__ void k_misplaced_right(const int* __restrict__ elem_seg, const Segment* __restrict__ segs,
    const int* __restrict__ flags, const int* __restrict__ scan, int n, int* __restrict__ from_right) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = elem_seg[i];
  if (seg < 0 || !segs[seg].partition) return;
  const Segment& g = segs[seg];
  if (i >= g.mid && flags[i]) {
    int k = (scan[g.end] - scan[i]) - 1;  // trues to the right of i
    from_right[g.start + k] = i;
  }
}
// swap the k-th misplaced element from the left (predicate false, left of the boundary) with its partner
__global__ void k_swap_pairs(const int* __restrict__ elem_seg, const Segment* __restrict__ segs, const int* __restrict__ flags,
    const int* __restrict__ scan, int n, const int* __restrict__ from_right, int* __restrict__ prims) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = elem_seg[i];
  if (seg < 0 || !segs[seg].partition) return;
  const Segment& g = segs[seg];
  if (i < g.mid && !flags[i]) {
    int k = (i - g.start) - (scan[i] - scan[g.start]);  // falses to the left of i
    int j = from_right[g.start + k];
    int a = prims[i], b = prims[j];
    prims[i] = b, prims[j] = a;
  }
}

// close the level: every open segment becomes an internal node with two children; children with more than four
// primitives are the open segments of the next level
__global__ void k_emit_children(const Segment* __restrict__ segs, int nsegs, TempNode* __restrict__ nodes, int first_child,
    Segment* __restrict__ next, int* __restrict__ next_count, int* __restrict__ child_seg) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs) return;
  const Segment& g = segs[s];
  const int left   = first_child + 2 * s;
  nodes[g.node].left = left, nodes[g.node].axis = g.axis;
  for (int c = 0; c < 2; c++) {
    const int a = c ? g.mid : g.start, b = c ? g.end : g.mid;
    nodes[left + c] = TempNode{a, b, -1, 0, 0, 0, 0};
    int slot        = -1;
    if (b - a > kMaxPrimsPerLeaf) {
      slot           = atomicAdd(next_count, 1);
      Segment child  = {};
      child.start = a, child.end = b, child.node = left + c;
      for (int k = 0; k < 3; k++) child.cmin[k] = 0xffffffffu, child.cmax[k] = 0u;
      next[slot] = child;
    }
    child_seg[2 * s + c] = slot;
  }
}
__global__ void k_assign_segments(int* __restrict__ elem_seg, const Segment* __restrict__ segs, const int* __restrict__ child_seg,
    int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = elem_seg[i];
  if (seg < 0) return;
  elem_seg[i] = child_seg[2 * seg + (i >= segs[seg].mid ? 1 : 0)];
}

// ---- numbering: bottom-up internal counts, top-down ranks and final indices, one launch per level ----
__global__ void k_count_internal(TempNode* __restrict__ nodes, int first, int count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  TempNode& n = nodes[first + k];
  n.icount    = n.left < 0 ? 0 : 1 + nodes[n.left].icount + nodes[n.left + 1].icount;
}
__global__ void k_rank(TempNode* __restrict__ nodes, int first, int count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const TempNode& n = nodes[first + k];
  if (n.left < 0) return;
  TempNode &l = nodes[n.left], &r = nodes[n.left + 1];
  r.rank = n.rank + 1;                // the reference visits the right child first (LIFO, yocto_bvh.cpp:284-285)
  l.rank = n.rank + 1 + r.icount;
  l.final_index = 1 + 2 * n.rank, r.final_index = 2 + 2 * n.rank;
}
// final nodes (bvh_node, yocto_shape.h:474-480), boxes bottom-up
__global__ void k_write_nodes(const TempNode* __restrict__ nodes, int first, int count, const float* __restrict__ boxes,
    const int* __restrict__ prims, ygl_bvh_node* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const TempNode& n = nodes[first + k];
  ygl_bvh_node    o;
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  if (n.left < 0) {
    for (int i = n.start; i < n.end; i++) {
      const float* b = boxes + 6 * (size_t)prims[i];
      for (int c = 0; c < 3; c++) mn[c] = ymin_(mn[c], b[c]), mx[c] = ymax_(mx[c], b[3 + c]);
    }
    o.start = n.start, o.num = (int16_t)(n.end - n.start), o.axis = 0, o.internal = 0;
  } else {
    const ygl_bvh_node& l = out[nodes[n.left].final_index];
    const ygl_bvh_node& r = out[nodes[n.left + 1].final_index];
    for (int c = 0; c < 3; c++) {
      mn[c] = ymin_(ymin_(mn[c], l.bbox_min[c]), r.bbox_min[c]);
      mx[c] = ymax_(ymax_(mx[c], l.bbox_max[c]), r.bbox_max[c]);
    }
    o.start = 1 + 2 * n.rank, o.num = 2, o.axis = (int8_t)n.axis, o.internal = 1;
  }
  for (int c = 0; c < 3; c++) o.bbox_min[c] = mn[c], o.bbox_max[c] = mx[c];
  out[n.final_index] = o;
}

struct DeviceBuffers {
  std::vector<void*> ptrs;
  template <class T>
  bool alloc(T*& p, size_t count, std::string& error) {
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(1, count) * sizeof(T)) != cudaSuccess) return error = "device bvh build: out of memory", false;
    ptrs.push_back(q);
    p = (T*)q;
    return true;
  }
  ~DeviceBuffers() {
    for (auto p : ptrs) cudaFree(p);
  }
};

}  // namespace

// boxes: n x {min.xyz, max.xyz}. Returns the tree in the reference's layout (nodes, primitives, depth).
bool build_tree_device(void* stream_, const float* boxes, int n, HostTree& tree, std::string& error) {
  cudaStream_t s = (cudaStream_t)stream_;
  tree           = HostTree{};
  tree.prims.resize(n);
  if (n <= kMaxPrimsPerLeaf) return error = "device bvh build: tree too small (use the host build)", false;
  DeviceBuffers buf;
  float *   d_boxes, *d_centers;
  int *     d_prims, *d_elem_seg, *d_flags, *d_scan, *d_block_sums, *d_from_right, *d_child_seg, *d_counters;
  TempNode* d_nodes;
  Segment * d_segs[2];
  ygl_bvh_node* d_out;
  const int nblocks_scan = (n + kScanBlock - 1) / kScanBlock;
  const size_t max_nodes = 2 * (size_t)n + 1, max_segs = (size_t)n / 2 + 2;
  if (!buf.alloc(d_boxes, (size_t)n * 6, error) || !buf.alloc(d_centers, (size_t)n * 3, error) ||
      !buf.alloc(d_prims, n, error) || !buf.alloc(d_elem_seg, n, error) || !buf.alloc(d_flags, n, error) ||
      !buf.alloc(d_scan, (size_t)n + 1, error) || !buf.alloc(d_block_sums, nblocks_scan, error) ||
      !buf.alloc(d_from_right, n, error) || !buf.alloc(d_child_seg, 2 * max_segs, error) ||
      !buf.alloc(d_counters, 4, error) || !buf.alloc(d_nodes, max_nodes, error) || !buf.alloc(d_segs[0], max_segs, error) ||
      !buf.alloc(d_segs[1], max_segs, error) || !buf.alloc(d_out, max_nodes, error))
    return false;
  YGL_BVH_TRY(cudaMemcpyAsync(d_boxes, boxes, (size_t)n * 24, cudaMemcpyHostToDevice, s));
  YGL_BVH_TRY(cudaMemsetAsync(d_counters, 0, 4 * sizeof(int), s));
  const int T = 256, B = (n + T - 1) / T;
  k_centers<<<B, T, 0, s>>>(d_boxes, n, d_centers, d_prims, d_elem_seg, d_counters + 1);
  // root
  TempNode root = {0, n, -1, 0, 0, 0, 0};
  Segment  seg0 = {};
  seg0.start = 0, seg0.end = n, seg0.node = 0;
  for (int k = 0; k < 3; k++) seg0.cmin[k] = 0xffffffffu, seg0.cmax[k] = 0u;
  YGL_BVH_TRY(cudaMemcpyAsync(d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice, s));
  YGL_BVH_TRY(cudaMemcpyAsync(d_segs[0], &seg0, sizeof(seg0), cudaMemcpyHostToDevice, s));

  std::vector<std::pair<int, int>> levels = {{0, 1}};  // (first temp node, count) per level
  int nsegs = 1, num_nodes = 1, cur = 0;
  while (nsegs > 0) {
    if ((int)levels.size() > kMaxTreeDepth) return error = "BVH depth exceeds the traversal stack (128 entries, as in the reference)", false;
    Segment* segs = d_segs[cur];
    Segment* next = d_segs[1 - cur];
    const int SB  = (nsegs + T - 1) / T;
    k_centroid_bounds<<<B, T, 0, s>>>(d_centers, d_prims, d_elem_seg, segs, n);
    k_choose_split<<<SB, T, 0, s>>>(segs, nsegs);
    k_flags<<<B, T, 0, s>>>(d_centers, d_prims, d_elem_seg, segs, n, d_flags);
    k_scan_blocks<<<nblocks_scan, kScanBlock, 0, s>>>(d_flags, n, d_scan, d_block_sums);
    k_scan_sums<<<1, 1024, 0, s>>>(d_block_sums, nblocks_scan);
    k_scan_add<<<nblocks_scan, kScanBlock, 0, s>>>(d_scan, n, d_block_sums, nullptr);
    k_scan_last<<<1, 1, 0, s>>>(d_flags, d_scan, n);
    k_boundaries<<<SB, T, 0, s>>>(segs, nsegs, d_scan);
    k_misplaced_right<<<B, T, 0, s>>>(d_elem_seg, segs, d_flags, d_scan, n, d_from_right);
    k_swap_pairs<<<B, T, 0, s>>>(d_elem_seg, segs, d_flags, d_scan, n, d_from_right, d_prims);
    YGL_BVH_TRY(cudaMemsetAsync(d_counters, 0, sizeof(int), s));
    k_emit_children<<<SB, T, 0, s>>>(segs, nsegs, d_nodes, num_nodes, next, d_counters, d_child_seg);
    k_assign_segments<<<B, T, 0, s>>>(d_elem_seg, segs, d_child_seg, n);
    levels.push_back({num_nodes, 2 * nsegs});
    num_nodes += 2 * nsegs;
    int counters[2] = {0, 0};
    YGL_BVH_TRY(cudaMemcpyAsync(counters, d_counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    YGL_BVH_TRY(cudaStreamSynchronize(s));
    if (counters[1]) return error = "device bvh build: non-finite primitive bounds", false;
    nsegs = counters[0];
    cur   = 1 - cur;
  }
  // numbering and boxes
  for (int l = (int)levels.size() - 1; l >= 0; l--)
    k_count_internal<<<(levels[l].second + T - 1) / T, T, 0, s>>>(d_nodes, levels[l].first, levels[l].second);
  for (size_t l = 0; l < levels.size(); l++)
    k_rank<<<(levels[l].second + T - 1) / T, T, 0, s>>>(d_nodes, levels[l].first, levels[l].second);
  for (int l = (int)levels.size() - 1; l >= 0; l--)
    k_write_nodes<<<(levels[l].second + T - 1) / T, T, 0, s>>>(d_nodes, levels[l].first, levels[l].second, d_boxes, d_prims, d_out);
  tree.nodes.resize(num_nodes);
  YGL_BVH_TRY(cudaMemcpyAsync(tree.nodes.data(), d_out, (size_t)num_nodes * sizeof(ygl_bvh_node), cudaMemcpyDeviceToHost, s));
  YGL_BVH_TRY(cudaMemcpyAsync(tree.prims.data(), d_prims, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  YGL_BVH_TRY(cudaStreamSynchronize(s));
  YGL_BVH_TRY(cudaGetLastError());
  tree.max_stack = (int)levels.size();
  return true;
}

}  // namespace ygl
