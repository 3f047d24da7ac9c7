// ygl_build.cpp — host-side preparation for the hot path (not timed, but its output order
// defines hit-id parity): the two-level BVH in the reference's node and primitive order
// (libs/yocto/yocto_bvh.cpp:108-396), the traversal packets described in ygl_scene.cuh, the
// light CDFs (yocto_trace.cpp:1528-1581) and the per-pixel rng table (yocto_trace.cpp:1495-1520).
// Compiled with -ffp-contract=off so every float op rounds exactly like the reference build.
#include "ygl_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <functional>
#include <sched.h>

#include <condition_variable>
#include <cstdio>
#include <limits>
#include <mutex>
#include <thread>

#include "ygl_sampling.cuh"

namespace ygl {

namespace {

// Cores this process may really use: the smaller of the affinity mask and the cgroup CPU quota (a container lease of
// 16 CPUs on a 128-thread host reports 128 from hardware_concurrency; starting 128 threads there only adds switching)
inline int host_parallelism() {
  static const int cores = []() {
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = std::min(n > 0 ? n : CPU_COUNT(&set), CPU_COUNT(&set));
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
      long long quota = 0, period = 0;
      if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0)
        n = std::min<long long>(n > 0 ? n : 1, (quota + period - 1) / period);
      fclose(f);
    }
    return std::max(1, n);
  }();
  return cores;
}

constexpr float kLowest = -kFltMax;  // flt_min = numeric_limits<float>::lowest(), yocto_math.h:78

struct box3 {
  f3 min = {kFltMax, kFltMax, kFltMax};  // invalidb3f, yocto_geometry.h:77-87
  f3 max = {kLowest, kLowest, kLowest};
};

box3 merge(const box3& a, const f3& b) { return {vmin(a.min, b), vmax(a.max, b)}; }
box3 merge(const box3& a, const box3& b) { return {vmin(a.min, b.min), vmax(a.max, b.max)}; }
f3   center(const box3& a) { return (a.min + a.max) / 2; }

f3 P(const float* positions, int i) { return {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]}; }

// primitive bounds, yocto_geometry.h:475-498
box3 point_bounds(const f3& p, float r) { return {vmin(p - r, p + r), vmax(p - r, p + r)}; }
box3 line_bounds(const f3& p0, const f3& p1, float r0, float r1) {
  return {vmin(p0 - r0, p1 - r1), vmax(p0 + r0, p1 + r1)};
}
box3 triangle_bounds(const f3& p0, const f3& p1, const f3& p2) {
  return {vmin(p0, vmin(p1, p2)), vmax(p0, vmax(p1, p2))};
}
box3 quad_bounds(const f3& p0, const f3& p1, const f3& p2, const f3& p3) {
  return {vmin(p0, vmin(p1, vmin(p2, p3))), vmax(p0, vmax(p1, vmax(p2, p3)))};
}

// split_middle, yocto_bvh.cpp:202-232
std::pair<int, int> split_middle(std::vector<int>& prims, const std::vector<f3>& centers, int start, int end) {
  box3 cb;
  for (int i = start; i < end; i++) cb = merge(cb, centers[prims[i]]);
  f3 csize = cb.max - cb.min;
  if (is_zero(csize)) return {(start + end) / 2, 0};
  int axis = 0;
  if (csize.x >= csize.y && csize.x >= csize.z) axis = 0;
  if (csize.y >= csize.x && csize.y >= csize.z) axis = 1;
  if (csize.z >= csize.x && csize.z >= csize.y) axis = 2;
  float split  = comp(center(cb), axis);
  int   middle = (int)(std::partition(prims.data() + start, prims.data() + end,
                         [&](int prim) { return comp(centers[prim], axis) < split; }) -
                     prims.data());
  if (middle == start || middle == end) return {(start + end) / 2, axis};
  return {middle, axis};
}

// split_sah, yocto_bvh.cpp:108-164 (16 bins per axis; cost normalised by the centroid box area)
// (`spread`, if given, runs the 45 independent (axis, bin) evaluations of a large node on several cores; each
// evaluation still folds its boxes over the primitives in order, and the minimum is taken in the reference's order)
using Spread = std::function<void(int, const std::function<void(int)>&)>;
std::pair<int, int> split_sah(std::vector<int>& prims, const std::vector<box3>& bboxes, const std::vector<f3>& centers,
    int start, int end, const Spread* spread = nullptr) {
  box3 cb;
  for (int i = start; i < end; i++) cb = merge(cb, centers[prims[i]]);
  f3 csize = cb.max - cb.min;
  if (is_zero(csize)) return {(start + end) / 2, 0};
  int       axis     = 0;
  const int nbins    = 16;
  float     split    = 0.0f;
  float     min_cost = kFltMax;
  auto      area     = [](const box3& b) {
    f3 size = b.max - b.min;
    return 1e-12f + 2 * size.x * size.y + 2 * size.x * size.z + 2 * size.y * size.z;
  };
  auto evaluate = [&](int saxis, int b, float& bsplit) {
    bsplit = comp(cb.min, saxis) + b * comp(csize, saxis) / nbins;
    box3 lbox, rbox;
    int  ln = 0, rn = 0;
    for (int i = start; i < end; i++) {
      if (comp(centers[prims[i]], saxis) < bsplit) {
        lbox = merge(lbox, bboxes[prims[i]]);
        ln += 1;
      } else {
        rbox = merge(rbox, bboxes[prims[i]]);
        rn += 1;
      }
    }
    return 1 + ln * area(lbox) / area(cb) + rn * area(rbox) / area(cb);
  };
  float costs[3][nbins], splits[3][nbins];
  if (spread) {
    (*spread)(3 * (nbins - 1), [&](int k) {
      const int saxis = k / (nbins - 1), b = 1 + k % (nbins - 1);
      costs[saxis][b] = evaluate(saxis, b, splits[saxis][b]);
    });
  }
  for (int saxis = 0; saxis < 3; saxis++) {
    for (int b = 1; b < nbins; b++) {
      float bsplit = 0;
      float cost   = spread ? (bsplit = splits[saxis][b], costs[saxis][b]) : evaluate(saxis, b, bsplit);
      if (cost < min_cost) {
        min_cost = cost;
        split    = bsplit;
        axis     = saxis;
      }
    }
  }
  int middle = (int)(std::partition(prims.data() + start, prims.data() + end,
                         [&](int prim) { return comp(centers[prim], axis) < split; }) -
                     prims.data());
  if (middle == start || middle == end) return {(start + end) / 2, axis};
  return {middle, axis};
}

// make_bvh, yocto_bvh.cpp:238-302: explicit LIFO of {node, start, end}; children allocated
// adjacently when the parent is visited; leaves hold <= 4 primitives.
constexpr int kMaxPrims = 4;
// the reference's loop over the primitives prims[first, last): `nodes` receives the subtree with its root at index 0
void build_range(std::vector<int>& prims, const std::vector<box3>& bboxes, const std::vector<f3>& centers, int first, int last,
    bool highquality, std::vector<ygl_bvh_node>& nodes) {
  struct item {
    int node, start, end;
  };
  std::vector<item> stack = {{0, first, last}};
  nodes.reserve((size_t)(last - first) * 2 + 1);
  nodes.push_back(ygl_bvh_node{{kFltMax, kFltMax, kFltMax}, {kLowest, kLowest, kLowest}, 0, 0, 0, 0});
  while (!stack.empty()) {
    item it = stack.back();
    stack.pop_back();
    box3 bb;
    for (int i = it.start; i < it.end; i++) bb = merge(bb, bboxes[prims[i]]);
    ygl_bvh_node node = {{bb.min.x, bb.min.y, bb.min.z}, {bb.max.x, bb.max.y, bb.max.z}, 0, 0, 0, 0};
    if (it.end - it.start > kMaxPrims) {
      auto [mid, axis] = highquality ? split_sah(prims, bboxes, centers, it.start, it.end)
                                     : split_middle(prims, centers, it.start, it.end);
      node.internal = 1;
      node.axis     = (int8_t)axis;
      node.num      = 2;
      node.start    = (int)nodes.size();
      nodes.push_back(ygl_bvh_node{{kFltMax, kFltMax, kFltMax}, {kLowest, kLowest, kLowest}, 0, 0, 0, 0});
      nodes.push_back(ygl_bvh_node{{kFltMax, kFltMax, kFltMax}, {kLowest, kLowest, kLowest}, 0, 0, 0, 0});
      stack.push_back({node.start + 0, it.start, mid});
      stack.push_back({node.start + 1, mid, it.end});
    } else {
      node.internal = 0;
      node.num      = (int16_t)(it.end - it.start);
      node.start    = it.start;
    }
    nodes[it.node] = node;
  }
}

// A few threads that stay up for the length of one large build: run(count, fn) hands fn(0..count-1) out to them and
// to the caller and returns when all are done.
class WorkCrew {
 public:
  explicit WorkCrew(int threads) {
    for (int t = 1; t < threads; t++) crew_.emplace_back([this]() { serve(); });
  }
  ~WorkCrew() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      quit_ = true;
    }
    wake_.notify_all();
    for (auto& t : crew_) t.join();
  }
  void run(int count, const std::function<void(int)>& fn) {
    if (count <= 0) return;
    {
      std::lock_guard<std::mutex> lock(mutex_);
      fn_ = &fn, count_ = count, next_ = 0, pending_ = count, round_++;
    }
    wake_.notify_all();
    work();
    std::unique_lock<std::mutex> lock(mutex_);
    done_.wait(lock, [&]() { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work() {
    while (true) {
      const std::function<void(int)>* fn = nullptr;
      int                             k  = 0;
      {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!fn_ || next_ >= count_) return;
        fn = fn_, k = next_++;
      }
      (*fn)(k);
      std::lock_guard<std::mutex> lock(mutex_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void serve() {
    uint64_t seen = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&]() { return quit_ || round_ != seen; });
        if (quit_) return;
        seen = round_;
      }
      work();
    }
  }
  std::vector<std::thread>        crew_;
  std::mutex                      mutex_;
  std::condition_variable         wake_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int                             count_ = 0, next_ = 0, pending_ = 0;
  uint64_t                        round_ = 0;
  bool                            quit_  = false;
};

constexpr int kParallelBuildMin = 32768;  // primitives from which one tree is built on several cores

// The same tree on several cores. The reference's loop is a pre-order walk that visits a node's right child first and
// gives a node's two children the next two free slots when it visits the node. Two consequences make the result
// reproducible out of order: (1) what a node does depends only on the primitives in its own range, which only its
// ancestors have permuted - disjoint subtrees can be built at the same time; (2) everything below a node is visited, and
// therefore numbered, in one uninterrupted stretch - a subtree built on its own has its descendants in one block, in the
// order the full walk would give them. So: split the nodes near the root one after the other (for SAH with the 45
// candidate evaluations of a node spread over the cores), build the subtrees below a size limit independently with the
// serial loop, then replay the walk over the few top nodes to hand out slots and copy each subtree's block behind it.
HostTree make_tree_parallel(const std::vector<box3>& bboxes, bool highquality, int threads) {
  HostTree  tree;
  const int n = (int)bboxes.size();
  tree.prims.resize(n);
  for (int i = 0; i < n; i++) tree.prims[i] = i;
  std::vector<f3> centers(n);
  for (int i = 0; i < n; i++) centers[i] = center(bboxes[i]);
  WorkCrew     crew(threads);
  const Spread spread = [&](int count, const std::function<void(int)>& fn) { crew.run(count, fn); };
  struct Top {
    int          start, end, left = -1, right = -1, subtree = -1;
    ygl_bvh_node node;
  };
  std::vector<Top> top = {{0, n}};
  const int        limit = std::max(4096, n / (threads * 8));
  std::vector<int> roots;  // top entries that are built as independent subtrees
  for (size_t t = 0; t < top.size(); t++) {
    const int first = top[t].start, last = top[t].end;
    if (last - first <= limit) {
      top[t].subtree = (int)roots.size();
      roots.push_back((int)t);
      continue;
    }
    box3 bb;
    for (int i = first; i < last; i++) bb = merge(bb, bboxes[tree.prims[i]]);
    auto [mid, axis] = highquality ? split_sah(tree.prims, bboxes, centers, first, last, &spread)
                                   : split_middle(tree.prims, centers, first, last);
    top[t].node      = {{bb.min.x, bb.min.y, bb.min.z}, {bb.max.x, bb.max.y, bb.max.z}, 0, 2, (int8_t)axis, 1};
    const int left = (int)top.size();
    top.push_back({first, mid});
    top.push_back({mid, last});
    top[t].left = left, top[t].right = left + 1;
  }
  std::vector<std::vector<ygl_bvh_node>> subtrees(roots.size());
  // largest first: the ranges differ in size
  std::vector<int> order(roots.size());
  for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
  std::stable_sort(order.begin(), order.end(),
      [&](int a, int b) { return top[roots[a]].end - top[roots[a]].start > top[roots[b]].end - top[roots[b]].start; });
  crew.run((int)roots.size(), [&](int k) {
    const Top& root = top[roots[order[k]]];
    build_range(tree.prims, bboxes, centers, root.start, root.end, highquality, subtrees[order[k]]);
  });
  // replay the walk over the top nodes
  size_t total = top.size() - roots.size();
  for (auto& nodes : subtrees) total += nodes.size();
  tree.nodes.resize(total);
  std::vector<std::pair<int, int>> stack = {{0, 0}};  // (top entry, slot)
  int                              cursor = 1;
  while (!stack.empty()) {
    auto [t, slot] = stack.back();
    stack.pop_back();
    if (top[t].subtree < 0) {
      ygl_bvh_node node = top[t].node;
      node.start        = cursor;
      cursor += 2;
      tree.nodes[slot] = node;
      stack.push_back({top[t].left, node.start});
      stack.push_back({top[t].right, node.start + 1});
      continue;
    }
    const auto& nodes = subtrees[top[t].subtree];
    const int   base  = cursor - 1;  // local index k >= 1 lands at base + k
    for (size_t k = 0; k < nodes.size(); k++) {
      ygl_bvh_node node = nodes[k];
      if (node.internal) node.start += base;
      tree.nodes[k == 0 ? slot : base + (int)k] = node;
    }
    cursor += (int)nodes.size() - 1;
  }
  return tree;
}

HostTree make_tree(const std::vector<box3>& bboxes, bool highquality, int threads = 1) {
  HostTree tree;
  if (threads > 1 && (int)bboxes.size() >= kParallelBuildMin) {
    tree = make_tree_parallel(bboxes, highquality, threads);
  } else {
    tree.prims.resize(bboxes.size());
    for (size_t i = 0; i < bboxes.size(); i++) tree.prims[i] = (int)i;
    std::vector<f3> centers(bboxes.size());
    for (size_t i = 0; i < bboxes.size(); i++) centers[i] = center(bboxes[i]);
    build_range(tree.prims, bboxes, centers, 0, (int)bboxes.size(), highquality, tree.nodes);
  }
  // stack need of a traversal = deepest level + 1
  if (!tree.nodes.empty()) {
    std::vector<std::pair<int, int>> todo = {{0, 1}};
    while (!todo.empty()) {
      auto [n, depth] = todo.back();
      todo.pop_back();
      tree.max_stack = std::max(tree.max_stack, depth);
      if (tree.nodes[n].internal) {
        todo.push_back({tree.nodes[n].start, depth + 1});
        todo.push_back({tree.nodes[n].start + 1, depth + 1});
      }
    }
  }
  return tree;
}

// run fn(0..n-1) on the host cores (dynamic: shapes differ wildly in size)
void parallel_shapes(int n, const std::function<void(int)>& fn) {
  const int nthreads = std::max(1, std::min<int>(n, host_parallelism()));
  if (nthreads <= 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<int>         next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([&]() {
      for (int i = next++; i < n; i = next++) fn(i);
    });
  for (auto& t : pool) t.join();
}

float as_float(int v) {
  float f;
  memcpy(&f, &v, 4);
  return f;
}

// node word (ygl_scene.cuh): internal = 0x80000000 | axis << 28 | first child (28 bits); leaf = num << 26 | first
// primitive (26 bits). make_tree / measure_tree guarantee the ranges (kMaxTreeNodes, kMaxTreePrims, leaves <= 4).
int node_word(const ygl_bvh_node& n) {
  if (n.internal) return (int)(0x80000000u | ((unsigned)(n.axis & 3) << 28) | ((unsigned)n.start & 0x0fffffffu));
  return (int)(((unsigned)n.num & 7u) << 26 | ((unsigned)n.start & 0x03ffffffu));
}
std::vector<float4h> pack_nodes(const HostTree& tree) {
  std::vector<float4h> out(tree.nodes.size() * 2);
  for (size_t i = 0; i < tree.nodes.size(); i++) {
    auto& n        = tree.nodes[i];
    out[2 * i + 0] = {n.bbox_min[0], n.bbox_min[1], n.bbox_min[2], n.bbox_max[0]};
    out[2 * i + 1] = {n.bbox_max[1], n.bbox_max[2], as_float(n.start), as_float(node_word(n))};
  }
  return out;
}
bool tree_fits_node_words(const HostTree& tree, std::string& error) {
  if (tree.nodes.size() > (size_t)kMaxTreeNodes || tree.prims.size() > (size_t)kMaxTreePrims)
    return error = "tree too large for the device node words (2^28 nodes, 2^26 primitives per tree)", false;
  return true;
}

frame3 to_frame(const ygl_frame3f& f) {
  return {{f.x[0], f.x[1], f.x[2]}, {f.y[0], f.y[1], f.y[2]}, {f.z[0], f.z[1], f.z[2]}, {f.o[0], f.o[1], f.o[2]}};
}

void pack_instance(float4h* out, const ygl_instance& inst, int id, int kind, int num_nodes, const float4h* root) {
  frame3 inv = frame_inverse(to_frame(inst.frame), true);  // inverse(frame, true), yocto_bvh.cpp:602
  out[0]     = {inv.x.x, inv.x.y, inv.x.z, inv.y.x};
  out[1]     = {inv.y.y, inv.y.z, inv.z.x, inv.z.y};
  out[2]     = {inv.z.z, inv.o.x, inv.o.y, inv.o.z};
  out[3]     = {as_float(inst.shape), as_float(id), as_float(kind), as_float(num_nodes)};
  out[4]     = {0, 0, 0, 0};  // device pointers of the shape's tree: patched at upload (ygl_api.cpp)
  out[5]     = {0, 0, 0, 0};
  if (kInstancePacketQuads == 8) {  // pair-visit build only: the root node of the shape's tree (object space), so that
    out[6] = root[0];               // entering an instance tests it without another dependent load
    out[7] = root[1];
  }
}

}  // namespace

namespace {

// element type a shape's tree is built over, make_shape_bvh, yocto_bvh.cpp:321-362: points > lines > triangles > quads
int shape_bvh_kind(const ygl_shape& s) {
  return s.num_points > 0 ? 1 : s.num_lines > 0 ? 2 : s.num_triangles > 0 ? 3 : s.num_quads > 0 ? 4 : 0;
}

bool shape_bounds(const ygl_shape& s, int si, int kind, std::vector<box3>& bboxes, std::string& error) {
  auto check = [&](int v) { return v >= 0 && v < s.num_positions; };
  if (kind == 1) {
    bboxes.resize(s.num_points);
    if (s.num_radius < s.num_positions) return error = "shape " + std::to_string(si) + ": points need a radius per vertex", false;
    for (int i = 0; i < s.num_points; i++) {
      int p = s.points[i];
      if (!check(p)) return error = "point index out of range", false;
      bboxes[i] = point_bounds(P(s.positions, p), s.radius[p]);
    }
  } else if (kind == 2) {
    bboxes.resize(s.num_lines);
    if (s.num_radius < s.num_positions) return error = "shape " + std::to_string(si) + ": lines need a radius per vertex", false;
    for (int i = 0; i < s.num_lines; i++) {
      int a = s.lines[2 * i], b = s.lines[2 * i + 1];
      if (!check(a) || !check(b)) return error = "line index out of range", false;
      bboxes[i] = line_bounds(P(s.positions, a), P(s.positions, b), s.radius[a], s.radius[b]);
    }
  } else if (kind == 3) {
    bboxes.resize(s.num_triangles);
    for (int i = 0; i < s.num_triangles; i++) {
      int a = s.triangles[3 * i], b = s.triangles[3 * i + 1], c = s.triangles[3 * i + 2];
      if (!check(a) || !check(b) || !check(c)) return error = "triangle index out of range", false;
      bboxes[i] = triangle_bounds(P(s.positions, a), P(s.positions, b), P(s.positions, c));
    }
  } else if (kind == 4) {
    bboxes.resize(s.num_quads);
    for (int i = 0; i < s.num_quads; i++) {
      const int* q = s.quads + 4 * i;
      if (!check(q[0]) || !check(q[1]) || !check(q[2]) || !check(q[3])) return error = "quad index out of range", false;
      bboxes[i] = quad_bounds(P(s.positions, q[0]), P(s.positions, q[1]), P(s.positions, q[2]), P(s.positions, q[3]));
    }
  }
  return true;
}

// leaf packets in primitive order (ygl_scene.cuh)
std::vector<float4h> pack_leaves(const ygl_shape& s, int kind, const HostTree& tree) {
  std::vector<float4h> pk;
  if (kind == 3) {
    pk.resize(tree.prims.size() * 3);
    for (size_t k = 0; k < tree.prims.size(); k++) {
      const int* t  = s.triangles + 3 * tree.prims[k];
      f3         p0 = P(s.positions, t[0]), e1 = P(s.positions, t[1]) - p0, e2 = P(s.positions, t[2]) - p0;
      pk[3 * k + 0] = {p0.x, p0.y, p0.z, e1.x};
      pk[3 * k + 1] = {e1.y, e1.z, e2.x, e2.y};
      pk[3 * k + 2] = {e2.z, 0, 0, 0};
    }
  } else if (kind == 4) {
    pk.resize(tree.prims.size() * 4);
    for (size_t k = 0; k < tree.prims.size(); k++) {
      const int* q = s.quads + 4 * tree.prims[k];
      for (int c = 0; c < 4; c++) {
        f3 p          = P(s.positions, q[c]);
        pk[4 * k + c] = {p.x, p.y, p.z, 0};
      }
    }
  } else if (kind == 2) {
    pk.resize(tree.prims.size() * 2);
    for (size_t k = 0; k < tree.prims.size(); k++) {
      const int* l  = s.lines + 2 * tree.prims[k];
      f3         p0 = P(s.positions, l[0]), p1 = P(s.positions, l[1]);
      pk[2 * k + 0] = {p0.x, p0.y, p0.z, s.radius[l[0]]};
      pk[2 * k + 1] = {p1.x, p1.y, p1.z, s.radius[l[1]]};
    }
  } else if (kind == 1) {
    pk.resize(tree.prims.size());
    for (size_t k = 0; k < tree.prims.size(); k++) {
      int p = s.points[tree.prims[k]];
      f3  v = P(s.positions, p);
      pk[k] = {v.x, v.y, v.z, s.radius[p]};
    }
  }
  return pk;
}

// instance boxes, yocto_bvh.cpp:382-389 (transform_bbox: merge of the 8 transformed corners)
bool instance_bounds(const ygl_scene_desc& desc, const HostBvh& out, std::vector<box3>& ibox, std::string& error) {
  ibox.resize(desc.num_instances);
  for (int i = 0; i < desc.num_instances; i++) {
    const ygl_instance& inst = desc.instances[i];
    if (inst.shape < 0 || inst.shape >= desc.num_shapes) return error = "instance shape id out of range", false;
    if (inst.material < 0 || inst.material >= desc.num_materials)
      return error = "instance material id out of range", false;
    const HostTree& st = out.shapes[inst.shape];  // never empty: make_bvh always emits a root
    auto&  n  = st.nodes[0];
    f3     mn = {n.bbox_min[0], n.bbox_min[1], n.bbox_min[2]}, mx = {n.bbox_max[0], n.bbox_max[1], n.bbox_max[2]};
    frame3 fr = to_frame(inst.frame);
    f3     corners[8] = {{mn.x, mn.y, mn.z}, {mn.x, mn.y, mx.z}, {mn.x, mx.y, mn.z}, {mn.x, mx.y, mx.z},
            {mx.x, mn.y, mn.z}, {mx.x, mn.y, mx.z}, {mx.x, mx.y, mn.z}, {mx.x, mx.y, mx.z}};
    box3   xf;
    for (auto& c : corners) xf = merge(xf, transform_point(fr, c));
    ibox[i] = xf;
  }
  return true;
}

// the device-ready data of the instance level (node float4s + instance packets)
void pack_top(const ygl_scene_desc& desc, HostBvh& out) {
  out.top_nodes = pack_nodes(out.top);
  auto pack = [&](float4h* dst, int id) {
    const ygl_instance& inst = desc.instances[id];
    pack_instance(dst, inst, id, out.shape_kind[inst.shape], (int)out.shapes[inst.shape].nodes.size(),
        out.shape_nodes[inst.shape].data());
  };
  out.top_packets.resize(out.top.prims.size() * kInstancePacketQuads);
  for (size_t k = 0; k < out.top.prims.size(); k++) pack(&out.top_packets[kInstancePacketQuads * k], out.top.prims[k]);
  out.inst_packets.resize((size_t)desc.num_instances * kInstancePacketQuads);
  for (int i = 0; i < desc.num_instances; i++) pack(&out.inst_packets[kInstancePacketQuads * (size_t)i], i);
}

// depth of a tree (stack entries a traversal can need) + structural checks of an adopted tree
bool measure_tree(HostTree& tree, int num_prims, const char* what, std::string& error) {
  tree.max_stack = 0;
  if (tree.nodes.empty()) return error = std::string(what) + ": a tree needs at least its root", false;
  if ((int)tree.prims.size() != num_prims) return error = std::string(what) + ": primitive count differs from the shape", false;
  for (int v : tree.prims)
    if (v < 0 || v >= num_prims) return error = std::string(what) + ": primitive id out of range", false;
  std::vector<std::pair<int, int>> todo = {{0, 1}};
  size_t visited = 0;
  while (!todo.empty()) {
    auto [n, depth] = todo.back();
    todo.pop_back();
    if (++visited > tree.nodes.size()) return error = std::string(what) + ": node graph is not a tree", false;
    tree.max_stack = std::max(tree.max_stack, depth);
    const ygl_bvh_node& node = tree.nodes[n];
    if (node.internal) {
      if (node.start < 1 || node.start + 1 >= (int)tree.nodes.size()) return error = std::string(what) + ": child index out of range", false;
      todo.push_back({node.start, depth + 1});
      todo.push_back({node.start + 1, depth + 1});
    } else {
      if (node.num < 0 || node.num > 4 || node.start < 0 || node.start + node.num > num_prims)
        return error = std::string(what) + ": leaf range invalid (leaves hold <= 4 primitives, yocto_bvh.cpp:236)", false;
    }
  }
  if (tree.max_stack > kMaxTreeDepth)
    return error = std::string(what) + ": BVH depth " + std::to_string(tree.max_stack) +
                   " exceeds the traversal stack (128 entries, as in the reference)", false;
  return true;
}

}  // namespace

bool build_scene_bvh(const ygl_scene_desc& desc, bool highquality, HostBvh& out, std::string& error, void* device_stream,
    bool use_device) {
  const int nshapes = desc.num_shapes;
  out.shapes.resize(nshapes);
  out.shape_kind.assign(nshapes, 0);
  out.shape_nodes.resize(nshapes);
  out.shape_packets.resize(nshapes);
  use_device = use_device && !highquality;
  static_assert(sizeof(box3) == 24, "box3 is {min.xyz, max.xyz}");
  auto element_count = [](const ygl_shape& s) {
    return s.num_points > 0 ? s.num_points : s.num_lines > 0 ? s.num_lines : s.num_triangles > 0 ? s.num_triangles : s.num_quads;
  };
  // large trees first, one after the other, on the device; meanwhile nothing else runs, the small ones follow on the host
  std::vector<std::string> errors(nshapes);
  std::vector<char>        done(nshapes, 0);
  if (use_device) {
    for (int si = 0; si < nshapes; si++) {
      const ygl_shape& s = desc.shapes[si];
      if (element_count(s) < kDeviceBuildMin) continue;
      const int         kind = shape_bvh_kind(s);
      std::vector<box3> bboxes;
      if (!shape_bounds(s, si, kind, bboxes, error)) return false;
      out.shape_kind[si] = kind;
      HostTree& tree     = out.shapes[si];
      if (!build_tree_device(device_stream, (const float*)bboxes.data(), (int)bboxes.size(), tree, error)) return false;
      if (!tree_fits_node_words(tree, error)) return false;
      out.shape_nodes[si]   = pack_nodes(tree);
      out.shape_packets[si] = pack_leaves(s, kind, tree);
      done[si]              = 1;
    }
  }
  // one tree per shape: independent, built on all host cores (the reference does the same, yocto_bvh.cpp:376-378);
  // the large ones first, one after the other, each on all cores (make_tree_parallel), then the small ones side by side
  const int host_threads = host_parallelism();
  auto build_shape = [&](int si, int threads) {
    const ygl_shape&  s    = desc.shapes[si];
    const int         kind = shape_bvh_kind(s);
    std::vector<box3> bboxes;
    if (!shape_bounds(s, si, kind, bboxes, errors[si])) return;
    out.shape_kind[si] = kind;
    HostTree& tree     = out.shapes[si];
    tree               = make_tree(bboxes, highquality, threads);
    // (an empty shape still gets a single empty root leaf, exactly like the reference's make_bvh)
    if (tree.max_stack > kMaxTreeDepth) {
      errors[si] = "shape " + std::to_string(si) + ": BVH depth " + std::to_string(tree.max_stack) +
                   " exceeds the traversal stack (128 entries, as in the reference)";
      return;
    }
    if (!tree_fits_node_words(tree, errors[si])) return;
    out.shape_nodes[si]   = pack_nodes(tree);
    out.shape_packets[si] = pack_leaves(s, kind, tree);
  };
  if (host_threads > 1)
    for (int si = 0; si < nshapes; si++)
      if (!done[si] && element_count(desc.shapes[si]) >= kParallelBuildMin) build_shape(si, host_threads), done[si] = 1;
  parallel_shapes(nshapes, [&](int si) {
    if (!done[si]) build_shape(si, 1);
  });
  for (auto& e : errors)
    if (!e.empty()) return error = e, false;

  std::vector<box3> ibox;
  if (!instance_bounds(desc, out, ibox, error)) return false;
  if (use_device && (int)ibox.size() >= kDeviceBuildMin) {
    if (!build_tree_device(device_stream, (const float*)ibox.data(), (int)ibox.size(), out.top, error)) return false;
  } else {
    out.top = make_tree(ibox, highquality, host_threads);
  }
  if (out.top.max_stack > kMaxTreeDepth) return error = "instance BVH too deep for the traversal stack (128 entries)", false;
  if (!tree_fits_node_words(out.top, error)) return false;
  pack_top(desc, out);
  return true;
}

// refit_bvh, yocto_bvh.cpp:303-318: boxes again from the leaves up; children always follow their parent in the node
// array, so one pass from the last node to the root sees every child before its parent
static void refit_tree(HostTree& tree, const std::vector<box3>& bboxes) {
  for (int id = (int)tree.nodes.size() - 1; id >= 0; id--) {
    ygl_bvh_node& node = tree.nodes[id];
    box3          box;
    if (node.internal) {
      for (int k = 0; k < 2; k++) {
        const ygl_bvh_node& c = tree.nodes[node.start + k];
        box = merge(box, box3{{c.bbox_min[0], c.bbox_min[1], c.bbox_min[2]}, {c.bbox_max[0], c.bbox_max[1], c.bbox_max[2]}});
      }
    } else {
      for (int k = 0; k < node.num; k++) box = merge(box, bboxes[tree.prims[node.start + k]]);
    }
    node.bbox_min[0] = box.min.x, node.bbox_min[1] = box.min.y, node.bbox_min[2] = box.min.z;
    node.bbox_max[0] = box.max.x, node.bbox_max[1] = box.max.y, node.bbox_max[2] = box.max.z;
  }
}

// update_scene_bvh, yocto_bvh.cpp:434-451 (+ update_shape_bvh :398-432): same topology, new boxes. Like the reference,
// every instance box is recomputed whatever `updated_instances` holds.
bool update_scene_bvh(const ygl_scene_desc& desc, const int* updated_shapes, int num_updated_shapes, HostBvh& bvh,
    std::string& error) {
  if ((int)bvh.shapes.size() != desc.num_shapes) return error = "bvh was built for a scene with another shape count", false;
  if ((int)bvh.top.prims.size() != desc.num_instances)
    return error = "bvh was built for a scene with another instance count", false;
  for (int k = 0; k < num_updated_shapes; k++) {
    const int si = updated_shapes[k];
    if (si < 0 || si >= desc.num_shapes) return error = "updated shape id out of range", false;
    const ygl_shape& s    = desc.shapes[si];
    const int        kind = shape_bvh_kind(s);
    std::vector<box3> bboxes;
    if (!shape_bounds(s, si, kind, bboxes, error)) return false;
    HostTree& tree = bvh.shapes[si];
    if (kind != bvh.shape_kind[si] || bboxes.size() != tree.prims.size())
      return error = "shape " + std::to_string(si) + ": element type or count changed since the build (refit keeps the topology)", false;
    refit_tree(tree, bboxes);
    bvh.shape_nodes[si]   = pack_nodes(tree);
    bvh.shape_packets[si] = pack_leaves(s, kind, tree);
  }
  std::vector<box3> ibox;
  if (!instance_bounds(desc, bvh, ibox, error)) return false;
  refit_tree(bvh.top, ibox);
  pack_top(desc, bvh);
  return true;
}

// Adopt trees built elsewhere (the reference's make_scene_bvh output): checks + packets only.
bool adopt_scene_bvh(const ygl_scene_desc& desc, const ygl_bvh_node* top_nodes, int num_top_nodes,
    const int32_t* top_prims, int num_top_prims, const ygl_bvh_node* const* shape_nodes, const int* shape_num_nodes,
    const int32_t* const* shape_prims, const int* shape_num_prims, HostBvh& out, std::string& error) {
  const int nshapes = desc.num_shapes;
  out.shapes.resize(nshapes);
  out.shape_kind.assign(nshapes, 0);
  out.shape_nodes.resize(nshapes);
  out.shape_packets.resize(nshapes);
  for (int si = 0; si < nshapes; si++) {
    const ygl_shape&  s    = desc.shapes[si];
    const int         kind = shape_bvh_kind(s);
    std::vector<box3> bboxes;  // only to validate the element indices
    if (!shape_bounds(s, si, kind, bboxes, error)) return false;
    out.shape_kind[si] = kind;
    HostTree& tree     = out.shapes[si];
    tree.nodes.assign(shape_nodes[si], shape_nodes[si] + shape_num_nodes[si]);
    tree.prims.assign(shape_prims[si], shape_prims[si] + shape_num_prims[si]);
    std::string what = "shape " + std::to_string(si);
    if (!measure_tree(tree, (int)bboxes.size(), what.c_str(), error) || !tree_fits_node_words(tree, error)) return false;
    out.shape_nodes[si]   = pack_nodes(tree);
    out.shape_packets[si] = pack_leaves(s, kind, tree);
  }
  std::vector<box3> ibox;  // validates instance ids
  if (!instance_bounds(desc, out, ibox, error)) return false;
  out.top.nodes.assign(top_nodes, top_nodes + num_top_nodes);
  out.top.prims.assign(top_prims, top_prims + num_top_prims);
  if (!measure_tree(out.top, desc.num_instances, "instance tree", error) || !tree_fits_node_words(out.top, error)) return false;
  pack_top(desc, out);
  return true;
}

// make_trace_lights, yocto_trace.cpp:1528-1581. Host libm (glibc sinf) as in the reference.
void build_lights(const ygl_scene_desc& desc, std::vector<HostLight>& lights) {
  lights.clear();
  auto tri_area = [](const f3& p0, const f3& p1, const f3& p2) { return length(cross(p1 - p0, p2 - p0)) / 2; };
  for (int i = 0; i < desc.num_instances; i++) {
    const ygl_instance& inst = desc.instances[i];
    const ygl_material& mat  = desc.materials[inst.material];
    if (mat.emission[0] == 0 && mat.emission[1] == 0 && mat.emission[2] == 0) continue;
    const ygl_shape& s = desc.shapes[inst.shape];
    if (s.num_triangles == 0 && s.num_quads == 0) continue;
    HostLight light;
    light.instance = i;
    if (s.num_triangles > 0) {
      light.cdf.resize(s.num_triangles);
      for (int e = 0; e < s.num_triangles; e++) {
        const int* t = s.triangles + 3 * e;
        light.cdf[e] = tri_area(P(s.positions, t[0]), P(s.positions, t[1]), P(s.positions, t[2]));
        if (e != 0) light.cdf[e] += light.cdf[e - 1];
      }
    }
    if (s.num_quads > 0) {
      light.cdf.assign(s.num_quads, 0.0f);
      for (int e = 0; e < s.num_quads; e++) {
        const int* q = s.quads + 4 * e;
        f3 p0 = P(s.positions, q[0]), p1 = P(s.positions, q[1]), p2 = P(s.positions, q[2]), p3 = P(s.positions, q[3]);
        light.cdf[e] = tri_area(p0, p1, p3) + tri_area(p2, p3, p1);
        if (e != 0) light.cdf[e] += light.cdf[e - 1];
      }
    }
    lights.push_back(std::move(light));
  }
  for (int i = 0; i < desc.num_environments; i++) {
    const ygl_environment& env = desc.environments[i];
    if (env.emission[0] == 0 && env.emission[1] == 0 && env.emission[2] == 0) continue;
    HostLight light;
    light.environment = i;
    if (env.emission_tex >= 0) {
      const ygl_texture& tex = desc.textures[env.emission_tex];
      light.cdf.resize((size_t)tex.width * tex.height);
      for (size_t idx = 0; idx < light.cdf.size(); idx++) {
        int   ix = (int)idx % tex.width, iy = (int)idx / tex.width;
        float th = (iy + 0.5f) * kPi / tex.height;
        float v[4];
        if (tex.pixelsf) {
          for (int c = 0; c < 4; c++) v[c] = tex.pixelsf[4 * ((size_t)iy * tex.width + ix) + c];
        } else {
          for (int c = 0; c < 4; c++) v[c] = tex.pixelsb[4 * ((size_t)iy * tex.width + ix) + c] / 255.0f;
        }
        float m        = ymax(ymax(ymax(v[0], v[1]), v[2]), v[3]);  // max(vec4f), yocto_math.h:1519
        light.cdf[idx] = m * std::sin(th);
        if (idx != 0) light.cdf[idx] += light.cdf[idx - 1];
      }
    }
    lights.push_back(std::move(light));
  }
}

bool state_size(const ygl_scene_desc& desc, const ygl_trace_params& params, int& width, int& height,
    std::string& error) {
  if (params.camera < 0 || params.camera >= desc.num_cameras) return error = "camera id out of range", false;
  if (params.resolution <= 0) return error = "resolution must be positive", false;
  const ygl_camera& camera = desc.cameras[params.camera];
  if (camera.aspect >= 1) {
    width  = params.resolution;
    height = (int)std::round(params.resolution / camera.aspect);
  } else {
    height = params.resolution;
    width  = (int)std::round(params.resolution * camera.aspect);
  }
  return true;
}

void state_rngs(const ygl_trace_params& params, int width, int height, uint64_t* rngs) {
  rng_t seq = rng_make(1301081, 1);
  for (size_t i = 0; i < (size_t)width * height; i++) {
    rng_t rng       = rng_make(params.seed, (uint64_t)(rand1i(seq, (int)(1u << 31)) / 2 + 1));
    rngs[2 * i + 0] = rng.state;
    rngs[2 * i + 1] = rng.inc;
  }
}

}  // namespace ygl
