// ygl_traverse.cuh — two-level BVH closest-hit / any-hit traversal and ray-primitive tests.
// Behavioural contract: libs/yocto/yocto_bvh.cpp:460-628 (intersect_shape_bvh,
// intersect_scene_bvh, intersect_instance_bvh) and libs/yocto/yocto_geometry.h:697-864
// (intersect_point/line/triangle/quad/bbox). Node visit order, tie-breaking ("last tested
// t <= tmax wins") and every rounding step are the reference's, so hit ids are bit-exact.
#pragma once

#include "ygl_scene.cuh"

namespace ygl {

constexpr int kStackSize = 64;  // per level; host build rejects deeper trees (reference: 128, unchecked)

struct hit_t {
  int   instance, element;
  f2    uv;
  float distance;
  bool  hit;
};

struct trav_counters {  // per-thread traversal statistics (SURVEY.md §8d algorithmic bytes)
  unsigned top_nodes, bot_nodes, instances, prims;
};

// intersect_bbox(ray, ray_dinv, bbox), yocto_geometry.h:854-864
YGL_D bool slab_test(const f3& o, const f3& dinv, float tmin, float tmax, const float4& n0, const float4& n1) {
  auto bmin   = f3{n0.x, n0.y, n0.z};
  auto bmax   = f3{n0.w, n1.x, n1.y};
  auto it_min = (bmin - o) * dinv;
  auto it_max = (bmax - o) * dinv;
  auto lo     = vmin(it_min, it_max);
  auto hi     = vmax(it_min, it_max);
  auto t0     = ymax(max3(lo), tmin);
  auto t1     = ymin(min3(hi), tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}

// intersect_triangle, yocto_geometry.h:794-825 (edges precomputed: e1 = p1 - p0, e2 = p2 - p0)
YGL_D bool hit_triangle(const f3& o, const f3& d, float tmin, float tmax, const f3& p0, const f3& edge1,
    const f3& edge2, f2& uv, float& dist) {
  auto pvec = cross(d, edge2);
  auto det  = dot(edge1, pvec);
  if (det == 0) return false;
  auto inv_det = 1.0f / det;
  auto tvec    = o - p0;
  auto u       = dot(tvec, pvec) * inv_det;
  if (u < 0 || u > 1) return false;
  auto qvec = cross(tvec, edge1);
  auto v    = dot(d, qvec) * inv_det;
  if (v < 0 || u + v > 1) return false;
  auto t = dot(edge2, qvec) * inv_det;
  if (t < tmin || t > tmax) return false;
  uv   = {u, v};
  dist = t;
  return true;
}

// intersect_quad, yocto_geometry.h:828-835
YGL_D bool hit_quad(const f3& o, const f3& d, float tmin, float tmax, const f3& p0, const f3& p1, const f3& p2,
    const f3& p3, f2& uv, float& dist) {
  if (p2 == p3) return hit_triangle(o, d, tmin, tmax, p0, p1 - p0, p3 - p0, uv, dist);
  f2    uv1 = {0, 0}, uv2 = {0, 0};
  float d1 = kFltMax, d2 = kFltMax;  // prim_intersection{}.distance == flt_max
  bool  h1 = hit_triangle(o, d, tmin, tmax, p0, p1 - p0, p3 - p0, uv1, d1);
  bool  h2 = hit_triangle(o, d, tmin, tmax, p2, p3 - p2, p1 - p2, uv2, d2);
  if (h2) uv2 = f2{1 - uv2.x, 1 - uv2.y};
  if (d1 < d2) {
    uv   = uv1;
    dist = d1;
    return h1;
  }
  uv   = uv2;
  dist = d2;
  return h2;
}

// intersect_line, yocto_geometry.h:716-757
YGL_D bool hit_line(const f3& o, const f3& dir, float tmin, float tmax, const f3& p0, const f3& p1, float r0,
    float r1, f2& uv, float& dist) {
  auto u   = dir;
  auto v   = p1 - p0;
  auto w   = o - p0;
  auto a   = dot(u, u);
  auto b   = dot(u, v);
  auto c   = dot(v, v);
  auto d   = dot(u, w);
  auto e   = dot(v, w);
  auto det = a * c - b * b;
  if (det == 0) return false;
  auto t = (b * e - c * d) / det;
  auto s = (a * e - b * d) / det;
  if (t < tmin || t > tmax) return false;
  s        = yclamp(s, (float)0, (float)1);
  auto pr  = o + dir * t;
  auto pl  = p0 + (p1 - p0) * s;
  auto prl = pr - pl;
  auto d2  = dot(prl, prl);
  auto r   = r0 * (1 - s) + r1 * s;
  if (d2 > r * r) return false;
  uv   = {s, ysqrt(d2) / r};
  dist = t;
  return true;
}

// intersect_point, yocto_geometry.h:697-713
YGL_D bool hit_point(const f3& o, const f3& d, float tmin, float tmax, const f3& p, float r, f2& uv, float& dist) {
  auto w = p - o;
  auto t = dot(w, d) / dot(d, d);
  if (t < tmin || t > tmax) return false;
  auto rp  = o + d * t;
  auto prp = p - rp;
  if (dot(prp, prp) > r * r) return false;
  uv   = {0, 0};
  dist = t;
  return true;
}

// intersect_shape_bvh, yocto_bvh.cpp:460-552. `tmax` is the running ray.tmax (shrinks on hits).
template <bool COUNT>
YGL_D bool traverse_shape(const DShape& shape, const f3& o, const f3& d, float tmin, float tmax, bool find_any,
    int* stack, int& element, f2& uv, float& distance, trav_counters& cnt) {
  if (shape.num_nodes == 0) return false;
  const float4* __restrict__ nodes   = shape.nodes;
  const float4* __restrict__ packets = shape.packets;
  const int kind = shape.bvh_kind;

  auto dinv = f3{1 / d.x, 1 / d.y, 1 / d.z};
  // ray_dsign as a bit mask (bit a set <=> dinv[a] < 0)
  unsigned sgn = ((dinv.x < 0) ? 1u : 0u) | ((dinv.y < 0) ? 2u : 0u) | ((dinv.z < 0) ? 4u : 0u);

  bool hit = false;
  int  sp  = 0;
  int  cur = 0;  // node to visit next (popped)
  while (true) {
    float4 n0 = __ldg(nodes + 2 * cur), n1 = __ldg(nodes + 2 * cur + 1);
    if (COUNT) cnt.bot_nodes++;
    if (slab_test(o, dinv, tmin, tmax, n0, n1)) {
      int      start = __float_as_int(n1.z);
      unsigned meta  = (unsigned)__float_as_int(n1.w);
      if (meta >> 24) {
        // internal: reference pushes (start, start+1) or (start+1, start) by ray_dsign[axis] and pops the
        // last pushed; visiting that child directly and stacking the other is the same order.
        int neg     = (sgn >> ((meta >> 16) & 0xff)) & 1;
        stack[sp++] = start + 1 - neg;  // popped second
        cur         = start + neg;      // popped first
        continue;
      }
      int num = meta & 0xffff;
      if (kind == kElemTriangles) {
        for (int idx = start; idx < start + num; idx++) {
          float4 a = __ldg(packets + 3 * idx), b = __ldg(packets + 3 * idx + 1), c = __ldg(packets + 3 * idx + 2);
          if (COUNT) cnt.prims++;
          f2    puv;
          float pd;
          if (!hit_triangle(o, d, tmin, tmax, f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y}, f3{b.z, b.w, c.x}, puv, pd))
            continue;
          hit = true, element = __ldg(shape.prims + idx), uv = puv, distance = pd, tmax = pd;
        }
      } else if (kind == kElemQuads) {
        for (int idx = start; idx < start + num; idx++) {
          float4 a = __ldg(packets + 4 * idx), b = __ldg(packets + 4 * idx + 1), c = __ldg(packets + 4 * idx + 2),
                 e = __ldg(packets + 4 * idx + 3);
          if (COUNT) cnt.prims++;
          f2    puv;
          float pd;
          if (!hit_quad(o, d, tmin, tmax, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, f3{c.x, c.y, c.z},
                  f3{e.x, e.y, e.z}, puv, pd))
            continue;
          hit = true, element = __ldg(shape.prims + idx), uv = puv, distance = pd, tmax = pd;
        }
      } else if (kind == kElemLines) {
        for (int idx = start; idx < start + num; idx++) {
          float4 a = __ldg(packets + 2 * idx), b = __ldg(packets + 2 * idx + 1);
          if (COUNT) cnt.prims++;
          f2    puv;
          float pd;
          if (!hit_line(o, d, tmin, tmax, f3{a.x, a.y, a.z}, f3{b.x, b.y, b.z}, a.w, b.w, puv, pd)) continue;
          hit = true, element = __ldg(shape.prims + idx), uv = puv, distance = pd, tmax = pd;
        }
      } else if (kind == kElemPoints) {
        for (int idx = start; idx < start + num; idx++) {
          float4 a = __ldg(packets + idx);
          if (COUNT) cnt.prims++;
          f2    puv;
          float pd;
          if (!hit_point(o, d, tmin, tmax, f3{a.x, a.y, a.z}, a.w, puv, pd)) continue;
          hit = true, element = __ldg(shape.prims + idx), uv = puv, distance = pd, tmax = pd;
        }
      }
      if (find_any && hit) return true;
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return hit;
}

// one instance visit: transform_ray(inverse(frame, true), ray) + intersect_shape_bvh,
// yocto_bvh.cpp:601-604 / :621-624
template <bool COUNT>
YGL_D bool traverse_instance(const DScene& scene, const DInstancePacket* pk, const f3& o, const f3& d, float tmin,
    float tmax, bool find_any, int* stack, int& instance, int& element, f2& uv, float& distance,
    trav_counters& cnt) {
  float4 a = __ldg(&pk->a), b = __ldg(&pk->b), c = __ldg(&pk->c), e = __ldg(&pk->d);
  frame3 inv = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
  auto   lo  = transform_point(inv, o);
  auto   ld  = transform_vector(inv, d);
  if (COUNT) cnt.instances++;
  const DShape& shape = scene.shapes[__float_as_int(e.x)];
  if (!traverse_shape<COUNT>(shape, lo, ld, tmin, tmax, find_any, stack, element, uv, distance, cnt)) return false;
  instance = __float_as_int(e.y);
  return true;
}

// intersect_scene_bvh, yocto_bvh.cpp:554-617
template <bool COUNT>
YGL_D hit_t traverse_scene(const DScene& scene, const f3& o, const f3& d, float tmin, float tmax, bool find_any,
    trav_counters& cnt) {
  hit_t res = {-1, -1, {0, 0}, 0, false};
  if (scene.top_num_nodes == 0) return res;
  int stack[2 * kStackSize];
  const float4* __restrict__ nodes = scene.top_nodes;

  auto dinv   = f3{1 / d.x, 1 / d.y, 1 / d.z};
  unsigned sgn = ((dinv.x < 0) ? 1u : 0u) | ((dinv.y < 0) ? 2u : 0u) | ((dinv.z < 0) ? 4u : 0u);
  int      sp = 0, cur = 0;
  while (true) {
    float4 n0 = __ldg(nodes + 2 * cur), n1 = __ldg(nodes + 2 * cur + 1);
    if (COUNT) cnt.top_nodes++;
    if (slab_test(o, dinv, tmin, tmax, n0, n1)) {
      int      start = __float_as_int(n1.z);
      unsigned meta  = (unsigned)__float_as_int(n1.w);
      if (meta >> 24) {
        int neg     = (sgn >> ((meta >> 16) & 0xff)) & 1;
        stack[sp++] = start + 1 - neg;
        cur         = start + neg;
        continue;
      }
      int num = meta & 0xffff;
      for (int idx = start; idx < start + num; idx++) {
        int   inst, elem;
        f2    uv;
        float dist;
        if (!traverse_instance<COUNT>(scene, scene.top_packets + idx, o, d, tmin, tmax, find_any, stack + kStackSize,
                inst, elem, uv, dist, cnt))
          continue;
        res  = {inst, elem, uv, dist, true};
        tmax = dist;
      }
      if (find_any && res.hit) return res;
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return res;
}

// intersect_instance_bvh, yocto_bvh.cpp:619-628
template <bool COUNT>
YGL_D hit_t traverse_single_instance(const DScene& scene, int instance, const f3& o, const f3& d, float tmin,
    float tmax, bool find_any, trav_counters& cnt) {
  int   stack[kStackSize];
  hit_t res = {-1, -1, {0, 0}, 0, false};
  int   inst, elem;
  f2    uv;
  float dist;
  if (traverse_instance<COUNT>(scene, scene.inst_packets + instance, o, d, tmin, tmax, find_any, stack, inst, elem,
          uv, dist, cnt))
    res = {instance, elem, uv, dist, true};
  return res;
}

}  // namespace ygl
