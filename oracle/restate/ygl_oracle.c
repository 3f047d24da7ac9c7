/*
 * oracle/restate/ygl_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's algorithm for the hot path, written from the
 * behavioural contract, one function per reference function, each citing the
 * reference file:line it follows (paths under libs/yocto/). It uses the host's float libm exactly
 * like the reference does, so — compiled with -ffp-contract=off — it reproduces the reference
 * bit for bit; tests/test_oracle_restatement.py pins it against the real reference
 * (oracle/_ref/libyocto_ref.so) and against the committed golden fixtures.
 *
 * Coverage: PCG32 streams and the per-pixel seeding of make_trace_state; make_bvh (split_middle and
 * split_sah) for the two-level BVH; intersect_point/line/triangle/quad/bbox; intersect_shape_bvh /
 * intersect_scene_bvh / intersect_instance_bvh; sample_camera/eval_camera; eval_position / normals / texcoords /
 * colors / normal maps / material; textures (bilinear, sRGB decode); environments; every lobe of yocto_shading.h the
 * renderer uses; make_trace_lights, sample_lights, sample_lights_pdf for area lights and (textured) environments;
 * trace_path, trace_pathdirect, trace_pathmis, trace_pathtest, trace_naive, trace_eyelight, trace_diagram,
 * trace_furnace, trace_falsecolor, trace_sample, trace_samples, trace_image, incl. opacity pass-through, nocaustics,
 * the tent filter, transmission lobes, participating media, textures, vertex colors, normal maps and textured
 * environments: the whole hot path of SURVEY.md 8a.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ygl_b200.h"

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
typedef struct { v3 x, y, z, o; } fr3;
typedef struct { v3 x, y, z; } m3;

static const float pif = (float)3.14159265358979323846;

/* ---- yocto_math.h:1045-1372: scalar and vector helpers, operation order preserved ---- */
static float absf_(float a) { return a < 0 ? -a : a; }
static float minf_(float a, float b) { return (a < b) ? a : b; }
static float maxf_(float a, float b) { return (a > b) ? a : b; }
static float clampf_(float a, float lo, float hi) { return minf_(maxf_(a, lo), hi); }
static int   clampi_(int a, int lo, int hi) { int m = a > lo ? a : lo; return m < hi ? m : hi; }
static v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static v3 add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 muls(v3 a, float b) { return V3(a.x * b, a.y * b, a.z * b); }
static v3 smul(float a, v3 b) { return V3(a * b.x, a * b.y, a * b.z); }
static v3 divv(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static v3 divs(v3 a, float b) { return V3(a.x / b, a.y / b, a.z / b); }
static v3 adds(v3 a, float b) { return V3(a.x + b, a.y + b, a.z + b); }
static v3 subs(v3 a, float b) { return V3(a.x - b, a.y - b, a.z - b); }
static v3 ssub(float a, v3 b) { return V3(a - b.x, a - b.y, a - b.z); }
static v3 sadd(float a, v3 b) { return V3(a + b.x, a + b.y, a + b.z); }
static int eq3(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static int zero3(v3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float length(v3 a) { return sqrtf(dot(a, a)); }
static v3 normalize(v3 a) { float l = length(a); return (l != 0) ? divs(a, l) : a; }
static float max3(v3 a) { return maxf_(maxf_(a.x, a.y), a.z); }
static float min3(v3 a) { return minf_(minf_(a.x, a.y), a.z); }
static v3 vmin(v3 a, v3 b) { return V3(minf_(a.x, b.x), minf_(a.y, b.y), minf_(a.z, b.z)); }
static v3 vmax(v3 a, v3 b) { return V3(maxf_(a.x, b.x), maxf_(a.y, b.y), maxf_(a.z, b.z)); }
static v3 vclamp(v3 a, float lo, float hi) { return V3(clampf_(a.x, lo, hi), clampf_(a.y, lo, hi), clampf_(a.z, lo, hi)); }
static v3 vsqrt(v3 a) { return V3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
static int finite3(v3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
static v3 lerp3(v3 a, v3 b, float u) { return add(muls(a, 1 - u), muls(b, u)); }
static float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static v3 reflect(v3 w, v3 n) { return add(neg(w), smul(2 * dot(n, w), n)); } /* yocto_math.h:1336 */
static fr3 to_frame(const ygl_frame3f* f) {
  fr3 r = {{f->x[0], f->x[1], f->x[2]}, {f->y[0], f->y[1], f->y[2]}, {f->z[0], f->z[1], f->z[2]}, {f->o[0], f->o[1], f->o[2]}};
  return r;
}
/* transform_point / vector / direction, yocto_math.h:2263-2271 */
static v3 xf_vector(const fr3* a, v3 b) { return add(add(muls(a->x, b.x), muls(a->y, b.y)), muls(a->z, b.z)); }
static v3 xf_point(const fr3* a, v3 b) { return add(xf_vector(a, b), a->o); }
static v3 xf_direction(const fr3* a, v3 b) { return normalize(xf_vector(a, b)); }
static v3 m3_mul(const m3* a, v3 b) { return add(add(muls(a->x, b.x), muls(a->y, b.y)), muls(a->z, b.z)); }
/* inverse(frame, non_rigid = true), yocto_math.h:2114-2122 + :1965-1972 */
static fr3 frame_inverse(const fr3* a) {
  v3    cyz = cross(a->y, a->z), czx = cross(a->z, a->x), cxy = cross(a->x, a->y);
  float det = dot(a->x, cross(a->y, a->z)), idt = 1 / det;
  m3    minv = {muls(V3(cyz.x, czx.x, cxy.x), idt), muls(V3(cyz.y, czx.y, cxy.y), idt), muls(V3(cyz.z, czx.z, cxy.z), idt)};
  fr3   r = {minv.x, minv.y, minv.z, neg(m3_mul(&minv, a->o))};
  return r;
}
/* inverse(frame, non_rigid = false): transpose of the rotation, yocto_math.h:2114-2122 */
static fr3 frame_inverse_rigid(const fr3* a) {
  m3  minv = {V3(a->x.x, a->y.x, a->z.x), V3(a->x.y, a->y.y, a->z.y), V3(a->x.z, a->y.z, a->z.z)};
  fr3 r = {minv.x, minv.y, minv.z, neg(m3_mul(&minv, a->o))};
  return r;
}
/* basis_fromz, yocto_math.h:1977-1986 */
static m3 basis_fromz(v3 v) {
  v3    z = normalize(v);
  float sign = copysignf(1.0f, z.z), a = -1.0f / (sign + z.z), b = z.x * z.y * a;
  m3    r = {V3(1.0f + sign * z.x * z.x * a, sign * b, -sign * z.x), V3(b, sign + z.y * z.y * a, -z.y), z};
  return r;
}

/* ---- PCG32, yocto_sampling.h:187-232 ---- */
typedef struct { uint64_t state, inc; } rng_t;
static uint32_t rng_next(rng_t* r) {
  uint64_t old = r->state;
  r->state     = old * 6364136223846793005ULL + r->inc;
  uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}
static rng_t make_rng(uint64_t seed, uint64_t seq) {
  rng_t r = {0, (seq << 1u) | 1u};
  rng_next(&r);
  r.state += seed;
  rng_next(&r);
  return r;
}
static float rand1f(rng_t* r) {
  uint32_t u = (rng_next(r) >> 9) | 0x3f800000u;
  float    f;
  memcpy(&f, &u, 4);
  return f - 1.0f;
}
static v2 rand2f(rng_t* r) { v2 v; v.x = rand1f(r); v.y = rand1f(r); return v; }

/* ---- scene access ---- */
typedef struct { ygl_bvh_node* nodes; int32_t* prims; int num_nodes, num_prims; } tree_t;
typedef struct { int instance, environment; float* cdf; int n; } light_t;
typedef struct oracle_scene {
  const ygl_scene_desc* d;
  tree_t  top, *shapes;
  light_t* lights;
  int      num_lights;
} oracle_scene;

static v3 P(const ygl_shape* s, int i) { return V3(s->positions[3 * i], s->positions[3 * i + 1], s->positions[3 * i + 2]); }
static v3 N(const ygl_shape* s, int i) { return V3(s->normals[3 * i], s->normals[3 * i + 1], s->normals[3 * i + 2]); }

/* ---- make_bvh, yocto_bvh.cpp:108-302 ---- */
typedef struct { v3 min, max; } box3;
static box3 box_invalid(void) { box3 b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; return b; }
static box3 box_merge_p(box3 a, v3 p) { box3 b = {vmin(a.min, p), vmax(a.max, p)}; return b; }
static box3 box_merge(box3 a, box3 c) { box3 b = {vmin(a.min, c.min), vmax(a.max, c.max)}; return b; }
static v3   box_center(box3 a) { return divs(add(a.min, a.max), 2); }
/* libstdc++ std::partition for bidirectional iterators */
static int partition_lt(int32_t* prims, const v3* centers, int start, int end, int axis, float split) {
  int first = start, last = end;
  while (1) {
    while (1) {
      if (first == last) return first;
      else if (comp(centers[prims[first]], axis) < split) ++first;
      else break;
    }
    --last;
    while (1) {
      if (first == last) return first;
      else if (!(comp(centers[prims[last]], axis) < split)) --last;
      else break;
    }
    int32_t t = prims[first]; prims[first] = prims[last]; prims[last] = t;
    ++first;
  }
}
static float box_area(box3 b) { /* bbox_area lambda, yocto_bvh.cpp:125-129 */
  v3 s = sub(b.max, b.min);
  return 1e-12f + 2 * s.x * s.y + 2 * s.x * s.z + 2 * s.y * s.z;
}
static void split_node(int32_t* prims, const box3* bboxes, const v3* centers, int start, int end, int hq, int* mid, int* axis_out) {
  box3 cb = box_invalid();
  for (int i = start; i < end; i++) cb = box_merge_p(cb, centers[prims[i]]);
  v3 cs = sub(cb.max, cb.min);
  if (zero3(cs)) { *mid = (start + end) / 2; *axis_out = 0; return; }
  int axis = 0; float split = 0.0f;
  if (!hq) { /* split_middle, yocto_bvh.cpp:202-232 */
    if (cs.x >= cs.y && cs.x >= cs.z) axis = 0;
    if (cs.y >= cs.x && cs.y >= cs.z) axis = 1;
    if (cs.z >= cs.x && cs.z >= cs.y) axis = 2;
    split = comp(box_center(cb), axis);
  } else { /* split_sah, yocto_bvh.cpp:108-164 */
    const int nbins = 16; float min_cost = FLT_MAX;
    for (int sa = 0; sa < 3; sa++) for (int b = 1; b < nbins; b++) {
      float bsplit = comp(cb.min, sa) + b * comp(cs, sa) / nbins;
      box3  lb = box_invalid(), rb = box_invalid(); int ln = 0, rn = 0;
      for (int i = start; i < end; i++) {
        if (comp(centers[prims[i]], sa) < bsplit) { lb = box_merge(lb, bboxes[prims[i]]); ln += 1; }
        else { rb = box_merge(rb, bboxes[prims[i]]); rn += 1; }
      }
      float cost = 1 + ln * box_area(lb) / box_area(cb) + rn * box_area(rb) / box_area(cb);
      if (cost < min_cost) { min_cost = cost; split = bsplit; axis = sa; }
    }
  }
  int m = partition_lt(prims, centers, start, end, axis, split);
  if (m == start || m == end) m = (start + end) / 2;
  *mid = m; *axis_out = axis;
}
static tree_t make_tree(const box3* bboxes, int n, int hq) {
  tree_t t; t.num_prims = n; t.prims = malloc(sizeof(int32_t) * (n > 0 ? n : 1));
  t.nodes = calloc((size_t)2 * n + 1, sizeof(ygl_bvh_node)); t.num_nodes = 1;
  v3* centers = malloc(sizeof(v3) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { t.prims[i] = i; centers[i] = box_center(bboxes[i]); }
  int (*stack)[3] = malloc(sizeof(int[3]) * ((size_t)2 * n + 2)); int sp = 0;
  stack[sp][0] = 0; stack[sp][1] = 0; stack[sp][2] = n; sp++;
  while (sp) {
    sp--; int id = stack[sp][0], start = stack[sp][1], end = stack[sp][2];
    box3 bb = box_invalid();
    for (int i = start; i < end; i++) bb = box_merge(bb, bboxes[t.prims[i]]);
    ygl_bvh_node node; memset(&node, 0, sizeof(node));
    node.bbox_min[0] = bb.min.x; node.bbox_min[1] = bb.min.y; node.bbox_min[2] = bb.min.z;
    node.bbox_max[0] = bb.max.x; node.bbox_max[1] = bb.max.y; node.bbox_max[2] = bb.max.z;
    if (end - start > 4) {
      int mid, axis; split_node(t.prims, bboxes, centers, start, end, hq, &mid, &axis);
      node.internal = 1; node.axis = (int8_t)axis; node.num = 2; node.start = t.num_nodes; t.num_nodes += 2;
      stack[sp][0] = node.start; stack[sp][1] = start; stack[sp][2] = mid; sp++;
      stack[sp][0] = node.start + 1; stack[sp][1] = mid; stack[sp][2] = end; sp++;
    } else { node.internal = 0; node.num = (int16_t)(end - start); node.start = start; }
    t.nodes[id] = node;
  }
  free(centers); free(stack);
  return t;
}
static tree_t make_shape_tree(const ygl_shape* s, int hq) { /* make_shape_bvh, yocto_bvh.cpp:321-362 */
  int n = s->num_points ? s->num_points : s->num_lines ? s->num_lines : s->num_triangles ? s->num_triangles : s->num_quads;
  box3* b = malloc(sizeof(box3) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    if (s->num_points) { int p = s->points[i]; float r = s->radius[p]; v3 q = P(s, p);
      b[i].min = vmin(subs(q, r), adds(q, r)); b[i].max = vmax(subs(q, r), adds(q, r)); }
    else if (s->num_lines) { int a = s->lines[2 * i], c = s->lines[2 * i + 1]; float r0 = s->radius[a], r1 = s->radius[c];
      b[i].min = vmin(subs(P(s, a), r0), subs(P(s, c), r1)); b[i].max = vmax(adds(P(s, a), r0), adds(P(s, c), r1)); }
    else if (s->num_triangles) { const int32_t* t = s->triangles + 3 * i;
      b[i].min = vmin(P(s, t[0]), vmin(P(s, t[1]), P(s, t[2]))); b[i].max = vmax(P(s, t[0]), vmax(P(s, t[1]), P(s, t[2]))); }
    else { const int32_t* q = s->quads + 4 * i;
      b[i].min = vmin(P(s, q[0]), vmin(P(s, q[1]), vmin(P(s, q[2]), P(s, q[3]))));
      b[i].max = vmax(P(s, q[0]), vmax(P(s, q[1]), vmax(P(s, q[2]), P(s, q[3])))); }
  }
  tree_t t = make_tree(b, n, hq); free(b); return t;
}

/* ---- ray-primitive tests, yocto_geometry.h:697-864 ---- */
typedef struct { v3 o, d; float tmin, tmax; } ray_t;
typedef struct { v2 uv; float distance; int hit; } prim_hit;
static prim_hit miss(void) { prim_hit h = {{0, 0}, FLT_MAX, 0}; return h; }
static prim_hit hit_triangle(const ray_t* r, v3 p0, v3 p1, v3 p2) { /* :794-825 */
  v3 e1 = sub(p1, p0), e2 = sub(p2, p0), pvec = cross(r->d, e2); float det = dot(e1, pvec);
  if (det == 0) return miss();
  float inv_det = 1.0f / det; v3 tvec = sub(r->o, p0);
  float u = dot(tvec, pvec) * inv_det; if (u < 0 || u > 1) return miss();
  v3 qvec = cross(tvec, e1); float v = dot(r->d, qvec) * inv_det; if (v < 0 || u + v > 1) return miss();
  float t = dot(e2, qvec) * inv_det; if (t < r->tmin || t > r->tmax) return miss();
  prim_hit h = {{u, v}, t, 1}; return h;
}
static prim_hit hit_quad(const ray_t* r, v3 p0, v3 p1, v3 p2, v3 p3) { /* :828-835 */
  if (eq3(p2, p3)) return hit_triangle(r, p0, p1, p3);
  prim_hit a = hit_triangle(r, p0, p1, p3), b = hit_triangle(r, p2, p3, p1);
  if (b.hit) { b.uv.x = 1 - b.uv.x; b.uv.y = 1 - b.uv.y; }
  return a.distance < b.distance ? a : b;
}
static prim_hit hit_line(const ray_t* r, v3 p0, v3 p1, float r0, float r1) { /* :716-757 */
  v3 u = r->d, v = sub(p1, p0), w = sub(r->o, p0);
  float a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w), det = a * c - b * b;
  if (det == 0) return miss();
  float t = (b * e - c * d) / det, s = (a * e - b * d) / det;
  if (t < r->tmin || t > r->tmax) return miss();
  s = clampf_(s, (float)0, (float)1);
  v3 pr = add(r->o, muls(r->d, t)), pl = add(p0, muls(sub(p1, p0), s)), prl = sub(pr, pl);
  float d2 = dot(prl, prl), rr = r0 * (1 - s) + r1 * s;
  if (d2 > rr * rr) return miss();
  prim_hit h = {{s, sqrtf(d2) / rr}, t, 1}; return h;
}
static prim_hit hit_point(const ray_t* r, v3 p, float rad) { /* :697-713 */
  v3 w = sub(p, r->o); float t = dot(w, r->d) / dot(r->d, r->d);
  if (t < r->tmin || t > r->tmax) return miss();
  v3 rp = add(r->o, muls(r->d, t)), prp = sub(p, rp);
  if (dot(prp, prp) > rad * rad) return miss();
  prim_hit h = {{0, 0}, t, 1}; return h;
}
static int hit_bbox(const ray_t* r, v3 dinv, const ygl_bvh_node* n) { /* :854-864 */
  v3 bmin = V3(n->bbox_min[0], n->bbox_min[1], n->bbox_min[2]), bmax = V3(n->bbox_max[0], n->bbox_max[1], n->bbox_max[2]);
  v3 a = mul(sub(bmin, r->o), dinv), b = mul(sub(bmax, r->o), dinv), lo = vmin(a, b), hi = vmax(a, b);
  float t0 = maxf_(max3(lo), r->tmin), t1 = minf_(min3(hi), r->tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}

/* ---- intersect_shape_bvh, yocto_bvh.cpp:460-552 ---- */
typedef struct { int element; v2 uv; float distance; int hit; } shape_hit;
static shape_hit intersect_shape(const tree_t* t, const ygl_shape* s, ray_t ray, int find_any) {
  shape_hit res = {-1, {0, 0}, 0, 0};
  if (t->num_nodes == 0) return res;
  int stack[128], sp = 0; stack[sp++] = 0;
  v3  dinv = V3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
  int dsign[3] = {dinv.x < 0, dinv.y < 0, dinv.z < 0};
  while (sp) {
    const ygl_bvh_node* n = &t->nodes[stack[--sp]];
    if (!hit_bbox(&ray, dinv, n)) continue;
    if (n->internal) {
      if (dsign[n->axis]) { stack[sp++] = n->start; stack[sp++] = n->start + 1; }
      else { stack[sp++] = n->start + 1; stack[sp++] = n->start; }
    } else {
      for (int idx = n->start; idx < n->start + n->num; idx++) {
        int e = t->prims[idx]; prim_hit h;
        if (s->num_points) { int p = s->points[e]; h = hit_point(&ray, P(s, p), s->radius[p]); }
        else if (s->num_lines) { int a = s->lines[2 * e], b = s->lines[2 * e + 1]; h = hit_line(&ray, P(s, a), P(s, b), s->radius[a], s->radius[b]); }
        else if (s->num_triangles) { const int32_t* q = s->triangles + 3 * e; h = hit_triangle(&ray, P(s, q[0]), P(s, q[1]), P(s, q[2])); }
        else { const int32_t* q = s->quads + 4 * e; h = hit_quad(&ray, P(s, q[0]), P(s, q[1]), P(s, q[2]), P(s, q[3])); }
        if (!h.hit) continue;
        res.element = e; res.uv = h.uv; res.distance = h.distance; res.hit = 1;
        ray.tmax = h.distance;
      }
    }
    if (find_any && res.hit) return res;
  }
  return res;
}
static ygl_intersection no_hit(void) { ygl_intersection h = {-1, -1, {0, 0}, 0, 0}; return h; }
static ray_t transform_ray(const fr3* f, const ray_t* r) { ray_t o = {xf_point(f, r->o), xf_vector(f, r->d), r->tmin, r->tmax}; return o; }
/* intersect_instance_bvh, yocto_bvh.cpp:619-628 */
static ygl_intersection intersect_instance(const oracle_scene* sc, int instance, ray_t ray, int find_any) {
  const ygl_instance* in = &sc->d->instances[instance];
  fr3 f = to_frame(&in->frame), inv = frame_inverse(&f);
  shape_hit h = intersect_shape(&sc->shapes[in->shape], &sc->d->shapes[in->shape], transform_ray(&inv, &ray), find_any);
  if (!h.hit) return no_hit();
  ygl_intersection r = {instance, h.element, {h.uv.x, h.uv.y}, h.distance, 1};
  return r;
}
/* intersect_scene_bvh, yocto_bvh.cpp:554-617 */
static ygl_intersection intersect_scene(const oracle_scene* sc, ray_t ray, int find_any) {
  ygl_intersection res = no_hit();
  const tree_t* t = &sc->top;
  if (t->num_nodes == 0) return res;
  int stack[128], sp = 0; stack[sp++] = 0;
  v3  dinv = V3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
  int dsign[3] = {dinv.x < 0, dinv.y < 0, dinv.z < 0};
  while (sp) {
    const ygl_bvh_node* n = &t->nodes[stack[--sp]];
    if (!hit_bbox(&ray, dinv, n)) continue;
    if (n->internal) {
      if (dsign[n->axis]) { stack[sp++] = n->start; stack[sp++] = n->start + 1; }
      else { stack[sp++] = n->start + 1; stack[sp++] = n->start; }
    } else {
      for (int idx = n->start; idx < n->start + n->num; idx++) {
        int i = t->prims[idx];
        const ygl_instance* in = &sc->d->instances[i];
        fr3 f = to_frame(&in->frame), inv = frame_inverse(&f);
        shape_hit h = intersect_shape(&sc->shapes[in->shape], &sc->d->shapes[in->shape], transform_ray(&inv, &ray), find_any);
        if (!h.hit) continue;
        res.instance = i; res.element = h.element; res.uv[0] = h.uv.x; res.uv[1] = h.uv.y; res.distance = h.distance; res.hit = 1;
        ray.tmax = h.distance;
      }
    }
    if (find_any && res.hit) return res;
  }
  return res;
}

/* ---- scene build: make_scene_bvh (yocto_bvh.cpp:364-396), make_trace_lights (yocto_trace.cpp:1528-1581) ---- */
static float tri_area(v3 p0, v3 p1, v3 p2) { return length(cross(sub(p1, p0), sub(p2, p0))) / 2; }
oracle_scene* oracle_scene_create(const ygl_scene_desc* d, int highquality) {
  oracle_scene* sc = calloc(1, sizeof(*sc));
  sc->d = d;
  sc->shapes = calloc(d->num_shapes > 0 ? d->num_shapes : 1, sizeof(tree_t));
  for (int i = 0; i < d->num_shapes; i++) sc->shapes[i] = make_shape_tree(&d->shapes[i], highquality);
  box3* ib = malloc(sizeof(box3) * (d->num_instances > 0 ? d->num_instances : 1));
  for (int i = 0; i < d->num_instances; i++) {
    const ygl_instance* in = &d->instances[i]; const ygl_bvh_node* n = &sc->shapes[in->shape].nodes[0];
    fr3 f = to_frame(&in->frame);
    v3 mn = V3(n->bbox_min[0], n->bbox_min[1], n->bbox_min[2]), mx = V3(n->bbox_max[0], n->bbox_max[1], n->bbox_max[2]);
    v3 c[8] = {{mn.x, mn.y, mn.z}, {mn.x, mn.y, mx.z}, {mn.x, mx.y, mn.z}, {mn.x, mx.y, mx.z},
               {mx.x, mn.y, mn.z}, {mx.x, mn.y, mx.z}, {mx.x, mx.y, mn.z}, {mx.x, mx.y, mx.z}};
    box3 b = box_invalid();
    for (int k = 0; k < 8; k++) b = box_merge_p(b, xf_point(&f, c[k]));
    ib[i] = b;
  }
  sc->top = make_tree(ib, d->num_instances, highquality);
  free(ib);
  sc->lights = calloc((size_t)d->num_instances + d->num_environments + 1, sizeof(light_t));
  for (int i = 0; i < d->num_instances; i++) {
    const ygl_instance* in = &d->instances[i]; const ygl_material* m = &d->materials[in->material];
    if (m->emission[0] == 0 && m->emission[1] == 0 && m->emission[2] == 0) continue;
    const ygl_shape* s = &d->shapes[in->shape];
    if (!s->num_triangles && !s->num_quads) continue;
    light_t* l = &sc->lights[sc->num_lights++]; l->instance = i; l->environment = -1;
    if (s->num_triangles) {
      l->n = s->num_triangles; l->cdf = malloc(sizeof(float) * l->n);
      for (int e = 0; e < l->n; e++) { const int32_t* t = s->triangles + 3 * e;
        l->cdf[e] = tri_area(P(s, t[0]), P(s, t[1]), P(s, t[2])); if (e) l->cdf[e] += l->cdf[e - 1]; }
    }
    if (s->num_quads) {
      free(l->cdf); l->n = s->num_quads; l->cdf = malloc(sizeof(float) * l->n);
      for (int e = 0; e < l->n; e++) { const int32_t* q = s->quads + 4 * e;
        l->cdf[e] = tri_area(P(s, q[0]), P(s, q[1]), P(s, q[3])) + tri_area(P(s, q[2]), P(s, q[3]), P(s, q[1]));
        if (e) l->cdf[e] += l->cdf[e - 1]; }
    }
  }
  for (int i = 0; i < d->num_environments; i++) {
    const ygl_environment* e = &d->environments[i];
    if (e->emission[0] == 0 && e->emission[1] == 0 && e->emission[2] == 0) continue;
    light_t* l = &sc->lights[sc->num_lights++]; l->instance = -1; l->environment = i; l->cdf = NULL; l->n = 0;
    if (e->emission_tex >= 0) { /* make_trace_lights, yocto_trace.cpp:1562-1576: max(texel) * sin(theta), cumulative */
      const ygl_texture* t = &d->textures[e->emission_tex];
      l->n = t->width * t->height; l->cdf = malloc(sizeof(float) * (size_t)(l->n > 0 ? l->n : 1));
      for (int idx = 0; idx < l->n; idx++) {
        int ix = idx % t->width, iy = idx / t->width;
        float th = (iy + 0.5f) * pif / t->height;
        float v[4];
        size_t k = 4 * ((size_t)iy * t->width + ix);
        for (int c = 0; c < 4; c++) v[c] = t->pixelsf ? t->pixelsf[k + c] : t->pixelsb[k + c] / 255.0f;
        float mx = maxf_(maxf_(maxf_(v[0], v[1]), v[2]), v[3]);
        l->cdf[idx] = mx * sinf(th);
        if (idx != 0) l->cdf[idx] += l->cdf[idx - 1];
      }
    }
  }
  return sc;
}
void oracle_scene_destroy(oracle_scene* sc) {
  if (!sc) return;
  for (int i = 0; i < sc->d->num_shapes; i++) { free(sc->shapes[i].nodes); free(sc->shapes[i].prims); }
  free(sc->shapes); free(sc->top.nodes); free(sc->top.prims);
  for (int i = 0; i < sc->num_lights; i++) free(sc->lights[i].cdf);
  free(sc->lights); free(sc);
}
/* 1 if every feature the scene uses is restated here (all material types, textures, vertex colors, normal maps,
 * textured environments: everything ygl_scene_desc can describe) */
int oracle_supported(const ygl_scene_desc* d) {
  for (int i = 0; i < d->num_materials; i++)
    if (d->materials[i].type < YGL_MATERIAL_MATTE || d->materials[i].type > YGL_MATERIAL_GLTFPBR) return 0;
  return 1;
}
void oracle_tree_size(const oracle_scene* sc, int shape, int* nn, int* np) {
  const tree_t* t = shape < 0 ? &sc->top : &sc->shapes[shape]; *nn = t->num_nodes; *np = t->num_prims;
}
void oracle_tree_get(const oracle_scene* sc, int shape, ygl_bvh_node* nodes, int32_t* prims) {
  const tree_t* t = shape < 0 ? &sc->top : &sc->shapes[shape];
  memcpy(nodes, t->nodes, sizeof(ygl_bvh_node) * t->num_nodes); memcpy(prims, t->prims, sizeof(int32_t) * t->num_prims);
}
void oracle_intersect_rays(const oracle_scene* sc, const ygl_ray* rays, int64_t n, int instance, int find_any, ygl_intersection* out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; i++) {
    ray_t r = {{rays[i].o[0], rays[i].o[1], rays[i].o[2]}, {rays[i].d[0], rays[i].d[1], rays[i].d[2]}, rays[i].tmin, rays[i].tmax};
    out[i] = instance < 0 ? intersect_scene(sc, r, find_any) : intersect_instance(sc, instance, r, find_any);
  }
}

/* ---- eval_*, yocto_scene.cpp:288-613 (untextured subset) ---- */
static v3 interp_tri(v3 p0, v3 p1, v3 p2, v2 uv) { return add(add(muls(p0, 1 - uv.x - uv.y), muls(p1, uv.x)), muls(p2, uv.y)); }
static v3 interp_quad(v3 p0, v3 p1, v3 p2, v3 p3, v2 uv) {
  if (uv.x + uv.y <= 1) return interp_tri(p0, p1, p3, uv);
  v2 w = {1 - uv.x, 1 - uv.y}; return interp_tri(p2, p3, p1, w);
}
static v3 eval_position(const oracle_scene* sc, int instance, int e, v2 uv) { /* :288-312 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape]; fr3 f = to_frame(&in->frame);
  if (s->num_triangles) { const int32_t* t = s->triangles + 3 * e; return xf_point(&f, interp_tri(P(s, t[0]), P(s, t[1]), P(s, t[2]), uv)); }
  if (s->num_quads) { const int32_t* q = s->quads + 4 * e; return xf_point(&f, interp_quad(P(s, q[0]), P(s, q[1]), P(s, q[2]), P(s, q[3]), uv)); }
  if (s->num_lines) { const int32_t* l = s->lines + 2 * e; return xf_point(&f, add(muls(P(s, l[0]), 1 - uv.x), muls(P(s, l[1]), uv.x))); }
  if (s->num_points) return xf_point(&f, P(s, s->points[e]));
  return V3(0, 0, 0);
}
static v3 tri_normal(v3 p0, v3 p1, v3 p2) { return normalize(cross(sub(p1, p0), sub(p2, p0))); }
static v3 eval_element_normal(const oracle_scene* sc, int instance, int e) { /* :315-337 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape]; fr3 f = to_frame(&in->frame);
  if (s->num_triangles) { const int32_t* t = s->triangles + 3 * e; return normalize(xf_vector(&f, tri_normal(P(s, t[0]), P(s, t[1]), P(s, t[2])))); }
  if (s->num_quads) { const int32_t* q = s->quads + 4 * e;
    return normalize(xf_vector(&f, normalize(add(tri_normal(P(s, q[0]), P(s, q[1]), P(s, q[3])), tri_normal(P(s, q[2]), P(s, q[3]), P(s, q[1])))))); }
  if (s->num_lines) { const int32_t* l = s->lines + 2 * e; return normalize(xf_vector(&f, normalize(sub(P(s, l[1]), P(s, l[0]))))); }
  if (s->num_points) return V3(0, 0, 1);
  return V3(0, 0, 0);
}
static v3 eval_normal(const oracle_scene* sc, int instance, int e, v2 uv) { /* :340-366 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape]; fr3 f = to_frame(&in->frame);
  if (!s->num_normals) return eval_element_normal(sc, instance, e);
  if (s->num_triangles) { const int32_t* t = s->triangles + 3 * e; return normalize(xf_vector(&f, normalize(interp_tri(N(s, t[0]), N(s, t[1]), N(s, t[2]), uv)))); }
  if (s->num_quads) { const int32_t* q = s->quads + 4 * e; return normalize(xf_vector(&f, normalize(interp_quad(N(s, q[0]), N(s, q[1]), N(s, q[2]), N(s, q[3]), uv)))); }
  if (s->num_lines) { const int32_t* l = s->lines + 2 * e; return normalize(xf_vector(&f, normalize(add(muls(N(s, l[0]), 1 - uv.x), muls(N(s, l[1]), uv.x))))); }
  if (s->num_points) return normalize(xf_vector(&f, normalize(N(s, s->points[e]))));
  return V3(0, 0, 0);
}
/* ---- textures, yocto_scene.cpp:111-171, yocto_color.h:235-238 ---- */
typedef struct { float x, y, z, w; } v4;
static v4 V4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static v4 add4(v4 a, v4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static v4 muls4(v4 a, float b) { return V4(a.x * b, a.y * b, a.z * b, a.w * b); }
static float srgb_to_rgb1(float srgb) { /* yocto_color.h:235-238: the threshold is a double literal */
  return ((double)srgb <= 0.04045) ? srgb / 12.92f : powf((srgb + 0.055f) / (1.0f + 0.055f), 2.4f);
}
static v4 lookup_texture(const ygl_texture* t, int i, int j, int as_linear) {
  v4 c;
  size_t k = 4 * ((size_t)j * t->width + i);
  if (t->pixelsf) c = V4(t->pixelsf[k], t->pixelsf[k + 1], t->pixelsf[k + 2], t->pixelsf[k + 3]);
  else c = V4(t->pixelsb[k] / 255.0f, t->pixelsb[k + 1] / 255.0f, t->pixelsb[k + 2] / 255.0f, t->pixelsb[k + 3] / 255.0f);
  if (as_linear && !t->linear) return V4(srgb_to_rgb1(c.x), srgb_to_rgb1(c.y), srgb_to_rgb1(c.z), c.w);
  return c;
}
static v4 eval_texture_raw(const ygl_texture* t, v2 uv, int as_linear, int no_interpolation, int clamp_to_edge) { /* :127-160 */
  if (t->width == 0 || t->height == 0) return V4(0, 0, 0, 0);
  int sx = t->width, sy = t->height;
  float s_, t_;
  if (clamp_to_edge) { s_ = clampf_(uv.x, 0.0f, 1.0f) * sx; t_ = clampf_(uv.y, 0.0f, 1.0f) * sy; }
  else { s_ = fmodf(uv.x, 1.0f) * sx; if (s_ < 0) s_ += sx; t_ = fmodf(uv.y, 1.0f) * sy; if (t_ < 0) t_ += sy; }
  int i = clampi_((int)s_, 0, sx - 1), j = clampi_((int)t_, 0, sy - 1);
  int ii = (i + 1) % sx, jj = (j + 1) % sy;
  float u = s_ - i, v = t_ - j;
  if (no_interpolation) return lookup_texture(t, i, j, as_linear);
  return add4(add4(add4(muls4(muls4(lookup_texture(t, i, j, as_linear), 1 - u), 1 - v),
                        muls4(muls4(lookup_texture(t, i, jj, as_linear), 1 - u), v)),
                   muls4(muls4(lookup_texture(t, ii, j, as_linear), u), 1 - v)),
              muls4(muls4(lookup_texture(t, ii, jj, as_linear), u), v));
}
static v4 eval_texture(const oracle_scene* sc, int texture, v2 uv, int as_linear) { /* :167-171 */
  if (texture < 0) return V4(1, 1, 1, 1);
  const ygl_texture* t = &sc->d->textures[texture];
  return eval_texture_raw(t, uv, as_linear, t->nearest != 0, t->clamp != 0);
}
static v2 eval_texcoord(const oracle_scene* sc, int instance, int e, v2 uv) { /* yocto_scene.cpp:369-391 */
  const ygl_shape* s = &sc->d->shapes[sc->d->instances[instance].shape];
  if (!s->num_texcoords) return uv;
  const float* tc = s->texcoords;
#define TC(i) V3(tc[2 * (i)], tc[2 * (i) + 1], 0)
  v3 r = V3(0, 0, 0);
  if (s->num_triangles) { const int32_t* t = s->triangles + 3 * e; r = interp_tri(TC(t[0]), TC(t[1]), TC(t[2]), uv); }
  else if (s->num_quads) { const int32_t* q = s->quads + 4 * e; r = interp_quad(TC(q[0]), TC(q[1]), TC(q[2]), TC(q[3]), uv); }
  else if (s->num_lines) { const int32_t* l = s->lines + 2 * e; r = add(muls(TC(l[0]), 1 - uv.x), muls(TC(l[1]), uv.x)); }
  else if (s->num_points) r = TC(s->points[e]);
#undef TC
  v2 o = {r.x, r.y};
  return o;
}
static v4 eval_color(const oracle_scene* sc, int instance, int e, v2 uv) { /* yocto_scene.cpp:394-416 */
  const ygl_shape* s = &sc->d->shapes[sc->d->instances[instance].shape];
  if (!s->num_colors) return V4(1, 1, 1, 1);
  const float* c = s->colors;
#define CL(i) V4(c[4 * (i)], c[4 * (i) + 1], c[4 * (i) + 2], c[4 * (i) + 3])
#define TRI4(a, b, cc, uv_) add4(add4(muls4(a, 1 - (uv_).x - (uv_).y), muls4(b, (uv_).x)), muls4(cc, (uv_).y))
  if (s->num_triangles) { const int32_t* t = s->triangles + 3 * e; return TRI4(CL(t[0]), CL(t[1]), CL(t[2]), uv); }
  if (s->num_quads) { const int32_t* q = s->quads + 4 * e;
    if (uv.x + uv.y <= 1) return TRI4(CL(q[0]), CL(q[1]), CL(q[3]), uv);
    v2 w = {1 - uv.x, 1 - uv.y}; return TRI4(CL(q[2]), CL(q[3]), CL(q[1]), w); }
  if (s->num_lines) { const int32_t* l = s->lines + 2 * e; return add4(muls4(CL(l[0]), 1 - uv.x), muls4(CL(l[1]), uv.x)); }
  if (s->num_points) return CL(s->points[e]);
#undef TRI4
#undef CL
  return V4(0, 0, 0, 0);
}
/* eval_element_tangents (yocto_scene.cpp:425-446; quads use their first triangle) and triangle_tangents_fromuv
 * (yocto_geometry.h:504-522) */
static v2 TCv(const ygl_shape* s, int i) { v2 r = {s->texcoords[2 * i], s->texcoords[2 * i + 1]}; return r; }
static void triangle_tangents_fromuv(v3 p0, v3 p1, v3 p2, v2 uv0, v2 uv1, v2 uv2, v3* tu, v3* tv) {
  v3 p = sub(p1, p0), q = sub(p2, p0);
  v2 s_ = {uv1.x - uv0.x, uv2.x - uv0.x}, t_ = {uv1.y - uv0.y, uv2.y - uv0.y};
  float div = s_.x * t_.y - s_.y * t_.x;
  if (div != 0) {
    *tu = divs(V3(t_.y * p.x - t_.x * q.x, t_.y * p.y - t_.x * q.y, t_.y * p.z - t_.x * q.z), div);
    *tv = divs(V3(s_.x * q.x - s_.y * p.x, s_.x * q.y - s_.y * p.y, s_.x * q.z - s_.y * p.z), div);
  } else { *tu = V3(1, 0, 0); *tv = V3(0, 1, 0); }
}
static void eval_element_tangents(const oracle_scene* sc, int instance, int e, v3* tu, v3* tv) {
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape]; fr3 f = to_frame(&in->frame);
  v3 a, b;
  if (s->num_triangles && s->num_texcoords) { const int32_t* t = s->triangles + 3 * e;
    triangle_tangents_fromuv(P(s, t[0]), P(s, t[1]), P(s, t[2]), TCv(s, t[0]), TCv(s, t[1]), TCv(s, t[2]), &a, &b);
  } else if (s->num_quads && s->num_texcoords) { const int32_t* q = s->quads + 4 * e;
    triangle_tangents_fromuv(P(s, q[0]), P(s, q[1]), P(s, q[3]), TCv(s, q[0]), TCv(s, q[1]), TCv(s, q[3]), &a, &b);
  } else { *tu = V3(0, 0, 0); *tv = V3(0, 0, 0); return; }
  *tu = xf_direction(&f, a); *tv = xf_direction(&f, b);
}
static v3 eval_normalmap(const oracle_scene* sc, int instance, int e, v2 uv) { /* yocto_scene.cpp:448-468 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape];
  const ygl_material* m = &sc->d->materials[in->material];
  v3 normal = eval_normal(sc, instance, e, uv);
  v2 texcoord = eval_texcoord(sc, instance, e, uv);
  if (m->normal_tex >= 0 && (s->num_triangles || s->num_quads)) {
    const ygl_texture* t = &sc->d->textures[m->normal_tex];
    v4 tx = eval_texture_raw(t, texcoord, 0, t->nearest != 0, t->clamp != 0);
    v3 nmap = V3(-1 + 2 * tx.x, -1 + 2 * tx.y, -1 + 2 * tx.z);
    v3 tu, tv; eval_element_tangents(sc, instance, e, &tu, &tv);
    v3 fx = normalize(sub(tu, muls(normal, dot(tu, normal)))); /* orthonormalize(tu, normal) */
    v3 fy = normalize(cross(normal, fx));
    int flip_v = dot(fy, tv) < 0;
    nmap.y *= flip_v ? 1 : -1;
    normal = normalize(add(add(muls(fx, nmap.x), muls(fy, nmap.y)), muls(normal, nmap.z)));
  }
  return normal;
}
static v3 eval_shading_normal(const oracle_scene* sc, int instance, int e, v2 uv, v3 outgoing) { /* :486-505 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape];
  const ygl_material* m = &sc->d->materials[in->material];
  if (s->num_triangles || s->num_quads) {
    v3 n = eval_normal(sc, instance, e, uv);
    if (m->normal_tex >= 0) n = eval_normalmap(sc, instance, e, uv);
    if (m->type == YGL_MATERIAL_REFRACTIVE) return n;
    return dot(n, outgoing) >= 0 ? n : neg(n);
  } else if (s->num_lines) { v3 n = eval_normal(sc, instance, e, uv); return normalize(sub(outgoing, muls(n, dot(outgoing, n)))); }
  else if (s->num_points) return outgoing;
  return V3(0, 0, 0);
}
static v3 eval_shading_position(const oracle_scene* sc, int instance, int e, v2 uv) { /* :471-483 */
  const ygl_instance* in = &sc->d->instances[instance]; const ygl_shape* s = &sc->d->shapes[in->shape];
  if (s->num_triangles || s->num_quads || s->num_lines) return eval_position(sc, instance, e, uv);
  if (s->num_points) return P(s, s->points[e]); /* object space: reference quirk */
  return V3(0, 0, 0);
}

typedef struct { int type; v3 emission, color; float opacity, roughness, metallic, ior; v3 density, scattering; float scanisotropy, trdepth; } mpoint;
static mpoint eval_material(const oracle_scene* sc, int instance, int element, v2 uv) { /* yocto_scene.cpp:531-581 */
  const ygl_material* m = &sc->d->materials[sc->d->instances[instance].material];
  v2 texcoord = eval_texcoord(sc, instance, element, uv);
  v4 emission_tex = eval_texture(sc, m->emission_tex, texcoord, 1);
  v4 color_shp = eval_color(sc, instance, element, uv);
  v4 color_tex = eval_texture(sc, m->color_tex, texcoord, 1);
  v4 roughness_tex = eval_texture(sc, m->roughness_tex, texcoord, 0);
  v4 scattering_tex = eval_texture(sc, m->scattering_tex, texcoord, 1);
  mpoint p; p.type = m->type;
  p.emission = mul(mul(V3(m->emission[0], m->emission[1], m->emission[2]), V3(emission_tex.x, emission_tex.y, emission_tex.z)),
      V3(color_shp.x, color_shp.y, color_shp.z));
  p.color = mul(mul(V3(m->color[0], m->color[1], m->color[2]), V3(color_tex.x, color_tex.y, color_tex.z)),
      V3(color_shp.x, color_shp.y, color_shp.z));
  p.opacity = m->opacity * color_tex.w * color_shp.w; p.metallic = m->metallic * roughness_tex.z;
  p.roughness = m->roughness * roughness_tex.y; p.roughness = p.roughness * p.roughness; p.ior = m->ior;
  p.scattering = mul(V3(m->scattering[0], m->scattering[1], m->scattering[2]), V3(scattering_tex.x, scattering_tex.y, scattering_tex.z));
  p.scanisotropy = m->scanisotropy; p.trdepth = m->trdepth;
  if (p.type == YGL_MATERIAL_REFRACTIVE || p.type == YGL_MATERIAL_VOLUMETRIC || p.type == YGL_MATERIAL_SUBSURFACE) {
    v3 c = vclamp(p.color, 0.0001f, 1.0f);
    p.density = divs(neg(V3(logf(c.x), logf(c.y), logf(c.z))), p.trdepth);
  } else p.density = V3(0, 0, 0);
  const float min_roughness = 0.03f * 0.03f;
  if (p.type == YGL_MATERIAL_MATTE || p.type == YGL_MATERIAL_GLTFPBR || p.type == YGL_MATERIAL_GLOSSY) p.roughness = clampf_(p.roughness, min_roughness, 1.0f);
  else if (p.type == YGL_MATERIAL_VOLUMETRIC) p.roughness = 0;
  else if (p.roughness < min_roughness) p.roughness = 0;
  return p;
}
static v3 eval_environment(const oracle_scene* sc, v3 direction) { /* yocto_scene.cpp:596-613 */
  v3 e = V3(0, 0, 0);
  for (int i = 0; i < sc->d->num_environments; i++) {
    const ygl_environment* env = &sc->d->environments[i];
    fr3 f = to_frame(&env->frame), inv = frame_inverse_rigid(&f);
    v3 wl = xf_direction(&inv, direction);
    v2 texcoord = {atan2f(wl.z, wl.x) / (2 * pif), acosf(clampf_(wl.y, -1.0f, 1.0f)) / pif};
    if (texcoord.x < 0) texcoord.x += 1;
    v4 t = eval_texture(sc, env->emission_tex, texcoord, 0);
    e = add(e, mul(V3(env->emission[0], env->emission[1], env->emission[2]), V3(t.x, t.y, t.z)));
  }
  return e;
}

/* ---- sampling, yocto_sampling.h:252-398 ---- */
static v3 sample_hemisphere_cos(v3 normal, v2 ruv) {
  float z = sqrtf(ruv.y), r = sqrtf(1 - z * z), phi = 2 * pif * ruv.x;
  m3 b = basis_fromz(normal);
  return normalize(m3_mul(&b, V3(r * cosf(phi), r * sinf(phi), z)));
}
static float sample_hemisphere_cos_pdf(v3 normal, v3 direction) { float c = dot(normal, direction); return (c <= 0) ? 0 : c / pif; }
static v3 sample_sphere(v2 ruv) { float z = 2 * ruv.y - 1, r = sqrtf(clampf_(1 - z * z, 0.0f, 1.0f)), phi = 2 * pif * ruv.x; return V3(r * cosf(phi), r * sinf(phi), z); }
static v2 sample_disk(v2 ruv) { float r = sqrtf(ruv.y), phi = 2 * pif * ruv.x; v2 o = {cosf(phi) * r, sinf(phi) * r}; return o; }
static int sample_discrete(const float* cdf, int n, float r) {
  float last = cdf[n - 1];
  r = clampf_(r * last, (float)0, last - (float)0.00001);
  int lo = 0, cnt = n; /* std::upper_bound */
  while (cnt > 0) { int step = cnt / 2, mid = lo + step; if (!(r < cdf[mid])) { lo = mid + 1; cnt -= step + 1; } else cnt = step; }
  return clampi_(lo, 0, n - 1);
}

/* ---- shading, yocto_shading.h:303-788 (matte, glossy, reflective, gltfpbr) ---- */
static v3 up_of(v3 n, v3 o) { return dot(n, o) <= 0 ? neg(n) : n; }
static float fresnel_dielectric(float eta, v3 n, v3 o) {
  float cosw = absf_(dot(n, o)), sin2 = 1 - cosw * cosw, eta2 = eta * eta, cos2t = 1 - sin2 / eta2;
  if (cos2t < 0) return 1;
  float t0 = sqrtf(cos2t), t1 = eta * t0, t2 = eta * cosw, rs = (cosw - t1) / (cosw + t1), rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2;
}
static v3 fresnel_conductor(v3 eta, v3 etak, v3 n, v3 o) {
  float cosw = dot(n, o); if (cosw <= 0) return V3(0, 0, 0);
  cosw = clampf_(cosw, (float)-1, (float)1);
  float cos2 = cosw * cosw, sin2 = clampf_(1 - cos2, (float)0, (float)1);
  v3 eta2 = mul(eta, eta), etak2 = mul(etak, etak), t0 = subs(sub(eta2, etak2), sin2);
  v3 a2plusb2 = vsqrt(add(mul(t0, t0), mul(smul(4, eta2), etak2))), t1 = adds(a2plusb2, cos2);
  v3 a = vsqrt(divs(add(a2plusb2, t0), 2)), t2 = muls(smul(2, a), cosw), rs = divv(sub(t1, t2), add(t1, t2));
  v3 t3 = adds(smul(cos2, a2plusb2), sin2 * sin2), t4 = muls(t2, sin2), rp = divv(mul(rs, sub(t3, t4)), add(t3, t4));
  return divs(add(rp, rs), 2);
}
static v3 fresnel_schlick(v3 spec, v3 n, v3 o) {
  if (zero3(spec)) return V3(0, 0, 0);
  float c = dot(n, o);
  return add(spec, muls(ssub(1, spec), powf(clampf_(1 - absf_(c), 0.0f, 1.0f), 5.0f)));
}
static v3 reflectivity_to_eta(v3 r_) { v3 r = vclamp(r_, 0.0f, 0.99f); return divv(sadd(1, vsqrt(r)), ssub(1, vsqrt(r))); }
static v3 eta_to_reflectivity(v3 eta) { return divv(mul(subs(eta, 1), subs(eta, 1)), mul(adds(eta, 1), adds(eta, 1))); }
static float mf_distribution(float rough, v3 n, v3 h) {
  float c = dot(n, h); if (c <= 0) return 0;
  float r2 = rough * rough, c2 = c * c;
  return r2 / (pif * (c2 * r2 + 1 - c2) * (c2 * r2 + 1 - c2));
}
static float mf_shadowing1(float rough, v3 n, v3 h, v3 d) {
  float c = dot(n, d), ch = dot(h, d); if (c * ch <= 0) return 0;
  float r2 = rough * rough, c2 = c * c;
  return 2 * absf_(c) / (absf_(c) + sqrtf(c2 - r2 * c2 + r2));
}
static float mf_shadowing(float rough, v3 n, v3 h, v3 o, v3 i) { return mf_shadowing1(rough, n, h, o) * mf_shadowing1(rough, n, h, i); }
static v3 sample_microfacet(float rough, v3 n, v2 rn) {
  float phi = 2 * pif * rn.x, theta = atanf(rough * sqrtf(rn.y / (1 - rn.y)));
  m3 b = basis_fromz(n);
  return normalize(m3_mul(&b, V3(cosf(phi) * sinf(theta), sinf(phi) * sinf(theta), cosf(theta))));
}
static float sample_microfacet_pdf(float rough, v3 n, v3 h) { float c = dot(n, h); if (c < 0) return 0; return mf_distribution(rough, n, h) * c; }
static int same_hemisphere(v3 n, v3 o, v3 i) { return dot(n, o) * dot(n, i) >= 0; }


/* ---- transmission lobes, yocto_shading.h:791-1048, and refract (yocto_math.h:1339) ---- */
static v3 refract_(v3 w, v3 n, float inv_eta) {
  float cosine = dot(n, w), k = 1 + inv_eta * inv_eta * (cosine * cosine - 1);
  if (k < 0) return V3(0, 0, 0);
  return add(muls(neg(w), inv_eta), muls(n, inv_eta * cosine - sqrtf(k)));
}
static v3 ones_times(float a) { return V3(1 * a, 1 * a, 1 * a); }
static v3 eval_transparent(v3 color, float ior, float rough, v3 n, v3 o, v3 i) { /* :791-812 */
  v3 up = up_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    v3 h = normalize(add(i, o));
    float F = fresnel_dielectric(ior, h, o), D = mf_distribution(rough, up, h), G = mf_shadowing(rough, up, h, o, i);
    return muls(divs(muls(muls(ones_times(F), D), G), 4 * dot(up, o) * dot(up, i)), absf_(dot(up, i)));
  } else {
    v3 refl = reflect(neg(i), up), h = normalize(add(refl, o));
    float F = fresnel_dielectric(ior, h, o), D = mf_distribution(rough, up, h), G = mf_shadowing(rough, up, h, o, refl);
    return muls(divs(muls(muls(muls(color, 1 - F), D), G), 4 * dot(up, o) * dot(up, refl)), absf_(dot(up, refl)));
  }
}
static v3 sample_transparent(float ior, float rough, v3 n, v3 o, float rnl, v2 rn) { /* :815-832 */
  v3 up = up_of(n, o), h = sample_microfacet(rough, up, rn);
  if (rnl < fresnel_dielectric(ior, h, o)) {
    v3 inc = reflect(o, h);
    if (!same_hemisphere(up, o, inc)) return V3(0, 0, 0);
    return inc;
  } else {
    v3 refl = reflect(o, h), inc = neg(reflect(refl, up));
    if (same_hemisphere(up, o, inc)) return V3(0, 0, 0);
    return inc;
  }
}
static float sample_transparent_pdf(float ior, float rough, v3 n, v3 o, v3 i) { /* :835-849 */
  v3 up = up_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) {
    v3 h = normalize(add(i, o));
    return fresnel_dielectric(ior, h, o) * sample_microfacet_pdf(rough, up, h) / (4 * absf_(dot(o, h)));
  } else {
    v3 refl = reflect(neg(i), up), h = normalize(add(refl, o));
    float d = (1 - fresnel_dielectric(ior, h, o)) * sample_microfacet_pdf(rough, up, h);
    return d / (4 * absf_(dot(o, h)));
  }
}
static v3 eval_transparent_delta(v3 color, float ior, v3 n, v3 o, v3 i) { /* :852-862 */
  v3 up = up_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) return ones_times(fresnel_dielectric(ior, up, o));
  return muls(color, 1 - fresnel_dielectric(ior, up, o));
}
static v3 sample_transparent_delta(float ior, v3 n, v3 o, float rnl) { /* :865-872 */
  v3 up = up_of(n, o);
  if (rnl < fresnel_dielectric(ior, up, o)) return reflect(o, up);
  return neg(o);
}
static float sample_transparent_delta_pdf(float ior, v3 n, v3 o, v3 i) { /* :875-881 */
  v3 up = up_of(n, o);
  if (dot(n, i) * dot(n, o) >= 0) return fresnel_dielectric(ior, up, o);
  return 1 - fresnel_dielectric(ior, up, o);
}
/* g++ -O2 expands pow(x, 2.0f) to x * x (no libm call): yocto_shading.h:907,952 */
static float sqr_(float a) { return a * a; }
static v3 eval_refractive(float ior, float rough, v3 n, v3 o, v3 i) { /* :884-911 */
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n); float rel = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    v3 h = normalize(add(i, o));
    float F = fresnel_dielectric(rel, h, o), D = mf_distribution(rough, up, h), G = mf_shadowing(rough, up, h, o, i);
    return muls(divs(muls(muls(ones_times(F), D), G), absf_(4 * dot(n, o) * dot(n, i))), absf_(dot(n, i)));
  } else {
    v3 h = muls(neg(normalize(add(muls(i, rel), o))), entering ? 1.0f : -1.0f);
    float F = fresnel_dielectric(rel, h, o), D = mf_distribution(rough, up, h), G = mf_shadowing(rough, up, h, o, i);
    v3 a = ones_times(absf_((dot(o, h) * dot(i, h)) / (dot(o, n) * dot(i, n))));
    return muls(divs(muls(muls(muls(a, 1 - F), D), G), sqr_(rel * dot(h, i) + dot(h, o))), absf_(dot(n, i)));
  }
}
static v3 sample_refractive(float ior, float rough, v3 n, v3 o, float rnl, v2 rn) { /* :914-932 */
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n);
  v3 h = sample_microfacet(rough, up, rn);
  if (rnl < fresnel_dielectric(entering ? ior : (1 / ior), h, o)) {
    v3 inc = reflect(o, h);
    if (!same_hemisphere(up, o, inc)) return V3(0, 0, 0);
    return inc;
  } else {
    v3 inc = refract_(o, h, entering ? (1 / ior) : ior);
    if (same_hemisphere(up, o, inc)) return V3(0, 0, 0);
    return inc;
  }
}
static float sample_refractive_pdf(float ior, float rough, v3 n, v3 o, v3 i) { /* :935-957 */
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n); float rel = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) {
    v3 h = normalize(add(i, o));
    return fresnel_dielectric(rel, h, o) * sample_microfacet_pdf(rough, up, h) / (4 * absf_(dot(o, h)));
  } else {
    v3 h = muls(neg(normalize(add(muls(i, rel), o))), entering ? 1.0f : -1.0f);
    return (1 - fresnel_dielectric(rel, h, o)) * sample_microfacet_pdf(rough, up, h) * absf_(dot(h, i)) / sqr_(rel * dot(h, i) + dot(h, o));
  }
}
static int ior_is_one(float ior) { return (double)absf_(ior - 1) < 1e-3; } /* the literal is a double, :961 */
static v3 eval_refractive_delta(float ior, v3 n, v3 o, v3 i) { /* :960-975 */
  if (ior_is_one(ior)) return dot(n, i) * dot(n, o) <= 0 ? V3(1, 1, 1) : V3(0, 0, 0);
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n); float rel = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) return ones_times(fresnel_dielectric(rel, up, o));
  return muls(ones_times(1 / (rel * rel)), 1 - fresnel_dielectric(rel, up, o));
}
static v3 sample_refractive_delta(float ior, v3 n, v3 o, float rnl) { /* :978-989 */
  if (ior_is_one(ior)) return neg(o);
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n); float rel = entering ? ior : (1 / ior);
  if (rnl < fresnel_dielectric(rel, up, o)) return reflect(o, up);
  return refract_(o, up, 1 / rel);
}
static float sample_refractive_delta_pdf(float ior, v3 n, v3 o, v3 i) { /* :992-1005 */
  if (ior_is_one(ior)) return dot(n, i) * dot(n, o) < 0 ? 1.0f : 0.0f;
  int entering = dot(n, o) >= 0; v3 up = entering ? n : neg(n); float rel = entering ? ior : (1 / ior);
  if (dot(n, i) * dot(n, o) >= 0) return fresnel_dielectric(rel, up, o);
  return 1 - fresnel_dielectric(rel, up, o);
}
/* ---- volumes, yocto_shading.h:1056-1111 ---- */
static v3 vexp_(v3 a) { return V3(expf(a.x), expf(a.y), expf(a.z)); }
static v3 eval_transmittance(v3 density, float distance) { return vexp_(muls(neg(density), distance)); }
static float sample_transmittance(v3 density, float max_distance, float rl, float rd) {
  int channel = clampi_((int)(rl * 3), 0, 2);
  float dc = comp(density, channel);
  float distance = (dc == 0) ? FLT_MAX : -logf(1 - rd) / dc;
  return minf_(distance, max_distance);
}
static float sample_transmittance_pdf(v3 density, float distance, float max_distance) {
  if (distance < max_distance) { v3 t = mul(density, vexp_(muls(neg(density), distance))); return (t.x + t.y + t.z) / 3; }
  v3 t = vexp_(muls(neg(density), max_distance)); return (t.x + t.y + t.z) / 3;
}
static float eval_phasefunction(float g, v3 o, v3 i) {
  float cosine = -dot(o, i), denom = 1 + g * g - 2 * g * cosine;
  return (1 - g * g) / (4 * pif * denom * sqrtf(denom));
}
static v3 sample_phasefunction(float g, v3 o, v2 rn) {
  float cos_theta;
  if (absf_(g) < 1e-3f) cos_theta = 1 - 2 * rn.y;
  else { float square = (1 - g * g) / (1 + g - 2 * g * rn.y); cos_theta = (1 + g * g - square * square) / (2 * g); }
  float sin_theta = sqrtf(maxf_(0.0f, 1 - cos_theta * cos_theta)), phi = 2 * pif * rn.x;
  m3 b = basis_fromz(neg(o));
  return m3_mul(&b, V3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta));
}
typedef struct { v3 density, scattering; float scanisotropy; } vsdf_t;
static v3 eval_scattering(const vsdf_t* v, v3 o, v3 i) { if (zero3(v->density)) return V3(0, 0, 0); return muls(mul(v->scattering, v->density), eval_phasefunction(v->scanisotropy, o, i)); }
static v3 sample_scattering(const vsdf_t* v, v3 o, v2 rn) { if (zero3(v->density)) return V3(0, 0, 0); return sample_phasefunction(v->scanisotropy, o, rn); }
static float sample_scattering_pdf(const vsdf_t* v, v3 o, v3 i) { if (zero3(v->density)) return 0; return eval_phasefunction(v->scanisotropy, o, i); }
static int is_volumetric_type(int t) { return t == YGL_MATERIAL_REFRACTIVE || t == YGL_MATERIAL_VOLUMETRIC || t == YGL_MATERIAL_SUBSURFACE; } /* yocto_scene.cpp:257-261 */

static v3 eval_bsdfcos(const mpoint* m, v3 n, v3 o, v3 i) { /* yocto_trace.cpp:173-198 */
  if (m->roughness == 0) return V3(0, 0, 0);
  if (m->type == YGL_MATERIAL_TRANSPARENT) return eval_transparent(m->color, m->ior, m->roughness, n, o, i);
  if (m->type == YGL_MATERIAL_REFRACTIVE || m->type == YGL_MATERIAL_SUBSURFACE) return eval_refractive(m->ior, m->roughness, n, o, i);
  if (m->type == YGL_MATERIAL_VOLUMETRIC) return V3(0, 0, 0);
  if (dot(n, i) * dot(n, o) <= 0) return V3(0, 0, 0);
  v3 up = up_of(n, o);
  if (m->type == YGL_MATERIAL_MATTE) return muls(divs(m->color, pif), absf_(dot(n, i)));
  v3 h = normalize(add(i, o));
  float D = mf_distribution(m->roughness, up, h), G = mf_shadowing(m->roughness, up, h, o, i);
  if (m->type == YGL_MATERIAL_GLOSSY) {
    float F1 = fresnel_dielectric(m->ior, up, o), F = fresnel_dielectric(m->ior, h, i);
    return add(muls(divs(muls(m->color, 1 - F1), pif), absf_(dot(up, i))),
        muls(divs(muls(muls(muls(V3(1, 1, 1), F), D), G), 4 * dot(up, o) * dot(up, i)), absf_(dot(up, i))));
  }
  if (m->type == YGL_MATERIAL_REFLECTIVE) {
    v3 F = fresnel_conductor(reflectivity_to_eta(m->color), V3(0, 0, 0), h, i);
    return muls(divs(muls(muls(F, D), G), 4 * dot(up, o) * dot(up, i)), absf_(dot(up, i)));
  }
  if (m->type == YGL_MATERIAL_GLTFPBR) {
    v3 refl = lerp3(eta_to_reflectivity(V3(m->ior, m->ior, m->ior)), m->color, m->metallic);
    v3 F1 = fresnel_schlick(refl, up, o), F = fresnel_schlick(refl, h, i);
    return add(muls(divs(mul(muls(m->color, 1 - m->metallic), ssub(1, F1)), pif), absf_(dot(up, i))),
        muls(divs(muls(muls(F, D), G), 4 * dot(up, o) * dot(up, i)), absf_(dot(up, i))));
  }
  return V3(0, 0, 0);
}
static v3 sample_bsdfcos(const mpoint* m, v3 n, v3 o, float rnl, v2 rn) { /* yocto_trace.cpp:221-246 */
  if (m->roughness == 0) return V3(0, 0, 0);
  if (m->type == YGL_MATERIAL_TRANSPARENT) return sample_transparent(m->ior, m->roughness, n, o, rnl, rn);
  if (m->type == YGL_MATERIAL_REFRACTIVE || m->type == YGL_MATERIAL_SUBSURFACE) return sample_refractive(m->ior, m->roughness, n, o, rnl, rn);
  v3 up = up_of(n, o);
  if (m->type == YGL_MATERIAL_MATTE) return sample_hemisphere_cos(up, rn);
  int specular;
  if (m->type == YGL_MATERIAL_GLOSSY) specular = rnl < fresnel_dielectric(m->ior, up, o);
  else if (m->type == YGL_MATERIAL_GLTFPBR) {
    v3 refl = lerp3(eta_to_reflectivity(V3(m->ior, m->ior, m->ior)), m->color, m->metallic);
    v3 F = fresnel_schlick(refl, up, o); specular = rnl < (F.x + F.y + F.z) / 3;
  } else if (m->type == YGL_MATERIAL_REFLECTIVE) specular = 1;
  else return V3(0, 0, 0);
  if (specular) {
    v3 h = sample_microfacet(m->roughness, up, rn), inc = reflect(o, h);
    if (!same_hemisphere(up, o, inc)) return V3(0, 0, 0);
    return inc;
  }
  return sample_hemisphere_cos(up, rn);
}
static float sample_bsdfcos_pdf(const mpoint* m, v3 n, v3 o, v3 i) { /* yocto_trace.cpp:266-291 */
  if (m->roughness == 0) return 0;
  if (m->type == YGL_MATERIAL_TRANSPARENT) return sample_transparent_pdf(m->ior, m->roughness, n, o, i);
  if (m->type == YGL_MATERIAL_REFRACTIVE || m->type == YGL_MATERIAL_SUBSURFACE) return sample_refractive_pdf(m->ior, m->roughness, n, o, i);
  if (m->type == YGL_MATERIAL_VOLUMETRIC) return 0;
  if (dot(n, i) * dot(n, o) <= 0) return 0;
  v3 up = up_of(n, o);
  if (m->type == YGL_MATERIAL_MATTE) return sample_hemisphere_cos_pdf(up, i);
  v3 h = normalize(add(o, i));
  if (m->type == YGL_MATERIAL_REFLECTIVE) return sample_microfacet_pdf(m->roughness, up, h) / (4 * absf_(dot(o, h)));
  float F;
  if (m->type == YGL_MATERIAL_GLOSSY) F = fresnel_dielectric(m->ior, up, o);
  else { v3 refl = lerp3(eta_to_reflectivity(V3(m->ior, m->ior, m->ior)), m->color, m->metallic); v3 f = fresnel_schlick(refl, up, o); F = (f.x + f.y + f.z) / 3; }
  return F * sample_microfacet_pdf(m->roughness, up, h) / (4 * absf_(dot(o, h))) + (1 - F) * sample_hemisphere_cos_pdf(up, i);
}
static int is_delta(const mpoint* m) { /* yocto_scene.cpp:263-271 */
  return (m->type == YGL_MATERIAL_REFLECTIVE && m->roughness == 0) || (m->type == YGL_MATERIAL_REFRACTIVE && m->roughness == 0) ||
         (m->type == YGL_MATERIAL_TRANSPARENT && m->roughness == 0) || m->type == YGL_MATERIAL_VOLUMETRIC;
}
/* delta lobes: reflective yocto_shading.h:693-712, transparent :852-881, refractive :960-1005, passthrough :1028-1048 */
static v3 eval_delta(const mpoint* m, v3 n, v3 o, v3 i) { /* yocto_trace.cpp:200-219 */
  if (m->roughness != 0) return V3(0, 0, 0);
  switch (m->type) {
    case YGL_MATERIAL_REFLECTIVE:
      if (dot(n, i) * dot(n, o) <= 0) return V3(0, 0, 0);
      return fresnel_conductor(reflectivity_to_eta(m->color), V3(0, 0, 0), up_of(n, o), o);
    case YGL_MATERIAL_TRANSPARENT: return eval_transparent_delta(m->color, m->ior, n, o, i);
    case YGL_MATERIAL_REFRACTIVE: return eval_refractive_delta(m->ior, n, o, i);
    case YGL_MATERIAL_VOLUMETRIC: return (dot(n, i) * dot(n, o) >= 0) ? V3(0, 0, 0) : V3(1, 1, 1);
    default: return V3(0, 0, 0);
  }
}
static v3 sample_delta(const mpoint* m, v3 n, v3 o, float rnl) { /* yocto_trace.cpp:248-264 */
  if (m->roughness != 0) return V3(0, 0, 0);
  switch (m->type) {
    case YGL_MATERIAL_REFLECTIVE: return reflect(o, up_of(n, o));
    case YGL_MATERIAL_TRANSPARENT: return sample_transparent_delta(m->ior, n, o, rnl);
    case YGL_MATERIAL_REFRACTIVE: return sample_refractive_delta(m->ior, n, o, rnl);
    case YGL_MATERIAL_VOLUMETRIC: return neg(o);
    default: return V3(0, 0, 0);
  }
}
static float sample_delta_pdf(const mpoint* m, v3 n, v3 o, v3 i) { /* yocto_trace.cpp:293-310 */
  if (m->roughness != 0) return 0;
  switch (m->type) {
    case YGL_MATERIAL_REFLECTIVE: return (dot(n, i) * dot(n, o) <= 0) ? 0.0f : 1.0f;
    case YGL_MATERIAL_TRANSPARENT: return sample_transparent_delta_pdf(m->ior, n, o, i);
    case YGL_MATERIAL_REFRACTIVE: return sample_refractive_delta_pdf(m->ior, n, o, i);
    case YGL_MATERIAL_VOLUMETRIC: return (dot(n, i) * dot(n, o) >= 0) ? 0.0f : 1.0f;
    default: return 0;
  }
}

/* ---- lights, yocto_trace.cpp:361-443 ---- */
static v3 sample_lights(const oracle_scene* sc, v3 position, float rl, float rel, v2 ruv) {
  int id = clampi_((int)(rl * sc->num_lights), 0, sc->num_lights - 1);
  const light_t* l = &sc->lights[id];
  if (l->instance >= 0) {
    const ygl_shape* s = &sc->d->shapes[sc->d->instances[l->instance].shape];
    int e = sample_discrete(l->cdf, l->n, rel);
    v2 uv = ruv;
    if (s->num_triangles) { uv.x = 1 - sqrtf(ruv.x); uv.y = ruv.y * sqrtf(ruv.x); }
    return normalize(sub(eval_position(sc, l->instance, e, uv), position));
  }
  if (l->environment >= 0) { /* :376-386 */
    const ygl_environment* env = &sc->d->environments[l->environment];
    if (env->emission_tex >= 0) {
      const ygl_texture* t = &sc->d->textures[env->emission_tex];
      int idx = sample_discrete(l->cdf, l->n, rel);
      v2 uv = {((idx % t->width) + 0.5f) / t->width, ((idx / t->width) + 0.5f) / t->height};
      fr3 f = to_frame(&env->frame);
      return xf_direction(&f, V3(cosf(uv.x * 2 * pif) * sinf(uv.y * pif), cosf(uv.y * pif), sinf(uv.x * 2 * pif) * sinf(uv.y * pif)));
    }
    return sample_sphere(ruv);
  }
  return V3(0, 0, 0);
}
static float sample_lights_pdf(const oracle_scene* sc, v3 position, v3 direction) {
  float pdf = 0.0f;
  for (int li = 0; li < sc->num_lights; li++) {
    const light_t* l = &sc->lights[li];
    if (l->instance >= 0) {
      float lpdf = 0.0f; v3 next = position;
      for (int b = 0; b < 100; b++) {
        ray_t r = {next, direction, 1e-4f, FLT_MAX};
        ygl_intersection h = intersect_instance(sc, l->instance, r, 0);
        if (!h.hit) break;
        v2 uv = {h.uv[0], h.uv[1]};
        v3 lp = eval_position(sc, l->instance, h.element, uv), ln = eval_element_normal(sc, l->instance, h.element);
        float area = l->cdf[l->n - 1];
        v3 dd = sub(lp, position);
        lpdf += dot(dd, dd) / (absf_(dot(ln, direction)) * area);
        next = add(lp, muls(direction, 1e-3f));
      }
      pdf += lpdf;
    } else if (sc->d->environments[l->environment].emission_tex >= 0) { /* :424-437 */
      const ygl_environment* env = &sc->d->environments[l->environment];
      const ygl_texture* t = &sc->d->textures[env->emission_tex];
      fr3 f = to_frame(&env->frame), inv = frame_inverse_rigid(&f);
      v3 wl = xf_direction(&inv, direction);
      v2 texcoord = {atan2f(wl.z, wl.x) / (2 * pif), acosf(clampf_(wl.y, -1.0f, 1.0f)) / pif};
      if (texcoord.x < 0) texcoord.x += 1;
      int i = clampi_((int)(texcoord.x * t->width), 0, t->width - 1), j = clampi_((int)(texcoord.y * t->height), 0, t->height - 1);
      int idx = j * t->width + i;
      float prob = (idx == 0 ? l->cdf[0] : l->cdf[idx] - l->cdf[idx - 1]) / l->cdf[l->n - 1];
      float angle = (2 * pif / t->width) * (pif / t->height) * sinf(pif * (j + 0.5f) / t->height);
      pdf += prob / angle;
    } else pdf += 1 / (4 * pif);
  }
  pdf *= (float)1 / (float)sc->num_lights;
  return pdf;
}

/* ---- camera, yocto_scene.cpp:66-101, yocto_trace.cpp:338-358 ---- */
static ray_t eval_camera(const ygl_camera* c, v2 iuv, v2 luv) {
  fr3 f = to_frame(&c->frame);
  v2 film; if (c->aspect >= 1) { film.x = c->film; film.y = c->film / c->aspect; } else { film.x = c->film * c->aspect; film.y = c->film; }
  ray_t r; r.tmin = 1e-4f; r.tmax = FLT_MAX;
  if (!c->orthographic) {
    v3 q = V3(film.x * (0.5f - iuv.x), film.y * (iuv.y - 0.5f), c->lens), dc = neg(normalize(q));
    v3 e = V3(luv.x * c->aperture / 2, luv.y * c->aperture / 2, 0), p = divs(muls(dc, c->focus), absf_(dc.z));
    r.o = xf_point(&f, e); r.d = xf_direction(&f, normalize(sub(p, e)));
  } else {
    float scale = 1 / c->lens;
    v3 q = V3(film.x * (0.5f - iuv.x) * scale, film.y * (iuv.y - 0.5f) * scale, c->lens);
    v3 e = add(V3(-q.x, -q.y, 0), V3(luv.x * c->aperture / 2, luv.y * c->aperture / 2, 0)), p = V3(-q.x, -q.y, -c->focus);
    r.o = xf_point(&f, e); r.d = xf_direction(&f, normalize(sub(p, e)));
  }
  return r;
}


/* opacity pass-through of the sampler loops (e.g. yocto_trace.cpp:505-510): the rng is drawn only when opacity < 1 */
#define OPACITY_PASS(m, position)                                                  \
  if ((m).opacity < 1 && rand1f(rng) >= (m).opacity) {                               \
    if (opbounce++ > 128) break;                                                   \
    ray.o = add(position, muls(ray.d, 1e-2f)); ray.tmin = 1e-4f; ray.tmax = FLT_MAX; \
    bounce -= 1;                                                                   \
    continue;                                                                      \
  }
/* nocaustics roughness clamp (yocto_trace.cpp:496-500) */
#define NOCAUSTICS(m)                                                \
  if (p->nocaustics) {                                               \
    max_roughness = maxf_((m).roughness, max_roughness);             \
    (m).roughness = max_roughness;                                   \
  }

/* participating media of the path samplers (e.g. yocto_trace.cpp:476-488, 545-579): the reference's volume stack never
 * holds more than one entry, so it is a slot. Draw order of sample_transmittance's arguments (g++): rd, then rl. */
#define VOLUME_TRANSMITTANCE                                                                              \
  int in_volume = 0;                                                                                      \
  if (has_volume) {                                                                                       \
    float rd_ = rand1f(rng);                                                                              \
    float rl_ = rand1f(rng);                                                                              \
    float dist_ = sample_transmittance(vsdf.density, isec.distance, rl_, rd_);                            \
    weight = mul(weight, divs(eval_transmittance(vsdf.density, dist_),                                    \
                             sample_transmittance_pdf(vsdf.density, dist_, isec.distance)));             \
    in_volume     = dist_ < isec.distance;                                                                \
    isec.distance = dist_;                                                                                \
  }
#define VOLUME_STACK_UPDATE                                                                                         \
  if (is_volumetric_type(sc->d->materials[sc->d->instances[isec.instance].material].type) &&                        \
      dot(normal, outgoing) * dot(normal, incoming) < 0) {                                                          \
    if (!has_volume) {                                                                                              \
      mpoint vm = eval_material(sc, isec.instance, isec.element, uv);                                                                 \
      vsdf.density = vm.density; vsdf.scattering = vm.scattering; vsdf.scanisotropy = vm.scanisotropy;             \
      has_volume = 1;                                                                                               \
    } else {                                                                                                        \
      has_volume = 0;                                                                                               \
    }                                                                                                               \
  }
/* next direction of a scattering event inside the medium (:557-572): phase function or lights, one-sample MIS */
static v3 sample_volume_direction(const oracle_scene* sc, const vsdf_t* vsdf, v3 position, v3 outgoing, rng_t* rng) {
  if (rand1f(rng) < 0.5f) { v2 rn = rand2f(rng); (void)rand1f(rng); return sample_scattering(vsdf, outgoing, rn); }
  v2 ruv = rand2f(rng); float rel = rand1f(rng); float rl = rand1f(rng);
  return sample_lights(sc, position, rl, rel, ruv);
}
static v3 volume_weight(const oracle_scene* sc, const vsdf_t* vsdf, v3 position, v3 outgoing, v3 incoming) { /* :574-576 */
  return divs(eval_scattering(vsdf, outgoing, incoming),
      0.5f * sample_scattering_pdf(vsdf, outgoing, incoming) + 0.5f * sample_lights_pdf(sc, position, incoming));
}
/* ---- trace_path, yocto_trace.cpp:453-596. g++ evaluates call arguments right to left: the rand2f of
 * sample_bsdfcos / sample_lights is drawn before the rand1f's. ---- */
typedef struct { v3 radiance; int hit; v3 albedo, normal; } trace_result;
static trace_result trace_path(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) {
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, opbounce = 0, has_volume = 0; float max_roughness = 0.0f;
  vsdf_t vsdf = {{0, 0, 0}, {0, 0, 0}, 0};
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if (bounce > 0 || !p->envhidden) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    VOLUME_TRANSMITTANCE
    if (!in_volume) {
      v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
      v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
      v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
      mpoint m = eval_material(sc, isec.instance, isec.element, uv);
      NOCAUSTICS(m)
      OPACITY_PASS(m, position)
      if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
      radiance = add(radiance, mul(weight, dot(normal, outgoing) >= 0 ? m.emission : V3(0, 0, 0)));
      v3 incoming;
      if (!is_delta(&m)) {
        if (rand1f(rng) < 0.5f) { v2 rn = rand2f(rng); float rnl = rand1f(rng); incoming = sample_bsdfcos(&m, normal, outgoing, rnl, rn); }
        else { v2 ruv = rand2f(rng); float rel = rand1f(rng); float rl = rand1f(rng); incoming = sample_lights(sc, position, rl, rel, ruv); }
        if (zero3(incoming)) break;
        weight = mul(weight, divs(eval_bsdfcos(&m, normal, outgoing, incoming),
            0.5f * sample_bsdfcos_pdf(&m, normal, outgoing, incoming) + 0.5f * sample_lights_pdf(sc, position, incoming)));
      } else {
        incoming = sample_delta(&m, normal, outgoing, rand1f(rng));
        weight = mul(weight, divs(eval_delta(&m, normal, outgoing, incoming), sample_delta_pdf(&m, normal, outgoing, incoming)));
      }
      VOLUME_STACK_UPDATE
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    } else {
      v3 outgoing = neg(ray.d), position = add(ray.o, muls(ray.d, isec.distance));
      v3 incoming = sample_volume_direction(sc, &vsdf, position, outgoing, rng);
      if (zero3(incoming)) break;
      weight = mul(weight, volume_weight(sc, &vsdf, position, outgoing, incoming));
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    }
    if (zero3(weight) || !finite3(weight)) break;
    if (bounce > 3) {
      float rr = minf_((float)0.99, max3(weight));
      if (rand1f(rng) >= rr) break;
      weight = muls(weight, 1 / rr);
    }
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}


/* ---- the other samplers of get_trace_sampler_func (yocto_trace.cpp:1422-1438), restated for the scenes
 * oracle_supported() admits: no textures, no volumes (so the volume branches of the reference loops are not taken). Same right-to-left draw order as trace_path above. ---- */
static v3 eval_emission(const mpoint* m, v3 n, v3 o) { return dot(n, o) >= 0 ? m->emission : V3(0, 0, 0); } /* yocto_trace.cpp:166-170 */
static int finish_bounce(v3* weight, int bounce, rng_t* rng) { /* weight check + russian roulette, :581-591; 0 = break */
  if (zero3(*weight) || !finite3(*weight)) return 0;
  if (bounce > 3) {
    float rr = minf_((float)0.99, max3(*weight));
    if (rand1f(rng) >= rr) return 0;
    *weight = muls(*weight, 1 / rr);
  }
  return 1;
}
/* emission seen along a shadow ray, :667-677 / :874-886 */
static v3 shadow_emission(const oracle_scene* sc, v3 position, v3 incoming, ygl_intersection* out) {
  ray_t sr = {position, incoming, 1e-4f, FLT_MAX};
  ygl_intersection is = intersect_scene(sc, sr, 0);
  if (out) *out = is;
  if (!is.hit) return eval_environment(sc, incoming);
  v2 uv = {is.uv[0], is.uv[1]};
  mpoint m = eval_material(sc, is.instance, is.element, uv);
  return eval_emission(&m, eval_shading_normal(sc, is.instance, is.element, uv, neg(incoming)), neg(incoming));
}
/* one-sample MIS next direction of trace_path / pathdirect / pathtest, :522-542. Returns 0 on `break`. */
static int next_direction(const oracle_scene* sc, const mpoint* m, v3 position, v3 normal, v3 outgoing, rng_t* rng,
    v3* incoming, v3* weight, int delta_zero_check) {
  if (!is_delta(m)) {
    if (rand1f(rng) < 0.5f) { v2 rn = rand2f(rng); float rnl = rand1f(rng); *incoming = sample_bsdfcos(m, normal, outgoing, rnl, rn); }
    else { v2 ruv = rand2f(rng); float rel = rand1f(rng); float rl = rand1f(rng); *incoming = sample_lights(sc, position, rl, rel, ruv); }
    if (zero3(*incoming)) return 0;
    *weight = mul(*weight, divs(eval_bsdfcos(m, normal, outgoing, *incoming),
        0.5f * sample_bsdfcos_pdf(m, normal, outgoing, *incoming) + 0.5f * sample_lights_pdf(sc, position, *incoming)));
  } else {
    *incoming = sample_delta(m, normal, outgoing, rand1f(rng));
    if (delta_zero_check && zero3(*incoming)) return 0;
    *weight = mul(*weight, divs(eval_delta(m, normal, outgoing, *incoming), sample_delta_pdf(m, normal, outgoing, *incoming)));
  }
  return 1;
}
static trace_result trace_pathdirect(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) { /* :599-767 */
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, next_emission = 1, opbounce = 0, has_volume = 0; float max_roughness = 0.0f;
  vsdf_t vsdf = {{0, 0, 0}, {0, 0, 0}, 0};
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if ((bounce > 0 || !p->envhidden) && next_emission) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    VOLUME_TRANSMITTANCE
    if (!in_volume) {
      v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
      v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
      v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
      mpoint m = eval_material(sc, isec.instance, isec.element, uv);
      NOCAUSTICS(m)
      OPACITY_PASS(m, position)
      if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
      if (next_emission) radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
      if (!is_delta(&m)) {
        v2 ruv = rand2f(rng); float rel = rand1f(rng); float rl = rand1f(rng);
        v3 dl = sample_lights(sc, position, rl, rel, ruv);
        float pdf = sample_lights_pdf(sc, position, dl);
        v3 bsdfcos = eval_bsdfcos(&m, normal, outgoing, dl);
        if (!zero3(bsdfcos) && pdf > 0) {
          v3 emission = shadow_emission(sc, position, dl, 0);
          radiance = add(radiance, divs(mul(mul(weight, bsdfcos), emission), pdf));
        }
        next_emission = 0;
      } else {
        next_emission = 1;
      }
      v3 incoming;
      if (!next_direction(sc, &m, position, normal, outgoing, rng, &incoming, &weight, 1)) break;
      VOLUME_STACK_UPDATE
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    } else {
      v3 outgoing = neg(ray.d), position = add(ray.o, muls(ray.d, isec.distance));
      v3 incoming = sample_volume_direction(sc, &vsdf, position, outgoing, rng);
      if (zero3(incoming)) break;
      weight = mul(weight, volume_weight(sc, &vsdf, position, outgoing, incoming));
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    }
    if (!finish_bounce(&weight, bounce, rng)) break;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}
static float mis_heuristic(float this_pdf, float other_pdf) { return (this_pdf * this_pdf) / (this_pdf * this_pdf + other_pdf * other_pdf); } /* :785-788 */
static trace_result trace_pathmis(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) { /* :770-950 */
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, next_emission = 1, opbounce = 0, has_volume = 0; float max_roughness = 0.0f;
  vsdf_t vsdf = {{0, 0, 0}, {0, 0, 0}, 0};
  ygl_intersection next_intersection = no_hit();
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    ygl_intersection isec = next_emission ? intersect_scene(sc, ray, 0) : next_intersection;
    if (!isec.hit) {
      if ((bounce > 0 || !p->envhidden) && next_emission) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    VOLUME_TRANSMITTANCE
    if (!in_volume) {
      v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
      v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
      v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
      mpoint m = eval_material(sc, isec.instance, isec.element, uv);
      NOCAUSTICS(m)
      OPACITY_PASS(m, position)
      if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
      if (next_emission) radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
      v3 incoming = V3(0, 0, 0);
      if (!is_delta(&m)) {
        for (int k = 0; k < 2; k++) { /* sample_light : {true, false} */
          int sample_light = k == 0;
          if (sample_light) { v2 ruv = rand2f(rng); float rel = rand1f(rng); float rl = rand1f(rng); incoming = sample_lights(sc, position, rl, rel, ruv); }
          else { v2 rn = rand2f(rng); float rnl = rand1f(rng); incoming = sample_bsdfcos(&m, normal, outgoing, rnl, rn); }
          if (zero3(incoming)) break;
          v3 bsdfcos = eval_bsdfcos(&m, normal, outgoing, incoming);
          float light_pdf = sample_lights_pdf(sc, position, incoming);
          float bsdf_pdf = sample_bsdfcos_pdf(&m, normal, outgoing, incoming);
          float mis_weight = sample_light ? mis_heuristic(light_pdf, bsdf_pdf) / light_pdf : mis_heuristic(bsdf_pdf, light_pdf) / bsdf_pdf;
          if (!zero3(bsdfcos) && mis_weight != 0) {
            ygl_intersection sh;
            v3 emission = shadow_emission(sc, position, incoming, &sh);
            if (!sample_light) next_intersection = sh;
            radiance = add(radiance, muls(mul(mul(weight, bsdfcos), emission), mis_weight));
          }
        }
        weight = mul(weight, divs(eval_bsdfcos(&m, normal, outgoing, incoming), sample_bsdfcos_pdf(&m, normal, outgoing, incoming)));
        next_emission = 0;
      } else {
        incoming = sample_delta(&m, normal, outgoing, rand1f(rng));
        weight = mul(weight, divs(eval_delta(&m, normal, outgoing, incoming), sample_delta_pdf(&m, normal, outgoing, incoming)));
        next_emission = 1;
      }
      VOLUME_STACK_UPDATE
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    } else {
      v3 outgoing = neg(ray.d), position = add(ray.o, muls(ray.d, isec.distance));
      v3 incoming = sample_volume_direction(sc, &vsdf, position, outgoing, rng); /* no zero check here, :921-928 */
      next_emission = 1;
      weight = mul(weight, volume_weight(sc, &vsdf, position, outgoing, incoming));
      ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    }
    if (!finish_bounce(&weight, bounce, rng)) break;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}
static trace_result trace_pathtest(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) { /* :953-1029 */
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0;
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if (bounce > 0 || !p->envhidden) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
    v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
    v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
    mpoint m = eval_material(sc, isec.instance, isec.element, uv);
    m.type = YGL_MATERIAL_MATTE; /* :981, after eval_material: the roughness keeps the original type's clamp */
    if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
    radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
    v3 incoming;
    if (!next_direction(sc, &m, position, normal, outgoing, rng, &incoming, &weight, 0)) break;
    ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
    if (!finish_bounce(&weight, bounce, rng)) break;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}
static trace_result trace_naive(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) { /* :1032-1108 */
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, opbounce = 0;
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if (bounce > 0 || !p->envhidden) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
    v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
    v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
    mpoint m = eval_material(sc, isec.instance, isec.element, uv);
    OPACITY_PASS(m, position)
    if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
    radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
    v3 incoming;
    if (m.roughness != 0) {
      v2 rn = rand2f(rng); float rnl = rand1f(rng);
      incoming = sample_bsdfcos(&m, normal, outgoing, rnl, rn);
      if (zero3(incoming)) break;
      weight = mul(weight, divs(eval_bsdfcos(&m, normal, outgoing, incoming), sample_bsdfcos_pdf(&m, normal, outgoing, incoming)));
    } else {
      incoming = sample_delta(&m, normal, outgoing, rand1f(rng));
      if (zero3(incoming)) break;
      weight = mul(weight, divs(eval_delta(&m, normal, outgoing, incoming), sample_delta_pdf(&m, normal, outgoing, incoming)));
    }
    if (!finish_bounce(&weight, bounce, rng)) break;
    ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}
/* trace_eyelight (:1111-1175) and trace_diagram (:1178-1244): they differ only on a miss */
static trace_result trace_eyelight_like(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p, int diagram) {
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, opbounce = 0, nb = p->bounces > 4 ? p->bounces : 4;
  for (int bounce = 0; bounce < nb; bounce++) {
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if (diagram) { radiance = add(radiance, mul(weight, V3(1, 1, 1))); hit = 1; }
      else if (bounce > 0 || !p->envhidden) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
    v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
    v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
    mpoint m = eval_material(sc, isec.instance, isec.element, uv);
    OPACITY_PASS(m, position)
    if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
    v3 incoming = outgoing;
    radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
    radiance = add(radiance, mul(muls(weight, pif), eval_bsdfcos(&m, normal, outgoing, incoming)));
    if (!is_delta(&m)) break;
    incoming = sample_delta(&m, normal, outgoing, rand1f(rng));
    if (zero3(incoming)) break;
    weight = mul(weight, divs(eval_delta(&m, normal, outgoing, incoming), sample_delta_pdf(&m, normal, outgoing, incoming)));
    if (zero3(weight) || !finite3(weight)) break;
    ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}

/* ---- trace_furnace, yocto_trace.cpp:1247-1338 (opacity is 1 on supported scenes) ---- */
static trace_result trace_furnace(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) {
  v3 radiance = V3(0, 0, 0), weight = V3(1, 1, 1), hit_albedo = V3(0, 0, 0), hit_normal = V3(0, 0, 0);
  int hit = 0, in_volume = 0, opbounce = 0;
  for (int bounce = 0; bounce < p->bounces; bounce++) {
    if (bounce > 0 && !in_volume) { radiance = add(radiance, mul(weight, eval_environment(sc, ray.d))); break; }
    ygl_intersection isec = intersect_scene(sc, ray, 0);
    if (!isec.hit) {
      if (bounce > 0 || !p->envhidden) radiance = add(radiance, mul(weight, eval_environment(sc, ray.d)));
      break;
    }
    v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
    v3 position = eval_position(sc, isec.instance, isec.element, uv); /* :1281: eval_position, not eval_shading_position */
    v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
    mpoint m = eval_material(sc, isec.instance, isec.element, uv);
    OPACITY_PASS(m, position)
    if (bounce == 0) { hit = 1; hit_albedo = m.color; hit_normal = normal; }
    radiance = add(radiance, mul(weight, eval_emission(&m, normal, outgoing)));
    v3 incoming;
    if (m.roughness != 0) {
      v2 rn = rand2f(rng); float rnl = rand1f(rng);
      incoming = sample_bsdfcos(&m, normal, outgoing, rnl, rn);
      if (zero3(incoming)) break;
      weight = mul(weight, divs(eval_bsdfcos(&m, normal, outgoing, incoming), sample_bsdfcos_pdf(&m, normal, outgoing, incoming)));
    } else {
      incoming = sample_delta(&m, normal, outgoing, rand1f(rng));
      if (zero3(incoming)) break;
      weight = mul(weight, divs(eval_delta(&m, normal, outgoing, incoming), sample_delta_pdf(&m, normal, outgoing, incoming)));
    }
    if (!finish_bounce(&weight, bounce, rng)) break;
    if (dot(normal, outgoing) * dot(normal, incoming) < 0) in_volume = !in_volume;
    ray.o = position; ray.d = incoming; ray.tmin = 1e-4f; ray.tmax = FLT_MAX;
  }
  trace_result r = {radiance, hit, hit_albedo, hit_normal};
  return r;
}
/* ---- trace_falsecolor, yocto_trace.cpp:1341-1419 ---- */
static v3 hashed_color(int id) { /* :1358-1362: std::hash<int> is the identity; rand3f draws x, y, z in order */
  rng_t r = make_rng(961748941, (uint64_t)(size_t)id);
  float x = rand1f(&r), y = rand1f(&r), z = rand1f(&r);
  return V3(powf(0.5f + 0.5f * x, 2.2f), powf(0.5f + 0.5f * y, 2.2f), powf(0.5f + 0.5f * z, 2.2f));
}
static trace_result trace_falsecolor(const oracle_scene* sc, ray_t ray, const ygl_trace_params* p) {
  trace_result none = {V3(0, 0, 0), 0, V3(0, 0, 0), V3(0, 0, 0)};
  ygl_intersection isec = intersect_scene(sc, ray, 0);
  if (!isec.hit) return none;
  v3 outgoing = neg(ray.d); v2 uv = {isec.uv[0], isec.uv[1]};
  v3 position = eval_shading_position(sc, isec.instance, isec.element, uv);
  v3 normal = eval_shading_normal(sc, isec.instance, isec.element, uv, outgoing);
  v3 gnormal = eval_element_normal(sc, isec.instance, isec.element);
  v2 texcoord = eval_texcoord(sc, isec.instance, isec.element, uv);
  mpoint m = eval_material(sc, isec.instance, isec.element, uv);
  float delta = is_delta(&m) ? 1.0f : 0.0f;
  const ygl_instance* in = &sc->d->instances[isec.instance];
  v3 result = V3(0, 0, 0);
  switch (p->falsecolor) {
    case 0: result = adds(muls(position, 0.5f), 0.5f); break;
    case 1: result = adds(muls(normal, 0.5f), 0.5f); break;
    case 2: result = dot(normal, neg(ray.d)) > 0 ? V3(0, 1, 0) : V3(1, 0, 0); break;
    case 3: result = adds(muls(gnormal, 0.5f), 0.5f); break;
    case 4: result = dot(gnormal, neg(ray.d)) > 0 ? V3(0, 1, 0) : V3(1, 0, 0); break;
    case 5: result = V3(fmodf(texcoord.x, 1.0f), fmodf(texcoord.y, 1.0f), 0); break;
    case 6: result = hashed_color(m.type); break;
    case 7: result = m.color; break;
    case 8: result = m.emission; break;
    case 9: result = V3(m.roughness, m.roughness, m.roughness); break;
    case 10: result = V3(m.opacity, m.opacity, m.opacity); break;
    case 11: result = V3(m.metallic, m.metallic, m.metallic); break;
    case 12: result = V3(delta, delta, delta); break;
    case 13: result = hashed_color(isec.instance); break;
    case 14: result = hashed_color(in->shape); break;
    case 15: result = hashed_color(in->material); break;
    case 16: result = hashed_color(isec.element); break;
    case 17: { if (zero3(m.emission)) m.emission = V3(0.2f, 0.2f, 0.2f); result = muls(m.emission, absf_(dot(neg(ray.d), normal))); } break;
    default: result = V3(0, 0, 0);
  }
  trace_result r = {V3(srgb_to_rgb1(result.x), srgb_to_rgb1(result.y), srgb_to_rgb1(result.z)), 1, m.color, normal};
  return r;
}
static trace_result trace_any(const oracle_scene* sc, ray_t ray, rng_t* rng, const ygl_trace_params* p) { /* get_trace_sampler_func, :1422-1438 */
  switch (p->sampler) {
    case YGL_SAMPLER_PATH: return trace_path(sc, ray, rng, p);
    case YGL_SAMPLER_PATHDIRECT: return trace_pathdirect(sc, ray, rng, p);
    case YGL_SAMPLER_PATHMIS: return trace_pathmis(sc, ray, rng, p);
    case YGL_SAMPLER_PATHTEST: return trace_pathtest(sc, ray, rng, p);
    case YGL_SAMPLER_NAIVE: return trace_naive(sc, ray, rng, p);
    case YGL_SAMPLER_EYELIGHT: return trace_eyelight_like(sc, ray, rng, p, 0);
    case YGL_SAMPLER_DIAGRAM: return trace_eyelight_like(sc, ray, rng, p, 1);
    case YGL_SAMPLER_FURNACE: return trace_furnace(sc, ray, rng, p);
    default: return trace_falsecolor(sc, ray, p); /* YGL_SAMPLER_FALSECOLOR (the caller filters unknown values) */
  }
}

/* ---- make_trace_state + trace_samples + trace_image, yocto_trace.cpp:1461-1619 ---- */
void oracle_state_size(const ygl_scene_desc* d, const ygl_trace_params* p, int* w, int* h) {
  const ygl_camera* c = &d->cameras[p->camera];
  if (c->aspect >= 1) { *w = p->resolution; *h = (int)roundf(p->resolution / c->aspect); }
  else { *h = p->resolution; *w = (int)roundf(p->resolution * c->aspect); }
}
void oracle_state_rngs(const ygl_trace_params* p, int w, int h, uint64_t* rngs) {
  rng_t seq = make_rng(1301081, 1);
  for (int64_t i = 0; i < (int64_t)w * h; i++) {
    rng_t r = make_rng(p->seed, (uint64_t)((int)(rng_next(&seq) % 2147483648u) / 2 + 1));
    rngs[2 * i] = r.state; rngs[2 * i + 1] = r.inc;
  }
}
/* image: w*h*4 floats, any of the nine samplers. Returns 0 on success, -1 for anything outside the restatement. */
int oracle_trace_image(const oracle_scene* sc, const ygl_trace_params* p, int nsamples, float* image) {
  if (p->sampler < 0 || p->sampler > YGL_SAMPLER_FALSECOLOR || !oracle_supported(sc->d)) return -1;
  int w, h; oracle_state_size(sc->d, p, &w, &h);
  uint64_t* rngs = malloc(sizeof(uint64_t) * 2 * (size_t)w * h);
  oracle_state_rngs(p, w, h, rngs);
  memset(image, 0, sizeof(float) * 4 * (size_t)w * h);
  const ygl_camera* cam = &sc->d->cameras[p->camera];
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
    size_t idx = (size_t)w * j + i;
    rng_t rng = {rngs[2 * idx], rngs[2 * idx + 1]};
    for (int s = 0; s < nsamples; s++) {
      v2 luv = rand2f(&rng), puv = rand2f(&rng); /* right-to-left: luv first */
      v2 fuv = puv;
      if (p->tentfilter) { /* sample_camera, yocto_trace.cpp:345-357 */
        const float width = 2.0f, offset = 0.5f;
        fuv.x = width * (puv.x < 0.5f ? sqrtf(2 * puv.x) - 1 : 1 - sqrtf(2 - 2 * puv.x)) + offset;
        fuv.y = width * (puv.y < 0.5f ? sqrtf(2 * puv.y) - 1 : 1 - sqrtf(2 - 2 * puv.y)) + offset;
      }
      v2 uv = {(i + fuv.x) / w, (j + fuv.y) / h};
      ray_t ray = eval_camera(cam, uv, sample_disk(luv));
      trace_result r = trace_any(sc, ray, &rng, p);
      v3 rad = r.radiance;
      if (!finite3(rad)) rad = V3(0, 0, 0);
      if (max3(rad) > p->clamp) rad = muls(rad, p->clamp / max3(rad));
      float wgt = 1.0f / (s + 1);
      float* px = image + 4 * idx;
      float src[4] = {0, 0, 0, 0};
      if (r.hit || (!p->envhidden && sc->d->num_environments > 0)) { src[0] = rad.x; src[1] = rad.y; src[2] = rad.z; src[3] = 1; }
      for (int c = 0; c < 4; c++) px[c] = px[c] * (1 - wgt) + src[c] * wgt;
    }
  }
  free(rngs);
  return 0;
}
