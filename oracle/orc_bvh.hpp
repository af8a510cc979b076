// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  BVH build / flatten / traversal.
// Follows /root/reference/rustracer-core/src/bvh/mod.rs.  Third-party arithmetic restated:
// `itertools::partition` (itertools 0.10.3, Cargo.lock:530-531; source not in the mount) — published
// algorithm: scan from the front for the first element failing the predicate, from the back for the
// first element passing it, swap, repeat; return the number of passing elements.
#pragma once
#include "orc_shapes.hpp"

namespace orc {

struct Primitive {                 // primitive.rs:34-75 GeometricPrimitive
  std::shared_ptr<Shape> shape;
  int material = -1;               // index into Scene::materials
  int area_light = -1;             // index into Scene::lights (the DiffuseAreaLight bound to this primitive)
};

struct LinearNode {                // bvh/mod.rs:582-598
  Bounds3 bounds;
  bool leaf; int axis;
  size_t offset;                   // leaf: primitives_offset ; interior: second_child_offset
  size_t n_prims;
};

template <class It, class Pred> inline size_t itertools_partition(It first, It last, Pred pred) {
  size_t split_index = 0;
  while (first != last) {
    if (!pred(*first)) {
      bool found = false;
      while (first != last) {
        --last;
        if (first == last) break;
        if (pred(*last)) { std::swap(*first, *last); found = true; break; }
      }
      if (!found) break;
    }
    split_index++;
    ++first;
  }
  return split_index;
}

struct BVH {
  std::vector<int> ordered;        // ordered_prims -> prim_number (bvh/mod.rs:120)
  std::vector<LinearNode> nodes;
  const std::vector<Primitive>* prims = nullptr;

  struct Info { size_t prim_number; V3 centroid; Bounds3 bounds; };
  struct Build { Bounds3 bounds; Build* c[2] = {nullptr, nullptr}; int axis = 0; size_t first = 0, n = 0; bool leaf = false; };
  std::vector<Build*> pool;
  Build* mk() { Build* b = new Build(); pool.push_back(b); return b; }

  void build(const std::vector<Primitive>& primitives, int max_prims_per_node, int split_method) {    // :80-135
    prims = &primitives;
    nodes.clear(); ordered.clear();
    if (primitives.empty()) return;
    std::vector<Info> info(primitives.size());
    for (size_t i = 0; i < primitives.size(); i++) {
      Bounds3 bb = primitives[i].shape->world_bounds();
      info[i].prim_number = i;
      info[i].centroid = 0.5f * bb.p_min + 0.5f * bb.p_max;                                           // :541-547
      info[i].bounds = bb;
    }
    size_t total = 0;
    ordered.reserve(primitives.size());
    Build* root = recursive_build(info, 0, primitives.size(), (size_t)max_prims_per_node, total, split_method);
    nodes.reserve(total);
    flatten(root);
    for (Build* b : pool) delete b;
    pool.clear();
  }

  Build* leaf(std::vector<Info>& info, size_t start, size_t end, const Bounds3& bounds) {
    Build* b = mk();
    b->leaf = true; b->first = ordered.size(); b->n = end - start; b->bounds = bounds;
    for (size_t i = start; i < end; i++) ordered.push_back((int)info[i].prim_number);
    return b;
  }

  Build* recursive_build(std::vector<Info>& info, size_t start, size_t end, size_t max_prims, size_t& total, int split_method) {  // :137-312
    total += 1;
    size_t n_primitives = end - start;
    Bounds3 bounds;
    for (size_t i = start; i < end; i++) bounds = bunion(bounds, info[i].bounds);
    if (n_primitives == 1) return leaf(info, start, end, bounds);
    Bounds3 cb;
    for (size_t i = start; i < end; i++) cb = bunion_point(cb, info[i].centroid);
    int dim = cb.maximum_extent();
    if (cb.p_min[dim] == cb.p_max[dim]) return leaf(info, start, end, bounds);
    size_t mid;
    if (split_method == 1) {                                                                          // Middle :182-200 (Q2: double `start`)
      float pmid = 0.5f * (cb.p_min[dim] + cb.p_max[dim]);
      mid = start + itertools_partition(info.begin() + start, info.begin() + end, [&](const Info& pi) { return pi.centroid[dim] < pmid; }) + start;
      if (mid == start || mid == end) {
        std::stable_sort(info.begin() + start, info.begin() + end, [&](const Info& a, const Info& b) { return a.centroid[dim] < b.centroid[dim]; });
        mid = (start + end) / 2;
      }
    } else {                                                                                          // SAH :202-287
      if (n_primitives <= 2) {
        mid = (start + end) / 2;
        if (start != end - 1 && info[end - 1].centroid[dim] < info[start].centroid[dim]) std::swap(info[start], info[end - 1]);
      } else {
        const size_t NB = 12;
        struct Bucket { size_t count = 0; Bounds3 bounds; } buckets[NB];
        auto bucket_of = [&](const Info& pi) {
          size_t b = (size_t)f2usize((float)NB * cb.offset(pi.centroid)[dim]);
          if (b == NB) b = NB - 1;
          return b;
        };
        for (size_t i = start; i < end; i++) {
          size_t b = bucket_of(info[i]);
          buckets[b].count += 1;
          buckets[b].bounds = bunion(buckets[b].bounds, info[i].bounds);
        }
        float cost[NB - 1];
        for (size_t i = 0; i < NB - 1; i++) {
          Bounds3 b0, b1; size_t c0 = 0, c1 = 0;
          for (size_t j = 0; j <= i; j++) { b0 = bunion(b0, buckets[j].bounds); c0 += buckets[j].count; }
          for (size_t j = i + 1; j < NB; j++) { b1 = bunion(b1, buckets[j].bounds); c1 += buckets[j].count; }
          cost[i] = 1.0f + ((float)c0 * b0.surface_area() + (float)c1 * b1.surface_area()) / bounds.surface_area();
        }
        float min_cost = cost[0]; size_t min_b = 0;
        for (size_t i = 1; i < NB - 1; i++) if (cost[i] < min_cost) { min_cost = cost[i]; min_b = i; }
        float leaf_cost = (float)n_primitives;
        if (n_primitives > max_prims || min_cost < leaf_cost) {
          mid = start + itertools_partition(info.begin() + start, info.begin() + end, [&](const Info& pi) { return bucket_of(pi) <= min_b; });
        } else {
          return leaf(info, start, end, bounds);
        }
      }
    }
    // right subtree first (:290-309, Q3)
    Build* right = recursive_build(info, mid, end, max_prims, total, split_method);
    Build* left = recursive_build(info, start, mid, max_prims, total, split_method);
    Build* b = mk();
    b->bounds = bunion(left->bounds, right->bounds);                                                  // :565-573
    b->c[0] = left; b->c[1] = right; b->axis = dim;
    return b;
  }

  size_t flatten(Build* node) {                                                                        // :314-358
    size_t offset = nodes.size();
    LinearNode ln;
    ln.bounds = node->bounds;
    if (node->leaf) {
      ln.leaf = true; ln.axis = 0; ln.offset = node->first; ln.n_prims = node->n;
      nodes.push_back(ln);
    } else {
      ln.leaf = false; ln.axis = node->axis; ln.offset = 0; ln.n_prims = 0;
      nodes.push_back(ln);
      flatten(node->c[0]);
      size_t second = flatten(node->c[1]);
      nodes[offset].offset = second;
    }
    return offset;
  }

  Bounds3 world_bounds() const { return nodes[0].bounds; }                                             // :362-364

  // :366-433.  Returns prim_number of the hit (or -1); fills si/t; shrinks ray.t_max.
  int intersect(Ray& ray, SurfaceInteraction& result) const {
    if (nodes.empty()) return -1;
    int hit_prim = -1;
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    V3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
    int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    Counters& cnt = tls_counters();
    while (true) {
      const LinearNode& node = nodes[cur];
      cnt.nodes_visited++;
      if (bounds_intersect_p_fast(node.bounds, ray, inv_dir, dir_is_neg)) {
        if (node.leaf) {
          for (size_t i = 0; i < node.n_prims; i++) {
            int pn = ordered[node.offset + i];
            SurfaceInteraction si; float t;
            cnt.prims_tested++;
            if ((*prims)[pn].shape->intersect(ray, si, t)) {       // primitive.rs:45-51 ; `.or(result)` keeps the later hit (Q9)
              si.prim = pn; ray.t_max = t; result = si; hit_prim = pn;
            }
          }
          if (to_visit == 0) break;
          cur = stack[--to_visit];
        } else {
          if (dir_is_neg[node.axis]) { stack[to_visit++] = cur + 1; cur = node.offset; }
          else { stack[to_visit++] = node.offset; cur = cur + 1; }
        }
      } else {
        if (to_visit == 0) break;
        cur = stack[--to_visit];
      }
    }
    return hit_prim;
  }

  // :435-501
  bool intersect_p(const Ray& ray) const {
    if (nodes.empty()) return false;
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    V3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
    int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    Counters& cnt = tls_counters();
    while (true) {
      const LinearNode& node = nodes[cur];
      cnt.nodes_visited++;
      if (bounds_intersect_p_fast(node.bounds, ray, inv_dir, dir_is_neg)) {
        if (node.leaf) {
          for (size_t i = 0; i < node.n_prims; i++) {
            cnt.prims_tested++;
            if ((*prims)[ordered[node.offset + i]].shape->intersect_p(ray)) return true;
          }
          if (to_visit == 0) break;
          cur = stack[--to_visit];
        } else {
          if (dir_is_neg[node.axis]) { stack[to_visit++] = cur + 1; cur = node.offset; }
          else { stack[to_visit++] = node.offset; cur = cur + 1; }
        }
      } else {
        if (to_visit == 0) break;
        cur = stack[--to_visit];
      }
    }
    return false;
  }
};

// ---- object instancing: TransformedPrimitive (primitive.rs:79-118) over the aggregate built at the first
// ObjectInstance (api.rs:1071-1080: a BVH when the definition holds more than one primitive, else the primitive itself) ----
struct ObjectDef {
  std::vector<Primitive> prims;
  BVH bvh;
  bool aggregate = false;
  void finish(int max_prims_per_node, int split_method) {
    aggregate = prims.size() > 1;
    if (aggregate) bvh.build(prims, max_prims_per_node, split_method);
  }
  Bounds3 world_bounds() const { return aggregate ? bvh.world_bounds() : prims[0].shape->world_bounds(); }
};

// SurfaceInteraction::transform (interaction.rs:156-190)
inline SurfaceInteraction transform_interaction(const SurfaceInteraction& s, const Transform& t) {
  SurfaceInteraction o = s;
  V3 p_err;
  V3 p = t.point_with_error(s.hit.p, s.hit.p_error, p_err);
  o.hit = Interaction::make(p, p_err, normalize(t.vector(s.hit.wo)), normalize(t.normal(s.hit.n)));
  o.dpdu = t.vector(s.dpdu); o.dpdv = t.vector(s.dpdv);
  o.shading.n = normalize(t.normal(s.shading.n));
  o.shading.dpdu = t.vector(s.shading.dpdu); o.shading.dpdv = t.vector(s.shading.dpdv);
  o.shading.n = face_forward(o.shading.n, o.hit.n);
  return o;
}

struct InstanceShape : Shape {
  std::shared_ptr<ObjectDef> def;
  Transform p2w;
  InstanceShape(std::shared_ptr<ObjectDef> d, const Transform& t) : def(std::move(d)), p2w(t) {}
  Bounds3 world_bounds() const override { return p2w.bounds(def->world_bounds()); }                   // primitive.rs:86-88
  // `primitive_to_world.inverse() * ray` (ray.rs:83-93): origin and direction only, no error offset, t_max kept
  Ray to_object(const Ray& ray) const { Transform w2p = p2w.inverse(); return Ray(w2p.point(ray.o), w2p.vector(ray.d), ray.t_max); }
  bool intersect(const Ray& ray, SurfaceInteraction& si, float& t) const override {                   // primitive.rs:90-97
    Ray r = to_object(ray);
    SurfaceInteraction inner;
    int material;
    if (def->aggregate) {
      int pn = def->bvh.intersect(r, inner);
      if (pn < 0) return false;
      material = def->prims[pn].material;
      t = r.t_max;
    } else {
      tls_counters().prims_tested++;
      if (!def->prims[0].shape->intersect(r, inner, t)) return false;
      material = def->prims[0].material;
    }
    si = transform_interaction(inner, p2w);
    si.material_override = material;
    return true;
  }
  bool intersect_p(const Ray& ray) const override {                                                   // primitive.rs:99-102
    Ray r = to_object(ray);
    if (def->aggregate) return def->bvh.intersect_p(r);
    tls_counters().prims_tested++;
    return def->prims[0].shape->intersect_p(r);
  }
  float area() const override { return 0.0f; }
  void sample(P2, Interaction&, float& pdf) const override { pdf = 0.0f; }                            // never a light (primitive.rs:104-106)
};

}  // namespace orc
