// BVH traversal on the device: BVH::intersect / BVH::intersect_p (rustracer-core/src/bvh/mod.rs:366-501)
// over the flattened LinearBVHNode array.  One thread walks one ray through the SAME 2-wide tree in the SAME
// order as the reference (near child first by dir_is_neg[axis], leaves tested immediately, `later hit at equal
// t wins`), because tie-breaks and grazing culls depend on the visit order.
#pragma once
#include "shapes.cuh"

namespace rt {

// Closest-hit record, 16 B.  slot = ordered-primitive slot (0xffffffff = miss); b1,b2 = mesh.rs:296-299.
struct __align__(16) HitRec { float t; uint32_t slot; float b1, b2; };

constexpr uint32_t kMiss = 0xffffffffu;
constexpr int kStackSize = 64;           // bvh/mod.rs:372

struct TravStats { uint32_t nodes, prims; };

// Bounds3::intersect_p_fast (bounds.rs:127-157): no (1+2*gamma3) widening, NaN compares false.
RT_DEV bool slab_test(float4 lo, float4 hi, V3 o, V3 inv_dir, bool nx, bool ny, bool nz, float t_max) {
  float tmin = ((nx ? hi.x : lo.x) - o.x) * inv_dir.x;
  float tmax = ((nx ? lo.x : hi.x) - o.x) * inv_dir.x;
  float tymin = ((ny ? hi.y : lo.y) - o.y) * inv_dir.y;
  float tymax = ((ny ? lo.y : hi.y) - o.y) * inv_dir.y;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  float tzmin = ((nz ? hi.z : lo.z) - o.z) * inv_dir.z;
  float tzmax = ((nz ? lo.z : hi.z) - o.z) * inv_dir.z;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  return tmin < t_max && tmax > 0.0f;
}

// One primitive slot against a ray (GeometricPrimitive::intersect / intersect_p, primitive.rs:45-60).
template <bool STATS>
RT_DEV bool slot_hit_test(const DScene& sc, uint32_t slot, const TriRay& tr, const Ray& ray, float& t, float& b1, float& b2, TravStats* st) {
  const float4 g0 = __ldg(&sc.geom[3 * (size_t)slot]);
  const uint32_t kind_bits = __float_as_uint(g0.w);
  if (STATS) st->prims++;
  if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
    const float4 g1 = __ldg(&sc.geom[3 * (size_t)slot + 1]);
    const float4 g2 = __ldg(&sc.geom[3 * (size_t)slot + 2]);
    float b0;
    return tri_hit_test_pre(tr, ray.t_max, v3(g0), v3(g1), v3(g2), b0, b1, b2, t);
  }
  b1 = 0.0f; b2 = 0.0f;
  return quadric_intersect(sc.quadrics[kind_bits >> 2], ray, t, false, nullptr);
}

// ANY = false: BVH::intersect (closest hit; ray.t_max shrinks; returns hit.slot != kMiss)
// ANY = true : BVH::intersect_p (first accepted hit ends the walk)
// root: node index of the tree's root (0 = the scene's; an object definition's own tree otherwise).  `inst_out`: instance row
// of the hit (kNoInst at the top level); may be null.  Instances recurse one level: TransformedPrimitive (primitive.rs:79-118).
// TOP: the scene's own tree (the only one that holds instances); TOP = false walks a definition's tree and keeps `hit`.
template <bool ANY, bool STATS, bool TOP>
RT_DEV bool bvh_walk(const DScene& sc, Ray& ray, HitRec& hit, TravStats* st, uint32_t root, uint32_t* inst_out) {
  if (TOP) { hit.t = inf_f(); hit.slot = kMiss; hit.b1 = 0.0f; hit.b2 = 0.0f; if (inst_out) *inst_out = kNoInst; }
  if (sc.n_nodes == 0) return false;
  const V3 inv_dir = v3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);                       // bvh/mod.rs:375-380
  const bool nx = inv_dir.x < 0.0f, ny = inv_dir.y < 0.0f, nz = inv_dir.z < 0.0f;
  const TriRay tr = make_tri_ray(ray);
  uint32_t stack[kStackSize];
  int sp = 0;
  uint32_t cur = root;
  bool found = false;
  const float4* __restrict__ nodes = sc.nodes;
  const float4* __restrict__ geom = sc.geom;
  while (true) {
    const float4 lo = __ldg(&nodes[2 * (size_t)cur]);
    const float4 hi = __ldg(&nodes[2 * (size_t)cur + 1]);
    if (STATS) st->nodes++;
    bool descend = false;
    if (slab_test(lo, hi, ray.o, inv_dir, nx, ny, nz, ray.t_max)) {
      const uint32_t meta = __float_as_uint(hi.w), off = __float_as_uint(lo.w);
      const uint32_t n_prims = meta >> 2;
      if (n_prims > 0) {                                                                       // leaf :387-401
        for (uint32_t i = 0; i < n_prims; i++) {
          const uint32_t slot = off + i;
          if (TOP && sc.n_instances && (__float_as_uint(__ldg(&geom[3 * (size_t)slot + 1]).w) & 2u)) {
            // TransformedPrimitive::intersect / intersect_p (primitive.rs:90-102)
            const uint32_t row = __float_as_uint(__ldg(&geom[3 * (size_t)slot + 2]).w);
            const rtgpu_instance& I = sc.instances[row];
            if (STATS) st->prims++;
            Ray r = instance_ray(I, ray);
            bool ok;
            if (I.root_node != 0xffffffffu) ok = bvh_walk<ANY, STATS, false>(sc, r, hit, st, I.root_node, nullptr);
            else {
              float t, b1, b2;
              ok = slot_hit_test<STATS>(sc, I.first_slot, make_tri_ray(r), r, t, b1, b2, st);
              if (ok) { r.t_max = t; hit.t = t; hit.slot = I.first_slot; hit.b1 = b1; hit.b2 = b2; }
            }
            if (ok) {
              if (inst_out) *inst_out = row;
              if (ANY) return true;
              ray.t_max = r.t_max;                                                             // primitive.rs:93-94
              found = true;
            }
            continue;
          }
          float t, b1, b2;
          if (slot_hit_test<STATS>(sc, slot, tr, ray, t, b1, b2, st)) {
            if (ANY) { hit.t = t; hit.slot = slot; return true; }
            ray.t_max = t;                                                                     // primitive.rs:45-51
            hit.t = t; hit.slot = slot; hit.b1 = b1; hit.b2 = b2;                              // `.or(result)`: later hit wins
            if (TOP && inst_out) *inst_out = kNoInst;
            found = true;
          }
        }
      } else {                                                                                 // interior :403-422
        const uint32_t axis = meta & 3u;
        const bool neg = axis == 0 ? nx : (axis == 1 ? ny : nz);
        if (neg) { stack[sp++] = cur + 1; cur = off; }
        else { stack[sp++] = off; cur = cur + 1; }
        descend = true;
      }
    }
    if (!descend) {
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
  return found;
}
template <bool ANY, bool STATS>
RT_DEV bool bvh_traverse(const DScene& sc, Ray& ray, HitRec& hit, TravStats* st, uint32_t* inst_out = nullptr) {
  return bvh_walk<ANY, STATS, true>(sc, ray, hit, st, 0u, inst_out);
}

}  // namespace rt
