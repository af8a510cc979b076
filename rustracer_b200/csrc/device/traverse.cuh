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

// ANY = false: BVH::intersect (closest hit; ray.t_max shrinks; returns hit.slot != kMiss)
// ANY = true : BVH::intersect_p (first accepted hit ends the walk)
template <bool ANY, bool STATS>
RT_DEV bool bvh_traverse(const DScene& sc, Ray& ray, HitRec& hit, TravStats* st) {
  hit.t = inf_f(); hit.slot = kMiss; hit.b1 = 0.0f; hit.b2 = 0.0f;
  if (sc.n_nodes == 0) return false;
  const V3 inv_dir = v3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);                       // bvh/mod.rs:375-380
  const bool nx = inv_dir.x < 0.0f, ny = inv_dir.y < 0.0f, nz = inv_dir.z < 0.0f;
  const TriRay tr = make_tri_ray(ray);
  uint32_t stack[kStackSize];
  int sp = 0;
  uint32_t cur = 0;
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
          const float4 g0 = __ldg(&geom[3 * (size_t)slot]);
          const uint32_t kind_bits = __float_as_uint(g0.w);
          if (STATS) st->prims++;
          float t, b0, b1, b2;
          bool ok;
          if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
            const float4 g1 = __ldg(&geom[3 * (size_t)slot + 1]);
            const float4 g2 = __ldg(&geom[3 * (size_t)slot + 2]);
            ok = tri_hit_test_pre(tr, ray.t_max, v3(g0), v3(g1), v3(g2), b0, b1, b2, t);
          } else {
            b1 = 0.0f; b2 = 0.0f;
            ok = quadric_intersect(sc.quadrics[kind_bits >> 2], ray, t, false, nullptr);
          }
          if (ok) {
            if (ANY) { hit.t = t; hit.slot = slot; return true; }
            ray.t_max = t;                                                                     // primitive.rs:45-51
            hit.t = t; hit.slot = slot; hit.b1 = b1; hit.b2 = b2;                              // `.or(result)`: later hit wins
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
  return hit.slot != kMiss;
}

}  // namespace rt
