// Persistent while-while BVH traversal engine (product code, sm_100a).
//
// Same walk as bvh_traverse (traverse.cuh) — rustracer's BVH::intersect / intersect_p (bvh/mod.rs:366-501) — for every
// individual ray: same tree, near child first by dir_is_neg[axis], leaves tested the moment they are reached, later hit
// wins at equal t.  What changes is how the 32 lanes of a warp share the work (profiles/r01a: the one-thread-one-ray loop
// ran with 5.5 of 32 lanes active):
//   * wide nodes: one 64-byte record per interior node holds BOTH children's boxes, so one fetch feeds two slab tests
//     and children that miss are never pushed.  A pushed child keeps its entry distance; it is re-checked against the
//     ray's current t_max when popped, which is exactly the test the reference performs at that moment (its slab test
//     depends on t_max only through the final `tmin < t_max`).
//   * scheduled while-while: each round the warp runs ONE node step for the lanes standing on an interior node, as long
//     as at least `tune_node_threshold` lanes want one; otherwise it serves the lanes standing on a leaf (all primitive
//     tests of that leaf).  Lanes never wait for the slowest lane's whole descent (the plain while-while measured 9 of
//     32 lanes active: the node loop lasted until the last lane found its leaf).
//   * persistent lanes: a lane whose ray is finished commits its result and pulls the next ray from the queue's atomic
//     cursor as soon as enough lanes are idle (ballot + popc aggregated), instead of waiting for the slowest ray.
#pragma once
#include "traverse.cuh"

namespace rt {

#ifndef RT_ENGINE_MIN_BLOCKS
#define RT_ENGINE_MIN_BLOCKS 8      // resident 128-thread blocks per SM the engine kernels are compiled for (register cap 64)
#endif
RT_DEV uint32_t lane_id_() { return threadIdx.x & 31u; }
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kDoneRef = 0xffffffffu;        // (a leaf ref never has all 31 payload bits set: slots < 2^31 - 1)

// Bounds3::intersect_p_fast (bounds.rs:127-157) split into its t_max-independent part and the entry distance.
RT_DEV bool slab_interval(float4 lo, float4 hi, V3 o, V3 inv_dir, bool nx, bool ny, bool nz, float t_max, float& tmin_out) {
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
  tmin_out = tmin;
  return tmin < t_max && tmax > 0.0f;
}

// Policy: RT_DEV void load(uint32_t idx, Ray& ray)          — fetch queue entry idx (lane-private bookkeeping inside)
//         RT_DEV void commit(bool has, uint32_t idx, const HitRec& h)   — called by ALL lanes of the warp, converged;
//                                                               `has` marks lanes with a finished ray
template <bool ANY, class Policy>
RT_DEV void trace_engine(const DScene& sc, uint32_t* cursor, uint32_t n, Policy& pol) {
  enum { NEED = 0, ACTIVE = 1, FINISHED = 2, EXHAUSTED = 3 };
  const unsigned FULL = 0xffffffffu;
  const float4* __restrict__ wide = sc.wide;
  const float4* __restrict__ geom = sc.geom;
  int st = NEED;
  uint32_t idx = 0, cur = kDoneRef;
  Ray ray; V3 inv_dir = v3(0, 0, 0); bool nx = false, ny = false, nz = false;
  TriRay tr; tr.o = v3(0, 0, 0); tr.kx = 0; tr.ky = 1; tr.kz = 2; tr.sx = tr.sy = tr.sz = 0.0f;
  ray = make_ray(v3(0, 0, 0), v3(0, 0, 1), 0.0f);
  HitRec hit; hit.t = inf_f(); hit.slot = kMiss; hit.b1 = hit.b2 = 0.0f;
  uint2 stack[kStackSize];
  int sp = 0;
  bool queue_empty = false;
  const int node_threshold = sc.tune_node_threshold, refill_threshold = sc.tune_refill_threshold;

  while (true) {
    // ---- commit finished rays and pull new ones ------------------------------------------------------------
    const unsigned waiting = __ballot_sync(FULL, st == NEED || st == FINISHED);
    const unsigned active = __ballot_sync(FULL, st == ACTIVE);
    if (active == 0 || __popc(waiting) >= refill_threshold) {
      if (waiting == 0) break;                                         // nothing active, nothing to commit or fetch
      pol.commit(st == FINISHED, idx, hit);
      if (st == FINISHED) st = NEED;
      if (!queue_empty) {
        const bool want = st == NEED;
        const unsigned wmask = __ballot_sync(FULL, want);
        uint32_t base = 0;
        const int leader = __ffs(wmask) - 1;
        if ((int)lane_id_() == leader) base = atomicAdd(cursor, (uint32_t)__popc(wmask));
        base = __shfl_sync(FULL, base, leader);
        if (want) {
          idx = base + (uint32_t)__popc(wmask & ((1u << lane_id_()) - 1u));
          if (idx < n) {
            pol.load(idx, ray);
            inv_dir = v3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);                    // bvh/mod.rs:375-380
            nx = inv_dir.x < 0.0f; ny = inv_dir.y < 0.0f; nz = inv_dir.z < 0.0f;
            tr = make_tri_ray(ray);
            hit.t = inf_f(); hit.slot = kMiss; hit.b1 = 0.0f; hit.b2 = 0.0f;
            sp = 0;
            st = ACTIVE;
            // root: the reference tests the root's own bounds first
            float t0;
            const float4 rlo = make_float4(sc.world_lo[0], sc.world_lo[1], sc.world_lo[2], 0.0f);
            const float4 rhi = make_float4(sc.world_hi[0], sc.world_hi[1], sc.world_hi[2], 0.0f);
            cur = (sc.n_nodes > 0 && slab_interval(rlo, rhi, ray.o, inv_dir, nx, ny, nz, ray.t_max, t0)) ? sc.root_ref : kDoneRef;
          }
        }
        if (__ballot_sync(FULL, want && idx >= n)) queue_empty = true;  // the cursor ran past the end
      }
      if (st == NEED) st = EXHAUSTED;
      if (__ballot_sync(FULL, st == ACTIVE) == 0) {
        if (queue_empty) break;
        continue;
      }
    }

    // ---- schedule: node steps while enough lanes want one, otherwise serve the lanes standing on a leaf ---------
    const bool is_n = st == ACTIVE && cur != kDoneRef && !(cur & kLeafBit);
    const bool is_l = st == ACTIVE && cur != kDoneRef && (cur & kLeafBit);
    const unsigned m_n = __ballot_sync(FULL, is_n), m_l = __ballot_sync(FULL, is_l);
    if (m_l == 0 || __popc(m_n) >= node_threshold) {
      // ---- node step: one wide node = both children's slab tests ----------------------------------------------
      if (is_n) {
        const float4 a = __ldg(&wide[4 * (size_t)cur]);
        const float4 b = __ldg(&wide[4 * (size_t)cur + 1]);
        const float4 c = __ldg(&wide[4 * (size_t)cur + 2]);
        const float4 d = __ldg(&wide[4 * (size_t)cur + 3]);
        float tl, trr;
        const bool hl = slab_interval(a, b, ray.o, inv_dir, nx, ny, nz, ray.t_max, tl);
        const bool hr = slab_interval(c, d, ray.o, inv_dir, nx, ny, nz, ray.t_max, trr);
        const uint32_t rl = __float_as_uint(a.w), rr = __float_as_uint(b.w), axis = __float_as_uint(c.w);
        const bool neg = axis == 0 ? nx : (axis == 1 ? ny : nz);       // bvh/mod.rs:408-421: right child first when negative
        const uint32_t first = neg ? rr : rl, second = neg ? rl : rr;
        const bool hfirst = neg ? hr : hl, hsecond = neg ? hl : hr;
        const float tsecond = neg ? tl : trr;
        if (hfirst) {
          if (hsecond) stack[sp++] = make_uint2(second, __float_as_uint(tsecond));
          cur = first;
        } else if (hsecond) cur = second;
        else {
          cur = kDoneRef;
          while (sp > 0) {
            const uint2 e = stack[--sp];
            if (ANY || __uint_as_float(e.y) < ray.t_max) { cur = e.x; break; }   // the reference's test at visit time
          }
        }
      }
    } else if (is_l) {
      // ---- leaf step: every primitive of the leaf, in slot order (bvh/mod.rs:392-396) ---------------------------
      uint32_t slot = cur & ~kLeafBit;
      bool last = false, done = false;
      do {
        const float4 g0 = __ldg(&geom[3 * (size_t)slot]);
        const float4 g1 = __ldg(&geom[3 * (size_t)slot + 1]);
        const uint32_t kind_bits = __float_as_uint(g0.w);
        last = (__float_as_uint(g1.w) & 1u) != 0;
        float t, b0, b1, b2;
        bool ok;
        if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
          const float4 g2 = __ldg(&geom[3 * (size_t)slot + 2]);
          ok = tri_hit_test_pre(tr, ray.t_max, v3(g0), v3(g1), v3(g2), b0, b1, b2, t);
        } else {
          b1 = 0.0f; b2 = 0.0f;
          ok = quadric_intersect(sc.quadrics[kind_bits >> 2], ray, t, false, nullptr);
        }
        if (ok) {
          hit.t = t; hit.slot = slot; hit.b1 = b1; hit.b2 = b2;
          if (ANY) { done = true; break; }
          ray.t_max = t;
        }
        slot++;
      } while (!last);
      cur = kDoneRef;
      if (!done) {
        while (sp > 0) {
          const uint2 e = stack[--sp];
          if (ANY || __uint_as_float(e.y) < ray.t_max) { cur = e.x; break; }
        }
      }
    }
    if (st == ACTIVE && cur == kDoneRef) st = FINISHED;
  }
}

}  // namespace rt
