// Persistent while-while BVH traversal engine (product code, sm_100a).
//
// Same walk as bvh_traverse (traverse.cuh) — rustracer's BVH::intersect / intersect_p (bvh/mod.rs:366-501) — for every
// individual ray: same tree, near child first by dir_is_neg[axis], leaves tested the moment they are reached, later hit
// wins at equal t.  What changes is how the 32 lanes of a warp share the work (profiles/r01a: the one-thread-one-ray loop
// ran with 5.5 of 32 lanes active):
//   * wide nodes: one 64-byte record per interior node holds BOTH children's boxes, so one fetch feeds two slab tests
//     and children that miss are never pushed.  With RT_ENGINE_WIDE4 (default) two levels of the reference's binary tree are
//     collapsed into one 128-byte record with the boxes of the four grandchildren c0..c3 and the three split axes: the step
//     tests all four and visits them in the order the binary near-first rule would reach them (the near child's near and far
//     grandchild, then the far child's), so the sequence of leaves a ray visits is the reference's.  The skipped intermediate
//     node needs no test of its own: Bounds3::intersect_p_fast is monotone in the box (every product (plane - o) * inv_dir
//     rounds monotonically, NaN compares false on both sides alike), so a grandchild that passes implies its parent passes
//     at the same t_max.  47 node steps per ray become ~26 dependent fetches (profiles/r02c).  A pushed child keeps its entry distance; it is re-checked against the
//     ray's current t_max when popped, which is exactly the test the reference performs at that moment (its slab test
//     depends on t_max only through the final `tmin < t_max`).
//   * scheduled while-while: each round the warp runs ONE node step for the lanes standing on an interior node, as long
//     as at least `tune_node_threshold` lanes want one; otherwise it serves the lanes standing on a leaf (all primitive
//     tests of that leaf).  Lanes never wait for the slowest lane's whole descent.
//   * persistent lanes: a lane whose ray is finished pulls the next ray from the queue's atomic cursor as soon as
//     `tune_refill_threshold` lanes are idle (ballot + popc aggregated).  A refill is cheap on purpose — the policy's
//     commit is a plain store executed by the finished lanes only, and everything that needs the whole warp (material
//     classification) lives in its own streaming kernel (k_classify) — so the threshold can be low and few lanes idle.
//   * the lane state is encoded in `cur` (interior ref / leaf ref / kDoneRef), two ballots per round (profiles/r01c: the
//     six-ballot version spent 20 % of its issue slots on scheduling).
#pragma once
#include "traverse.cuh"

namespace rt {

#ifndef RT_ENGINE_MIN_BLOCKS
#define RT_ENGINE_MIN_BLOCKS 8      // resident 128-thread blocks per SM the engine kernels are compiled for (register cap 64)
#endif
#ifndef RT_ENGINE_PREFETCH
#define RT_ENGINE_PREFETCH 0        // 1 / 2: prefetch a pushed subtree's record into L1 / L2 (measured: see profiles/)
#endif
#ifndef RT_ENGINE_LDG256
#define RT_ENGINE_LDG256 1          // collapsed nodes are fetched with four 256-bit loads instead of eight 128-bit ones
#endif
#ifndef RT_ENGINE_ANY_FIXED_ORDER
#define RT_ENGINE_ANY_FIXED_ORDER 1 // 1: any-hit walks visit a collapsed node's children in storage order (2: near pair first, storage order inside a pair) (the occlusion answer is order-independent; profiles/r02n: C3 +3.6 %, AO +12 %)
#endif
#ifndef RT_ENGINE_SMEM_DEPTH
#define RT_ENGINE_SMEM_DEPTH 8      // traversal-stack entries per lane kept in shared memory; deeper entries go to local memory
#endif
RT_DEV uint32_t lane_id_() { return threadIdx.x & 31u; }
// One 256-bit read-only load (sm_100: LDG.E.256.CONSTANT) of two consecutive float4; p must be 32-byte aligned.  The engine is bound
// by the L1 data pipe (l1tex__data_pipe_lsu_wavefronts 89 % of peak with 128-bit loads, profiles/r02f): a lane's node costs one
// wavefront per load instruction, so half as many instructions per node are half as many wavefronts.
RT_DEV void ldg256(const float4* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kDoneRef = 0xffffffffu;        // (a leaf ref never has all 31 payload bits set: slots < 2^31 - 1)
constexpr uint32_t kHitSlotMask = (1u << kHitSlotBits) - 1u;   // inside the engine a hit slot carries its shade-queue id in the top 3 bits
constexpr uint32_t kExitInstance = 0xfffffffeu;   // stack marker: the walk below this entry happens in world space again

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

// Bounds3::intersect_p_fast, branch-free: the same comparisons on the same values (NaN compares false, exactly like
// the early returns of the reference), evaluated for every lane so that the two boxes of a wide node interleave.
RT_DEV bool slab_interval_bf(float4 lo, float4 hi, V3 o, V3 inv_dir, bool nx, bool ny, bool nz, float t_max, float& tmin_out) {
  float tmin = ((nx ? hi.x : lo.x) - o.x) * inv_dir.x;
  float tmax = ((nx ? lo.x : hi.x) - o.x) * inv_dir.x;
  const float tymin = ((ny ? hi.y : lo.y) - o.y) * inv_dir.y;
  const float tymax = ((ny ? lo.y : hi.y) - o.y) * inv_dir.y;
  const bool r1 = (tmin > tymax) | (tymin > tmax);
  tmin = tymin > tmin ? tymin : tmin;
  tmax = tymax < tmax ? tymax : tmax;
  const float tzmin = ((nz ? hi.z : lo.z) - o.z) * inv_dir.z;
  const float tzmax = ((nz ? lo.z : hi.z) - o.z) * inv_dir.z;
  const bool r2 = (tmin > tzmax) | (tzmin > tmax);
  tmin = tzmin > tmin ? tzmin : tmin;
  tmax = tzmax < tmax ? tzmax : tmax;
  tmin_out = tmin;
  return !r1 & !r2 & (tmin < t_max) & (tmax > 0.0f);
}

// Policy: RT_DEV void load(uint32_t idx, Ray& ray)                            — fetch queue entry idx (lane-private bookkeeping inside)
//         RT_DEV void commit(uint32_t idx, const HitRec& h, float t_hit, uint32_t inst, uint32_t cls)
//                                                                             — called by the lanes holding a finished ray (divergent);
//                                                                               h.t holds the FIRST barycentric of a triangle hit (the distance is t_hit:
//                                                                               during a closest-hit walk it is ray.t_max, one register instead of two),
//                                                                               inst = instance row of the hit or kNoInst
// INST: the scene holds object instances (TransformedPrimitive, primitive.rs:79-118).  A leaf primitive flagged as an
// instance suspends the leaf: an exit marker with the resume point goes on the stack, the ray moves to object space
// (`primitive_to_world.inverse() * ray`, same t parametrisation) and the walk continues at the definition's root; popping
// the marker restores the world-space ray (re-read from the queue) and resumes the leaf.  Same visit order as the
// reference's nested BVH::intersect call.  Compiled out entirely for scenes without instances.
template <bool ANY, bool INST, class Policy>
RT_DEV void trace_engine(const DScene& sc, uint32_t* cursor, uint32_t n, Policy& pol) {
  const unsigned FULL = 0xffffffffu;
  const float4* __restrict__ wide = sc.wide;
  const float4* __restrict__ geom = sc.geom;
  const uint32_t lane = lane_id_(), lane_lt = (1u << lane) - 1u;
  uint32_t idx = 0, cur = kDoneRef;         // cur: interior ref (top bit clear) | leaf ref (kLeafBit | slot) | kDoneRef (no ray in flight)
  bool pending = false;                     // a finished ray whose result is not committed yet
  Ray ray = make_ray(v3(0, 0, 0), v3(0, 0, 1), 0.0f);
  V3 inv_dir = v3(0, 0, 0); bool nx = false, ny = false, nz = false;
  TriRay tr; tr.o = v3(0, 0, 0); tr.kx = 0; tr.ky = 1; tr.kz = 2; tr.sx = tr.sy = tr.sz = 0.0f;
  HitRec hit; hit.t = 0.0f; hit.slot = kMiss; hit.b1 = hit.b2 = 0.0f;   // hit.t: barycentric b0 of the best hit (see Policy::commit)
  // Traversal stack (64 entries, bvh/mod.rs:372): the bottom RT_ENGINE_SMEM_DEPTH entries of every lane live in shared
  // memory as [entry][thread] (conflict-free, ~25-cycle pops, no L1 traffic: profiles/r01c-v2a showed more local-memory
  // sectors than global ones and 19 % of the stall samples on the pop), the rarely used rest in local memory.
  __shared__ uint2 s_stack[RT_ENGINE_SMEM_DEPTH][128];
  // capacity: the reference's 64 entries per tree (bvh/mod.rs:372); with instances the scene's tree, the exit marker and the
  // definition's tree share this one stack
#if RT_ENGINE_WIDE4
  constexpr int kTreeStack = (kStackSize / 2) * 3;   // up to three postponed grandchildren per collapsed pair of levels
#else
  constexpr int kTreeStack = kStackSize;
#endif
  uint2 stack_l[(INST ? 2 * kTreeStack + 1 : kTreeStack) - RT_ENGINE_SMEM_DEPTH];
  const uint32_t tid = threadIdx.x;
#if RT_ENGINE_TOP_NODES > 0
  // top levels of the tree (breadth-first numbered at upload) staged in shared memory: every ray walks them
  __shared__ float4 s_top[4 * RT_ENGINE_TOP_NODES];
  const uint32_t n_top = sc.n_top;
  for (uint32_t k = tid; k < 4u * n_top; k += blockDim.x) s_top[k] = __ldg(&wide[k]);
  __syncthreads();
#endif
  int sp = 0;
  bool queue_empty = n == 0;
  uint32_t negmask = 0;                     // bit k = dir_is_neg[k]
  uint32_t inst = kNoInst, hit_inst = kNoInst;   // INST: instance the walk is inside of / instance of the best hit
  const int node_threshold = sc.tune_node_threshold, refill_threshold = sc.tune_refill_threshold;

  // (re)derive the per-ray constants of the walk from a ray's origin and direction; ray.t_max is left alone
#define RT_ENGINE_SET_RAY(r_) do { \
    ray.o = (r_).o; \
    inv_dir = v3(1.0f / (r_).d.x, 1.0f / (r_).d.y, 1.0f / (r_).d.z);                         /* bvh/mod.rs:375-380 */ \
    nx = inv_dir.x < 0.0f; ny = inv_dir.y < 0.0f; nz = inv_dir.z < 0.0f; \
    negmask = (nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u); \
    tr = make_tri_ray(r_); } while (0)
#define RT_ENGINE_PUSH(e_) do { \
    const uint2 pe_ = (e_); \
    if (sp < RT_ENGINE_SMEM_DEPTH) s_stack[sp][tid] = pe_; else stack_l[sp - RT_ENGINE_SMEM_DEPTH] = pe_; \
    sp++; } while (0)
  // next subtree whose entry distance is still in front of the hit (the reference's test at visit time), or finished
#define RT_ENGINE_POP() do { \
    cur = kDoneRef; \
    while (sp > 0) { \
      --sp; \
      const uint2 e_ = sp < RT_ENGINE_SMEM_DEPTH ? s_stack[sp][tid] : stack_l[sp - RT_ENGINE_SMEM_DEPTH]; \
      if (INST && e_.x == kExitInstance) { \
        Ray wr_; pol.load(idx, wr_); \
        RT_ENGINE_SET_RAY(wr_); \
        inst = kNoInst; \
        if (e_.y != kDoneRef) { cur = e_.y; break; } \
        continue; } \
      if (ANY || __uint_as_float(e_.y) < ray.t_max) { cur = e_.x; break; } } \
    pending = cur == kDoneRef; } while (0)

  while (true) {
    const unsigned m_n = __ballot_sync(FULL, !(cur & kLeafBit));
    const unsigned m_l = __ballot_sync(FULL, (cur & kLeafBit) != 0u && cur != kDoneRef);
    const unsigned busy = m_n | m_l;
    if (busy == 0u || (!queue_empty && busy != FULL && 32 - __popc(busy) >= refill_threshold)) {
      // ---- commit finished rays and pull new ones ----------------------------------------------------------
      if (pending) {
        HitRec hc = hit; uint32_t cls = (uint32_t)Q_MISS_CLASS;
        if (hit.slot != kMiss) { cls = hit.slot >> kHitSlotBits; hc.slot = hit.slot & kHitSlotMask; }
        pol.commit(idx, hc, hit.slot != kMiss ? ray.t_max : inf_f(), hit_inst, cls);   // (an any-hit policy ignores the distance)
        pending = false;
      }
      if (queue_empty) break;                                          // busy == 0 and nothing left to fetch
      const unsigned wmask = ~busy;
      const int leader = __ffs(wmask) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(wmask));
      base = __shfl_sync(FULL, base, leader);
      if ((wmask >> lane) & 1u) {
        idx = base + (uint32_t)__popc(wmask & lane_lt);
        if (idx < n) {
          pol.load(idx, ray);
          RT_ENGINE_SET_RAY(ray);
          hit.t = 0.0f; hit.slot = kMiss; hit.b1 = 0.0f; hit.b2 = 0.0f;
          sp = 0;
          if (INST) { inst = kNoInst; hit_inst = kNoInst; }
          // root: the reference tests the root's own bounds first
          float t0;
          const float4 rlo = make_float4(sc.world_lo[0], sc.world_lo[1], sc.world_lo[2], 0.0f);
          const float4 rhi = make_float4(sc.world_hi[0], sc.world_hi[1], sc.world_hi[2], 0.0f);
          const bool in = sc.n_nodes > 0 && slab_interval(rlo, rhi, ray.o, inv_dir, nx, ny, nz, ray.t_max, t0);
          cur = in ? sc.root_ref : kDoneRef;
          pending = !in;                                               // a ray that misses the world is finished: commit the miss
        }
      }
      if (base + (uint32_t)__popc(wmask) >= n) queue_empty = true;     // warp-uniform: the cursor ran past the end
      continue;
    }

    // ---- schedule: node steps while enough lanes want one, otherwise serve the lanes standing on a leaf ---------
    if (m_l == 0u || __popc(m_n) >= node_threshold) {
      // ---- node step: one wide node = both children's slab tests ----------------------------------------------
      if (!(cur & kLeafBit)) {
        // the stack top, read by every lane of the step while the node is in flight: a lane whose two children both miss takes it
        // from here instead of running the pop loop on its own (profiles/r01r: the divergent pop was 8 % of the issue slots at 5 lanes)
        const bool top_in_smem = sp > 0 && sp <= RT_ENGINE_SMEM_DEPTH;
        const uint2 top = top_in_smem ? s_stack[sp - 1][tid] : make_uint2(kExitInstance, 0u);
#if RT_ENGINE_WIDE4
        const float4* __restrict__ nd = wide + 8 * (size_t)cur;
        float4 q0, q1, q2, q3, q4, q5, q6, q7;
#if RT_ENGINE_LDG256
        ldg256(nd, q0, q1); ldg256(nd + 2, q2, q3); ldg256(nd + 4, q4, q5); ldg256(nd + 6, q6, q7);
#else
        q0 = __ldg(nd); q1 = __ldg(nd + 1); q2 = __ldg(nd + 2); q3 = __ldg(nd + 3);
        q4 = __ldg(nd + 4); q5 = __ldg(nd + 5); q6 = __ldg(nd + 6); q7 = __ldg(nd + 7);
#endif
        const uint32_t r0 = __float_as_uint(q0.w), r1 = __float_as_uint(q1.w), axes = __float_as_uint(q2.w), r2 = __float_as_uint(q3.w), r3 = __float_as_uint(q4.w);
        float t0, t1, t2, t3;
        const bool h0 = slab_interval_bf(q0, q1, ray.o, inv_dir, nx, ny, nz, ray.t_max, t0);
        const bool h1 = slab_interval_bf(q2, q3, ray.o, inv_dir, nx, ny, nz, ray.t_max, t1) & (r1 != kDoneRef);
        const bool h2 = slab_interval_bf(q4, q5, ray.o, inv_dir, nx, ny, nz, ray.t_max, t2);
        const bool h3 = slab_interval_bf(q6, q7, ray.o, inv_dir, nx, ny, nz, ray.t_max, t3) & (r3 != kDoneRef);
#if RT_ENGINE_ANY_FIXED_ORDER
        // intersect_p answers "is anything in the way": the answer does not depend on the order the subtrees are searched in (no t_max
        // shrinks during an any-hit walk, so the set of leaves that can be reached is the same), and the near-first permutation below is a
        // tenth of the node step's instructions.  Any-hit walks therefore take the children in storage order.
        const bool fixed_order = ANY;
#else
        const bool fixed_order = false;
#endif
        // bvh/mod.rs:408-421 at the binary node and at each of its children: the second child first when the ray is negative along the split axis
        const bool negA = (!fixed_order || RT_ENGINE_ANY_FIXED_ORDER == 2) && ((negmask >> (axes & 3u)) & 1u) != 0u, negL = !fixed_order && ((negmask >> ((axes >> 2) & 3u)) & 1u) != 0u,
                   negR = !fixed_order && ((negmask >> ((axes >> 4) & 3u)) & 1u) != 0u;
        const uint32_t la_r = negL ? r1 : r0, lb_r = negL ? r0 : r1, ra_r = negR ? r3 : r2, rb_r = negR ? r2 : r3;
        const float la_t = negL ? t1 : t0, lb_t = negL ? t0 : t1, ra_t = negR ? t3 : t2, rb_t = negR ? t2 : t3;
        const bool la_h = negL ? h1 : h0, lb_h = negL ? h0 : h1, ra_h = negR ? h3 : h2, rb_h = negR ? h2 : h3;
        const uint32_t e0r = negA ? ra_r : la_r, e1r = negA ? rb_r : lb_r, e2r = negA ? la_r : ra_r, e3r = negA ? lb_r : rb_r;
        const float e1t = negA ? rb_t : lb_t, e2t = negA ? la_t : ra_t, e3t = negA ? lb_t : rb_t;
        const bool e0h = negA ? ra_h : la_h, e1h = negA ? rb_h : lb_h, e2h = negA ? la_h : ra_h, e3h = negA ? lb_h : rb_h;
        // the first entry that passes is visited now; the others wait on the stack, nearest on top, each with its entry distance
        if (e3h & (e0h | e1h | e2h)) RT_ENGINE_PUSH(make_uint2(e3r, __float_as_uint(e3t)));
        if (e2h & (e0h | e1h)) RT_ENGINE_PUSH(make_uint2(e2r, __float_as_uint(e2t)));
        if (e1h & e0h) RT_ENGINE_PUSH(make_uint2(e1r, __float_as_uint(e1t)));
        if (e0h) cur = e0r;
        else if (e1h) cur = e1r;
        else if (e2h) cur = e2r;
        else if (e3h) cur = e3r;
        else if (top_in_smem && top.x != kExitInstance && (ANY || __uint_as_float(top.y) < ray.t_max)) { cur = top.x; --sp; }   // == the first iteration of the pop loop
        else RT_ENGINE_POP();
#else
        float4 a, b, c, d;
#if RT_ENGINE_TOP_NODES > 0
        if (cur < n_top) { a = s_top[4 * cur]; b = s_top[4 * cur + 1]; c = s_top[4 * cur + 2]; d = s_top[4 * cur + 3]; }
        else
#endif
        {
          a = __ldg(&wide[4 * (size_t)cur]);
          b = __ldg(&wide[4 * (size_t)cur + 1]);
          c = __ldg(&wide[4 * (size_t)cur + 2]);
          d = __ldg(&wide[4 * (size_t)cur + 3]);
        }
        float tl, trr;
        const bool hl = slab_interval_bf(a, b, ray.o, inv_dir, nx, ny, nz, ray.t_max, tl);
        const bool hr = slab_interval_bf(c, d, ray.o, inv_dir, nx, ny, nz, ray.t_max, trr);
        const uint32_t rl = __float_as_uint(a.w), rr = __float_as_uint(b.w), axis = __float_as_uint(c.w);
        const bool neg = ((negmask >> axis) & 1u) != 0u;               // bvh/mod.rs:408-421: right child first when negative
        const uint32_t first = neg ? rr : rl, second = neg ? rl : rr;
        const bool hfirst = neg ? hr : hl, hsecond = neg ? hl : hr;
        const float tsecond = neg ? tl : trr;
        if (hfirst) {
          if (hsecond) {
            RT_ENGINE_PUSH(make_uint2(second, __float_as_uint(tsecond)));
#if RT_ENGINE_PREFETCH
            {   // the postponed subtree's record will be needed after the near subtree: start its fetch now
              const char* pa = (second & kLeafBit) ? (const char*)&geom[3 * (size_t)(second & ~kLeafBit)] : (const char*)&wide[4 * (size_t)second];
#if RT_ENGINE_PREFETCH == 2
              asm volatile("prefetch.global.L2 [%0];" :: "l"(pa));
              asm volatile("prefetch.global.L2 [%0];" :: "l"(pa + 32));
#else
              asm volatile("prefetch.global.L1 [%0];" :: "l"(pa));
              asm volatile("prefetch.global.L1 [%0];" :: "l"(pa + 32));
#endif
            }
#endif
          }
          cur = first;
        } else if (hsecond) cur = second;
        else if (top_in_smem && top.x != kExitInstance && (ANY || __uint_as_float(top.y) < ray.t_max)) { cur = top.x; --sp; }   // == the first iteration of the pop loop
        else RT_ENGINE_POP();
#endif
      }
    } else if ((cur & kLeafBit) != 0u && cur != kDoneRef) {
      // ---- leaf step: every primitive of the leaf, in slot order (bvh/mod.rs:392-396) ---------------------------
      uint32_t slot = cur & ~kLeafBit;
      bool last = false, done = false, entered = false;
      do {
        const float4 g0 = __ldg(&geom[3 * (size_t)slot]);
        const float4 g1 = __ldg(&geom[3 * (size_t)slot + 1]);
        const uint32_t kind_bits = __float_as_uint(g0.w);
        last = (__float_as_uint(g1.w) & 1u) != 0;
        float t, b0, b1, b2;
        bool ok;
        if (INST && (__float_as_uint(g1.w) & 2u)) {                    // TransformedPrimitive::intersect / intersect_p (primitive.rs:90-102)
          const uint32_t row = __float_as_uint(__ldg(&geom[3 * (size_t)slot + 2]).w);
          const rtgpu_instance& I = sc.instances[row];
          Ray wr; pol.load(idx, wr);
          const Ray orr = instance_ray(I, wr);
          const V3 iinv = v3(1.0f / orr.d.x, 1.0f / orr.d.y, 1.0f / orr.d.z);
          float t0;
          // an aggregate tests its root bounds first (bvh/mod.rs:382-386); a one-primitive definition is the primitive itself
          const bool in = I.root_node == 0xffffffffu ||
                          slab_interval(make_float4(I.lo[0], I.lo[1], I.lo[2], 0.0f), make_float4(I.hi[0], I.hi[1], I.hi[2], 0.0f), orr.o, iinv,
                                        iinv.x < 0.0f, iinv.y < 0.0f, iinv.z < 0.0f, ray.t_max, t0);
          if (in) {
            RT_ENGINE_PUSH(make_uint2(kExitInstance, last ? kDoneRef : (kLeafBit | (slot + 1u))));
            RT_ENGINE_SET_RAY(orr);
            inst = row; cur = I.root_ref;
            entered = true;
            break;
          }
          slot++;
          continue;
        }
        if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
          const float4 g2 = __ldg(&geom[3 * (size_t)slot + 2]);
          ok = tri_hit_test_pre(tr, ray.t_max, v3(g0), v3(g1), v3(g2), b0, b1, b2, t);
        } else {
          b0 = 0.0f; b1 = 0.0f; b2 = 0.0f;
          Ray rq; pol.load(idx, rq);                                   // the direction is not kept in registers across the walk:
          if (INST && inst != kNoInst) rq = instance_ray(sc.instances[inst], rq);
          rq.t_max = ray.t_max;                                        // re-read it for the (rare) quadric test
          ok = quadric_intersect(sc.quadrics[kind_bits >> 2], rq, t, false, nullptr);
        }
        if (ok) {
          hit.t = b0; hit.slot = slot | (((__float_as_uint(g1.w) >> kGeomClassShift) & 7u) << kHitSlotBits); hit.b1 = b1; hit.b2 = b2;
          if (INST) hit_inst = inst;
          if (ANY) { done = true; break; }
          ray.t_max = t;
        }
        slot++;
      } while (!last);
      if (done) { cur = kDoneRef; pending = true; }
      else if (!entered) RT_ENGINE_POP();
    }
  }
#undef RT_ENGINE_POP
#undef RT_ENGINE_PUSH
#undef RT_ENGINE_SET_RAY
}

}  // namespace rt
