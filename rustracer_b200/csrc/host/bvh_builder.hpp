// Host SAH BVH builder (product code).  Produces the same tree, node for node and slot for slot, as
// rustracer's `BVH::recursive_build` + `flatten_bvh` (rustracer-core/src/bvh/mod.rs:137-358), because
// closest-hit ties are resolved by visit order.  Unlike the reference it works on an
// index permutation over SoA bounds and builds independent subtrees on worker threads: a subtree over n
// primitives always owns exactly n consecutive `ordered_prims` slots, the right child's first
// (bvh/mod.rs:290-309), so slot ranges are known before the children are built.
#pragma once
#include <cstdint>
#include <vector>
#include "hmath.hpp"

namespace rth {

struct FlatBvh {
  uvec<float> node_lo, node_hi;          // float4 per node (see include/rtgpu.h)
  uvec<uint32_t> ordered;                // slot -> prim_number
  uint32_t n_nodes = 0, n_leaves = 0, max_leaf_prims = 0;
  double build_seconds = 0;
};

// bounds: one box per primitive in prim_number order.  split_method: RT_SPLIT_*.  threads <= 0: hardware concurrency.
void build_bvh(const std::vector<Box3>& bounds, int max_prims_per_node, int split_method, int threads, FlatBvh& out);

}  // namespace rth
