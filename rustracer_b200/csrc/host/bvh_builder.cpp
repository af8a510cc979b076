#include "bvh_builder.hpp"
#include <atomic>
#include <chrono>
#include <stdexcept>
#include <thread>

namespace rth {
namespace {

struct BuildNode {
  Box3 box;
  uint32_t child[2];        // interior: node ids ; leaf: {first slot, count}
  uint8_t axis, leaf;
};

struct Builder {
  const std::vector<Box3>& bounds;
  std::vector<float> cx, cy, cz;          // centroids by prim id
  std::vector<uint32_t> perm;             // working permutation of prim ids (the reference permutes `primitive_info`)
  uvec<uint32_t>& ordered;
  std::vector<BuildNode> pool; std::atomic<uint32_t> pool_next{0};
  size_t max_prims; int split_method;
  std::atomic<int> spare_threads{0};

  Builder(const std::vector<Box3>& b, uvec<uint32_t>& ord) : bounds(b), ordered(ord) {}
  const std::vector<float>& cen(int d) const { return d == 0 ? cx : (d == 1 ? cy : cz); }
  uint32_t alloc() { return pool_next.fetch_add(1); }

  // itertools::partition (0.10.3): front scan for a failing element, back scan for a passing one, swap.
  template <class Pred> size_t partition(size_t start, size_t end, Pred pred) {
    size_t front = start, back = end, passed = 0;
    while (front < back) {
      if (!pred(perm[front])) {
        bool swapped = false;
        while (front + 1 < back) {
          back--;
          if (pred(perm[back])) { std::swap(perm[front], perm[back]); swapped = true; break; }
        }
        if (!swapped) return passed;
      }
      passed++; front++;
    }
    return passed;
  }

  uint32_t make_leaf(size_t start, size_t end, size_t base, const Box3& box) {
    uint32_t id = alloc();
    BuildNode& n = pool[id];
    n.box = box; n.leaf = 1; n.axis = 0; n.child[0] = (uint32_t)base; n.child[1] = (uint32_t)(end - start);
    for (size_t i = start; i < end; i++) ordered[base + (i - start)] = perm[i];
    return id;
  }

  uint32_t build(size_t start, size_t end, size_t base, int depth) {
    const size_t n = end - start;
    Box3 box;
    for (size_t i = start; i < end; i++) box.merge(bounds[perm[i]]);
    if (n == 1) return make_leaf(start, end, base, box);
    Box3 cb;
    for (size_t i = start; i < end; i++) cb.grow(v3(cx[perm[i]], cy[perm[i]], cz[perm[i]]));
    const int dim = cb.widest_axis();
    if (cb.lo[dim] == cb.hi[dim]) return make_leaf(start, end, base, box);
    const std::vector<float>& c = cen(dim);
    size_t mid;
    if (split_method == RT_SPLIT_MIDDLE) {                         // bvh/mod.rs:182-200 (`start + partition + start`, Q2)
      float pmid = 0.5f * (cb.lo[dim] + cb.hi[dim]);
      mid = start + partition(start, end, [&](uint32_t id) { return c[id] < pmid; }) + start;
      if (mid == start || mid == end) {
        std::stable_sort(perm.begin() + start, perm.begin() + end, [&](uint32_t a, uint32_t b) { return c[a] < c[b]; });
        mid = (start + end) / 2;
      }
      if (mid <= start || mid >= end) throw std::runtime_error("splitmethod \"middle\": the reference's split index leaves the range here (bvh/mod.rs:186-190) and it panics");
    } else if (n <= 2) {                                           // :204-212
      mid = (start + end) / 2;
      if (start != end - 1 && c[perm[end - 1]] < c[perm[start]]) std::swap(perm[start], perm[end - 1]);
    } else {                                                       // :213-286
      constexpr int NB = 12;
      const float lo = cb.lo[dim], hi = cb.hi[dim];
      auto bucket = [&](uint32_t id) {
        float o = c[id] - lo;
        if (hi > lo) o /= hi - lo;                                 // Bounds3::offset (bounds.rs:177-190)
        float fb = (float)NB * o;
        int b = !(fb == fb) ? 0 : (fb <= 0.0f ? 0 : (fb >= 2147483648.0f ? INT32_MAX : (int)fb));   // saturating `as usize`
        return b == NB ? NB - 1 : b;
      };
      size_t count[NB] = {0}; Box3 bb[NB];
      for (size_t i = start; i < end; i++) { int b = bucket(perm[i]); count[b]++; bb[b].merge(bounds[perm[i]]); }
      // prefix / suffix unions: min/max are exact, so these equal the reference's per-split re-unions (:240-247)
      Box3 pre[NB], suf[NB]; size_t pc[NB], sc[NB];
      { Box3 acc; size_t k = 0; for (int i = 0; i < NB; i++) { acc.merge(bb[i]); k += count[i]; pre[i] = acc; pc[i] = k; } }
      { Box3 acc; size_t k = 0; for (int i = NB - 1; i >= 0; i--) { acc.merge(bb[i]); k += count[i]; suf[i] = acc; sc[i] = k; } }
      float best = 0; int best_b = 0;
      const float inv_total = box.half_area2();
      for (int i = 0; i < NB - 1; i++) {
        float cost = 1.0f + ((float)pc[i] * pre[i].half_area2() + (float)sc[i + 1] * suf[i + 1].half_area2()) / inv_total;
        if (i == 0 || cost < best) { best = cost; best_b = i; }
      }
      if (n > max_prims || best < (float)n) mid = start + partition(start, end, [&](uint32_t id) { return bucket(id) <= best_b; });
      else return make_leaf(start, end, base, box);
    }
    // children: right subtree is built first and owns the first (end - mid) slots
    uint32_t left_id, right_id;
    const size_t right_n = end - mid;
    bool forked = false;
    std::thread worker;
    if (n > 32768 && depth < 12) {
      int s = spare_threads.load();
      while (s > 0 && !spare_threads.compare_exchange_weak(s, s - 1)) {}
      if (s > 0) {
        forked = true;
        worker = std::thread([&, mid, end, base, depth]() { right_id = build(mid, end, base, depth + 1); spare_threads.fetch_add(1); });
      }
    }
    if (!forked) right_id = build(mid, end, base, depth + 1);
    left_id = build(start, mid, base + right_n, depth + 1);
    if (forked) worker.join();
    uint32_t id = alloc();
    BuildNode& nd = pool[id];
    nd.box = pool[left_id].box; nd.box.merge(pool[right_id].box);   // :565-573 union(child1, child2)
    nd.child[0] = left_id; nd.child[1] = right_id; nd.axis = (uint8_t)dim; nd.leaf = 0;
    return id;
  }
};

}  // namespace

void build_bvh(const std::vector<Box3>& bounds, int max_prims_per_node, int split_method, int threads, FlatBvh& out) {
  auto t0 = std::chrono::steady_clock::now();
  out = FlatBvh();
  const size_t N = bounds.size();
  if (N == 0) return;
  if (N >= (1ull << 30)) throw std::runtime_error("too many primitives for 32-bit node offsets");
  out.ordered.assign(N, 0);
  Builder b(bounds, out.ordered);
  b.max_prims = (size_t)std::max(0, max_prims_per_node); b.split_method = split_method;
  b.cx.resize(N); b.cy.resize(N); b.cz.resize(N); b.perm.resize(N);
  for (size_t i = 0; i < N; i++) {                                  // BVHPrimitiveInfo::new (:541-547): 0.5*min + 0.5*max
    b.cx[i] = 0.5f * bounds[i].lo.x + 0.5f * bounds[i].hi.x;
    b.cy[i] = 0.5f * bounds[i].lo.y + 0.5f * bounds[i].hi.y;
    b.cz[i] = 0.5f * bounds[i].lo.z + 0.5f * bounds[i].hi.z;
    b.perm[i] = (uint32_t)i;
  }
  b.pool.resize(2 * N);
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  b.spare_threads = std::max(0, threads - 1);
  uint32_t root = b.build(0, N, 0, 0);
  // flatten: pre-order, left child first (:314-358)
  const uint32_t n_nodes = b.pool_next.load();
  out.n_nodes = n_nodes;
  out.node_lo.resize((size_t)n_nodes * 4); out.node_hi.resize((size_t)n_nodes * 4);
  struct Item { uint32_t node; uint32_t parent_slot; };
  std::vector<Item> stack; stack.reserve(128);
  stack.push_back(Item{root, UINT32_MAX});
  uint32_t next = 0;
  auto put_bits = [](float* dst, uint32_t v) { std::memcpy(dst, &v, 4); };
  while (!stack.empty()) {
    Item it = stack.back(); stack.pop_back();
    const BuildNode& nd = b.pool[it.node];
    uint32_t slot = next++;
    if (it.parent_slot != UINT32_MAX) put_bits(&out.node_lo[(size_t)it.parent_slot * 4 + 3], slot);   // second_child_offset
    float* lo = &out.node_lo[(size_t)slot * 4]; float* hi = &out.node_hi[(size_t)slot * 4];
    lo[0] = nd.box.lo.x; lo[1] = nd.box.lo.y; lo[2] = nd.box.lo.z;
    hi[0] = nd.box.hi.x; hi[1] = nd.box.hi.y; hi[2] = nd.box.hi.z;
    if (nd.leaf) {
      put_bits(&lo[3], nd.child[0]); put_bits(&hi[3], nd.child[1] << 2);
      out.n_leaves++; out.max_leaf_prims = std::max(out.max_leaf_prims, nd.child[1]);
    } else {
      put_bits(&lo[3], 0u); put_bits(&hi[3], (uint32_t)nd.axis);
      stack.push_back(Item{nd.child[1], slot});       // right: visited after the whole left subtree, patches lo.w
      stack.push_back(Item{nd.child[0], UINT32_MAX}); // left: immediately next (slot + 1)
    }
  }
  out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace rth
