// SAH BVH construction on the device (product code, sm_100a): rtgpu_build_bvh.
//
// Builds THE SAME tree as rustracer's `BVH::recursive_build` + `flatten_bvh` (rustracer-core/src/bvh/mod.rs:137-358) — node for
// node, slot for slot — because closest-hit ties are resolved by visit order and the parity tests compare per-ray node counts.
// The reference recurses depth-first on one thread; this builder runs the same per-node decisions level by level:
//   * a node over n primitives always owns n consecutive `ordered_prims` slots, the right child's first (bvh/mod.rs:290-309),
//     so slot ranges are known top-down and the processing order does not matter;
//   * bounds are min / max reductions and bucket counts are integers: order-independent, done with warp-aggregated atomics;
//   * the 12-bucket SAH choice is evaluated per node with the reference's float expressions;
//   * `itertools::partition` (0.10.3: front scan for a failing element, back scan for a passing one, swap) has a closed form:
//     with m = number of passing elements, the k-th failing element of [0, m) in ascending order is exchanged with the k-th
//     passing element of [m, n) in descending order — two stream compactions (one prefix sum) and a pairwise swap;
//   * nodes of at most 32 primitives are handled by one thread each with the sequential code;
//   * the linear pre-order layout comes from subtree sizes (bottom-up over the levels) and offsets (top-down).
// `splitmethod "middle"` and the per-definition trees of object instances stay on the host builder.
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <cfloat>
#include <cstring>
#include <vector>
#include "context.hpp"

namespace rt {
namespace bvhb {

constexpr int NB = 12;
constexpr uint32_t kSmall = 32;
constexpr uint32_t kNone = 0xffffffffu;
enum : uint8_t { ST_PENDING = 0, ST_LEAF = 1, ST_SPLIT = 2 };

// monotonic float <-> int key for atomicMin / atomicMax
__host__ __device__ __forceinline__ int fkey(float f) { int i; memcpy(&i, &f, 4); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ __forceinline__ float funkey(int k) { int i = k >= 0 ? k : k ^ 0x7fffffff; float f; memcpy(&f, &i, 4); return f; }

struct Nodes {                 // build nodes, SoA, capacity 2N
  uint32_t *start, *end, *base, *left, *right, *mid, *bslot, *size, *off;
  uint8_t *state, *axis; int8_t* best;
  int* box;                    // 6 keys per node: lo.xyz, hi.xyz (large nodes: atomics; small nodes: written once)
  int* cbox;                   // centroid bounds, 6 keys per node (large nodes only use them across kernels)
};
struct Work {
  const float* bounds;         // 6 per primitive: lo.xyz, hi.xyz
  uint32_t n;
  uint32_t* perm;              // position -> primitive id (the reference permutes `primitive_info`)
  uint32_t* node_of;           // position -> build node of the current level (kNone: finished)
  unsigned long long* flags;   // n + 1: low = misplaced-failing, high = misplaced-passing; scanned in place (exclusive)
  uint32_t *flist, *tlist;
  uint32_t* ordered;           // slot -> primitive id
  uint32_t* bcount; int* bbox; // per large node of the level: 12 counts, 12 x 6 keys
  uint32_t* counters;          // [0] nodes allocated, [1] large nodes of the level (bucket slots), [2] depth overflow flag
  uint32_t max_prims;
};

__device__ __forceinline__ float centroid(const float* bounds, uint32_t id, int d) { return 0.5f * bounds[6 * (size_t)id + d] + 0.5f * bounds[6 * (size_t)id + 3 + d]; }   // bvh/mod.rs:541-547
__device__ __forceinline__ int widest_axis(float lx, float ly, float lz, float hx, float hy, float hz) {   // bounds.rs:77-90
  const float dx = hx - lx, dy = hy - ly, dz = hz - lz;
  return dx > dy ? (dx > dz ? 0 : 2) : (dy > dz ? 1 : 2);
}
__device__ __forceinline__ int bucket_of(float c, float lo, float hi) {   // bvh/mod.rs:223-232 + Bounds3::offset (bounds.rs:177-190)
  float o = c - lo;
  if (hi > lo) o /= hi - lo;
  const float fb = (float)NB * o;
  int b = __float2int_rz(fb);                                         // saturating `as usize`: NaN -> 0 ...
  if (b < 0) b = 0;                                                   // ... negative -> 0
  return b >= NB ? NB - 1 : b;                                        // `if b == n_buckets { b = n_buckets - 1 }`; larger cannot happen (the centroid lies inside its bounds)
}
__device__ __forceinline__ float half_area2(const float* lo, const float* hi) {   // bounds.rs:213-216
  const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return 2.0f * (dx * dy + dx * dz + dy * dz);
}
// The SAH choice of bvh/mod.rs:234-262 from the 12 bucket counts / boxes: returns the best bucket and its cost.
__device__ __forceinline__ void sah_choose(const uint32_t* count, const float (*blo)[3], const float (*bhi)[3], const float* box_lo, const float* box_hi,
                                           int& best_b, float& best, uint32_t* passed) {
  float plo[NB][3], phi[NB][3], slo[NB][3], shi[NB][3]; uint32_t pc[NB], sc[NB];
  { float al[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ah[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t k = 0;
    for (int i = 0; i < NB; i++) { for (int d = 0; d < 3; d++) { al[d] = al[d] < blo[i][d] ? al[d] : blo[i][d]; ah[d] = ah[d] > bhi[i][d] ? ah[d] : bhi[i][d]; plo[i][d] = al[d]; phi[i][d] = ah[d]; } k += count[i]; pc[i] = k; } }
  { float al[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ah[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t k = 0;
    for (int i = NB - 1; i >= 0; i--) { for (int d = 0; d < 3; d++) { al[d] = al[d] < blo[i][d] ? al[d] : blo[i][d]; ah[d] = ah[d] > bhi[i][d] ? ah[d] : bhi[i][d]; slo[i][d] = al[d]; shi[i][d] = ah[d]; } k += count[i]; sc[i] = k; } }
  const float inv_total = half_area2(box_lo, box_hi);
  best = 0.0f; best_b = 0;
  for (int i = 0; i < NB - 1; i++) {
    const float cost = 1.0f + ((float)pc[i] * half_area2(plo[i], phi[i]) + (float)sc[i + 1] * half_area2(slo[i + 1], shi[i + 1])) / inv_total;
    if (i == 0 || cost < best) { best = cost; best_b = i; }
  }
  *passed = pc[best_b];
}

__device__ __forceinline__ uint32_t alloc_children(const Nodes& nd, const Work& w, uint32_t node, uint32_t start, uint32_t mid, uint32_t end, uint32_t base) {
  const uint32_t id = atomicAdd(&w.counters[0], 2u);                  // left = id, right = id + 1
  const uint32_t right_n = end - mid;
  nd.start[id] = start; nd.end[id] = mid; nd.base[id] = base + right_n; nd.state[id] = ST_PENDING;   // the right subtree owns the first slots
  nd.start[id + 1] = mid; nd.end[id + 1] = end; nd.base[id + 1] = base; nd.state[id + 1] = ST_PENDING;
  nd.left[node] = id; nd.right[node] = id + 1; nd.mid[node] = mid; nd.state[node] = ST_SPLIT;
  return id;
}

// ---- per level, large nodes (n > kSmall) --------------------------------------------------------------------------------
__global__ void k_prepare(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  nd.bslot[node] = kNone;
  if (nd.end[node] - nd.start[node] <= kSmall) return;
  const uint32_t s = atomicAdd(&w.counters[1], 1u);
  nd.bslot[node] = s;
  for (int d = 0; d < 3; d++) { nd.box[6 * (size_t)node + d] = fkey(FLT_MAX); nd.box[6 * (size_t)node + 3 + d] = fkey(-FLT_MAX); nd.cbox[6 * (size_t)node + d] = fkey(FLT_MAX); nd.cbox[6 * (size_t)node + 3 + d] = fkey(-FLT_MAX); }
  for (int b = 0; b < NB; b++) {
    w.bcount[(size_t)s * NB + b] = 0;
    for (int d = 0; d < 3; d++) { w.bbox[((size_t)s * NB + b) * 6 + d] = fkey(FLT_MAX); w.bbox[((size_t)s * NB + b) * 6 + 3 + d] = fkey(-FLT_MAX); }
  }
}
// Reductions into per-node (or per-bucket) min / max keys and counts.  Lanes of a warp that share a destination reduce with
// __reduce_min/max/add_sync over their __match_any_sync group; when the whole block works on one node (the top levels, where a
// single node spans millions of positions) the warps meet in shared memory first, so a block issues one global atomic per
// destination instead of one per warp.
__device__ __forceinline__ void group_min_max(unsigned group, bool leader, int* dst_lo, int* dst_hi, int klo, int khi) {
  const int mn = __reduce_min_sync(group, klo), mx = __reduce_max_sync(group, khi);
  if (leader) { atomicMin(dst_lo, mn); atomicMax(dst_hi, mx); }
}
// box and centroid bounds of every large node of the level (bvh/mod.rs:153-159, :170-174)
__global__ void __launch_bounds__(256) k_bounds(Nodes nd, Work w, uint32_t first, uint32_t count) {
  __shared__ int sh[12];
  __shared__ uint32_t sh_node;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t node = i < w.n ? w.node_of[i] : kNone;
  if (node != kNone && (node < first || node >= first + count || nd.bslot[node] == kNone)) node = kNone;
  if (threadIdx.x == 0) sh_node = node;
  if (threadIdx.x < 12) sh[threadIdx.x] = threadIdx.x % 6 < 3 ? fkey(FLT_MAX) : fkey(-FLT_MAX);   // {box lo, box hi, cbox lo, cbox hi}
  __syncthreads();
  const bool uniform = __syncthreads_and(node == sh_node || i >= w.n) && sh_node != kNone;
  int kb[6] = {0, 0, 0, 0, 0, 0}, kc[3] = {0, 0, 0};
  if (node != kNone) {
    const uint32_t id = w.perm[i];
    for (int d = 0; d < 3; d++) {
      const float lo = w.bounds[6 * (size_t)id + d], hi = w.bounds[6 * (size_t)id + 3 + d];
      kb[d] = fkey(lo); kb[3 + d] = fkey(hi); kc[d] = fkey(0.5f * lo + 0.5f * hi);
    }
  }
  const unsigned group = __match_any_sync(0xffffffffu, node);
  const bool leader = (int)(threadIdx.x & 31) == __ffs(group) - 1 && node != kNone;
  int* box = uniform ? sh : &nd.box[6 * (size_t)(node == kNone ? 0 : node)];
  int* cbox = uniform ? sh + 6 : &nd.cbox[6 * (size_t)(node == kNone ? 0 : node)];
  for (int d = 0; d < 3; d++) {
    group_min_max(group, leader, &box[d], &box[3 + d], kb[d], kb[3 + d]);
    group_min_max(group, leader, &cbox[d], &cbox[3 + d], kc[d], kc[d]);
  }
  if (uniform) {
    __syncthreads();
    if (threadIdx.x < 12) {
      int* dst = threadIdx.x < 6 ? &nd.box[6 * (size_t)sh_node + threadIdx.x] : &nd.cbox[6 * (size_t)sh_node + threadIdx.x - 6];
      if (threadIdx.x % 6 < 3) atomicMin(dst, sh[threadIdx.x]); else atomicMax(dst, sh[threadIdx.x]);
    }
  }
}
// split axis; a node whose centroids coincide on it becomes a leaf (bvh/mod.rs:175-180)
__global__ void k_choose_dim(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  if (nd.bslot[node] == kNone) return;
  float cl[3], ch[3];
  for (int d = 0; d < 3; d++) { cl[d] = funkey(nd.cbox[6 * (size_t)node + d]); ch[d] = funkey(nd.cbox[6 * (size_t)node + 3 + d]); }
  const int dim = widest_axis(cl[0], cl[1], cl[2], ch[0], ch[1], ch[2]);
  nd.axis[node] = (uint8_t)dim;
  if (cl[dim] == ch[dim]) nd.state[node] = ST_LEAF;
}
// bucket counts and bucket bounds (bvh/mod.rs:219-232)
__global__ void __launch_bounds__(256) k_bucket(Nodes nd, Work w, uint32_t first, uint32_t count) {
  __shared__ int sh_box[NB * 6];
  __shared__ uint32_t sh_cnt[NB];
  __shared__ uint32_t sh_node;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t node = i < w.n ? w.node_of[i] : kNone;
  if (node != kNone && (node < first || node >= first + count || nd.bslot[node] == kNone || nd.state[node] == ST_LEAF)) node = kNone;
  if (threadIdx.x == 0) sh_node = node;
  if (threadIdx.x < NB * 6) sh_box[threadIdx.x] = threadIdx.x % 6 < 3 ? fkey(FLT_MAX) : fkey(-FLT_MAX);
  if (threadIdx.x < NB) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  const bool uniform = __syncthreads_and(node == sh_node || i >= w.n) && sh_node != kNone;
  int b = 0, kb[6] = {0, 0, 0, 0, 0, 0};
  uint32_t dest = kNone;                                              // bucket row: bslot * NB + b
  if (node != kNone) {
    const int dim = nd.axis[node];
    const uint32_t id = w.perm[i];
    b = bucket_of(centroid(w.bounds, id, dim), funkey(nd.cbox[6 * (size_t)node + dim]), funkey(nd.cbox[6 * (size_t)node + 3 + dim]));
    dest = nd.bslot[node] * NB + (uint32_t)b;
    for (int d = 0; d < 6; d++) kb[d] = fkey(w.bounds[6 * (size_t)id + d]);
  }
  const unsigned group = __match_any_sync(0xffffffffu, dest);
  const bool leader = (int)(threadIdx.x & 31) == __ffs(group) - 1 && node != kNone;
  uint32_t* cnt = uniform ? &sh_cnt[b] : &w.bcount[dest == kNone ? 0 : dest];
  int* box = uniform ? &sh_box[b * 6] : &w.bbox[(size_t)(dest == kNone ? 0 : dest) * 6];
  if (leader) atomicAdd(cnt, (uint32_t)__popc(group));
  for (int d = 0; d < 3; d++) group_min_max(group, leader, &box[d], &box[3 + d], kb[d], kb[3 + d]);
  if (uniform) {
    __syncthreads();
    const size_t row = (size_t)nd.bslot[sh_node] * NB;
    if (threadIdx.x < NB && sh_cnt[threadIdx.x]) atomicAdd(&w.bcount[row + threadIdx.x], sh_cnt[threadIdx.x]);
    if (threadIdx.x < NB * 6 && sh_cnt[threadIdx.x / 6]) {
      if (threadIdx.x % 6 < 3) atomicMin(&w.bbox[row * 6 + threadIdx.x], sh_box[threadIdx.x]); else atomicMax(&w.bbox[row * 6 + threadIdx.x], sh_box[threadIdx.x]);
    }
  }
}
// SAH split or leaf (bvh/mod.rs:234-286)
__global__ void k_sah(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  if (nd.bslot[node] == kNone || nd.state[node] == ST_LEAF) return;
  const size_t s = (size_t)nd.bslot[node] * NB;
  uint32_t cnt[NB]; float blo[NB][3], bhi[NB][3], lo[3], hi[3];
  for (int b = 0; b < NB; b++) { cnt[b] = w.bcount[s + b]; for (int d = 0; d < 3; d++) { blo[b][d] = funkey(w.bbox[(s + b) * 6 + d]); bhi[b][d] = funkey(w.bbox[(s + b) * 6 + 3 + d]); } }
  for (int d = 0; d < 3; d++) { lo[d] = funkey(nd.box[6 * (size_t)node + d]); hi[d] = funkey(nd.box[6 * (size_t)node + 3 + d]); }
  int best_b; float best; uint32_t passed;
  sah_choose(cnt, blo, bhi, lo, hi, best_b, best, &passed);
  const uint32_t n = nd.end[node] - nd.start[node];
  if (n > w.max_prims || best < (float)n) { nd.best[node] = (int8_t)best_b; nd.mid[node] = nd.start[node] + passed; nd.state[node] = ST_SPLIT; }
  else nd.state[node] = ST_LEAF;
}
// elements on the wrong side of the split point
__global__ void k_flags(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > w.n) return;
  unsigned long long f = 0;
  if (i < w.n) {
    const uint32_t node = w.node_of[i];
    if (node != kNone && node >= first && node < first + count && nd.bslot[node] != kNone && nd.state[node] == ST_SPLIT) {
      const int dim = nd.axis[node];
      const int b = bucket_of(centroid(w.bounds, w.perm[i], dim), funkey(nd.cbox[6 * (size_t)node + dim]), funkey(nd.cbox[6 * (size_t)node + 3 + dim]));
      const bool pass = b <= (int)nd.best[node];
      if (i < nd.mid[node] && !pass) f = 1ull;
      else if (i >= nd.mid[node] && pass) f = 1ull << 32;
    }
  }
  w.flags[i] = f;
}
// ---- exclusive prefix sum of n 64-bit items (two 32-bit counters packed), three kernels -----------------------------------
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
__device__ unsigned long long block_exclusive(unsigned long long v, unsigned long long* total, unsigned long long* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = v;
  for (int off = 1; off < 32; off <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long s = lane < (kScanThreads / 32) ? sh[lane] : 0;
    for (int off = 1; off < 32; off <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, s, off); if (lane >= off) s += y; }
    if (lane < (kScanThreads / 32)) sh[lane] = s;
  }
  __syncthreads();
  const unsigned long long before = warp ? sh[warp - 1] : 0;
  if (total) *total = sh[kScanThreads / 32 - 1];
  __syncthreads();
  return before + x - v;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(unsigned long long* data, uint32_t n, unsigned long long* tile_sums) {
  __shared__ unsigned long long sh[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  unsigned long long v[kScanItems], sum = 0;
  for (int k = 0; k < kScanItems; k++) { v[k] = base + k < n ? data[base + k] : 0; sum += v[k]; }
  unsigned long long total;
  unsigned long long ex = block_exclusive(sum, &total, sh);
  for (int k = 0; k < kScanItems; k++) { if (base + k < n) data[base + k] = ex; ex += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_sums(unsigned long long* tile_sums, uint32_t n_tiles) {
  __shared__ unsigned long long sh[kScanThreads / 32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n_tiles; b0 += kScanThreads) {
    const uint32_t i = b0 + threadIdx.x;
    const unsigned long long v = i < n_tiles ? tile_sums[i] : 0;
    unsigned long long total;
    const unsigned long long ex = block_exclusive(v, &total, sh);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}
__global__ void k_scan_add(unsigned long long* data, uint32_t n, const unsigned long long* tile_sums) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) data[i] += tile_sums[i / kScanTile];
}
// compaction of the misplaced elements, in position order (flags now hold the exclusive prefix sums)
__global__ void k_scatter(Work w) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.n) return;
  const unsigned long long a = w.flags[i], b = w.flags[i + 1];
  if ((uint32_t)b != (uint32_t)a) w.flist[(uint32_t)a] = i;
  if ((uint32_t)(b >> 32) != (uint32_t)(a >> 32)) w.tlist[(uint32_t)(a >> 32)] = i;
}
// itertools::partition's exchanges: k-th misplaced failing element from the front <-> k-th misplaced passing element from the back
__global__ void k_swap(Nodes nd, Work w) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = (uint32_t)w.flags[w.n];
  if (k >= total) return;
  const uint32_t i = w.flist[k];
  const uint32_t node = w.node_of[i];
  const uint32_t seg_first = (uint32_t)w.flags[nd.start[node]], m = (uint32_t)w.flags[nd.end[node]] - seg_first;
  const uint32_t j = w.tlist[seg_first + (m - 1u - (k - seg_first))];
  const uint32_t a = w.perm[i]; w.perm[i] = w.perm[j]; w.perm[j] = a;
}
__global__ void k_children_large(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  if (nd.bslot[node] == kNone || nd.state[node] != ST_SPLIT) return;
  alloc_children(nd, w, node, nd.start[node], nd.mid[node], nd.end[node], nd.base[node]);
}
// positions follow their node: into the left / right child, or out (leaf: its primitives go to their `ordered_prims` slots)
__global__ void k_assign(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.n) return;
  const uint32_t node = w.node_of[i];
  if (node == kNone || node < first || node >= first + count) return;
  if (nd.bslot[node] == kNone) { w.node_of[i] = kNone; return; }     // small node: its thread takes the whole range from here on
  if (nd.state[node] == ST_LEAF) { w.ordered[nd.base[node] + (i - nd.start[node])] = w.perm[i]; w.node_of[i] = kNone; }
  else w.node_of[i] = i < nd.mid[node] ? nd.left[node] : nd.right[node];
}

// ---- per level, small nodes: one thread runs `recursive_build`'s body for its node (bvh/mod.rs:148-312) --------------------
__global__ void k_small(Nodes nd, Work w, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  const uint32_t start = nd.start[node], end = nd.end[node], base = nd.base[node], n = end - start;
  if (n > kSmall) return;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, cl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ch[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i = start; i < end; i++) {
    const float* b = w.bounds + 6 * (size_t)w.perm[i];
    for (int d = 0; d < 3; d++) {
      lo[d] = lo[d] < b[d] ? lo[d] : b[d]; hi[d] = hi[d] > b[3 + d] ? hi[d] : b[3 + d];
      const float c = 0.5f * b[d] + 0.5f * b[3 + d];
      if (c < cl[d]) cl[d] = c;
      if (c > ch[d]) ch[d] = c;
    }
  }
  for (int d = 0; d < 3; d++) { nd.box[6 * (size_t)node + d] = fkey(lo[d]); nd.box[6 * (size_t)node + 3 + d] = fkey(hi[d]); }
  const int dim = widest_axis(cl[0], cl[1], cl[2], ch[0], ch[1], ch[2]);
  nd.axis[node] = (uint8_t)dim;
  bool leaf = n == 1 || cl[dim] == ch[dim];
  uint32_t mid = 0;
  if (!leaf) {
    if (n <= 2) {                                                     // bvh/mod.rs:204-212
      mid = (start + end) / 2;
      if (start != end - 1 && centroid(w.bounds, w.perm[end - 1], dim) < centroid(w.bounds, w.perm[start], dim)) { const uint32_t a = w.perm[start]; w.perm[start] = w.perm[end - 1]; w.perm[end - 1] = a; }
    } else {
      uint32_t cnt[NB]; float blo[NB][3], bhi[NB][3];
      for (int b = 0; b < NB; b++) { cnt[b] = 0; for (int d = 0; d < 3; d++) { blo[b][d] = FLT_MAX; bhi[b][d] = -FLT_MAX; } }
      for (uint32_t i = start; i < end; i++) {
        const uint32_t id = w.perm[i];
        const int b = bucket_of(centroid(w.bounds, id, dim), cl[dim], ch[dim]);
        cnt[b]++;
        const float* pb = w.bounds + 6 * (size_t)id;
        for (int d = 0; d < 3; d++) { blo[b][d] = blo[b][d] < pb[d] ? blo[b][d] : pb[d]; bhi[b][d] = bhi[b][d] > pb[3 + d] ? bhi[b][d] : pb[3 + d]; }
      }
      int best_b; float best; uint32_t passed;
      sah_choose(cnt, blo, bhi, lo, hi, best_b, best, &passed);
      if (n > w.max_prims || best < (float)n) {
        // itertools::partition, sequentially
        uint32_t front = start, back = end;
        while (front < back) {
          if (!(bucket_of(centroid(w.bounds, w.perm[front], dim), cl[dim], ch[dim]) <= best_b)) {
            bool swapped = false;
            while (front + 1 < back) {
              back--;
              if (bucket_of(centroid(w.bounds, w.perm[back], dim), cl[dim], ch[dim]) <= best_b) { const uint32_t a = w.perm[front]; w.perm[front] = w.perm[back]; w.perm[back] = a; swapped = true; break; }
            }
            if (!swapped) break;
          }
          front++;
        }
        mid = start + passed;
      } else leaf = true;
    }
  }
  if (leaf) {
    nd.state[node] = ST_LEAF;
    for (uint32_t i = start; i < end; i++) w.ordered[base + (i - start)] = w.perm[i];
  } else alloc_children(nd, w, node, start, mid, end, base);
}

// ---- linear layout (flatten_bvh, bvh/mod.rs:314-358): pre-order, left child first ---------------------------------------------
__global__ void k_sizes(Nodes nd, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  nd.size[node] = nd.state[node] == ST_LEAF ? 1u : 1u + nd.size[nd.left[node]] + nd.size[nd.right[node]];
}
__global__ void k_offsets(Nodes nd, uint32_t first, uint32_t count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t node = first + k;
  if (nd.state[node] == ST_LEAF) return;
  nd.off[nd.left[node]] = nd.off[node] + 1u;
  nd.off[nd.right[node]] = nd.off[node] + 1u + nd.size[nd.left[node]];
}
__global__ void k_emit(Nodes nd, uint32_t n_nodes, float4* node_lo, float4* node_hi) {
  const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  const uint32_t o = nd.off[node];
  float lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = funkey(nd.box[6 * (size_t)node + d]); hi[d] = funkey(nd.box[6 * (size_t)node + 3 + d]); }
  if (nd.state[node] == ST_LEAF) {
    node_lo[o] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(nd.base[node]));
    node_hi[o] = make_float4(hi[0], hi[1], hi[2], __uint_as_float((nd.end[node] - nd.start[node]) << 2));
  } else {
    node_lo[o] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(nd.off[nd.right[node]]));   // second_child_offset
    node_hi[o] = make_float4(hi[0], hi[1], hi[2], __uint_as_float((uint32_t)nd.axis[node]));
  }
}
__global__ void k_init(Nodes nd, Work w) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.n) { w.perm[i] = i; w.node_of[i] = 0; }
  if (i == 0) { nd.start[0] = 0; nd.end[0] = w.n; nd.base[0] = 0; nd.state[0] = ST_PENDING; nd.off[0] = 0; w.counters[0] = 1; w.counters[1] = 0; w.counters[2] = 0; }
}

}  // namespace bvhb
}  // namespace rt

using namespace rt;
using namespace rt::bvhb;

extern "C" int rtgpu_build_bvh(rtgpu_ctx* ctx, const float* prim_bounds, uint64_t n_prims, int max_prims_per_node, float* node_lo, float* node_hi, uint32_t* ordered,
                               uint32_t* n_nodes_out, float* build_ms) {
  if (!ctx || !prim_bounds || !node_lo || !node_hi || !ordered || !n_nodes_out) return RTGPU_ERR_ARG;
  if (n_prims == 0 || n_prims >= (1ull << 30)) return fail(ctx, RTGPU_ERR_ARG, "rtgpu_build_bvh: primitive count must be in [1, 2^30)");
  RT_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t N = (uint32_t)n_prims, cap = 2 * N;
  const bool timing = std::getenv("RT_UPLOAD_TIMING") != nullptr;     // wall time of the steps around the build on stderr (tools/upload_probe.py)
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[build_bvh] %-25s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  // one arena for the ~35 work arrays: a cudaMalloc / cudaFree pair per array was 90 ms of a 26 ms build (profiles/r02v_upload_probe_c4.log)
  struct Req { void** p; size_t bytes; };
  std::vector<Req> reqs;
  char* arena = nullptr;
  auto release = [&]() { if (arena) cudaFree(arena); arena = nullptr; };
  int rc = 0;
#define DA(ptr, count) reqs.push_back(Req{(void**)&(ptr), sizeof(*(ptr)) * (size_t)(count)})
  Nodes nd{}; Work w{};
  DA(nd.start, cap); DA(nd.end, cap); DA(nd.base, cap); DA(nd.left, cap); DA(nd.right, cap); DA(nd.mid, cap); DA(nd.bslot, cap); DA(nd.size, cap); DA(nd.off, cap);
  DA(nd.state, cap); DA(nd.axis, cap); DA(nd.best, cap); DA(nd.box, (size_t)cap * 6); DA(nd.cbox, (size_t)cap * 6);
  float* d_bounds = nullptr; float4 *d_lo = nullptr, *d_hi = nullptr; unsigned long long* tile_sums = nullptr;
  const uint32_t max_large = N / (kSmall + 1) + 1, n_tiles = (N + 1 + kScanTile - 1) / kScanTile;
  DA(d_bounds, (size_t)N * 6); DA(w.perm, N); DA(w.node_of, N); DA(w.flags, (size_t)N + 1); DA(w.flist, N); DA(w.tlist, N); DA(w.ordered, N);
  DA(w.bcount, (size_t)max_large * NB); DA(w.bbox, (size_t)max_large * NB * 6); DA(w.counters, 4); DA(tile_sums, n_tiles); DA(d_lo, cap); DA(d_hi, cap);
  {
    size_t total_bytes = 0;
    for (const Req& r : reqs) total_bytes += (r.bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc((void**)&arena, total_bytes ? total_bytes : 256);
    if (e != cudaSuccess) { arena = nullptr; return check_cuda(ctx, e, "cudaMalloc (rtgpu_build_bvh)"); }
    size_t at = 0;
    for (const Req& r : reqs) { *r.p = arena + at; at += (r.bytes + 255) & ~(size_t)255; }
  }
#undef DA
  lap("device allocations");
  w.bounds = d_bounds; w.n = N; w.max_prims = (uint32_t)(max_prims_per_node < 0 ? 0 : max_prims_per_node);
  cudaStream_t s = ctx->stream;
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = check_cuda(ctx, _e, #call); release(); return rc; } } while (0)
  CK(cudaMemcpyAsync(d_bounds, prim_bounds, sizeof(float) * 6 * (size_t)N, cudaMemcpyHostToDevice, s));
  lap("bounds to the device");
  CK(cudaEventRecord(ctx->ev0, s));
  const unsigned pos_blocks = (N + 255) / 256, pos1_blocks = (N + 1 + 255) / 256;
  k_init<<<pos_blocks, 256, 0, s>>>(nd, w); ctx->launches++;
  std::vector<std::pair<uint32_t, uint32_t>> levels;
  uint32_t first = 0, count = 1, total = 1;
  while (count > 0) {
    if (levels.size() > 4096) { release(); return fail(ctx, RTGPU_ERR_UNSUPPORTED, "rtgpu_build_bvh: tree deeper than 4096 levels"); }
    levels.push_back({first, count});
    const unsigned nb = (count + 255) / 256;
    CK(cudaMemsetAsync(&w.counters[1], 0, 4, s));
    k_prepare<<<nb, 256, 0, s>>>(nd, w, first, count); ctx->launches++;
    uint32_t n_large = 0;
    CK(cudaMemcpyAsync(&n_large, &w.counters[1], 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_large > 0) {
      k_bounds<<<pos_blocks, 256, 0, s>>>(nd, w, first, count);
      k_choose_dim<<<nb, 256, 0, s>>>(nd, w, first, count);
      k_bucket<<<pos_blocks, 256, 0, s>>>(nd, w, first, count);
      k_sah<<<nb, 256, 0, s>>>(nd, w, first, count);
      k_flags<<<pos1_blocks, 256, 0, s>>>(nd, w, first, count);
      k_scan_tiles<<<n_tiles, kScanThreads, 0, s>>>(w.flags, N + 1, tile_sums);
      k_scan_sums<<<1, kScanThreads, 0, s>>>(tile_sums, n_tiles);
      k_scan_add<<<pos1_blocks, 256, 0, s>>>(w.flags, N + 1, tile_sums);
      k_scatter<<<pos_blocks, 256, 0, s>>>(w);
      k_swap<<<pos_blocks, 256, 0, s>>>(nd, w);
      k_children_large<<<nb, 256, 0, s>>>(nd, w, first, count);
      ctx->launches += 11;
    }
    k_small<<<(count + 63) / 64, 64, 0, s>>>(nd, w, first, count); ctx->launches++;
    if (n_large > 0) { k_assign<<<pos_blocks, 256, 0, s>>>(nd, w, first, count); ctx->launches++; }
    uint32_t new_total = 0;
    CK(cudaMemcpyAsync(&new_total, &w.counters[0], 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    first = total; count = new_total - total; total = new_total;
  }
  for (size_t l = levels.size(); l-- > 0;) { k_sizes<<<(levels[l].second + 255) / 256, 256, 0, s>>>(nd, levels[l].first, levels[l].second); ctx->launches++; }
  for (size_t l = 0; l < levels.size(); l++) { k_offsets<<<(levels[l].second + 255) / 256, 256, 0, s>>>(nd, levels[l].first, levels[l].second); ctx->launches++; }
  k_emit<<<(total + 255) / 256, 256, 0, s>>>(nd, total, d_lo, d_hi); ctx->launches++;
  CK(cudaEventRecord(ctx->ev1, s));
  lap("build kernels");
  CK(cudaMemcpyAsync(node_lo, d_lo, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(node_hi, d_hi, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ordered, w.ordered, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (build_ms) CK(cudaEventElapsedTime(build_ms, ctx->ev0, ctx->ev1));
#undef CK
  lap("nodes + order to the host");
  *n_nodes_out = total;
  release();
  lap("device frees");
  return RTGPU_OK;
}
