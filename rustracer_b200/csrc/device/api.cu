// librtgpu.so — context, scene upload and the batched BVH::intersect / BVH::intersect_p entry points
// (include/rtgpu.h).  Hand-written CUDA for sm_100a; compiled with -fmad=false.
#define RT_QUADRIC_INLINE 1   // shapes.cuh: quadric tests inlined (hot on sphere / disk / cylinder scenes)
#include <atomic>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include "context.hpp"
namespace rt { int configure_trace_engines(int carveout_percent); }   // tu_trace.cu
#include "trace_engine.cuh"
#include <cstring>
#include <algorithm>
#include <new>

using namespace rt;

namespace rt {

// ---------------------------------------------------------------------------------------------------------
// Ray binning: incoherent batches are reordered by (origin cell Morton code, direction octant) so that the
// 32 rays of a warp start in the same part of the tree.  Results are written back through the permutation,
// so the caller sees the original order; per-ray results do not depend on the order.
constexpr int kCellBits = 5;                       // 32^3 origin cells
constexpr int kSortBins = 1 << (3 * kCellBits + 3);

RT_DEV uint32_t spread3(uint32_t v) {              // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
RT_DEV uint32_t ray_sort_key(const float4 a, const float4 b, const float* lo, const float* inv_ext) {
  const float cells = (float)(1 << kCellBits);
  int cx = min(max(__float2int_rd((a.x - lo[0]) * inv_ext[0] * cells), 0), (1 << kCellBits) - 1);
  int cy = min(max(__float2int_rd((a.y - lo[1]) * inv_ext[1] * cells), 0), (1 << kCellBits) - 1);
  int cz = min(max(__float2int_rd((a.z - lo[2]) * inv_ext[2] * cells), 0), (1 << kCellBits) - 1);
  uint32_t morton = spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
  uint32_t oct = (b.x < 0.0f ? 1u : 0u) | (b.y < 0.0f ? 2u : 0u) | (b.z < 0.0f ? 4u : 0u);
  return (morton << 3) | oct;
}

struct SortParams { float lo[3], inv_ext[3]; };

__global__ void __launch_bounds__(256) k_sort_hist(const float4* __restrict__ rays, uint32_t n, SortParams sp, uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = rays[2 * (size_t)i], b = rays[2 * (size_t)i + 1];
  uint32_t k = ray_sort_key(a, b, sp.lo, sp.inv_ext);
  keys[i] = k;
  atomicAdd(&hist[k], 1u);
}
// exclusive scan of the kSortBins counters in three small launches (one block of 1024 threads walking all 262,144 bins took 431 us,
// 20x the traversal of a 20 k-ray batch): per-segment sums, a scan of the kSortSegments sums by one block, then every segment's own scan
constexpr int kSortSegment = 1024;                                   // bins per block
constexpr int kSortSegments = kSortBins / kSortSegment;              // 256
__global__ void __launch_bounds__(256) k_sort_scan_sums(const uint32_t* __restrict__ hist, uint32_t* __restrict__ sums) {
  __shared__ uint32_t warp_sum[8];
  const uint32_t base = blockIdx.x * kSortSegment;
  uint32_t v = 0;
  for (int i = threadIdx.x; i < kSortSegment; i += 256) v += hist[base + i];
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; w++) t += warp_sum[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(kSortSegments) k_sort_scan_top(uint32_t* __restrict__ sums) {   // exclusive scan of kSortSegments values, one block
  __shared__ uint32_t tmp[kSortSegments];
  const uint32_t own = sums[threadIdx.x];
  tmp[threadIdx.x] = own;
  __syncthreads();
  for (int off = 1; off < kSortSegments; off <<= 1) {
    const uint32_t v = (int)threadIdx.x >= off ? tmp[threadIdx.x - off] : 0;
    __syncthreads();
    tmp[threadIdx.x] += v;
    __syncthreads();
  }
  sums[threadIdx.x] = tmp[threadIdx.x] - own;
}
__global__ void __launch_bounds__(256) k_sort_scan_segments(uint32_t* __restrict__ hist, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t warp_sum[8];
  const uint32_t base = blockIdx.x * kSortSegment + threadIdx.x * 4;   // 4 consecutive bins per thread
  const uint4 c = *(const uint4*)&hist[base];
  const uint32_t own = c.x + c.y + c.z + c.w;
  uint32_t x = own;
  for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((int)(threadIdx.x & 31) >= off) x += y; }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t before = sums[blockIdx.x];
  for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_sum[w];
  const uint32_t run = before + x - own;
  *(uint4*)&hist[base] = make_uint4(run, run + c.x, run + c.x + c.y, run + c.x + c.y + c.z);
}
__global__ void __launch_bounds__(256) k_sort_scatter(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ offsets, uint32_t* __restrict__ perm) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t pos = atomicAdd(&offsets[keys[i]], 1u);
  perm[pos] = i;
}

// ---------------------------------------------------------------------------------------------------------
// Batched traversal kernels.  One thread per ray.  PERM: gather rays / scatter results through `perm`.
template <bool STATS>
__global__ void __launch_bounds__(128) k_closest_batch(DScene sc, const float4* __restrict__ rays, const uint32_t* __restrict__ perm, uint32_t n,
                                                        HitRec* __restrict__ hits, int to_prim_number, uint2* __restrict__ stats) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = perm ? perm[i] : i;
  const float4 a = rays[2 * (size_t)r], b = rays[2 * (size_t)r + 1];
  Ray ray = make_ray(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w);
  HitRec h; TravStats st; st.nodes = 0; st.prims = 0; uint32_t inst;
  bvh_traverse<false, STATS>(sc, ray, h, &st, &inst);
  // slot -> prim_number (bvh/mod.rs:92); a hit inside an object instance reports the TransformedPrimitive's number
  if (to_prim_number) h.slot = h.slot == kMiss ? kMiss : (inst != kNoInst ? sc.instances[inst].prim_number : sc.info[h.slot].x);
  hits[r] = h;
  if (STATS) stats[r] = make_uint2(st.nodes, st.prims);
}
template <bool STATS>
__global__ void __launch_bounds__(128) k_anyhit_batch(DScene sc, const float4* __restrict__ rays, const uint32_t* __restrict__ perm, uint32_t n,
                                                       uint8_t* __restrict__ occluded, uint2* __restrict__ stats) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = perm ? perm[i] : i;
  const float4 a = rays[2 * (size_t)r], b = rays[2 * (size_t)r + 1];
  Ray ray = make_ray(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w);
  HitRec h; TravStats st; st.nodes = 0; st.prims = 0;
  bool occ = bvh_traverse<true, STATS>(sc, ray, h, &st);
  occluded[r] = occ ? 1 : 0;
  if (STATS) stats[r] = make_uint2(st.nodes, st.prims);
}

// Engine variants (trace_engine.cuh): persistent warps, queue = the batch itself (optionally through `perm`).
struct BatchClosestPolicy {
  const float4* rays; const uint32_t* perm; HitRec* hits; const uint4* info; const rtgpu_instance* instances; uint32_t r;
  RT_DEV void load(uint32_t idx, Ray& ray) {
    r = perm ? perm[idx] : idx;
    const float4 a = rays[2 * (size_t)r], b = rays[2 * (size_t)r + 1];
    ray = make_ray(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w);
  }
  RT_DEV void commit(uint32_t, const HitRec& h, float t_hit, uint32_t inst, uint32_t) {
    HitRec o = h;
    o.t = t_hit;                                                      // the engine's record carries b0 in .t (trace_engine.cuh)
    o.slot = h.slot == kMiss ? kMiss : (inst != kNoInst ? instances[inst].prim_number : info[h.slot].x);   // slot -> prim_number (bvh/mod.rs:92)
    hits[r] = o;
  }
};
struct BatchAnyPolicy {
  const float4* rays; const uint32_t* perm; uint8_t* occluded; uint32_t r;
  RT_DEV void load(uint32_t idx, Ray& ray) {
    r = perm ? perm[idx] : idx;
    const float4 a = rays[2 * (size_t)r], b = rays[2 * (size_t)r + 1];
    ray = make_ray(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w);
  }
  RT_DEV void commit(uint32_t, const HitRec& h, float, uint32_t, uint32_t) { occluded[r] = h.slot != kMiss ? 1 : 0; }
};
template <bool INST>
__global__ void __launch_bounds__(128, RT_ENGINE_MIN_BLOCKS) k_closest_batch_engine(DScene sc, const float4* __restrict__ rays, const uint32_t* __restrict__ perm, uint32_t n,
                                                               HitRec* __restrict__ hits, uint32_t* cursor) {
  BatchClosestPolicy pol; pol.rays = rays; pol.perm = perm; pol.hits = hits; pol.info = sc.info; pol.instances = sc.instances; pol.r = 0;
  trace_engine<false, INST>(sc, cursor, n, pol);
}
template <bool INST>
__global__ void __launch_bounds__(128, RT_ENGINE_MIN_BLOCKS) k_anyhit_batch_engine(DScene sc, const float4* __restrict__ rays, const uint32_t* __restrict__ perm, uint32_t n,
                                                              uint8_t* __restrict__ occluded, uint32_t* cursor) {
  BatchAnyPolicy pol; pol.rays = rays; pol.perm = perm; pol.occluded = occluded; pol.r = 0;
  trace_engine<true, INST>(sc, cursor, n, pol);
}

static int ensure_sort_scratch(rtgpu_ctx* ctx, size_t n, uint32_t** keys, uint32_t** perm, uint32_t** hist) {
  // layout in one allocation: cursor[64] | hist[kSortBins] | segment sums[kSortSegments] | keys[n] | perm[n]
  size_t need = 256 + (size_t)kSortBins * 4 + (size_t)kSortSegments * 4 + n * 8;
  if (ctx->scratch_n < need) {
    if (ctx->scratch_rays) cudaFree(ctx->scratch_rays);
    ctx->scratch_rays = nullptr; ctx->scratch_n = 0;
    RT_CUDA(ctx, cudaMalloc(&ctx->scratch_rays, need));
    ctx->scratch_n = need;
  }
  *hist = (uint32_t*)ctx->scratch_rays + 64; *keys = *hist + kSortBins + kSortSegments; *perm = *keys + n;
  return 0;
}

static int build_perm(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint32_t** perm_out, uint32_t** cursor_out) {
  *perm_out = nullptr;
  uint32_t *keys, *perm, *hist;
  int rc = ensure_sort_scratch(ctx, n, &keys, &perm, &hist); if (rc) return rc;
  *cursor_out = (uint32_t*)ctx->scratch_rays;
  RT_CUDA(ctx, cudaMemsetAsync(*cursor_out, 0, 256, ctx->stream));
  // below the break-even the binning costs more than the incoherent walk it avoids (profiles/r02m_small_batches.log)
  if (!ctx->sort_rays || n < (size_t)ctx->sort_min_rays) return 0;
  SortParams sp;
  for (int k = 0; k < 3; k++) {
    float ext = ctx->scene.world_hi[k] - ctx->scene.world_lo[k];
    // rays may start outside the world bounds: bin over the bounds grown by 10 %
    sp.lo[k] = ctx->scene.world_lo[k] - 0.1f * ext;
    sp.inv_ext[k] = ext > 0.0f ? 1.0f / (1.2f * ext) : 0.0f;
  }
  RT_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)kSortBins * 4, ctx->stream));
  unsigned blocks = (unsigned)((n + 255) / 256);
  k_sort_hist<<<blocks, 256, 0, ctx->stream>>>((const float4*)d_rays, (uint32_t)n, sp, keys, hist);
  uint32_t* sums = hist + kSortBins;
  k_sort_scan_sums<<<kSortSegments, 256, 0, ctx->stream>>>(hist, sums);
  k_sort_scan_top<<<1, kSortSegments, 0, ctx->stream>>>(sums);
  k_sort_scan_segments<<<kSortSegments, 256, 0, ctx->stream>>>(hist, sums);
  k_sort_scatter<<<blocks, 256, 0, ctx->stream>>>(keys, (uint32_t)n, hist, perm);
  ctx->launches += 5;
  RT_CUDA(ctx, cudaGetLastError());
  *perm_out = perm;
  return 0;
}

}  // namespace rt

// Static split of [0, n) over the host threads for the per-node / per-slot staging loops of rtgpu_upload_scene.
template <class F> static void parallel_for(size_t n, F body) {
  int threads = (int)std::thread::hardware_concurrency();
  if (threads <= 1 || n < ((size_t)1 << 16)) { body((size_t)0, n); return; }
  const size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
  std::vector<std::thread> pool;
  for (size_t b = 0; b < n; b += chunk) pool.emplace_back([=, &body]() { body(b, std::min(n, b + chunk)); });
  for (auto& t : pool) t.join();
}

template <class T> static int upload(rtgpu_ctx* ctx, const T* host, size_t count, const T** dev) {
  *dev = nullptr;
  if (count == 0 || host == nullptr) return 0;
  void* p = nullptr;
  RT_CUDA(ctx, cudaMalloc(&p, count * sizeof(T)));
  ctx->scene_allocs.push_back(p);
  RT_CUDA(ctx, cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dev = (const T*)p;
  return 0;
}

static void free_scene(rtgpu_ctx* ctx) {
  for (void* p : ctx->scene_allocs) cudaFree(p);
  ctx->scene_allocs.clear();
  ctx->has_scene = false;
  ctx->h_lights.clear(); ctx->h_materials.clear();
  std::memset(&ctx->scene, 0, sizeof(ctx->scene));
  rt::free_lightgrid(ctx);
}

static void apply_engine_carveout(int pct) {
  if (pct < 0) return;
  rt::configure_trace_engines(pct);
  cudaFuncSetAttribute(k_closest_batch_engine<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_closest_batch_engine<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_anyhit_batch_engine<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_anyhit_batch_engine<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

extern "C" {

int rtgpu_create(int device, rtgpu_ctx** out) {
  if (!out) return RTGPU_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return RTGPU_ERR_CUDA;   // no CPU fallback exists
  if (device < 0 || device >= count) return RTGPU_ERR_ARG;
  rtgpu_ctx* ctx = new (std::nothrow) rtgpu_ctx();
  if (!ctx) return RTGPU_ERR_OOM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side_stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream_b, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->side_stream_b, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side_stream2_b, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork_b, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_join_b, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork2_b, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_join2_b, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_group_b, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    delete ctx; return RTGPU_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // L1 / shared-memory split of the traversal engines (tu_trace.cu configure_trace_engines); RT_ENGINE_CARVEOUT overrides for sweeps (-1: driver's choice)
  if (const char* e = std::getenv("RT_ENGINE_CARVEOUT")) ctx->engine_carveout = std::atoi(e);
  apply_engine_carveout(ctx->engine_carveout);
  // traversal stacks live in local memory; prefer L1 over shared memory for the traversal kernels
  *out = ctx;
  return RTGPU_OK;
}

int rtgpu_destroy(rtgpu_ctx* ctx) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->side_stream);
  cudaStreamSynchronize(ctx->side_stream2);
  cudaStreamSynchronize(ctx->stream_b); cudaStreamSynchronize(ctx->side_stream_b); cudaStreamSynchronize(ctx->side_stream2_b);
  rtgpu_comm_destroy(ctx);
  free_scene(ctx);
  rt::free_wave_buffers(ctx);
  if (ctx->film) cudaFree(ctx->film);
  if (ctx->scratch_rays) cudaFree(ctx->scratch_rays);
  if (ctx->scratch_hits) cudaFree(ctx->scratch_hits);
  for (int k = 0; k < 2; k++) { if (ctx->batch_rays[k]) cudaFree(ctx->batch_rays[k]); if (ctx->batch_out[k]) cudaFree(ctx->batch_out[k]); for (int e = 0; e < 3; e++) if (ctx->batch_ev[k][e]) cudaEventDestroy(ctx->batch_ev[k][e]); }
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); cudaEventDestroy(ctx->ev_fork2); cudaEventDestroy(ctx->ev_join2);
  cudaStreamDestroy(ctx->side_stream); cudaStreamDestroy(ctx->side_stream2);
  cudaEventDestroy(ctx->ev_fork_b); cudaEventDestroy(ctx->ev_join_b); cudaEventDestroy(ctx->ev_fork2_b); cudaEventDestroy(ctx->ev_join2_b); cudaEventDestroy(ctx->ev_group_b);
  cudaStreamDestroy(ctx->stream_b); cudaStreamDestroy(ctx->side_stream_b); cudaStreamDestroy(ctx->side_stream2_b);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return RTGPU_OK;
}

const char* rtgpu_last_error(rtgpu_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
uint64_t rtgpu_launch_count(rtgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rtgpu_set_option(rtgpu_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return RTGPU_ERR_ARG;
  if (std::strcmp(name, "sort_rays") == 0) { ctx->sort_rays = value; return RTGPU_OK; }
  if (std::strcmp(name, "sort_min_rays") == 0) { if (value < 0) return fail(ctx, RTGPU_ERR_ARG, "sort_min_rays must be >= 0"); ctx->sort_min_rays = value; return RTGPU_OK; }
  if (std::strcmp(name, "lightgrid_dense_mib") == 0 || std::strcmp(name, "lightgrid_sparse_mib") == 0) {
    if (value < 0 || value > (1 << 20)) return fail(ctx, RTGPU_ERR_ARG, "light grid budgets are in MiB, 0 .. 2^20");
    (name[10] == 'd' ? ctx->lightgrid_dense_mib : ctx->lightgrid_sparse_mib) = value;
    rt::free_lightgrid(ctx);                                           // rebuilt in the new mode by the next render
    return RTGPU_OK;
  }
  if (std::strcmp(name, "waves_in_flight") == 0) { if (value < 1 || value > 2) return fail(ctx, RTGPU_ERR_ARG, "waves_in_flight must be 1 or 2"); ctx->waves_in_flight = value; return RTGPU_OK; }
  if (std::strcmp(name, "sort_items") == 0) { ctx->sort_items = value; return RTGPU_OK; }
  if (std::strcmp(name, "overlap_bounces") == 0) { ctx->overlap_bounces = value; return RTGPU_OK; }
  if (std::strcmp(name, "sort_bounce_rays") == 0) { ctx->sort_bounce_rays = value; return RTGPU_OK; }
  if (std::strcmp(name, "node_threshold") == 0) {
    if (value < 0 || value > 32) return fail(ctx, RTGPU_ERR_ARG, "node_threshold must be in [0, 32]");
    ctx->node_threshold = value; ctx->scene.tune_node_threshold = value; return RTGPU_OK;
  }
  if (std::strcmp(name, "engine_carveout") == 0) {            // per cent of the SM's shared memory the traversal engines ask for (the rest is L1)
    if (value < -1 || value > 100) return fail(ctx, RTGPU_ERR_ARG, "engine_carveout must be in [-1, 100]");
    ctx->engine_carveout = value; cudaSetDevice(ctx->device); apply_engine_carveout(value); return RTGPU_OK;
  }
  if (std::strcmp(name, "refill_threshold") == 0) {           // <= 0 would enter the refill branch with no idle lane and spin
    if (value < 1 || value > 32) return fail(ctx, RTGPU_ERR_ARG, "refill_threshold must be in [1, 32]");
    ctx->refill_threshold = value; ctx->scene.tune_refill_threshold = value; return RTGPU_OK;
  }
  if (std::strcmp(name, "simple_traversal") == 0) { ctx->simple_traversal = value; return RTGPU_OK; }
  if (std::strcmp(name, "profile") == 0) { ctx->profile = value; return RTGPU_OK; }
  if (std::strcmp(name, "count_traversal") == 0) { ctx->count_traversal = value; return RTGPU_OK; }
  return fail(ctx, RTGPU_ERR_ARG, std::string("unknown option ") + name);
}

int rtgpu_upload_scene(rtgpu_ctx* ctx, const rtgpu_scene_desc* s) {
  if (!ctx || !s) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  free_scene(ctx);
  if (s->n_nodes > 0 && (!s->node_lo || !s->node_hi)) return fail(ctx, RTGPU_ERR_ARG, "node arrays missing");
  if (s->n_prims > 0 && (!s->prim_geom || !s->prim_info)) return fail(ctx, RTGPU_ERR_ARG, "primitive arrays missing");
  DScene& d = ctx->scene;
  // RT_UPLOAD_TIMING=1: wall time of each staging step on stderr (tools/upload_probe.py)
  const bool timing = std::getenv("RT_UPLOAD_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[upload] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  // interleave node_lo / node_hi into one 32-byte record per node (one DRAM sector per node visit)
  {
    // staging arrays are left uninitialised (new float[n]) and first touched by the threads that fill them: value-initialising 0.6 GB vectors on one
    // thread was half of the upload time of a 10 M-triangle scene (profiles/r02t_upload_probe_c4_before.log)
    std::unique_ptr<float[]> inter(new float[(size_t)s->n_nodes * 8]);
    parallel_for(s->n_nodes, [&](size_t i0, size_t i1) {
      for (size_t i = i0; i < i1; i++) {
        std::memcpy(&inter[i * 8], &s->node_lo[i * 4], 16);
        std::memcpy(&inter[i * 8 + 4], &s->node_hi[i * 4], 16);
      }
    });
    const float* p = nullptr;
    int rc = upload(ctx, inter.get(), (size_t)s->n_nodes * 8, &p); if (rc) return rc;
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `inter` dies at scope end
    d.nodes = (const float4*)p;
  }
  lap("nodes interleave + copy");
  int rc;
  const float* pf = nullptr; const uint32_t* pu = nullptr;
  // wide nodes for the traversal engine (trace_engine.cuh) + "last primitive of the leaf" marks in the geometry copy
  {
    const size_t nn = s->n_nodes;
    auto bits = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    auto fbits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
#if RT_ENGINE_WIDE4
    // Collapsed nodes (trace_engine.cuh): one record per interior node at EVEN depth below the root of its tree — the scene's tree and
    // every object definition's — numbered in pre-order; it holds the children's children (a leaf child stands for itself).
    auto is_leaf = [&](size_t i) { return (bits(s->node_hi[i * 4 + 3]) >> 2) != 0; };
    // Pre-order of the collapsed tree == array order of the nodes that get a record (the array is the binary tree's pre-order), so the numbering is
    // a mark pass (which interior nodes sit at even depth: disjoint subtrees walked by all host threads) and a prefix count over the array.
    // One thread walking a 10 M-triangle tree took 120 ms (profiles/r02v_upload_probe_c4.log).
    constexpr uint32_t kUnmarked = 0xffffffffu, kMarked = 0xfffffffeu;
    std::unique_ptr<uint32_t[]> interior_index(new uint32_t[nn ? nn : 1]);
    parallel_for(nn, [&](size_t i0, size_t i1) { for (size_t i = i0; i < i1; i++) interior_index[i] = kUnmarked; });
    uint32_t n_interior = 0;
    {
      std::vector<size_t> roots;
      if (nn > 0) roots.push_back(0);
      for (uint32_t k = 0; k < s->n_instances && s->instances; k++) if (s->instances[k].root_node != 0xffffffffu && s->instances[k].root_node < nn) roots.push_back(s->instances[k].root_node);
      std::sort(roots.begin(), roots.end());
      roots.erase(std::unique(roots.begin(), roots.end()), roots.end());
      std::atomic<int> bad_child{0};
      // marks node i and lists its interior grandchildren (the roots of the next collapsed level)
      auto visit = [&](size_t i, std::vector<size_t>& out) {
        if (interior_index[i] != kUnmarked) return;
        interior_index[i] = kMarked;
        const size_t kids[2] = {i + 1, (size_t)bits(s->node_lo[i * 4 + 3])};
        for (size_t c : kids) {
          if (c >= nn) { bad_child = 1; return; }
          if (is_leaf(c)) continue;
          const size_t g[2] = {c + 1, (size_t)bits(s->node_lo[c * 4 + 3])};
          for (size_t q : g) { if (q >= nn) { bad_child = 1; return; } if (!is_leaf(q)) out.push_back(q); }
        }
      };
      const size_t n_threads = std::max<size_t>(1, std::thread::hardware_concurrency());
      std::vector<size_t> frontier, next;
      for (size_t r : roots) if (!is_leaf(r)) frontier.push_back(r);
      while (!frontier.empty() && frontier.size() < 64 * n_threads && !bad_child) {   // the top levels, until there is a subtree per thread and then some
        next.clear();
        for (size_t i : frontier) visit(i, next);
        frontier.swap(next);
      }
      std::atomic<size_t> cursor{0};
      auto worker = [&]() {
        std::vector<size_t> todo, kids;
        while (true) {
          const size_t t = cursor.fetch_add(1);
          if (t >= frontier.size() || bad_child) return;
          todo.assign(1, frontier[t]);
          while (!todo.empty()) { const size_t i = todo.back(); todo.pop_back(); kids.clear(); visit(i, kids); todo.insert(todo.end(), kids.begin(), kids.end()); }
        }
      };
      {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < n_threads && t < frontier.size(); t++) pool.emplace_back(worker);
        worker();
        for (auto& t : pool) t.join();
      }
      if (bad_child) return fail(ctx, RTGPU_ERR_ARG, "interior node child index outside the node arrays");
      // prefix count of the marked nodes in array order
      const size_t chunk = (nn + n_threads - 1) / n_threads;
      std::vector<uint32_t> counts(n_threads + 1, 0);
      auto span = [&](size_t t, size_t& a, size_t& b) { a = std::min(nn, t * chunk); b = std::min(nn, a + chunk); };
      auto for_chunks = [&](auto body) {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < n_threads; t++) pool.emplace_back(body, t);
        body((size_t)0);
        for (auto& t : pool) t.join();
      };
      for_chunks([&](size_t t) { size_t a, b; span(t, a, b); uint32_t c = 0; for (size_t i = a; i < b; i++) c += interior_index[i] == kMarked; counts[t + 1] = c; });
      for (size_t t = 0; t < n_threads; t++) counts[t + 1] += counts[t];
      n_interior = counts[n_threads];
      for_chunks([&](size_t t) { size_t a, b; span(t, a, b); uint32_t c = counts[t]; for (size_t i = a; i < b; i++) if (interior_index[i] == kMarked) interior_index[i] = c++; });
      d.n_top = 0;
    }
    auto ref_of = [&](size_t i) -> uint32_t {
      const uint32_t n_prims = bits(s->node_hi[i * 4 + 3]) >> 2;
      return n_prims > 0 ? (0x80000000u | bits(s->node_lo[i * 4 + 3])) : interior_index[i];
    };
    lap("collapsed numbering (DFS)");
    const size_t wide_floats = (size_t)n_interior * 32;
    std::unique_ptr<float[]> wide(new float[wide_floats]);          // every record is written in full below
    lap("wide alloc");
#else
    // compact interior numbering: the top levels of the scene's tree first, breadth-first (the engine can stage them in shared
    // memory, RT_ENGINE_TOP_NODES), then every other interior node in array (pre-order) order
    std::vector<uint32_t> interior_index(nn, 0xffffffffu);
    uint32_t n_interior = 0;
    {
      const uint32_t want_top = 255;
      std::vector<size_t> level;
      if (nn > 0 && (bits(s->node_hi[3]) >> 2) == 0) level.push_back(0);
      while (!level.empty() && n_interior + level.size() <= want_top) {
        std::vector<size_t> next;
        for (size_t i : level) {
          interior_index[i] = n_interior++;
          const size_t L = i + 1, R = bits(s->node_lo[i * 4 + 3]);
          if (L < nn && (bits(s->node_hi[L * 4 + 3]) >> 2) == 0) next.push_back(L);
          if (R < nn && (bits(s->node_hi[R * 4 + 3]) >> 2) == 0) next.push_back(R);
        }
        level.swap(next);
      }
      d.n_top = std::min<uint32_t>(n_interior, (uint32_t)RT_ENGINE_TOP_NODES);
    }
    for (size_t i = 0; i < nn; i++) if ((bits(s->node_hi[i * 4 + 3]) >> 2) == 0 && interior_index[i] == 0xffffffffu) interior_index[i] = n_interior++;
    auto ref_of = [&](size_t i) -> uint32_t {
      const uint32_t n_prims = bits(s->node_hi[i * 4 + 3]) >> 2;
      return n_prims > 0 ? (0x80000000u | bits(s->node_lo[i * 4 + 3])) : interior_index[i];
    };
    const size_t wide_floats = (size_t)n_interior * 16;
    std::unique_ptr<float[]> wide(new float[wide_floats]());
#endif
    const size_t geom_floats = (size_t)s->n_prims * 12;
    std::unique_ptr<float[]> geom(new float[geom_floats]);
    // shade-queue id of every slot's material in bits 2..4 of the second float4's w: the engine has that word in a register
    // when it records a hit, so classification needs no second look-up (hit slots are packed into 29 bits next to it)
    if (s->n_prims >= (1u << kHitSlotBits)) return fail(ctx, RTGPU_ERR_ARG, "more than 2^29 primitive slots");
    parallel_for(s->n_prims, [&](size_t s0, size_t s1) {
      std::memcpy(&geom[s0 * 12], &s->prim_geom[s0 * 12], (s1 - s0) * 12 * sizeof(float));
      for (size_t slot = s0; slot < s1; slot++) {
        const uint32_t mrow = s->prim_info[slot * 4 + 1];
        const uint32_t type = (mrow < s->n_materials && s->materials) ? s->materials[mrow].type : (uint32_t)RTGPU_MAT_NONE;
        geom[slot * 12 + 7] = fbits(bits(geom[slot * 12 + 7]) | ((uint32_t)material_queue(type) << kGeomClassShift));
      }
    });
    lap("geom copy + class bits");
    std::atomic<int> bad_nodes{0};                                    // every node writes its own wide record / its own leaf's last slot
    parallel_for(nn, [&](size_t n0, size_t n1) {
      for (size_t i = n0; i < n1; i++) {
        const uint32_t meta = bits(s->node_hi[i * 4 + 3]), n_prims = meta >> 2, off = bits(s->node_lo[i * 4 + 3]);
        if (n_prims > 0) {
          if ((size_t)off + n_prims > s->n_prims) { bad_nodes = 1; continue; }
          geom[((size_t)off + n_prims - 1) * 12 + 7] = fbits(bits(geom[((size_t)off + n_prims - 1) * 12 + 7]) | kGeomLastBit);
          continue;
        }
#if RT_ENGINE_WIDE4
        if (interior_index[i] == 0xffffffffu) continue;                // an interior node of odd depth: absorbed by its parent's record
        float* w = &wide[(size_t)interior_index[i] * 32];
        std::memset(w, 0, 32 * sizeof(float));
        const size_t kids[2] = {i + 1, (size_t)off};
        uint32_t axes = meta & 3u, refs[4];
        for (int c = 0; c < 2; c++) {
          const size_t k = kids[c];
          size_t e[2]; int ne;
          if ((bits(s->node_hi[k * 4 + 3]) >> 2) != 0) { e[0] = k; ne = 1; }
          else { e[0] = k + 1; e[1] = bits(s->node_lo[k * 4 + 3]); ne = 2; axes |= (bits(s->node_hi[k * 4 + 3]) & 3u) << (2 + 2 * c); }
          for (int j = 0; j < 2; j++) {
            const int slot = 2 * c + j;
            float* lo = w + 8 * slot; float* hi = lo + 4;
            if (j < ne) {
              for (int x = 0; x < 3; x++) { lo[x] = s->node_lo[e[j] * 4 + x]; hi[x] = s->node_hi[e[j] * 4 + x]; }
              refs[slot] = ref_of(e[j]);
            } else refs[slot] = 0xffffffffu;
          }
        }
        // w components of the eight float4: ref c0, ref c1, axes, ref c2, ref c3, -, -, -
        w[3] = fbits(refs[0]); w[7] = fbits(refs[1]); w[11] = fbits(axes); w[15] = fbits(refs[2]); w[19] = fbits(refs[3]);
#else
        const size_t L = i + 1, R = off;
        if (L >= nn || R >= nn) { bad_nodes = 2; continue; }
        float* w = &wide[(size_t)interior_index[i] * 16];
        for (int k = 0; k < 3; k++) { w[k] = s->node_lo[L * 4 + k]; w[4 + k] = s->node_hi[L * 4 + k]; w[8 + k] = s->node_lo[R * 4 + k]; w[12 + k] = s->node_hi[R * 4 + k]; }
        w[3] = fbits(ref_of(L)); w[7] = fbits(ref_of(R)); w[11] = fbits(meta & 3u); w[15] = 0.0f;
#endif
      }
    });
    lap("wide records + last bits");
    if (bad_nodes == 1) return fail(ctx, RTGPU_ERR_ARG, "leaf primitive range outside the primitive arrays");
    if (bad_nodes == 2) return fail(ctx, RTGPU_ERR_ARG, "interior node child index outside the node arrays");
    d.root_ref = nn > 0 ? ref_of(0) : 0xffffffffu;
    // object instances: the engine's reference of each definition's root (a one-primitive definition is a one-slot leaf)
    if (s->n_instances && !s->instances) return fail(ctx, RTGPU_ERR_ARG, "instance table missing");
    std::vector<rtgpu_instance> inst(s->instances, s->instances + s->n_instances);
    for (rtgpu_instance& I : inst) {
      if (I.root_node == 0xffffffffu) {
        if (I.first_slot >= s->n_prims) return fail(ctx, RTGPU_ERR_ARG, "instance slot outside the primitive arrays");
        I.root_ref = 0x80000000u | I.first_slot;
      } else {
        if (I.root_node >= nn) return fail(ctx, RTGPU_ERR_ARG, "instance root outside the node arrays");
        I.root_ref = ref_of(I.root_node);
      }
    }
    if ((rc = upload(ctx, inst.data(), inst.size(), &d.instances))) return rc;
    d.n_instances = s->n_instances;
    if ((rc = upload(ctx, wide.get(), wide_floats, &pf))) return rc; d.wide = (const float4*)pf;
    if ((rc = upload(ctx, geom.get(), geom_floats, &pf))) return rc; d.geom = (const float4*)pf;
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the staging vectors die at scope end
    lap("wide + geom copy");
  }
  lap("staging vectors freed");
  if ((rc = upload(ctx, s->prim_info, (size_t)s->n_prims * 4, &pu))) return rc; d.info = (const uint4*)pu;
  if ((rc = upload(ctx, s->tri_n, s->tri_n ? (size_t)s->n_prims * 9 : 0, &d.tri_n))) return rc;
  if ((rc = upload(ctx, s->tri_s, s->tri_s ? (size_t)s->n_prims * 9 : 0, &d.tri_s))) return rc;
  if ((rc = upload(ctx, s->tri_uv, s->tri_uv ? (size_t)s->n_prims * 6 : 0, &d.tri_uv))) return rc;
  if ((rc = upload(ctx, s->quadrics, (size_t)s->n_quadrics, &d.quadrics))) return rc;
  if ((rc = upload(ctx, s->materials, (size_t)s->n_materials, &d.materials))) return rc;
  if ((rc = upload(ctx, s->lobes, s->lobes ? (size_t)s->n_lobes : 0, &d.lobes))) return rc;
  // textured materials (texture.cuh): neutral material rows, texture rows, MIP pyramids + EWA weight table
  if (s->n_texmats && (!s->texmats || s->n_texmats != s->n_materials)) return fail(ctx, RTGPU_ERR_ARG, "texmats must mirror the material table");
  if (s->n_texmats && (!s->tex_data || s->n_tex_floats < 128)) return fail(ctx, RTGPU_ERR_ARG, "texture pool missing");
  if ((rc = upload(ctx, s->texmats, s->texmats ? (size_t)s->n_texmats : 0, &d.texmats))) return rc;
  if ((rc = upload(ctx, s->textures, s->textures ? (size_t)s->n_textures : 0, &d.textures))) return rc;
  if ((rc = upload(ctx, s->tex_data, s->tex_data ? (size_t)s->n_tex_floats : 0, &d.tex_data))) return rc;
  d.n_textures = s->n_textures;
  for (uint32_t i = 0; i < s->n_lights && s->lights; i++) {           // the env-map lookup wraps with a mask (lights.cuh env_texel)
    const rtgpu_light& l = s->lights[i];
    if (l.kind == RTGPU_LIGHT_INFINITE && (l.env_w == 0 || l.env_h == 0 || (l.env_w & (l.env_w - 1)) || (l.env_h & (l.env_h - 1))))
      return fail(ctx, RTGPU_ERR_UNSUPPORTED, "infinite light: environment map sides must be powers of two (resample first, as MIPMap::new does)");
  }
  if ((rc = upload(ctx, s->lights, (size_t)s->n_lights, &d.lights))) return rc;
  if ((rc = upload(ctx, s->env_data, (size_t)s->n_env_floats, &d.env))) return rc;
  d.n_nodes = s->n_nodes; d.n_prims = s->n_prims; d.n_quadrics = s->n_quadrics; d.n_materials = s->n_materials; d.n_lights = s->n_lights;
  for (int i = 0; i < 3; i++) { d.world_lo[i] = s->world_lo[i]; d.world_hi[i] = s->world_hi[i]; }
  d.tune_node_threshold = ctx->node_threshold; d.tune_refill_threshold = ctx->refill_threshold;
  if (s->n_lights) ctx->h_lights.assign(s->lights, s->lights + s->n_lights);
  if (s->n_materials) ctx->h_materials.assign(s->materials, s->materials + s->n_materials);
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  lap("info + tables copy");
  ctx->has_scene = true;
  return RTGPU_OK;
}

int rtgpu_malloc(rtgpu_ctx* ctx, size_t bytes, void** d_ptr) {
  if (!ctx || !d_ptr) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaMalloc(d_ptr, bytes));
  return RTGPU_OK;
}
int rtgpu_free(rtgpu_ctx* ctx, void* d_ptr) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaFree(d_ptr));
  return RTGPU_OK;
}
int rtgpu_memcpy_h2d(rtgpu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RTGPU_OK;
}
int rtgpu_memcpy_d2h(rtgpu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RTGPU_OK;
}
int rtgpu_synchronize(rtgpu_ctx* ctx) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RTGPU_OK;
}

static int run_closest(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, rtgpu_hit* d_hits, uint32_t* d_stats, float* elapsed_ms) {
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  if (n == 0) { if (elapsed_ms) *elapsed_ms = 0; return RTGPU_OK; }
  if (n > 0xfffffff0ull) return fail(ctx, RTGPU_ERR_ARG, "batch too large (max 2^32-16 rays)");
  cudaSetDevice(ctx->device);
  if (elapsed_ms) RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  uint32_t *perm = nullptr, *cursor = nullptr;
  int rc = build_perm(ctx, d_rays, n, &perm, &cursor); if (rc) return rc;
  unsigned blocks = (unsigned)((n + 127) / 128);
  const unsigned pblocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, blocks);
  // the counting variant is the plain one-thread-one-ray reference walk; the fast path is the persistent engine
  if (d_stats) k_closest_batch<true><<<blocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, (HitRec*)d_hits, 1, (uint2*)d_stats);
  else if (ctx->simple_traversal) k_closest_batch<false><<<blocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, (HitRec*)d_hits, 1, nullptr);
  else if (ctx->scene.n_instances) k_closest_batch_engine<true><<<pblocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, (HitRec*)d_hits, cursor);
  else k_closest_batch_engine<false><<<pblocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, (HitRec*)d_hits, cursor);
  ctx->launches += 1;
  RT_CUDA(ctx, cudaGetLastError());
  if (elapsed_ms) {
    RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    RT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    RT_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
  }
  return RTGPU_OK;
}
static int run_anyhit(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint8_t* d_occ, uint32_t* d_stats, float* elapsed_ms) {
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  if (n == 0) { if (elapsed_ms) *elapsed_ms = 0; return RTGPU_OK; }
  if (n > 0xfffffff0ull) return fail(ctx, RTGPU_ERR_ARG, "batch too large (max 2^32-16 rays)");
  cudaSetDevice(ctx->device);
  if (elapsed_ms) RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  uint32_t *perm = nullptr, *cursor = nullptr;
  int rc = build_perm(ctx, d_rays, n, &perm, &cursor); if (rc) return rc;
  unsigned blocks = (unsigned)((n + 127) / 128);
  const unsigned pblocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, blocks);
  if (d_stats) k_anyhit_batch<true><<<blocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, d_occ, (uint2*)d_stats);
  else if (ctx->simple_traversal) k_anyhit_batch<false><<<blocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, d_occ, nullptr);
  else if (ctx->scene.n_instances) k_anyhit_batch_engine<true><<<pblocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, d_occ, cursor);
  else k_anyhit_batch_engine<false><<<pblocks, 128, 0, ctx->stream>>>(ctx->scene, (const float4*)d_rays, perm, (uint32_t)n, d_occ, cursor);
  ctx->launches += 1;
  RT_CUDA(ctx, cudaGetLastError());
  if (elapsed_ms) {
    RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    RT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    RT_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
  }
  return RTGPU_OK;
}

int rtgpu_intersect_device(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, rtgpu_hit* d_hits, float* elapsed_ms) {
  if (!ctx || (n && (!d_rays || !d_hits))) return RTGPU_ERR_ARG;
  return run_closest(ctx, d_rays, n, d_hits, nullptr, elapsed_ms);
}
int rtgpu_occluded_device(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint8_t* d_occluded, float* elapsed_ms) {
  if (!ctx || (n && (!d_rays || !d_occluded))) return RTGPU_ERR_ARG;
  return run_anyhit(ctx, d_rays, n, d_occluded, nullptr, elapsed_ms);
}
// Same as the _device calls plus per-ray {nodes visited, primitives tested} (2 x uint32 each): the N and T of
// the roofline's algorithmic bytes.  d_stats must hold 2*n uint32.
int rtgpu_intersect_device_stats(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, rtgpu_hit* d_hits, uint32_t* d_stats) {
  if (!ctx || (n && (!d_rays || !d_hits || !d_stats))) return RTGPU_ERR_ARG;
  return run_closest(ctx, d_rays, n, d_hits, d_stats, nullptr);
}
int rtgpu_occluded_device_stats(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint8_t* d_occluded, uint32_t* d_stats) {
  if (!ctx || (n && (!d_rays || !d_occluded || !d_stats))) return RTGPU_ERR_ARG;
  return run_anyhit(ctx, d_rays, n, d_occluded, d_stats, nullptr);
}

// Host-buffer variants: H2D + kernel + D2H inside (the reference-facing plugin call of the ray-batch config).  Chunks of 4 Mi rays go
// through a three-stage pipeline on three streams — copy-in of chunk c + 1 (side stream), traversal of chunk c (main stream), copy-out of
// chunk c - 1 (second side stream) — over two persistent device slots, so a large batch moves at the PCIe rate of its slower direction
// instead of serialising copy, kernel and copy (round 1: 90 Mrays/s against 2 Grays/s on the device).  Pinned host memory
// (rtgpu_host_alloc, or registered by the caller) is copied asynchronously at link rate; pageable memory works and is staged by the driver.
static int ensure_batch_slots(rtgpu_ctx* ctx, size_t cap_rays) {
  if (ctx->batch_cap >= cap_rays) return 0;
  for (int k = 0; k < 2; k++) { if (ctx->batch_rays[k]) cudaFree(ctx->batch_rays[k]); if (ctx->batch_out[k]) cudaFree(ctx->batch_out[k]); ctx->batch_rays[k] = ctx->batch_out[k] = nullptr; }
  ctx->batch_cap = 0;
  for (int k = 0; k < 2; k++) {
    RT_CUDA(ctx, cudaMalloc(&ctx->batch_rays[k], cap_rays * sizeof(rtgpu_ray)));
    RT_CUDA(ctx, cudaMalloc(&ctx->batch_out[k], cap_rays * sizeof(rtgpu_hit)));
    for (int e = 0; e < 3; e++) if (!ctx->batch_ev[k][e]) RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->batch_ev[k][e], cudaEventDisableTiming));
  }
  ctx->batch_cap = cap_rays;
  return 0;
}
static int host_batch(rtgpu_ctx* ctx, const rtgpu_ray* rays, size_t n, rtgpu_hit* hits, uint8_t* occluded) {
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  if (n == 0) return RTGPU_OK;
  cudaSetDevice(ctx->device);
  const size_t chunk = (size_t)1 << 22;
  const size_t cap = n < chunk ? std::max<size_t>(n, 4096) : chunk;
  int rc = ensure_batch_slots(ctx, cap); if (rc) return rc;
  const size_t out_size = hits ? sizeof(rtgpu_hit) : 1;
  cudaStream_t s_in = ctx->side_stream, s_run = ctx->stream, s_out = ctx->side_stream2;
  enum { EV_IN = 0, EV_RUN = 1, EV_OUT = 2 };
  size_t c = 0;
  for (size_t first = 0; first < n; first += cap, c++) {
    const int slot = (int)(c & 1);
    const size_t m = std::min(cap, n - first);
    if (c >= 2) RT_CUDA(ctx, cudaStreamWaitEvent(s_in, ctx->batch_ev[slot][EV_RUN], 0));          // the slot's previous rays have been traced
    RT_CUDA(ctx, cudaMemcpyAsync(ctx->batch_rays[slot], rays + first, m * sizeof(rtgpu_ray), cudaMemcpyHostToDevice, s_in));
    RT_CUDA(ctx, cudaEventRecord(ctx->batch_ev[slot][EV_IN], s_in));
    RT_CUDA(ctx, cudaStreamWaitEvent(s_run, ctx->batch_ev[slot][EV_IN], 0));
    if (c >= 2) RT_CUDA(ctx, cudaStreamWaitEvent(s_run, ctx->batch_ev[slot][EV_OUT], 0));         // the slot's previous results have left
    if (hits) rc = run_closest(ctx, (const rtgpu_ray*)ctx->batch_rays[slot], m, (rtgpu_hit*)ctx->batch_out[slot], nullptr, nullptr);
    else rc = run_anyhit(ctx, (const rtgpu_ray*)ctx->batch_rays[slot], m, (uint8_t*)ctx->batch_out[slot], nullptr, nullptr);
    if (rc) break;
    RT_CUDA(ctx, cudaEventRecord(ctx->batch_ev[slot][EV_RUN], s_run));
    RT_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->batch_ev[slot][EV_RUN], 0));
    void* dst = hits ? (void*)(hits + first) : (void*)(occluded + first);
    RT_CUDA(ctx, cudaMemcpyAsync(dst, ctx->batch_out[slot], m * out_size, cudaMemcpyDeviceToHost, s_out));
    RT_CUDA(ctx, cudaEventRecord(ctx->batch_ev[slot][EV_OUT], s_out));
  }
  cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_run), e3 = cudaStreamSynchronize(s_out);
  if (rc) return rc;
  if (e1 != cudaSuccess) return check_cuda(ctx, e1, "copy-in stream");
  if (e2 != cudaSuccess) return check_cuda(ctx, e2, "traversal stream");
  if (e3 != cudaSuccess) return check_cuda(ctx, e3, "copy-out stream");
  return RTGPU_OK;
}
// Pinned host memory for the host-buffer entry points (and the film read-back): page-locked, so copies run asynchronously at link rate.
int rtgpu_host_alloc(rtgpu_ctx* ctx, size_t bytes, void** h_ptr) {
  if (!ctx || !h_ptr) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaHostAlloc(h_ptr, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
  return RTGPU_OK;
}
int rtgpu_host_free(rtgpu_ctx* ctx, void* h_ptr) {
  if (!ctx) return RTGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  RT_CUDA(ctx, cudaFreeHost(h_ptr));
  return RTGPU_OK;
}
int rtgpu_intersect(rtgpu_ctx* ctx, const rtgpu_ray* rays, size_t n, rtgpu_hit* hits) {
  if (!ctx || (n && (!rays || !hits))) return RTGPU_ERR_ARG;
  return host_batch(ctx, rays, n, hits, nullptr);
}
int rtgpu_occluded(rtgpu_ctx* ctx, const rtgpu_ray* rays, size_t n, uint8_t* occluded) {
  if (!ctx || (n && (!rays || !occluded))) return RTGPU_ERR_ARG;
  return host_batch(ctx, rays, n, nullptr, occluded);
}

}  // extern "C"
