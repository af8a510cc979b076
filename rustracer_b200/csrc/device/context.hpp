// Per-GPU context of librtgpu.so (product code).  One context owns its CUDA streams (main + one side stream), the device copy of the
// flattened scene, the film accumulator and the wavefront queues.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "shapes.cuh"
#include "traverse.cuh"

struct WaveBuffers;   // render.cu

struct rtgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream2 = nullptr;       // ... and the closest-hit MIS rays beside the any-hit MIS rays
  cudaStream_t side_stream = nullptr;        // path integrator: secondary traces of a bounce, beside the next closest-hit launch
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  // second stream group: the path integrator keeps two waves in flight (render.cu), each on its own three streams
  cudaStream_t stream_b = nullptr, side_stream_b = nullptr, side_stream2_b = nullptr;
  cudaEvent_t ev_fork_b = nullptr, ev_join_b = nullptr, ev_fork2_b = nullptr, ev_join2_b = nullptr, ev_group_b = nullptr;
  int waves_in_flight = 2;
  std::string error;
  uint64_t launches = 0;
  int sm_count = 148;
  // scene
  bool has_scene = false;
  rt::DScene scene{};
  std::vector<void*> scene_allocs;
  std::vector<rtgpu_light> h_lights;          // host copies of the small tables (wave planning reads them)
  std::vector<rtgpu_material> h_materials;
  // film: 4 floats (sum r, sum g, sum b, sum weight) per cropped pixel
  float* film = nullptr; size_t film_pixels = 0; int film_w = 0, film_h = 0; float film_scale = 1.0f;
  // wavefront queues (render.cu)
  WaveBuffers* wave = nullptr;
  // spatial light distribution (render.cu)
  void* lightgrid = nullptr;
  // scratch for the batch API
  void* scratch_rays = nullptr; void* scratch_hits = nullptr; size_t scratch_n = 0;
  // host-buffer batches (api.cu host_batch): two persistent device slots {rays, results} and their {copied-in, traced, copied-out} events
  void* batch_rays[2] = {nullptr, nullptr}; void* batch_out[2] = {nullptr, nullptr}; size_t batch_cap = 0;
  cudaEvent_t batch_ev[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  int engine_carveout = 33;   // shared-memory carve-out of the traversal engines in per cent of the SM maximum (-1: the driver's choice, 132 KB; 33 -> the 100 KB configuration, 156 KB of L1: C5 +0.8 %, profiles/r03b)
  int sort_min_rays = 32768;  // batch API: batches below this are traced in the caller's order
  int profile = 0;            // rtgpu_render: time every launch with CUDA events, per kernel class (rtgpu_stats.ms_*)
  int count_traversal = 0;    // rtgpu_render: count BVH nodes visited / primitives tested (rtgpu_stats.nodes_* / prims_*)
  std::vector<cudaEvent_t> event_pool;
  int node_threshold = 12, refill_threshold = 16;   // trace_engine.cuh scheduling knobs
  int simple_traversal = 0;   // 1 = one-thread-one-ray reference walk everywhere (validation); 0 = persistent engine
  int sort_rays = 1;    // batch API: bin rays by origin cell + direction octant before traversal
  int sort_bounce_rays = 0;   // rtgpu_render (path): bin the rays of bounces >= 1 by origin cell + direction octant before tracing (profiles/r01q)
  int overlap_bounces = 2;    // rtgpu_render (path): >= 1 shadow / MIS traces of bounce b on a second stream, beside closest-hit + classify of bounce b + 1;
                              // 2: closest-hit MIS rays on a third stream beside the any-hit MIS rays
  void* comm = nullptr; int comm_rank = 0, comm_world = 1;   // ncclComm_t of rtgpu_comm_init (render.cu)
  int lightgrid_dense_mib = 2048;    // spatial light distribution: dense voxel table up to this size, sparse (rows on demand) beyond
  int lightgrid_sparse_mib = 8192;   // budget of the sparse table
  int sort_items = 1;   // rtgpu_render: counting sort of the listed-lobes queue / of the recursive integrators' items by material row
};

namespace rt {

inline int fail(rtgpu_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->error = msg;
  return code;
}
inline int check_cuda(rtgpu_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  int code = (e == cudaErrorMemoryAllocation) ? RTGPU_ERR_OOM : RTGPU_ERR_CUDA;
  return fail(ctx, code, m);
}
#define RT_CUDA(ctx, call) do { int _rc = rt::check_cuda((ctx), (call), #call); if (_rc) return _rc; } while (0)

void free_wave_buffers(rtgpu_ctx* ctx);   // render.cu
void free_lightgrid(rtgpu_ctx* ctx);      // render.cu

}  // namespace rt
