// Wavefront state: ray / path-state queues in coalesced float4 SoA buffers, device-side queue counters,
// warp-aggregated queue appends (product code).
#pragma once
#include "lights.cuh"
#include "sampler.cuh"
#include "traverse.cuh"

namespace rt {

// (the material queues Q_* and material_queue() live in shapes.cuh: the scene upload tags every geometry slot with its queue)
// device counters (uint32)
enum {
  C_LIVE0 = 0, C_LIVE1, C_MATQ0, C_SHADOW = C_MATQ0 + Q_COUNT, C_MIS, C_CUR_CLOSEST, C_CUR_ANY, C_CUR_MIS, C_OVERFLOW, C_MIS_ANY, C_MIS_SKIPPED, C_CUR_MISANY, C_COUNT = 32
};
// device statistics (uint64): the reference's counters (scene.rs:9-16, renderer.rs:17)
enum { S_CAMERA = 0, S_REGULAR, S_SHADOW, S_NODES_CLOSEST, S_PRIMS_CLOSEST, S_NODES_ANY, S_PRIMS_ANY, S_OVERFLOW, S_CLOSEST_RAYS, S_ANY_RAYS, S_VERTICES, S_COUNT = 11 };

// Spatial / uniform light distribution tables (lightdistrib.rs).  Per voxel: func[n], cdf[n+1], func_int.
struct LightGrid {
  const float* table;        // voxel table (dense: row = voxel; sparse: row = slots[voxel]), or the single uniform distribution when nv = {0,0,0}
  int nv[3];
  int n_lights;
  // Sparse mode (many lights: the dense table would not fit).  Like the reference's hash table (lightdistrib.rs:201-296) only the voxels
  // that path vertices actually fall into get a distribution: k_lightgrid_mark claims a row for every new voxel among a bounce's hit
  // points, k_lightgrid_contrib / _build fill the new rows, all before the bounce is shaded.  Rows live as long as the scene.
  int* slots;                // per voxel: row, -1 = not requested yet, -2 = being claimed; null in dense mode
  uint32_t* new_voxels;      // the rows claimed by the current bounce: {voxel, row} pairs
  uint32_t* grid_counters;   // [0] rows in use, [1] entries of new_voxels, [2] overflow flag
  uint32_t cap_rows;
};
enum { G_ROWS = 0, G_NEW = 1, G_OVERFLOW = 2 };

struct WaveView {            // device pointers, passed to kernels by value
  // items (path: item slot == sample slot)
  float4 *ray_o, *ray_d;     // {o.xyz, t_max}, {d.xyz, -}
  HitRec* hit;
  uint8_t* hit_class;        // shade queue of each hit (Q_*), written by the traversal engine from the class bits of the hit slot
  uint32_t* hit_inst;        // instance row of each hit (kNoInst at the top level); null when the scene has no object instances
  float4* beta;              // rgb throughput, w = eta_scale (path)
  uint4* pstate;             // x = sample slot, y = bounces (path) | node id (recursive), z = flags | depth, w = d1 | d2 << 16
  // second item buffer (recursive integrators ping-pong between levels)
  float4 *ray_o2, *ray_d2; float4* beta2; uint4* pstate2;
  float4 *rdiff, *rdiff2;    // ray differentials of the recursive integrators' items, 3 float4 each {rx_o, ry_o, rx_d, ry_d}; textured scenes only
  // camera samples
  float4* L;                 // rgb radiance accumulator, w = 1 if the sample exists
  float2* pfilm;
  uint2* sinfo;              // pixel hash, sample index
  // shadow (any-hit) queue and MIS (closest-hit) queue
  float4 *sh_o, *sh_d, *sh_c;          // ray; d.w = bits(sample slot); c = rgb contribution if unoccluded
  float4 *mi_o, *mi_d, *mi_c;          // ray; d.w = bits(sample slot); c = rgb weight, w = bits(light row)
  float4 *ma_o, *ma_d, *ma_c;          // MIS rays towards infinite lights, traced any-hit: c = rgb contribution if the ray escapes
  uint32_t* list[2];
  uint32_t* matq[Q_COUNT];
  uint32_t *matsort_hist, *matsort_out;   // material sort of the listed-lobes queue (path) / of a level's items (recursive integrators)
  uint32_t matsort_bins;
  rtgpu_lobe* tex_lobes;                  // path integrator, textured scenes: 8 lobe rows per item, written by k_eval_textured
  float4* tex_frame;                      //   and 2 float4 per item: {bumped shading normal, Bsdf::eta}, {bumped shading dpdu, bits(lobe count)}
  uint32_t *raysort_keys, *raysort_hist, *raysort_out;   // option "sort_bounce_rays": binning of a bounce's rays before the closest-hit launch
  const uint32_t* item_order;             // recursive integrators: processing order of the level's items (or null)
  uint32_t* counters;
  unsigned long long* stats;
  uint32_t cap_items, cap_samples, cap_shadow, cap_mis;
};

struct RenderParams {
  DScene sc;
  WaveView w;
  LightGrid grid;
  float r2c[16], c2w[16]; float lens_radius, focal_distance;
  int sample_bounds[4], pixel_bounds[4];
  SamplerCfg scfg; unsigned long long seed;
  int integrator, max_depth, direct_strategy, ao_samples; float rr_threshold;
  const uint32_t* n_light_samples;     // DirectLighting "all": per light, rounded to a power of two
  // wave decomposition: my tiles [tile_first, tile_first + n_tiles) x samples [sample_first, sample_first + n_samples)
  int tiles_x, tile_rank, tile_world, tile_first, n_tiles, sample_first, n_samples;
  const int32_t* explicit_pixels;      // li_samples mode: {x, y, sample} triples, else null
  int hit_t_is_b0;                     // 1: w.hit[].t holds the first barycentric of the hit (records written by the traversal engine), 0: the distance (reference walker)
  uint32_t n_items;
};

// k_film_add parameters (film.rs): cropped bounds, filter radius and the 16x16 filter table
struct FilmParams {
  float4* film; int crop[4]; float rx, ry, irx, iry; float max_lum; float ao_div;
  float table[256];
};

RT_DEV uint32_t lane_id() { return threadIdx.x & 31u; }
// Append to a device queue: one atomicAdd per warp (ballot + popc), lanes get consecutive entries.
RT_DEV uint32_t warp_append(uint32_t* counter, bool pred) {
  const unsigned active = __activemask();
  const unsigned mask = __ballot_sync(active, pred);
  if (!pred) return 0xffffffffu;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + (uint32_t)__popc(mask & ((1u << lane_id()) - 1u));
}
// Append to a device queue: one atomicAdd per BLOCK.  Every thread of the block must call it (it synchronises); blockDim.x <= 1024.
RT_DEV uint32_t block_append(uint32_t* counter, bool pred) {
  __shared__ uint32_t s_warp_count[32];
  __shared__ uint32_t s_block_base;
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31u) >> 5;
  if (lane == 0) s_warp_count[warp] = (uint32_t)__popc(mask);
  __syncthreads();
  if (warp == 0) {
    const uint32_t c = lane < n_warps ? s_warp_count[lane] : 0u;
    uint32_t x = c;                                                  // inclusive scan of the warp counts
    for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((int)lane >= off) x += y; }
    if (lane == 31) s_block_base = x ? atomicAdd(counter, x) : 0u;
    if (lane < n_warps) s_warp_count[lane] = x - c;                  // exclusive offset of each warp
  }
  __syncthreads();
  const uint32_t pos = s_block_base + s_warp_count[warp] + (uint32_t)__popc(mask & ((1u << lane) - 1u));
  __syncthreads();                                                   // the shared words are reused by the next call
  return pred ? pos : 0xffffffffu;
}
// Next packet of 32 queue entries for a persistent warp (atomic cursor).
RT_DEV uint32_t warp_fetch(uint32_t* cursor) {
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(cursor, 32u);
  return __shfl_sync(0xffffffffu, base, 0);
}

// Wave buffers are written by one kernel and read once by the next, gigabytes later: with RT_STREAM_HINTS their loads and stores carry the
// "streaming" cache operator (ld.global.cs / st.global.cs: evict-first in L1 and L2), so that they do not push the BVH out of the L2.
#ifndef RT_STREAM_HINTS
#define RT_STREAM_HINTS 1            // profiles/r03g: C5 855 -> 871 M samples/s, C3 798 -> 807 M (the gain is in the shade kernels: 122.6 -> 118.3 ms per C5 step)
#endif
#if RT_STREAM_HINTS
template <class T> RT_DEV T ld_stream(const T* p) { return __ldcs(p); }
template <class T> RT_DEV void st_stream(T* p, T v) { __stcs(p, v); }
#else
template <class T> RT_DEV T ld_stream(const T* p) { return *p; }
template <class T> RT_DEV void st_stream(T* p, T v) { *p = v; }
#endif
RT_DEV void store_ray(float4* o, float4* d, uint32_t i, const Ray& r, uint32_t tag) {
  st_stream(&o[i], make_float4(r.o.x, r.o.y, r.o.z, r.t_max));
  st_stream(&d[i], make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(tag)));
}
RT_DEV Ray load_ray(const float4* o, const float4* d, uint32_t i, uint32_t* tag) {
  const float4 a = ld_stream(&o[i]), b = ld_stream(&d[i]);
  if (tag) *tag = __float_as_uint(b.w);
  return make_ray(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w);
}

}  // namespace rt
