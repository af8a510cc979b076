// Translation unit: ray generation, queue traversal, bounce bookkeeping, film and light-grid kernels.
#define RT_QUADRIC_INLINE 1   // shapes.cuh: quadric tests inlined (hot on sphere / disk / cylinder scenes)
#include "kernels_trace.cuh"
#include "launch.hpp"
#include <algorithm>

namespace rt {

void launch_generate_rays(const RenderParams& p, const float4* samples, uint32_t n, float4* rays, cudaStream_t s) {
  k_generate_rays<<<(n + 255) / 256, 256, 0, s>>>(p, samples, n, rays);
}
void launch_raygen(const RenderParams& p, cudaStream_t s) { k_raygen<<<(p.n_items + 255) / 256, 256, 0, s>>>(p); }
// mode: TRACE_ENGINE = persistent while-while engine (production), TRACE_SIMPLE = one-thread-one-ray reference walk
// (validation), TRACE_COUNTING = the reference walk that also counts nodes visited / primitives tested.
void launch_trace_closest(int mode, const RenderParams& p, const float4* ray_o, const float4* ray_d, const uint32_t* list, int count_idx,
                          HitRec* hits, unsigned blocks, cudaStream_t s) {
  if (mode == TRACE_COUNTING) k_trace_closest<true><<<blocks, 128, 0, s>>>(p, ray_o, ray_d, list, count_idx, hits);
  else if (mode == TRACE_SIMPLE) k_trace_closest<false><<<blocks, 128, 0, s>>>(p, ray_o, ray_d, list, count_idx, hits);
  else if (p.sc.n_instances) k_trace_closest_engine<true><<<blocks, 128, 0, s>>>(p, ray_o, ray_d, list, count_idx, hits);
  else k_trace_closest_engine<false><<<blocks, 128, 0, s>>>(p, ray_o, ray_d, list, count_idx, hits);
}
void launch_classify(const RenderParams& p, const uint32_t* list, int count_idx, const HitRec* hits, bool from_hit_class, unsigned blocks, cudaStream_t s) {
  if (from_hit_class) k_classify<true><<<blocks, 256, 0, s>>>(p, list, count_idx, hits);
  else k_classify<false><<<blocks, 256, 0, s>>>(p, list, count_idx, hits);
}
// The engine kernels keep 8 KB of traversal stacks per block in shared memory and run 8 blocks per SM (74 KB with the per-block reserve).  Left alone the
// driver configured 132 KB of shared memory for them (it sizes for the 14 blocks shared memory alone would admit; registers admit 8), which leaves 121 KB
// of the SM's 256 KB as L1 for a kernel whose every node fetch goes through it (profiles/r02m: launch__shared_mem_config_size 135 KB, L1 hit rate 37 - 62 %).
int configure_trace_engines(int pct) {
  if (pct < 0) return 0;
  int bad = 0;
#define RT_CARVE(k) bad |= cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess
  RT_CARVE(k_trace_closest_engine<false>); RT_CARVE(k_trace_closest_engine<true>);
  RT_CARVE((k_trace_shadow_engine<false, false>)); RT_CARVE((k_trace_shadow_engine<false, true>)); RT_CARVE((k_trace_shadow_engine<true, false>)); RT_CARVE((k_trace_shadow_engine<true, true>));
  RT_CARVE((k_trace_mis_engine<false, false>)); RT_CARVE((k_trace_mis_engine<false, true>)); RT_CARVE((k_trace_mis_engine<true, false>)); RT_CARVE((k_trace_mis_engine<true, true>));
#undef RT_CARVE
  return bad;
}
// q: 0 = NEE shadow queue, 1 = queue of the MIS rays towards infinite lights (both any-hit)
void launch_trace_shadow(bool atomic, int mode, const RenderParams& p, int q, unsigned blocks, cudaStream_t s) {
  if (mode == TRACE_COUNTING) { if (atomic) k_trace_shadow<true, true><<<blocks, 128, 0, s>>>(p, q); else k_trace_shadow<false, true><<<blocks, 128, 0, s>>>(p, q); }
  else if (mode == TRACE_SIMPLE) { if (atomic) k_trace_shadow<true, false><<<blocks, 128, 0, s>>>(p, q); else k_trace_shadow<false, false><<<blocks, 128, 0, s>>>(p, q); }
  else if (p.sc.n_instances) { if (atomic) k_trace_shadow_engine<true, true><<<blocks, 128, 0, s>>>(p, q); else k_trace_shadow_engine<false, true><<<blocks, 128, 0, s>>>(p, q); }
  else { if (atomic) k_trace_shadow_engine<true, false><<<blocks, 128, 0, s>>>(p, q); else k_trace_shadow_engine<false, false><<<blocks, 128, 0, s>>>(p, q); }
}
void launch_trace_mis(bool atomic, int mode, const RenderParams& p, unsigned blocks, cudaStream_t s) {
  if (mode == TRACE_COUNTING) { if (atomic) k_trace_mis<true, true><<<blocks, 128, 0, s>>>(p); else k_trace_mis<false, true><<<blocks, 128, 0, s>>>(p); }
  else if (mode == TRACE_SIMPLE) { if (atomic) k_trace_mis<true, false><<<blocks, 128, 0, s>>>(p); else k_trace_mis<false, false><<<blocks, 128, 0, s>>>(p); }
  else if (p.sc.n_instances) { if (atomic) k_trace_mis_engine<true, true><<<blocks, 128, 0, s>>>(p); else k_trace_mis_engine<false, true><<<blocks, 128, 0, s>>>(p); }
  else { if (atomic) k_trace_mis_engine<true, false><<<blocks, 128, 0, s>>>(p); else k_trace_mis_engine<false, false><<<blocks, 128, 0, s>>>(p); }
}
// Sorts the queue (list / count_idx; list == null: the items 0..n-1) by the material row of each item's hit into `out`.
// hist: n_bins = n_materials + 1 counters.  Three launches + one memset, all on the stream.
void launch_material_sort(const RenderParams& p, const uint32_t* list, int count_idx, uint32_t* hist, uint32_t n_bins, uint32_t* out, unsigned blocks, cudaStream_t s) {
  cudaMemsetAsync(hist, 0, sizeof(uint32_t) * n_bins, s);
  k_matsort_hist<<<blocks, 256, 0, s>>>(p, list, count_idx, hist, n_bins);
  k_matsort_scan<<<1, 256, 0, s>>>(hist, n_bins);
  k_matsort_scatter<<<blocks, 256, 0, s>>>(p, list, count_idx, hist, n_bins, out);
}
// Sorts the live list of a bounce by (origin cell, direction octant) into `out`; keys: one uint32 per entry, hist: ray_sort_bins() counters.
uint32_t ray_sort_bins() { return (uint32_t)kWaveSortBins; }
void launch_ray_sort(const RenderParams& p, const uint32_t* list, int count_idx, uint32_t* keys, uint32_t* hist, uint32_t* out, unsigned blocks, cudaStream_t s) {
  WaveSortParams sp;
  for (int k = 0; k < 3; k++) {
    const float ext = p.sc.world_hi[k] - p.sc.world_lo[k];
    sp.lo[k] = p.sc.world_lo[k] - 0.05f * ext;
    sp.inv_ext[k] = ext > 0.0f ? 1.0f / (1.1f * ext) : 0.0f;
  }
  cudaMemsetAsync(hist, 0, sizeof(uint32_t) * kWaveSortBins, s);
  k_raysort_hist<<<blocks, 256, 0, s>>>(p, list, count_idx, sp, keys, hist);
  k_raysort_scan<<<1, 1024, 0, s>>>(hist);
  k_raysort_scatter<<<blocks, 256, 0, s>>>(p, list, count_idx, keys, hist, out);
}
void launch_shade_miss(const RenderParams& p, unsigned blocks, cudaStream_t s) { k_shade_miss<<<blocks, 128, 0, s>>>(p); }
void launch_next_bounce(const RenderParams& p, int live_idx, int count_camera, int part, cudaStream_t s) { k_next_bounce<<<1, 32, 0, s>>>(p, live_idx, count_camera, part); }
void launch_lightgrid(const DScene& sc, int nvx, int nvy, int nvz, float* table, cudaStream_t s) {
  const size_t n_voxels = (size_t)nvx * nvy * nvz, total = n_voxels * sc.n_lights;
  k_lightgrid_contrib<<<(unsigned)std::min<size_t>((total + 127) / 128, (size_t)1 << 30), 128, 0, s>>>(sc, nvx, nvy, nvz, table, nullptr, nullptr);
  k_lightgrid_build<<<(unsigned)((n_voxels + 127) / 128), 128, 0, s>>>((int)sc.n_lights, n_voxels, table, nullptr, nullptr);
}
// Sparse light grid, before shading a bounce: claim rows for the new voxels of the bounce's hit points and build them (4 launches).
void launch_lightgrid_bounce(const RenderParams& p, const uint32_t* list, int count_idx, float* table, unsigned blocks, cudaStream_t s) {
  k_lightgrid_mark<<<blocks, 256, 0, s>>>(p, list, count_idx);
  k_lightgrid_contrib<<<blocks * 2, 128, 0, s>>>(p.sc, p.grid.nv[0], p.grid.nv[1], p.grid.nv[2], table, p.grid.new_voxels, p.grid.grid_counters);
  k_lightgrid_build<<<blocks / 2, 128, 0, s>>>((int)p.sc.n_lights, 0, table, p.grid.new_voxels, p.grid.grid_counters);
  k_lightgrid_new_done<<<1, 32, 0, s>>>(p.grid.grid_counters);
}
void launch_film_add(const FilmParams& f, const float4* L, const float2* pfilm, uint32_t n, cudaStream_t s) { k_film_add<<<(n + 255) / 256, 256, 0, s>>>(f, L, pfilm, n); }
void launch_li_out(const float4* L, float ao_div, uint32_t n, float* out, cudaStream_t s) { k_li_out<<<(n + 255) / 256, 256, 0, s>>>(L, ao_div, n, out); }
void launch_film_xyz(const float4* film, size_t n, float4* out, cudaStream_t s) { k_film_xyz<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(film, n, out); }
void launch_film_resolve(const float4* film, size_t n, float scale, float* rgb, cudaStream_t s) { k_film_resolve<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(film, n, scale, rgb); }
void launch_film_accumulate(float4* dst, const float4* src, size_t n, cudaStream_t s) { k_film_accumulate<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, src, n); }

}  // namespace rt
