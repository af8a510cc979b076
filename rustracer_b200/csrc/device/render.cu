// librtgpu.so — host orchestration of the wavefront renderer: rtgpu_render / rtgpu_li_samples /
// rtgpu_generate_rays / film read-out and reduction (include/rtgpu.h).  Replaces the tile loop of
// renderer::render (rustracer-core/src/renderer.rs:22-143).  All kernel launches of one render are queued on the
// context's stream without host synchronisation: queue sizes live in device memory.
#include "context.hpp"
#include "launch.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>      // types only: the library is opened at run time (rtgpu_comm_init), see NcclApi below

using namespace rt;

struct WaveBuffers {
  WaveView v{};
  std::vector<void*> allocs;
  WaveView v2{};                    // second set of wave buffers: two waves of the path integrator in flight (run_render)
  std::vector<void*> allocs2;
  bool recursive = false;
  float* uniform_table = nullptr; int uniform_n = 0;          // UniformLightDistribution
  float* grid_table = nullptr; int grid_nv[3] = {0, 0, 0}; int grid_n = 0; bool grid_valid = false;
  // sparse mode of the spatial light distribution (wave.cuh LightGrid): voxel -> row map, the rows claimed per bounce, counters
  int* grid_slots = nullptr; uint32_t* grid_new = nullptr; uint32_t* grid_counters = nullptr; uint32_t grid_cap_rows = 0;
  uint32_t* n_light_samples = nullptr; uint32_t n_light_samples_cap = 0;
  float4* film_tmp = nullptr; size_t film_tmp_n = 0;
  unsigned long long* stats_backup = nullptr;
};

namespace rt {

static void release(WaveBuffers* w, int set = -1) {
  if (set != 1) { for (void* p : w->allocs) cudaFree(p); w->allocs.clear(); w->v = WaveView{}; }
  if (set != 0) { for (void* p : w->allocs2) cudaFree(p); w->allocs2.clear(); w->v2 = WaveView{}; }
}
void free_wave_buffers(rtgpu_ctx* ctx) {
  if (!ctx->wave) return;
  release(ctx->wave);
  if (ctx->wave->uniform_table) cudaFree(ctx->wave->uniform_table);
  if (ctx->wave->grid_table) cudaFree(ctx->wave->grid_table);
  if (ctx->wave->grid_slots) cudaFree(ctx->wave->grid_slots);
  if (ctx->wave->grid_new) cudaFree(ctx->wave->grid_new);
  if (ctx->wave->grid_counters) cudaFree(ctx->wave->grid_counters);
  if (ctx->wave->n_light_samples) cudaFree(ctx->wave->n_light_samples);
  if (ctx->wave->film_tmp) cudaFree(ctx->wave->film_tmp);
  delete ctx->wave;
  ctx->wave = nullptr;
}
void free_lightgrid(rtgpu_ctx* ctx) {
  if (!ctx->wave) return;
  WaveBuffers* w = ctx->wave;
  if (w->uniform_table) { cudaFree(w->uniform_table); w->uniform_table = nullptr; w->uniform_n = 0; }
  if (w->grid_table) { cudaFree(w->grid_table); w->grid_table = nullptr; }
  if (w->grid_slots) { cudaFree(w->grid_slots); w->grid_slots = nullptr; }
  if (w->grid_new) { cudaFree(w->grid_new); w->grid_new = nullptr; }
  if (w->grid_counters) { cudaFree(w->grid_counters); w->grid_counters = nullptr; }
  w->grid_cap_rows = 0;
  w->grid_valid = false;
}

template <class T> static int dalloc(rtgpu_ctx* ctx, std::vector<void*>& allocs, T** out, size_t count) {
  void* p = nullptr;
  RT_CUDA(ctx, cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  allocs.push_back(p);
  *out = (T*)p;
  return 0;
}

static int ensure_wave(rtgpu_ctx* ctx, uint32_t cap_items, uint32_t cap_samples, uint32_t cap_shadow, uint32_t cap_mis, bool recursive, int set = 0) {
  if (!ctx->wave) ctx->wave = new WaveBuffers();
  WaveBuffers* w = ctx->wave;
  WaveView& v = set ? w->v2 : w->v;
  std::vector<void*>& allocs = set ? w->allocs2 : w->allocs;
  if (v.cap_items >= cap_items && v.cap_samples >= cap_samples && v.cap_shadow >= cap_shadow && v.cap_mis >= cap_mis && (w->recursive || !recursive) && v.counters &&
      (v.hit_inst != nullptr || ctx->scene.n_instances == 0) && (v.rdiff != nullptr || !(recursive && ctx->scene.texmats)) &&
      v.matsort_out != nullptr && v.matsort_bins >= ctx->scene.n_materials + 1 && (v.tex_lobes != nullptr || recursive || !ctx->scene.texmats))
    return 0;
  release(w, set);
  int rc = 0;
#define A(field, n) if ((rc = dalloc(ctx, allocs, &v.field, (n)))) return rc
  A(ray_o, cap_items); A(ray_d, cap_items); A(hit, cap_items); A(beta, cap_items); A(pstate, cap_items);
  A(hit_class, cap_items);
  if (ctx->scene.n_instances) A(hit_inst, cap_items);
  if (recursive) { A(ray_o2, cap_items); A(ray_d2, cap_items); A(beta2, cap_items); A(pstate2, cap_items); }
  if (recursive && ctx->scene.texmats) { A(rdiff, (size_t)cap_items * 3); A(rdiff2, (size_t)cap_items * 3); }
  if (!recursive && ctx->scene.texmats) { A(tex_lobes, (size_t)cap_items * 8); A(tex_frame, (size_t)cap_items * 2); }
  A(raysort_keys, cap_items); A(raysort_hist, (size_t)ray_sort_bins()); A(raysort_out, cap_items);
  A(matsort_hist, (size_t)ctx->scene.n_materials + 1); A(matsort_out, cap_items); v.matsort_bins = ctx->scene.n_materials + 1;
  A(L, cap_samples); A(pfilm, cap_samples); A(sinfo, cap_samples);
  A(sh_o, cap_shadow); A(sh_d, cap_shadow); A(sh_c, cap_shadow);
  A(mi_o, cap_mis); A(mi_d, cap_mis); A(mi_c, cap_mis);
  A(ma_o, cap_mis); A(ma_d, cap_mis); A(ma_c, cap_mis);
  A(list[0], cap_items); A(list[1], cap_items);
  for (int k = 0; k < Q_COUNT; k++) A(matq[k], cap_items);
  A(counters, C_COUNT); A(stats, S_COUNT);
#undef A
  if (!set && (rc = dalloc(ctx, allocs, &w->stats_backup, (size_t)S_COUNT))) return rc;
  v.cap_items = cap_items; v.cap_samples = cap_samples; v.cap_shadow = cap_shadow; v.cap_mis = cap_mis;
  if (!set) w->recursive = recursive;
  return 0;
}

// UniformLightDistribution (lightdistrib.rs:37-54): Distribution1D over n ones, built as distribution1d.rs:11-45
static int ensure_uniform_table(rtgpu_ctx* ctx, int n) {
  WaveBuffers* w = ctx->wave;
  if (w->uniform_table && w->uniform_n == n) return 0;
  if (w->uniform_table) { cudaFree(w->uniform_table); w->uniform_table = nullptr; }
  std::vector<float> t((size_t)2 * n + 2, 0.0f);
  float* func = t.data(); float* cdf = func + n;
  for (int i = 0; i < n; i++) func[i] = 1.0f;
  cdf[0] = 0.0f;
  for (int i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] / (float)n;
  float func_int = cdf[n];
  if (func_int == 0.0f) for (int i = 1; i < n + 1; i++) cdf[i] = (float)i / (float)n;
  else for (int i = 1; i < n + 1; i++) cdf[i] /= func_int;
  t[(size_t)2 * n + 1] = func_int;
  RT_CUDA(ctx, cudaMalloc((void**)&w->uniform_table, t.size() * sizeof(float)));
  RT_CUDA(ctx, cudaMemcpyAsync(w->uniform_table, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  w->uniform_n = n;
  return 0;
}

// SpatialLightDistribution::new (lightdistrib.rs:67-99) + compute_distribution for every voxel (device prepass
// instead of the reference's lazily filled hash table: same per-voxel values).
static int ensure_light_grid(rtgpu_ctx* ctx) {
  WaveBuffers* w = ctx->wave;
  if (w->grid_valid) return 0;
  const DScene& sc = ctx->scene;
  float diag[3]; int widest = 0;
  for (int k = 0; k < 3; k++) diag[k] = sc.world_hi[k] - sc.world_lo[k];
  widest = diag[0] > diag[1] ? (diag[0] > diag[2] ? 0 : 2) : (diag[1] > diag[2] ? 1 : 2);       // Bounds3::maximum_extent
  const float b_max = diag[widest];
  size_t n_voxels = 1;
  for (int k = 0; k < 3; k++) {
    float r = std::round(diag[k] / b_max * 64.0f);
    uint32_t u = !(r == r) ? 0u : (r <= 0.0f ? 0u : (r >= 4294967296.0f ? 0xffffffffu : (uint32_t)r));   // saturating `as u32`
    w->grid_nv[k] = (int)std::max<uint32_t>(1u, std::min<uint32_t>(u, 64u));
    n_voxels *= (size_t)w->grid_nv[k];
  }
  const int n = (int)sc.n_lights;
  const size_t row_floats = (size_t)(2 * n + 2);
  const size_t floats = n_voxels * row_floats;
  if (w->grid_table) { cudaFree(w->grid_table); w->grid_table = nullptr; }
  if (w->grid_slots) { cudaFree(w->grid_slots); w->grid_slots = nullptr; }
  if (w->grid_new) { cudaFree(w->grid_new); w->grid_new = nullptr; }
  if (w->grid_counters) { cudaFree(w->grid_counters); w->grid_counters = nullptr; }
  w->grid_cap_rows = 0;
  const size_t dense_limit = (size_t)ctx->lightgrid_dense_mib << 20;
  if (floats * sizeof(float) <= dense_limit) {
    // few lights: every voxel's distribution up front (dense prepass, same per-voxel values as the reference's lazily filled table)
    RT_CUDA(ctx, cudaMalloc((void**)&w->grid_table, floats * sizeof(float)));
    launch_lightgrid(sc, w->grid_nv[0], w->grid_nv[1], w->grid_nv[2], w->grid_table, ctx->stream);
    ctx->launches += 2;
    RT_CUDA(ctx, cudaGetLastError());
  } else {
    // many lights (an emissive mesh): only the voxels path vertices fall into get a row, claimed and built bounce by bounce
    // (k_lightgrid_mark), as the reference's hash table fills on demand (lightdistrib.rs:221-296)
    const size_t budget = (size_t)ctx->lightgrid_sparse_mib << 20;
    const size_t rows = std::min<size_t>(n_voxels, budget / (row_floats * sizeof(float)));
    if (rows < 64) return fail(ctx, RTGPU_ERR_UNSUPPORTED, "spatial light distribution: one voxel's table over all lights does not fit the budget 64 times; "
                                                            "raise option lightgrid_sparse_mib or use lightsamplestrategy \"uniform\"");
    RT_CUDA(ctx, cudaMalloc((void**)&w->grid_table, rows * row_floats * sizeof(float)));
    RT_CUDA(ctx, cudaMalloc((void**)&w->grid_slots, n_voxels * sizeof(int)));
    RT_CUDA(ctx, cudaMalloc((void**)&w->grid_new, rows * 2 * sizeof(uint32_t)));
    RT_CUDA(ctx, cudaMalloc((void**)&w->grid_counters, 4 * sizeof(uint32_t)));
    RT_CUDA(ctx, cudaMemsetAsync(w->grid_slots, 0xff, n_voxels * sizeof(int), ctx->stream));
    RT_CUDA(ctx, cudaMemsetAsync(w->grid_counters, 0, 4 * sizeof(uint32_t), ctx->stream));
    w->grid_cap_rows = (uint32_t)rows;
  }
  w->grid_n = n; w->grid_valid = true;
  return 0;
}

static uint32_t next_pow2_u32(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

struct Plan {
  RenderParams p{};
  FilmParams film{};
  bool recursive = false;
  uint32_t samples_per_wave_cap = 0;
  bool mat_present[Q_COUNT] = {false, false, false, false, false, false, false, true};
  int extra_rounds = 0;
};

static int fill_params(rtgpu_ctx* ctx, const rtgpu_render_desc* rd, const rtgpu_material* host_materials_unused, Plan& plan) {
  (void)host_materials_unused;
  RenderParams& p = plan.p;
  p.sc = ctx->scene;
  std::memcpy(p.r2c, rd->raster_to_camera, 64); std::memcpy(p.c2w, rd->camera_to_world, 64);
  p.lens_radius = rd->lens_radius; p.focal_distance = rd->focal_distance;
  for (int i = 0; i < 4; i++) { p.sample_bounds[i] = rd->sample_bounds[i]; p.pixel_bounds[i] = rd->pixel_bounds[i]; }
  p.scfg.spp = (uint32_t)std::max(1, rd->spp); p.scfg.dims = (uint32_t)std::max(0, rd->sampler_dims); p.scfg.n_arrays = 0;
  p.seed = rd->seed;
  p.integrator = rd->integrator; p.max_depth = rd->max_depth; p.direct_strategy = rd->direct_strategy; p.ao_samples = rd->ao_samples;
  p.rr_threshold = rd->rr_threshold;
  p.tile_rank = rd->tile_rank; p.tile_world = std::max(1, rd->tile_world);
  if (p.tile_rank < 0 || p.tile_rank >= p.tile_world) return fail(ctx, RTGPU_ERR_ARG, "tile_rank outside [0, tile_world)");
  if (rd->integrator < RTGPU_INTEGRATOR_PATH || rd->integrator > RTGPU_INTEGRATOR_NORMAL) return fail(ctx, RTGPU_ERR_ARG, "unknown integrator");
  if (rd->integrator == RTGPU_INTEGRATOR_AO && rd->ao_samples <= 0) return fail(ctx, RTGPU_ERR_ARG, "ambient occlusion needs ao_samples > 0");
  plan.recursive = rd->integrator == RTGPU_INTEGRATOR_WHITTED || rd->integrator == RTGPU_INTEGRATOR_DIRECT;
  return 0;
}

// Everything between `renderer::render` entry and the film merge, for my (tile, sample) share or an explicit
// sample list.  d_explicit: device {x,y,s} triples (n_explicit of them) or null.  d_li_out: device rgb or null.
static int run_render(rtgpu_ctx* ctx, const rtgpu_render_desc* rd, const int32_t* d_explicit, size_t n_explicit, float* d_li_out, rtgpu_stats* stats) {
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  cudaSetDevice(ctx->device);
  Plan plan;
  int rc = fill_params(ctx, rd, nullptr, plan); if (rc) return rc;
  RenderParams& p = plan.p;
  const DScene& sc = ctx->scene;
  const uint32_t max_depth = (uint32_t)rd->max_depth & 0xffu;

  // ---- capacities --------------------------------------------------------------------------------------------
  // default wave size (tools/wave_sweep.py on B200): the path integrator keeps gaining up to 16 M paths per wave (fewer, fuller
  // launches); the recursive integrators peak at 4 M items (their per-sample atomics and level queues stay L2-resident)
  const bool recursive_integrator = rd->integrator == RTGPU_INTEGRATOR_WHITTED || rd->integrator == RTGPU_INTEGRATOR_DIRECT;
  // Jobs of more than 16 M camera samples on untextured scenes take 32 M-path waves (~10 GB of queues at ~305 B per path): the short
  // late-bounce launches are paid once per wave (profiles/r01zl_wave_sweep_large.log: C5 664 -> 688 M samples/s, C3 641 -> 652 M).
  const uint64_t job_samples = d_explicit ? (uint64_t)n_explicit :
      (uint64_t)std::max(0, rd->sample_bounds[2] - rd->sample_bounds[0]) * (uint64_t)std::max(0, rd->sample_bounds[3] - rd->sample_bounds[1]) *
      (uint64_t)std::max(0, std::min(rd->spp, rd->sample_end) - std::max(0, rd->sample_begin)) / (uint64_t)std::max(1, rd->tile_world);
  const uint32_t P_path = (rd->integrator == RTGPU_INTEGRATOR_PATH && !sc.texmats && job_samples > (1ull << 24)) ? (1u << 25) : (1u << 24);
  // the smallest wave is one 16x16 tile of one sample index: 256 items, whatever the caller asks for
  const uint32_t P = rd->wave_paths > 0 ? std::max(256u, (uint32_t)rd->wave_paths) : (recursive_integrator ? (1u << 22) : P_path);
  uint32_t cap_items = P, cap_samples = P, cap_shadow = P, cap_mis = P;
  uint32_t rays_per_item = 1;
  std::vector<uint32_t> nls;
  if (plan.recursive) {
    if (rd->integrator == RTGPU_INTEGRATOR_WHITTED) rays_per_item = std::max(1u, sc.n_lights);
    else if (rd->direct_strategy == 0) {
      // DirectLightingIntegrator::preprocess (directlighting.rs:70-87): n_samples rounded by the sampler, 2 arrays per light and depth
      const std::vector<rtgpu_light>& hl = ctx->h_lights;
      rays_per_item = 0;
      for (uint32_t j = 0; j < sc.n_lights; j++) { nls.push_back(next_pow2_u32(std::max(1u, hl[j].n_samples))); rays_per_item += nls.back(); }
      rays_per_item = std::max(1u, rays_per_item);
      p.scfg.n_arrays = max_depth * sc.n_lights * 2u;
    }
    // level queues hold P items; a wave starts with P/2 camera samples (the ray tree rarely doubles) and is split on overflow
    cap_items = P;
    cap_shadow = (uint32_t)std::min<size_t>((size_t)cap_items * rays_per_item, (size_t)4 * P);
    cap_samples = std::max(256u, std::min(P / 2, cap_shadow / std::max(1u, rays_per_item)));
    cap_mis = rd->integrator == RTGPU_INTEGRATOR_WHITTED ? 1 : cap_shadow;
  } else if (rd->integrator == RTGPU_INTEGRATOR_AO) {
    cap_shadow = 2 * P;
    cap_samples = cap_items = std::max(256u, cap_shadow / (uint32_t)rd->ao_samples);
    cap_mis = 1;
  } else if (rd->integrator == RTGPU_INTEGRATOR_NORMAL) { cap_shadow = 1; cap_mis = 1; }
  rc = ensure_wave(ctx, cap_items, cap_samples, cap_shadow, cap_mis, plan.recursive); if (rc) return rc;
  WaveBuffers* wb = ctx->wave;
  p.w = wb->v;
  // the capacities the sizing rules assume (buffers may be larger from an earlier render)
  p.w.cap_items = cap_items; p.w.cap_samples = cap_samples; p.w.cap_shadow = cap_shadow; p.w.cap_mis = cap_mis;

  // ---- light distribution (path.rs:86-94) and DirectLighting sample counts -------------------------------------------
  if (rd->integrator == RTGPU_INTEGRATOR_PATH && sc.n_lights > 0) {
    if (rd->light_strategy == 0 || sc.n_lights == 1) {
      rc = ensure_uniform_table(ctx, (int)sc.n_lights); if (rc) return rc;
      p.grid.table = wb->uniform_table; p.grid.nv[0] = p.grid.nv[1] = p.grid.nv[2] = 0; p.grid.n_lights = (int)sc.n_lights;
    } else {
      rc = ensure_light_grid(ctx); if (rc) return rc;
      p.grid.table = wb->grid_table; for (int k = 0; k < 3; k++) p.grid.nv[k] = wb->grid_nv[k]; p.grid.n_lights = wb->grid_n;
      p.grid.slots = wb->grid_slots; p.grid.new_voxels = wb->grid_new; p.grid.grid_counters = wb->grid_counters; p.grid.cap_rows = wb->grid_cap_rows;
    }
  }
  if (!nls.empty()) {
    if (wb->n_light_samples_cap < nls.size()) {
      if (wb->n_light_samples) cudaFree(wb->n_light_samples);
      RT_CUDA(ctx, cudaMalloc((void**)&wb->n_light_samples, nls.size() * 4)); wb->n_light_samples_cap = (uint32_t)nls.size();
    }
    RT_CUDA(ctx, cudaMemcpyAsync(wb->n_light_samples, nls.data(), nls.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p.n_light_samples = wb->n_light_samples;
  }
  // which material classes exist (skip empty shade launches)
  {
    const std::vector<rtgpu_material>& hm = ctx->h_materials;
    auto queue_of = [](uint32_t type) { return material_queue(type); };
    for (const rtgpu_material& m : hm) plan.mat_present[queue_of(m.type)] = true;
    plan.mat_present[Q_NONE] = true;     // primitives without a material row also land here
    bool any_none = false; for (const rtgpu_material& m : hm) any_none |= queue_of(m.type) == Q_NONE;
    plan.extra_rounds = any_none ? 4 : 0;
  }

  // ---- film --------------------------------------------------------------------------------------------------------
  const int fw = rd->cropped[2] - rd->cropped[0], fh = rd->cropped[3] - rd->cropped[1];
  FilmParams& fp = plan.film;
  if (!d_li_out) {
    if (fw <= 0 || fh <= 0) return fail(ctx, RTGPU_ERR_ARG, "empty film");
    const size_t npix = (size_t)fw * fh;
    if (ctx->film_pixels != npix || !ctx->film) {
      if (ctx->film) cudaFree(ctx->film);
      ctx->film = nullptr; ctx->film_pixels = 0;
      RT_CUDA(ctx, cudaMalloc((void**)&ctx->film, npix * sizeof(float4)));
      ctx->film_pixels = npix;
      RT_CUDA(ctx, cudaMemsetAsync(ctx->film, 0, npix * sizeof(float4), ctx->stream));
    } else if (rd->clear_film) RT_CUDA(ctx, cudaMemsetAsync(ctx->film, 0, npix * sizeof(float4), ctx->stream));
    ctx->film_w = fw; ctx->film_h = fh; ctx->film_scale = rd->scale;
    fp.film = (float4*)ctx->film;
    for (int i = 0; i < 4; i++) fp.crop[i] = rd->cropped[i];
    fp.rx = rd->filter_radius[0]; fp.ry = rd->filter_radius[1]; fp.irx = 1.0f / fp.rx; fp.iry = 1.0f / fp.ry;
    fp.max_lum = rd->max_sample_luminance;
    std::memcpy(fp.table, rd->filter_table, sizeof(fp.table));
  }
  fp.ao_div = rd->integrator == RTGPU_INTEGRATOR_AO ? (float)rd->ao_samples : 1.0f;

  // ---- work decomposition ----------------------------------------------------------------------------------------
  const int sbw = rd->sample_bounds[2] - rd->sample_bounds[0], sbh = rd->sample_bounds[3] - rd->sample_bounds[1];
  const int tiles_x = std::max(0, (sbw + 15) / 16), tiles_y = std::max(0, (sbh + 15) / 16);
  const long long n_tiles_total = (long long)tiles_x * tiles_y;
  const long long my_tiles = d_explicit ? 0 : (n_tiles_total > p.tile_rank ? (n_tiles_total - p.tile_rank + p.tile_world - 1) / p.tile_world : 0);
  const int s_begin = std::max(0, rd->sample_begin), s_end = std::min(rd->spp, rd->sample_end);
  p.tiles_x = std::max(1, tiles_x);

  RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  RT_CUDA(ctx, cudaMemsetAsync(p.w.stats, 0, S_COUNT * sizeof(unsigned long long), ctx->stream));
  uint64_t waves = 0, launches0 = ctx->launches, splits = 0;
  const unsigned pblocks = (unsigned)ctx->sm_count * 8u;                // persistent / grid-stride kernels: 8 x 128 threads per SM

  // optional per-class device timing (set_option "profile"): CUDA events around every launch on the context's stream
  enum { K_CLOSEST = 0, K_ANYHIT, K_SHADE, K_OTHER, K_CLASSES };
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  const bool prof = ctx->profile != 0;
  bool has_infinite = false;
  for (const rtgpu_light& l : ctx->h_lights) has_infinite |= l.kind == RTGPU_LIGHT_INFINITE;
  const int tstats = ctx->count_traversal ? TRACE_COUNTING : (ctx->simple_traversal ? TRACE_SIMPLE : TRACE_ENGINE);
  p.hit_t_is_b0 = tstats == TRACE_ENGINE ? 1 : 0;                     // what the closest-hit records of this render carry in .t (wave.cuh)
  size_t ev_used = 0;
  auto next_event = [&]() -> cudaEvent_t {
    if (ev_used == ctx->event_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->event_pool.push_back(e); }
    return ctx->event_pool[ev_used++];
  };
  uint64_t n_closest_launches = 0, n_anyhit_launches = 0;
#define RT_LAUNCH(kclass, call) do { \
    if (prof) { Span sp_; sp_.cls = (kclass); sp_.a = next_event(); sp_.b = next_event(); cudaEventRecord(sp_.a, ctx->stream); call; cudaEventRecord(sp_.b, ctx->stream); spans.push_back(sp_); } \
    else { call; } \
    ctx->launches++; if ((kclass) == K_CLOSEST) n_closest_launches++; else if ((kclass) == K_ANYHIT) n_anyhit_launches++; } while (0)

  struct StreamGroup { cudaStream_t main, side, side2; cudaEvent_t fork, join, fork2, join2; };
  const StreamGroup groups[2] = {{ctx->stream, ctx->side_stream, ctx->side_stream2, ctx->ev_fork, ctx->ev_join, ctx->ev_fork2, ctx->ev_join2},
                                 {ctx->stream_b, ctx->side_stream_b, ctx->side_stream2_b, ctx->ev_fork_b, ctx->ev_join_b, ctx->ev_fork2_b, ctx->ev_join2_b}};
  auto run_wave = [&](RenderParams& p, uint32_t n_items, const StreamGroup& G) -> int {
    if (n_items > p.w.cap_samples || n_items > p.w.cap_items) return fail(ctx, RTGPU_ERR_ARG, "internal: wave larger than its buffers");
    p.n_items = n_items;
    RT_CUDA(ctx, cudaMemsetAsync(p.w.counters, 0, C_COUNT * sizeof(uint32_t), G.main));
    RT_LAUNCH(K_OTHER, launch_raygen(p, G.main));
    if (rd->integrator == RTGPU_INTEGRATOR_PATH) {
      const uint32_t rounds = max_depth + 1 + (uint32_t)plan.extra_rounds;
      // Two streams: the shadow / MIS traces of bounce b (side stream, in the reference's order of additions to L) run beside the
      // closest-hit launch and the classification of bounce b + 1; they join before anything of bounce b + 1 touches L or the
      // shadow / MIS queues again.  The late bounces are short queues whose launches are bound by one warp's walk, not by the
      // machine (profiles/r01z launch list).  Per-class timing (`profile`) and ray binning keep the single-stream order.
      const bool overlap = ctx->overlap_bounces && !prof && !ctx->sort_bounce_rays && sc.n_lights > 0;
      bool joined = true;
      auto trace_closest = [&](uint32_t b) {
        const int in = (int)(b & 1u);
        const uint32_t* live = p.w.list[in];
        if (ctx->sort_bounce_rays && b >= 1) {                        // camera rays are coherent as generated
          RT_LAUNCH(K_CLOSEST, launch_ray_sort(p, live, C_LIVE0 + in, p.w.raysort_keys, p.w.raysort_hist, p.w.raysort_out, pblocks / 2, G.main));
          ctx->launches += 2;
          live = p.w.raysort_out;
        }
        RT_LAUNCH(K_CLOSEST, launch_trace_closest(tstats, p, p.w.ray_o, p.w.ray_d, live, C_LIVE0 + in, p.w.hit, pblocks, G.main));
        RT_LAUNCH(K_SHADE, launch_classify(p, live, C_LIVE0 + in, p.w.hit, tstats == TRACE_ENGINE, pblocks / 2, G.main));
      };
      for (uint32_t b = 0; b < rounds; b++) {
        const int in = (int)(b & 1u);
        if (b == 0 || !overlap) trace_closest(b);                     // overlap: bounce b >= 1 was traced beside the secondary rays of b - 1
        if (!joined) { cudaStreamWaitEvent(G.main, G.join, 0); joined = true; }
        if (p.grid.slots && b < max_depth) {                            // sparse light grid: rows for the voxels this bounce's vertices fall into
          RT_LAUNCH(K_SHADE, launch_lightgrid_bounce(p, p.w.list[in], C_LIVE0 + in, wb->grid_table, pblocks, G.main));
          ctx->launches += 3;
        }
        RT_LAUNCH(K_SHADE, launch_shade_miss(p, pblocks, G.main));
        if (plan.mat_present[Q_MATTE]) RT_LAUNCH(K_SHADE, launch_shade_path_0(p, in, pblocks, G.main));
        if (plan.mat_present[Q_PLASTIC]) RT_LAUNCH(K_SHADE, launch_shade_path_1(p, in, pblocks, G.main));
        if (plan.mat_present[Q_METAL]) RT_LAUNCH(K_SHADE, launch_shade_path_2(p, in, pblocks, G.main));
        if (plan.mat_present[Q_GLASS]) RT_LAUNCH(K_SHADE, launch_shade_path_3(p, in, pblocks, G.main));
        if (plan.mat_present[Q_MIRROR]) RT_LAUNCH(K_SHADE, launch_shade_path_4(p, in, pblocks, G.main));
        if (plan.extra_rounds) RT_LAUNCH(K_SHADE, launch_shade_path_5(p, in, pblocks, G.main));
        if (plan.mat_present[Q_LOBES]) {
          if (ctx->sort_items) {                                      // keep neighbouring warps on one material's lobe list / texture graph (kernels_trace.cuh)
            RT_LAUNCH(K_SHADE, launch_material_sort(p, p.w.matq[Q_LOBES], C_MATQ0 + Q_LOBES, p.w.matsort_hist, sc.n_materials + 1u, p.w.matsort_out, pblocks / 2, G.main));
            ctx->launches += 2;
            RenderParams ps = p; ps.w.matq[Q_LOBES] = p.w.matsort_out;
            if (sc.texmats) RT_LAUNCH(K_SHADE, launch_eval_textured(ps, ps.w.matq[Q_LOBES], pblocks, G.main));
            RT_LAUNCH(K_SHADE, launch_shade_path_6(ps, in, pblocks, G.main));
          } else {
            if (sc.texmats) RT_LAUNCH(K_SHADE, launch_eval_textured(p, p.w.matq[Q_LOBES], pblocks, G.main));
            RT_LAUNCH(K_SHADE, launch_shade_path_6(p, in, pblocks, G.main));
          }
        }
        cudaStream_t sec = G.main;
        if (overlap) { sec = G.side; cudaEventRecord(G.fork, G.main); cudaStreamWaitEvent(sec, G.fork, 0); }
        if (sc.n_lights > 0) {
          RT_LAUNCH(K_ANYHIT, launch_trace_shadow(false, tstats, p, 0, pblocks, sec));
          // The BSDF-sampled MIS ray of a vertex is either a closest-hit ray (area light chosen) or an any-hit ray (infinite light
          // chosen), never both (uniform_sample_one_light, integrator/mod.rs:145-177): the two launches add to disjoint samples of L,
          // after the light sample's shadow ray, so they may run side by side.
          const bool third = overlap && has_infinite && ctx->overlap_bounces > 1;
          if (third) { cudaEventRecord(G.fork2, sec); cudaStreamWaitEvent(G.side2, G.fork2, 0); }
          if (has_infinite) RT_LAUNCH(K_ANYHIT, launch_trace_shadow(false, tstats, p, 1, pblocks, sec));
          RT_LAUNCH(K_CLOSEST, launch_trace_mis(false, tstats, p, pblocks, third ? G.side2 : sec));
          if (third) { cudaEventRecord(G.join2, G.side2); cudaStreamWaitEvent(sec, G.join2, 0); }
        }
        if (!overlap) RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + in, b == 0 ? 1 : 0, 3, G.main));
        else {
          RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + in, 0, 2, sec));
          cudaEventRecord(G.join, sec); joined = false;
          RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + in, b == 0 ? 1 : 0, 1, G.main));
          if (b + 1 < rounds) trace_closest(b + 1);
        }
      }
      if (!joined) { cudaStreamWaitEvent(G.main, G.join, 0); joined = true; }
    } else if (plan.recursive) {
      const uint32_t rounds = std::max(1u, max_depth) + (uint32_t)plan.extra_rounds;
      // same two-stream schedule as the path integrator: the shadow / MIS traces of level l beside the closest-hit launch and the
      // material sort of level l + 1 (these integrators add to L with atomics, so only the queues need the join)
      const bool overlap = ctx->overlap_bounces && !prof;
      bool joined = true;
      auto trace_level = [&](uint32_t lvl) {
        const int par = (int)(lvl & 1u);
        RT_LAUNCH(K_CLOSEST, launch_trace_closest(tstats, p, par ? p.w.ray_o2 : p.w.ray_o, par ? p.w.ray_d2 : p.w.ray_d, nullptr, C_LIVE0 + par, p.w.hit, pblocks, G.main));
        if (ctx->sort_items) {
          RT_LAUNCH(K_SHADE, launch_material_sort(p, nullptr, C_LIVE0 + par, p.w.matsort_hist, sc.n_materials + 1u, p.w.matsort_out, pblocks / 2, G.main));
          ctx->launches += 2;
        }
      };
      for (uint32_t lvl = 0; lvl < rounds; lvl++) {
        const int par = (int)(lvl & 1u);
        if (lvl == 0 || !overlap) trace_level(lvl);
        if (!joined) { cudaStreamWaitEvent(G.main, G.join, 0); joined = true; }
        if (ctx->sort_items) {
          RenderParams ps = p; ps.w.item_order = p.w.matsort_out;
          RT_LAUNCH(K_SHADE, launch_shade_recursive(ps, par, pblocks, G.main));
        } else RT_LAUNCH(K_SHADE, launch_shade_recursive(p, par, pblocks, G.main));
        cudaStream_t sec = G.main;
        if (overlap) { sec = G.side; cudaEventRecord(G.fork, G.main); cudaStreamWaitEvent(sec, G.fork, 0); }
        RT_LAUNCH(K_ANYHIT, launch_trace_shadow(true, tstats, p, 0, pblocks, sec));
        if (rd->integrator == RTGPU_INTEGRATOR_DIRECT && has_infinite) RT_LAUNCH(K_ANYHIT, launch_trace_shadow(true, tstats, p, 1, pblocks, sec));
        if (rd->integrator == RTGPU_INTEGRATOR_DIRECT) RT_LAUNCH(K_CLOSEST, launch_trace_mis(true, tstats, p, pblocks, sec));
        if (!overlap) RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + par, lvl == 0 ? 1 : 0, 3, G.main));
        else {
          RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + par, 0, 2, sec));
          cudaEventRecord(G.join, sec); joined = false;
          RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0 + par, lvl == 0 ? 1 : 0, 1, G.main));
          if (lvl + 1 < rounds) trace_level(lvl + 1);
        }
      }
      if (!joined) { cudaStreamWaitEvent(G.main, G.join, 0); joined = true; }
    } else {
      RT_LAUNCH(K_CLOSEST, launch_trace_closest(tstats, p, p.w.ray_o, p.w.ray_d, p.w.list[0], C_LIVE0, p.w.hit, pblocks, G.main));
      RT_LAUNCH(K_SHADE, launch_shade_ao(p, pblocks, G.main));
      if (rd->integrator == RTGPU_INTEGRATOR_AO) RT_LAUNCH(K_ANYHIT, launch_trace_shadow(true, tstats, p, 0, pblocks, G.main));
      RT_LAUNCH(K_OTHER, launch_next_bounce(p, C_LIVE0, 1, 3, G.main));
    }
    waves++;
    return check_cuda(ctx, cudaGetLastError(), "kernel launch");
  };

  // Work list of waves.  The recursive integrators size their waves for the EXPECTED growth of the ray tree (most hits
  // spawn no specular children), not the 2^depth worst case; a wave whose queues overflow is discarded (nothing has
  // reached the film yet), the counters are rolled back and the wave is split in two.
  struct Chunk { long long a0, an; int s0, sn; };          // tiles [a0, a0+an) x samples [s0, s0+sn)  |  explicit items [a0, a0+an)
  std::vector<Chunk> work;
  if (d_explicit) {
    for (size_t first = n_explicit; first > 0;) {
      const size_t m = std::min<size_t>(cap_samples, first);
      first -= m;
      work.push_back(Chunk{(long long)first, (long long)m, 0, 0});
    }
  } else if (my_tiles > 0 && s_end > s_begin) {
    const long long per_sample = my_tiles * 256;
    long long tiles_per_wave = my_tiles, samples_per_wave = 1;
    if (per_sample <= (long long)cap_samples) samples_per_wave = std::max<long long>(1, (long long)cap_samples / per_sample);
    else tiles_per_wave = std::max<long long>(1, (long long)cap_samples / 256);
    std::vector<Chunk> fwd;
    for (long long t0 = 0; t0 < my_tiles; t0 += tiles_per_wave)
      for (int s0 = s_begin; s0 < s_end; s0 += (int)samples_per_wave)
        fwd.push_back(Chunk{t0, std::min(tiles_per_wave, my_tiles - t0), s0, (int)std::min<long long>(samples_per_wave, s_end - s0)});
    work.assign(fwd.rbegin(), fwd.rend());                  // processed from the back
  }
  // Two waves in flight (path integrator): wave k runs on stream group k & 1 with its own set of wave buffers, so the head of wave k + 1
  // (ray generation, the big first-bounce launches) fills the SMs under the tail of wave k, whose late-bounce launches are short queues
  // bound by one warp's walk.  Waves are independent: each adds its samples to the film with atomics (k_film_add).
  const bool two_sets = rd->integrator == RTGPU_INTEGRATOR_PATH && !d_explicit && !prof && !p.grid.slots && ctx->waves_in_flight > 1 && work.size() > 1 &&
                        !ctx->count_traversal;
  RenderParams p2 = p;
  if (two_sets) {
    rc = ensure_wave(ctx, cap_items, cap_samples, cap_shadow, cap_mis, false, 1); if (rc) return rc;
    p2.w = wb->v2;
    p2.w.cap_items = cap_items; p2.w.cap_samples = cap_samples; p2.w.cap_shadow = cap_shadow; p2.w.cap_mis = cap_mis;
    RT_CUDA(ctx, cudaStreamWaitEvent(groups[1].main, ctx->ev0, 0));    // after the film clear and the light tables
    RT_CUDA(ctx, cudaMemsetAsync(p2.w.stats, 0, S_COUNT * sizeof(unsigned long long), groups[1].main));
  }
  // Processing order and stream group of every wave.  With two sets the waves alternate between the groups, and the second group's
  // first wave is cut in two halves, one run first and one last: the groups then stay half a wave apart, so the short late-bounce launches
  // of one group's wave meet the long first-bounce launches of the other's instead of its tail.
  std::vector<int> group_of(work.size(), 0);                          // indexed like `work` (processed from the back)
  if (two_sets) {
    std::vector<Chunk> order(work.rbegin(), work.rend());
    Chunk h1 = order[1], h2 = order[1];
    const bool by_samples = !d_explicit && order[1].sn >= 2, by_tiles = order[1].an >= 2;
    if (by_samples) { h1.sn = order[1].sn / 2; h2.s0 = order[1].s0 + h1.sn; h2.sn = order[1].sn - h1.sn; }
    else if (by_tiles) { h1.an = order[1].an / 2; h2.a0 = order[1].a0 + h1.an; h2.an = order[1].an - h1.an; }
    std::vector<Chunk> seq; std::vector<int> grp;
    for (size_t k = 0; k < order.size(); k++) {
      if (k == 1 && (by_samples || by_tiles)) { seq.push_back(h1); grp.push_back(1); continue; }
      seq.push_back(order[k]); grp.push_back((int)(k & 1u));
    }
    if (by_samples || by_tiles) { seq.push_back(h2); grp.push_back(1); }
    work.assign(seq.rbegin(), seq.rend());
    group_of.assign(grp.rbegin(), grp.rend());
  }
  while (!work.empty()) {
    const Chunk c = work.back();
    work.pop_back();
    const int g = group_of.empty() ? 0 : group_of.back();
    if (!group_of.empty()) group_of.pop_back();
    RenderParams& pw = g ? p2 : p;
    const StreamGroup& G = groups[g];
    uint32_t n_items;
    if (d_explicit) { pw.explicit_pixels = d_explicit + 3 * c.a0; n_items = (uint32_t)c.an; }
    else { pw.tile_first = (int)c.a0; pw.n_tiles = (int)c.an; pw.sample_first = c.s0; pw.n_samples = c.sn; n_items = (uint32_t)(c.an * 256 * c.sn); }
    if (plan.recursive) RT_CUDA(ctx, cudaMemcpyAsync(wb->stats_backup, pw.w.stats, S_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, G.main));
    rc = run_wave(pw, n_items, G); if (rc) return rc;
    if (plan.recursive) {
      unsigned long long over = 0;
      RT_CUDA(ctx, cudaMemcpyAsync(&over, pw.w.stats + S_OVERFLOW, sizeof(over), cudaMemcpyDeviceToHost, G.main));
      RT_CUDA(ctx, cudaStreamSynchronize(G.main));
      if (over) {
        RT_CUDA(ctx, cudaMemcpyAsync(pw.w.stats, wb->stats_backup, S_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, G.main));
        if (d_explicit ? c.an < 2 : (c.sn < 2 && c.an < 2))
          return fail(ctx, RTGPU_ERR_QUEUE_OVERFLOW, "a wavefront queue overflowed on a minimal wave; raise wave_paths or lower the light sample counts");
        Chunk lo = c, hi = c;
        if (!d_explicit && c.sn >= 2) { lo.sn = c.sn / 2; hi.s0 = c.s0 + lo.sn; hi.sn = c.sn - lo.sn; }
        else { lo.an = c.an / 2; hi.a0 = c.a0 + lo.an; hi.an = c.an - lo.an; }
        work.push_back(hi); work.push_back(lo);
        group_of.push_back(0); group_of.push_back(0);
        splits++;
        continue;
      }
    }
    if (d_explicit) RT_LAUNCH(K_OTHER, launch_li_out(pw.w.L, fp.ao_div, n_items, d_li_out + 3 * c.a0, G.main));
    else RT_LAUNCH(K_OTHER, launch_film_add(fp, pw.w.L, pw.w.pfilm, n_items, G.main));
  }
  if (two_sets) {                                                      // the second group joins the main stream before the end event
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_group_b, groups[1].main));
    RT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_group_b, 0));
  }
  RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  RT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  RT_CUDA(ctx, cudaGetLastError());
  unsigned long long hs[S_COUNT];
  RT_CUDA(ctx, cudaMemcpy(hs, p.w.stats, sizeof(hs), cudaMemcpyDeviceToHost));
  if (two_sets) {
    unsigned long long hs2[S_COUNT];
    RT_CUDA(ctx, cudaMemcpy(hs2, p2.w.stats, sizeof(hs2), cudaMemcpyDeviceToHost));
    for (int k = 0; k < S_COUNT; k++) hs[k] += hs2[k];
  }
  (void)splits;
  if (hs[S_OVERFLOW]) return fail(ctx, RTGPU_ERR_QUEUE_OVERFLOW, "a wavefront queue overflowed; raise wave_paths or lower the light sample counts");
  uint64_t grid_rows = 0;
  if (p.grid.slots) {
    uint32_t gc[4];
    RT_CUDA(ctx, cudaMemcpy(gc, p.grid.grid_counters, sizeof(gc), cudaMemcpyDeviceToHost));
    if (gc[G_OVERFLOW]) {
      free_lightgrid(ctx);                                             // the film holds samples shaded with a wrong distribution: the caller must not use it
      return fail(ctx, RTGPU_ERR_UNSUPPORTED, "spatial light distribution: more occupied voxels than the sparse table holds; raise option lightgrid_sparse_mib "
                                               "or use lightsamplestrategy \"uniform\"");
    }
    grid_rows = gc[G_ROWS];
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->camera_rays = hs[S_CAMERA]; stats->regular_rays = hs[S_REGULAR]; stats->shadow_rays = hs[S_SHADOW];
    stats->waves = waves; stats->kernel_launches = ctx->launches - launches0;
    cudaEventElapsedTime(&stats->ms_total, ctx->ev0, ctx->ev1);
    stats->closest_launches = n_closest_launches; stats->anyhit_launches = n_anyhit_launches;
    stats->nodes_closest = hs[S_NODES_CLOSEST]; stats->prims_closest = hs[S_PRIMS_CLOSEST];
    stats->nodes_anyhit = hs[S_NODES_ANY]; stats->prims_anyhit = hs[S_PRIMS_ANY];
    stats->closest_rays = hs[S_CLOSEST_RAYS]; stats->anyhit_rays = hs[S_ANY_RAYS]; stats->shaded_items = hs[S_VERTICES]; stats->lightgrid_rows = grid_rows;
    float acc[K_CLASSES] = {0, 0, 0, 0};
    for (const Span& sp : spans) { float ms = 0; cudaEventElapsedTime(&ms, sp.a, sp.b); acc[sp.cls] += ms; }
    stats->ms_closest = acc[K_CLOSEST]; stats->ms_anyhit = acc[K_ANYHIT]; stats->ms_shade = acc[K_SHADE]; stats->ms_other = acc[K_OTHER];
  }
#undef RT_LAUNCH
  return RTGPU_OK;
}

}  // namespace rt

extern "C" {

int rtgpu_render(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, rtgpu_stats* stats) {
  if (!ctx || !desc) return RTGPU_ERR_ARG;
  return run_render(ctx, desc, nullptr, 0, nullptr, stats);
}

int rtgpu_li_samples(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, const int32_t* pixels, size_t n, float* rgb) {
  if (!ctx || !desc || (n && (!pixels || !rgb))) return RTGPU_ERR_ARG;
  if (n == 0) return RTGPU_OK;
  cudaSetDevice(ctx->device);
  int32_t* d_pix = nullptr; float* d_out = nullptr;
  RT_CUDA(ctx, cudaMalloc((void**)&d_pix, n * 3 * sizeof(int32_t)));
  cudaError_t e = cudaMalloc((void**)&d_out, n * 3 * sizeof(float));
  if (e != cudaSuccess) { cudaFree(d_pix); return check_cuda(ctx, e, "cudaMalloc"); }
  int rc = check_cuda(ctx, cudaMemcpy(d_pix, pixels, n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice), "cudaMemcpy");
  if (!rc) rc = run_render(ctx, desc, d_pix, n, d_out, nullptr);
  if (!rc) rc = check_cuda(ctx, cudaMemcpy(rgb, d_out, n * 3 * sizeof(float), cudaMemcpyDeviceToHost), "cudaMemcpy");
  cudaFree(d_pix); cudaFree(d_out);
  return rc;
}

int rtgpu_generate_rays(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, const float* samples, size_t n, rtgpu_ray* rays) {
  if (!ctx || !desc || (n && (!samples || !rays))) return RTGPU_ERR_ARG;
  if (n == 0) return RTGPU_OK;
  cudaSetDevice(ctx->device);
  RenderParams p{};
  std::memcpy(p.r2c, desc->raster_to_camera, 64); std::memcpy(p.c2w, desc->camera_to_world, 64);
  p.lens_radius = desc->lens_radius; p.focal_distance = desc->focal_distance;
  float4* d_s = nullptr; float4* d_r = nullptr;
  RT_CUDA(ctx, cudaMalloc((void**)&d_s, n * sizeof(float4)));
  cudaError_t e = cudaMalloc((void**)&d_r, n * 2 * sizeof(float4));
  if (e != cudaSuccess) { cudaFree(d_s); return check_cuda(ctx, e, "cudaMalloc"); }
  int rc = check_cuda(ctx, cudaMemcpyAsync(d_s, samples, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream), "h2d");
  if (!rc) {
    launch_generate_rays(p, d_s, (uint32_t)n, d_r, ctx->stream);
    ctx->launches++;
    rc = check_cuda(ctx, cudaMemcpyAsync(rays, d_r, n * 2 * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream), "d2h");
  }
  if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync");
  cudaFree(d_s); cudaFree(d_r);
  return rc;
}

static int film_tmp(rtgpu_ctx* ctx, size_t n_float4) {
  if (!ctx->wave) ctx->wave = new WaveBuffers();
  WaveBuffers* w = ctx->wave;
  if (w->film_tmp_n >= n_float4) return 0;
  if (w->film_tmp) cudaFree(w->film_tmp);
  w->film_tmp = nullptr; w->film_tmp_n = 0;
  RT_CUDA(ctx, cudaMalloc((void**)&w->film_tmp, n_float4 * sizeof(float4)));
  w->film_tmp_n = n_float4;
  return 0;
}

int rtgpu_read_film(rtgpu_ctx* ctx, float* xyzw) {
  if (!ctx || !xyzw) return RTGPU_ERR_ARG;
  if (!ctx->film) return fail(ctx, RTGPU_ERR_ARG, "no film: call rtgpu_render first");
  cudaSetDevice(ctx->device);
  const size_t n = ctx->film_pixels;
  int rc = film_tmp(ctx, n); if (rc) return rc;
  launch_film_xyz((const float4*)ctx->film, n, ctx->wave->film_tmp, ctx->stream);
  ctx->launches++;
  RT_CUDA(ctx, cudaMemcpyAsync(xyzw, ctx->wave->film_tmp, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RTGPU_OK;
}

int rtgpu_resolve_film(rtgpu_ctx* ctx, float* rgb) {
  if (!ctx || !rgb) return RTGPU_ERR_ARG;
  if (!ctx->film) return fail(ctx, RTGPU_ERR_ARG, "no film: call rtgpu_render first");
  cudaSetDevice(ctx->device);
  const size_t n = ctx->film_pixels;
  int rc = film_tmp(ctx, n); if (rc) return rc;
  launch_film_resolve((const float4*)ctx->film, n, ctx->film_scale, (float*)ctx->wave->film_tmp, ctx->stream);
  ctx->launches++;
  RT_CUDA(ctx, cudaMemcpyAsync(rgb, ctx->wave->film_tmp, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RTGPU_OK;
}

int rtgpu_film_device_ptr(rtgpu_ctx* ctx, void** d_ptr, size_t* n_floats) {
  if (!ctx || !d_ptr || !n_floats) return RTGPU_ERR_ARG;
  if (!ctx->film) return fail(ctx, RTGPU_ERR_ARG, "no film: call rtgpu_render first");
  *d_ptr = ctx->film; *n_floats = ctx->film_pixels * 4;
  return RTGPU_OK;
}

// Single-process multi-GPU: films of ctxs[0..n) summed into ctxs[root] (peer copy into a staging buffer + add).
int rtgpu_reduce_film(rtgpu_ctx** ctxs, int n, int root) {
  if (!ctxs || n <= 0 || root < 0 || root >= n || !ctxs[root]) return RTGPU_ERR_ARG;
  rtgpu_ctx* r = ctxs[root];
  if (!r->film) return fail(r, RTGPU_ERR_ARG, "root has no film");
  const size_t npix = r->film_pixels;
  cudaSetDevice(r->device);
  int rc = film_tmp(r, npix); if (rc) return rc;
  for (int i = 0; i < n; i++) {
    if (i == root) continue;
    rtgpu_ctx* c = ctxs[i];
    if (!c || !c->film || c->film_pixels != npix) return fail(r, RTGPU_ERR_ARG, "film sizes differ between contexts");
    cudaSetDevice(c->device);
    RT_CUDA(r, cudaStreamSynchronize(c->stream));
    cudaSetDevice(r->device);
    RT_CUDA(r, cudaMemcpyPeerAsync(r->wave->film_tmp, r->device, c->film, c->device, npix * sizeof(float4), r->stream));
    launch_film_accumulate((float4*)r->film, r->wave->film_tmp, npix, r->stream);
    r->launches++;
  }
  RT_CUDA(r, cudaStreamSynchronize(r->stream));
  return RTGPU_OK;
}


// ---- multi-process film reduce: one ncclReduce(sum) over NVLink (film.rs:177-194 is the merge it replaces) -----------
// NCCL is opened with dlopen at the first rtgpu_comm_* call instead of being a link-time dependency: a host process that already
// carries an NCCL (e.g. PyTorch's bundled libnccl.so.2) must keep exactly one copy, and a host without any (the Rust binary) gets the
// system library.  Only entry points whose ABI is stable across NCCL 2.x are used.
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
NcclApi* nccl_api() {
  static NcclApi api;
  if (api.handle || !api.error.empty()) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
  if (!api.handle) { api.error = std::string("NCCL not found: ") + dlerror(); return &api; }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.Reduce = (decltype(api.Reduce))dlsym(api.handle, "ncclReduce");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.Reduce || !api.CommDestroy || !api.GetErrorString) { api.error = "NCCL library lacks an expected entry point"; api.handle = nullptr; }
  return &api;
}
int nccl_fail(rtgpu_ctx* ctx, NcclApi* api, ncclResult_t r, const char* what) {
  return fail(ctx, RTGPU_ERR_CUDA, std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
}
}  // namespace

int rtgpu_comm_unique_id(void* id_out) {
  static_assert(sizeof(ncclUniqueId) == RTGPU_COMM_ID_BYTES, "ncclUniqueId size");
  if (!id_out) return RTGPU_ERR_ARG;
  NcclApi* api = nccl_api();
  if (!api->handle) return RTGPU_ERR_UNSUPPORTED;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return RTGPU_ERR_CUDA;
  std::memcpy(id_out, &id, sizeof(id));
  return RTGPU_OK;
}

int rtgpu_comm_init(rtgpu_ctx* ctx, const void* unique_id, int rank, int world) {
  if (!ctx || !unique_id || world < 1 || rank < 0 || rank >= world) return RTGPU_ERR_ARG;
  NcclApi* api = nccl_api();
  if (!api->handle) return fail(ctx, RTGPU_ERR_UNSUPPORTED, api->error);
  if (ctx->comm) { api->CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
  cudaSetDevice(ctx->device);
  ncclUniqueId id; std::memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api->CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, api, r, "ncclCommInitRank");
  ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
  return RTGPU_OK;
}

// The job's one collective: the raw accumulators (sum w*RGB, sum w per pixel — linear, film.rs:357-358) of every rank summed into
// `root`, in place, on the context's stream; blocks until done.  With disjoint tile shares the other ranks contribute zeros.
int rtgpu_reduce_film_nccl(rtgpu_ctx* ctx, int root, float* elapsed_ms) {
  if (!ctx) return RTGPU_ERR_ARG;
  if (!ctx->comm) return fail(ctx, RTGPU_ERR_ARG, "no communicator: call rtgpu_comm_init first");
  if (!ctx->film) return fail(ctx, RTGPU_ERR_ARG, "no film: call rtgpu_render first");
  if (root < 0 || root >= ctx->comm_world) return fail(ctx, RTGPU_ERR_ARG, "root outside the communicator");
  NcclApi* api = nccl_api();
  cudaSetDevice(ctx->device);
  if (elapsed_ms) RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  const ncclResult_t r = api->Reduce(ctx->film, ctx->film, ctx->film_pixels * 4, ncclFloat, ncclSum, root, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, api, r, "ncclReduce");
  if (elapsed_ms) RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (elapsed_ms) RT_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
  return RTGPU_OK;
}

int rtgpu_comm_destroy(rtgpu_ctx* ctx) {
  if (!ctx) return RTGPU_ERR_ARG;
  if (ctx->comm) { NcclApi* api = nccl_api(); if (api->handle) api->CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
  return RTGPU_OK;
}

}  // extern "C"
