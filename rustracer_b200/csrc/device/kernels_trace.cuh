// Wavefront kernels: ray generation, queue traversal (closest / shadow / MIS), bounce bookkeeping, film accumulation and
// resolve, spatial light-distribution prepass (product code, sm_100a).  Each kernel names the rustracer code it replaces.
#pragma once
#include "shade_common.cuh"
#include "trace_engine.cuh"

namespace rt {

__global__ void __launch_bounds__(256) k_generate_rays(RenderParams p, const float4* __restrict__ samples, uint32_t n, float4* __restrict__ rays) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 s = samples[i];
  Ray r = camera_ray(p.r2c, p.c2w, p.lens_radius, p.focal_distance, mk2(s.x, s.y), mk2(s.z, s.w));
  rays[2 * (size_t)i] = make_float4(r.o.x, r.o.y, r.o.z, r.t_max);
  rays[2 * (size_t)i + 1] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
}

// ---- k_raygen: get_camera_sample + generate_ray (renderer.rs:109-111, zerotwosequence.rs:182-192) ------------
// Item i of the wave -> (pixel, sample index): sample-major, then my tiles, then the 16x16 pixels of a tile, so
// the 32 lanes of a warp are two adjacent pixel rows of one tile.
__global__ void __launch_bounds__(256) k_raygen(RenderParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < p.n_items;
  int x = 0, y = 0; uint32_t s = 0;
  if (valid) {
    if (p.explicit_pixels) { x = p.explicit_pixels[3 * (size_t)i]; y = p.explicit_pixels[3 * (size_t)i + 1]; s = (uint32_t)p.explicit_pixels[3 * (size_t)i + 2]; }
    else {
      const uint32_t per_sample = (uint32_t)p.n_tiles * 256u;
      const uint32_t j = i / per_sample, rem = i - j * per_sample;
      const uint32_t tile_ord = (uint32_t)p.tile_first + rem / 256u, pix = rem & 255u;
      const uint32_t tile = (uint32_t)p.tile_rank + tile_ord * (uint32_t)p.tile_world;
      // tile number -> tile of the frame: row ty, column rotated by ty.  With a tile count per row that is a multiple of the rank count
      // (3840 / 16 = 240 columns over 8 ranks) the plain row-major numbering would hand every rank the same columns of every row —
      // vertical stripes, whose work follows the scene's layout (profiles/r02o: the slowest of 8 ranks 2 % behind); rotated, the stripes run diagonally
      const int ty = (int)(tile / (uint32_t)p.tiles_x), tx = (int)((tile % (uint32_t)p.tiles_x + (uint32_t)ty) % (uint32_t)p.tiles_x);
      x = p.sample_bounds[0] + tx * 16 + (int)(pix & 15u);
      y = p.sample_bounds[1] + ty * 16 + (int)(pix >> 4);
      s = (uint32_t)p.sample_first + j;
      valid = x < p.sample_bounds[2] && y < p.sample_bounds[3] &&
              x >= p.pixel_bounds[0] && x < p.pixel_bounds[2] && y >= p.pixel_bounds[1] && y < p.pixel_bounds[3];   // renderer.rs:103-105
    }
  }
  // queue positions for the valid items: one global atomic per block (a 32 M-sample wave issued 1 M same-address atomics per warp-level
  // append, which bounded the kernel: 61 ps per sample against 17 ps for its 100 bytes of traffic, profiles/r02f_launches_c5.csv)
  const uint32_t pos = block_append(&p.w.counters[C_LIVE0], valid);
  if (i < p.n_items) {
    SamplerState ss; ss.ph = pixel_hash(x, y, p.seed); ss.s = s; ss.d1 = 0; ss.d2 = 0; ss.da = 0;
    P2 u = ss.get_2d(p.scfg);
    P2 p_film = mk2((float)x + u.x, (float)y + u.y);
    (void)ss.get_1d(p.scfg);                                          // time: drawn, unused by the camera
    P2 p_lens = ss.get_2d(p.scfg);
    p.w.L[i] = make_float4(0.0f, 0.0f, 0.0f, valid ? 1.0f : 0.0f);
    st_stream(&p.w.pfilm[i], make_float2(p_film.x, p_film.y));
    st_stream(&p.w.sinfo[i], make_uint2(ss.ph, s));
    if (valid) {                                                      // items are compacted: item slot = queue position
      Ray ray = camera_ray(p.r2c, p.c2w, p.lens_radius, p.focal_distance, p_film, p_lens);
      store_ray(p.w.ray_o, p.w.ray_d, pos, ray, 0);
      st_stream(&p.w.beta[pos], make_float4(1.0f, 1.0f, 1.0f, 1.0f));
      // path: z bit 1 = "camera ray" (carries the ray differential); recursive: z = depth, w != 0 = camera ray
      p.w.pstate[pos] = make_uint4(i, p.integrator == RTGPU_INTEGRATOR_PATH ? 0u : 1u, p.integrator == RTGPU_INTEGRATOR_PATH ? 2u : 0u, ss.d1 | (ss.d2 << 16));
      st_stream(&p.w.list[0][pos], pos);
    }
  }
}

// ---- queue traversal ----------------------------------------------------------------------------------------
// Persistent warps pull 32-entry packets from the queue with an atomic cursor (validation / counting walk; the
// production path is the engine below).  Material classification is k_classify's job.
// STATS variants also count BVH nodes visited / primitives tested (the N and T of the roofline's bytes per ray).
RT_DEV void flush_trav_stats(const RenderParams& p, const TravStats& st, int s_nodes, int s_prims) {
  unsigned long long nn = st.nodes, np = st.prims;
  for (int off = 16; off > 0; off >>= 1) { nn += __shfl_down_sync(0xffffffffu, nn, off); np += __shfl_down_sync(0xffffffffu, np, off); }
  if (lane_id() == 0) { atomicAdd(&p.w.stats[s_nodes], nn); atomicAdd(&p.w.stats[s_prims], np); }
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_trace_closest(RenderParams p, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                        const uint32_t* __restrict__ list, int count_idx, HitRec* __restrict__ hits) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  TravStats st; st.nodes = 0; st.prims = 0;
  while (true) {
    const uint32_t base = warp_fetch(&p.w.counters[C_CUR_CLOSEST]);
    if (base >= n) break;
    const uint32_t i = base + lane_id();
    if (i < n) {
      const uint32_t slot = list ? list[i] : i;
      Ray ray = load_ray(ray_o, ray_d, slot, nullptr);
      HitRec h; uint32_t inst;
      bvh_traverse<false, STATS>(p.sc, ray, h, &st, &inst);
      hits[slot] = h;
      if (p.w.hit_inst) p.w.hit_inst[slot] = inst;
    }
  }
  if (STATS) flush_trav_stats(p, st, S_NODES_CLOSEST, S_PRIMS_CLOSEST);
}

// Shadow rays: VisibilityTester::unoccluded (light/mod.rs:52-55) -> Scene::intersect_p; adds the pending
// contribution to the sample's radiance when the segment is clear.
// An any-hit queue: q = 0 the NEE shadow rays, q = 1 the MIS rays towards infinite lights (shade_common.cuh).  Each
// camera sample has at most one entry per queue and bounce in the path integrator, so the plain read-modify-write of
// the non-ATOMIC variant is race-free and the sum order is fixed.
struct AnyQueue { const float4 *o, *d, *c; uint32_t n; uint32_t* cursor; };
RT_DEV AnyQueue any_queue(const RenderParams& p, int q) {
  AnyQueue a;
  if (q == 0) { a.o = p.w.sh_o; a.d = p.w.sh_d; a.c = p.w.sh_c; a.n = min(p.w.counters[C_SHADOW], p.w.cap_shadow); a.cursor = &p.w.counters[C_CUR_ANY]; }
  else { a.o = p.w.ma_o; a.d = p.w.ma_d; a.c = p.w.ma_c; a.n = min(p.w.counters[C_MIS_ANY], p.w.cap_mis); a.cursor = &p.w.counters[C_CUR_MISANY]; }
  return a;
}

template <bool ATOMIC, bool STATS>
__global__ void __launch_bounds__(128) k_trace_shadow(RenderParams p, int q) {
  const AnyQueue aq = any_queue(p, q);
  const uint32_t n = aq.n;
  TravStats st; st.nodes = 0; st.prims = 0;
  while (true) {
    const uint32_t base = warp_fetch(aq.cursor);
    if (base >= n) break;
    const uint32_t i = base + lane_id();
    if (i >= n) continue;
    uint32_t sample;
    Ray ray = load_ray(aq.o, aq.d, i, &sample);
    HitRec h;
    if (!bvh_traverse<true, STATS>(p.sc, ray, h, &st)) {
      const float4 c = aq.c[i];
      float4* L = &p.w.L[sample];
      if (ATOMIC) { atomicAdd(&L->x, c.x); atomicAdd(&L->y, c.y); atomicAdd(&L->z, c.z); }
      else { float4 v = *L; v.x += c.x; v.y += c.y; v.z += c.z; *L = v; }
    }
  }
  if (STATS) flush_trav_stats(p, st, S_NODES_ANY, S_PRIMS_ANY);
}

// MIS rays: second half of estimate_direct (integrator/mod.rs:291-313): closest hit of the BSDF-sampled ray; it
// contributes only if it lands on the sampled light (or escapes to the sampled infinite light).
template <bool ATOMIC, bool STATS>
__global__ void __launch_bounds__(128) k_trace_mis(RenderParams p) {
  const uint32_t n = min(p.w.counters[C_MIS], p.w.cap_mis);
  TravStats st; st.nodes = 0; st.prims = 0;
  while (true) {
    const uint32_t base = warp_fetch(&p.w.counters[C_CUR_MIS]);
    if (base >= n) break;
    const uint32_t i = base + lane_id();
    if (i >= n) continue;
    uint32_t sample;
    Ray ray = load_ray(p.w.mi_o, p.w.mi_d, i, &sample);
    const Ray ray0 = ray;
    const float4 c = p.w.mi_c[i];
    const uint32_t light_row = __float_as_uint(c.w);
    const rtgpu_light& light = p.sc.lights[light_row];
    HitRec h; uint32_t inst;
    Spec li = spec(0.0f);
    if (bvh_traverse<false, STATS>(p.sc, ray, h, &st, &inst)) {
      if (inst == kNoInst && p.sc.info[h.slot].z == light_row) {      // same light id (integrator/mod.rs:294-299)
        SurfHit si; float t;
        if (slot_intersect_surface(p.sc, h.slot, ray0, t, si)) li = area_L(light, si.n, -ray0.d);
      }
    } else li = light_le(p.sc, light, ray0.d);
    if (!is_black(li)) {
      float4* L = &p.w.L[sample];
      const float r = c.x * li.r, g = c.y * li.g, b = c.z * li.b;
      if (ATOMIC) { atomicAdd(&L->x, r); atomicAdd(&L->y, g); atomicAdd(&L->z, b); }
      else { float4 v = *L; v.x += r; v.y += g; v.z += b; *L = v; }
    }
  }
  if (STATS) flush_trav_stats(p, st, S_NODES_CLOSEST, S_PRIMS_CLOSEST);
}

// ---- the same three kernels on the persistent while-while engine (trace_engine.cuh): the production path ------------
template <bool INST>
struct ClosestPolicy {
  const float4* ray_o; const float4* ray_d; const uint32_t* list; HitRec* hits; uint32_t* hit_inst; uint8_t* hit_class; uint32_t slot;
  RT_DEV ClosestPolicy(const float4* o, const float4* d, const uint32_t* l, HitRec* h, uint32_t* hi, uint8_t* hc) : ray_o(o), ray_d(d), list(l), hits(h), hit_inst(hi), hit_class(hc), slot(0) {}
  RT_DEV void load(uint32_t idx, Ray& ray) { slot = list ? list[idx] : idx; ray = load_ray(ray_o, ray_d, slot, nullptr); }
  // the record keeps the three barycentrics {b0, slot, b1, b2} (RenderParams::hit_t_is_b0): the shade kernels rebuild the surface from them instead of
  // repeating the watertight test, and nothing downstream reads the distance
  RT_DEV void commit(uint32_t, const HitRec& h, float, uint32_t inst, uint32_t cls) { st_stream((float4*)&hits[slot], make_float4(h.t, __uint_as_float(h.slot), h.b1, h.b2)); hit_class[slot] = (uint8_t)cls; if (INST) hit_inst[slot] = inst; }
};
template <bool INST>
__global__ void __launch_bounds__(128, RT_ENGINE_MIN_BLOCKS) k_trace_closest_engine(RenderParams p, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                               const uint32_t* __restrict__ list, int count_idx, HitRec* __restrict__ hits) {
  ClosestPolicy<INST> pol(ray_o, ray_d, list, hits, p.w.hit_inst, p.w.hit_class);
  trace_engine<false, INST>(p.sc, &p.w.counters[C_CUR_CLOSEST], min(p.w.counters[count_idx], p.w.cap_items), pol);
}

// Material classification of the traced paths: appends each path to the queue of its hit material (or the miss queue)
// for the material-sorted shade kernels.  A streaming pass over the live list at full lane occupancy (one atomic per
// distinct class and warp via match_any) instead of seven ballot rounds inside every refill of the traversal engine.
// FROM_CLASS: the engine left each hit's queue id in hit_class (it rides in the hit slot's geometry record), so the pass
// reads 5 bytes per path instead of chasing hit -> primitive info -> material row.
// Queue positions are claimed per BLOCK: every thread classifies kClassifyPerThread paths, the block counts each class in shared
// memory and issues one global atomic per class for all of them (with one atomic per class and warp the eight counters took 17 M
// same-address atomics per C5 step and the pass ran at 32 ps per path instead of ~3: profiles/r02f_launches_c5.csv).
constexpr int kClassifyPerThread = 8;
template <bool FROM_CLASS>
__global__ void __launch_bounds__(256) k_classify(RenderParams p, const uint32_t* __restrict__ list, int count_idx, const HitRec* __restrict__ hits) {
  __shared__ uint32_t s_count[Q_COUNT], s_base[Q_COUNT];
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  const uint32_t lane = lane_id(), lane_lt = (1u << lane) - 1u;
  const uint32_t tile = 256u * kClassifyPerThread;
  for (uint32_t base = blockIdx.x * tile; base < n; base += gridDim.x * tile) {
    if (threadIdx.x < Q_COUNT) s_count[threadIdx.x] = 0;
    __syncthreads();
    int q[kClassifyPerThread]; uint32_t slot[kClassifyPerThread], off[kClassifyPerThread];
#pragma unroll
    for (int k = 0; k < kClassifyPerThread; k++) {
      const uint32_t i = base + (uint32_t)k * 256u + threadIdx.x;
      q[k] = -1; slot[k] = 0; off[k] = 0;
      if (i < n) {
        slot[k] = list ? ld_stream(&list[i]) : i;
        if (FROM_CLASS) q[k] = (int)ld_stream(&p.w.hit_class[slot[k]]);
        else {
          const uint32_t hslot = hits[slot[k]].slot;
          if (hslot == kMiss) q[k] = Q_MISS;
          else {
            const uint32_t mrow = p.sc.info[hslot].y;
            const uint32_t type = mrow < p.sc.n_materials ? p.sc.materials[mrow].type : (uint32_t)RTGPU_MAT_NONE;
            q[k] = material_queue(type);
          }
        }
      }
      const unsigned peers = __match_any_sync(0xffffffffu, q[k]);
      const int leader = __ffs(peers) - 1;
      uint32_t wbase = 0;
      if (q[k] >= 0 && (int)lane == leader) wbase = atomicAdd(&s_count[q[k]], (uint32_t)__popc(peers));   // shared-memory atomic
      wbase = __shfl_sync(0xffffffffu, wbase, leader);
      off[k] = wbase + (uint32_t)__popc(peers & lane_lt);
    }
    __syncthreads();
    if (threadIdx.x < Q_COUNT && s_count[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&p.w.counters[C_MATQ0 + threadIdx.x], s_count[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kClassifyPerThread; k++) if (q[k] >= 0) st_stream(&p.w.matq[q[k]][s_base[q[k]] + off[k]], slot[k]);
    __syncthreads();
  }
}

// ---- material sort of a shade queue (textured scenes) --------------------------------------------------------------------
// The listed-lobes shade kernel and the recursive shade kernel evaluate textured materials with very different code per
// material (texture graph, EWA loops, bump map, lobe list); unsorted, the warps of an SM wander through megabytes of code and
// starve on instruction fetch (profiles/r01k: 95 % of the stall samples "no instruction").  A counting sort of the queue by
// material row keeps neighbouring warps on the same material: histogram, exclusive scan, scatter.
// Entry i of the queue is item `list ? list[i] : i`; its key is the material row of its hit (n_bins - 1: miss / no material).
RT_DEV uint32_t matsort_key(const RenderParams& p, uint32_t item, uint32_t n_bins) {
  const uint32_t hslot = p.w.hit[item].slot;
  if (hslot == kMiss) return n_bins - 1u;
  const uint32_t mrow = p.sc.info[hslot].y;
  return mrow < n_bins - 1u ? mrow : n_bins - 1u;
}
__global__ void __launch_bounds__(256) k_matsort_hist(RenderParams p, const uint32_t* __restrict__ list, int count_idx, uint32_t* __restrict__ hist, uint32_t n_bins) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31u) & ~31u); i += gridDim.x * blockDim.x) {
    const uint32_t key = i < n ? matsort_key(p, list ? list[i] : i, n_bins) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (i < n && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&hist[key], (uint32_t)__popc(peers));
  }
}
__global__ void __launch_bounds__(256) k_matsort_scan(uint32_t* __restrict__ hist, uint32_t n_bins) {   // exclusive, in place, one block
  __shared__ uint32_t warp_sums[8];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n_bins; b0 += 256) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < n_bins ? hist[i] : 0;
    uint32_t x = v;
    for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((int)lane_id() >= off) x += y; }
    if (lane_id() == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = carry;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_sums[w];
    if (i < n_bins) hist[i] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = before + x;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_matsort_scatter(RenderParams p, const uint32_t* __restrict__ list, int count_idx, uint32_t* __restrict__ cursor, uint32_t n_bins,
                                                         uint32_t* __restrict__ out) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  const uint32_t lane_lt = (1u << lane_id()) - 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31u) & ~31u); i += gridDim.x * blockDim.x) {
    const uint32_t item = i < n ? (list ? list[i] : i) : 0u;
    const uint32_t key = i < n ? matsort_key(p, item, n_bins) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (i < n && (int)lane_id() == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) out[base + (uint32_t)__popc(peers & lane_lt)] = item;
  }
}

// ---- binning of a bounce's rays (option "sort_bounce_rays") ------------------------------------------------------------------
// Same key as the batch API's ray binning (api.cu): Morton code of the origin's cell in a 32^3 grid over the world bounds, then
// the direction octant.  A counting sort of the live list by that key before the closest-hit launch of a bounce.
constexpr int kWaveCellBits = 5;
constexpr int kWaveSortBins = 1 << (3 * kWaveCellBits + 3);
struct WaveSortParams { float lo[3], inv_ext[3]; };
RT_DEV uint32_t wave_spread3(uint32_t v) { v &= 0x3ffu; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v; }
RT_DEV uint32_t wave_ray_key(const float4 o, const float4 d, const WaveSortParams& sp) {
  const float cells = (float)(1 << kWaveCellBits);
  const int hi = (1 << kWaveCellBits) - 1;
  const int cx = min(max(__float2int_rd((o.x - sp.lo[0]) * sp.inv_ext[0] * cells), 0), hi);
  const int cy = min(max(__float2int_rd((o.y - sp.lo[1]) * sp.inv_ext[1] * cells), 0), hi);
  const int cz = min(max(__float2int_rd((o.z - sp.lo[2]) * sp.inv_ext[2] * cells), 0), hi);
  const uint32_t morton = wave_spread3((uint32_t)cx) | (wave_spread3((uint32_t)cy) << 1) | (wave_spread3((uint32_t)cz) << 2);
  return (morton << 3) | (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
}
__global__ void __launch_bounds__(256) k_raysort_hist(RenderParams p, const uint32_t* __restrict__ list, int count_idx, WaveSortParams sp, uint32_t* __restrict__ keys,
                                                      uint32_t* __restrict__ hist) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = list[i];
    const uint32_t k = wave_ray_key(p.w.ray_o[slot], p.w.ray_d[slot], sp);
    keys[i] = k;
    atomicAdd(&hist[k], 1u);
  }
}
__global__ void __launch_bounds__(1024) k_raysort_scan(uint32_t* __restrict__ hist) {    // exclusive scan of kWaveSortBins counters, one block
  __shared__ uint32_t partial[1024];
  constexpr int per = kWaveSortBins / 1024;
  const uint32_t base = threadIdx.x * per;
  uint32_t sum = 0;
  for (int i = 0; i < per; i++) sum += hist[base + i];
  partial[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = threadIdx.x >= (uint32_t)off ? partial[threadIdx.x - off] : 0;
    __syncthreads();
    partial[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = partial[threadIdx.x] - sum;
  for (int i = 0; i < per; i++) { const uint32_t c = hist[base + i]; hist[base + i] = run; run += c; }
}
__global__ void __launch_bounds__(256) k_raysort_scatter(RenderParams p, const uint32_t* __restrict__ list, int count_idx, const uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ offsets, uint32_t* __restrict__ out) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);   // a level queue that overflowed keeps counting past its capacity
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[atomicAdd(&offsets[keys[i]], 1u)] = list[i];
}

template <bool ATOMIC>
struct ShadowPolicy {
  const RenderParams& p; AnyQueue aq; uint32_t sample;
  RT_DEV ShadowPolicy(const RenderParams& p_, const AnyQueue& a) : p(p_), aq(a), sample(0) {}
  RT_DEV void load(uint32_t idx, Ray& ray) { ray = load_ray(aq.o, aq.d, idx, &sample); }
  RT_DEV void commit(uint32_t idx, const HitRec& h, float, uint32_t, uint32_t) {
    if (h.slot != kMiss) return;
    const float4 c = ld_stream(&aq.c[idx]);
    float4* L = &p.w.L[sample];
    if (ATOMIC) { atomicAdd(&L->x, c.x); atomicAdd(&L->y, c.y); atomicAdd(&L->z, c.z); }
    else { float4 v = *L; v.x += c.x; v.y += c.y; v.z += c.z; *L = v; }
  }
};
template <bool ATOMIC, bool INST>
__global__ void __launch_bounds__(128, RT_ENGINE_MIN_BLOCKS) k_trace_shadow_engine(RenderParams p, int q) {
  const AnyQueue aq = any_queue(p, q);
  ShadowPolicy<ATOMIC> pol(p, aq);
  trace_engine<true, INST>(p.sc, aq.cursor, aq.n, pol);
}

template <bool ATOMIC>
struct MisPolicy {
  const RenderParams& p; uint32_t sample;
  RT_DEV MisPolicy(const RenderParams& p_) : p(p_), sample(0) {}
  RT_DEV void load(uint32_t idx, Ray& ray) { ray = load_ray(p.w.mi_o, p.w.mi_d, idx, &sample); }
  RT_DEV void commit(uint32_t idx, const HitRec& h, float, uint32_t inst, uint32_t) {
    const float4 c = ld_stream(&p.w.mi_c[idx]);
    const uint32_t light_row = __float_as_uint(c.w);
    const rtgpu_light& light = p.sc.lights[light_row];
    const Ray ray0 = load_ray(p.w.mi_o, p.w.mi_d, idx, nullptr);
    Spec li = spec(0.0f);
    if (h.slot != kMiss) {
      // an area light is never inside an object instance (the front end rejects it), so a hit there cannot be the sampled light
      if (inst == kNoInst && p.sc.info[h.slot].z == light_row) {        // same light id (integrator/mod.rs:294-299)
        SurfHit si; float t;
        if (slot_intersect_surface(p.sc, h.slot, ray0, t, si)) li = area_L(light, si.n, -ray0.d);
      }
    } else li = light_le(p.sc, light, ray0.d);
    if (!is_black(li)) {
      float4* L = &p.w.L[sample];
      const float r = c.x * li.r, g = c.y * li.g, b = c.z * li.b;
      if (ATOMIC) { atomicAdd(&L->x, r); atomicAdd(&L->y, g); atomicAdd(&L->z, b); }
      else { float4 v = *L; v.x += r; v.y += g; v.z += b; *L = v; }
    }
  }
};
template <bool ATOMIC, bool INST>
__global__ void __launch_bounds__(128, RT_ENGINE_MIN_BLOCKS) k_trace_mis_engine(RenderParams p) {
  MisPolicy<ATOMIC> pol(p);
  trace_engine<false, INST>(p.sc, &p.w.counters[C_CUR_MIS], min(p.w.counters[C_MIS], p.w.cap_mis), pol);
}

// compute_distribution (lightdistrib.rs:101-179), first half: one thread per (voxel, light) accumulates the
// 128 Halton-point estimates sequentially, in the reference's order.  Dense mode: every voxel of the grid, row = voxel.
// Sparse mode (`list` non-null): the {voxel, row} pairs k_lightgrid_mark claimed for this bounce, count read from the device.
RT_DEV float lightgrid_contrib(const DScene& sc, int nvx, int nvy, int nvz, size_t voxel, int j) {
  const int pz = (int)(voxel % nvz), py = (int)((voxel / nvz) % nvy), px = (int)(voxel / ((size_t)nvz * nvy));
  const float lo[3] = {sc.world_lo[0], sc.world_lo[1], sc.world_lo[2]}, hi[3] = {sc.world_hi[0], sc.world_hi[1], sc.world_hi[2]};
  const int pi[3] = {px, py, pz}, nv[3] = {nvx, nvy, nvz};
  float vlo[3], vhi[3];
  for (int k = 0; k < 3; k++) {
    float t0 = (float)pi[k] / (float)nv[k], t1 = ((float)pi[k] + 1.0f) / (float)nv[k];
    float a = lo[k] * (1.0f - t0) + hi[k] * t0, b = lo[k] * (1.0f - t1) + hi[k] * t1;   // Bounds3::lerp
    vlo[k] = pmin(a, b); vhi[k] = pmax(a, b);                                             // Bounds3::from_points
  }
  const rtgpu_light& light = sc.lights[j];
  float contrib = 0.0f;
  for (uint64_t i = 0; i < 128; i++) {
    float t[3] = {radical_inverse(0, i), radical_inverse(1, i), radical_inverse(2, i)};
    V3 po = v3(vlo[0] * (1.0f - t[0]) + vhi[0] * t[0], vlo[1] * (1.0f - t[1]) + vhi[1] * t[1], vlo[2] * (1.0f - t[2]) + vhi[2] * t[2]);
    Inter intr = inter_point(po);
    P2 u = mk2(radical_inverse(3, i), radical_inverse(4, i));
    V3 wi; float pdf; Inter p1;
    Spec li = light_sample_li(sc, light, intr, u, wi, pdf, p1);
    if (pdf > 0.0f) contrib += lum(li) / pdf;
  }
  return contrib;
}
__global__ void __launch_bounds__(128) k_lightgrid_contrib(DScene sc, int nvx, int nvy, int nvz, float* __restrict__ table, const uint32_t* __restrict__ list,
                                                            const uint32_t* __restrict__ grid_counters) {
  const int n = (int)sc.n_lights;
  const size_t n_rows = list ? (size_t)grid_counters[G_NEW] : (size_t)nvx * nvy * nvz;
  const size_t total = n_rows * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % n);
    const size_t k = idx / n;
    const size_t voxel = list ? (size_t)list[2 * k] : k, row = list ? (size_t)list[2 * k + 1] : k;
    table[row * (size_t)(2 * n + 2) + j] = lightgrid_contrib(sc, nvx, nvy, nvz, voxel, j);
  }
}
// second half: floor at 0.1 % of the average, then Distribution1D::new (distribution1d.rs:11-45)
__global__ void __launch_bounds__(128) k_lightgrid_build(int n, size_t n_voxels, float* __restrict__ table, const uint32_t* __restrict__ list,
                                                          const uint32_t* __restrict__ grid_counters) {
  const size_t n_rows = list ? (size_t)grid_counters[G_NEW] : n_voxels;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_rows; k += (size_t)gridDim.x * blockDim.x) {
    const size_t row = list ? (size_t)list[2 * k + 1] : k;
    float* func = table + row * (size_t)(2 * n + 2);
    float* cdf = func + n;
    float sum = 0.0f;
    for (int j = 0; j < n; j++) sum += func[j];
    const float avg = sum / (float)(128ull * (unsigned long long)n);
    const float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (int j = 0; j < n; j++) func[j] = fmaxf(func[j], min_contrib);
    cdf[0] = 0.0f;
    for (int i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] / (float)n;
    const float func_int = cdf[n];
    if (func_int == 0.0f) for (int i = 1; i < n + 1; i++) cdf[i] = (float)i / (float)n;
    else for (int i = 1; i < n + 1; i++) cdf[i] /= func_int;
    func[2 * n + 1] = func_int;
  }
}
// Sparse mode, before a bounce is shaded: every hit point of the bounce's live list claims the row of its voxel if the voxel has none
// yet (the device-side counterpart of the reference's lazy insert under compare-and-swap, lightdistrib.rs:221-296).  The point is the
// one the shade kernel will look up (hit_surface's p), so the voxel is the same.
__global__ void __launch_bounds__(256) k_lightgrid_mark(RenderParams p, const uint32_t* __restrict__ list, int count_idx) {
  const uint32_t n = min(p.w.counters[count_idx], p.w.cap_items);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = list ? list[i] : i;
    const HitRec h = p.w.hit[slot];
    if (h.slot == kMiss) continue;
    Ray ray = load_ray(p.w.ray_o, p.w.ray_d, slot, nullptr);
    ray.t_max = inf_f();
    SurfHit si;
    hit_surface_bary(p.sc, h.slot, p.w.hit_inst ? p.w.hit_inst[slot] : kNoInst, ray, p.hit_t_is_b0 != 0, h.t, h.b1, h.b2, si);
    const size_t voxel = grid_voxel(p.sc, p.grid, si.p);
    if (p.grid.slots[voxel] != -1) continue;
    if (atomicCAS(&p.grid.slots[voxel], -1, -2) != -1) continue;      // somebody else claims it
    const uint32_t row = atomicAdd(&p.grid.grid_counters[G_ROWS], 1u);
    if (row >= p.grid.cap_rows) { p.grid.grid_counters[G_OVERFLOW] = 1; p.grid.slots[voxel] = 0; continue; }
    const uint32_t k = atomicAdd(&p.grid.grid_counters[G_NEW], 1u);
    p.grid.new_voxels[2 * (size_t)k] = (uint32_t)voxel; p.grid.new_voxels[2 * (size_t)k + 1] = row;
    p.grid.slots[voxel] = (int)row;                                   // read by the shade kernels, which run after the rows are built
  }
}
__global__ void k_lightgrid_new_done(uint32_t* grid_counters) { if (threadIdx.x == 0 && blockIdx.x == 0) grid_counters[G_NEW] = 0; }

// Escaped paths: emission of the infinite lights, then the path ends (path.rs:127-141).
__global__ void __launch_bounds__(128) k_shade_miss(RenderParams p) {
  const uint32_t n = p.w.counters[C_MATQ0 + Q_MISS];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = p.w.matq[Q_MISS][i];
    const uint4 ps = p.w.pstate[slot];
    const uint32_t bounces = ps.y & 0xffu; const bool specular_bounce = (ps.z & 1u) != 0;
    if (bounces == 0 || specular_bounce) {
      const float4 bt = p.w.beta[slot];
      const Spec beta = spec(bt.x, bt.y, bt.z);
      const float4 d = p.w.ray_d[slot];
      float4 L = p.w.L[ps.x];
      for (uint32_t j = 0; j < p.sc.n_lights; j++) {
        if (p.sc.lights[j].kind != RTGPU_LIGHT_INFINITE) continue;
        Spec c = beta * light_le(p.sc, p.sc.lights[j], v3(d.x, d.y, d.z));
        L.x += c.r; L.y += c.g; L.z += c.b;
      }
      p.w.L[ps.x] = L;
    }
  }
}

// End of a bounce: fold the queue sizes into the reference's ray counters and reset the per-bounce queues.
// part bit 0: the live list, the material queues and the closest-hit cursor (what the next bounce's closest-hit launch needs);
// part bit 1: the shadow / MIS queues and their cursors.  The path integrator runs the two halves on two streams
// (render.cu: the secondary traces of bounce b overlap the closest-hit launch of bounce b + 1), hence the atomic adds.
__global__ void k_next_bounce(RenderParams p, int live_idx, int count_camera, int part) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t* c = p.w.counters;
  if (c[C_OVERFLOW]) p.w.stats[S_OVERFLOW] = 1;
  if (part & 1) {
    const uint32_t live = min(c[live_idx], p.w.cap_items);
    atomicAdd(&p.w.stats[S_REGULAR], (unsigned long long)live);
    atomicAdd(&p.w.stats[S_CLOSEST_RAYS], (unsigned long long)live);
    p.w.stats[S_VERTICES] += live;                                    // items handed to the shade kernels (part 1 runs on one stream)
    if (count_camera) p.w.stats[S_CAMERA] += live;
    c[live_idx] = 0;
    for (int k = 0; k < Q_COUNT; k++) c[C_MATQ0 + k] = 0;
    c[C_CUR_CLOSEST] = 0;
  }
  if (part & 2) {
    // MIS rays towards infinite lights travel in the shadow queue (shade_common.cuh) but are "regular" rays for the reference
    atomicAdd(&p.w.stats[S_REGULAR], (unsigned long long)min(c[C_MIS], p.w.cap_mis) + min(c[C_MIS_ANY], p.w.cap_mis) + c[C_MIS_SKIPPED]);
    p.w.stats[S_SHADOW] += min(c[C_SHADOW], p.w.cap_shadow);
    atomicAdd(&p.w.stats[S_CLOSEST_RAYS], (unsigned long long)min(c[C_MIS], p.w.cap_mis));
    p.w.stats[S_ANY_RAYS] += (unsigned long long)min(c[C_SHADOW], p.w.cap_shadow) + min(c[C_MIS_ANY], p.w.cap_mis);
    c[C_SHADOW] = 0; c[C_MIS] = 0; c[C_MIS_ANY] = 0; c[C_MIS_SKIPPED] = 0; c[C_CUR_ANY] = 0; c[C_CUR_MIS] = 0; c[C_CUR_MISANY] = 0;
  }
}

// ---- film (film.rs) -----------------------------------------------------------------------------------------------
// renderer.rs:115-126 guards + FilmTile::add_sample (film.rs:298-361) + merge (film.rs:177-194, kept as RGB sums)
__global__ void __launch_bounds__(256) k_film_add(FilmParams f, const float4* __restrict__ L, const float2* __restrict__ pfilm, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 l = L[i];
  if (l.w == 0.0f) return;
  Spec c = spec(l.x, l.y, l.z) / f.ao_div;                             // AO: n_clear / n_samples (ao.rs:57); 1 otherwise
  if (has_nan(c)) c = spec(0.0f);
  if (lum(c) < -1e-5f) c = spec(0.0f);
  if (isinf(lum(c))) c = spec(0.0f);
  if (lum(c) > f.max_lum) c = c * f.max_lum / lum(c);
  const float2 pf = pfilm[i];
  const float dx = pf.x - 0.5f, dy = pf.y - 0.5f;
  const float p0x = ceilf(dx - f.rx), p0y = ceilf(dy - f.ry);
  const float p1x = floorf(dx + f.rx + 1.0f), p1y = floorf(dy + f.ry + 1.0f);
  const int x0 = f2i32(pmax(pmin(p0x, p1x), (float)f.crop[0])), y0 = f2i32(pmax(pmin(p0y, p1y), (float)f.crop[1]));
  const int x1 = f2i32(pmin(pmax(p0x, p1x), (float)f.crop[2])), y1 = f2i32(pmin(pmax(p0y, p1y), (float)f.crop[3]));
  const int w = f.crop[2] - f.crop[0];
  for (int y = y0; y < y1; y++) {
    const float fy = fabsf(((float)y - dy) * f.iry * 16.0f);
    const int iy = (int)f2u32(fminf(floorf(fy), 15.0f));
    for (int x = x0; x < x1; x++) {
      const float fx = fabsf(((float)x - dx) * f.irx * 16.0f);
      const int ix = (int)f2u32(fminf(floorf(fx), 15.0f));
      const float wgt = f.table[iy * 16 + ix];
      const Spec cw = c * wgt;
      atomicAdd(&f.film[(size_t)(y - f.crop[1]) * w + (x - f.crop[0])], make_float4(cw.r, cw.g, cw.b, wgt));
    }
  }
}
__global__ void __launch_bounds__(256) k_li_out(const float4* __restrict__ L, float ao_div, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 l = L[i];
  out[3 * (size_t)i] = l.x / ao_div; out[3 * (size_t)i + 1] = l.y / ao_div; out[3 * (size_t)i + 2] = l.z / ao_div;
}
// Film pixel accumulators as the reference keeps them: XYZ + weight (film.rs:38-43,187-192)
__global__ void __launch_bounds__(256) k_film_xyz(const float4* __restrict__ film, size_t n, float4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = film[i];
  float xyz[3]; to_xyz(spec(v.x, v.y, v.z), xyz);
  out[i] = make_float4(xyz[0], xyz[1], xyz[2], v.w);
}
// Film::write_image arithmetic (film.rs:196-234)
__global__ void __launch_bounds__(256) k_film_resolve(const float4* __restrict__ film, size_t n, float scale, float* __restrict__ rgb) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = film[i];
  float xyz[3]; to_xyz(spec(v.x, v.y, v.z), xyz);
  Spec c = from_xyz(xyz[0], xyz[1], xyz[2]);
  if (v.w != 0.0f) { const float inv = 1.0f / v.w; c = spec(fmaxf(0.0f, c.r * inv), fmaxf(0.0f, c.g * inv), fmaxf(0.0f, c.b * inv)); }
  const Spec sp = from_xyz(0.0f, 0.0f, 0.0f);                          // splat term (film.rs:222-230): zero on this path
  c = spec(c.r + 1.0f * sp.r, c.g + 1.0f * sp.g, c.b + 1.0f * sp.b);
  rgb[3 * i] = c.r * scale; rgb[3 * i + 1] = c.g * scale; rgb[3 * i + 2] = c.b * scale;
}
__global__ void __launch_bounds__(256) k_film_accumulate(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = dst[i]; const float4 b = src[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  dst[i] = a;
}

}  // namespace rt
