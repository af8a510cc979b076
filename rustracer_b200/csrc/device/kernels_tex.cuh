// Texture pass of the path integrator (product code, sm_100a): evaluates the textured materials of a bounce before the listed-lobes
// shade kernel runs.  Kept in its own kernel and translation unit on purpose: inlined into k_shade_path<Q_LOBES>, the texture graph,
// the bump map and the lobe listing pushed that kernel's per-path instruction footprint past the instruction cache and it ran at
// 3-8 % issue utilisation (profiles/r01l_SUMMARY.md).  Here: surface + differentials + bump + parameter textures + lobe listing
// (texture.cuh) for every path of the (material-sorted) listed-lobes queue whose material is RTGPU_MAT_TEXTURED; the lobes
// (<= 8 x 112 B) and the bumped shading frame go to per-path slots the shade kernel picks up.
#pragma once
#include "shade_common.cuh"
#include "texture.cuh"

namespace rt {

#ifndef RT_TEX_THREADS
#define RT_TEX_THREADS 512           // the warps of a block start every iteration together (instruction-cache reuse, see kernels_path.cuh)
#endif
__global__ void __launch_bounds__(RT_TEX_THREADS, 1024 / RT_TEX_THREADS) k_eval_textured(RenderParams p, const uint32_t* __restrict__ list) {
  const uint32_t n = p.w.counters[C_MATQ0 + Q_LOBES];
  const uint32_t max_depth = (uint32_t)p.max_depth & 0xffu;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    __syncthreads();
    const uint32_t i = base + threadIdx.x;
    if (i >= n) continue;
    const uint32_t slot = list[i];
    const HitRec h = p.w.hit[slot];
    const uint4 info = p.sc.info[h.slot];
    if (p.sc.materials[info.y].type != RTGPU_MAT_TEXTURED) continue;
    const uint4 ps = p.w.pstate[slot];
    if ((ps.y & 0xffu) >= max_depth) continue;                        // path.rs:139-141: no scattering functions past the last bounce
    Ray ray = load_ray(p.w.ray_o, p.w.ray_d, slot, nullptr);
    ray.t_max = inf_f();
    SurfHit si; SurfTex st;
    hit_surface_bary(p.sc, h.slot, p.w.hit_inst ? p.w.hit_inst[slot] : kNoInst, ray, p.hit_t_is_b0 != 0, h.t, h.b1, h.b2, si, &st);
    RayDiff rd = no_diff();
    if (ps.z & 2u) {                                                  // still the ray k_raygen made: it has a differential (renderer.rs:110-111)
      const uint32_t sample = ps.x;
      const float2 pf = p.w.pfilm[sample];
      const uint2 sinf = p.w.sinfo[sample];
      rd = camera_ray_diff(p.r2c, p.c2w, p.lens_radius, p.focal_distance, mk2(pf.x, pf.y), draw_2d(sinf.x, sinf.y, p.scfg, 1u), ray,
                           1.0f / sqrtf((float)p.scfg.spp));
    }
    rtgpu_lobe lobes[rtml::kMaxLobes];
    Bsdf bsdf;
    make_bsdf_textured(p.sc, info.y, si, st, rd, true, lobes, bsdf);
    rtgpu_lobe* out = p.w.tex_lobes + (size_t)slot * rtml::kMaxLobes;
    for (int k = 0; k < bsdf.n; k++) out[k] = lobes[k];
    p.w.tex_frame[2 * (size_t)slot] = make_float4(si.ns.x, si.ns.y, si.ns.z, bsdf.eta);
    p.w.tex_frame[2 * (size_t)slot + 1] = make_float4(si.dpdu_s.x, si.dpdu_s.y, si.dpdu_s.z, __uint_as_float((uint32_t)bsdf.n));
  }
}

}  // namespace rt
