// Material-sorted shading kernels of the path integrator (product code, sm_100a).
#pragma once
#include "shade_common.cuh"

namespace rt {

// ---- path integrator: one bounce of PathIntegrator::li (integrator/path.rs:96-215) -------------------------------
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS 4      // resident 128-thread blocks per SM the path shade kernels are compiled for (register cap 128)
#endif
#ifndef RT_SHADE_MIN_BLOCKS_LOBES
#define RT_SHADE_MIN_BLOCKS_LOBES 8   // the listed-lobes / textured kernel is instruction-fetch bound: more resident warps (64 registers) win 5-20 % (profiles/r01o)
#endif
#ifndef RT_LOBES_THREADS
#define RT_LOBES_THREADS 512         // block size of the listed-lobes kernel; its warps start every iteration together (below)
#endif
#ifndef RT_SHADE_THREADS
#define RT_SHADE_THREADS 512         // block size of the per-material kernels; above 128 their warps also start every iteration together
#endif
template <int MAT>
__global__ void __launch_bounds__(MAT == Q_LOBES ? RT_LOBES_THREADS : RT_SHADE_THREADS,
                                  MAT == Q_LOBES ? RT_SHADE_MIN_BLOCKS_LOBES * 128 / RT_LOBES_THREADS : RT_SHADE_MIN_BLOCKS * 128 / RT_SHADE_THREADS)
k_shade_path(RenderParams p, int parity) {
  const uint32_t n = p.w.counters[C_MATQ0 + MAT];
  uint32_t* out_list = p.w.list[1 - parity];
  uint32_t* out_count = &p.w.counters[C_LIVE0 + (1 - parity)];
  const uint32_t max_depth = (uint32_t)p.max_depth & 0xffu;            // `max_ray_depth as u8` (path.rs:42)
  // Block-uniform trip count.  The shade kernels run a long straight-line body and are (partly) instruction-fetch bound
  // (profiles/r01l_SUMMARY.md): the 16 warps of a 512-thread block start every iteration together, so an instruction-cache line is
  // fetched once per block instead of once per warp.  128 -> 512 threads with the barrier: listed lobes -20 %, textured -19 %, the
  // per-material kernels -14 % on a five-material scene and unchanged on the all-matte C3 scene.
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    if (MAT == Q_LOBES || RT_SHADE_THREADS > 128) __syncthreads();
    const uint32_t i = base + threadIdx.x;
    bool alive = false;
    uint32_t slot = 0;
    if (i < n) {
      slot = ld_stream(&p.w.matq[MAT][i]);
      Ray ray = load_ray(p.w.ray_o, p.w.ray_d, slot, nullptr);
      ray.t_max = inf_f();
      const float4 hv = ld_stream((const float4*)&p.w.hit[slot]);
      HitRec h; h.t = hv.x; h.slot = __float_as_uint(hv.y); h.b1 = hv.z; h.b2 = hv.w;
      const float4 bt = ld_stream(&p.w.beta[slot]);
      Spec beta = spec(bt.x, bt.y, bt.z); float eta_scale = bt.w;
      uint4 ps = ld_stream(&p.w.pstate[slot]);
      const uint32_t sample = ps.x;
      uint32_t bounces = ps.y & 0xffu; bool specular_bounce = (ps.z & 1u) != 0;
      const uint2 sinf = ld_stream(&p.w.sinfo[sample]);
      SamplerState ss; ss.ph = sinf.x; ss.s = sinf.y; ss.d1 = ps.w & 0xffffu; ss.d2 = ps.w >> 16; ss.da = 0;
      SurfHit si;
      hit_surface_bary(p.sc, h.slot, p.w.hit_inst ? p.w.hit_inst[slot] : kNoInst, ray, p.hit_t_is_b0 != 0, h.t, h.b1, h.b2, si);
      const uint4 info = p.sc.info[h.slot];
      Spec l_add = spec(0.0f);
      if ((bounces == 0 || specular_bounce) && info.z != kNoLight) l_add = beta * area_L(p.sc.lights[info.z], si.n, -ray.d);   // path.rs:127-131
      if (bounces < max_depth) {                                        // path.rs:139-141
        const Inter it = inter_of(si);
        if (MAT == Q_NONE) {                                            // path.rs:146-152 (u8 wrap)
          Ray nr = spawn_ray(it, ray.d);
          store_ray(p.w.ray_o, p.w.ray_d, slot, nr, 0);
          bounces = (bounces - 1u) & 0xffu;
          alive = true;
        } else {
          Bsdf bsdf;
          if (MAT == Q_LOBES && p.sc.materials[info.y].type == RTGPU_MAT_TEXTURED) {
            // textures, bump map and lobe list were evaluated by the texture pass (kernels_tex.cuh): pick up the lobes and the bumped frame
            const float4 f0 = p.w.tex_frame[2 * (size_t)slot], f1 = p.w.tex_frame[2 * (size_t)slot + 1];
            si.ns = v3(f0.x, f0.y, f0.z); si.dpdu_s = v3(f1.x, f1.y, f1.z);
            bsdf_init(bsdf, si, f0.w);
            bsdf.g = p.w.tex_lobes + (size_t)slot * 8; bsdf.n = (int)__float_as_uint(f1.w);
          } else
          make_bsdf(MAT == Q_LOBES ? (uint32_t)RTGPU_MAT_LOBES : (uint32_t)MAT, p.sc.materials[info.y], p.sc.lobes, si, true, bsdf);
          if (bsdf_num_components(bsdf, kBsdfNonSpecular) > 0 && p.sc.n_lights > 0) {   // path.rs:161-171 -> uniform_sample_one_light
            const Distrib dist = lookup_distrib(p, si.p);
            const float s = ss.get_1d(p.scfg);
            float light_pdf;
            const int light_num = dist1d_sample_discrete(dist.func, dist.cdf, dist.n, dist.func_int, s, light_pdf);
            if (light_pdf != 0.0f) {
              const P2 u_light = ss.get_2d(p.scfg);
              const P2 u_scattering = ss.get_2d(p.scfg);
              estimate_direct(p, si, bsdf, u_scattering, (uint32_t)light_num, u_light, beta / light_pdf, sample);
            }
          }
          const V3 wo = -ray.d;
          Spec f; V3 wi; float pdf; uint32_t flags;
          bsdf_sample_f(bsdf, wo, ss.get_2d(p.scfg), BSDF_ALL, f, wi, pdf, flags);
          if (!(is_black(f) || pdf <= 0.0f)) {                          // path.rs:174-176
            beta = beta * f * fabsf(dot(wi, si.ns)) / pdf;
            specular_bounce = (flags & BSDF_SPECULAR) != 0;
            if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
              const float eta = bsdf.eta;
              eta_scale *= dot(wo, si.n) > 0.0f ? eta * eta : 1.0f / (eta * eta);
            }
            Ray nr = spawn_ray(it, wi);
            alive = true;
            const Spec rr_beta = beta * eta_scale;                      // path.rs:199-209
            if (max_component_value(rr_beta) < p.rr_threshold && bounces > 3) {
              const float q = fmaxf(1.0f - max_component_value(rr_beta), 0.05f);
              if (ss.get_1d(p.scfg) < q) alive = false;
              beta = beta / (1.0f - q);
            }
            bounces = (bounces + 1u) & 0xffu;
            if (alive) store_ray(p.w.ray_o, p.w.ray_d, slot, nr, 0);
          }
        }
        if (alive) {
          st_stream(&p.w.beta[slot], make_float4(beta.r, beta.g, beta.b, eta_scale));
          st_stream(&p.w.pstate[slot], make_uint4(sample, bounces, specular_bounce ? 1u : 0u, (ss.d1 & 0xffffu) | (ss.d2 << 16)));
        }
      }
      if (!is_black(l_add)) { float4 L = p.w.L[sample]; L.x += l_add.r; L.y += l_add.g; L.z += l_add.b; p.w.L[sample] = L; }
    }
    const uint32_t pos = warp_append(out_count, alive);
    if (alive) st_stream(&out_list[pos], slot);
  }
}

}  // namespace rt
