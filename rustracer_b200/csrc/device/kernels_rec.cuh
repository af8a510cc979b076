// Shading kernels of the Whitted / DirectLighting / AmbientOcclusion / Normal integrators (product code, sm_100a).
#pragma once
#include "shade_common.cuh"
#include "texture.cuh"

namespace rt {

// ---- Whitted / DirectLighting: one level of the recursion (whitted.rs:41-99, directlighting.rs:89-143,
// integrator/mod.rs:49-142).  Radiance is linear, so every item carries its throughput and adds into its
// camera sample; specular reflection / transmission spawn child items for the next level.
// TEX: the scene has textured materials — every item carries its ray differential (integrator/mod.rs:64-83, :107-136) in
// the rdiff buffers and textured materials are evaluated at the hit (texture.cuh); scenes without them run the lean variant.
#ifndef RT_REC_MIN_BLOCKS
#define RT_REC_MIN_BLOCKS 8          // 64 registers: 4 / 8 / 12 / 16 blocks per SM measured in profiles/r01o
#endif
#ifndef RT_REC_THREADS
#define RT_REC_THREADS 512           // the warps of a block start every iteration together (instruction-cache reuse, see kernels_path.cuh)
#endif
template <bool TEX>
__global__ void __launch_bounds__(RT_REC_THREADS, RT_REC_MIN_BLOCKS * 128 / RT_REC_THREADS) k_shade_recursive(RenderParams p, int parity) {
  const uint32_t n = min(p.w.counters[C_LIVE0 + parity], p.w.cap_items);   // an overflowed level (the wave is discarded and split) must not read past its buffers
  const float4* ray_o = parity ? p.w.ray_o2 : p.w.ray_o; const float4* ray_d = parity ? p.w.ray_d2 : p.w.ray_d;
  const float4* beta_in = parity ? p.w.beta2 : p.w.beta; const uint4* ps_in = parity ? p.w.pstate2 : p.w.pstate;
  float4* oray_o = parity ? p.w.ray_o : p.w.ray_o2; float4* oray_d = parity ? p.w.ray_d : p.w.ray_d2;
  float4* obeta = parity ? p.w.beta : p.w.beta2; uint4* ops = parity ? p.w.pstate : p.w.pstate2;
  uint32_t* out_count = &p.w.counters[C_LIVE0 + (1 - parity)];
  const uint32_t max_depth = (uint32_t)p.max_depth & 0xffu;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    __syncthreads();
    const uint32_t k = base + threadIdx.x;
    if (k >= n) continue;
    const uint32_t i = p.w.item_order ? p.w.item_order[k] : k;           // items in material order (k_matsort_*)
    Ray ray = load_ray(ray_o, ray_d, i, nullptr);
    ray.t_max = inf_f();
    const HitRec h = p.w.hit[i];
    const float4 bt = beta_in[i];
    const Spec beta = spec(bt.x, bt.y, bt.z);
    const uint4 ps = ps_in[i];
    const uint32_t sample = ps.x, node = ps.y, depth = ps.z;
    float4* L = &p.w.L[sample];
    Spec colour = spec(0.0f);
    if (h.slot == kMiss) {                                              // whitted.rs:91-94: every light's le(ray)
      for (uint32_t j = 0; j < p.sc.n_lights; j++) colour = colour + light_le(p.sc, p.sc.lights[j], ray.d);
    } else {
      SurfHit si;
      SurfTex st; RayDiff rd; rtgpu_lobe hit_lobes[TEX ? rtml::kMaxLobes : 1];
      hit_surface_bary(p.sc, h.slot, p.w.hit_inst ? p.w.hit_inst[i] : kNoInst, ray, p.hit_t_is_b0 != 0, h.t, h.b1, h.b2, si, TEX ? &st : nullptr);
      const uint4 info = p.sc.info[h.slot];
      const uint32_t mtype = info.y < p.sc.n_materials ? p.sc.materials[info.y].type : (uint32_t)RTGPU_MAT_NONE;
      const V3 n_before = si.ns;                                        // whitted.rs:53: shading normal BEFORE the scattering functions (bump map)
      Bsdf bsdf;
      bool have_bsdf = material_queue(mtype) != Q_NONE;
      if (TEX) {
        rd = no_diff();
        if (ps.w >> 16) {                                               // the camera ray itself: differential from the camera sample
          const float2 pf = p.w.pfilm[sample]; const uint2 sf = p.w.sinfo[sample];
          rd = camera_ray_diff(p.r2c, p.c2w, p.lens_radius, p.focal_distance, mk2(pf.x, pf.y), draw_2d(sf.x, sf.y, p.scfg, 1u), ray, 1.0f / sqrtf((float)p.scfg.spp));
        } else if (ps.w & 1u) {
          const float4* q = (parity ? p.w.rdiff2 : p.w.rdiff) + 3 * (size_t)i;
          const float4 a = q[0], b = q[1], c = q[2];
          rd.has = true; rd.rx_o = v3(a.x, a.y, a.z); rd.ry_o = v3(a.w, b.x, b.y); rd.rx_d = v3(b.z, b.w, c.x); rd.ry_d = v3(c.y, c.z, c.w);
        }
        if (have_bsdf && mtype == RTGPU_MAT_TEXTURED) make_bsdf_textured(p.sc, info.y, si, st, rd, false, hit_lobes, bsdf);
        else { compute_differential(si, st, rd); if (have_bsdf) have_bsdf = make_bsdf(mtype, p.sc.materials[info.y], p.sc.lobes, si, false, bsdf); }
      } else if (have_bsdf) have_bsdf = make_bsdf(mtype, p.sc.materials[info.y], p.sc.lobes, si, false, bsdf);
      const Inter it = inter_of(si);
      if (!have_bsdf) {
        // no material: continue the same node through the surface (whitted.rs:60-63); the spawned ray has no differential
        const uint32_t pos = warp_append(out_count, true);
        if (pos < p.w.cap_items) { store_ray(oray_o, oray_d, pos, spawn_ray(it, ray.d), 0); obeta[pos] = bt; ops[pos] = make_uint4(ps.x, ps.y, ps.z, 0u); }
        else p.w.counters[C_OVERFLOW] = 1;
        continue;
      }
      const uint2 sinf = p.w.sinfo[sample];
      SamplerState ss; ss.ph = sinf.x; ss.s = sinf.y;
      if (node == 1) { ss.d1 = 1; ss.d2 = 2; ss.da = 0; } else { ss.d1 = ss.d2 = ss.da = 64u * node; }   // oracle-twin node keying
      const V3 wo = si.wo, ns = si.ns;
      if (info.z != kNoLight) colour = colour + area_L(p.sc.lights[info.z], si.n, wo);
      if (p.integrator == RTGPU_INTEGRATOR_WHITTED) {                   // whitted.rs:70-82
        for (uint32_t j = 0; j < p.sc.n_lights; j++) {
          V3 wi; float pdf; Inter p1;
          Spec li = light_sample_li(p.sc, p.sc.lights[j], it, ss.get_2d(p.scfg), wi, pdf, p1);
          if (is_black(li) || pdf == 0.0f) continue;
          Spec f = bsdf_f(bsdf, wo, wi, BSDF_ALL);
          if (!is_black(f)) push_shadow(p, spawn_ray_to(it, p1), sample, beta * (f * li * fabsf(dot(wi, n_before)) / pdf));
        }
      } else if (p.sc.n_lights > 0) {
        if (p.direct_strategy == 0) {                                   // uniform_sample_all_light (integrator/mod.rs:145-184)
          for (uint32_t j = 0; j < p.sc.n_lights; j++) {
            const uint32_t ns_j = p.n_light_samples[j];
            const uint32_t ca = ss.da++, cb = ss.da++;
            for (uint32_t k = 0; k < ns_j; k++) {
              P2 ul = draw_2d_array(ss.ph, ss.s, p.scfg, ns_j, k, ca);
              P2 us = draw_2d_array(ss.ph, ss.s, p.scfg, ns_j, k, cb);
              estimate_direct(p, si, bsdf, us, j, ul, beta / (float)ns_j, sample);
            }
          }
        } else {                                                        // uniform_sample_one_light, no distribution (:186-220)
          const uint32_t nl = p.sc.n_lights;
          const float s = ss.get_1d(p.scfg);
          const uint32_t light_num = min(nl - 1u, f2u32(s * (float)nl));
          const float light_pdf = 1.0f / (float)nl;
          const P2 u_light = ss.get_2d(p.scfg);
          const P2 u_scattering = ss.get_2d(p.scfg);
          estimate_direct(p, si, bsdf, u_scattering, light_num, u_light, beta / light_pdf, sample);
        }
      }
      if (depth + 1 < max_depth) {                                      // whitted.rs:83-88
#pragma unroll 1
        for (uint32_t pass = 0; pass < 2; pass++) {
          const uint32_t flags = (pass == 0 ? BSDF_REFLECTION : BSDF_TRANSMISSION) | BSDF_SPECULAR;
          Spec f; V3 wi; float pdf; uint32_t sampled_type;
          bsdf_sample_f(bsdf, wo, ss.get_2d(p.scfg), flags, f, wi, pdf, sampled_type);
          if (pdf > 0.0f && !is_black(f) && fabsf(dot(wi, ns)) != 0.0f) {
            const Spec cb = beta * (f * fabsf(dot(wi, ns)) / pdf);
            const uint32_t pos = warp_append(out_count, true);
            if (pos < p.w.cap_items) {
              store_ray(oray_o, oray_d, pos, spawn_ray(it, wi), 0);
              obeta[pos] = make_float4(cb.r, cb.g, cb.b, 1.0f);
              ops[pos] = make_uint4(sample, node * 2u + pass, depth + 1u, (TEX && rd.has) ? 1u : 0u);
              if (TEX && rd.has) {                                      // integrator/mod.rs:64-83 (reflection), :107-136 (transmission)
                const V3 zero_n = v3(0, 0, 0);                          // shading.dndu / dndv / isect.dndv: always zero (SurfTex)
                const V3 dndx = zero_n * st.dudx + zero_n * st.dvdx, dndy = zero_n * st.dudy + zero_n * st.dvdy;
                const V3 dwodx = -rd.rx_d - wo, dwody = -rd.ry_d - wo;
                const float dDNdx = dot(dwodx, ns) + dot(wo, dndx), dDNdy = dot(dwody, ns) + dot(wo, dndy);
                const V3 rx_o = si.p + st.dpdx, ry_o = si.p + st.dpdy;
                V3 rx_d, ry_d;
                if (pass == 0) {
                  rx_d = wi - dwodx + 2.0f * (dot(wo, ns) * dndx + dDNdx * ns);
                  ry_d = wi - dwody + 2.0f * (dot(wo, ns) * dndy + dDNdy * ns);
                } else {
                  float eta = bsdf.eta;
                  const V3 w = -wo;
                  if (dot(wo, ns) < 0.0f) eta = 1.0f / eta;
                  const float mu = eta * dot(w, ns) - dot(wi, ns);
                  rx_d = wi + eta * dwodx - (mu * dndx + dDNdx * ns);
                  ry_d = wi + eta * dwody - (mu * dndy + dDNdy * ns);
                }
                float4* q = (parity ? p.w.rdiff : p.w.rdiff2) + 3 * (size_t)pos;
                q[0] = make_float4(rx_o.x, rx_o.y, rx_o.z, ry_o.x); q[1] = make_float4(ry_o.y, ry_o.z, rx_d.x, rx_d.y); q[2] = make_float4(rx_d.z, ry_d.x, ry_d.y, ry_d.z);
              }
            } else p.w.counters[C_OVERFLOW] = 1;
          }
        }
      }
    }
    if (!is_black(colour)) { const Spec c = beta * colour; atomicAdd(&L->x, c.r); atomicAdd(&L->y, c.g); atomicAdd(&L->z, c.b); }
  }
}

// ---- AmbientOcclusion::li (integrator/ao.rs:32-58) and Normal::li (normal.rs:20-34) ---------------------------
__global__ void __launch_bounds__(128) k_shade_ao(RenderParams p) {
  const uint32_t n = min(p.w.counters[C_LIVE0], p.w.cap_items);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = p.w.list[0][i];
    const HitRec h = p.w.hit[slot];
    if (h.slot == kMiss) continue;
    Ray ray = load_ray(p.w.ray_o, p.w.ray_d, slot, nullptr);
    ray.t_max = inf_f();
    SurfHit si;
    hit_surface_bary(p.sc, h.slot, p.w.hit_inst ? p.w.hit_inst[slot] : kNoInst, ray, p.hit_t_is_b0 != 0, h.t, h.b1, h.b2, si);
    const Inter it = inter_of(si);
    const uint32_t sample = p.w.pstate[slot].x;
    const uint2 sinf = p.w.sinfo[sample];
    SamplerState ss; ss.ph = sinf.x; ss.s = sinf.y; ss.d1 = 1; ss.d2 = 2; ss.da = 0;
    if (p.integrator == RTGPU_INTEGRATOR_NORMAL) {
      const float v = fabsf(dot(ray.d, si.n));
      float4 L = p.w.L[sample]; L.x = v; L.y = v; L.z = v; p.w.L[sample] = L;
      continue;
    }
    for (int k = 0; k < p.ao_samples; k++) {
      V3 w = uniform_sample_sphere(ss.get_2d(p.scfg));
      if (dot(w, si.n) < 0.0f) w = -w;
      push_shadow(p, spawn_ray(it, w), sample, spec(1.0f));              // counts clear rays; k_film_add divides by n_samples
    }
  }
}

}  // namespace rt
