// Device functions shared by the wavefront kernels (product code): camera ray, light-distribution lookup and
// next-event estimation (queues the shadow and MIS rays).
#pragma once
#include "wave.cuh"

namespace rt {

constexpr uint32_t kNoLight = 0xffffffffu;
constexpr uint32_t kBsdfNonSpecular = BSDF_ALL & ~BSDF_SPECULAR;

// ---- camera (camera.rs:131-202, differential-free part) ---------------------------------------------------
RT_DEV Ray camera_ray(const float* r2c, const float* c2w, float lens_radius, float focal_distance, P2 p_film, P2 p_lens) {
  V3 p_camera = xf_point(r2c, v3(p_film.x, p_film.y, 0.0f));
  Ray ray = make_ray(v3(0, 0, 0), normalize(p_camera), inf_f());
  if (lens_radius > 0.0f) {
    P2 d = concentric_sample_disk(p_lens);
    P2 pl = mk2(lens_radius * d.x, lens_radius * d.y);
    float ft = focal_distance / ray.d.z;
    V3 p_focus = ray_at(ray, ft);
    ray.o = v3(pl.x, pl.y, 0.0f);
    ray.d = normalize(p_focus - ray.o);
  }
  V3 oe, de;
  return ray_transform(ray, c2w, oe, de);
}

// ---- light distributions (lightdistrib.rs) ---------------------------------------------------------------------
struct Distrib { const float* func; const float* cdf; float func_int; int n; };
// voxel of a point (SpatialLightDistribution::lookup, lightdistrib.rs:183-201)
RT_DEV size_t grid_voxel(const DScene& sc, const LightGrid& g, V3 pt) {
  const float* lo = sc.world_lo; const float* hi = sc.world_hi;
  float ox = pt.x - lo[0], oy = pt.y - lo[1], oz = pt.z - lo[2];      // Bounds3::offset (bounds.rs:177-190)
  if (hi[0] > lo[0]) ox /= hi[0] - lo[0];
  if (hi[1] > lo[1]) oy /= hi[1] - lo[1];
  if (hi[2] > lo[2]) oz /= hi[2] - lo[2];
  const int px = min(max(f2i32(ox * (float)g.nv[0]), 0), g.nv[0] - 1);
  const int py = min(max(f2i32(oy * (float)g.nv[1]), 0), g.nv[1] - 1);
  const int pz = min(max(f2i32(oz * (float)g.nv[2]), 0), g.nv[2] - 1);
  return ((size_t)px * g.nv[1] + py) * g.nv[2] + pz;
}
RT_DEV Distrib lookup_distrib(const RenderParams& p, V3 pt) {
  const LightGrid& g = p.grid;
  size_t voxel = 0;
  if (g.nv[0] > 0) {
    voxel = grid_voxel(p.sc, g, pt);
    if (g.slots) voxel = (size_t)max(g.slots[voxel], 0);             // sparse mode: the row k_lightgrid_mark claimed for this voxel before the bounce was shaded
  }
  const int n = g.n_lights;
  const float* base = g.table + voxel * (size_t)(2 * n + 2);
  Distrib d; d.func = base; d.cdf = base + n; d.func_int = base[2 * n + 1]; d.n = n;
  return d;
}

// ---- next-event estimation ------------------------------------------------------------------------------------
RT_DEV Inter inter_of(const SurfHit& si) { Inter it; it.p = si.p; it.p_error = si.p_error; it.n = si.n; return it; }

RT_DEV void push_shadow(const RenderParams& p, const Ray& ray, uint32_t sample, Spec c) {
  const uint32_t pos = warp_append(&p.w.counters[C_SHADOW], true);
  if (pos < p.w.cap_shadow) { store_ray(p.w.sh_o, p.w.sh_d, pos, ray, sample); st_stream(&p.w.sh_c[pos], make_float4(c.r, c.g, c.b, 0.0f)); }
  else p.w.counters[C_OVERFLOW] = 1;
}
RT_DEV void push_mis(const RenderParams& p, const Ray& ray, uint32_t sample, Spec c, uint32_t light_row) {
  const uint32_t pos = warp_append(&p.w.counters[C_MIS], true);
  if (pos < p.w.cap_mis) { store_ray(p.w.mi_o, p.w.mi_d, pos, ray, sample); st_stream(&p.w.mi_c[pos], make_float4(c.r, c.g, c.b, __uint_as_float(light_row))); }
  else p.w.counters[C_OVERFLOW] = 1;
}

// estimate_direct (integrator/mod.rs:222-318), deferred: instead of tracing, it queues the shadow ray with the
// light-sampling term and the MIS ray with the BSDF-sampling weight, both pre-multiplied by `scale`
// (= path throughput / light-choice pdf / sample count).
RT_DEV void estimate_direct(const RenderParams& p, const SurfHit& si, const Bsdf& bsdf, P2 u_scattering, uint32_t light_row, P2 u_light,
                            Spec scale, uint32_t sample) {
  const rtgpu_light& light = p.sc.lights[light_row];
  const Inter it = inter_of(si);
  V3 wi; float light_pdf; Inter p1;
  Spec li = light_sample_li(p.sc, light, it, u_light, wi, light_pdf, p1);
  if (light_pdf > 0.0f && !is_black(li)) {
    Spec f = bsdf_f(bsdf, si.wo, wi, kBsdfNonSpecular) * fabsf(dot(wi, si.ns));
    float scattering_pdf = bsdf_pdf(bsdf, si.wo, wi, kBsdfNonSpecular);
    if (!is_black(f)) {
      Spec c;
      if (light_is_delta(light)) c = f * li / light_pdf;
      else { float weight = power_heuristic(light_pdf, scattering_pdf); c = f * li * weight / light_pdf; }
      push_shadow(p, spawn_ray_to(it, p1), sample, scale * c);
    }
  }
  if (!light_is_delta(light)) {
    Spec f; V3 wi2; float scattering_pdf; uint32_t sampled;
    bsdf_sample_f(bsdf, si.wo, u_scattering, kBsdfNonSpecular, f, wi2, scattering_pdf, sampled);
    f = f * fabsf(dot(wi2, si.ns));
    const bool sampled_specular = (sampled & BSDF_SPECULAR) != 0;
    if (!is_black(f) && scattering_pdf > 0.0f) {
      float weight = 1.0f;
      if (!sampled_specular) {
        float lp = light_pdf_li(p.sc, light, it, wi2);
        if (lp == 0.0f) return;
        weight = power_heuristic(scattering_pdf, lp);
      }
      const Spec w = scale * (f * weight / scattering_pdf);
      if (light.kind == RTGPU_LIGHT_INFINITE) {
        // The MIS ray towards an infinite light contributes light.le(ray) exactly when its closest-hit query finds
        // nothing (a hit can never carry the infinite light's id, integrator/mod.rs:293-310) — i.e. when an any-hit
        // query finds nothing.  Same ray, same result, cheaper walk: queue it with the shadow rays, radiance folded in.
        // It stays a "regular" ray in the reference's counters (C_MIS_ANY).
        const Spec le = light_le(p.sc, light, wi2);
        if (!is_black(le)) {
          const uint32_t pos = warp_append(&p.w.counters[C_MIS_ANY], true);
          if (pos < p.w.cap_mis) {
            store_ray(p.w.ma_o, p.w.ma_d, pos, spawn_ray(it, wi2), sample);
            st_stream(&p.w.ma_c[pos], make_float4(w.r * le.r, w.g * le.g, w.b * le.b, 0.0f));
          } else p.w.counters[C_OVERFLOW] = 1;
        } else warp_append(&p.w.counters[C_MIS_SKIPPED], true);       // traced by the reference, contributes nothing
      } else push_mis(p, spawn_ray(it, wi2), sample, w, light_row);
    }
  }
}

}  // namespace rt
