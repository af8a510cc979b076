// Material -> BxDF list (product code, compiled for the host flattener and for the device shading kernels).
//
// `Material::compute_scattering_functions` of rustracer's materials (material/{matte,plastic,metal,glass,mirror,uber,
// substrate,translucent,mixmat}.rs) once the parameter textures have been evaluated: which BxDFs are added, in which order,
// with which constants.  With constant textures the list is a per-material constant and the host builds it once
// (scene_build.cpp); with image / procedural textures or a bump map the device builds it per hit from the evaluated
// parameters (texture.cuh).  One implementation, so both paths list identical lobes.
// Reference line numbers are relative to rustracer-core/src/.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../../include/rt_scene.h"
#include "../../../include/rtgpu.h"

#ifdef __CUDACC__
#define RTML_HD __host__ __device__ __forceinline__
#define RTML_LIST __host__ __device__ __noinline__      /* one function per mix depth: inlined, the nested listings multiply the code */
#else
#define RTML_HD inline
#define RTML_LIST inline
#endif

namespace rtml {

constexpr int kMaxLobes = 8;          // BxDFHolder capacity (bsdf/mod.rs:41-46)
constexpr int kMaxMixDepth = 2;       // ScaledBxDF wrappers a lobe row can carry (rtgpu_lobe.scale)
enum { kOk = 0, kTooManyLobes = 1, kMixTooDeep = 2, kNoBsdf = 3 };

struct LobeList { rtgpu_lobe* out; int n; int error; };

struct Rgb3 { float v[3]; };
RTML_HD Rgb3 rgb(const float* c) { Rgb3 r; r.v[0] = c[0]; r.v[1] = c[1]; r.v[2] = c[2]; return r; }
RTML_HD float clamp0(float x) { const float hi = INFINITY; return x < 0.0f ? 0.0f : (x > hi ? hi : x); }   // lib.rs clamp(v, 0, inf): NaN stays NaN
RTML_HD Rgb3 clamp_rgb(Rgb3 c) { c.v[0] = clamp0(c.v[0]); c.v[1] = clamp0(c.v[1]); c.v[2] = clamp0(c.v[2]); return c; }   // Spectrum::clamp (spectrum.rs:156-162)
RTML_HD Rgb3 mul(Rgb3 a, Rgb3 b) { Rgb3 r; r.v[0] = a.v[0] * b.v[0]; r.v[1] = a.v[1] * b.v[1]; r.v[2] = a.v[2] * b.v[2]; return r; }
RTML_HD Rgb3 one_minus(Rgb3 a) { Rgb3 r; r.v[0] = 1.0f - a.v[0]; r.v[1] = 1.0f - a.v[1]; r.v[2] = 1.0f - a.v[2]; return r; }
RTML_HD bool black(Rgb3 a) { return a.v[0] == 0.0f && a.v[1] == 0.0f && a.v[2] == 0.0f; }
RTML_HD void put3(float* d, Rgb3 c) { d[0] = c.v[0]; d[1] = c.v[1]; d[2] = c.v[2]; }
RTML_HD float clampf01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

RTML_HD float roughness_to_alpha(float roughness) {                   // bsdf/microfacet.rs:485-493
  roughness = fmaxf(roughness, 1e-3f);
  float x = logf(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

RTML_HD rtgpu_lobe new_lobe(uint32_t kind) {
  rtgpu_lobe l;
  l.kind = kind; l.n_scales = 0;
  for (int i = 0; i < 3; i++) { l.scale[0][i] = 0.0f; l.scale[1][i] = 0.0f; l.r[i] = 0.0f; l.t[i] = 0.0f; l.c_eta_t[i] = 1.0f; l.c_k[i] = 0.0f; }
  l.on_a = 0.0f; l.on_b = 0.0f; l.fr_kind = 0; l.fr_eta_i = 1.0f; l.fr_eta_t = 1.0f; l.ax = 0.0f; l.ay = 0.0f; l.eta_a = 1.0f; l.eta_b = 1.0f;
  return l;
}
RTML_HD void set_dielectric(rtgpu_lobe& l, float eta_i, float eta_t) { l.fr_kind = 1; l.fr_eta_i = eta_i; l.fr_eta_t = eta_t; }
RTML_HD void push(LobeList& L, const rtgpu_lobe& l) {
  if (L.n >= kMaxLobes) { L.error = kTooManyLobes; return; }         // the reference's BxDFHolder indexes out of bounds here
  L.out[L.n++] = l;
}

// Appends the BxDFs material `m` adds, in the reference's order; returns Bsdf::eta.  `m` has its parameters evaluated.
// `children(row, first)` returns the evaluated material of a mix child (first = mat1, whose bump map stays on the caller's
// surface; mat2 works on a clone, mixmat.rs:43-47).
template <int DEPTH, class Children>
RTML_LIST float list_lobes(const rt_material& m, bool allow_multiple_lobes, Children& children, LobeList& L) {
  switch (m.type) {
    case RT_MAT_MATTE: {                                              // matte.rs:37-62 ; oren_nayar.rs:17-26 (sigma in degrees)
      const Rgb3 r = clamp_rgb(rgb(m.kd));
      const float sigma = clampf01(m.sigma);
      if (!black(r)) {
        rtgpu_lobe l = new_lobe(sigma == 0.0f ? RTGPU_LOBE_LAMBERT_R : RTGPU_LOBE_OREN_NAYAR); put3(l.r, r);
        if (sigma != 0.0f) {
          const float sr = sigma * (3.14159265358979323846f / 180.0f), s2 = sr * sr;
          l.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
          l.on_b = 0.45f * s2 / (s2 + 0.09f);
        }
        push(L, l);
      }
      return 1.0f;
    }
    case RT_MAT_PLASTIC: {                                            // plastic.rs:45-74 (no clamp)
      const Rgb3 kd = rgb(m.kd), ks = rgb(m.ks);
      if (!black(kd)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_LAMBERT_R); put3(l.r, kd); push(L, l); }
      if (!black(ks)) {
        float rough = m.roughness;
        if (m.remap_roughness) rough = roughness_to_alpha(rough);
        rtgpu_lobe l = new_lobe(RTGPU_LOBE_MICRO_REFL); put3(l.r, ks); set_dielectric(l, 1.5f, 1.0f); l.ax = rough; l.ay = rough; push(L, l);
      }
      return 1.0f;
    }
    case RT_MAT_METAL: {                                              // metal.rs:50-81
      float ur = m.has_uroughness ? m.uroughness : m.roughness, vr = m.has_vroughness ? m.vroughness : m.roughness;
      if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
      rtgpu_lobe l = new_lobe(RTGPU_LOBE_MICRO_REFL); l.r[0] = l.r[1] = l.r[2] = 1.0f; l.fr_kind = 2; put3(l.c_eta_t, rgb(m.eta_rgb)); put3(l.c_k, rgb(m.k_rgb));
      l.ax = ur; l.ay = vr; push(L, l);
      return 1.0f;
    }
    case RT_MAT_GLASS: {                                              // glass.rs:53-106
      const float eta = m.eta;
      float ur = m.uroughness, vr = m.vroughness;
      const Rgb3 r = rgb(m.kr), t = rgb(m.kt);
      if (!black(r) || !black(t)) {
        const bool is_specular = ur == 0.0f && vr == 0.0f;
        if (is_specular && allow_multiple_lobes) {
          rtgpu_lobe l = new_lobe(RTGPU_LOBE_FRESNEL_SPEC); put3(l.r, r); put3(l.t, t); l.eta_a = 1.0f; l.eta_b = eta; push(L, l);
        } else {
          if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
          if (!black(r)) {
            rtgpu_lobe l = new_lobe(is_specular ? RTGPU_LOBE_SPEC_REFL : RTGPU_LOBE_MICRO_REFL); put3(l.r, r); set_dielectric(l, 1.0f, eta);
            l.ax = ur; l.ay = vr; push(L, l);
          }
          if (!black(t)) {
            rtgpu_lobe l = new_lobe(is_specular ? RTGPU_LOBE_SPEC_TRANS : RTGPU_LOBE_MICRO_TRANS);
            put3(l.t, is_specular ? t : r);                           // the rough transmission lobe is built with Kr (glass.rs:97)
            l.eta_a = 1.0f; l.eta_b = eta; set_dielectric(l, 1.0f, eta); l.ax = ur; l.ay = vr; push(L, l);
          }
        }
      }
      return eta;
    }
    case RT_MAT_MIRROR: {                                             // mirror.rs:30-48
      const Rgb3 r = clamp_rgb(rgb(m.kr));
      if (!black(r)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_SPEC_REFL); put3(l.r, r); l.fr_kind = 0; push(L, l); }
      return 1.0f;
    }
    case RT_MAT_UBER: {                                               // uber.rs:62-125
      const float e = m.eta;
      const Rgb3 op = clamp_rgb(rgb(m.opacity)), t = clamp_rgb(one_minus(op));
      float eta = e;
      if (!black(t)) {
        eta = 1.0f;
        rtgpu_lobe l = new_lobe(RTGPU_LOBE_SPEC_TRANS); put3(l.t, t); l.eta_a = 1.0f; l.eta_b = 1.0f; set_dielectric(l, 1.0f, 1.0f); push(L, l);
      }
      const Rgb3 kd = mul(op, clamp_rgb(rgb(m.kd)));
      if (!black(kd)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_LAMBERT_R); put3(l.r, kd); push(L, l); }
      const Rgb3 ks = mul(op, clamp_rgb(rgb(m.ks)));
      if (!black(ks)) {
        float ru = m.has_uroughness ? m.uroughness : m.roughness, rv = m.has_vroughness ? m.vroughness : m.roughness;
        if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
        rtgpu_lobe l = new_lobe(RTGPU_LOBE_MICRO_REFL); put3(l.r, ks); set_dielectric(l, 1.0f, e); l.ax = ru; l.ay = rv; push(L, l);
      }
      const Rgb3 kr = mul(op, clamp_rgb(rgb(m.kr)));
      if (!black(kr)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_SPEC_REFL); put3(l.r, kr); set_dielectric(l, 1.0f, e); push(L, l); }
      const Rgb3 kt = mul(op, clamp_rgb(rgb(m.kt)));
      if (!black(kt)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_SPEC_TRANS); put3(l.t, kt); l.eta_a = 1.0f; l.eta_b = e; set_dielectric(l, 1.0f, e); push(L, l); }
      return eta;
    }
    case RT_MAT_SUBSTRATE: {                                          // substrate.rs:42-71
      const Rgb3 d = clamp_rgb(rgb(m.kd)), s = clamp_rgb(rgb(m.ks));
      if (!black(d) || !black(s)) {
        float ru = m.uroughness, rv = m.vroughness;
        if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
        rtgpu_lobe l = new_lobe(RTGPU_LOBE_FRESNEL_BLEND); put3(l.r, s); put3(l.t, d); l.ax = ru; l.ay = rv; push(L, l);   // FresnelBlend::new(rs, rd, ..)
      }
      return 1.0f;
    }
    case RT_MAT_TRANSLUCENT: {                                        // translucent.rs:48-101
      const float eta = 1.5f;
      const Rgb3 r = clamp_rgb(rgb(m.reflect)), t = clamp_rgb(rgb(m.transmit));
      if (!black(r) || !black(t)) {
        const Rgb3 kd = clamp_rgb(rgb(m.kd));
        if (!black(kd)) {
          if (!black(r)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_LAMBERT_R); put3(l.r, mul(r, kd)); push(L, l); }
          if (!black(t)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_LAMBERT_T); put3(l.t, mul(t, kd)); push(L, l); }
        }
        const Rgb3 ks = clamp_rgb(rgb(m.ks));
        if (!black(ks)) {
          float rough = m.roughness;
          if (m.remap_roughness) rough = roughness_to_alpha(rough);
          if (!black(r)) { rtgpu_lobe l = new_lobe(RTGPU_LOBE_MICRO_REFL); put3(l.r, mul(r, ks)); set_dielectric(l, 1.0f, eta); l.ax = l.ay = rough; push(L, l); }
          if (!black(t)) {
            rtgpu_lobe l = new_lobe(RTGPU_LOBE_MICRO_TRANS); put3(l.t, mul(t, ks)); l.eta_a = 1.0f; l.eta_b = eta; set_dielectric(l, 1.0f, eta); l.ax = l.ay = rough;
            push(L, l);
          }
        }
      }
      return eta;
    }
    case RT_MAT_MIX: {                                                // mixmat.rs:34-64: ScaledBxDF(b, s1) for mat1's, ScaledBxDF(b, s2) for mat2's
      if constexpr (DEPTH >= kMaxMixDepth) { L.error = kMixTooDeep; return 1.0f; }
      else {
        const Rgb3 s1 = clamp_rgb(rgb(m.amount)), s2 = clamp_rgb(one_minus(s1));
        const int first = L.n;
        const rt_material a = children(m.mix_a, true);
        const float eta = list_lobes<DEPTH + 1>(a, allow_multiple_lobes, children, L);   // the Bsdf object (eta, frame) stays mat1's
        const int mid = L.n;
        const rt_material b = children(m.mix_b, false);
        list_lobes<DEPTH + 1>(b, allow_multiple_lobes, children, L);
        for (int i = first; i < L.n; i++) {
          rtgpu_lobe& l = L.out[i];
          if (l.n_scales >= 2) { L.error = kMixTooDeep; continue; }
          put3(l.scale[l.n_scales++], i < mid ? s1 : s2);
        }
        return eta;
      }
    }
    default: L.error = kNoBsdf; return 1.0f;
  }
}

}  // namespace rtml
