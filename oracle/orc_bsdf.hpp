// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  BSDF, BxDFs, materials.
// Follows /root/reference/rustracer-core/src/bsdf/*.rs and material/{matte,plastic,metal,glass,mirror}.rs.
#pragma once
#include <stdexcept>
#include "orc_texture.hpp"
#include "../include/rt_scene.h"

namespace orc {

enum : uint32_t { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 }; // bsdf/mod.rs:24-32

// geometry/mod.rs:15-95 (local shading frame helpers)
inline float cos_theta(V3 w) { return w.z; }
inline float cos2_theta(V3 w) { return w.z * w.z; }
inline float abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline float sin2_theta(V3 w) { return fmax_(1.0f - cos2_theta(w), 0.0f); }
inline float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline float cos_phi(V3 w) { float st = sin_theta(w); return st == 0.0f ? 1.0f : clampv(w.x / st, -1.0f, 1.0f); }
inline float sin_phi(V3 w) { float st = sin_theta(w); return st == 0.0f ? 0.0f : clampv(w.y / st, -1.0f, 1.0f); }
inline float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
inline float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
inline float spherical_theta(V3 v) { return std::acos(clampv(v.z, -1.0f, 1.0f)); }
inline float spherical_phi(V3 v) { float p = std::atan2(v.y, v.x); return p < 0.0f ? p + 2.0f * PI : p; }

// bsdf/fresnel.rs:14-31
inline V3 reflect(V3 wo, V3 n) { return -wo + n * 2.0f * dot(wo, n); }
inline bool refract(V3 i, V3 n, float eta, V3& wt) {
  float cos_theta_i = dot(n, i);
  float sin2theta_i = fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f);
  float sin2theta_t = eta * eta * sin2theta_i;
  if (sin2theta_t >= 1.0f) return false;
  float cos_theta_t = std::sqrt(1.0f - sin2theta_t);
  wt = eta * -i + (eta * cos_theta_i - cos_theta_t) * n;
  return true;
}
// fresnel.rs:33-58
inline float fr_dielectric(float cos_theta_i, float eta_i, float eta_t) {
  cos_theta_i = clampv(cos_theta_i, -1.0f, 1.0f);
  if (cos_theta_i <= 0.0f) { std::swap(eta_i, eta_t); cos_theta_i = std::fabs(cos_theta_i); }
  float sin_theta_i = std::sqrt(fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f));
  float sin_theta_t = eta_i / eta_t * sin_theta_i;
  if (sin_theta_t >= 1.0f) return 1.0f;
  float cos_theta_t = std::sqrt(fmax_(1.0f - sin_theta_t * sin_theta_t, 0.0f));
  float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
  float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
  return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// fresnel.rs:60-82
inline Spectrum fr_conductor(float cos_theta_i, Spectrum eta_i, Spectrum eta_t, Spectrum k) {
  cos_theta_i = clampv(cos_theta_i, -1.0f, 1.0f);
  Spectrum eta = eta_t / eta_i, eta_k = k / eta_i;
  float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.0f - cos2;
  Spectrum eta2 = eta * eta, eta_k2 = eta_k * eta_k;
  Spectrum t0 = eta2 - eta_k2 - sin2;
  Spectrum a2plusb2 = (t0 * t0 + 4.0f * eta2 * eta_k2).sqrt();
  Spectrum t1 = a2plusb2 + cos2;
  Spectrum a = (0.5f * (a2plusb2 + t0)).sqrt();
  Spectrum t2 = 2.0f * cos_theta_i * a;
  Spectrum r_s = (t1 - t2) / (t1 + t2);
  Spectrum t3 = cos2 * a2plusb2 + sin2 * sin2;
  Spectrum t4 = t2 * sin2;
  Spectrum r_p = r_s * (t3 - t4) / (t3 + t4);
  return 0.5f * (r_p + r_s);
}

struct Fresnel {                                                 // fresnel.rs:84-138 (evaluate takes abs(cos) for both)
  int kind = 0;                 // 0 no-op, 1 dielectric, 2 conductor
  float eta_i = 1, eta_t = 1;
  Spectrum c_eta_i, c_eta_t, c_k;
  Spectrum evaluate(float cos_theta_i) const {
    if (kind == 1) return Spectrum(fr_dielectric(std::fabs(cos_theta_i), eta_i, eta_t));
    if (kind == 2) return fr_conductor(std::fabs(cos_theta_i), c_eta_i, c_eta_t, c_k);
    return Spectrum(1.0f);
  }
};

struct TrowbridgeReitz {                                         // microfacet.rs:469-650
  float ax = 0, ay = 0;
  static float roughness_to_alpha(float roughness) {             // :485-493
    roughness = fmax_(roughness, 1e-3f);
    float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
  }
  float d(V3 wh) const {                                         // :576-588
    float tan2 = tan2_theta(wh);
    if (std::isinf(tan2)) return 0.0f;
    float cos4 = cos2_theta(wh) * cos2_theta(wh);
    float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * tan2;
    return 1.0f / (PI * ax * ay * cos4 * (1.0f + e) * (1.0f + e));
  }
  float lambda(V3 w) const {                                     // :590-602
    float abs_tan = std::fabs(tan_theta(w));
    if (std::isinf(abs_tan)) return 0.0f;
    float alpha = std::sqrt(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
    return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
  }
  float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }                     // :235-237
  float g(V3 wi, V3 wo) const { return 1.0f / (1.0f + lambda(wi) + lambda(wo)); } // :239-241
  float pdf(V3 wo, V3 wh) const { return d(wh) * g1(wo) * std::fabs(dot(wo, wh)) / abs_cos_theta(wo); } // :243-249 (sample_visible_area = true)
  static void sample11(float cos_t, float u1, float u2, float& slope_x, float& slope_y) {   // :517-572
    if (cos_t > 0.9999f) {
      float r = std::sqrt(u1 / (1.0f - u1));
      float phi = 6.28318530717958647692f * u2;
      slope_x = r * std::cos(phi); slope_y = r * std::sin(phi);
      return;
    }
    float sin_t = std::sqrt(fmax_(1.0f - cos_t * cos_t, 0.0f));
    float tan_t = sin_t / cos_t;
    float a = 1.0f / tan_t;
    float G1 = 2.0f / (1.0f + std::sqrt(1.0f + 1.0f / (a * a)));
    float A = 2.0f * u1 / G1 - 1.0f;
    float tmp = 1.0f / (A * A - 1.0f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tan_t;
    float D = std::sqrt(fmax_(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.0f));
    float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
    slope_x = (A < 0.0f || slope_x_2 > 1.0f / tan_t) ? slope_x_1 : slope_x_2;
    float S;
    if (u2 > 0.5f) { S = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
    else { S = -1.0f; u2 = 2.0f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    slope_y = S * z * std::sqrt(1.0f + slope_x * slope_x);
  }
  V3 sample(V3 wi, float u1, float u2) const {                   // :495-515
    V3 wis = normalize(V3(ax * wi.x, ay * wi.y, wi.z));
    float sx, sy; sample11(cos_theta(wis), u1, u2, sx, sy);
    float tmp = cos_phi(wis) * sx - sin_phi(wis) * sy;
    sy = sin_phi(wis) * sx + cos_phi(wis) * sy;
    sx = tmp;
    sx *= ax; sy *= ay;
    return normalize(V3(-sx, -sy, 1.0f));
  }
  V3 sample_wh(V3 wo, P2 u) const {                              // :604-645 (visible-area branch)
    bool flip = wo.z < 0.0f;
    V3 w = flip ? -wo : wo;
    V3 wh = sample(w, u.x, u.y);
    if (flip) wh = -wh;
    return wh;
  }
};

enum LobeKind { LOBE_LAMBERT_R, LOBE_OREN_NAYAR, LOBE_SPEC_REFL, LOBE_SPEC_TRANS, LOBE_FRESNEL_SPEC, LOBE_MICRO_REFL, LOBE_MICRO_TRANS,
                LOBE_LAMBERT_T, LOBE_FRESNEL_BLEND };
inline float pow5(float v) { return (v * v) * (v * v) * v; }                    // fresnel.rs:414-417

struct Lobe {
  LobeKind kind;
  Spectrum r, t;
  float on_a = 0, on_b = 0;         // OrenNayar A, B
  Fresnel fresnel;
  TrowbridgeReitz dist;
  float eta_a = 1, eta_b = 1;
  // LOBE_FRESNEL_BLEND (fresnel.rs:335-412): rs in `r`, rd in `t`.
  // ScaledBxDF wrappers of MixMaterial (bxdf.rs:48-71), innermost first: f and sample_f are scaled, get_type is the
  // wrapped lobe's, and pdf() is NOT forwarded — it is the trait's default cosine pdf (bxdf.rs:38-44).
  Spectrum scale[2]; int n_scales = 0;
  uint32_t type() const {
    switch (kind) {
      case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: return BSDF_DIFFUSE | BSDF_REFLECTION;
      case LOBE_SPEC_REFL: return BSDF_SPECULAR | BSDF_REFLECTION;
      case LOBE_SPEC_TRANS: return BSDF_SPECULAR | BSDF_TRANSMISSION;
      case LOBE_FRESNEL_SPEC: return BSDF_SPECULAR | BSDF_REFLECTION | BSDF_TRANSMISSION;
      case LOBE_MICRO_REFL: case LOBE_FRESNEL_BLEND: return BSDF_REFLECTION | BSDF_GLOSSY;
      case LOBE_LAMBERT_T: return BSDF_DIFFUSE | BSDF_TRANSMISSION;           // lambertian.rs:43-45
      default: return BSDF_TRANSMISSION | BSDF_GLOSSY;
    }
  }
  bool matches(uint32_t flags) const { return (type() & flags) == type(); }   // bxdf.rs:29-31
  Spectrum f_inner(V3 wo, V3 wi) const {
    switch (kind) {
      case LOBE_LAMBERT_R: return r * INV_PI;                                 // lambertian.rs:19-21
      case LOBE_LAMBERT_T: return t * INV_PI;                                 // lambertian.rs:39-41
      case LOBE_FRESNEL_BLEND: {                                              // fresnel.rs:358-375 (rs = r, rd = t)
        Spectrum diffuse = (28.0f / (23.0f * PI)) * t * (Spectrum(1.0f) - r) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                           (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
        V3 wh = wi + wo;
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
        wh = normalize(wh);
        float cwh = dot(wi, wh);
        Spectrum schlick = r + pow5(1.0f - cwh) * (Spectrum(1.0f) - r);       // :352-354
        Spectrum specular = dist.d(wh) / (4.0f * std::fabs(dot(wi, wh)) * fmax_(abs_cos_theta(wi), abs_cos_theta(wo))) * schlick;
        return diffuse + specular;
      }
      case LOBE_OREN_NAYAR: {                                                 // oren_nayar.rs:30-52
        float sti = sin_theta(wi), sto = sin_theta(wo);
        float max_cos = 0.0f;
        if (sti > 1e-4f && sto > 1e-4f) {
          float d_cos = sin_phi(wi) * sin_phi(wo) + cos_phi(wi) * cos_phi(wo);
          max_cos = fmax_(d_cos, 0.0f);
        }
        float sin_alpha, tan_beta;
        if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sto; tan_beta = sti / abs_cos_theta(wi); }
        else { sin_alpha = sti; tan_beta = sto / abs_cos_theta(wo); }
        return r * INV_PI * (on_a + on_b * max_cos * sin_alpha * tan_beta);
      }
      case LOBE_MICRO_REFL: {                                                 // microfacet.rs:36-53
        float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
        V3 wh = wi + wo;
        if (cto == 0.0f || cti == 0.0f) return Spectrum(0.0f);
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
        wh = normalize(wh);
        Spectrum F = fresnel.evaluate(dot(wi, wh));
        return r * dist.d(wh) * dist.g(wo, wi) * F / (4.0f * cti * cto);
      }
      case LOBE_MICRO_TRANS: {                                                // microfacet.rs:125-169
        if (same_hemisphere(wo, wi)) return Spectrum(0.0f);
        float cto = cos_theta(wo), cti = cos_theta(wi);
        if (cto == 0.0f || cti == 0.0f) return Spectrum(0.0f);
        float eta = cto > 0.0f ? eta_b / eta_a : eta_a / eta_b;
        V3 wh = normalize(wo + wi * eta);
        if (wh.z < 0.0f) wh = -wh;
        Spectrum F = fresnel.evaluate(dot(wo, wh));
        float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        float factor = 1.0f / eta;   // TransportMode::RADIANCE on this path
        return (Spectrum(1.0f) - F) * t *
               std::fabs(dist.d(wh) * dist.g(wo, wi) * eta * eta * std::fabs(dot(wi, wh)) * std::fabs(dot(wo, wh)) * factor * factor /
                         (cti * cto * sqrt_denom * sqrt_denom));
      }
      default: return Spectrum(0.0f);                                         // specular lobes: fresnel.rs:153-157 etc.
    }
  }
  float pdf_inner(V3 wo, V3 wi) const {
    switch (kind) {
      case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: case LOBE_LAMBERT_T:         // bxdf.rs:38-44 (LambertianTransmission does not override it)
        return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
      case LOBE_FRESNEL_BLEND: {                                              // fresnel.rs:377-385
        if (!same_hemisphere(wo, wi)) return 0.0f;
        V3 wh = normalize(wo + wi);
        float pdf_wh = dist.pdf(wo, wh);
        return 0.5f * (abs_cos_theta(wi) * INV_PI + pdf_wh / (4.0f * dot(wo, wh)));
      }
      case LOBE_MICRO_REFL: {                                                 // microfacet.rs:87-94
        if (!same_hemisphere(wo, wi)) return 0.0f;
        V3 wh = normalize(wo + wi);
        return dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
      }
      case LOBE_MICRO_TRANS: {                                                // microfacet.rs:207-222
        if (same_hemisphere(wo, wi)) return 0.0f;
        float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
        V3 wh = normalize(wo + wi * eta);
        float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        float dwh_dwi = std::fabs((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
        return dist.pdf(wo, wh) * dwh_dwi;
      }
      default: return 0.0f;
    }
  }
  // returns sampled type (bxdf.rs:18-25: default returns EMPTY — Q17)
  void sample_f_inner(V3 wo, P2 u, Spectrum& f_out, V3& wi, float& pdf_out, uint32_t& sampled) const {
    switch (kind) {
      case LOBE_FRESNEL_BLEND: {                                              // fresnel.rs:387-407
        sampled = type();
        if (u.x < 0.5f) {
          u.x = fmin_(2.0f * u.x, ONE_MINUS_EPSILON);
          wi = cosine_sample_hemisphere(u);
          if (wo.z < 0.0f) wi.z *= -1.0f;
        } else {
          u.x = fmin_(2.0f * (u.x - 0.5f), ONE_MINUS_EPSILON);
          V3 wh = dist.sample_wh(wo, u);
          wi = reflect(wo, wh);
          if (!same_hemisphere(wo, wi)) { f_out = Spectrum(0.0f); pdf_out = 0.0f; return; }
        }
        f_out = f_inner(wo, wi); pdf_out = pdf_inner(wo, wi);
        return;
      }
      case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: case LOBE_LAMBERT_T: {       // bxdf.rs:18-25 (same hemisphere as wo, also for LambertianTransmission)
        wi = cosine_sample_hemisphere(u);
        if (wo.z < 0.0f) wi.z *= -1.0f;
        pdf_out = pdf_inner(wo, wi); f_out = f_inner(wo, wi); sampled = 0;
        return;
      }
      case LOBE_SPEC_REFL: {                                                  // fresnel.rs:159-164
        wi = V3(-wo.x, -wo.y, wo.z);
        f_out = fresnel.evaluate(cos_theta(wi)) * r / abs_cos_theta(wi);
        pdf_out = 1.0f; sampled = type();
        return;
      }
      case LOBE_SPEC_TRANS: {                                                 // fresnel.rs:203-230
        bool entering = cos_theta(wo) > 0.0f;
        float ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
        V3 w;
        if (refract(wo, face_forward(V3(0, 0, 1), wo), ei / et, w)) {
          wi = w;
          Spectrum ft = t * (Spectrum(1.0f) - fresnel.evaluate(cos_theta(wi)));
          ft = ft * (ei * ei) / (et * et);
          f_out = ft / abs_cos_theta(wi); pdf_out = 1.0f; sampled = type();
        } else { f_out = Spectrum(1.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; sampled = 0; }
        return;
      }
      case LOBE_FRESNEL_SPEC: {                                               // fresnel.rs:273-322
        float fr = fr_dielectric(cos_theta(wo), eta_a, eta_b);
        if (u.x < fr) {
          wi = V3(-wo.x, -wo.y, wo.z);
          f_out = fr * r / abs_cos_theta(wi); pdf_out = fr; sampled = BSDF_SPECULAR | BSDF_REFLECTION;
        } else {
          bool entering = cos_theta(wo) > 0.0f;
          float ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
          V3 w;
          if (refract(wo, face_forward(V3(0, 0, 1), wo), ei / et, w)) {
            wi = w;
            Spectrum ft = t * (1.0f - fr);
            ft = ft * ((ei * ei) / (et * et));
            f_out = ft / abs_cos_theta(wi); pdf_out = 1.0f - fr; sampled = BSDF_SPECULAR | BSDF_TRANSMISSION;
          } else { f_out = Spectrum(0.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; sampled = 0; }
        }
        return;
      }
      case LOBE_MICRO_REFL: {                                                 // microfacet.rs:61-85
        sampled = type();
        if (wo.z == 0.0f) { f_out = Spectrum(0.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; return; }
        V3 wh = dist.sample_wh(wo, u);
        wi = reflect(wo, wh);
        if (!same_hemisphere(wo, wi)) { f_out = Spectrum(0.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; return; }
        pdf_out = dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
        f_out = f_inner(wo, wi);
        return;
      }
      default: {                                                              // LOBE_MICRO_TRANS microfacet.rs:177-205
        sampled = type();
        if (wo.z == 0.0f) { f_out = Spectrum(0.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; return; }
        V3 wh = dist.sample_wh(wo, u);
        float eta = cos_theta(wo) > 0.0f ? eta_a / eta_b : eta_b / eta_a;
        V3 w;
        if (refract(wo, wh, eta, w)) { wi = w; pdf_out = pdf_inner(wo, wi); f_out = f_inner(wo, wi); }
        else { f_out = Spectrum(0.0f); wi = V3(0, 0, 0); pdf_out = 0.0f; }
        return;
      }
    }
  }
  // ---- the BxDF interface as Bsdf sees it: the lobe itself, or the lobe behind its ScaledBxDF wrappers (bxdf.rs:48-71) ----
  Spectrum f(V3 wo, V3 wi) const { Spectrum v = f_inner(wo, wi); for (int i = 0; i < n_scales; i++) v = v * scale[i]; return v; }
  float pdf(V3 wo, V3 wi) const {
    if (n_scales > 0) return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
    return pdf_inner(wo, wi);
  }
  void sample_f(V3 wo, P2 u, Spectrum& f_out, V3& wi, float& pdf_out, uint32_t& sampled) const {
    sample_f_inner(wo, u, f_out, wi, pdf_out, sampled);
    for (int i = 0; i < n_scales; i++) f_out = f_out * scale[i];
  }
};

struct Bsdf {                                                    // bsdf/mod.rs:64-269
  float eta = 1.0f;
  V3 ns, ng, ss, ts;
  Lobe lobes[8]; int n = 0;
  void init(const SurfaceInteraction& isect, float eta_) {       // :77-92
    eta = eta_;
    ss = normalize(isect.shading.dpdu);
    ns = isect.shading.n; ng = isect.hit.n;
    ts = cross(isect.shading.n, ss);
  }
  void add(const Lobe& l) { lobes[n++] = l; }
  V3 world_to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }   // :253-255
  V3 local_to_world(V3 v) const {                                                   // :257-263
    return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
  }
  int num_components(uint32_t flags) const { int c = 0; for (int i = 0; i < n; i++) if (lobes[i].matches(flags)) c++; return c; } // :265-268
  Spectrum f(V3 wo_w, V3 wi_w, uint32_t flags) const {           // :94-112
    V3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
    if (wo.z == 0.0f) return Spectrum(0.0f);
    bool refl = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
    Spectrum c(0.0f);
    for (int i = 0; i < n; i++) {
      const Lobe& b = lobes[i];
      if (b.matches(flags) && ((refl && (b.type() & BSDF_REFLECTION)) || (!refl && (b.type() & BSDF_TRANSMISSION)))) c = c + b.f(wo, wi);
    }
    return c;
  }
  float pdf(V3 wo_w, V3 wi_w, uint32_t flags) const {            // :114-136
    if (n == 0) return 0.0f;
    V3 wo = world_to_local(wo_w);
    if (wo.z == 0.0f) return 0.0f;
    V3 wi = world_to_local(wi_w);
    int matched = 0; float p = 0.0f;
    for (int i = 0; i < n; i++) if (lobes[i].matches(flags)) { matched++; p += lobes[i].pdf(wo, wi); }
    return matched == 0 ? 0.0f : p / (float)matched;
  }
  void sample_f(V3 wo_w, P2 u, uint32_t flags, Spectrum& f_out, V3& wi_w, float& pdf_out, uint32_t& sampled) const {   // :138-251
    const Lobe* m[8]; int nm = 0;
    for (int i = 0; i < n; i++) if (lobes[i].matches(flags)) m[nm++] = &lobes[i];
    if (nm == 0) { f_out = Spectrum(0.0f); wi_w = V3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
    int comp = (int)pmin<int64_t>(f2usize(std::floor(u.x * (float)nm)), nm - 1);
    const Lobe* bxdf = m[comp];
    P2 ur(fmin_(u.x * (float)nm - (float)comp, ONE_MINUS_EPSILON), u.y);
    V3 wo = world_to_local(wo_w);
    if (wo.z == 0.0f) { f_out = Spectrum(0.0f); wi_w = V3(0, 0, 0); pdf_out = 0.0f; sampled = bxdf->type(); return; }
    Spectrum f; V3 wi; float pdf;
    bxdf->sample_f(wo, ur, f, wi, pdf, sampled);
    if (pdf == 0.0f) { f_out = Spectrum(0.0f); wi_w = V3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
    wi_w = local_to_world(wi);
    if (!(bxdf->type() & BSDF_SPECULAR) && nm > 1)
      for (int i = 0; i < nm; i++) if (i != comp) pdf += m[i]->pdf(wo, wi);
    if (nm > 1) pdf /= (float)nm;
    if (!(bxdf->type() & BSDF_SPECULAR)) {
      bool refl = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
      f = Spectrum(0.0f);
      for (int i = 0; i < nm; i++)
        if ((refl && (m[i]->type() & BSDF_REFLECTION)) || (!refl && (m[i]->type() & BSDF_TRANSMISSION))) f = f + m[i]->f(wo, wi);
    }
    f_out = f; pdf_out = pdf;
  }
};

// material/*.rs with every texture constant.  `allow_multiple_lobes`: path=true (path.rs:145), whitted/direct=false.
// `table`: the scene's material rows (MixMaterial refers to its two children by row).
// `ts`: the scene's textures; a material with textured parameters is evaluated at `si` first (and its bump map applied to `si`).
inline bool compute_scattering_functions(const rt_material* table, const rt_material& mt_in, SurfaceInteraction& si, bool allow_multiple_lobes, Bsdf& bsdf,
                                         const TextureSet* ts = nullptr) {
  bsdf.n = 0;
  const rt_material mt = (ts && mt_in.textured) ? ts->resolve(mt_in, si) : mt_in;
  auto S = [](const float* c) { return Spectrum(c[0], c[1], c[2]); };
  switch (mt.type) {
    case RT_MAT_MATTE: {                                          // matte.rs:37-62
      Spectrum r = S(mt.kd).clamp0();
      float sigma = clampv(mt.sigma, 0.0f, 1.0f);
      bsdf.init(si, 1.0f);
      if (!r.is_black()) {
        Lobe l;
        if (sigma == 0.0f) { l.kind = LOBE_LAMBERT_R; l.r = r; }
        else {                                                    // oren_nayar.rs:17-26 (sigma in DEGREES — Q32)
          l.kind = LOBE_OREN_NAYAR; l.r = r;
          float sr = to_radians(sigma), s2 = sr * sr;
          l.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
          l.on_b = 0.45f * s2 / (s2 + 0.09f);
        }
        bsdf.add(l);
      }
      return true;
    }
    case RT_MAT_PLASTIC: {                                        // plastic.rs:45-74 (no clamp — Q19)
      Spectrum kd = S(mt.kd), ks = S(mt.ks);
      bsdf.init(si, 1.0f);
      if (!kd.is_black()) { Lobe l; l.kind = LOBE_LAMBERT_R; l.r = kd; bsdf.add(l); }
      if (!ks.is_black()) {
        Lobe l; l.kind = LOBE_MICRO_REFL; l.r = ks;
        l.fresnel.kind = 1; l.fresnel.eta_i = 1.5f; l.fresnel.eta_t = 1.0f;
        float rough = mt.roughness;
        if (mt.remap_roughness) rough = TrowbridgeReitz::roughness_to_alpha(rough);
        l.dist.ax = rough; l.dist.ay = rough;
        bsdf.add(l);
      }
      return true;
    }
    case RT_MAT_METAL: {                                          // metal.rs:50-81
      float ur = mt.has_uroughness ? mt.uroughness : mt.roughness;
      float vr = mt.has_vroughness ? mt.vroughness : mt.roughness;
      if (mt.remap_roughness) { ur = TrowbridgeReitz::roughness_to_alpha(ur); vr = TrowbridgeReitz::roughness_to_alpha(vr); }
      Lobe l; l.kind = LOBE_MICRO_REFL; l.r = Spectrum(1.0f);
      l.fresnel.kind = 2; l.fresnel.c_eta_i = Spectrum(1.0f); l.fresnel.c_eta_t = S(mt.eta_rgb); l.fresnel.c_k = S(mt.k_rgb);
      l.dist.ax = ur; l.dist.ay = vr;
      bsdf.init(si, 1.0f);
      bsdf.add(l);
      return true;
    }
    case RT_MAT_GLASS: {                                          // glass.rs:53-106
      float eta = mt.eta, u_rough = mt.uroughness, v_rough = mt.vroughness;
      Spectrum r = S(mt.kr), t = S(mt.kt);
      bsdf.init(si, eta);
      if (!r.is_black() || !t.is_black()) {
        bool is_specular = u_rough == 0.0f && v_rough == 0.0f;
        if (is_specular && allow_multiple_lobes) {
          Lobe l; l.kind = LOBE_FRESNEL_SPEC; l.r = r; l.t = t; l.eta_a = 1.0f; l.eta_b = eta; bsdf.add(l);
        } else {
          if (mt.remap_roughness) { u_rough = TrowbridgeReitz::roughness_to_alpha(u_rough); v_rough = TrowbridgeReitz::roughness_to_alpha(v_rough); }
          if (!r.is_black()) {
            Lobe l; l.r = r; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = eta;
            if (is_specular) l.kind = LOBE_SPEC_REFL; else { l.kind = LOBE_MICRO_REFL; l.dist.ax = u_rough; l.dist.ay = v_rough; }
            bsdf.add(l);
          }
          if (!t.is_black()) {
            Lobe l; l.eta_a = 1.0f; l.eta_b = eta; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = eta;
            if (is_specular) { l.kind = LOBE_SPEC_TRANS; l.t = t; }
            else { l.kind = LOBE_MICRO_TRANS; l.t = r; l.dist.ax = u_rough; l.dist.ay = v_rough; }   // built with `r` — glass.rs:97, Q18
            bsdf.add(l);
          }
        }
      }
      return true;
    }
    case RT_MAT_MIRROR: {                                         // mirror.rs:30-48
      Spectrum R = S(mt.kr).clamp0();
      bsdf.init(si, 1.0f);
      if (!R.is_black()) { Lobe l; l.kind = LOBE_SPEC_REFL; l.r = R; l.fresnel.kind = 0; bsdf.add(l); }
      return true;
    }
    case RT_MAT_UBER: {                                           // uber.rs:62-125
      float e = mt.eta;
      Spectrum op = S(mt.opacity).clamp0();
      Spectrum t = (Spectrum(1.0f) - op).clamp0();
      float eta = e;
      if (!t.is_black()) {
        eta = 1.0f;
        Lobe l; l.kind = LOBE_SPEC_TRANS; l.t = t; l.eta_a = 1.0f; l.eta_b = 1.0f; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = 1.0f;
        bsdf.add(l);
      }
      Spectrum kd = op * S(mt.kd).clamp0();
      if (!kd.is_black()) { Lobe l; l.kind = LOBE_LAMBERT_R; l.r = kd; bsdf.add(l); }
      Spectrum ks = op * S(mt.ks).clamp0();
      if (!ks.is_black()) {
        float roughu = mt.has_uroughness ? mt.uroughness : mt.roughness, roughv = mt.has_vroughness ? mt.vroughness : mt.roughness;
        if (mt.remap_roughness) { roughu = TrowbridgeReitz::roughness_to_alpha(roughu); roughv = TrowbridgeReitz::roughness_to_alpha(roughv); }
        Lobe l; l.kind = LOBE_MICRO_REFL; l.r = ks; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = e; l.dist.ax = roughu; l.dist.ay = roughv;
        bsdf.add(l);
      }
      Spectrum kr = op * S(mt.kr).clamp0();
      if (!kr.is_black()) { Lobe l; l.kind = LOBE_SPEC_REFL; l.r = kr; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = e; bsdf.add(l); }
      Spectrum kt = op * S(mt.kt).clamp0();
      if (!kt.is_black()) {
        Lobe l; l.kind = LOBE_SPEC_TRANS; l.t = kt; l.eta_a = 1.0f; l.eta_b = e; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = e;
        bsdf.add(l);
      }
      bsdf.init(si, eta);
      return true;
    }
    case RT_MAT_SUBSTRATE: {                                      // substrate.rs:42-71
      Spectrum d = S(mt.kd).clamp0(), sp = S(mt.ks).clamp0();
      float roughu = mt.uroughness, roughv = mt.vroughness;
      if (!d.is_black() || !sp.is_black()) {
        if (mt.remap_roughness) { roughu = TrowbridgeReitz::roughness_to_alpha(roughu); roughv = TrowbridgeReitz::roughness_to_alpha(roughv); }
        Lobe l; l.kind = LOBE_FRESNEL_BLEND; l.r = sp; l.t = d; l.dist.ax = roughu; l.dist.ay = roughv;
        bsdf.add(l);
      }
      bsdf.init(si, 1.0f);
      return true;
    }
    case RT_MAT_TRANSLUCENT: {                                    // translucent.rs:48-101
      const float eta = 1.5f;
      Spectrum r = S(mt.reflect).clamp0(), t = S(mt.transmit).clamp0();
      if (!r.is_black() || !t.is_black()) {
        Spectrum kd = S(mt.kd).clamp0();
        if (!kd.is_black()) {
          if (!r.is_black()) { Lobe l; l.kind = LOBE_LAMBERT_R; l.r = r * kd; bsdf.add(l); }
          if (!t.is_black()) { Lobe l; l.kind = LOBE_LAMBERT_T; l.t = t * kd; bsdf.add(l); }
        }
        Spectrum ks = S(mt.ks).clamp0();
        if (!ks.is_black() && (!r.is_black() || !t.is_black())) {
          float rough = mt.roughness;
          if (mt.remap_roughness) rough = TrowbridgeReitz::roughness_to_alpha(rough);
          if (!r.is_black()) {
            Lobe l; l.kind = LOBE_MICRO_REFL; l.r = r * ks; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = eta; l.dist.ax = rough; l.dist.ay = rough;
            bsdf.add(l);
          }
          if (!t.is_black()) {
            Lobe l; l.kind = LOBE_MICRO_TRANS; l.t = t * ks; l.eta_a = 1.0f; l.eta_b = eta; l.fresnel.kind = 1; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = eta;
            l.dist.ax = rough; l.dist.ay = rough;
            bsdf.add(l);
          }
        }
      }
      bsdf.init(si, eta);
      return true;
    }
    case RT_MAT_MIX: {                                            // mixmat.rs:34-64
      Spectrum s1 = S(mt.amount).clamp0();
      Spectrum s2 = (Spectrum(1.0f) - s1).clamp0();
      Bsdf b2;
      // both children always yield a Bsdf on this path (every material/*.rs sets si.bsdf); the result keeps mat1's Bsdf
      // (its eta and frame) and replaces the lobe list by the scaled lobes of both
      SurfaceInteraction si2 = si;                               // mixmat.rs:43: mat2 works on a clone (its bump map never reaches the caller)
      if (!compute_scattering_functions(table, table[mt.mix_a], si, allow_multiple_lobes, bsdf, ts)) return false;
      if (!compute_scattering_functions(table, table[mt.mix_b], si2, allow_multiple_lobes, b2, ts)) return false;
      const int n1 = bsdf.n;
      for (int i = 0; i < n1; i++) {
        Lobe& l = bsdf.lobes[i];
        if (l.n_scales >= 2) throw std::runtime_error("oracle: MixMaterial nested deeper than two levels");
        l.scale[l.n_scales++] = s1;
      }
      if (n1 + b2.n > 8) throw std::runtime_error("oracle: more than 8 BxDFs (the reference's BxDFHolder panics, bsdf/mod.rs:41-52)");
      for (int i = 0; i < b2.n; i++) {
        Lobe l = b2.lobes[i];
        if (l.n_scales >= 2) throw std::runtime_error("oracle: MixMaterial nested deeper than two levels");
        l.scale[l.n_scales++] = s2;
        bsdf.add(l);
      }
      return true;
    }
    default: return false;                                        // no material: bsdf = None (path.rs:146-152)
  }
}

}  // namespace orc
