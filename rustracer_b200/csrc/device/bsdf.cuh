// Device BSDF: Bsdf frame + lobe selection (bsdf/mod.rs:64-269), the BxDFs used by Matte / Plastic / Metal /
// Glass / Mirror (bsdf/{lambertian,oren_nayar,fresnel,microfacet}.rs) and the material -> lobe mapping
// (material/{matte,plastic,metal,glass,mirror}.rs).  Product code; reference lines relative to rustracer-core/src/.
#pragma once
#include "shapes.cuh"

namespace rt {

enum : uint32_t { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };   // bsdf/mod.rs:24-32

// geometry/mod.rs:15-95 (local shading frame)
RT_DEV float cos_theta(V3 w) { return w.z; }
RT_DEV float cos2_theta(V3 w) { return w.z * w.z; }
RT_DEV float abs_cos_theta(V3 w) { return fabsf(w.z); }
RT_DEV float sin2_theta(V3 w) { return fmaxf(1.0f - cos2_theta(w), 0.0f); }
RT_DEV float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
RT_DEV float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
RT_DEV float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
RT_DEV float cos_phi(V3 w) { float st = sin_theta(w); return st == 0.0f ? 1.0f : clampf(w.x / st, -1.0f, 1.0f); }
RT_DEV float sin_phi(V3 w) { float st = sin_theta(w); return st == 0.0f ? 0.0f : clampf(w.y / st, -1.0f, 1.0f); }
RT_DEV float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
RT_DEV float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
RT_DEV bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
RT_DEV float spherical_theta(V3 v) { return acosf(clampf(v.z, -1.0f, 1.0f)); }
RT_DEV float spherical_phi(V3 v) { float p = atan2f(v.y, v.x); return p < 0.0f ? p + 2.0f * kPi : p; }

// sampling/mod.rs
RT_DEV V3 uniform_sample_sphere(P2 u) {                                          // :14-20
  float z = 1.0f - 2.0f * u.x;
  float r = sqrtf(fmaxf(1.0f - z * z, 0.0f));
  float phi = 2.0f * kPi * u.y;
  return v3(r * cosf(phi), r * sinf(phi), z);
}
RT_DEV P2 concentric_sample_disk(P2 u) {                                         // :28-47
  const float kFracPi4 = kFracPi2 / 2.0f;
  float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
  if (ox == 0.0f && oy == 0.0f) return mk2(0.0f, 0.0f);
  float r, theta;
  if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kFracPi4 * (oy / ox); }
  else { r = oy; theta = kFracPi2 - kFracPi4 * (ox / oy); }
  return mk2(r * cosf(theta), r * sinf(theta));
}
RT_DEV V3 cosine_sample_hemisphere(P2 u) {                                       // :22-26
  P2 d = concentric_sample_disk(u);
  float z = sqrtf(fmaxf(1.0f - d.x * d.x - d.y * d.y, 0.0f));
  return v3(d.x, d.y, z);
}
RT_DEV P2 uniform_sample_triangle(P2 u) { float su0 = sqrtf(u.x); return mk2(1.0f - su0, u.y * su0); }   // :49-52
RT_DEV float uniform_cone_pdf(float cos_theta_max) { return 1.0f / (2.0f * kPi * (1.0f - cos_theta_max)); } // :54-56
RT_DEV float power_heuristic(float f_pdf, float g_pdf) {                                                  // :58-63 with nf = ng = 1
  float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
  return (f * f) / (f * f + g * g);
}

// bsdf/fresnel.rs:14-31
RT_DEV V3 reflect(V3 wo, V3 n) { return -wo + n * 2.0f * dot(wo, n); }
RT_DEV bool refract(V3 i, V3 n, float eta, V3& wt) {
  float cos_theta_i = dot(n, i);
  float sin2theta_i = fmaxf(1.0f - cos_theta_i * cos_theta_i, 0.0f);
  float sin2theta_t = eta * eta * sin2theta_i;
  if (sin2theta_t >= 1.0f) return false;
  float cos_theta_t = sqrtf(1.0f - sin2theta_t);
  wt = eta * -i + (eta * cos_theta_i - cos_theta_t) * n;
  return true;
}
// fresnel.rs:33-58
RT_DEV float fr_dielectric(float cos_theta_i, float eta_i, float eta_t) {
  cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
  if (cos_theta_i <= 0.0f) { float tmp = eta_i; eta_i = eta_t; eta_t = tmp; cos_theta_i = fabsf(cos_theta_i); }
  float sin_theta_i = sqrtf(fmaxf(1.0f - cos_theta_i * cos_theta_i, 0.0f));
  float sin_theta_t = eta_i / eta_t * sin_theta_i;
  if (sin_theta_t >= 1.0f) return 1.0f;
  float cos_theta_t = sqrtf(fmaxf(1.0f - sin_theta_t * sin_theta_t, 0.0f));
  float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
  float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
  return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// fresnel.rs:60-82
RT_DEV Spec fr_conductor(float cos_theta_i, Spec eta_i, Spec eta_t, Spec k) {
  cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
  Spec eta = eta_t / eta_i, eta_k = k / eta_i;
  float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.0f - cos2;
  Spec eta2 = eta * eta, eta_k2 = eta_k * eta_k;
  Spec t0 = eta2 - eta_k2 - sin2;
  Spec a2plusb2 = spec_sqrt(t0 * t0 + 4.0f * eta2 * eta_k2);
  Spec t1 = a2plusb2 + cos2;
  Spec a = spec_sqrt(0.5f * (a2plusb2 + t0));
  Spec t2 = 2.0f * cos_theta_i * a;
  Spec r_s = (t1 - t2) / (t1 + t2);
  Spec t3 = cos2 * a2plusb2 + sin2 * sin2;
  Spec t4 = t2 * sin2;
  Spec r_p = r_s * (t3 - t4) / (t3 + t4);
  return 0.5f * (r_p + r_s);
}

enum { FR_NOOP = 0, FR_DIELECTRIC = 1, FR_CONDUCTOR = 2 };
// same numbering as RTGPU_LOBE_* (include/rtgpu.h)
enum { LOBE_LAMBERT_R = 0, LOBE_OREN_NAYAR, LOBE_SPEC_REFL, LOBE_SPEC_TRANS, LOBE_FRESNEL_SPEC, LOBE_MICRO_REFL, LOBE_MICRO_TRANS,
       LOBE_LAMBERT_T, LOBE_FRESNEL_BLEND };

// Lobe kinds and lobe counts a translation unit can meet (bit k = LOBE_* kind k).  The per-material path kernels (tu_path.cu) narrow
// them: their `switch (kind)` dispatch then drops the other lobes' code — the matte kernel carried 61 k SASS instructions of which a
// warp ever executed 4.9 k (profiles/r02f_ncu_full_shade_c3_raw.csv: 877 KB never executed, 18 % "no instruction" stalls).
#ifndef RT_LOBE_KINDS
#define RT_LOBE_KINDS 0x1ffu
#endif
#ifndef RT_MAX_BUILT_LOBES
#define RT_MAX_BUILT_LOBES 8
#endif
// RT_LOBE_OFF(K...) as the first statement of a `case`: a compile-time test, so the optimiser deletes the case of a kind this TU never meets
#define RT_LOBE_ON(K) (((RT_LOBE_KINDS) >> (K)) & 1u)
#define RT_LOBE_OFF1(A) if (!RT_LOBE_ON(A)) __builtin_unreachable();
#define RT_LOBE_OFF2(A, B) if (!(RT_LOBE_ON(A) | RT_LOBE_ON(B))) __builtin_unreachable();
#define RT_LOBE_OFF3(A, B, C) if (!(RT_LOBE_ON(A) | RT_LOBE_ON(B) | RT_LOBE_ON(C))) __builtin_unreachable();
RT_DEV int lobe_kind_known(int kind) { return kind; }
RT_DEV int lobe_count_known(int n) { if (n > RT_MAX_BUILT_LOBES) __builtin_unreachable(); return n; }

struct Lobe {
  int kind;
  Spec r, t;                        // FresnelBlend: rs in r, rd in t
  float on_a, on_b;                 // OrenNayar A, B
  int fr_kind; float fr_eta_i, fr_eta_t; Spec c_eta_t, c_k;   // Fresnel (conductor eta_i is always 1: metal.rs:70-75)
  float ax, ay;                     // TrowbridgeReitz alpha
  float eta_a, eta_b;
};

RT_DEV Spec fresnel_evaluate(const Lobe& l, float cos_theta_i) {                 // fresnel.rs:84-138 (abs(cos) for both)
  if (l.fr_kind == FR_DIELECTRIC) return spec(fr_dielectric(fabsf(cos_theta_i), l.fr_eta_i, l.fr_eta_t));
  if (l.fr_kind == FR_CONDUCTOR) return fr_conductor(fabsf(cos_theta_i), spec(1.0f), l.c_eta_t, l.c_k);
  return spec(1.0f);
}

// TrowbridgeReitzDistribution (microfacet.rs:469-650), sample_visible_area = true
RT_DEV float tr_d(float ax, float ay, V3 wh) {                                   // :576-588
  float tan2 = tan2_theta(wh);
  if (isinf(tan2)) return 0.0f;
  float cos4 = cos2_theta(wh) * cos2_theta(wh);
  float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * tan2;
  return 1.0f / (kPi * ax * ay * cos4 * (1.0f + e) * (1.0f + e));
}
RT_DEV float tr_lambda(float ax, float ay, V3 w) {                               // :590-602
  float abs_tan = fabsf(tan_theta(w));
  if (isinf(abs_tan)) return 0.0f;
  float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
  float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
  return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
RT_DEV float tr_g1(float ax, float ay, V3 w) { return 1.0f / (1.0f + tr_lambda(ax, ay, w)); }                          // :235-237
RT_DEV float tr_g(float ax, float ay, V3 wi, V3 wo) { return 1.0f / (1.0f + tr_lambda(ax, ay, wi) + tr_lambda(ax, ay, wo)); }   // :239-241
RT_DEV float tr_pdf(float ax, float ay, V3 wo, V3 wh) { return tr_d(ax, ay, wh) * tr_g1(ax, ay, wo) * fabsf(dot(wo, wh)) / abs_cos_theta(wo); }   // :243-249
RT_DEV void tr_sample11(float cos_t, float u1, float u2, float& slope_x, float& slope_y) {                             // :517-572
  if (cos_t > 0.9999f) {
    float r = sqrtf(u1 / (1.0f - u1));
    float phi = 6.28318530717958647692f * u2;
    slope_x = r * cosf(phi); slope_y = r * sinf(phi);
    return;
  }
  float sin_t = sqrtf(fmaxf(1.0f - cos_t * cos_t, 0.0f));
  float tan_t = sin_t / cos_t;
  float a = 1.0f / tan_t;
  float G1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
  float A = 2.0f * u1 / G1 - 1.0f;
  float tmp = 1.0f / (A * A - 1.0f);
  if (tmp > 1e10f) tmp = 1e10f;
  float B = tan_t;
  float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.0f));
  float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
  slope_x = (A < 0.0f || slope_x_2 > 1.0f / tan_t) ? slope_x_1 : slope_x_2;
  float S;
  if (u2 > 0.5f) { S = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
  else { S = -1.0f; u2 = 2.0f * (0.5f - u2); }
  float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  slope_y = S * z * sqrtf(1.0f + slope_x * slope_x);
}
RT_DEV V3 tr_sample(float ax, float ay, V3 wi, float u1, float u2) {             // :495-515
  V3 wis = normalize(v3(ax * wi.x, ay * wi.y, wi.z));
  float sx, sy; tr_sample11(cos_theta(wis), u1, u2, sx, sy);
  float tmp = cos_phi(wis) * sx - sin_phi(wis) * sy;
  sy = sin_phi(wis) * sx + cos_phi(wis) * sy;
  sx = tmp;
  sx *= ax; sy *= ay;
  return normalize(v3(-sx, -sy, 1.0f));
}
RT_DEV V3 tr_sample_wh(float ax, float ay, V3 wo, P2 u) {                        // :604-645 (visible-area branch)
  bool flip = wo.z < 0.0f;
  V3 w = flip ? -wo : wo;
  V3 wh = tr_sample(ax, ay, w, u.x, u.y);
  if (flip) wh = -wh;
  return wh;
}

RT_DEV uint32_t lobe_type(int kind) {
  switch (kind) {
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: return BSDF_DIFFUSE | BSDF_REFLECTION;
    case LOBE_SPEC_REFL: return BSDF_SPECULAR | BSDF_REFLECTION;
    case LOBE_SPEC_TRANS: return BSDF_SPECULAR | BSDF_TRANSMISSION;
    case LOBE_FRESNEL_SPEC: return BSDF_SPECULAR | BSDF_REFLECTION | BSDF_TRANSMISSION;
    case LOBE_MICRO_REFL: case LOBE_FRESNEL_BLEND: return BSDF_REFLECTION | BSDF_GLOSSY;
    case LOBE_LAMBERT_T: return BSDF_DIFFUSE | BSDF_TRANSMISSION;                // lambertian.rs:43-45
    default: return BSDF_TRANSMISSION | BSDF_GLOSSY;
  }
}
RT_DEV bool lobe_matches(int kind, uint32_t flags) { uint32_t t = lobe_type(kind); return (t & flags) == t; }   // bxdf.rs:29-31

RT_DEV float pow5(float v) { return (v * v) * (v * v) * v; }                     // fresnel.rs:414-417
RT_DEV Spec lobe_f_inner(const Lobe& l, V3 wo, V3 wi) {
  switch (lobe_kind_known(l.kind)) {
    case LOBE_LAMBERT_R: RT_LOBE_OFF1(LOBE_LAMBERT_R) return l.r * kInvPi;                                    // lambertian.rs:19-21
    case LOBE_LAMBERT_T: RT_LOBE_OFF1(LOBE_LAMBERT_T) return l.t * kInvPi;                                    // lambertian.rs:39-41
    case LOBE_FRESNEL_BLEND: { RT_LOBE_OFF1(LOBE_FRESNEL_BLEND)                   // fresnel.rs:358-375 (rs = r, rd = t)
      Spec diffuse = (28.0f / (23.0f * kPi)) * l.t * (spec(1.0f) - l.r) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                     (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
      V3 wh = wi + wo;
      if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return spec(0.0f);
      wh = normalize(wh);
      Spec schlick = l.r + pow5(1.0f - dot(wi, wh)) * (spec(1.0f) - l.r);        // :352-354
      Spec specular = tr_d(l.ax, l.ay, wh) / (4.0f * fabsf(dot(wi, wh)) * fmaxf(abs_cos_theta(wi), abs_cos_theta(wo))) * schlick;
      return diffuse + specular;
    }
    case LOBE_OREN_NAYAR: { RT_LOBE_OFF1(LOBE_OREN_NAYAR)                        // oren_nayar.rs:30-52
      float sti = sin_theta(wi), sto = sin_theta(wo);
      float max_cos = 0.0f;
      if (sti > 1e-4f && sto > 1e-4f) {
        float d_cos = sin_phi(wi) * sin_phi(wo) + cos_phi(wi) * cos_phi(wo);
        max_cos = fmaxf(d_cos, 0.0f);
      }
      float sin_alpha, tan_beta;
      if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sto; tan_beta = sti / abs_cos_theta(wi); }
      else { sin_alpha = sti; tan_beta = sto / abs_cos_theta(wo); }
      return l.r * kInvPi * (l.on_a + l.on_b * max_cos * sin_alpha * tan_beta);
    }
    case LOBE_MICRO_REFL: { RT_LOBE_OFF1(LOBE_MICRO_REFL)                        // microfacet.rs:36-53
      float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
      V3 wh = wi + wo;
      if (cto == 0.0f || cti == 0.0f) return spec(0.0f);
      if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return spec(0.0f);
      wh = normalize(wh);
      Spec F = fresnel_evaluate(l, dot(wi, wh));
      return l.r * tr_d(l.ax, l.ay, wh) * tr_g(l.ax, l.ay, wo, wi) * F / (4.0f * cti * cto);
    }
    case LOBE_MICRO_TRANS: { RT_LOBE_OFF1(LOBE_MICRO_TRANS)                      // microfacet.rs:125-169
      if (same_hemisphere(wo, wi)) return spec(0.0f);
      float cto = cos_theta(wo), cti = cos_theta(wi);
      if (cto == 0.0f || cti == 0.0f) return spec(0.0f);
      float eta = cto > 0.0f ? l.eta_b / l.eta_a : l.eta_a / l.eta_b;
      V3 wh = normalize(wo + wi * eta);
      if (wh.z < 0.0f) wh = -wh;
      Spec F = fresnel_evaluate(l, dot(wo, wh));
      float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
      float factor = 1.0f / eta;                                                 // TransportMode::RADIANCE on this path
      return (spec(1.0f) - F) * l.t *
             fabsf(tr_d(l.ax, l.ay, wh) * tr_g(l.ax, l.ay, wo, wi) * eta * eta * fabsf(dot(wi, wh)) * fabsf(dot(wo, wh)) * factor * factor /
                   (cti * cto * sqrt_denom * sqrt_denom));
    }
    default: return spec(0.0f);                                                  // specular lobes: f() is zero
  }
}
RT_DEV float lobe_pdf_inner(const Lobe& l, V3 wo, V3 wi) {
  switch (lobe_kind_known(l.kind)) {
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: case LOBE_LAMBERT_T:              // bxdf.rs:38-44 (LambertianTransmission keeps the default)
      RT_LOBE_OFF3(LOBE_LAMBERT_R, LOBE_OREN_NAYAR, LOBE_LAMBERT_T)
      return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0.0f;
    case LOBE_FRESNEL_BLEND: { RT_LOBE_OFF1(LOBE_FRESNEL_BLEND)                   // fresnel.rs:377-385
      if (!same_hemisphere(wo, wi)) return 0.0f;
      V3 wh = normalize(wo + wi);
      float pdf_wh = tr_pdf(l.ax, l.ay, wo, wh);
      return 0.5f * (abs_cos_theta(wi) * kInvPi + pdf_wh / (4.0f * dot(wo, wh)));
    }
    case LOBE_MICRO_REFL: { RT_LOBE_OFF1(LOBE_MICRO_REFL)                        // microfacet.rs:87-94
      if (!same_hemisphere(wo, wi)) return 0.0f;
      V3 wh = normalize(wo + wi);
      return tr_pdf(l.ax, l.ay, wo, wh) / (4.0f * dot(wo, wh));
    }
    case LOBE_MICRO_TRANS: { RT_LOBE_OFF1(LOBE_MICRO_TRANS)                      // microfacet.rs:207-222
      if (same_hemisphere(wo, wi)) return 0.0f;
      float eta = cos_theta(wo) > 0.0f ? l.eta_b / l.eta_a : l.eta_a / l.eta_b;
      V3 wh = normalize(wo + wi * eta);
      float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
      float dwh_dwi = fabsf((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
      return tr_pdf(l.ax, l.ay, wo, wh) * dwh_dwi;
    }
    default: return 0.0f;
  }
}
// `sampled` is the BxdfType the lobe reports (bxdf.rs:18-25: the default sample_f reports EMPTY)
RT_DEV void lobe_sample_f_inner(const Lobe& l, V3 wo, P2 u, Spec& f_out, V3& wi, float& pdf_out, uint32_t& sampled) {
  switch (lobe_kind_known(l.kind)) {
    case LOBE_FRESNEL_BLEND: { RT_LOBE_OFF1(LOBE_FRESNEL_BLEND)                   // fresnel.rs:387-407
      sampled = lobe_type(l.kind);
      if (u.x < 0.5f) {
        u.x = fminf(2.0f * u.x, kOneMinusEpsilon);
        wi = cosine_sample_hemisphere(u);
        if (wo.z < 0.0f) wi.z *= -1.0f;
      } else {
        u.x = fminf(2.0f * (u.x - 0.5f), kOneMinusEpsilon);
        V3 wh = tr_sample_wh(l.ax, l.ay, wo, u);
        wi = reflect(wo, wh);
        if (!same_hemisphere(wo, wi)) { f_out = spec(0.0f); pdf_out = 0.0f; return; }
      }
      f_out = lobe_f_inner(l, wo, wi); pdf_out = lobe_pdf_inner(l, wo, wi);
      return;
    }
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: case LOBE_LAMBERT_T: { RT_LOBE_OFF3(LOBE_LAMBERT_R, LOBE_OREN_NAYAR, LOBE_LAMBERT_T)   // bxdf.rs:18-25: the hemisphere of wo, also for LambertianTransmission
      wi = cosine_sample_hemisphere(u);
      if (wo.z < 0.0f) wi.z *= -1.0f;
      pdf_out = lobe_pdf_inner(l, wo, wi); f_out = lobe_f_inner(l, wo, wi); sampled = 0;
      return;
    }
    case LOBE_SPEC_REFL: { RT_LOBE_OFF1(LOBE_SPEC_REFL)                          // fresnel.rs:159-164
      wi = v3(-wo.x, -wo.y, wo.z);
      f_out = fresnel_evaluate(l, cos_theta(wi)) * l.r / abs_cos_theta(wi);
      pdf_out = 1.0f; sampled = lobe_type(l.kind);
      return;
    }
    case LOBE_SPEC_TRANS: { RT_LOBE_OFF1(LOBE_SPEC_TRANS)                        // fresnel.rs:203-230
      bool entering = cos_theta(wo) > 0.0f;
      float ei = entering ? l.eta_a : l.eta_b, et = entering ? l.eta_b : l.eta_a;
      V3 w;
      if (refract(wo, face_forward(v3(0, 0, 1), wo), ei / et, w)) {
        wi = w;
        Spec ft = l.t * (spec(1.0f) - fresnel_evaluate(l, cos_theta(wi)));
        ft = ft * (ei * ei) / (et * et);
        f_out = ft / abs_cos_theta(wi); pdf_out = 1.0f; sampled = lobe_type(l.kind);
      } else { f_out = spec(1.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; }
      return;
    }
    case LOBE_FRESNEL_SPEC: { RT_LOBE_OFF1(LOBE_FRESNEL_SPEC)                    // fresnel.rs:273-322
      float fr = fr_dielectric(cos_theta(wo), l.eta_a, l.eta_b);
      if (u.x < fr) {
        wi = v3(-wo.x, -wo.y, wo.z);
        f_out = fr * l.r / abs_cos_theta(wi); pdf_out = fr; sampled = BSDF_SPECULAR | BSDF_REFLECTION;
      } else {
        bool entering = cos_theta(wo) > 0.0f;
        float ei = entering ? l.eta_a : l.eta_b, et = entering ? l.eta_b : l.eta_a;
        V3 w;
        if (refract(wo, face_forward(v3(0, 0, 1), wo), ei / et, w)) {
          wi = w;
          Spec ft = l.t * (1.0f - fr);
          ft = ft * ((ei * ei) / (et * et));
          f_out = ft / abs_cos_theta(wi); pdf_out = 1.0f - fr; sampled = BSDF_SPECULAR | BSDF_TRANSMISSION;
        } else { f_out = spec(0.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; }
      }
      return;
    }
    case LOBE_MICRO_REFL: { RT_LOBE_OFF1(LOBE_MICRO_REFL)                        // microfacet.rs:61-85
      sampled = lobe_type(l.kind);
      if (wo.z == 0.0f) { f_out = spec(0.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; return; }
      V3 wh = tr_sample_wh(l.ax, l.ay, wo, u);
      wi = reflect(wo, wh);
      if (!same_hemisphere(wo, wi)) { f_out = spec(0.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; return; }
      pdf_out = tr_pdf(l.ax, l.ay, wo, wh) / (4.0f * dot(wo, wh));
      f_out = lobe_f_inner(l, wo, wi);
      return;
    }
    default: { RT_LOBE_OFF1(LOBE_MICRO_TRANS)                                    // LOBE_MICRO_TRANS microfacet.rs:177-205
      sampled = lobe_type(l.kind);
      if (wo.z == 0.0f) { f_out = spec(0.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; return; }
      V3 wh = tr_sample_wh(l.ax, l.ay, wo, u);
      float eta = cos_theta(wo) > 0.0f ? l.eta_a / l.eta_b : l.eta_b / l.eta_a;
      V3 w;
      if (refract(wo, wh, eta, w)) { wi = w; pdf_out = lobe_pdf_inner(l, wo, wi); f_out = lobe_f_inner(l, wo, wi); }
      else { f_out = spec(0.0f); wi = v3(0, 0, 0); pdf_out = 0.0f; }
      return;
    }
  }
}

constexpr int kMaxLobes = 2;     // the five materials with their own shade kernels build at most two lobes
// Translation units that only ever shade one of those five materials (tu_path.cu with RT_PATH_MAT != Q_LOBES) compile the
// host-listed-lobes path out (it cost the matte kernel 10 % when left to a run-time test: profiles/r01g).
#if defined(RT_PATH_MAT) && RT_PATH_MAT != 6
constexpr bool kListedLobes = false;
#else
constexpr bool kListedLobes = true;
#endif

// Bsdf (bsdf/mod.rs:64-269).  The lobes are either built in place (`lobes`, the five materials above) or are the
// host-listed rows of an RTGPU_MAT_LOBES material (`g`, up to 8: uber / substrate / translucent / mix).
struct Bsdf {
  float eta;
  V3 ns, ng, ss, ts;
  Lobe lobes[kMaxLobes]; int n;
  const rtgpu_lobe* g;
};
RT_DEV Lobe load_lobe(const rtgpu_lobe& g) {
  Lobe l;
  l.kind = (int)g.kind;
  l.r = spec3(g.r); l.t = spec3(g.t); l.on_a = g.on_a; l.on_b = g.on_b;
  l.fr_kind = (int)g.fr_kind; l.fr_eta_i = g.fr_eta_i; l.fr_eta_t = g.fr_eta_t; l.c_eta_t = spec3(g.c_eta_t); l.c_k = spec3(g.c_k);
  l.ax = g.ax; l.ay = g.ay; l.eta_a = g.eta_a; l.eta_b = g.eta_b;
  return l;
}
RT_DEV Lobe bsdf_lobe(const Bsdf& b, int i) { return (kListedLobes && b.g) ? load_lobe(b.g[i]) : b.lobes[i]; }
RT_DEV int bsdf_lobe_kind(const Bsdf& b, int i) { return (kListedLobes && b.g) ? (int)b.g[i].kind : b.lobes[i].kind; }
// ---- the BxDF interface as Bsdf sees it: lobe i itself, or lobe i behind its ScaledBxDF wrappers (bxdf.rs:48-71; only the
// host-listed lobes of a mix have any): f and sample_f are scaled, the type is the wrapped lobe's, and pdf() is NOT
// forwarded — it is the trait's default cosine pdf (bxdf.rs:38-44).
RT_DEV Spec scale_by_wrappers(const Bsdf& b, int i, Spec v) {
  if (kListedLobes && b.g) {
    const rtgpu_lobe& g = b.g[i];
    if (g.n_scales > 0) v = v * spec3(g.scale[0]);
    if (g.n_scales > 1) v = v * spec3(g.scale[1]);
  }
  return v;
}
RT_DEV Spec lobe_f(const Bsdf& b, int i, V3 wo, V3 wi) { return scale_by_wrappers(b, i, lobe_f_inner(bsdf_lobe(b, i), wo, wi)); }
RT_DEV float lobe_pdf(const Bsdf& b, int i, V3 wo, V3 wi) {
  if (kListedLobes && b.g && b.g[i].n_scales > 0) return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0.0f;
  return lobe_pdf_inner(bsdf_lobe(b, i), wo, wi);
}
RT_DEV void lobe_sample_f(const Bsdf& b, int i, V3 wo, P2 u, Spec& f_out, V3& wi, float& pdf_out, uint32_t& sampled) {
  lobe_sample_f_inner(bsdf_lobe(b, i), wo, u, f_out, wi, pdf_out, sampled);
  f_out = scale_by_wrappers(b, i, f_out);
}
RT_DEV void bsdf_init(Bsdf& b, const SurfHit& si, float eta) {                   // :77-92
  b.eta = eta;
  b.ss = normalize(si.dpdu_s);
  b.ns = si.ns; b.ng = si.n;
  b.ts = cross(si.ns, b.ss);
  b.n = 0; b.g = nullptr;
}
RT_DEV V3 world_to_local(const Bsdf& b, V3 v) { return v3(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }   // :253-255
RT_DEV V3 local_to_world(const Bsdf& b, V3 v) {                                                        // :257-263
  return v3(b.ss.x * v.x + b.ts.x * v.y + b.ns.x * v.z, b.ss.y * v.x + b.ts.y * v.y + b.ns.y * v.z, b.ss.z * v.x + b.ts.z * v.y + b.ns.z * v.z);
}
RT_DEV int bsdf_num_components(const Bsdf& b, uint32_t flags) {                  // :265-268
  int c = 0;
  for (int i = 0; i < lobe_count_known(b.n); i++) if (lobe_matches(bsdf_lobe_kind(b, i), flags)) c++;
  return c;
}
RT_DEV Spec bsdf_f(const Bsdf& b, V3 wo_w, V3 wi_w, uint32_t flags) {            // :94-112
  V3 wi = world_to_local(b, wi_w), wo = world_to_local(b, wo_w);
  if (wo.z == 0.0f) return spec(0.0f);
  bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
  Spec c = spec(0.0f);
  for (int i = 0; i < lobe_count_known(b.n); i++) {
    const int kind = bsdf_lobe_kind(b, i);
    uint32_t t = lobe_type(kind);
    if (lobe_matches(kind, flags) && ((refl && (t & BSDF_REFLECTION)) || (!refl && (t & BSDF_TRANSMISSION)))) c = c + lobe_f(b, i, wo, wi);
  }
  return c;
}
RT_DEV float bsdf_pdf(const Bsdf& b, V3 wo_w, V3 wi_w, uint32_t flags) {         // :114-136
  if (b.n == 0) return 0.0f;
  V3 wo = world_to_local(b, wo_w);
  if (wo.z == 0.0f) return 0.0f;
  V3 wi = world_to_local(b, wi_w);
  int matched = 0; float p = 0.0f;
  for (int i = 0; i < lobe_count_known(b.n); i++) if (lobe_matches(bsdf_lobe_kind(b, i), flags)) { matched++; p += lobe_pdf(b, i, wo, wi); }
  return matched == 0 ? 0.0f : p / (float)matched;
}
// the lobes built in place: the matching lobes are listed once (kernels that never see host-listed lobes use this form)
RT_DEV void bsdf_sample_f_built(const Bsdf& b, V3 wo_w, P2 u, uint32_t flags, Spec& f_out, V3& wi_w, float& pdf_out, uint32_t& sampled) {   // :138-251
  int m[RT_MAX_BUILT_LOBES]; int nm = 0;
  for (int i = 0; i < lobe_count_known(b.n); i++) if (lobe_matches(b.lobes[i].kind, flags)) m[nm++] = i;
  if (nm == 0) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
  int comp = (int)min(f2u32(floorf(u.x * (float)nm)), (uint32_t)(nm - 1));
  const Lobe& bxdf = b.lobes[m[comp]];
  const uint32_t btype = lobe_type(bxdf.kind);
  P2 ur = mk2(fminf(u.x * (float)nm - (float)comp, kOneMinusEpsilon), u.y);
  V3 wo = world_to_local(b, wo_w);
  if (wo.z == 0.0f) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = btype; return; }
  Spec f; V3 wi; float pdf;
  lobe_sample_f_inner(bxdf, wo, ur, f, wi, pdf, sampled);
  if (pdf == 0.0f) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
  wi_w = local_to_world(b, wi);
  if (!(btype & BSDF_SPECULAR) && nm > 1)
    for (int i = 0; i < nm; i++) if (i != comp) pdf += lobe_pdf_inner(b.lobes[m[i]], wo, wi);
  if (nm > 1) pdf /= (float)nm;
  if (!(btype & BSDF_SPECULAR)) {
    bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
    f = spec(0.0f);
    for (int i = 0; i < nm; i++) {
      uint32_t t = lobe_type(b.lobes[m[i]].kind);
      if ((refl && (t & BSDF_REFLECTION)) || (!refl && (t & BSDF_TRANSMISSION))) f = f + lobe_f_inner(b.lobes[m[i]], wo, wi);
    }
  }
  f_out = f; pdf_out = pdf;
}
RT_DEV void bsdf_sample_f(const Bsdf& b, V3 wo_w, P2 u, uint32_t flags, Spec& f_out, V3& wi_w, float& pdf_out, uint32_t& sampled) {   // :138-251
  if (!kListedLobes) { bsdf_sample_f_built(b, wo_w, u, flags, f_out, wi_w, pdf_out, sampled); return; }
  const int nm = bsdf_num_components(b, flags);
  if (nm == 0) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
  const int comp = (int)min(f2u32(floorf(u.x * (float)nm)), (uint32_t)(nm - 1));
  int chosen = 0;                                                                // the comp-th matching lobe (:150-160)
  for (int i = 0, c = comp; i < b.n; i++) if (lobe_matches(bsdf_lobe_kind(b, i), flags)) { if (c == 0) { chosen = i; break; } c--; }
  const uint32_t btype = lobe_type(bsdf_lobe_kind(b, chosen));
  P2 ur = mk2(fminf(u.x * (float)nm - (float)comp, kOneMinusEpsilon), u.y);
  V3 wo = world_to_local(b, wo_w);
  if (wo.z == 0.0f) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = btype; return; }
  Spec f; V3 wi; float pdf;
  lobe_sample_f(b, chosen, wo, ur, f, wi, pdf, sampled);
  if (pdf == 0.0f) { f_out = spec(0.0f); wi_w = v3(0, 0, 0); pdf_out = 0.0f; sampled = 0; return; }
  wi_w = local_to_world(b, wi);
  if (!(btype & BSDF_SPECULAR) && nm > 1)
    for (int i = 0; i < b.n; i++) if (i != chosen && lobe_matches(bsdf_lobe_kind(b, i), flags)) pdf += lobe_pdf(b, i, wo, wi);
  if (nm > 1) pdf /= (float)nm;
  if (!(btype & BSDF_SPECULAR)) {
    bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
    f = spec(0.0f);
    for (int i = 0; i < b.n; i++) {
      const int kind = bsdf_lobe_kind(b, i);
      uint32_t t = lobe_type(kind);
      if (lobe_matches(kind, flags) && ((refl && (t & BSDF_REFLECTION)) || (!refl && (t & BSDF_TRANSMISSION)))) f = f + lobe_f(b, i, wo, wi);
    }
  }
  f_out = f; pdf_out = pdf;
}

RT_DEV Lobe blank_lobe() {
  Lobe l;
  l.kind = LOBE_LAMBERT_R; l.r = spec(0.0f); l.t = spec(0.0f); l.on_a = 0.0f; l.on_b = 0.0f;
  l.fr_kind = FR_NOOP; l.fr_eta_i = 1.0f; l.fr_eta_t = 1.0f; l.c_eta_t = spec(1.0f); l.c_k = spec(0.0f);
  l.ax = 0.0f; l.ay = 0.0f; l.eta_a = 1.0f; l.eta_b = 1.0f;
  return l;
}

// Material::compute_scattering_functions for constant textures.  The host already evaluated the textures and
// the libm-dependent scalars (roughness_to_alpha, OrenNayar A/B): see rtgpu_material.  allow_multiple_lobes:
// Path passes true (path.rs:145), Whitted / DirectLighting false (whitted.rs:57, directlighting.rs:104).
// Returns false for "no material" (bsdf = None, path.rs:146-152).
// `type` is mt.type; the material-sorted shade kernels pass it as a compile-time constant so the switch folds.
// `lobe_table`: DScene::lobes (rows of the RTGPU_MAT_LOBES materials).
RT_DEV bool make_bsdf(uint32_t type, const rtgpu_material& mt, const rtgpu_lobe* lobe_table, const SurfHit& si, bool allow_multiple_lobes, Bsdf& bsdf) {
  switch (type) {
    case RTGPU_MAT_LOBES: {                                                      // uber.rs / substrate.rs / translucent.rs / mixmat.rs
      const int a = allow_multiple_lobes ? 1 : 0;
      bsdf_init(bsdf, si, mt.bsdf_eta);
      bsdf.g = lobe_table + mt.lobe_first[a]; bsdf.n = (int)mt.lobe_count[a];
      return true;
    }
    case RTGPU_MAT_MATTE: {                                                      // matte.rs:37-62
      Spec r = spec3(mt.kd);
      bsdf_init(bsdf, si, 1.0f);
      if (!is_black(r)) {
        Lobe l = blank_lobe();
        l.r = r;
        if (!mt.use_oren_nayar) l.kind = LOBE_LAMBERT_R;
        else { l.kind = LOBE_OREN_NAYAR; l.on_a = mt.oren_a; l.on_b = mt.oren_b; }
        bsdf.lobes[bsdf.n++] = l;
      }
      return true;
    }
    case RTGPU_MAT_PLASTIC: {                                                    // plastic.rs:45-74
      Spec kd = spec3(mt.kd), ks = spec3(mt.ks);
      bsdf_init(bsdf, si, 1.0f);
      if (!is_black(kd)) { Lobe l = blank_lobe(); l.kind = LOBE_LAMBERT_R; l.r = kd; bsdf.lobes[bsdf.n++] = l; }
      if (!is_black(ks)) {
        Lobe l = blank_lobe(); l.kind = LOBE_MICRO_REFL; l.r = ks;
        l.fr_kind = FR_DIELECTRIC; l.fr_eta_i = 1.5f; l.fr_eta_t = 1.0f;
        l.ax = mt.alpha_u; l.ay = mt.alpha_v;
        bsdf.lobes[bsdf.n++] = l;
      }
      return true;
    }
    case RTGPU_MAT_METAL: {                                                      // metal.rs:50-81
      Lobe l = blank_lobe(); l.kind = LOBE_MICRO_REFL; l.r = spec(1.0f);
      l.fr_kind = FR_CONDUCTOR; l.c_eta_t = spec3(mt.eta_rgb); l.c_k = spec3(mt.k_rgb);
      l.ax = mt.alpha_u; l.ay = mt.alpha_v;
      bsdf_init(bsdf, si, 1.0f);
      bsdf.lobes[bsdf.n++] = l;
      return true;
    }
    case RTGPU_MAT_GLASS: {                                                      // glass.rs:53-106
      const float eta = mt.eta;
      Spec r = spec3(mt.kr), t = spec3(mt.kt);
      bsdf_init(bsdf, si, eta);
      if (!is_black(r) || !is_black(t)) {
        const bool is_specular = mt.glass_specular != 0;
        if (is_specular && allow_multiple_lobes) {
          Lobe l = blank_lobe(); l.kind = LOBE_FRESNEL_SPEC; l.r = r; l.t = t; l.eta_a = 1.0f; l.eta_b = eta;
          bsdf.lobes[bsdf.n++] = l;
        } else {
          if (!is_black(r)) {
            Lobe l = blank_lobe(); l.r = r; l.fr_kind = FR_DIELECTRIC; l.fr_eta_i = 1.0f; l.fr_eta_t = eta;
            if (is_specular) l.kind = LOBE_SPEC_REFL; else { l.kind = LOBE_MICRO_REFL; l.ax = mt.alpha_u; l.ay = mt.alpha_v; }
            bsdf.lobes[bsdf.n++] = l;
          }
          if (!is_black(t)) {
            Lobe l = blank_lobe(); l.eta_a = 1.0f; l.eta_b = eta; l.fr_kind = FR_DIELECTRIC; l.fr_eta_i = 1.0f; l.fr_eta_t = eta;
            if (is_specular) { l.kind = LOBE_SPEC_TRANS; l.t = t; }
            else { l.kind = LOBE_MICRO_TRANS; l.t = r; l.ax = mt.alpha_u; l.ay = mt.alpha_v; }   // built with Kr: glass.rs:97
            bsdf.lobes[bsdf.n++] = l;
          }
        }
      }
      return true;
    }
    case RTGPU_MAT_MIRROR: {                                                     // mirror.rs:30-48
      Spec R = spec3(mt.kr);
      bsdf_init(bsdf, si, 1.0f);
      if (!is_black(R)) { Lobe l = blank_lobe(); l.kind = LOBE_SPEC_REFL; l.r = R; l.fr_kind = FR_NOOP; bsdf.lobes[bsdf.n++] = l; }
      return true;
    }
    default: return false;
  }
}

}  // namespace rt
