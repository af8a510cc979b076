// Device math for the B200 wavefront renderer (product code, sm_100a).
//
// Arithmetic contract: f32 everywhere, IEEE round-to-nearest + - * / sqrt, NO fused
// multiply-add (the translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false),
// expression order exactly as rustracer writes it, f64 only where rustracer uses f64.  Reference line
// numbers are relative to rustracer-core/src/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rt {

#define RT_DEV __device__ __forceinline__
#define RT_HD __host__ __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kFracPi2 = 1.57079632679489661923f;
constexpr float kMachineEpsilon = 5.9604644775390625e-08f;     // f32::EPSILON * 0.5 (lib.rs:87)
constexpr float kOneMinusEpsilon = 0.99999994f;                // lib.rs:95
constexpr float kU32ToUnit = 2.3283064365386963e-10f;          // 2^-32

RT_DEV float inf_f() { return __int_as_float(0x7f800000); }
// lib.rs:88-92: gamma(n) = n*eps / (1 - n*eps), evaluated in f32.  n is a compile-time constant at every
// call site, so the compiler folds this with IEEE semantics (same bits as the runtime expression).
RT_DEV constexpr float gamma_f(int n) { return ((float)n * kMachineEpsilon) / (1.0f - (float)n * kMachineEpsilon); }

// lib.rs:226-244
RT_DEV float next_float_up(float v) {
  if (v == inf_f()) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = __float_as_uint(v);
  if (v >= 0.0f) ui += 1; else ui -= 1;
  return __uint_as_float(ui);
}
// lib.rs:246-262
RT_DEV float next_float_down(float v) {
  if (v == -inf_f()) return v;
  if (v == 0.0f) v = -0.0f;
  uint32_t ui = __float_as_uint(v);
  if (v > 0.0f) ui -= 1; else ui += 1;
  return __uint_as_float(ui);
}

// lib.rs:191-207 PartialOrd min / max and lib.rs:264-275 clamp (NOT fminf/fmaxf: `if a < b {a} else {b}`)
RT_DEV float pmin(float a, float b) { return a < b ? a : b; }
RT_DEV float pmax(float a, float b) { return a > b ? a : b; }
RT_DEV float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Rust float -> integer `as` casts saturate and map NaN to 0.
RT_DEV int32_t f2i32(float f) { return __float2int_rz(f); }          // cvt.rzi.s32.f32 saturates, NaN -> 0
RT_DEV uint32_t f2u32(float f) { return __float2uint_rz(f); }        // cvt.rzi.u32.f32 saturates, NaN -> 0
RT_DEV float to_radians(float deg) { return deg * (kPi / 180.0f); }

struct V3 { float x, y, z; };
RT_DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_DEV V3 v3(float4 a) { return v3(a.x, a.y, a.z); }
RT_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
RT_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
RT_DEV V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RT_DEV V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }          // vector.rs:354-360: true division
RT_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             // vector.rs:241-243
RT_DEV V3 cross(V3 a, V3 b) {                                                           // vector.rs:280-286
  return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
RT_DEV float length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
RT_DEV float length(V3 a) { return sqrtf(length_squared(a)); }
RT_DEV V3 normalize(V3 a) { return a / length(a); }                                    // vector.rs:276-278
RT_DEV V3 vabs(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
RT_DEV float distance_squared(V3 p1, V3 p2) { return length_squared(p2 - p1); }        // geometry/mod.rs:222-224
RT_DEV float distance(V3 p1, V3 p2) { return length(p2 - p1); }
RT_DEV float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
RT_DEV int max_dimension(V3 v) { return v.x > v.y ? (v.x > v.z ? 0 : 2) : (v.y > v.z ? 1 : 2); }   // lib.rs:119-135
RT_DEV float max_component(V3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }               // lib.rs:137-139
RT_DEV V3 permute(V3 v, int x, int y, int z) { return v3(comp(v, x), comp(v, y), comp(v, z)); }
// lib.rs:158-168
RT_DEV void coordinate_system(V3 v1, V3& v2, V3& v3_) {
  if (fabsf(v1.x) > fabsf(v1.y)) v2 = v3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
  else v2 = v3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
  v3_ = cross(v1, v2);
}
RT_DEV V3 face_forward(V3 v1, V3 v2) { return dot(v1, v2) < 0.0f ? -v1 : v1; }        // geometry/mod.rs:127-143

struct P2 { float x, y; };
RT_DEV P2 mk2(float x, float y) { P2 r; r.x = x; r.y = y; return r; }

// geometry/mod.rs:203-220
RT_DEV V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
  float d = dot(vabs(n), p_error);
  V3 offset = d * n;
  if (dot(w, n) < 0.0f) offset = -offset;
  V3 po = p + offset;
  if (offset.x > 0.0f) po.x = next_float_up(po.x); else if (offset.x < 0.0f) po.x = next_float_down(po.x);
  if (offset.y > 0.0f) po.y = next_float_up(po.y); else if (offset.y < 0.0f) po.y = next_float_down(po.y);
  if (offset.z > 0.0f) po.z = next_float_up(po.z); else if (offset.z < 0.0f) po.z = next_float_down(po.z);
  return po;
}

// spectrum.rs:15-165, operators :222-393 (component-wise RGB)
struct Spec { float r, g, b; };
RT_DEV Spec spec(float v) { Spec s; s.r = v; s.g = v; s.b = v; return s; }
RT_DEV Spec spec(float r, float g, float b) { Spec s; s.r = r; s.g = g; s.b = b; return s; }
RT_DEV Spec spec3(const float* c) { return spec(c[0], c[1], c[2]); }
RT_DEV Spec operator+(Spec a, Spec b) { return spec(a.r + b.r, a.g + b.g, a.b + b.b); }
RT_DEV Spec operator-(Spec a, Spec b) { return spec(a.r - b.r, a.g - b.g, a.b - b.b); }
RT_DEV Spec operator*(Spec a, Spec b) { return spec(a.r * b.r, a.g * b.g, a.b * b.b); }
RT_DEV Spec operator/(Spec a, Spec b) { return spec(a.r / b.r, a.g / b.g, a.b / b.b); }
RT_DEV Spec operator*(Spec a, float s) { return spec(a.r * s, a.g * s, a.b * s); }
RT_DEV Spec operator*(float s, Spec a) { return spec(s * a.r, s * a.g, s * a.b); }
RT_DEV Spec operator/(Spec a, float s) { return spec(a.r / s, a.g / s, a.b / s); }
RT_DEV Spec operator+(Spec a, float s) { return spec(a.r + s, a.g + s, a.b + s); }
RT_DEV Spec operator-(Spec a, float s) { return spec(a.r - s, a.g - s, a.b - s); }
RT_DEV bool is_black(Spec s) { return s.r == 0.0f && s.g == 0.0f && s.b == 0.0f; }
RT_DEV bool has_nan(Spec s) { return isnan(s.r) || isnan(s.g) || isnan(s.b); }
RT_DEV float lum(Spec s) { return 0.212671f * s.r + 0.715160f * s.g + 0.072169f * s.b; }    // spectrum.rs:147-150
RT_DEV float max_component_value(Spec s) { return fmaxf(fmaxf(s.r, s.g), s.b); }             // spectrum.rs:152-154
RT_DEV Spec spec_sqrt(Spec s) { return spec(sqrtf(s.r), sqrtf(s.g), sqrtf(s.b)); }
RT_DEV void to_xyz(Spec s, float xyz[3]) {                                                   // spectrum.rs:99-107
  xyz[0] = 0.412453f * s.r + 0.357580f * s.g + 0.180423f * s.b;
  xyz[1] = 0.212671f * s.r + 0.715160f * s.g + 0.072169f * s.b;
  xyz[2] = 0.019334f * s.r + 0.119193f * s.g + 0.950227f * s.b;
}
RT_DEV Spec from_xyz(float x, float y, float z) {                                            // spectrum.rs:92-97
  return spec(3.240479f * x - 1.537150f * y - 0.498535f * z, -0.969256f * x + 1.875991f * y + 0.041556f * z,
              0.055648f * x - 0.204043f * y + 1.057311f * z);
}

struct Ray { V3 o, d; float t_max; };
RT_DEV Ray make_ray(V3 o, V3 d, float t_max) { Ray r; r.o = o; r.d = d; r.t_max = t_max; return r; }
RT_DEV V3 ray_at(const Ray& r, float t) { return r.o + t * r.d; }

// ---- 4x4 transforms stored row-major (float[16]); transform.rs ------------------------------------
struct Mat { const float* m; };
RT_DEV V3 xf_point(const float* m, V3 p) {                                                   // transform.rs:263-287
  float x = p.x, y = p.y, z = p.z;
  float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
  float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
  float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
  float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
  if (wp == 1.0f) return v3(xp, yp, zp);
  return v3(xp, yp, zp) / wp;
}
RT_DEV V3 xf_vector(const float* m, V3 v) {                                                  // transform.rs:289-304
  float x = v.x, y = v.y, z = v.z;
  return v3(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
// affine transform given by its rows 0..2 (row 3 = 0 0 0 1, so transform.rs:263-287 takes its `wp == 1` branch)
RT_DEV V3 xf_point_affine(const float* m, V3 p) {
  float x = p.x, y = p.y, z = p.z;
  return v3(m[0] * x + m[1] * y + m[2] * z + m[3], m[4] * x + m[5] * y + m[6] * z + m[7], m[8] * x + m[9] * y + m[10] * z + m[11]);
}
// Normal transform uses the transpose of the inverse: pass m_inv (transform.rs:244-254, :306-320).
RT_DEV V3 xf_normal(const float* mi, V3 n) {
  float x = n.x, y = n.y, z = n.z;
  return v3(mi[0] * x + mi[4] * y + mi[8] * z, mi[1] * x + mi[5] * y + mi[9] * z, mi[2] * x + mi[6] * y + mi[10] * z);
}
RT_DEV V3 xf_abs_sum(const float* m, V3 p) {                                                 // shared by :175-189 and :222-242
  float x = p.x, y = p.y, z = p.z;
  return v3(fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z) + fabsf(m[3]),
            fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z) + fabsf(m[7]),
            fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z) + fabsf(m[11]));
}
RT_DEV V3 xf_point_err(const float* m, V3 p, V3& p_err) { p_err = gamma_f(3) * xf_abs_sum(m, p); return xf_point(m, p); }      // :175-189
RT_DEV V3 xf_vector_err(const float* m, V3 v, V3& v_err) { v_err = gamma_f(3) * xf_abs_sum(m, v); return xf_vector(m, v); }   // :222-242 (keeps the |m[i][3]| term)
// AFFINE: m holds rows 0..2 only (row 3 = 0 0 0 1)
template <bool AFFINE = false>
RT_DEV V3 xf_point_with_error(const float* m, V3 p, V3 pe, V3& out_err) {                     // :191-220
  float x = p.x, y = p.y, z = p.z;
  V3 tp = AFFINE ? xf_point_affine(m, p) : xf_point(m, p);
  float e[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    e[i] = (gamma_f(3) + 1.0f) * (fabsf(m[4 * i] * pe.x) + fabsf(m[4 * i + 1] * pe.y) + fabsf(m[4 * i + 2] * pe.z)) +
           gamma_f(3) * (fabsf(m[4 * i] * x) + fabsf(m[4 * i + 1] * y) + fabsf(m[4 * i + 2] * z) + fabsf(m[4 * i + 3]));
  }
  out_err = v3(e[0], e[1], e[2]);
  return tp;
}
// ray.rs:46-71
RT_DEV Ray ray_transform(const Ray& r, const float* m, V3& o_error, V3& d_error) {
  V3 o = xf_point_err(m, r.o, o_error);
  V3 d = xf_vector_err(m, r.d, d_error);
  float ls = length_squared(d);
  if (ls > 0.0f) {
    float dt = dot(vabs(d), o_error) / ls;
    o = o + d * dt;
  }
  return make_ray(o, d, r.t_max);
}

// ---- efloat.rs ---------------------------------------------------------------------------------------
struct EFloat { float v, low, high; };
RT_DEV EFloat ef(float v, float err) {                                                       // :16-27
  EFloat e; e.v = v;
  if (err == 0.0f) { e.low = v; e.high = v; }
  else { e.low = next_float_down(v - err); e.high = next_float_up(v + err); }
  return e;
}
RT_DEV EFloat ef_raw(float v, float lo, float hi) { EFloat e; e.v = v; e.low = lo; e.high = hi; return e; }
RT_DEV EFloat operator+(EFloat a, EFloat f) { return ef_raw(a.v + f.v, next_float_down(a.low + f.low), next_float_up(a.high + f.high)); }   // :129-141
RT_DEV EFloat operator-(EFloat a, EFloat f) { return ef_raw(a.v - f.v, next_float_down(a.low - f.high), next_float_up(a.high - f.low)); }   // :143-155
RT_DEV EFloat operator*(EFloat a, EFloat f) {                                                // :157-183
  float p0 = a.low * f.low, p1 = a.high * f.low, p2 = a.low * f.high, p3 = a.high * f.high;
  return ef_raw(a.v * f.v, next_float_down(fminf(fminf(p0, p1), fminf(p2, p3))), next_float_up(fmaxf(fmaxf(p0, p1), fmaxf(p2, p3))));
}
RT_DEV EFloat operator/(EFloat a, EFloat f) {                                                // :185-210
  float lo, hi;
  if (f.low < 0.0f && f.high > 0.0f) { lo = -inf_f(); hi = inf_f(); }
  else {
    float d0 = a.low / f.low, d1 = a.high / f.low, d2 = a.low / f.high, d3 = a.high / f.high;
    lo = next_float_down(fminf(fminf(d0, d1), fminf(d2, d3)));
    hi = next_float_up(fmaxf(fmaxf(d0, d1), fmaxf(d2, d3)));
  }
  return ef_raw(a.v / f.v, lo, hi);
}
RT_DEV EFloat operator*(float s, EFloat f) { return ef(s, 0.0f) * f; }                       // :268-274
// efloat.rs:97-119 (f64 discriminant and square root)
RT_DEV bool solve_quadratic(EFloat a, EFloat b, EFloat c, EFloat& t0, EFloat& t1) {
  double discrim = (double)b.v * (double)b.v - 4.0 * (double)a.v * (double)c.v;
  if (discrim < 0.0) return false;
  double root_discrim = sqrt(discrim);
  EFloat frd = ef((float)root_discrim, kMachineEpsilon * (float)root_discrim);
  EFloat q = (b.v < 0.0f) ? (-0.5f * (b - frd)) : (-0.5f * (b + frd));
  t0 = q / a;
  t1 = c / q;
  if (t0.v > t1.v) { EFloat tmp = t0; t0 = t1; t1 = tmp; }
  return true;
}

// lib.rs:171-189 find_interval over a float array with predicate `a[i] <= x`
RT_DEV int find_interval_le(const float* a, int size, float x) {
  int first = 0, len = size;
  while (len > 0) {
    int half = len >> 1, middle = first + half;
    if (a[middle] <= x) { first = middle + 1; len -= half + 1; }
    else len = half;
  }
  int r = first - 1;
  return r < 0 ? 0 : (r > size - 2 ? size - 2 : r);
}

}  // namespace rt
