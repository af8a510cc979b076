// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under rustracer_b200/
// may include, link or call this.  CPU restatement (C++17, scalar, no FMA contraction) of
// rustracer's math layer.  Every function cites the reference file:line it follows
// (paths relative to /root/reference/rustracer-core/src/).
//
// PARITY UNPINNED for BVH hit ids / triangle t / radiance: the reference holds no golden
// vector for those.  What the reference DOES pin (EFloat containment,
// sphere no-re-intersection, Distribution1D KAT, find_interval KAT, Bounds2i order, lexer and
// parser KATs) is restated in tests/.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

constexpr float PI = 3.14159265358979323846f;
constexpr float INV_PI = 0.31830988618379067154f;   // std::f32::consts::FRAC_1_PI
constexpr float FRAC_PI_2 = 1.57079632679489661923f;
constexpr float INF = std::numeric_limits<float>::infinity();
// lib.rs:87-95
constexpr float MACHINE_EPSILON = std::numeric_limits<float>::epsilon() * 0.5f;
constexpr float ONE_MINUS_EPSILON = 0.99999994f;
inline float gamma_f(uint32_t n) { return ((float)n * MACHINE_EPSILON) / (1.0f - (float)n * MACHINE_EPSILON); }

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// lib.rs:226-244
inline float next_float_up(float v) {
  if (std::isinf(v) && v > 0.0f) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = f2u(v);
  if (v >= 0.0f) ui += 1; else ui -= 1;
  return u2f(ui);
}
// lib.rs:246-262
inline float next_float_down(float v) {
  if (std::isinf(v) && v < 0.0f) return v;
  if (v == 0.0f) v = -0.0f;
  uint32_t ui = f2u(v);
  if (v > 0.0f) ui -= 1; else ui += 1;
  return u2f(ui);
}

// lib.rs:191-207: PartialOrd min/max (NOT fminf: `if a<b {a} else {b}`)
template <class T> inline T pmin(T a, T b) { return a < b ? a : b; }
template <class T> inline T pmax(T a, T b) { return a > b ? a : b; }
// lib.rs:264-275
template <class T> inline T clampv(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Rust f32::min/max: NaN-ignoring (== fminf/fmaxf)
inline float fmin_(float a, float b) { return std::fmin(a, b); }
inline float fmax_(float a, float b) { return std::fmax(a, b); }
// Rust `as usize` / `as i32` from f32 saturates, NaN -> 0
inline int64_t f2usize(float f) { if (!(f == f)) return 0; if (f <= 0.0f) return 0; if (f >= 9.2e18f) return INT64_MAX; return (int64_t)f; }
inline int32_t f2i32(float f) { if (!(f == f)) return 0; if (f <= -2147483648.0f) return INT32_MIN; if (f >= 2147483648.0f) return INT32_MAX; return (int32_t)f; }
inline uint32_t f2u32(float f) { if (!(f == f)) return 0; if (f <= 0.0f) return 0; if (f >= 4294967296.0f) return UINT32_MAX; return (uint32_t)f; }
inline float to_radians(float deg) { return deg * (PI / 180.0f); }   // Rust f32::to_radians: value * (PI/180)

// ---------------------------------------------------------------------------------------
// Vector3 / Point3 / Normal3 share one layout here (geometry/vector.rs:227-420 etc.)
struct V3 {
  float x = 0, y = 0, z = 0;
  V3() {}
  V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }   // vector.rs:354-360 true division
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }     // vector.rs:241-243
inline V3 cross(V3 a, V3 b) {                                                   // vector.rs:280-286
  return V3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
inline float length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float length(V3 a) { return std::sqrt(length_squared(a)); }
inline V3 normalize(V3 a) { return a / length(a); }                             // vector.rs:276-278
inline V3 vabs(V3 a) { return V3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline float distance_squared(V3 p1, V3 p2) { return length_squared(p2 - p1); } // geometry/mod.rs:222-224
inline float distance(V3 p1, V3 p2) { return length(p2 - p1); }
// lib.rs:119-135
inline int max_dimension(V3 v) { return v.x > v.y ? (v.x > v.z ? 0 : 2) : (v.y > v.z ? 1 : 2); }
inline float max_component(V3 v) { return fmax_(v.x, fmax_(v.y, v.z)); }        // lib.rs:137-139
inline V3 permute(V3 v, int x, int y, int z) { return V3(v[x], v[y], v[z]); }
// lib.rs:158-168
inline void coordinate_system(V3 v1, V3& v2, V3& v3) {
  if (std::fabs(v1.x) > std::fabs(v1.y)) v2 = V3(-v1.z, 0.0f, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
  else v2 = V3(0.0f, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
  v3 = cross(v1, v2);
}
inline V3 face_forward(V3 v1, V3 v2) { return dot(v1, v2) < 0.0f ? -v1 : v1; } // geometry/mod.rs:127-143

struct P2 { float x = 0, y = 0; P2() {} P2(float x_, float y_) : x(x_), y(y_) {} float operator[](int i) const { return i == 0 ? x : y; } };

// geometry/mod.rs:203-220
inline V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
  float d = dot(vabs(n), p_error);
  V3 offset = d * n;
  if (dot(w, n) < 0.0f) offset = -offset;
  V3 po = p + offset;
  for (int i = 0; i < 3; i++) {
    if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
    else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
  }
  return po;
}

// ---------------------------------------------------------------------------------------
// spectrum.rs:15-165 + operators :222-393 (component-wise RGB)
struct Spectrum {
  float r = 0, g = 0, b = 0;
  Spectrum() {}
  explicit Spectrum(float v) : r(v), g(v), b(v) {}
  Spectrum(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
  bool is_black() const { return r == 0.0f && g == 0.0f && b == 0.0f; }
  bool has_nan() const { return std::isnan(r) || std::isnan(g) || std::isnan(b); }
  float y() const { return 0.212671f * r + 0.715160f * g + 0.072169f * b; }   // spectrum.rs:147-150
  float max_component_value() const { return fmax_(fmax_(r, g), b); }          // spectrum.rs:152-154
  Spectrum clamp0() const { return Spectrum(clampv(r, 0.0f, INF), clampv(g, 0.0f, INF), clampv(b, 0.0f, INF)); } // :156-162
  Spectrum sqrt() const { return Spectrum(std::sqrt(r), std::sqrt(g), std::sqrt(b)); }
};
inline Spectrum operator+(Spectrum a, Spectrum b) { return Spectrum(a.r + b.r, a.g + b.g, a.b + b.b); }
inline Spectrum operator-(Spectrum a, Spectrum b) { return Spectrum(a.r - b.r, a.g - b.g, a.b - b.b); }
inline Spectrum operator*(Spectrum a, Spectrum b) { return Spectrum(a.r * b.r, a.g * b.g, a.b * b.b); }
inline Spectrum operator/(Spectrum a, Spectrum b) { return Spectrum(a.r / b.r, a.g / b.g, a.b / b.b); }
inline Spectrum operator*(Spectrum a, float s) { return Spectrum(a.r * s, a.g * s, a.b * s); }
inline Spectrum operator*(float s, Spectrum a) { return Spectrum(s * a.r, s * a.g, s * a.b); }
inline Spectrum operator/(Spectrum a, float s) { return Spectrum(a.r / s, a.g / s, a.b / s); }
inline Spectrum operator+(Spectrum a, float s) { return Spectrum(a.r + s, a.g + s, a.b + s); }
inline Spectrum operator-(Spectrum a, float s) { return Spectrum(a.r - s, a.g - s, a.b - s); }
inline Spectrum& operator+=(Spectrum& a, Spectrum b) { a = a + b; return a; }
inline void to_xyz(Spectrum s, float xyz[3]) {                                  // spectrum.rs:99-107
  xyz[0] = 0.412453f * s.r + 0.357580f * s.g + 0.180423f * s.b;
  xyz[1] = 0.212671f * s.r + 0.715160f * s.g + 0.072169f * s.b;
  xyz[2] = 0.019334f * s.r + 0.119193f * s.g + 0.950227f * s.b;
}
inline Spectrum from_xyz(const float xyz[3]) {                                  // spectrum.rs:92-97
  return Spectrum(3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2],
                  -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2],
                  0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2]);
}

// ---------------------------------------------------------------------------------------
// geometry/matrix.rs
struct Matrix4 {
  float m[4][4];
  Matrix4() { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = (i == j) ? 1.0f : 0.0f; }
  static Matrix4 from(const float* a) { Matrix4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a[i * 4 + j]; return r; }
  Matrix4 transpose() const { Matrix4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = m[j][i]; return r; }
  // matrix.rs:72-145 Gauss-Jordan, full pivoting, `>=` pivot choice
  Matrix4 inverse() const {
    int indxc[4] = {0, 0, 0, 0}, indxr[4] = {0, 0, 0, 0}, ipiv[4] = {0, 0, 0, 0};
    float minv[4][4];
    std::memcpy(minv, m, sizeof(minv));
    for (int i = 0; i < 4; i++) {
      int irow = 0, icol = 0;
      float big = 0.0f;
      for (int j = 0; j < 4; j++) {
        if (ipiv[j] != 1) {
          for (int k = 0; k < 4; k++) {
            if (ipiv[k] == 0) {
              if (std::fabs(minv[j][k]) >= big) { big = std::fabs(minv[j][k]); irow = j; icol = k; }
            }
          }
        }
      }
      ipiv[icol] += 1;
      if (irow != icol) for (int k = 0; k < 4; k++) std::swap(minv[irow][k], minv[icol][k]);
      indxr[i] = irow; indxc[i] = icol;
      float pivinv = 1.0f / minv[icol][icol];
      minv[icol][icol] = 1.0f;
      for (int j = 0; j < 4; j++) minv[icol][j] *= pivinv;
      for (int j = 0; j < 4; j++) {
        if (j != icol) {
          float save = minv[j][icol];
          minv[j][icol] = 0.0f;
          for (int k = 0; k < 4; k++) minv[j][k] -= minv[icol][k] * save;
        }
      }
    }
    for (int j = 3; j >= 0; j--) {
      if (indxr[j] != indxc[j]) for (int k = 0; k < 4; k++) std::swap(minv[k][indxr[j]], minv[k][indxc[j]]);
    }
    Matrix4 r; std::memcpy(r.m, minv, sizeof(minv)); return r;
  }
};
inline Matrix4 mul(const Matrix4& a, const Matrix4& b) {                        // matrix.rs:154-169
  Matrix4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
    r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
  return r;
}

struct Bounds3 {                                                                // bounds.rs:14-32
  V3 p_min, p_max;
  Bounds3() : p_min(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()),
              p_max(std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()) {}
  Bounds3(V3 a, V3 b) : p_min(a), p_max(b) {}
  static Bounds3 from_point(V3 p) { return Bounds3(p, p); }
  static Bounds3 from_points(V3 a, V3 b) {                                     // bounds.rs:41-46
    return Bounds3(V3(pmin(a.x, b.x), pmin(a.y, b.y), pmin(a.z, b.z)), V3(pmax(a.x, b.x), pmax(a.y, b.y), pmax(a.z, b.z)));
  }
  const V3& operator[](int i) const { return i == 0 ? p_min : p_max; }
  void extend(V3 p) {                                                           // bounds.rs:56-75
    if (p.x < p_min.x) p_min.x = p.x;
    if (p.y < p_min.y) p_min.y = p.y;
    if (p.z < p_min.z) p_min.z = p.z;
    if (p.x > p_max.x) p_max.x = p.x;
    if (p.y > p_max.y) p_max.y = p.y;
    if (p.z > p_max.z) p_max.z = p.z;
  }
  int maximum_extent() const {                                                  // bounds.rs:77-90
    V3 v = p_max - p_min;
    return v.x > v.y ? (v.x > v.z ? 0 : 2) : (v.y > v.z ? 1 : 2);
  }
  V3 diagonal() const { return p_max - p_min; }
  float surface_area() const { V3 d = diagonal(); return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); } // bounds.rs:213-216
  V3 offset(V3 p) const {                                                       // bounds.rs:177-190
    V3 o = p - p_min;
    if (p_max.x > p_min.x) o.x /= p_max.x - p_min.x;
    if (p_max.y > p_min.y) o.y /= p_max.y - p_min.y;
    if (p_max.z > p_min.z) o.z /= p_max.z - p_min.z;
    return o;
  }
  static float lerp1(float t, float a, float b) { return a * (1.0f - t) + b * t; }   // lib.rs:107-117
  V3 lerp(V3 t) const { return V3(lerp1(t.x, p_min.x, p_max.x), lerp1(t.y, p_min.y, p_max.y), lerp1(t.z, p_min.z, p_max.z)); }
  bool inside(V3 p) const { return p.x >= p_min.x && p.x <= p_max.x && p.y >= p_min.y && p.y <= p_max.y && p.z >= p_min.z && p.z <= p_max.z; }
  void bounding_sphere(V3& c, float& r) const {                                 // bounds.rs:197-211
    c = V3((p_min.x + p_max.x) / 2.0f, (p_min.y + p_max.y) / 2.0f, (p_min.z + p_max.z) / 2.0f);
    r = inside(c) ? length(p_max - c) : 0.0f;
  }
};
inline Bounds3 bunion(const Bounds3& a, const Bounds3& b) {                     // bounds.rs:92-109
  return Bounds3(V3(pmin(a.p_min.x, b.p_min.x), pmin(a.p_min.y, b.p_min.y), pmin(a.p_min.z, b.p_min.z)),
                 V3(pmax(a.p_max.x, b.p_max.x), pmax(a.p_max.y, b.p_max.y), pmax(a.p_max.z, b.p_max.z)));
}
inline Bounds3 bunion_point(const Bounds3& a, V3 p) { Bounds3 b = a; b.extend(p); return b; } // bounds.rs:111-115

struct Ray {                                                                    // ray.rs:9-44
  V3 o, d; float t_max = INF;
  bool has_diff = false; V3 rx_o, ry_o, rx_d, ry_d;                             // `differential: Option<RayDifferential>` (ray.rs:96-102)
  Ray() {}
  Ray(V3 o_, V3 d_, float t = INF) : o(o_), d(d_), t_max(t) {}
  V3 at(float t) const { return o + t * d; }
  void scale_differentials(float s) {                                           // ray.rs:73-80
    if (!has_diff) return;
    rx_o = o + (rx_o - o) * s; ry_o = o + (ry_o - o) * s;
    rx_d = d + (rx_d - d) * s; ry_d = d + (ry_d - d) * s;
  }
};

// bounds.rs:127-157 — no (1+2γ3) widening, NaN compares false
inline bool bounds_intersect_p_fast(const Bounds3& b, const Ray& ray, V3 inv_dir, const int dir_is_neg[3]) {
  float tmin = (b[dir_is_neg[0]].x - ray.o.x) * inv_dir.x;
  float tmax = (b[1 - dir_is_neg[0]].x - ray.o.x) * inv_dir.x;
  float tymin = (b[dir_is_neg[1]].y - ray.o.y) * inv_dir.y;
  float tymax = (b[1 - dir_is_neg[1]].y - ray.o.y) * inv_dir.y;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  float tzmin = (b[dir_is_neg[2]].z - ray.o.z) * inv_dir.z;
  float tzmax = (b[1 - dir_is_neg[2]].z - ray.o.z) * inv_dir.z;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  return tmin < ray.t_max && tmax > 0.0f;
}

// ---------------------------------------------------------------------------------------
// transform.rs
struct Transform {
  Matrix4 m, m_inv;
  Transform() {}
  Transform(const Matrix4& a, const Matrix4& b) : m(a), m_inv(b) {}
  static Transform from_matrix(const Matrix4& a) { return Transform(a, a.inverse()); }   // :23-28
  Transform inverse() const { return Transform(m_inv, m); }                               // :168-173
  static Transform translate(V3 d) {                                                      // :69-79
    Matrix4 a, b;
    a.m[0][3] = d.x; a.m[1][3] = d.y; a.m[2][3] = d.z;
    b.m[0][3] = -d.x; b.m[1][3] = -d.y; b.m[2][3] = -d.z;
    return Transform(a, b);
  }
  static Transform scale(float sx, float sy, float sz) {                                   // :93-116
    Matrix4 a, b;
    a.m[0][0] = sx; a.m[1][1] = sy; a.m[2][2] = sz;
    b.m[0][0] = 1.0f / sx; b.m[1][1] = 1.0f / sy; b.m[2][2] = 1.0f / sz;
    return Transform(a, b);
  }
  static Transform perspective(float fov, float n, float f) {                              // :155-166
    Matrix4 persp;
    persp.m[2][2] = f / (f - n); persp.m[2][3] = -f * n / (f - n);
    persp.m[3][2] = 1.0f; persp.m[3][3] = 0.0f;
    float inv_tan_ang = 1.0f / std::tan(to_radians(fov) / 2.0f);
    return mulT(scale(inv_tan_ang, inv_tan_ang, 1.0f), from_matrix(persp));
  }
  static Transform mulT(const Transform& a, const Transform& b) { return Transform(mul(a.m, b.m), mul(b.m_inv, a.m_inv)); } // :332-351
  V3 point(V3 p) const {                                                                   // :263-287
    float x = p.x, y = p.y, z = p.z;
    float xp = m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3];
    float yp = m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3];
    float zp = m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3];
    float wp = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
    if (wp == 1.0f) return V3(xp, yp, zp);
    return V3(xp, yp, zp) / wp;
  }
  V3 vector(V3 v) const {                                                                  // :289-304
    float x = v.x, y = v.y, z = v.z;
    return V3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z, m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
              m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
  }
  V3 normal(V3 n) const {                                                                  // :244-254, :306-320
    float x = n.x, y = n.y, z = n.z;
    return V3(m_inv.m[0][0] * x + m_inv.m[1][0] * y + m_inv.m[2][0] * z, m_inv.m[0][1] * x + m_inv.m[1][1] * y + m_inv.m[2][1] * z,
              m_inv.m[0][2] * x + m_inv.m[1][2] * y + m_inv.m[2][2] * z);
  }
  // :175-189
  V3 point_err(V3 p, V3& p_err) const {
    float x = p.x, y = p.y, z = p.z;
    V3 tp = point(p);
    float xs = std::fabs(m.m[0][0] * x) + std::fabs(m.m[0][1] * y) + std::fabs(m.m[0][2] * z) + std::fabs(m.m[0][3]);
    float ys = std::fabs(m.m[1][0] * x) + std::fabs(m.m[1][1] * y) + std::fabs(m.m[1][2] * z) + std::fabs(m.m[1][3]);
    float zs = std::fabs(m.m[2][0] * x) + std::fabs(m.m[2][1] * y) + std::fabs(m.m[2][2] * z) + std::fabs(m.m[2][3]);
    p_err = gamma_f(3) * V3(xs, ys, zs);
    return tp;
  }
  // :191-220
  V3 point_with_error(V3 p, V3 pe, V3& out_err) const {
    float x = p.x, y = p.y, z = p.z;
    V3 tp = point(p);
    float e[3];
    for (int i = 0; i < 3; i++) {
      e[i] = (gamma_f(3) + 1.0f) * (std::fabs(m.m[i][0] * pe.x) + std::fabs(m.m[i][1] * pe.y) + std::fabs(m.m[i][2] * pe.z)) +
             gamma_f(3) * (std::fabs(m.m[i][0] * x) + std::fabs(m.m[i][1] * y) + std::fabs(m.m[i][2] * z) + std::fabs(m.m[i][3]));
    }
    out_err = V3(e[0], e[1], e[2]);
    return tp;
  }
  // :222-242 (note the `+ abs(m[i][3])` term the reference keeps for vectors too)
  V3 vector_err(V3 v, V3& v_err) const {
    float x = v.x, y = v.y, z = v.z;
    V3 tv = vector(v);
    float xs = std::fabs(m.m[0][0] * x) + std::fabs(m.m[0][1] * y) + std::fabs(m.m[0][2] * z) + std::fabs(m.m[0][3]);
    float ys = std::fabs(m.m[1][0] * x) + std::fabs(m.m[1][1] * y) + std::fabs(m.m[1][2] * z) + std::fabs(m.m[1][3]);
    float zs = std::fabs(m.m[2][0] * x) + std::fabs(m.m[2][1] * y) + std::fabs(m.m[2][2] * z) + std::fabs(m.m[2][3]);
    v_err = gamma_f(3) * V3(xs, ys, zs);
    return tv;
  }
  bool swaps_handedness() const {                                                          // :256-262
    const auto& a = m.m;
    float det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    return det < 0.0f;
  }
  Bounds3 bounds(const Bounds3& b) const {                                                 // :353-389
    Bounds3 ret = Bounds3::from_point(point(V3(b.p_min.x, b.p_min.y, b.p_min.z)));
    ret = bunion_point(ret, point(V3(b.p_max.x, b.p_min.y, b.p_min.z)));
    ret = bunion_point(ret, point(V3(b.p_min.x, b.p_max.y, b.p_min.z)));
    ret = bunion_point(ret, point(V3(b.p_min.x, b.p_min.y, b.p_max.z)));
    ret = bunion_point(ret, point(V3(b.p_min.x, b.p_max.y, b.p_max.z)));
    ret = bunion_point(ret, point(V3(b.p_max.x, b.p_max.y, b.p_min.z)));
    ret = bunion_point(ret, point(V3(b.p_max.x, b.p_min.y, b.p_max.z)));
    ret = bunion_point(ret, point(V3(b.p_max.x, b.p_max.y, b.p_max.z)));
    return ret;
  }
};

// ray.rs:46-71
inline Ray ray_transform(const Ray& r, const Transform& t, V3& o_error, V3& d_error) {
  V3 o = t.point_err(r.o, o_error);
  V3 d = t.vector_err(r.d, d_error);
  float ls = length_squared(d);
  if (ls > 0.0f) {
    float dt = dot(vabs(d), o_error) / ls;
    o = o + d * dt;
  }
  Ray out(o, d, r.t_max);
  if (r.has_diff) {                                                             // ray.rs:57-62
    out.has_diff = true;
    out.rx_o = t.point(r.rx_o); out.ry_o = t.point(r.ry_o);
    out.rx_d = t.vector(r.rx_d); out.ry_d = t.vector(r.ry_d);
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// efloat.rs
struct EFloat {
  float v = 0, low = 0, high = 0;
  EFloat() {}
  EFloat(float v_, float err) {                                                  // :16-27
    v = v_;
    if (err == 0.0f) { low = v; high = v; }
    else { low = next_float_down(v - err); high = next_float_up(v + err); }
  }
  static EFloat raw(float v, float lo, float hi) { EFloat e; e.v = v; e.low = lo; e.high = hi; return e; }
  float lower_bound() const { return low; }
  float upper_bound() const { return high; }
};
inline EFloat operator+(EFloat a, EFloat f) {                                    // :129-141
  return EFloat::raw(a.v + f.v, next_float_down(a.low + f.low), next_float_up(a.high + f.high));
}
inline EFloat operator-(EFloat a, EFloat f) {                                    // :143-155
  return EFloat::raw(a.v - f.v, next_float_down(a.low - f.high), next_float_up(a.high - f.low));
}
inline EFloat operator*(EFloat a, EFloat f) {                                    // :157-183
  float prod[4] = {a.low * f.low, a.high * f.low, a.low * f.high, a.high * f.high};
  return EFloat::raw(a.v * f.v, next_float_down(fmin_(fmin_(prod[0], prod[1]), fmin_(prod[2], prod[3]))),
                     next_float_up(fmax_(fmax_(prod[0], prod[1]), fmax_(prod[2], prod[3]))));
}
inline EFloat operator/(EFloat a, EFloat f) {                                    // :185-210
  float lo, hi;
  if (f.low < 0.0f && f.high > 0.0f) { lo = -INF; hi = INF; }
  else {
    float d[4] = {a.low / f.low, a.high / f.low, a.low / f.high, a.high / f.high};
    lo = next_float_down(fmin_(fmin_(d[0], d[1]), fmin_(d[2], d[3])));
    hi = next_float_up(fmax_(fmax_(d[0], d[1]), fmax_(d[2], d[3])));
  }
  return EFloat::raw(a.v / f.v, lo, hi);
}
inline EFloat operator*(float s, EFloat f) { return EFloat(s, 0.0f) * f; }      // :268-274
inline EFloat ef_abs(EFloat a) {                                                 // :49-71
  if (a.low >= 0.0f) return a;
  if (a.high <= 0.0f) return EFloat::raw(-a.v, -a.high, -a.low);
  return EFloat::raw(std::fabs(a.v), 0.0f, fmax_(-a.low, a.high));
}
inline EFloat ef_sqrt(EFloat a) {                                                // :40-47
  return EFloat::raw(std::sqrt(a.v), next_float_down(std::sqrt(a.low)), next_float_up(std::sqrt(a.high)));
}
// efloat.rs:97-119
inline bool solve_quadratic(EFloat a, EFloat b, EFloat c, EFloat& t0, EFloat& t1) {
  double discrim = (double)b.v * (double)b.v - 4.0 * (double)a.v * (double)c.v;
  if (discrim < 0.0) return false;
  double root_discrim = std::sqrt(discrim);
  EFloat frd((float)root_discrim, MACHINE_EPSILON * (float)root_discrim);
  EFloat q = (b.v < 0.0f) ? (-0.5f * (b - frd)) : (-0.5f * (b + frd));
  t0 = q / a;
  t1 = c / q;
  if (t0.v > t1.v) std::swap(t0, t1);
  return true;
}

// lib.rs:171-189
template <class P> inline size_t find_interval(size_t size, P pred) {
  size_t first = 0, len = size;
  while (len > 0) {
    size_t half = len >> 1, middle = first + half;
    if (pred(middle)) { first = middle + 1; len -= half + 1; }
    else len = half;
  }
  return (size_t)clampv<int64_t>((int64_t)first - 1, 0, (int64_t)size - 2);
}
inline bool is_power_of_2(int32_t v) { return (v != 0) && (v & (v - 1)) == 0; }  // lib.rs:209-212
inline int32_t round_up_pow_2(int32_t v) {                                       // lib.rs:214-224
  v -= 1; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1;
}
inline size_t next_power_of_two(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; } // usize::next_power_of_two

}  // namespace orc
