// Host-side math for the scene front end and the BVH builder (product code).
// Mirrors the arithmetic of rustracer-core/src/{transform.rs, geometry/matrix.rs, bounds.rs} so that
// the matrices, world-space vertices and bounds fed to the device are bit-identical to the reference's.
// Compiled with -ffp-contract=off (Rust never contracts a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>
#include "../../../include/rt_scene.h"

namespace rth {

// std::vector whose resize() leaves new elements uninitialised: the per-primitive and per-node arrays of a 10 M-triangle scene are gigabytes that a
// parallel loop fills right after the allocation; value-initialising them first was a single-threaded pass over every page (profiles/r02t: 0.5 s of 1 s).
template <class T> struct default_init_alloc : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_alloc<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using uvec = std::vector<T, default_init_alloc<T>>;

constexpr float kPi = 3.14159265358979323846f;
inline float radians(float deg) { return deg * (kPi / 180.0f); }   // f32::to_radians
inline float pmin(float a, float b) { return a < b ? a : b; }      // lib.rs:191-207 (PartialOrd, not fminf)
inline float pmax(float a, float b) { return a > b ? a : b; }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
constexpr float kMachineEps = std::numeric_limits<float>::epsilon() * 0.5f;
inline float gamma_n(uint32_t n) { return ((float)n * kMachineEps) / (1.0f - (float)n * kMachineEps); }

struct Vec3 { float x, y, z; float operator[](int i) const { return (&x)[i]; } float& operator[](int i) { return (&x)[i]; } };
inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 sub(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 add(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 scale(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline Vec3 divs(Vec3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline Vec3 cross(Vec3 a, Vec3 b) { return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
inline float len2(Vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float len(Vec3 a) { return std::sqrt(len2(a)); }
inline Vec3 unit(Vec3 a) { return divs(a, len(a)); }               // vector.rs:276-278

// row-major 4x4, m[r*4+c]
struct Mat4 {
  float m[16];
  static Mat4 identity() { Mat4 r; for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f; return r; }
  float& at(int r, int c) { return m[r * 4 + c]; }
  float at(int r, int c) const { return m[r * 4 + c]; }
};
inline Mat4 transpose(const Mat4& a) { Mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.at(i, j) = a.at(j, i); return r; }
inline Mat4 matmul(const Mat4& a, const Mat4& b) {                 // matrix.rs:154-169
  Mat4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
    r.at(i, j) = a.at(i, 0) * b.at(0, j) + a.at(i, 1) * b.at(1, j) + a.at(i, 2) * b.at(2, j) + a.at(i, 3) * b.at(3, j);
  return r;
}
// matrix.rs:72-145: Gauss-Jordan elimination with full pivoting (pivot chosen with >=)
inline Mat4 invert(const Mat4& src) {
  int colIdx[4] = {0, 0, 0, 0}, rowIdx[4] = {0, 0, 0, 0}, used[4] = {0, 0, 0, 0};
  Mat4 w = src;
  for (int step = 0; step < 4; step++) {
    int pr = 0, pc = 0; float best = 0.0f;
    for (int r = 0; r < 4; r++) {
      if (used[r] == 1) continue;
      for (int c = 0; c < 4; c++) {
        if (used[c] != 0) continue;
        float a = std::fabs(w.at(r, c));
        if (a >= best) { best = a; pr = r; pc = c; }
      }
    }
    used[pc] += 1;
    if (pr != pc) for (int c = 0; c < 4; c++) std::swap(w.at(pr, c), w.at(pc, c));
    rowIdx[step] = pr; colIdx[step] = pc;
    float inv = 1.0f / w.at(pc, pc);
    w.at(pc, pc) = 1.0f;
    for (int c = 0; c < 4; c++) w.at(pc, c) *= inv;
    for (int r = 0; r < 4; r++) {
      if (r == pc) continue;
      float f = w.at(r, pc);
      w.at(r, pc) = 0.0f;
      for (int c = 0; c < 4; c++) w.at(r, c) -= w.at(pc, c) * f;
    }
  }
  for (int step = 3; step >= 0; step--)
    if (rowIdx[step] != colIdx[step]) for (int r = 0; r < 4; r++) std::swap(w.at(r, rowIdx[step]), w.at(r, colIdx[step]));
  return w;
}

// transform.rs:9-13
struct Xform {
  Mat4 m, inv;
  static Xform identity() { return Xform{Mat4::identity(), Mat4::identity()}; }
  Xform inverse() const { return Xform{inv, m}; }                  // :168-173
};
inline Xform compose(const Xform& a, const Xform& b) { return Xform{matmul(a.m, b.m), matmul(b.inv, a.inv)}; }   // :332-351
inline Xform from_matrix(const Mat4& a) { return Xform{a, invert(a)}; }
inline Xform translate(Vec3 d) {                                    // :69-79
  Xform t = Xform::identity();
  t.m.at(0, 3) = d.x; t.m.at(1, 3) = d.y; t.m.at(2, 3) = d.z;
  t.inv.at(0, 3) = -d.x; t.inv.at(1, 3) = -d.y; t.inv.at(2, 3) = -d.z;
  return t;
}
inline Xform scaling(float sx, float sy, float sz) {                // :93-116
  Xform t = Xform::identity();
  t.m.at(0, 0) = sx; t.m.at(1, 1) = sy; t.m.at(2, 2) = sz;
  t.inv.at(0, 0) = 1.0f / sx; t.inv.at(1, 1) = 1.0f / sy; t.inv.at(2, 2) = 1.0f / sz;
  return t;
}
inline Xform rotate(float theta_deg, Vec3 axis) {                   // :30-55
  Vec3 a = unit(axis);
  float s = std::sin(radians(theta_deg)), c = std::cos(radians(theta_deg));
  Mat4 m = Mat4::identity();
  m.at(0, 0) = a.x * a.x + (1.0f - a.x * a.x) * c;
  m.at(0, 1) = a.x * a.y * (1.0f - c) - a.z * s;
  m.at(0, 2) = a.x * a.z * (1.0f - c) + a.y * s;
  m.at(0, 3) = 0.0f;
  m.at(1, 0) = a.x * a.y * (1.0f - c) + a.z * s;
  m.at(1, 1) = a.y * a.y + (1.0f - a.y * a.y) * c;
  m.at(1, 2) = a.y * a.z * (1.0f - c) - a.x * s;
  m.at(1, 3) = 0.0f;
  m.at(2, 0) = a.x * a.z * (1.0f - c) - a.y * s;
  m.at(2, 1) = a.y * a.z * (1.0f - c) + a.x * s;
  m.at(2, 2) = a.z * a.z + (1.0f - a.z * a.z) * c;
  m.at(2, 3) = 0.0f;
  return Xform{m, transpose(m)};
}
// :118-153.  Returns false (identity) when up and the viewing direction are parallel.
inline bool look_at(Vec3 pos, Vec3 look, Vec3 up, Xform& out) {
  Mat4 c2w = Mat4::identity();
  c2w.at(0, 3) = pos.x; c2w.at(1, 3) = pos.y; c2w.at(2, 3) = pos.z; c2w.at(3, 3) = 1.0f;
  Vec3 dir = unit(sub(look, pos));
  if (len(cross(unit(up), dir)) == 0.0f) { out = Xform::identity(); return false; }
  Vec3 left = unit(cross(unit(up), dir));
  Vec3 new_up = cross(dir, left);
  c2w.at(0, 0) = left.x; c2w.at(1, 0) = left.y; c2w.at(2, 0) = left.z; c2w.at(3, 0) = 0.0f;
  c2w.at(0, 1) = new_up.x; c2w.at(1, 1) = new_up.y; c2w.at(2, 1) = new_up.z; c2w.at(3, 1) = 0.0f;
  c2w.at(0, 2) = dir.x; c2w.at(1, 2) = dir.y; c2w.at(2, 2) = dir.z; c2w.at(3, 2) = 0.0f;
  out = Xform{invert(c2w), c2w};
  return true;
}
inline Xform perspective(float fov, float n, float f) {             // :155-166
  Mat4 p = Mat4::identity();
  p.at(2, 2) = f / (f - n); p.at(2, 3) = -f * n / (f - n);
  p.at(3, 2) = 1.0f; p.at(3, 3) = 0.0f;
  float inv_tan = 1.0f / std::tan(radians(fov) / 2.0f);
  return compose(scaling(inv_tan, inv_tan, 1.0f), from_matrix(p));
}
inline Vec3 xf_point(const Mat4& m, Vec3 p) {                       // :263-287
  float xp = m.at(0, 0) * p.x + m.at(0, 1) * p.y + m.at(0, 2) * p.z + m.at(0, 3);
  float yp = m.at(1, 0) * p.x + m.at(1, 1) * p.y + m.at(1, 2) * p.z + m.at(1, 3);
  float zp = m.at(2, 0) * p.x + m.at(2, 1) * p.y + m.at(2, 2) * p.z + m.at(2, 3);
  float wp = m.at(3, 0) * p.x + m.at(3, 1) * p.y + m.at(3, 2) * p.z + m.at(3, 3);
  if (wp == 1.0f) return v3(xp, yp, zp);
  return divs(v3(xp, yp, zp), wp);
}
inline Vec3 xf_vector(const Mat4& m, Vec3 v) {                      // :289-304
  return v3(m.at(0, 0) * v.x + m.at(0, 1) * v.y + m.at(0, 2) * v.z, m.at(1, 0) * v.x + m.at(1, 1) * v.y + m.at(1, 2) * v.z,
            m.at(2, 0) * v.x + m.at(2, 1) * v.y + m.at(2, 2) * v.z);
}
inline bool swaps_handedness(const Mat4& m) {                       // :256-262
  float det = m.at(0, 0) * (m.at(1, 1) * m.at(2, 2) - m.at(1, 2) * m.at(2, 1)) - m.at(0, 1) * (m.at(1, 0) * m.at(2, 2) - m.at(1, 2) * m.at(2, 0)) +
              m.at(0, 2) * (m.at(1, 0) * m.at(2, 1) - m.at(1, 1) * m.at(2, 0));
  return det < 0.0f;
}
inline rt_transform to_ir(const Xform& t) { rt_transform r; std::memcpy(r.m, t.m.m, 64); std::memcpy(r.m_inv, t.inv.m, 64); return r; }
inline Xform from_ir(const rt_transform& t) { Xform r; std::memcpy(r.m.m, t.m, 64); std::memcpy(r.inv.m, t.m_inv, 64); return r; }

// bounds.rs:14-32: empty box is (f32::MAX, f32::MIN), not infinities
struct Box3 {
  Vec3 lo{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  Vec3 hi{std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
  void grow(Vec3 p) {                                               // bounds.rs:56-75
    if (p.x < lo.x) lo.x = p.x;
    if (p.y < lo.y) lo.y = p.y;
    if (p.z < lo.z) lo.z = p.z;
    if (p.x > hi.x) hi.x = p.x;
    if (p.y > hi.y) hi.y = p.y;
    if (p.z > hi.z) hi.z = p.z;
  }
  void merge(const Box3& b) {                                       // bounds.rs:92-109
    lo = v3(pmin(lo.x, b.lo.x), pmin(lo.y, b.lo.y), pmin(lo.z, b.lo.z));
    hi = v3(pmax(hi.x, b.hi.x), pmax(hi.y, b.hi.y), pmax(hi.z, b.hi.z));
  }
  int widest_axis() const {                                         // bounds.rs:77-90
    Vec3 d = sub(hi, lo);
    return d.x > d.y ? (d.x > d.z ? 0 : 2) : (d.y > d.z ? 1 : 2);
  }
  float half_area2() const { Vec3 d = sub(hi, lo); return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }   // bounds.rs:213-216
};
inline Box3 box_of_points(Vec3 a, Vec3 b) {                         // bounds.rs:41-46
  Box3 r; r.lo = v3(pmin(a.x, b.x), pmin(a.y, b.y), pmin(a.z, b.z)); r.hi = v3(pmax(a.x, b.x), pmax(a.y, b.y), pmax(a.z, b.z)); return r;
}

}  // namespace rth
