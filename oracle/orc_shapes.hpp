// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Shapes, interactions, sampling
// helpers.  Paths cited are relative to /root/reference/rustracer-core/src/.
#pragma once
#include "orc_math.hpp"
#include <memory>
#include <vector>

namespace orc {

// Per-thread counters restating the reference's stats (scene.rs:9-16, shapes/mesh.rs:21, renderer.rs:17).
struct Counters {
  uint64_t regular_rays = 0, shadow_rays = 0, tri_tests = 0, tri_hits = 0, camera_rays = 0;
  uint64_t nodes_visited = 0, prims_tested = 0;   // oracle-only: N̄ / T̄ for the roofline
  void add(const Counters& o) {
    regular_rays += o.regular_rays; shadow_rays += o.shadow_rays; tri_tests += o.tri_tests; tri_hits += o.tri_hits;
    camera_rays += o.camera_rays; nodes_visited += o.nodes_visited; prims_tested += o.prims_tested;
  }
};
inline Counters& tls_counters() { static thread_local Counters c; return c; }

// sampling/mod.rs
inline V3 uniform_sample_sphere(P2 u) {                                          // :14-20
  float z = 1.0f - 2.0f * u.x;
  float r = std::sqrt(fmax_(1.0f - z * z, 0.0f));
  float phi = 2.0f * PI * u.y;
  return V3(r * std::cos(phi), r * std::sin(phi), z);
}
inline P2 concentric_sample_disk(P2 u) {                                         // :28-47
  const float FRAC_PI_4 = FRAC_PI_2 / 2.0f;
  float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
  if (ox == 0.0f && oy == 0.0f) return P2(0.0f, 0.0f);
  float r, theta;
  if (std::fabs(ox) > std::fabs(oy)) { r = ox; theta = FRAC_PI_4 * (oy / ox); }
  else { r = oy; theta = FRAC_PI_2 - FRAC_PI_4 * (ox / oy); }
  return P2(r * std::cos(theta), r * std::sin(theta));
}
inline V3 cosine_sample_hemisphere(P2 u) {                                       // :22-26
  P2 d = concentric_sample_disk(u);
  float z = std::sqrt(fmax_(1.0f - d.x * d.x - d.y * d.y, 0.0f));
  return V3(d.x, d.y, z);
}
inline P2 uniform_sample_triangle(P2 u) { float su0 = std::sqrt(u.x); return P2(1.0f - su0, u.y * su0); } // :49-52
inline float uniform_cone_pdf(float cos_theta_max) { return 1.0f / (2.0f * PI * (1.0f - cos_theta_max)); } // :54-56
inline float power_heuristic(uint32_t nf, float f_pdf, uint32_t ng, float g_pdf) {                          // :58-63
  float f = (float)nf * f_pdf, g = (float)ng * g_pdf;
  return (f * f) / (f * f + g * g);
}

// interaction.rs:17-75
struct Interaction {
  V3 p, p_error, wo, n;
  static Interaction make(V3 p, V3 p_error, V3 wo, V3 n) { Interaction i; i.p = p; i.p_error = p_error; i.wo = normalize(wo); i.n = n; return i; } // :38-45
  static Interaction from_point(V3 p) { Interaction i; i.p = p; return i; }
  Ray spawn_ray(V3 dir) const { return Ray(offset_ray_origin(p, p_error, n, dir), dir); }                                 // :56-60
  Ray spawn_ray_to_interaction(const Interaction& it) const {                                                             // :69-74
    V3 origin = offset_ray_origin(p, p_error, n, it.p - p);
    V3 target = offset_ray_origin(it.p, it.p_error, it.n, origin - it.p);
    return Ray(origin, target - origin, 1.0f - 1e-4f);
  }
};

// interaction.rs:78-147.  dndu / dndv are not stored: every shape of the reference hands a zero to SurfaceInteraction::new
// (mesh.rs:371-372) or goes through SurfaceInteraction::transform, which zeroes them (interaction.rs:168-169, :182-183 — Q15);
// the places that read them (material/mod.rs:66-88, integrator/mod.rs:71-72) use kZeroN.
struct Shape;
struct SurfaceInteraction {
  Interaction hit;
  P2 uv;
  V3 dpdu, dpdv;
  V3 dpdx, dpdy; float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;   // filled by compute_differential (interaction.rs:245-314)
  struct { V3 n, dpdu, dpdv; } shading;
  inline void compute_differential(const Ray& ray);
  inline void set_shading_geometry(V3 dpdus, V3 dpdvs, bool is_orientation_authoritative);
  int prim = -1;           // prim_number of the GeometricPrimitive (primitive.rs:45-51)
  const Shape* shape = nullptr;
  // hit inside an object instance: `isect.primitive` is the instance's inner GeometricPrimitive (primitive.rs:91-97 maps the
  // inner interaction), so the material is the inner one's; `prim` stays the TransformedPrimitive's number.  -2 = not set.
  int material_override = -2;
};

struct Shape {
  bool reverse_orientation = false, swaps_handedness = false;
  virtual ~Shape() {}
  virtual bool intersect(const Ray& ray, SurfaceInteraction& si, float& t) const = 0;
  virtual bool intersect_p(const Ray& ray) const { SurfaceInteraction si; float t; return intersect(ray, si, t); }   // shapes/mod.rs:27-29
  virtual float area() const = 0;
  virtual Bounds3 world_bounds() const = 0;
  virtual void sample(P2 u, Interaction& it, float& pdf) const = 0;
  virtual void sample_si(const Interaction& si, P2 u, Interaction& intr, float& pdf) const {                          // shapes/mod.rs:39-53
    sample(u, intr, pdf);
    V3 wi = intr.p - si.p;
    if (length_squared(wi) == 0.0f) pdf = 0.0f;
    else {
      wi = normalize(wi);
      pdf *= distance_squared(si.p, intr.p) / std::fabs(dot(intr.n, -wi));
      if (std::isinf(pdf)) pdf = 0.0f;
    }
  }
  virtual float pdf_wi(const Interaction& si, V3 wi) const {                                                           // shapes/mod.rs:59-68
    Ray ray = si.spawn_ray(wi);
    SurfaceInteraction il; float t;
    if (intersect(ray, il, t)) return distance_squared(si.p, il.hit.p) / (std::fabs(dot(il.hit.n, -wi)) * area());
    return 0.0f;
  }
};

// interaction.rs:218-242 (dndu / dndv: see above)
inline void SurfaceInteraction::set_shading_geometry(V3 dpdus, V3 dpdvs, bool is_orientation_authoritative) {
  shading.n = normalize(cross(dpdus, dpdvs));
  if (shape->reverse_orientation ^ shape->swaps_handedness) shading.n = shading.n * -1.0f;
  if (is_orientation_authoritative) hit.n = face_forward(hit.n, shading.n);
  else shading.n = face_forward(shading.n, hit.n);
  shading.dpdu = dpdus; shading.dpdv = dpdvs;
}
// transform.rs:382-394
inline bool solve_linear_system2x2(const float A[2][2], float B0, float B1, float& x0, float& x1) {
  float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  if (std::fabs(det) < 1e-10f) return false;
  x0 = (A[1][1] * B0 - A[0][1] * B1) / det;
  x1 = (A[0][0] * B1 - A[1][0] * B0) / det;
  if (std::isnan(x0) || std::isnan(x1)) return false;
  return true;
}
// interaction.rs:245-314
inline void SurfaceInteraction::compute_differential(const Ray& ray) {
  dudx = dvdx = dudy = dvdy = 0.0f; dpdx = V3(0, 0, 0); dpdy = V3(0, 0, 0);
  if (!ray.has_diff) return;
  const V3 n = hit.n, p = hit.p;
  float d = dot(n, V3(p.x, p.y, p.z));
  float tx = -(dot(n, ray.rx_o) - d) / dot(n, ray.rx_d);
  float ty = -(dot(n, ray.ry_o) - d) / dot(n, ray.ry_d);
  if (std::isinf(tx) || std::isnan(tx) || std::isinf(ty) || std::isnan(ty)) return;
  V3 px = ray.rx_o + tx * ray.rx_d, py = ray.ry_o + ty * ray.ry_d;
  dpdx = px - p; dpdy = py - p;
  int dim[2];
  if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.z)) { dim[0] = 1; dim[1] = 2; }
  else if (std::fabs(n.y) > std::fabs(n.z)) { dim[0] = 0; dim[1] = 2; }
  else { dim[0] = 0; dim[1] = 1; }
  const float A[2][2] = {{dpdu[dim[0]], dpdv[dim[0]]}, {dpdu[dim[1]], dpdv[dim[1]]}};
  const float Bx0 = px[dim[0]] - p[dim[0]], Bx1 = px[dim[1]] - p[dim[1]];
  const float By0 = py[dim[0]] - p[dim[0]], By1 = py[dim[1]] - p[dim[1]];
  if (!solve_linear_system2x2(A, Bx0, Bx1, dudx, dvdx)) { dudx = 0.0f; dvdx = 0.0f; }
  if (!solve_linear_system2x2(A, By0, By1, dudy, dvdy)) { dudy = 0.0f; dvdy = 0.0f; }
}

// interaction.rs:103-147
inline SurfaceInteraction make_si(V3 p, V3 p_error, P2 uv, V3 wo, V3 dpdu, V3 dpdv, const Shape* shape) {
  V3 n = normalize(cross(dpdu, dpdv));
  if (shape->reverse_orientation ^ shape->swaps_handedness) n = n * -1.0f;
  SurfaceInteraction si;
  si.hit = Interaction::make(p, p_error, normalize(wo), n);   // wo normalised twice, as in the reference (:120)
  si.uv = uv; si.dpdu = dpdu; si.dpdv = dpdv;
  si.shading.n = n; si.shading.dpdu = dpdu; si.shading.dpdv = dpdv;
  si.shape = shape;
  return si;
}
// interaction.rs:156-190 (dndu/dndv zeroed there — Q15)
inline SurfaceInteraction si_transform(const SurfaceInteraction& s, const Transform& t) {
  SurfaceInteraction r;
  V3 p_err;
  V3 p = t.point_with_error(s.hit.p, s.hit.p_error, p_err);
  r.hit = Interaction::make(p, p_err, normalize(t.vector(s.hit.wo)), normalize(t.normal(s.hit.n)));
  r.uv = s.uv;
  r.dpdu = t.vector(s.dpdu); r.dpdv = t.vector(s.dpdv);
  r.shading.n = normalize(t.normal(s.shading.n));
  r.shading.dpdu = t.vector(s.shading.dpdu); r.shading.dpdv = t.vector(s.shading.dpdv);
  r.shading.n = face_forward(r.shading.n, r.hit.n);
  r.prim = s.prim; r.shape = s.shape;
  return r;
}

// ---------------------------------------------------------------------------------------
// shapes/mesh.rs
struct TriangleMesh {
  Transform o2w;
  std::vector<int64_t> vi;
  std::vector<V3> p;          // world space (:61)
  std::vector<V3> n, s;       // untransformed (:67-68, Q8)
  std::vector<P2> uv;
  bool has_n = false, has_s = false, has_uv = false;
};

struct Triangle : Shape {
  std::shared_ptr<TriangleMesh> mesh;
  size_t v0i;
  Triangle(std::shared_ptr<TriangleMesh> m, size_t tri, bool rev) : mesh(m), v0i(tri * 3) {          // :181-192
    reverse_orientation = rev; swaps_handedness = m->o2w.swaps_handedness();
  }
  size_t v(int i) const { return (size_t)mesh->vi[v0i + i]; }
  void get_uvs(P2 uv[3]) const {                                                                    // :199-210
    if (mesh->has_uv) { uv[0] = mesh->uv[v(0)]; uv[1] = mesh->uv[v(1)]; uv[2] = mesh->uv[v(2)]; }
    else { uv[0] = P2(0, 0); uv[1] = P2(1, 0); uv[2] = P2(1, 1); }
  }
  // Shared front half of intersect / intersect_p (:215-319 == :428-536). Outputs b0..b2, t.
  // Also exposes the edge proximity used by the harness to flag near-edge rays.
  bool hit_test(const Ray& ray, float& b0, float& b1, float& b2, float& t, float* edge_prox = nullptr) const {
    tls_counters().tri_tests++;
    const V3 p0 = mesh->p[v(0)], p1 = mesh->p[v(1)], p2 = mesh->p[v(2)];
    V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
    int kz = max_dimension(vabs(ray.d));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    V3 d = permute(ray.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz); p1t = permute(p1t, kx, ky, kz); p2t = permute(p2t, kx, ky, kz);
    float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
      double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
      e0 = (float)(p2typ1tx - p2txp1ty);
      double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
      e1 = (float)(p0typ2tx - p0txp2ty);
      double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
      e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if (edge_prox) *edge_prox = INF;
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray.t_max * det)) || (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray.t_max * det)))
      return false;
    float inv_det = 1.0f / det;
    b0 = e0 * inv_det; b1 = e1 * inv_det; b2 = e2 * inv_det;
    t = t_scaled * inv_det;
    float maxzt = max_component(vabs(V3(p0t.z, p1t.z, p2t.z)));
    float delta_z = gamma_f(3) * maxzt;
    float maxxt = max_component(vabs(V3(p0t.x, p1t.x, p2t.x)));
    float maxyt = max_component(vabs(V3(p0t.y, p1t.y, p2t.y)));
    float delta_x = gamma_f(5) * (maxxt + maxzt);
    float delta_y = gamma_f(5) * (maxyt + maxzt);
    float delta_e = 2.0f * (gamma_f(2) * maxxt * maxyt + delta_y * maxxt + delta_x * maxyt);
    float max_e = max_component(vabs(V3(e0, e1, e2)));
    float delta_t = 3.0f * (gamma_f(3) * max_e * maxzt + delta_e * maxzt + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return false;
    if (edge_prox) *edge_prox = std::fmin(std::fabs(e0), std::fmin(std::fabs(e1), std::fabs(e2))) / std::fabs(det);
    return true;
  }
  bool intersect(const Ray& ray, SurfaceInteraction& out, float& t) const override {               // :215-426
    float b0, b1, b2;
    if (!hit_test(ray, b0, b1, b2, t)) return false;
    const V3 p0 = mesh->p[v(0)], p1 = mesh->p[v(1)], p2 = mesh->p[v(2)];
    V3 dpdu(0, 0, 0), dpdv(0, 0, 0);
    P2 uv[3]; get_uvs(uv);
    float duv02x = uv[0].x - uv[2].x, duv02y = uv[0].y - uv[2].y;
    float duv12x = uv[1].x - uv[2].x, duv12y = uv[1].y - uv[2].y;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = duv02x * duv12y - duv02y * duv12x;
    bool degenerate_uv = std::fabs(determinant) < 1e-8f;
    if (!degenerate_uv) {
      float inv_det = 1.0f / determinant;
      dpdu = (duv12y * dp02 - duv02y * dp12) * inv_det;
      dpdv = (-duv12x * dp02 + duv02x * dp12) * inv_det;
    }
    if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
      V3 a, b; coordinate_system(normalize(cross(p2 - p0, p1 - p0)), a, b);
      dpdu = a; dpdv = b;
    }
    float xs = std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x);
    float ys = std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y);
    float zs = std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z);
    V3 p_error = gamma_f(7) * V3(xs, ys, zs);
    V3 p_hit = p0 * b0 + p1 * b1 + p2 * b2;
    P2 uv_hit(uv[0].x * b0 + uv[1].x * b1 + uv[2].x * b2, uv[0].y * b0 + uv[1].y * b1 + uv[2].y * b2);
    SurfaceInteraction isect = make_si(p_hit, p_error, uv_hit, -ray.d, dpdu, dpdv, this);
    V3 n = normalize(cross(dp02, dp12));
    isect.hit.n = n; isect.shading.n = n;
    V3 ns = mesh->has_n ? normalize(mesh->n[v(0)] * b0 + mesh->n[v(1)] * b1 + mesh->n[v(2)] * b2) : isect.hit.n;
    V3 ss = mesh->has_s ? normalize(mesh->s[v(0)] * b0 + mesh->s[v(1)] * b1 + mesh->s[v(2)] * b2) : normalize(isect.dpdu);
    V3 ts = cross(ss, ns);
    if (length_squared(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
    else { V3 a, b; coordinate_system(ns, a, b); ss = a; ts = b; }
    isect.shading.n = ns; isect.shading.dpdu = ss; isect.shading.dpdv = ts;
    if (mesh->has_n) isect.hit.n = face_forward(isect.hit.n, isect.shading.n);
    else if (reverse_orientation ^ swaps_handedness) { isect.hit.n = -isect.hit.n; isect.shading.n = isect.hit.n; }
    tls_counters().tri_hits++;
    out = isect;
    return true;
  }
  bool intersect_p(const Ray& ray) const override {                                                 // :428-586 (no alpha masks on this path)
    float b0, b1, b2, t;
    if (!hit_test(ray, b0, b1, b2, t)) return false;
    tls_counters().tri_hits++;
    return true;
  }
  float area() const override {                                                                     // :588-594
    V3 p0 = mesh->p[v(0)], p1 = mesh->p[v(1)], p2 = mesh->p[v(2)];
    return 0.5f * length(cross(p1 - p0, p2 - p0));
  }
  Bounds3 world_bounds() const override {                                                           // :603-608
    return bunion_point(Bounds3::from_points(mesh->p[v(0)], mesh->p[v(1)]), mesh->p[v(2)]);
  }
  void sample(P2 u, Interaction& it, float& pdf) const override {                                   // :610-634
    P2 b = uniform_sample_triangle(u);
    V3 p0 = mesh->p[v(0)], p1 = mesh->p[v(1)], p2 = mesh->p[v(2)];
    V3 p = (b.x * p0) + (b.y * p1) + ((1.0f - b.x - b.y) * p2);
    V3 normal = normalize(cross(p1 - p0, p2 - p0));
    if (mesh->has_n) {
      V3 ns = b.x * mesh->n[v(0)] + b.y * mesh->n[v(1)] + (1.0f - b.x - b.y) * mesh->n[v(2)];
      normal = face_forward(normal, ns);
    } else if (reverse_orientation ^ swaps_handedness) normal = normal * -1.0f;
    V3 p_abs_sum = vabs(b.x * p0) + vabs(b.y * p1) + vabs((1.0f - b.x - b.y) * p2);
    V3 p_error = gamma_f(6) * p_abs_sum;
    it = Interaction::make(p, p_error, V3(0, 0, 0), normal);
    pdf = 1.0f / area();
  }
};

// ---------------------------------------------------------------------------------------
// shapes/sphere.rs
struct Sphere : Shape {
  Transform o2w, w2o;
  float radius, z_min, z_max, theta_min, theta_max, phi_max;
  Sphere(const Transform& t, float r, float zmin, float zmax, float phimax, bool rev) {            // :30-51
    o2w = t; w2o = t.inverse(); radius = r;
    z_min = clampv(fmin_(zmin, zmax), -r, r);
    z_max = clampv(fmax_(zmin, zmax), -r, r);
    theta_min = std::acos(clampv(fmin_(zmin, zmax) / r, -1.0f, 1.0f));
    theta_max = std::acos(clampv(fmax_(zmin, zmax) / r, -1.0f, 1.0f));
    phi_max = to_radians(clampv(phimax, 0.0f, 360.0f));
    reverse_orientation = rev; swaps_handedness = t.swaps_handedness();
  }
  bool intersect(const Ray& ray, SurfaceInteraction& out, float& t_out) const override {          // :71-203
    V3 o_err, d_err;
    Ray r = ray_transform(ray, w2o, o_err, d_err);
    EFloat ox(r.o.x, o_err.x), oy(r.o.y, o_err.y), oz(r.o.z, o_err.z);
    EFloat dx(r.d.x, d_err.x), dy(r.d.y, d_err.y), dz(r.d.z, d_err.z);
    EFloat a = dx * dx + dy * dy + dz * dz;
    EFloat b = 2.0f * (dx * ox + dy * oy + dz * oz);
    EFloat c = (ox * ox + oy * oy + oz * oz) - EFloat(radius, 0.0f) * EFloat(radius, 0.0f);
    EFloat t0, t1;
    if (!solve_quadratic(a, b, c, t0, t1)) return false;
    if (t0.upper_bound() > r.t_max || t1.lower_bound() <= 0.0f) return false;
    EFloat t_shape_hit = t0;
    if (t_shape_hit.lower_bound() <= 0.0f) {
      t_shape_hit = t1;
      if (t_shape_hit.upper_bound() > r.t_max) return false;
    }
    V3 p_hit = r.at(t_shape_hit.v);
    p_hit = p_hit * (radius / length(p_hit));
    if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * radius;
    float phi = std::atan2(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * PI;
    if ((z_min > -radius && p_hit.z < z_min) || (z_max < radius && p_hit.z > z_max) || phi > phi_max) {
      if (t_shape_hit.v == t1.v) return false;
      if (t1.upper_bound() > ray.t_max) return false;
      t_shape_hit = t1;
      p_hit = r.at(t_shape_hit.v);
      p_hit = p_hit * (radius / length(p_hit));
      if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * radius;
      phi = std::atan2(p_hit.x, p_hit.y);   // swapped args, as in the reference (:133, Q11)
      if (phi < 0.0f) phi += 2.0f * PI;
      if ((z_min > -radius && p_hit.z < z_min) || (z_max < radius && p_hit.z > z_max) || phi > phi_max) return false;
    }
    float u = phi / phi_max;
    float theta = std::acos(clampv(p_hit.z / radius, -1.0f, 1.0f));
    float v = (theta - theta_min) / (theta_max - theta_min);
    float z_radius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    float inv_z_radius = 1.0f / z_radius;
    float cos_phi = p_hit.x * inv_z_radius, sin_phi = p_hit.y * inv_z_radius;
    V3 dpdu(-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f);
    V3 dpdv = (theta_max - theta_min) * V3(p_hit.z * cos_phi, p_hit.z * sin_phi, -radius * std::sin(theta));
    V3 p_error = gamma_f(5) * vabs(p_hit);
    SurfaceInteraction isect = make_si(p_hit, p_error, P2(u, v), -r.d, dpdu, dpdv, this);
    out = si_transform(isect, o2w);
    t_out = t_shape_hit.v;
    return true;
  }
  float area() const override { return phi_max * radius * (z_max - z_min); }                       // :336-338
  Bounds3 world_bounds() const override {                                                           // :205-225
    Bounds3 bounds;
    V3 b0(-radius, -radius, z_min), b1(radius, radius, z_max);
    bounds.extend(o2w.point(V3(b0.x, b0.y, b0.z)));
    bounds.extend(o2w.point(V3(b1.x, b0.y, b0.z)));
    bounds.extend(o2w.point(V3(b0.x, b1.y, b0.z)));
    bounds.extend(o2w.point(V3(b0.x, b0.y, b1.z)));
    bounds.extend(o2w.point(V3(b1.x, b1.y, b0.z)));
    bounds.extend(o2w.point(V3(b1.x, b0.y, b1.z)));
    bounds.extend(o2w.point(V3(b0.x, b1.y, b1.z)));
    bounds.extend(o2w.point(V3(b1.x, b1.y, b1.z)));
    return bounds;
  }
  void sample(P2 u, Interaction& it, float& pdf) const override {                                   // :227-243 (no reverse flip — Q12)
    V3 p_obj = V3(0, 0, 0) + radius * uniform_sample_sphere(u);
    it = Interaction();
    it.n = normalize(o2w.normal(V3(p_obj.x, p_obj.y, p_obj.z)));
    p_obj = p_obj * radius / distance(p_obj, V3(0, 0, 0));
    V3 p_obj_error = gamma_f(5) * vabs(p_obj);
    it.p = o2w.point_with_error(p_obj, p_obj_error, it.p_error);
    pdf = 1.0f / area();
  }
  void sample_si(const Interaction& si, P2 u, Interaction& it, float& pdf) const override {        // :245-308
    V3 p_center = o2w.point(V3(0, 0, 0));
    V3 p_origin = offset_ray_origin(si.p, si.p_error, si.n, p_center - si.p);
    if (distance_squared(p_origin, p_center) <= radius * radius) {
      sample(u, it, pdf);
      V3 wi = it.p - si.p;
      if (length_squared(wi) == 0.0f) pdf = 0.0f;
      else { wi = normalize(wi); pdf *= distance_squared(si.p, it.p) / std::fabs(dot(it.n, -wi)); }
      if (std::isinf(pdf)) pdf = 0.0f;
      return;
    }
    V3 wc = normalize(p_center - si.p);
    V3 wc_x, wc_y; coordinate_system(wc, wc_x, wc_y);
    float sin_theta_max_2 = radius * radius / distance_squared(si.p, p_center);
    float cos_theta_max = std::sqrt(fmax_(0.0f, 1.0f - sin_theta_max_2));
    float cos_theta = (1.0f - u.x) + u.x * cos_theta_max;
    float sin_theta = std::sqrt(fmax_(0.0f, 1.0f - cos_theta * cos_theta));
    float phi = u.y * 2.0f * PI;
    float dc = distance(si.p, p_center);
    float ds = dc * cos_theta - std::sqrt(fmax_(0.0f, radius * radius - dc * dc * sin_theta * sin_theta));
    float cos_alpha = (dc * dc + radius * radius - ds * ds) / (2.0f * dc * radius);
    float sin_alpha = std::sqrt(fmax_(0.0f, 1.0f - cos_alpha * cos_alpha));
    // geometry/mod.rs:115-125 spherical_direction_vec
    V3 n_world = sin_alpha * std::cos(phi) * (-wc_x) + sin_alpha * std::sin(phi) * (-wc_y) + cos_alpha * (-wc);
    V3 p_world = p_center + radius * V3(n_world.x, n_world.y, n_world.z);
    it = Interaction();
    it.p = p_world;
    it.p_error = gamma_f(5) * vabs(p_world);
    it.n = n_world;
    if (reverse_orientation) it.n = it.n * -1.0f;
    pdf = 1.0f / (2.0f * PI * (1.0f - cos_theta_max));
  }
  float pdf_wi(const Interaction& si, V3 wi) const override {                                       // :310-334
    V3 p_center = o2w.point(V3(0, 0, 0));
    V3 p_origin = offset_ray_origin(si.p, si.p_error, si.n, p_center - si.p);
    if (distance_squared(p_origin, p_center) <= radius * radius) return Shape::pdf_wi(si, wi);
    float sin_theta_max_2 = radius * radius / distance_squared(si.p, p_center);
    float cos_theta_max = std::sqrt(fmax_(0.0f, 1.0f - sin_theta_max_2));
    return uniform_cone_pdf(cos_theta_max);
  }
};

// ---------------------------------------------------------------------------------------
// shapes/disk.rs
struct Disk : Shape {
  Transform o2w, w2o;
  float height, radius, inner_radius, phi_max;
  Disk(float h, float r, float ir, float phimax, const Transform& t, bool rev) {                   // :25-45
    height = h; radius = r; inner_radius = ir; phi_max = to_radians(clampv(phimax, 0.0f, 360.0f));
    o2w = t; w2o = t.inverse(); reverse_orientation = rev; swaps_handedness = t.swaps_handedness();
  }
  bool intersect(const Ray& r, SurfaceInteraction& out, float& t_out) const override {            // :65-120
    V3 oe, de;
    Ray ray = ray_transform(r, w2o, oe, de);
    if (ray.d.z == 0.0f) return false;
    float t_shape_hit = (height - ray.o.z) / ray.d.z;
    if (t_shape_hit <= 0.0f || t_shape_hit > ray.t_max) return false;
    V3 p_hit = ray.at(t_shape_hit);
    float dist2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
    if (dist2 > radius * radius || dist2 < inner_radius * inner_radius) return false;
    float phi = std::atan2(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * PI;
    if (phi > phi_max) return false;
    float u = phi / phi_max;
    float r_hit = std::sqrt(dist2);
    float one_minus_v = (r_hit - inner_radius) / (radius - inner_radius);
    float v = 1.0f - one_minus_v;
    V3 dpdu(-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f);
    V3 dpdv = V3(p_hit.x, p_hit.y, 0.0f) * (radius - inner_radius) / r_hit;
    p_hit.z = height;
    SurfaceInteraction isect = make_si(p_hit, V3(0, 0, 0), P2(u, v), -ray.d, dpdu, dpdv, this);
    out = si_transform(isect, o2w);
    t_out = t_shape_hit;
    return true;
  }
  float area() const override { return phi_max * 0.5f * (radius * radius - inner_radius * inner_radius); } // :156-158
  Bounds3 world_bounds() const override {                                                           // :129-136 (Q14: 2 corners only)
    V3 p1 = o2w.point(V3(-radius, -radius, height)), p2 = o2w.point(V3(radius, radius, height));
    V3 lo(fmin_(p1.x, p2.x), fmin_(p1.y, p2.y), fmin_(p1.z, p2.z)), hi(fmax_(p1.x, p2.x), fmax_(p1.y, p2.y), fmax_(p1.z, p2.z));
    return Bounds3::from_points(lo, hi);
  }
  void sample(P2 u, Interaction& it, float& pdf) const override {                                   // :138-154
    P2 pd = concentric_sample_disk(u);
    V3 p_obj(pd.x * radius, pd.y * radius, height);
    it = Interaction();
    it.n = normalize(o2w.normal(V3(0, 0, 1)));
    if (reverse_orientation) it.n = -it.n;
    it.p = o2w.point_with_error(p_obj, V3(0, 0, 0), it.p_error);
    pdf = 1.0f / area();
  }
};

// ---------------------------------------------------------------------------------------
// shapes/cylinder.rs
struct Cylinder : Shape {
  Transform o2w, w2o;
  float radius, z_min, z_max, phi_max;
  Cylinder(const Transform& t, float r, float zmin, float zmax, float phimax, bool rev) {         // :26-46 (no min/max sort)
    o2w = t; w2o = t.inverse(); radius = r; z_min = zmin; z_max = zmax;
    phi_max = to_radians(clampv(phimax, 0.0f, 360.0f));
    reverse_orientation = rev; swaps_handedness = t.swaps_handedness();
  }
  bool solve(const Ray& r, Ray& ray, V3& p_hit, float& phi, EFloat& t_shape_hit) const {          // common :62-127 / :180-241
    V3 o_err, d_err;
    ray = ray_transform(r, w2o, o_err, d_err);
    EFloat ox(ray.o.x, o_err.x), oy(ray.o.y, o_err.y);
    EFloat dx(ray.d.x, d_err.x), dy(ray.d.y, d_err.y);
    EFloat a = dx * dx + dy * dy;
    EFloat b = 2.0f * (dx * ox + dy * oy);
    EFloat c = ox * ox + oy * oy - EFloat(radius, 0.0f) * EFloat(radius, 0.0f);
    EFloat t0, t1;
    if (!solve_quadratic(a, b, c, t0, t1)) return false;
    if (t0.upper_bound() > ray.t_max || t1.lower_bound() <= 0.0f) return false;
    t_shape_hit = t0;
    if (t_shape_hit.lower_bound() <= 0.0f) {
      t_shape_hit = t1;
      if (t_shape_hit.upper_bound() > ray.t_max) return false;
    }
    p_hit = ray.at(t_shape_hit.v);
    float hit_rad = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    p_hit.x *= radius / hit_rad; p_hit.y *= radius / hit_rad;
    phi = std::atan2(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * PI;
    if (p_hit.z < z_min || p_hit.z > z_max || phi > phi_max) {
      if (t_shape_hit.v == t1.v) return false;
      t_shape_hit = t1;
      if (t1.upper_bound() > ray.t_max) return false;
      p_hit = ray.at(t_shape_hit.v);
      hit_rad = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
      p_hit.x *= radius / hit_rad; p_hit.y *= radius / hit_rad;
      phi = std::atan2(p_hit.y, p_hit.x);
      if (phi < 0.0f) phi += 2.0f * PI;
      if (p_hit.z < z_min || p_hit.z > z_max || phi > phi_max) return false;
    }
    return true;
  }
  bool intersect(const Ray& r, SurfaceInteraction& out, float& t_out) const override {            // :62-176
    Ray ray; V3 p_hit; float phi; EFloat th;
    if (!solve(r, ray, p_hit, phi, th)) return false;
    float u = phi / phi_max;
    float v = (p_hit.z - z_min) / (z_max / z_min);     // as written in the reference (:132, Q13)
    V3 dpdu(-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f);
    V3 dpdv(0.0f, 0.0f, z_max - z_min);
    V3 p_error = gamma_f(3) * V3(std::fabs(p_hit.x), std::fabs(p_hit.y), 0.0f);
    SurfaceInteraction isect = make_si(p_hit, p_error, P2(u, v), -ray.d, dpdu, dpdv, this);
    out = si_transform(isect, o2w);
    t_out = th.v;
    return true;
  }
  bool intersect_p(const Ray& r) const override {                                                   // :178-249
    Ray ray; V3 p_hit; float phi; EFloat th;
    return solve(r, ray, p_hit, phi, th);
  }
  float area() const override { return (z_max - z_min) * radius * phi_max; }                        // :251-253
  Bounds3 world_bounds() const override {                                                           // :50-59
    return o2w.bounds(Bounds3::from_points(V3(-radius, -radius, z_min), V3(radius, radius, z_max)));
  }
  void sample(P2 u, Interaction& it, float& pdf) const override {                                   // :255-274
    float z = Bounds3::lerp1(u.x, z_min, z_max);
    float phi = u.y * phi_max;
    V3 p_obj(radius * std::cos(phi), radius * std::sin(phi), z);
    V3 n = normalize(o2w.normal(V3(p_obj.x, p_obj.y, 0.0f)));
    if (reverse_orientation) n = n * -1.0f;
    float hit_rad = std::sqrt(p_obj.x * p_obj.x + p_obj.y * p_obj.y);
    p_obj.x *= radius / hit_rad; p_obj.y *= radius / hit_rad;
    V3 p_obj_error = gamma_f(3) * V3(std::fabs(p_obj.x), std::fabs(p_obj.y), 0.0f);
    V3 p_error;
    V3 p = o2w.point_with_error(p_obj, p_obj_error, p_error);
    it = Interaction::make(p, p_error, V3(0, 0, 0), n);
    pdf = 1.0f / area();
  }
};

}  // namespace orc
