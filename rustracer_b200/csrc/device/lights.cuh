// Device lights: DiffuseAreaLight / PointLight / DistantLight / InfiniteAreaLight sample_li, pdf_li, le
// (light/{diffuse,point,distant,infinite}.rs), the shape sampling they call (shapes/mod.rs:39-68,
// mesh.rs:610-634, sphere.rs:227-334, disk.rs:138-154, cylinder.rs:255-274) and the piecewise-constant
// distributions (sampling/distribution{1d,2d}.rs).  Product code.
#pragma once
#include "bsdf.cuh"

namespace rt {

// interaction.rs:17-75 — the part of Interaction that light sampling needs
struct Inter { V3 p, p_error, n; };
RT_DEV Inter inter_point(V3 p) { Inter i; i.p = p; i.p_error = v3(0, 0, 0); i.n = v3(0, 0, 0); return i; }
RT_DEV Ray spawn_ray(const Inter& it, V3 dir) { return make_ray(offset_ray_origin(it.p, it.p_error, it.n, dir), dir, inf_f()); }   // :56-60
RT_DEV Ray spawn_ray_to(const Inter& a, const Inter& b) {                                                                       // :69-74
  V3 origin = offset_ray_origin(a.p, a.p_error, a.n, b.p - a.p);
  V3 target = offset_ray_origin(b.p, b.p_error, b.n, origin - b.p);
  return make_ray(origin, target - origin, 1.0f - 1e-4f);
}

// ---- Shape::sample (area-measure) ------------------------------------------------------------------------
RT_DEV void tri_vertices(const DScene& sc, uint32_t slot, V3& p0, V3& p1, V3& p2) {
  p0 = v3(sc.geom[(size_t)slot * 3]); p1 = v3(sc.geom[(size_t)slot * 3 + 1]); p2 = v3(sc.geom[(size_t)slot * 3 + 2]);
}
// Triangle::sample (mesh.rs:610-634); `area` = Triangle::area (mesh.rs:588-594), precomputed by the host
RT_DEV void tri_sample(const DScene& sc, uint32_t slot, float area, P2 u, Inter& it, float& pdf) {
  P2 b = uniform_sample_triangle(u);
  V3 p0, p1, p2; tri_vertices(sc, slot, p0, p1, p2);
  const uint32_t flags = sc.info[slot].w;
  V3 p = (b.x * p0) + (b.y * p1) + ((1.0f - b.x - b.y) * p2);
  V3 normal = normalize(cross(p1 - p0, p2 - p0));
  if ((flags & RTGPU_PRIMFLAG_HAS_N) && sc.tri_n) {
    const float* a = sc.tri_n + (size_t)slot * 9;
    V3 ns = b.x * v3(a[0], a[1], a[2]) + b.y * v3(a[3], a[4], a[5]) + (1.0f - b.x - b.y) * v3(a[6], a[7], a[8]);
    normal = face_forward(normal, ns);
  } else if (flags & RTGPU_PRIMFLAG_FLIP) normal = normal * -1.0f;
  V3 p_abs_sum = vabs(b.x * p0) + vabs(b.y * p1) + vabs((1.0f - b.x - b.y) * p2);
  it.p = p; it.p_error = gamma_f(6) * p_abs_sum; it.n = normal;
  pdf = 1.0f / area;
}
// Sphere::sample (sphere.rs:227-243): no reverse-orientation flip
RT_DEV void sphere_sample(const rtgpu_quadric& q, P2 u, Inter& it, float& pdf) {
  V3 p_obj = v3(0, 0, 0) + q.radius * uniform_sample_sphere(u);
  it.n = normalize(xf_normal(q.w2o, p_obj));
  p_obj = p_obj * q.radius / distance(p_obj, v3(0, 0, 0));
  V3 p_obj_error = gamma_f(5) * vabs(p_obj);
  it.p = xf_point_with_error(q.o2w, p_obj, p_obj_error, it.p_error);
  pdf = 1.0f / q.area;
}
// Disk::sample (disk.rs:138-154)
RT_DEV void disk_sample(const rtgpu_quadric& q, P2 u, Inter& it, float& pdf) {
  P2 pd = concentric_sample_disk(u);
  V3 p_obj = v3(pd.x * q.radius, pd.y * q.radius, q.height);
  it.n = normalize(xf_normal(q.w2o, v3(0, 0, 1)));
  if (q.flags & RTGPU_PRIMFLAG_REVERSE) it.n = -it.n;
  it.p = xf_point_with_error(q.o2w, p_obj, v3(0, 0, 0), it.p_error);
  pdf = 1.0f / q.area;
}
// Cylinder::sample (cylinder.rs:255-274)
RT_DEV void cylinder_sample(const rtgpu_quadric& q, P2 u, Inter& it, float& pdf) {
  float z = q.z_min * (1.0f - u.x) + q.z_max * u.x;                         // lerp (lib.rs:107-117)
  float phi = u.y * q.phi_max;
  V3 p_obj = v3(q.radius * cosf(phi), q.radius * sinf(phi), z);
  V3 n = normalize(xf_normal(q.w2o, v3(p_obj.x, p_obj.y, 0.0f)));
  if (q.flags & RTGPU_PRIMFLAG_REVERSE) n = n * -1.0f;
  float hit_rad = sqrtf(p_obj.x * p_obj.x + p_obj.y * p_obj.y);
  p_obj.x *= q.radius / hit_rad; p_obj.y *= q.radius / hit_rad;
  V3 p_obj_error = gamma_f(3) * v3(fabsf(p_obj.x), fabsf(p_obj.y), 0.0f);
  it.p = xf_point_with_error(q.o2w, p_obj, p_obj_error, it.p_error);
  it.n = n;
  pdf = 1.0f / q.area;
}
static __device__ __noinline__ void quadric_sample(const rtgpu_quadric& q, P2 u, Inter& it, float& pdf) {   // out of line: cold on triangle scenes
  if (q.kind == RTGPU_PRIM_SPHERE) sphere_sample(q, u, it, pdf);
  else if (q.kind == RTGPU_PRIM_DISK) disk_sample(q, u, it, pdf);
  else cylinder_sample(q, u, it, pdf);
}
RT_DEV void shape_sample(const DScene& sc, uint32_t slot, float area, P2 u, Inter& it, float& pdf) {
  const uint32_t kind_bits = __float_as_uint(sc.geom[(size_t)slot * 3].w);
  if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) { tri_sample(sc, slot, area, u, it, pdf); return; }
  quadric_sample(sc.quadrics[kind_bits >> 2], u, it, pdf);
}
// Shape::sample_si default (shapes/mod.rs:39-53): area pdf -> solid angle
RT_DEV void area_to_solid_angle(const Inter& ref, const Inter& it, float& pdf, bool check_inf_inside) {
  V3 wi = it.p - ref.p;
  if (length_squared(wi) == 0.0f) pdf = 0.0f;
  else {
    wi = normalize(wi);
    pdf *= distance_squared(ref.p, it.p) / fabsf(dot(it.n, -wi));
    if (check_inf_inside && isinf(pdf)) pdf = 0.0f;
  }
}
// Shape::sample_si (default) and Sphere::sample_si (sphere.rs:245-308)
static __device__ __noinline__ void sphere_sample_si(const rtgpu_quadric& q, const Inter& ref, P2 u, Inter& it, float& pdf) {
  {
    const float radius = q.radius;
    V3 p_center = xf_point(q.o2w, v3(0, 0, 0));
    V3 p_origin = offset_ray_origin(ref.p, ref.p_error, ref.n, p_center - ref.p);
    if (distance_squared(p_origin, p_center) <= radius * radius) {
      sphere_sample(q, u, it, pdf);
      area_to_solid_angle(ref, it, pdf, false);
      if (isinf(pdf)) pdf = 0.0f;                                           // sphere.rs:262-264 (outside the else branch)
      return;
    }
    V3 wc = normalize(p_center - ref.p);
    V3 wc_x, wc_y; coordinate_system(wc, wc_x, wc_y);
    float sin_theta_max_2 = radius * radius / distance_squared(ref.p, p_center);
    float cos_theta_max = sqrtf(fmaxf(0.0f, 1.0f - sin_theta_max_2));
    float cos_theta_v = (1.0f - u.x) + u.x * cos_theta_max;
    float sin_theta_v = sqrtf(fmaxf(0.0f, 1.0f - cos_theta_v * cos_theta_v));
    float phi = u.y * 2.0f * kPi;
    float dc = distance(ref.p, p_center);
    float ds = dc * cos_theta_v - sqrtf(fmaxf(0.0f, radius * radius - dc * dc * sin_theta_v * sin_theta_v));
    float cos_alpha = (dc * dc + radius * radius - ds * ds) / (2.0f * dc * radius);
    float sin_alpha = sqrtf(fmaxf(0.0f, 1.0f - cos_alpha * cos_alpha));
    V3 n_world = sin_alpha * cosf(phi) * (-wc_x) + sin_alpha * sinf(phi) * (-wc_y) + cos_alpha * (-wc);   // geometry/mod.rs:115-125
    V3 p_world = p_center + radius * v3(n_world.x, n_world.y, n_world.z);
    it.p = p_world;
    it.p_error = gamma_f(5) * vabs(p_world);
    it.n = n_world;
    if (q.flags & RTGPU_PRIMFLAG_REVERSE) it.n = it.n * -1.0f;
    pdf = 1.0f / (2.0f * kPi * (1.0f - cos_theta_max));
    return;
  }
}
RT_DEV void shape_sample_si(const DScene& sc, uint32_t slot, float area, const Inter& ref, P2 u, Inter& it, float& pdf) {
  const uint32_t kind_bits = __float_as_uint(sc.geom[(size_t)slot * 3].w);
  if ((kind_bits & 3u) == RTGPU_PRIM_SPHERE) { sphere_sample_si(sc.quadrics[kind_bits >> 2], ref, u, it, pdf); return; }
  shape_sample(sc, slot, area, u, it, pdf);
  area_to_solid_angle(ref, it, pdf, true);
}
// Shape::pdf_wi default (shapes/mod.rs:59-68) and Sphere::pdf_wi (sphere.rs:310-334)
RT_DEV float shape_pdf_wi(const DScene& sc, uint32_t slot, float area, const Inter& ref, V3 wi) {
  const uint32_t kind_bits = __float_as_uint(sc.geom[(size_t)slot * 3].w);
  if ((kind_bits & 3u) == RTGPU_PRIM_SPHERE) {
    const rtgpu_quadric& q = sc.quadrics[kind_bits >> 2];
    V3 p_center = xf_point(q.o2w, v3(0, 0, 0));
    V3 p_origin = offset_ray_origin(ref.p, ref.p_error, ref.n, p_center - ref.p);
    if (!(distance_squared(p_origin, p_center) <= q.radius * q.radius)) {
      float sin_theta_max_2 = q.radius * q.radius / distance_squared(ref.p, p_center);
      float cos_theta_max = sqrtf(fmaxf(0.0f, 1.0f - sin_theta_max_2));
      return uniform_cone_pdf(cos_theta_max);
    }
  }
  Ray ray = spawn_ray(ref, wi);
  SurfHit il; float t;
  if (slot_intersect_surface(sc, slot, ray, t, il)) return distance_squared(ref.p, il.p) / (fabsf(dot(il.n, -wi)) * area);
  return 0.0f;
}

// ---- distributions --------------------------------------------------------------------------------------
// Distribution1D::sample_continuous (distribution1d.rs:51-68) over func[n], cdf[n+1]
RT_DEV float dist1d_sample_continuous(const float* func, const float* cdf, int n, float func_int, float u, float& pdf, int& offset) {
  offset = find_interval_le(cdf, n + 1, u);
  float du = u - cdf[offset];
  if (cdf[offset + 1] - cdf[offset] > 0.0f) du /= cdf[offset + 1] - cdf[offset];
  pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
  return ((float)offset + du) / (float)n;
}
// Distribution1D::sample_discrete (distribution1d.rs:70-79)
RT_DEV int dist1d_sample_discrete(const float* func, const float* cdf, int n, float func_int, float u, float& pdf) {
  int offset = find_interval_le(cdf, n + 1, u);
  pdf = func_int > 0.0f ? func[offset] / (func_int * (float)n) : 0.0f;
  return offset;
}

// ---- InfiniteAreaLight (light/infinite.rs) -------------------------------------------------------------------
// MIPMap::lookup(st, 0.0) == level-0 bilinear with Repeat wrap on this path (mipmap.rs:227-245,285-309)
RT_DEV Spec env_texel(const float* tex, int w, int h, int s, int t) {
  const int ss = s & (w - 1), tt = t & (h - 1);                       // `modulo` (mipmap.rs:455-462): map sides are powers of two (checked by the host)
  const float* p = tex + ((size_t)tt * w + ss) * 3;
  return spec(p[0], p[1], p[2]);
}
RT_DEV Spec env_lookup(const DScene& sc, const rtgpu_light& l, P2 st) {
  const float* tex = sc.env + l.env_texels;
  const int w = (int)l.env_w, h = (int)l.env_h;
  float s = st.x * (float)w - 0.5f, t = st.y * (float)h - 0.5f;
  float fs = floorf(s), ft = floorf(t);
  int s0 = f2i32(fs), t0 = f2i32(ft);
  float ds = s - fs, dt = t - ft;
  return env_texel(tex, w, h, s0, t0) * (1.0f - ds) * (1.0f - dt) + env_texel(tex, w, h, s0, t0 + 1) * (1.0f - ds) * dt +
         env_texel(tex, w, h, s0 + 1, t0) * ds * (1.0f - dt) + env_texel(tex, w, h, s0 + 1, t0 + 1) * ds * dt;
}
RT_DEV V3 m3_vector(const float* m, V3 v) { return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }

// Light::le(ray) (infinite.rs:210-219; black for every other light: light/mod.rs:92-94)
RT_DEV Spec light_le(const DScene& sc, const rtgpu_light& l, V3 ray_d) {
  if (l.kind != RTGPU_LIGHT_INFINITE) return spec(0.0f);
  V3 w = normalize(m3_vector(l.w2l, ray_d));
  P2 st = mk2(spherical_phi(w) * kInvPi * 0.5f, spherical_theta(w) * kInvPi);
  return env_lookup(sc, l, st);
}
// DiffuseAreaLight::l (diffuse.rs:91-97)
RT_DEV Spec area_L(const rtgpu_light& l, V3 n, V3 w) { return (l.two_sided || dot(n, w) > 0.0f) ? spec3(l.I) : spec(0.0f); }
RT_DEV bool light_is_delta(const rtgpu_light& l) { return l.kind == RTGPU_LIGHT_POINT || l.kind == RTGPU_LIGHT_DISTANT; }   // light/mod.rs:38-40

// Light::sample_li -> Li; wi, pdf and p1 (the far end of the VisibilityTester, light/mod.rs:42-56)
RT_DEV Spec light_sample_li(const DScene& sc, const rtgpu_light& l, const Inter& ref, P2 u, V3& wi, float& pdf, Inter& p1) {
  switch (l.kind) {
    case RTGPU_LIGHT_POINT: {                                               // point.rs:43-54
      V3 pos = v3(l.pos[0], l.pos[1], l.pos[2]);
      V3 w = pos - ref.p;
      float r2 = length_squared(w);
      Spec li = spec3(l.I) / (4.0f * kPi * r2);
      p1 = inter_point(pos);
      wi = normalize(w); pdf = 1.0f;
      return li;
    }
    case RTGPU_LIGHT_DISTANT: {                                             // distant.rs:58-71
      V3 dir = v3(l.dir[0], l.dir[1], l.dir[2]);
      p1 = inter_point(ref.p + dir * (2.0f * l.world_radius));
      wi = dir; pdf = 1.0f;
      return spec3(l.I);
    }
    case RTGPU_LIGHT_AREA: {                                                // diffuse.rs:59-70
      shape_sample_si(sc, l.prim_slot, l.area, ref, u, p1, pdf);
      wi = normalize(p1.p - ref.p);
      return area_L(l, p1.n, -wi);
    }
    default: {                                                              // infinite.rs:143-183
      const int W2 = 2 * (int)l.env_w, H2 = 2 * (int)l.env_h;
      float p_marg, p_cond; int v, dummy;
      float d1 = dist1d_sample_continuous(sc.env + l.env_mfunc, sc.env + l.env_mcdf, H2, l.env_mfunc_int, u.y, p_marg, v);
      float d0 = dist1d_sample_continuous(sc.env + l.env_func + (size_t)v * W2, sc.env + l.env_cdf + (size_t)v * (W2 + 1), W2,
                                          sc.env[l.env_func_int + v], u.x, p_cond, dummy);
      float map_pdf = p_cond * p_marg;                                      // distribution2d.rs:30-35
      if (map_pdf == 0.0f) { wi = v3(0, 0, 0); pdf = 0.0f; p1 = inter_point(v3(0, 0, 0)); return spec(0.0f); }
      float theta = d1 * kPi, phi = d0 * 2.0f * kPi;
      float cos_t = cosf(theta), sin_t = sinf(theta), cos_p = cosf(phi), sin_p = sinf(phi);
      wi = m3_vector(l.l2w, v3(sin_t * cos_p, sin_t * sin_p, cos_t));
      pdf = sin_t == 0.0f ? 0.0f : map_pdf / (2.0f * kPi * kPi * sin_t);
      p1 = inter_point(ref.p + wi * (2.0f * l.world_radius));
      return env_lookup(sc, l, mk2(d0, d1));
    }
  }
}
// Light::pdf_li (diffuse.rs:72-74, infinite.rs:185-198; 0 for delta lights)
RT_DEV float light_pdf_li(const DScene& sc, const rtgpu_light& l, const Inter& ref, V3 w) {
  if (l.kind == RTGPU_LIGHT_AREA) return shape_pdf_wi(sc, l.prim_slot, l.area, ref, w);
  if (l.kind == RTGPU_LIGHT_INFINITE) {
    V3 wi = m3_vector(l.w2l, w);
    float theta = spherical_theta(wi), phi = spherical_phi(wi);
    float sin_t = sinf(theta);
    if (sin_t == 0.0f) return 0.0f;
    const int W2 = 2 * (int)l.env_w, H2 = 2 * (int)l.env_h;
    P2 p = mk2(phi * kInvPi * 0.5f, theta * kInvPi);                        // Distribution2D::pdf (distribution2d.rs:37-49)
    int iu = min(max((int)min(f2u32(p.x * (float)W2), 0x7fffffffu), 0), W2 - 1);
    int iv = min(max((int)min(f2u32(p.y * (float)H2), 0x7fffffffu), 0), H2 - 1);
    return (sc.env[l.env_func + (size_t)iv * W2 + iu] / l.env_mfunc_int) / (2.0f * kPi * kPi * sin_t);
  }
  return 0.0f;
}

}  // namespace rt
