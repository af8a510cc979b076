// Device shape routines: watertight triangle, sphere, disk, cylinder (product code, sm_100a).
// Each routine restates the arithmetic of the rustracer function cited beside it, operation for operation,
// because closest-hit primitive ids must match the reference bit-exactly.
#pragma once
#include "dmath.cuh"
#include "../../../include/rtgpu.h"

#ifndef RT_ENGINE_WIDE4
#define RT_ENGINE_WIDE4 1           // 1: the traversal engine walks 128-byte nodes that hold the FOUR grandchildren of a binary node (two levels of the
#endif                              //    reference's tree collapsed, visit order preserved: trace_engine.cuh); 0: 64-byte nodes with the two children
#ifndef RT_ENGINE_TOP_NODES
#define RT_ENGINE_TOP_NODES 0       // interior nodes of the top levels the traversal engine stages in shared memory (0: none; profiles/r01n)
#endif

namespace rt {

// Device-resident scene (all pointers are device pointers).  Layout: DESIGN.md "Data layout in HBM".
struct DScene {
  const float4* nodes;       // 2 float4 per LinearBVHNode: {min.xyz, bits(offset)}, {max.xyz, bits(n_prims<<2 | axis)}
  const float4* wide;        // RT_ENGINE_WIDE4 == 0: 4 float4 per INTERIOR node (compact numbering): {L.min, bits(ref L)}, {L.max, bits(ref R)},
                             //   {R.min, bits(axis)}, {R.max, 0}; ref = interior index, or 0x80000000 | first slot for a leaf.
                             // RT_ENGINE_WIDE4 == 1: 8 float4 per interior node of EVEN depth below its tree's root (compact numbering), children
                             //   c0, c1 = the left child's children (or the left child itself + an empty entry when it is a leaf), c2, c3 likewise
                             //   for the right child: {c0.min, ref c0}, {c0.max, ref c1}, {c1.min, bits(axis | axisL << 2 | axisR << 4)},
                             //   {c1.max, ref c2}, {c2.min, ref c3}, {c2.max, 0}, {c3.min, 0}, {c3.max, 0}; an empty entry has ref 0xffffffff
  uint32_t root_ref;         // ref of the root node
  uint32_t n_top;            // interior nodes [0, n_top) are the top levels of the scene's tree in breadth-first order (staged in shared memory
                             //   by the traversal engine when it is built with RT_ENGINE_TOP_NODES > 0)
  const float4* geom;        // 3 float4 per ordered slot: triangle v0,v1,v2 (w of v0 = bits(kind | quadric<<2),
                             //   w of v1 bit 0 = last primitive of its leaf)
  const uint4* info;         // per slot: prim_number, material row, light row (0xffffffff none), RTGPU_PRIMFLAG_*
  const float* tri_n;        // 9 per slot or null
  const float* tri_s;        // 9 per slot or null
  const float* tri_uv;       // 6 per slot or null
  const rtgpu_quadric* quadrics;
  const rtgpu_material* materials;
  const rtgpu_lobe* lobes;   // lobe lists of the RTGPU_MAT_LOBES materials (or null)
  const rt_material* texmats;       // neutral material rows for the RTGPU_MAT_TEXTURED materials (or null): evaluated per hit
  const rtgpu_texture* textures;     // texture rows (or null)
  const float* tex_data;             // EWA weight table [0,128) + MIP pyramids
  uint32_t n_textures;
  const rtgpu_instance* instances;   // object instances (or null); n_instances > 0 selects the instance-aware kernels
  uint32_t n_instances;
  const rtgpu_light* lights;
  const float* env;
  uint32_t n_nodes, n_prims, n_quadrics, n_materials, n_lights;
  float world_lo[3], world_hi[3];
  int tune_node_threshold, tune_refill_threshold;   // trace_engine.cuh scheduling knobs (rtgpu_set_option)
};

// What the integrators need of `SurfaceInteraction` (interaction.rs:78-147).  Ray differentials, dndu/dndv and
// uv are dropped: on this path they only feed texture filtering and every texture is constant.
struct SurfHit {
  V3 p, p_error, n, wo;      // Interaction (n = geometric normal after orientation flips)
  V3 ns, dpdu_s;             // shading.n, shading.dpdu
};
// The rest of `SurfaceInteraction` that texture lookups and bump mapping read (interaction.rs:78-99); only the kernels that
// shade textured materials ask for it.  dndu / dndv do not exist: every shape hands zeros to SurfaceInteraction::new
// (mesh.rs:371-372) or goes through SurfaceInteraction::transform, which zeroes them (interaction.rs:168-169, :182-183).
struct SurfTex {
  P2 uv; V3 dpdu, dpdv;      // parametric position and partials
  V3 dpdv_s;                 // shading.dpdv
  bool flip;                 // shape.reverse_orientation ^ shape.transform_swaps_handedness
  V3 dpdx, dpdy; float dudx, dvdx, dudy, dvdy;   // compute_differential (interaction.rs:245-314)
};

// ---- Triangle (shapes/mesh.rs) -----------------------------------------------------------------------
// Front half shared by Triangle::intersect (:215-319) and Triangle::intersect_p (:428-536): returns the
// barycentrics and t of an accepted hit.
// Vector3::permute(kx, ky, kz) for the watertight test's cyclic choice kx = kz + 1, ky = kx + 1 (mesh.rs:229-238): a
// rotation of the components selected by two predicates (six selects per vector instead of three dynamic index chains).
RT_DEV V3 permute_cyclic(V3 v, int kz) {
  const bool r1 = kz == 0, r2 = kz == 1;
  return v3(r1 ? v.y : (r2 ? v.z : v.x), r1 ? v.z : (r2 ? v.x : v.y), r1 ? v.x : (r2 ? v.y : v.z));
}
RT_DEV bool tri_hit_test(V3 p0, V3 p1, V3 p2, const Ray& ray, float& b0, float& b1, float& b2, float& t) {
  V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
  int kz = max_dimension(vabs(ray.d));
  V3 d = permute_cyclic(ray.d, kz);                                   // kx = kz + 1, ky = kx + 1 (mod 3)
  p0t = permute_cyclic(p0t, kz); p1t = permute_cyclic(p1t, kz); p2t = permute_cyclic(p2t, kz);
  float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
  p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
  p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
  p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {                                  // :260-270 f64 products
    double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
    e0 = (float)(p2typ1tx - p2txp1ty);
    double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
    e1 = (float)(p0typ2tx - p0txp2ty);
    double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
    e2 = (float)(p1typ0tx - p1txp0ty);
  }
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
  float maxzt = max_component(vabs(v3(p0t.z, p1t.z, p2t.z)));                    // :300-319 conservative t > delta_t
  float delta_z = gamma_f(3) * maxzt;
  float maxxt = max_component(vabs(v3(p0t.x, p1t.x, p2t.x)));
  float maxyt = max_component(vabs(v3(p0t.y, p1t.y, p2t.y)));
  float delta_x = gamma_f(5) * (maxxt + maxzt);
  float delta_y = gamma_f(5) * (maxyt + maxzt);
  float delta_e = 2.0f * (gamma_f(2) * maxxt * maxyt + delta_y * maxxt + delta_x * maxyt);
  float max_e = max_component(vabs(v3(e0, e1, e2)));
  float delta_t = 3.0f * (gamma_f(3) * max_e * maxzt + delta_e * maxzt + delta_z * max_e) * fabsf(inv_det);
  if (t <= delta_t) return false;
  return true;
}

// Ray-constant part of the watertight test hoisted out of the leaf loop: the permutation and the shear
// depend on the ray only.  Produces the same floating-point values as tri_hit_test (same operations on the
// same operands, only computed once per ray instead of once per triangle).
struct TriRay {
  V3 o; int kx, ky, kz; float sx, sy, sz;
};
RT_DEV TriRay make_tri_ray(const Ray& ray) {
  TriRay tr; tr.o = ray.o;
  tr.kz = max_dimension(vabs(ray.d));
  tr.kx = tr.kz + 1; if (tr.kx == 3) tr.kx = 0;
  tr.ky = tr.kx + 1; if (tr.ky == 3) tr.ky = 0;
  V3 d = permute_cyclic(ray.d, tr.kz);
  tr.sx = -d.x / d.z; tr.sy = -d.y / d.z; tr.sz = 1.0f / d.z;
  return tr;
}
RT_DEV bool tri_hit_test_pre(const TriRay& tr, float t_max, V3 p0, V3 p1, V3 p2, float& b0, float& b1, float& b2, float& t) {
  V3 p0t = p0 - tr.o, p1t = p1 - tr.o, p2t = p2 - tr.o;
  p0t = permute_cyclic(p0t, tr.kz); p1t = permute_cyclic(p1t, tr.kz); p2t = permute_cyclic(p2t, tr.kz);
  p0t.x += tr.sx * p0t.z; p0t.y += tr.sy * p0t.z;
  p1t.x += tr.sx * p1t.z; p1t.y += tr.sy * p1t.z;
  p2t.x += tr.sx * p2t.z; p2t.y += tr.sy * p2t.z;
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
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  p0t.z *= tr.sz; p1t.z *= tr.sz; p2t.z *= tr.sz;
  float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) || (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)))
    return false;
  float inv_det = 1.0f / det;
  b0 = e0 * inv_det; b1 = e1 * inv_det; b2 = e2 * inv_det;
  t = t_scaled * inv_det;
  float maxzt = max_component(vabs(v3(p0t.z, p1t.z, p2t.z)));
  float delta_z = gamma_f(3) * maxzt;
  float maxxt = max_component(vabs(v3(p0t.x, p1t.x, p2t.x)));
  float maxyt = max_component(vabs(v3(p0t.y, p1t.y, p2t.y)));
  float delta_x = gamma_f(5) * (maxxt + maxzt);
  float delta_y = gamma_f(5) * (maxyt + maxzt);
  float delta_e = 2.0f * (gamma_f(2) * maxxt * maxyt + delta_y * maxxt + delta_x * maxyt);
  float max_e = max_component(vabs(v3(e0, e1, e2)));
  float delta_t = 3.0f * (gamma_f(3) * max_e * maxzt + delta_e * maxzt + delta_z * max_e) * fabsf(inv_det);
  return !(t <= delta_t);
}

// Back half of Triangle::intersect (mesh.rs:321-425): the SurfaceInteraction of an accepted hit.
RT_DEV void tri_surface(const DScene& sc, uint32_t slot, uint32_t flags, V3 p0, V3 p1, V3 p2, float b0, float b1, float b2, V3 ray_d, SurfHit& out, SurfTex* ex = nullptr) {
  V3 dpdu = v3(0, 0, 0), dpdv = v3(0, 0, 0);
  P2 uv0 = mk2(0.0f, 0.0f), uv1 = mk2(1.0f, 0.0f), uv2 = mk2(1.0f, 1.0f);                      // :199-210
  if ((flags & RTGPU_PRIMFLAG_HAS_UV) && sc.tri_uv) {
    const float* u = sc.tri_uv + (size_t)slot * 6;
    uv0 = mk2(u[0], u[1]); uv1 = mk2(u[2], u[3]); uv2 = mk2(u[4], u[5]);
  }
  float duv02x = uv0.x - uv2.x, duv02y = uv0.y - uv2.y;
  float duv12x = uv1.x - uv2.x, duv12y = uv1.y - uv2.y;
  V3 dp02 = p0 - p2, dp12 = p1 - p2;
  float determinant = duv02x * duv12y - duv02y * duv12x;
  bool degenerate_uv = fabsf(determinant) < 1e-8f;
  if (!degenerate_uv) {
    float inv_det = 1.0f / determinant;
    dpdu = (duv12y * dp02 - duv02y * dp12) * inv_det;
    dpdv = (-duv12x * dp02 + duv02x * dp12) * inv_det;
  }
  if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
    V3 a, b; coordinate_system(normalize(cross(p2 - p0, p1 - p0)), a, b);
    dpdu = a; dpdv = b;
  }
  float xs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
  float ys = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
  float zs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
  out.p_error = gamma_f(7) * v3(xs, ys, zs);
  out.p = p0 * b0 + p1 * b1 + p2 * b2;
  out.wo = normalize(normalize(-ray_d));                                                     // interaction.rs:120 + :38-45 (normalised twice)
  V3 n = normalize(cross(dp02, dp12));                                                       // :364 overrides the dpdu x dpdv normal
  V3 ns = n, ss;
  const bool has_n = (flags & RTGPU_PRIMFLAG_HAS_N) && sc.tri_n;
  if (has_n) {
    const float* a = sc.tri_n + (size_t)slot * 9;
    ns = normalize(v3(a[0], a[1], a[2]) * b0 + v3(a[3], a[4], a[5]) * b1 + v3(a[6], a[7], a[8]) * b2);
  }
  if ((flags & RTGPU_PRIMFLAG_HAS_S) && sc.tri_s) {
    const float* a = sc.tri_s + (size_t)slot * 9;
    ss = normalize(v3(a[0], a[1], a[2]) * b0 + v3(a[3], a[4], a[5]) * b1 + v3(a[6], a[7], a[8]) * b2);
  } else ss = normalize(dpdu);
  V3 ts = cross(ss, ns);
  if (length_squared(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
  else { V3 a, b; coordinate_system(ns, a, b); ss = a; ts = b; }
  if (has_n) n = face_forward(n, ns);
  else if (flags & RTGPU_PRIMFLAG_FLIP) { n = -n; ns = n; }
  out.n = n; out.ns = ns; out.dpdu_s = ss;
  if (ex) {
    ex->uv = mk2(uv0.x * b0 + uv1.x * b1 + uv2.x * b2, uv0.y * b0 + uv1.y * b1 + uv2.y * b2);   // mesh.rs:352
    ex->dpdu = dpdu; ex->dpdv = dpdv; ex->dpdv_s = ts; ex->flip = (flags & RTGPU_PRIMFLAG_FLIP) != 0;
  }
}

// ---- interaction.rs:103-147 + :156-190 for quadrics: object-space hit -> world-space SurfHit ----------
RT_DEV void quadric_surface(const rtgpu_quadric& q, V3 p_hit, V3 p_error, V3 neg_ray_d_obj, V3 dpdu, V3 dpdv, SurfHit& out, float u = 0.0f, float v = 0.0f,
                            SurfTex* ex = nullptr) {
  V3 n = normalize(cross(dpdu, dpdv));
  if (q.flags & RTGPU_PRIMFLAG_FLIP) n = n * -1.0f;
  V3 wo = normalize(normalize(neg_ray_d_obj));
  // SurfaceInteraction::transform (interaction.rs:156-190)
  out.p = xf_point_with_error(q.o2w, p_hit, p_error, out.p_error);
  out.wo = normalize(normalize(xf_vector(q.o2w, wo)));
  out.n = normalize(xf_normal(q.w2o, n));
  V3 ns = normalize(xf_normal(q.w2o, n));
  out.dpdu_s = xf_vector(q.o2w, dpdu);
  out.ns = face_forward(ns, out.n);
  if (ex) {
    ex->uv = mk2(u, v);
    ex->dpdu = xf_vector(q.o2w, dpdu); ex->dpdv = xf_vector(q.o2w, dpdv); ex->dpdv_s = ex->dpdv;
    ex->flip = (q.flags & RTGPU_PRIMFLAG_FLIP) != 0;
  }
}

// Sphere::intersect (shapes/sphere.rs:71-203).  want_surface=false stops after t is known.
RT_DEV bool sphere_intersect(const rtgpu_quadric& q, const Ray& ray, float& t_out, bool want_surface, SurfHit* out, SurfTex* ex = nullptr) {
  V3 o_err, d_err;
  Ray r = ray_transform(ray, q.w2o, o_err, d_err);
  EFloat ox = ef(r.o.x, o_err.x), oy = ef(r.o.y, o_err.y), oz = ef(r.o.z, o_err.z);
  EFloat dx = ef(r.d.x, d_err.x), dy = ef(r.d.y, d_err.y), dz = ef(r.d.z, d_err.z);
  const float radius = q.radius;
  EFloat a = dx * dx + dy * dy + dz * dz;
  EFloat b = 2.0f * (dx * ox + dy * oy + dz * oz);
  EFloat c = (ox * ox + oy * oy + oz * oz) - ef(radius, 0.0f) * ef(radius, 0.0f);
  EFloat t0, t1;
  if (!solve_quadratic(a, b, c, t0, t1)) return false;
  if (t0.high > r.t_max || t1.low <= 0.0f) return false;
  EFloat t_shape_hit = t0;
  if (t_shape_hit.low <= 0.0f) {
    t_shape_hit = t1;
    if (t_shape_hit.high > r.t_max) return false;
  }
  V3 p_hit = ray_at(r, t_shape_hit.v);
  p_hit = p_hit * (radius / length(p_hit));
  if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * radius;
  float phi = atan2f(p_hit.y, p_hit.x);
  if (phi < 0.0f) phi += 2.0f * kPi;
  if ((q.z_min > -radius && p_hit.z < q.z_min) || (q.z_max < radius && p_hit.z > q.z_max) || phi > q.phi_max) {
    if (t_shape_hit.v == t1.v) return false;
    if (t1.high > ray.t_max) return false;
    t_shape_hit = t1;
    p_hit = ray_at(r, t_shape_hit.v);
    p_hit = p_hit * (radius / length(p_hit));
    if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * radius;
    phi = atan2f(p_hit.x, p_hit.y);                                    // swapped arguments, as in the reference (:133)
    if (phi < 0.0f) phi += 2.0f * kPi;
    if ((q.z_min > -radius && p_hit.z < q.z_min) || (q.z_max < radius && p_hit.z > q.z_max) || phi > q.phi_max) return false;
  }
  t_out = t_shape_hit.v;
  if (!want_surface) return true;
  float theta = acosf(clampf(p_hit.z / radius, -1.0f, 1.0f));
  float z_radius = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
  float inv_z_radius = 1.0f / z_radius;
  float cos_phi = p_hit.x * inv_z_radius, sin_phi = p_hit.y * inv_z_radius;
  V3 dpdu = v3(-q.phi_max * p_hit.y, q.phi_max * p_hit.x, 0.0f);
  V3 dpdv = (q.theta_max - q.theta_min) * v3(p_hit.z * cos_phi, p_hit.z * sin_phi, -radius * sinf(theta));
  V3 p_error = gamma_f(5) * vabs(p_hit);
  quadric_surface(q, p_hit, p_error, -r.d, dpdu, dpdv, *out, phi / q.phi_max, (theta - q.theta_min) / (q.theta_max - q.theta_min), ex);   // u, v: sphere.rs:141-148
  return true;
}

// Disk::intersect (shapes/disk.rs:65-120)
RT_DEV bool disk_intersect(const rtgpu_quadric& q, const Ray& r, float& t_out, bool want_surface, SurfHit* out, SurfTex* ex = nullptr) {
  V3 oe, de;
  Ray ray = ray_transform(r, q.w2o, oe, de);
  if (ray.d.z == 0.0f) return false;
  float t_shape_hit = (q.height - ray.o.z) / ray.d.z;
  if (t_shape_hit <= 0.0f || t_shape_hit > ray.t_max) return false;
  V3 p_hit = ray_at(ray, t_shape_hit);
  float dist2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
  if (dist2 > q.radius * q.radius || dist2 < q.inner_radius * q.inner_radius) return false;
  float phi = atan2f(p_hit.y, p_hit.x);
  if (phi < 0.0f) phi += 2.0f * kPi;
  if (phi > q.phi_max) return false;
  t_out = t_shape_hit;
  if (!want_surface) return true;
  float r_hit = sqrtf(dist2);
  V3 dpdu = v3(-q.phi_max * p_hit.y, q.phi_max * p_hit.x, 0.0f);
  V3 dpdv = v3(p_hit.x, p_hit.y, 0.0f) * (q.radius - q.inner_radius) / r_hit;
  p_hit.z = q.height;
  const float one_minus_v = (r_hit - q.inner_radius) / (q.radius - q.inner_radius);              // disk.rs:89-92
  quadric_surface(q, p_hit, v3(0, 0, 0), -ray.d, dpdu, dpdv, *out, phi / q.phi_max, 1.0f - one_minus_v, ex);
  return true;
}

// Cylinder::intersect / intersect_p (shapes/cylinder.rs:62-176 / :178-249)
RT_DEV bool cylinder_intersect(const rtgpu_quadric& q, const Ray& r, float& t_out, bool want_surface, SurfHit* out, SurfTex* ex = nullptr) {
  V3 o_err, d_err;
  Ray ray = ray_transform(r, q.w2o, o_err, d_err);
  EFloat ox = ef(ray.o.x, o_err.x), oy = ef(ray.o.y, o_err.y);
  EFloat dx = ef(ray.d.x, d_err.x), dy = ef(ray.d.y, d_err.y);
  const float radius = q.radius;
  EFloat a = dx * dx + dy * dy;
  EFloat b = 2.0f * (dx * ox + dy * oy);
  EFloat c = ox * ox + oy * oy - ef(radius, 0.0f) * ef(radius, 0.0f);
  EFloat t0, t1;
  if (!solve_quadratic(a, b, c, t0, t1)) return false;
  if (t0.high > ray.t_max || t1.low <= 0.0f) return false;
  EFloat t_shape_hit = t0;
  if (t_shape_hit.low <= 0.0f) {
    t_shape_hit = t1;
    if (t_shape_hit.high > ray.t_max) return false;
  }
  V3 p_hit = ray_at(ray, t_shape_hit.v);
  float hit_rad = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
  p_hit.x *= radius / hit_rad; p_hit.y *= radius / hit_rad;
  float phi = atan2f(p_hit.y, p_hit.x);
  if (phi < 0.0f) phi += 2.0f * kPi;
  if (p_hit.z < q.z_min || p_hit.z > q.z_max || phi > q.phi_max) {
    if (t_shape_hit.v == t1.v) return false;
    t_shape_hit = t1;
    if (t1.high > ray.t_max) return false;
    p_hit = ray_at(ray, t_shape_hit.v);
    hit_rad = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    p_hit.x *= radius / hit_rad; p_hit.y *= radius / hit_rad;
    phi = atan2f(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * kPi;
    if (p_hit.z < q.z_min || p_hit.z > q.z_max || phi > q.phi_max) return false;
  }
  t_out = t_shape_hit.v;
  if (!want_surface) return true;
  V3 dpdu = v3(-q.phi_max * p_hit.y, q.phi_max * p_hit.x, 0.0f);
  V3 dpdv = v3(0.0f, 0.0f, q.z_max - q.z_min);
  V3 p_error = gamma_f(3) * v3(fabsf(p_hit.x), fabsf(p_hit.y), 0.0f);
  quadric_surface(q, p_hit, p_error, -ray.d, dpdu, dpdv, *out, phi / q.phi_max, (p_hit.z - q.z_min) / (q.z_max / q.z_min), ex);   // v as written in the reference (cylinder.rs:132)
  return true;
}

// A real function, not inlined: the three quadric tests (EFloat intervals, f64 discriminants, phi / z clipping, partials) are thousands of
// instructions, used to be inlined two or three times into every kernel that can meet a quadric, and are cold on triangle scenes
// (profiles/r02f: 877 KB of the matte shade kernel's 978 KB of SASS never executed).
// Translation units whose kernels run it on their hot path for quadric scenes (the traversal engine, the recursive integrators: C2 lost
// 9 % to the call) define RT_QUADRIC_INLINE; the per-material path kernels take the call.
#ifdef RT_QUADRIC_INLINE
RT_DEV bool quadric_intersect(const rtgpu_quadric& q, const Ray& ray, float& t, bool want_surface, SurfHit* out, SurfTex* ex = nullptr) {
#else
static __device__ __noinline__ bool quadric_intersect(const rtgpu_quadric& q, const Ray& ray, float& t, bool want_surface, SurfHit* out, SurfTex* ex = nullptr) {
#endif
  if (q.kind == RTGPU_PRIM_SPHERE) return sphere_intersect(q, ray, t, want_surface, out, ex);
  if (q.kind == RTGPU_PRIM_DISK) return disk_intersect(q, ray, t, want_surface, out, ex);
  return cylinder_intersect(q, ray, t, want_surface, out, ex);
}

// Shape::intersect on one ordered slot with the full surface record (GeometricPrimitive::intersect,
// primitive.rs:45-51).  Used by the shading kernels on the final closest hit and by pdf_wi.
RT_DEV bool slot_intersect_surface(const DScene& sc, uint32_t slot, const Ray& ray, float& t, SurfHit& out, SurfTex* ex = nullptr) {
  float4 g0 = sc.geom[(size_t)slot * 3];
  uint32_t kind_bits = __float_as_uint(g0.w);
  if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
    float4 g1 = sc.geom[(size_t)slot * 3 + 1], g2 = sc.geom[(size_t)slot * 3 + 2];
    V3 p0 = v3(g0), p1 = v3(g1), p2 = v3(g2);
    float b0, b1, b2;
    if (!tri_hit_test(p0, p1, p2, ray, b0, b1, b2, t)) return false;
    tri_surface(sc, slot, sc.info[slot].w, p0, p1, p2, b0, b1, b2, ray.d, out, ex);
    return true;
  }
  return quadric_intersect(sc.quadrics[kind_bits >> 2], ray, t, true, &out, ex);
}

// ---- object instances: TransformedPrimitive (primitive.rs:79-118) -----------------------------------------------------
constexpr uint32_t kNoInst = 0xffffffffu;
// second float4 of a geometry slot, w: bit 0 = last primitive of its leaf, bit 1 = object instance, bits 2..4 = shade queue
// of the slot's material (wave.cuh Q_*; set at upload)
constexpr uint32_t kGeomLastBit = 1u, kGeomInstanceBit = 2u, kGeomClassShift = 2u;
constexpr uint32_t kHitSlotBits = 29;
constexpr int Q_MISS_CLASS = 7;
// material queues of the path integrator (one shade launch per non-empty class)
enum { Q_MATTE = 0, Q_PLASTIC, Q_METAL, Q_GLASS, Q_MIRROR, Q_NONE, Q_LOBES, Q_MISS, Q_COUNT };   // Q_LOBES: uber / substrate / translucent / mix, and every textured material
// material type (rtgpu_material.type, or RTGPU_MAT_NONE for a primitive without material row) -> shade queue
__host__ __device__ __forceinline__ int material_queue(uint32_t type) { return type <= RTGPU_MAT_MIRROR ? (int)type : ((type == RTGPU_MAT_LOBES || type == RTGPU_MAT_TEXTURED) ? Q_LOBES : Q_NONE); }

static_assert(Q_MISS == Q_MISS_CLASS && Q_COUNT <= 8, "the shade-queue id travels in three bits of the hit slot");
// `primitive_to_world.inverse() * ray` (ray.rs:83-93): origin and direction only, no error offset, t_max kept
RT_DEV Ray instance_ray(const rtgpu_instance& I, const Ray& ray) { return make_ray(xf_point_affine(I.w2o, ray.o), xf_vector(I.w2o, ray.d), ray.t_max); }
// SurfaceInteraction::transform (interaction.rs:156-190) for the fields SurfHit keeps
RT_DEV void instance_surface(const rtgpu_instance& I, const SurfHit& o, SurfHit& si, SurfTex* ex = nullptr) {
  if (ex) { ex->dpdu = xf_vector(I.o2w, ex->dpdu); ex->dpdv = xf_vector(I.o2w, ex->dpdv); ex->dpdv_s = xf_vector(I.o2w, ex->dpdv_s); }   // uv, shape kept
  si.p = xf_point_with_error<true>(I.o2w, o.p, o.p_error, si.p_error);
  si.wo = normalize(normalize(xf_vector(I.o2w, o.wo)));                          // `(t * wo).normalize()`, then Interaction::new normalises again
  si.n = normalize(xf_normal(I.w2o, o.n));
  si.ns = normalize(xf_normal(I.w2o, o.ns));
  si.dpdu_s = xf_vector(I.o2w, o.dpdu_s);
  si.ns = face_forward(si.ns, si.n);
}
// The surface record of a final closest hit: slot_intersect_surface, through the instance's transform when the hit lies
// inside an object instance (inst = row of DScene::instances, or kNoInst).
RT_DEV bool hit_surface(const DScene& sc, uint32_t slot, uint32_t inst, const Ray& ray, float& t, SurfHit& si, SurfTex* ex = nullptr) {
  if (inst == kNoInst) return slot_intersect_surface(sc, slot, ray, t, si, ex);
  const rtgpu_instance& I = sc.instances[inst];
  SurfHit o;
  if (!slot_intersect_surface(sc, slot, instance_ray(I, ray), t, o, ex)) return false;
  instance_surface(I, o, si, ex);
  return true;
}

// The same record for a hit whose barycentrics the traversal engine kept (HitRec {b0, slot, b1, b2}, trace_engine.cuh): a triangle's surface follows from
// them directly — they are the values the watertight test produced (mesh.rs:321-329), so nothing is tested twice; a quadric, or a record written by the
// reference walker (have_bary == false), takes the full intersection above.
RT_DEV bool hit_surface_bary(const DScene& sc, uint32_t slot, uint32_t inst, const Ray& ray, bool have_bary, float b0, float b1, float b2, SurfHit& si, SurfTex* ex = nullptr) {
  const float4 g0 = sc.geom[(size_t)slot * 3];
  const uint32_t kind_bits = __float_as_uint(g0.w);
  const bool instanced = inst != kNoInst;
  Ray r = ray;
  if (instanced) r = instance_ray(sc.instances[inst], ray);
  SurfHit o;
  SurfHit& dst = instanced ? o : si;
  float t;
  if ((kind_bits & 3u) == RTGPU_PRIM_TRIANGLE) {
    const float4 g1 = sc.geom[(size_t)slot * 3 + 1], g2 = sc.geom[(size_t)slot * 3 + 2];
    const V3 p0 = v3(g0), p1 = v3(g1), p2 = v3(g2);
    if (!have_bary && !tri_hit_test(p0, p1, p2, r, b0, b1, b2, t)) return false;
    tri_surface(sc, slot, sc.info[slot].w, p0, p1, p2, b0, b1, b2, r.d, dst, ex);
  } else if (!quadric_intersect(sc.quadrics[kind_bits >> 2], r, t, true, &dst, ex)) return false;
  if (instanced) instance_surface(sc.instances[inst], o, si, ex);
  return true;
}

}  // namespace rt
