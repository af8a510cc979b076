// Textures, MIP-map lookups, noise, bump mapping and per-hit material evaluation on the device (product code, sm_100a).
// Restates rustracer-core/src/{texture/*.rs, mipmap.rs:194-381, noise.rs, material/mod.rs:50-92, interaction.rs:218-314,
// camera.rs:150-202 (the differential part), ray.rs:73-80}; line numbers are relative to rustracer-core/src/.
// The MIP pyramids are built by the host (csrc/host/texture_build.cpp); only kernels that shade RTGPU_MAT_TEXTURED
// materials include any of this in their hot path (the Q_LOBES shade kernel and k_shade_recursive).
#pragma once
#include "bsdf.cuh"
#include "sampler.cuh"
#include "../common/material_lobes.hpp"

namespace rt {

constexpr int kMaxTexDepth = 4;      // nesting of scale / mix / checkerboard textures the evaluator unrolls (host-checked)

// `Ray::differential` (ray.rs:96-102)
struct RayDiff { bool has; V3 rx_o, ry_o, rx_d, ry_d; };
RT_DEV RayDiff no_diff() { RayDiff d; d.has = false; d.rx_o = d.ry_o = d.rx_d = d.ry_d = v3(0, 0, 0); return d; }

// The differential of PerspectiveCamera::generate_ray_differential (camera.rs:150-202) in world space, scaled by
// `1 / sqrt(spp)` (renderer.rs:111, ray.rs:73-80).  `ray` is the camera ray as k_raygen stored it.
RT_DEV RayDiff camera_ray_diff(const float* r2c, const float* c2w, float lens_radius, float focal_distance, P2 p_film, P2 p_lens, const Ray& ray, float scale) {
  const V3 p_camera = xf_point(r2c, v3(p_film.x, p_film.y, 0.0f));
  const V3 dx_camera = xf_point(r2c, v3(1, 0, 0)) - xf_point(r2c, v3(0, 0, 0));            // camera.rs:62-65
  const V3 dy_camera = xf_point(r2c, v3(0, 1, 0)) - xf_point(r2c, v3(0, 0, 0));
  V3 o = v3(0, 0, 0), rx_d, ry_d;
  if (lens_radius > 0.0f) {
    const P2 d = concentric_sample_disk(p_lens);
    o = v3(lens_radius * d.x, lens_radius * d.y, 0.0f);
    const V3 dx = normalize(p_camera + dx_camera);
    const float ft_x = focal_distance / dx.z;
    const V3 p_focus_x = ft_x * dx;
    const V3 dy = normalize(p_camera + dy_camera);
    const float ft_y = focal_distance / dy.z;
    const V3 p_focus_y = ft_y * dy;
    rx_d = normalize(p_focus_x - o); ry_d = normalize(p_focus_y - o);
  } else {
    rx_d = normalize(p_camera + dx_camera); ry_d = normalize(p_camera + dy_camera);
  }
  RayDiff r; r.has = true;
  r.rx_o = xf_point(c2w, o); r.ry_o = r.rx_o;                                               // ray.rs:57-62
  r.rx_d = xf_vector(c2w, rx_d); r.ry_d = xf_vector(c2w, ry_d);
  r.rx_o = ray.o + (r.rx_o - ray.o) * scale; r.ry_o = ray.o + (r.ry_o - ray.o) * scale;     // ray.rs:73-80
  r.rx_d = ray.d + (r.rx_d - ray.d) * scale; r.ry_d = ray.d + (r.ry_d - ray.d) * scale;
  return r;
}

// transform.rs:382-394
RT_DEV bool solve_linear_system2x2(float a00, float a01, float a10, float a11, float b0, float b1, float& x0, float& x1) {
  const float det = a00 * a11 - a01 * a10;
  if (fabsf(det) < 1e-10f) return false;
  x0 = (a11 * b0 - a01 * b1) / det;
  x1 = (a00 * b1 - a10 * b0) / det;
  return !(isnan(x0) || isnan(x1));
}
// SurfaceInteraction::compute_differential (interaction.rs:245-314)
RT_DEV void compute_differential(const SurfHit& si, SurfTex& st, const RayDiff& rd) {
  st.dudx = st.dvdx = st.dudy = st.dvdy = 0.0f; st.dpdx = v3(0, 0, 0); st.dpdy = v3(0, 0, 0);
  if (!rd.has) return;
  const V3 n = si.n, p = si.p;
  const float d = dot(n, v3(p.x, p.y, p.z));
  const float tx = -(dot(n, rd.rx_o) - d) / dot(n, rd.rx_d);
  const float ty = -(dot(n, rd.ry_o) - d) / dot(n, rd.ry_d);
  if (isinf(tx) || isnan(tx) || isinf(ty) || isnan(ty)) return;
  const V3 px = rd.rx_o + tx * rd.rx_d, py = rd.ry_o + ty * rd.ry_d;
  st.dpdx = px - p; st.dpdy = py - p;
  int d0, d1;
  if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { d0 = 1; d1 = 2; }
  else if (fabsf(n.y) > fabsf(n.z)) { d0 = 0; d1 = 2; }
  else { d0 = 0; d1 = 1; }
  const float a00 = comp(st.dpdu, d0), a01 = comp(st.dpdv, d0), a10 = comp(st.dpdu, d1), a11 = comp(st.dpdv, d1);
  const float bx0 = comp(px, d0) - comp(p, d0), bx1 = comp(px, d1) - comp(p, d1);
  const float by0 = comp(py, d0) - comp(p, d0), by1 = comp(py, d1) - comp(p, d1);
  if (!solve_linear_system2x2(a00, a01, a10, a11, bx0, bx1, st.dudx, st.dvdx)) { st.dudx = 0.0f; st.dvdx = 0.0f; }
  if (!solve_linear_system2x2(a00, a01, a10, a11, by0, by1, st.dudy, st.dvdy)) { st.dudy = 0.0f; st.dvdy = 0.0f; }
}
// SurfaceInteraction::set_shading_geometry(.., is_orientation_authoritative = false) (interaction.rs:218-242)
RT_DEV void set_shading_geometry(SurfHit& si, SurfTex& st, V3 dpdus, V3 dpdvs) {
  V3 ns = normalize(cross(dpdus, dpdvs));
  if (st.flip) ns = ns * -1.0f;
  si.ns = face_forward(ns, si.n);
  si.dpdu_s = dpdus; st.dpdv_s = dpdvs;
}

// ---- noise.rs -----------------------------------------------------------------------------------------------------------
static __constant__ uint8_t kNoisePerm[256] = {                              // Ken Perlin's reference permutation (noise.rs:96-119 repeats it twice)
    151, 160, 137, 91,  90,  15,  131, 13,  201, 95,  96,  53,  194, 233, 7,   225, 140, 36,  103, 30,  69,  142, 8,   99,  37,  240,
    21,  10,  23,  190, 6,   148, 247, 120, 234, 75,  0,   26,  197, 62,  94,  252, 219, 203, 117, 35,  11,  32,  57,  177, 33,  88,
    237, 149, 56,  87,  174, 20,  125, 136, 171, 168, 68,  175, 74,  165, 71,  134, 139, 48,  27,  166, 77,  146, 158, 231, 83,  111,
    229, 122, 60,  211, 133, 230, 220, 105, 92,  41,  55,  46,  245, 40,  244, 102, 143, 54,  65,  25,  63,  161, 1,   216, 80,  73,
    209, 76,  132, 187, 208, 89,  18,  169, 200, 196, 135, 130, 116, 188, 159, 86,  164, 100, 109, 198, 173, 186, 3,   64,  52,  217,
    226, 250, 124, 123, 5,   202, 38,  147, 118, 126, 255, 82,  85,  212, 207, 206, 59,  227, 47,  16,  58,  17,  182, 189, 28,  42,
    223, 183, 170, 213, 119, 248, 152, 2,   44,  154, 163, 70,  221, 153, 101, 155, 167, 43,  172, 9,   129, 22,  39,  253, 19,  98,
    108, 110, 79,  113, 224, 232, 178, 185, 112, 104, 218, 246, 97,  228, 251, 34,  242, 193, 238, 210, 144, 12,  191, 179, 162, 241,
    81,  51,  145, 235, 249, 14,  239, 107, 49,  192, 214, 31,  181, 199, 106, 157, 184, 84,  204, 176, 115, 121, 50,  45,  127, 4,
    150, 254, 138, 236, 205, 93,  222, 114, 67,  29,  24,  72,  243, 141, 128, 195, 78,  66,  215, 61,  156, 180};
RT_DEV int noise_perm(int i) { return kNoisePerm[i & 255]; }
RT_DEV float noise_grad(int x, int y, int z, float dx, float dy, float dz) {   // :67-75
  int h = noise_perm(noise_perm(noise_perm(x) + y) + z);
  h &= 15;
  const float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
  const float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
  return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
RT_DEV float noise_weight(float t) { const float t3 = t * t * t, t4 = t3 * t; return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3; }   // :77-82
RT_DEV float lerpf(float t, float a, float b) { return a * (1.0f - t) + b * t; }   // lib.rs:107-117
RT_DEV float noise3(float x, float y, float z) {                       // :7-41
  int ix = f2i32(floorf(x)), iy = f2i32(floorf(y)), iz = f2i32(floorf(z));
  const float dx = x - (float)ix, dy = y - (float)iy, dz = z - (float)iz;
  ix &= 255; iy &= 255; iz &= 255;
  const float w000 = noise_grad(ix, iy, iz, dx, dy, dz), w100 = noise_grad(ix + 1, iy, iz, dx - 1.0f, dy, dz);
  const float w010 = noise_grad(ix, iy + 1, iz, dx, dy - 1.0f, dz), w110 = noise_grad(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
  const float w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz - 1.0f), w101 = noise_grad(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
  const float w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f), w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
  const float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
  const float x00 = lerpf(wx, w000, w100), x10 = lerpf(wx, w010, w110), x01 = lerpf(wx, w001, w101), x11 = lerpf(wx, w011, w111);
  const float y0 = lerpf(wy, x00, x10), y1 = lerpf(wy, x01, x11);
  return lerpf(wz, y0, y1);
}
RT_DEV float smooth_step(float mn, float mx, float value) { const float v = clampf((value - mn) / (mx - mn), 0.0f, 1.0f); return v * v * (-2.0f * v + 3.0f); }   // :84-89
RT_DEV float fbm(V3 p, V3 dpdx, V3 dpdy, float omega, uint32_t max_octaves) {   // :43-61
  const float len2 = fmaxf(length_squared(dpdx), length_squared(dpdy));
  const float n = clampf(-1.0f - 0.5f * log2f(len2), 0.0f, (float)max_octaves);
  const uint32_t n_int = f2u32(floorf(n));
  float sum = 0.0f, lambda = 1.0f, o = 1.0f;
#pragma unroll 1
  for (uint32_t i = 0; i < n_int; i++) {
    const V3 q = lambda * p;
    sum += o * noise3(q.x, q.y, q.z);
    lambda *= 1.99f; o *= omega;
  }
  const float n_partial = n - (float)n_int;
  const V3 q = lambda * p;
  sum += o * smooth_step(0.3f, 0.7f, n_partial) * noise3(q.x, q.y, q.z);
  return sum;
}

// ---- mipmap.rs: lookups --------------------------------------------------------------------------------------------------
// :455-462 `modulo`: every pyramid level has power-of-two sides (MIPMap::new resamples first), so the non-negative remainder is a mask
RT_DEV int mip_modulo(int a, int b) { return a & (b - 1); }
RT_DEV Spec mip_texel(const rtgpu_texture& t, const float* __restrict__ pool, int level, int s, int tt) {   // :194-210
  const int u = t.level_u[level], v = t.level_v[level];
  if (t.wrap == RT_WRAP_REPEAT) { s = mip_modulo(s, u); tt = mip_modulo(tt, v); }
  else if (t.wrap == RT_WRAP_CLAMP) { s = min(max(s, 0), u - 1); tt = min(max(tt, 0), v - 1); }
  else if (s < 0 || s >= u || tt < 0 || tt >= v) return spec(0.0f);
  const float* px = pool + t.level_offset[level] + ((size_t)tt * u + (size_t)s) * t.channels;
  return t.channels == 1 ? spec(px[0]) : spec(px[0], px[1], px[2]);
}
RT_DEV Spec mip_lerp(float x, Spec a, Spec b) { return a * (1.0f - x) + b * x; }
RT_DEV Spec mip_triangle(const rtgpu_texture& t, const float* __restrict__ pool, int level, P2 st) {   // :270-294
  level = min(max(level, 0), t.n_levels - 1);
  const float s = st.x * (float)t.level_u[level] - 0.5f, tt = st.y * (float)t.level_v[level] - 0.5f;
  const int s0 = f2i32(floorf(s)), t0 = f2i32(floorf(tt));
  const float ds = s - (float)s0, dt = tt - (float)t0;
  return mip_texel(t, pool, level, s0, t0) * (1.0f - ds) * (1.0f - dt) + mip_texel(t, pool, level, s0, t0 + 1) * (1.0f - ds) * dt +
         mip_texel(t, pool, level, s0 + 1, t0) * ds * (1.0f - dt) + mip_texel(t, pool, level, s0 + 1, t0 + 1) * ds * dt;
}
RT_DEV Spec mip_lookup(const rtgpu_texture& t, const float* __restrict__ pool, P2 st, float width) {   // :212-229
  const float level = (float)t.n_levels - 1.0f + log2f(fmaxf(width, 1e-8f));
  if (level < 0.0f) return mip_triangle(t, pool, 0, st);
  if (level >= (float)t.n_levels - 1.0f) return mip_texel(t, pool, t.n_levels - 1, 0, 0);
  const float i_level = floorf(level);
  const float delta = level - i_level;
  return mip_lerp(delta, mip_triangle(t, pool, f2i32(i_level), st), mip_triangle(t, pool, f2i32(i_level) + 1, st));
}
RT_DEV Spec mip_ewa(const rtgpu_texture& t, const float* __restrict__ pool, int level, P2 st, P2 dst0, P2 dst1) {   // :296-381
  if (level >= t.n_levels) return mip_texel(t, pool, t.n_levels - 1, 0, 0);
  const float us = (float)t.level_u[level], vs = (float)t.level_v[level];
  st.x = st.x * us - 0.5f; st.y = st.y * vs - 0.5f;
  dst0.x *= us; dst0.y *= vs; dst1.x *= us; dst1.y *= vs;
  float A = dst0.y * dst0.y + dst1.y * dst1.y + 1.0f;
  float B = -2.0f * (dst0.x * dst0.y + dst1.x * dst1.y);
  float C = dst0.x * dst0.x + dst1.x * dst1.x + 1.0f;
  const float inv_f = 1.0f / (A * C - B * B * 0.25f);
  A *= inv_f; B *= inv_f; C *= inv_f;
  const float det = -B * B + 4.0f * A * C;
  const float inv_det = 1.0f / det;
  const float u_sqrt = sqrtf(det * C), v_sqrt = sqrtf(A * det);
  const int s0 = f2i32(ceilf(st.x - 2.0f * inv_det * u_sqrt)), s1 = f2i32(floorf(st.x + 2.0f * inv_det * u_sqrt));
  const int t0 = f2i32(ceilf(st.y - 2.0f * inv_det * v_sqrt)), t1 = f2i32(floorf(st.y + 2.0f * inv_det * v_sqrt));
  Spec sum = spec(0.0f); float sum_wts = 0.0f;
#pragma unroll 1
  for (int it = t0; it < t1 + 1; it++) {
    const float tt = (float)it - st.y;
#pragma unroll 1
    for (int is = s0; is < s1 + 1; is++) {
      const float ss = (float)is - st.x;
      const float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
      if (r2 < 1.0f) {
        const int index = min(f2i32(r2 * 128.0f), 127);
        const float weight = pool[index];                            // WEIGHT_LUT (:35-45), first 128 floats of the pool
        sum = sum + mip_texel(t, pool, level, is, it) * weight;
        sum_wts += weight;
      }
    }
  }
  return sum / sum_wts;
}
RT_DEV Spec mip_lookup_diff(const rtgpu_texture& t, const float* __restrict__ pool, P2 st, P2 dst0, P2 dst1) {   // :231-268
  if (t.trilinear) {
    const float width = fmaxf(fmaxf(fabsf(dst0.x), fabsf(dst0.y)), fmaxf(fabsf(dst1.x), fabsf(dst1.y)));
    return mip_lookup(t, pool, st, 2.0f * width);
  }
  if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { const P2 tmp = dst0; dst0 = dst1; dst1 = tmp; }
  const float major_length = sqrtf(dst0.x * dst0.x + dst0.y * dst0.y);
  float minor_length = sqrtf(dst1.x * dst1.x + dst1.y * dst1.y);
  if ((minor_length * t.max_aniso) < major_length && minor_length > 0.0f) {
    const float scale = major_length / (minor_length * t.max_aniso);
    dst1.x *= scale; dst1.y *= scale;
    minor_length *= scale;
  }
  if (minor_length == 0.0f) return mip_triangle(t, pool, 0, st);
  const float lod = fmaxf(0.0f, (float)t.n_levels - 1.0f + log2f(minor_length));
  const int ilod = f2i32(floorf(lod));
  return mip_lerp(lod - (float)ilod, mip_ewa(t, pool, ilod, st, dst0, dst1), mip_ewa(t, pool, ilod + 1, st, dst0, dst1));
}

// ---- texture/*.rs ----------------------------------------------------------------------------------------------------------
// What a texture reads of the SurfaceInteraction
struct TexPoint { P2 uv; V3 p, dpdx, dpdy; float dudx, dvdx, dudy, dvdy; };
RT_DEV TexPoint tex_point(const SurfHit& si, const SurfTex& st) {
  TexPoint tp; tp.uv = st.uv; tp.p = si.p; tp.dpdx = st.dpdx; tp.dpdy = st.dpdy; tp.dudx = st.dudx; tp.dvdx = st.dvdx; tp.dudy = st.dudy; tp.dvdy = st.dvdy;
  return tp;
}
// TextureMapping2D::map (texture/mod.rs:32-83)
RT_DEV void tex_map2d(const rtgpu_texture& t, const TexPoint& tp, P2& st, P2& dstdx, P2& dstdy) {
  if (t.mapping == RT_TEXMAP_PLANAR) {
    const V3 vs = v3(t.vs[0], t.vs[1], t.vs[2]), vt = v3(t.vt[0], t.vt[1], t.vt[2]);
    st = mk2(t.du + dot(tp.p, vs), t.dv + dot(tp.p, vt));
    dstdx = mk2(dot(tp.dpdx, vs), dot(tp.dpdx, vt));
    dstdy = mk2(dot(tp.dpdy, vs), dot(tp.dpdy, vt));
  } else {
    st = mk2(t.su * tp.uv.x + t.du, t.sv * tp.uv.y + t.dv);
    dstdx = mk2(t.su * tp.dudx, t.sv * tp.dvdx);
    dstdy = mk2(t.su * tp.dudy, t.sv * tp.dvdy);
  }
}
// constant / uv / imagemap / fbm: one real function, called from every nesting level of the combinators below
static __device__ __noinline__ Spec tex_leaf(const rtgpu_texture* __restrict__ textures, const float* __restrict__ pool, int row, const TexPoint& tp) {
  const rtgpu_texture& t = textures[row];
  switch (t.kind) {
    case RT_TEX_CONSTANT: return spec(t.value[0], t.value[1], t.value[2]);                 // constant.rs:36-38
    case RT_TEX_UV: {                                                                        // uv.rs:51-54
      P2 st, dx, dy; tex_map2d(t, tp, st, dx, dy);
      return spec(st.x - floorf(st.x), st.y - floorf(st.y), 0.0f);
    }
    case RT_TEX_IMAGEMAP: {                                                                  // imagemap.rs:231-234
      P2 st, dx, dy; tex_map2d(t, tp, st, dx, dy);
      return mip_lookup_diff(t, pool, st, dx, dy);
    }
    case RT_TEX_FBM: {                                                                       // fbm.rs:19-22, texture/mod.rs:102-110
      const V3 dpdx = xf_vector(t.w2t, tp.dpdx), dpdy = xf_vector(t.w2t, tp.dpdy), p = xf_point(t.w2t, tp.p);
      return spec(fbm(p, dpdx, dpdy, t.omega, (uint32_t)t.octaves));
    }
    default: return spec(0.0f);
  }
}
// Texture<T>::evaluate; a float texture returns its value in all three channels.
// (one real function per nesting level: inlined, the nine child call sites of a level would multiply into megabytes of code)
template <int D>
static __device__ __noinline__ Spec tex_eval(const rtgpu_texture* __restrict__ textures, const float* __restrict__ pool, int row, const TexPoint& tp) {
  const rtgpu_texture& t = textures[row];
  const int kind = t.kind;
  if (kind == RT_TEX_SCALE || kind == RT_TEX_MIX || kind == RT_TEX_CHECKERBOARD) {
    if constexpr (D + 1 < kMaxTexDepth) {
      if (kind == RT_TEX_SCALE) return tex_eval<D + 1>(textures, pool, t.tex1, tp) * tex_eval<D + 1>(textures, pool, t.tex2, tp);   // scale.rs:24-26
      if (kind == RT_TEX_MIX) {                                                              // mix.rs:24-30
        const Spec t1 = tex_eval<D + 1>(textures, pool, t.tex1, tp), t2 = tex_eval<D + 1>(textures, pool, t.tex2, tp);
        const float amt = tex_eval<D + 1>(textures, pool, t.amount, tp).r;
        return t1 * (1.0f - amt) + t2 * amt;
      }
      P2 st, dstdx, dstdy; tex_map2d(t, tp, st, dstdx, dstdy);                               // checkerboard.rs:106-143
      float area2 = 0.0f; bool first = true, blend = false;
      if (t.aa_none) first = ((f2u32(floorf(st.x)) + f2u32(floorf(st.y))) % 2u) == 0u;
      else {
        const float ds = fmaxf(fabsf(dstdx.x), fabsf(dstdy.x)), dt = fmaxf(fabsf(dstdx.y), fabsf(dstdy.y));
        const float s0 = st.x - ds, s1 = st.x + ds, t0 = st.y - dt, t1 = st.y + dt;
        if (floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1)) {
          const int k = (int)((uint32_t)f2i32(floorf(st.x)) + (uint32_t)f2i32(floorf(st.y)));
          first = (k % 2) == 0;
        } else {
          auto bump_int = [](float x) { return floorf(x / 2.0f) + 2.0f * fmaxf(x / 2.0f - floorf(x / 2.0f) - 0.5f, 0.0f); };
          const float sint = (bump_int(s1) - bump_int(s0)) / (2.0f * ds);
          const float tint = (bump_int(t1) - bump_int(t0)) / (2.0f * dt);
          area2 = sint + tint - 2.0f * sint * tint;
          if (ds > 1.0f || dt > 1.0f) area2 = 0.5f;
          blend = true;
        }
      }
      if (!blend) return first ? tex_eval<D + 1>(textures, pool, t.tex1, tp) : tex_eval<D + 1>(textures, pool, t.tex2, tp);
      return tex_eval<D + 1>(textures, pool, t.tex1, tp) * (1.0f - area2) + tex_eval<D + 1>(textures, pool, t.tex2, tp) * area2;
    } else return spec(0.0f);                                                                // deeper graphs are rejected by the host
  }
  return tex_leaf(textures, pool, row, tp);
}
RT_DEV Spec tex_evaluate(const rtgpu_texture* __restrict__ textures, const float* __restrict__ pool, int row, const TexPoint& tp) {
  return tex_eval<0>(textures, pool, row, tp);
}

// material/mod.rs:50-92 (dndu = dndv = 0: see SurfTex)
static __device__ __noinline__ void bump_map(const rtgpu_texture* __restrict__ textures, const float* __restrict__ tex_data, int row, SurfHit& si, SurfTex& st) {
  const V3 zero_n = v3(0, 0, 0);
  TexPoint tp = tex_point(si, st);
  float du = 0.5f * (fabsf(st.dudx) + fabsf(st.dudy));
  if (du == 0.0f) du = 0.0005f;
  tp.p = si.p + du * si.dpdu_s; tp.uv = mk2(st.uv.x + du, st.uv.y + 0.0f);
  const float u_displace = tex_evaluate(textures, tex_data, row, tp).r;
  float dv = 0.5f * (fabsf(st.dvdx) + fabsf(st.dvdy));
  if (dv == 0.0f) dv = 0.0005f;
  tp.p = si.p + dv * st.dpdv_s; tp.uv = mk2(st.uv.x + 0.0f, st.uv.y + dv);
  const float v_displace = tex_evaluate(textures, tex_data, row, tp).r;
  tp.p = si.p; tp.uv = st.uv;
  const float displace = tex_evaluate(textures, tex_data, row, tp).r;
  const V3 dpdu = si.dpdu_s + (u_displace - displace) / du * si.ns + displace * zero_n;
  const V3 dpdv = st.dpdv_s + (v_displace - displace) / dv * si.ns + displace * zero_n;
  set_shading_geometry(si, st, dpdu, dpdv);
}

// The material row with every textured parameter evaluated at the hit (the `.evaluate(si)` calls of material/*.rs);
// `apply_bump`: the material's bump map displaces the caller's shading geometry first.
static __device__ __noinline__ void resolve_material(const rt_material* __restrict__ texmats, const rtgpu_texture* __restrict__ textures, const float* __restrict__ tex_data,
                                                    int row, SurfHit& si, SurfTex& st, bool apply_bump, rt_material& m) {
  m = texmats[row];
  if (!m.textured) return;
  if (apply_bump && m.type != RT_MAT_MIX && m.tex[RT_TS_BUMP]) bump_map(textures, tex_data, m.tex[RT_TS_BUMP] - 1, si, st);
  const TexPoint tp = tex_point(si, st);
#pragma unroll 1
  for (int slot = 0; slot < RT_TS_BUMP; slot++) {
    if (!m.tex[slot]) continue;
    const Spec v = tex_evaluate(textures, tex_data, m.tex[slot] - 1, tp);
    float* d3 = nullptr; float* d1 = nullptr;
    switch (slot) {
      case RT_TS_KD: d3 = m.kd; break;           case RT_TS_KS: d3 = m.ks; break;             case RT_TS_KR: d3 = m.kr; break;
      case RT_TS_KT: d3 = m.kt; break;           case RT_TS_ETA_RGB: d3 = m.eta_rgb; break;   case RT_TS_K_RGB: d3 = m.k_rgb; break;
      case RT_TS_OPACITY: d3 = m.opacity; break; case RT_TS_REFLECT: d3 = m.reflect; break;   case RT_TS_TRANSMIT: d3 = m.transmit; break;
      case RT_TS_AMOUNT: d3 = m.amount; break;
      case RT_TS_SIGMA: d1 = &m.sigma; break;    case RT_TS_ROUGHNESS: d1 = &m.roughness; break;
      case RT_TS_UROUGHNESS: d1 = &m.uroughness; break; case RT_TS_VROUGHNESS: d1 = &m.vroughness; break;
      default: d1 = &m.eta; break;               // RT_TS_ETA
    }
    if (d3) { d3[0] = v.r; d3[1] = v.g; d3[2] = v.b; } else *d1 = v.r;
  }
}

// Material::compute_scattering_functions for an RTGPU_MAT_TEXTURED material: differentials, bump map, parameter textures,
// then the same lobe listing the host uses for constant materials (common/material_lobes.hpp).  `scratch` holds the lobes
// (the Bsdf points at them).  MixMaterial: mat1 sees the caller's surface (its bump map stays), mat2 a clone whose
// shading geometry nobody reads afterwards (mixmat.rs:43-47) — so once a second child has been met no bump map applies.
struct DeviceMixChildren {
  const DScene& sc; SurfHit& si; SurfTex& st; bool live;
  RT_DEV rt_material operator()(int row, bool first) {
    const bool apply = live && first;
    if (!first) live = false;
    rt_material m;
    resolve_material(sc.texmats, sc.textures, sc.tex_data, row, si, st, apply, m);
    return m;
  }
};
static __device__ __noinline__ float list_material_lobes(const DScene& sc, const rt_material& m, SurfHit& si, SurfTex& st, bool allow_multiple_lobes, rtgpu_lobe* scratch, int& n) {
  rtml::LobeList L{scratch, 0, rtml::kOk};
  DeviceMixChildren children{sc, si, st, true};
  const float eta = rtml::list_lobes<0>(m, allow_multiple_lobes, children, L);
  n = L.n;
  return eta;
}
RT_DEV void make_bsdf_textured(const DScene& sc, uint32_t row, SurfHit& si, SurfTex& st, const RayDiff& rd, bool allow_multiple_lobes, rtgpu_lobe* scratch,
                               Bsdf& bsdf) {
  compute_differential(si, st, rd);
  rt_material m;
  resolve_material(sc.texmats, sc.textures, sc.tex_data, (int)row, si, st, true, m);
  int n;
  const float eta = list_material_lobes(sc, m, si, st, allow_multiple_lobes, scratch, n);
  bsdf_init(bsdf, si, eta);
  bsdf.g = scratch; bsdf.n = n;
}

}  // namespace rt
