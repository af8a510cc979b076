// TEST INFRASTRUCTURE — CPU oracle, not product code (see orc_api.cpp).  Textures, MIP maps, noise and bump mapping
// (SURVEY 8f rank 3).  Restates rustracer-core/src/{texture/*.rs, mipmap.rs, noise.rs, material/mod.rs:50-92}; line numbers
// are relative to rustracer-core/src/.  Parity unpinned: the reference has no test or fixture for any of these (its only
// tests in mipmap.rs check ndarray indexing); the self-tests live in tests/test_oracle_textures.py.
#pragma once
#include "orc_shapes.hpp"
#include "../include/rt_scene.h"
#include <memory>

namespace orc {

// ---------------------------------------------------------------------------------------
// noise.rs.  The permutation table is Ken Perlin's reference permutation (data), doubled (:96-119).
static const uint8_t kNoisePerm[256] = {
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
inline int noise_perm(int i) { return kNoisePerm[i & 255]; }   // NOISE_PERM[i] for i < 512 (second half repeats the first)
inline float noise_grad(int x, int y, int z, float dx, float dy, float dz) {   // :67-75
  int h = noise_perm(noise_perm(noise_perm(x) + y) + z);
  h &= 15;
  float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
  float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
  return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
inline float noise_weight(float t) { float t3 = t * t * t, t4 = t3 * t; return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3; }   // :77-82
inline float lerpf(float t, float a, float b) { return a * (1.0f - t) + b * t; }   // lib.rs:107-117
inline float noise(float x, float y, float z) {                 // :7-41
  int ix = f2i32(std::floor(x)), iy = f2i32(std::floor(y)), iz = f2i32(std::floor(z));
  float dx = x - (float)ix, dy = y - (float)iy, dz = z - (float)iz;
  ix &= 255; iy &= 255; iz &= 255;
  float w000 = noise_grad(ix, iy, iz, dx, dy, dz), w100 = noise_grad(ix + 1, iy, iz, dx - 1.0f, dy, dz);
  float w010 = noise_grad(ix, iy + 1, iz, dx, dy - 1.0f, dz), w110 = noise_grad(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
  float w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz - 1.0f), w101 = noise_grad(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
  float w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f), w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
  float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
  float x00 = lerpf(wx, w000, w100), x10 = lerpf(wx, w010, w110), x01 = lerpf(wx, w001, w101), x11 = lerpf(wx, w011, w111);
  float y0 = lerpf(wy, x00, x10), y1 = lerpf(wy, x01, x11);
  return lerpf(wz, y0, y1);
}
inline float smooth_step(float mn, float mx, float value) { float v = clampv((value - mn) / (mx - mn), 0.0f, 1.0f); return v * v * (-2.0f * v + 3.0f); }   // :84-89
inline float fbm(V3 p, V3 dpdx, V3 dpdy, float omega, uint32_t max_octaves) {   // :43-61
  float len2 = std::fmax(length_squared(dpdx), length_squared(dpdy));
  float n = clampv(-1.0f - 0.5f * std::log2(len2), 0.0f, (float)max_octaves);
  uint32_t n_int = f2u32(std::floor(n));
  float sum = 0.0f, lambda = 1.0f, o = 1.0f;
  for (uint32_t i = 0; i < n_int; i++) {
    V3 q = lambda * p;
    sum += o * noise(q.x, q.y, q.z);
    lambda *= 1.99f; o *= omega;
  }
  float n_partial = n - (float)n_int;
  V3 q = lambda * p;
  sum += o * smooth_step(0.3f, 0.7f, n_partial) * noise(q.x, q.y, q.z);
  return sum;
}

// ---------------------------------------------------------------------------------------
// mipmap.rs.  T = f32 or Spectrum: `nc` channels per texel, every operation channel-wise.
struct Texel { float c[3] = {0, 0, 0}; };
struct MIPMap {
  int nc = 1; bool do_trilinear = false; float max_anisotropy = 8.0f; int wrap = RT_WRAP_REPEAT;
  int res_x = 0, res_y = 0;
  struct Level { int u = 0, v = 0; std::vector<float> d; };
  std::vector<Level> pyramid;
  float weight_lut[128];

  static int64_t modulo(int64_t a, int64_t b) { int64_t r = a % b; return r < 0 ? r + b : r; }   // :455-462
  static float lanczos(float f) {                                 // :413-425
    const float tau = 2.0f;
    float x = std::fabs(f);
    if (x < 1e-5f) return 1.0f;
    if (x > 1.0f) return 0.0f;
    x *= PI;
    float s = std::sin(x * tau) / (x * tau);
    float l = std::sin(x) / x;
    return s * l;
  }
  struct ResampleWeight { int32_t first_texel; float w[4]; };
  static std::vector<ResampleWeight> resample_weights(size_t old_res, size_t new_res) {   // :383-411
    std::vector<ResampleWeight> wt;
    const float filter_width = 2.0f;
    for (size_t i = 0; i < new_res; i++) {
      float center = ((float)i + 0.5f) * (float)old_res / (float)new_res;
      float first = std::floor((center - filter_width) + 0.5f);
      ResampleWeight r;
      for (int j = 0; j < 4; j++) { float pos = first + (float)j + 0.5f; r.w[j] = lanczos((pos - center) / filter_width); }
      float inv = 1.0f / (r.w[0] + r.w[1] + r.w[2] + r.w[3]);
      for (int j = 0; j < 4; j++) r.w[j] *= inv;
      r.first_texel = f2i32(first);
      wt.push_back(r);
    }
    return wt;
  }
  int64_t wrap_index(int64_t i, int64_t n) const { return wrap == RT_WRAP_REPEAT ? modulo(i, n) : (wrap == RT_WRAP_CLAMP ? clampv<int64_t>(i, 0, n - 1) : i); }

  // MIPMap::new (:65-180)
  void build(int rx, int ry, const float* img, int channels, bool trilerp, float max_aniso, int wrap_mode) {
    nc = channels; do_trilinear = trilerp; max_anisotropy = max_aniso; wrap = wrap_mode;
    for (int i = 0; i < 128; i++) {                               // :35-45
      const float alpha = 2.0f;
      float r2 = (float)i / (128.0f - 1.0f);
      weight_lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
    }
    Level l0;
    if (!is_power_of_2(rx) || !is_power_of_2(ry)) {
      const int px = round_up_pow_2(rx), py = round_up_pow_2(ry);
      std::vector<float> out((size_t)px * py * nc, 0.0f);
      std::vector<ResampleWeight> sw = resample_weights((size_t)rx, (size_t)px);
      for (int t = 0; t < ry; t++)                                // :87-106 (rows t >= res.y stay zero)
        for (int s = 0; s < px; s++)
          for (int j = 0; j < 4; j++) {
            int64_t orig_s = wrap_index((int64_t)sw[s].first_texel + j, rx);
            if (orig_s >= 0 && orig_s < rx)
              for (int c = 0; c < nc; c++) out[((size_t)t * px + s) * nc + c] += img[((size_t)t * rx + (size_t)orig_s) * nc + c] * sw[s].w[j];
          }
      std::vector<ResampleWeight> tw = resample_weights((size_t)ry, (size_t)py);
      std::vector<float> work((size_t)py * nc);
      for (int s = 0; s < px; s++) {                              // :108-131
        std::fill(work.begin(), work.end(), 0.0f);
        for (int t = 0; t < py; t++)
          for (int j = 0; j < 4; j++) {
            int64_t off = wrap_index((int64_t)tw[t].first_texel + j, ry);
            if (off >= 0 && off < ry)
              for (int c = 0; c < nc; c++) work[(size_t)t * nc + c] += out[((size_t)off * px + s) * nc + c] * tw[t].w[j];
          }
        for (int t = 0; t < py; t++)
          for (int c = 0; c < nc; c++) out[((size_t)t * px + s) * nc + c] = clampv(work[(size_t)t * nc + c], 0.0f, INF);
      }
      l0.u = px; l0.v = py; l0.d = std::move(out);
    } else {
      l0.u = rx; l0.v = ry; l0.d.assign(img, img + (size_t)rx * ry * nc);
    }
    res_x = l0.u; res_y = l0.v;
    pyramid.clear();
    pyramid.push_back(std::move(l0));
    const size_t n_levels = 1 + (size_t)f2usize(std::log2((float)std::max(res_x, res_y)));   // :150
    for (size_t i = 1; i < n_levels; i++) {                        // :157-175
      Level l; l.u = std::max(1, pyramid[i - 1].u / 2); l.v = std::max(1, pyramid[i - 1].v / 2);
      l.d.resize((size_t)l.u * l.v * nc);
      for (int t = 0; t < l.v; t++) for (int s = 0; s < l.u; s++) {
        Texel a = texel(i - 1, 2 * s, 2 * t), b = texel(i - 1, 2 * s + 1, 2 * t), c = texel(i - 1, 2 * s, 2 * t + 1), d = texel(i - 1, 2 * s + 1, 2 * t + 1);
        for (int k = 0; k < nc; k++) l.d[((size_t)t * l.u + s) * nc + k] = (((a.c[k] + b.c[k]) + c.c[k]) + d.c[k]) * 0.25f;
      }
      pyramid.push_back(std::move(l));
    }
  }
  size_t levels() const { return pyramid.size(); }
  Texel texel(size_t level, int64_t s, int64_t t) const {          // :194-210
    const Level& l = pyramid[level];
    Texel r;
    if (wrap == RT_WRAP_REPEAT) { s = modulo(s, l.u); t = modulo(t, l.v); }
    else if (wrap == RT_WRAP_CLAMP) { s = clampv<int64_t>(s, 0, l.u - 1); t = clampv<int64_t>(t, 0, l.v - 1); }
    else if (s < 0 || s >= l.u || t < 0 || t >= l.v) return r;   // black
    for (int k = 0; k < nc; k++) r.c[k] = l.d[((size_t)t * l.u + (size_t)s) * nc + k];
    return r;
  }
  static Texel tlerp(float t, Texel a, Texel b) { Texel r; for (int k = 0; k < 3; k++) r.c[k] = a.c[k] * (1.0f - t) + b.c[k] * t; return r; }
  Texel triangle(size_t level, P2 st) const {                      // :270-294
    level = clampv<size_t>(level, 0, levels() - 1);
    float s = st.x * (float)pyramid[level].u - 0.5f, t = st.y * (float)pyramid[level].v - 0.5f;
    int64_t s0 = (int64_t)f2i64(std::floor(s)), t0 = (int64_t)f2i64(std::floor(t));
    float ds = s - (float)s0, dt = t - (float)t0;
    Texel a = texel(level, s0, t0), b = texel(level, s0, t0 + 1), c = texel(level, s0 + 1, t0), d = texel(level, s0 + 1, t0 + 1), r;
    for (int k = 0; k < 3; k++)
      r.c[k] = ((a.c[k] * (1.0f - ds) * (1.0f - dt) + b.c[k] * (1.0f - ds) * dt) + c.c[k] * ds * (1.0f - dt)) + d.c[k] * ds * dt;
    return r;
  }
  Texel lookup(P2 st, float width) const {                         // :212-229
    float level = (float)levels() - 1.0f + std::log2(std::fmax(width, 1e-8f));
    if (level < 0.0f) return triangle(0, st);
    if (level >= (float)levels() - 1.0f) return texel(levels() - 1, 0, 0);
    float i_level = std::floor(level);
    float delta = level - i_level;
    return tlerp(delta, triangle((size_t)f2usize(i_level), st), triangle((size_t)f2usize(i_level) + 1, st));
  }
  Texel ewa(size_t level, P2 st, P2 dst0, P2 dst1) const {         // :296-381
    if (level >= levels()) return texel(levels() - 1, 0, 0);
    const float us = (float)pyramid[level].u, vs = (float)pyramid[level].v;
    st.x = st.x * us - 0.5f; st.y = st.y * vs - 0.5f;
    dst0.x *= us; dst0.y *= vs; dst1.x *= us; dst1.y *= vs;
    float A = dst0.y * dst0.y + dst1.y * dst1.y + 1.0f;
    float B = -2.0f * (dst0.x * dst0.y + dst1.x * dst1.y);
    float C = dst0.x * dst0.x + dst1.x * dst1.x + 1.0f;
    float inv_f = 1.0f / (A * C - B * B * 0.25f);
    A *= inv_f; B *= inv_f; C *= inv_f;
    float det = -B * B + 4.0f * A * C;
    float inv_det = 1.0f / det;
    float u_sqrt = std::sqrt(det * C), v_sqrt = std::sqrt(A * det);
    int64_t s0 = f2i64(std::ceil(st.x - 2.0f * inv_det * u_sqrt)), s1 = f2i64(std::floor(st.x + 2.0f * inv_det * u_sqrt));
    int64_t t0 = f2i64(std::ceil(st.y - 2.0f * inv_det * v_sqrt)), t1 = f2i64(std::floor(st.y + 2.0f * inv_det * v_sqrt));
    Texel sum; float sum_wts = 0.0f;
    for (int64_t it = t0; it < t1 + 1; it++) {
      float tt = (float)it - st.y;
      for (int64_t is = s0; is < s1 + 1; is++) {
        float ss = (float)is - st.x;
        float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
        if (r2 < 1.0f) {
          int64_t index = pmin<int64_t>(f2usize(r2 * 128.0f), 127);
          float weight = weight_lut[index];
          Texel tx = texel(level, is, it);
          for (int k = 0; k < 3; k++) sum.c[k] += tx.c[k] * weight;
          sum_wts += weight;
        }
      }
    }
    for (int k = 0; k < 3; k++) sum.c[k] = sum.c[k] / sum_wts;
    return sum;
  }
  Texel lookup_diff(P2 st, P2 dst0, P2 dst1) const {               // :231-268
    if (do_trilinear) {
      float width = std::fmax(std::fmax(std::fabs(dst0.x), std::fabs(dst0.y)), std::fmax(std::fabs(dst1.x), std::fabs(dst1.y)));
      return lookup(st, 2.0f * width);
    }
    auto len2 = [](P2 v) { return v.x * v.x + v.y * v.y; };
    if (len2(dst0) < len2(dst1)) std::swap(dst0, dst1);
    float major_length = std::sqrt(len2(dst0));
    float minor_length = std::sqrt(len2(dst1));
    if ((minor_length * max_anisotropy) < major_length && minor_length > 0.0f) {
      float scale = major_length / (minor_length * max_anisotropy);
      dst1.x *= scale; dst1.y *= scale;
      minor_length *= scale;
    }
    if (minor_length == 0.0f) return triangle(0, st);
    float lod = std::fmax(0.0f, (float)levels() - 1.0f + std::log2(minor_length));
    size_t ilod = (size_t)f2usize(std::floor(lod));
    return tlerp(lod - (float)ilod, ewa(ilod, st, dst0, dst1), ewa(ilod + 1, st, dst0, dst1));
  }
  static int64_t f2i64(float f) { if (!(f == f)) return 0; if (f <= -9.2e18f) return INT64_MIN; if (f >= 9.2e18f) return INT64_MAX; return (int64_t)f; }   // `as isize` saturates
};

// ---------------------------------------------------------------------------------------
// texture/*.rs: the scene's texture rows (include/rt_scene.h rt_texture) with their MIP maps built
struct TextureSet {
  std::vector<rt_texture> rows;
  std::vector<std::unique_ptr<MIPMap>> mips;                      // per row, imagemap only
  void init(const rt_texture* t, size_t n) {
    rows.assign(t, t + n); mips.clear(); mips.resize(n);
    for (size_t i = 0; i < n; i++)
      if (rows[i].kind == RT_TEX_IMAGEMAP) {
        mips[i].reset(new MIPMap());
        mips[i]->build(rows[i].img_w, rows[i].img_h, rows[i].texels, rows[i].is_float ? 1 : 3, rows[i].trilinear != 0, rows[i].max_aniso, rows[i].wrap);
      }
  }
  // TextureMapping2D::map (texture/mod.rs:32-83)
  static void map2d(const rt_texture& t, const SurfaceInteraction& si, P2& st, P2& dstdx, P2& dstdy) {
    if (t.mapping == RT_TEXMAP_PLANAR) {
      V3 vs(t.vs[0], t.vs[1], t.vs[2]), vt(t.vt[0], t.vt[1], t.vt[2]);
      V3 vec = si.hit.p;
      st = P2(t.du + dot(vec, vs), t.dv + dot(vec, vt));
      dstdx = P2(dot(si.dpdx, vs), dot(si.dpdx, vt));
      dstdy = P2(dot(si.dpdy, vs), dot(si.dpdy, vt));
    } else {
      st = P2(t.su * si.uv.x + t.du, t.sv * si.uv.y + t.dv);
      dstdx = P2(t.su * si.dudx, t.sv * si.dvdx);
      dstdy = P2(t.su * si.dudy, t.sv * si.dvdy);
    }
  }
  // Texture<T>::evaluate.  A float texture returns its value in all three channels.
  Spectrum eval(int row, const SurfaceInteraction& si) const {
    const rt_texture& t = rows[(size_t)row];
    switch (t.kind) {
      case RT_TEX_CONSTANT: return Spectrum(t.value[0], t.value[1], t.value[2]);          // constant.rs:36-38
      case RT_TEX_SCALE: return eval(t.tex1, si) * eval(t.tex2, si);                      // scale.rs:24-26
      case RT_TEX_MIX: {                                                                   // mix.rs:24-30
        Spectrum t1 = eval(t.tex1, si), t2 = eval(t.tex2, si);
        float amt = eval(t.amount, si).r;
        return t1 * (1.0f - amt) + t2 * amt;
      }
      case RT_TEX_UV: {                                                                    // uv.rs:51-54
        P2 st, dx, dy; map2d(t, si, st, dx, dy);
        return Spectrum(st.x - std::floor(st.x), st.y - std::floor(st.y), 0.0f);
      }
      case RT_TEX_CHECKERBOARD: {                                                          // checkerboard.rs:106-143
        P2 st, dstdx, dstdy; map2d(t, si, st, dstdx, dstdy);
        if (t.aa_none) {
          uint32_t k = f2u32(std::floor(st.x)) + f2u32(std::floor(st.y));                 // `as u32` saturates, the sum wraps in release builds
          return (k % 2 == 0) ? eval(t.tex1, si) : eval(t.tex2, si);
        }
        float ds = std::fmax(std::fabs(dstdx.x), std::fabs(dstdy.x));
        float dt = std::fmax(std::fabs(dstdx.y), std::fabs(dstdy.y));
        float s0 = st.x - ds, s1 = st.x + ds, t0 = st.y - dt, t1 = st.y + dt;
        if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1)) {
          int32_t k = (int32_t)((uint32_t)f2i32(std::floor(st.x)) + (uint32_t)f2i32(std::floor(st.y)));
          return (k % 2 == 0) ? eval(t.tex1, si) : eval(t.tex2, si);
        }
        auto bump_int = [](float x) { return std::floor(x / 2.0f) + 2.0f * std::fmax(x / 2.0f - std::floor(x / 2.0f) - 0.5f, 0.0f); };
        float sint = (bump_int(s1) - bump_int(s0)) / (2.0f * ds);
        float tint = (bump_int(t1) - bump_int(t0)) / (2.0f * dt);
        float area2 = sint + tint - 2.0f * sint * tint;
        if (ds > 1.0f || dt > 1.0f) area2 = 0.5f;
        return eval(t.tex1, si) * (1.0f - area2) + eval(t.tex2, si) * area2;
      }
      case RT_TEX_IMAGEMAP: {                                                              // imagemap.rs:231-234
        P2 st, dstdx, dstdy; map2d(t, si, st, dstdx, dstdy);
        Texel x = mips[(size_t)row]->lookup_diff(st, dstdx, dstdy);
        return t.is_float ? Spectrum(x.c[0]) : Spectrum(x.c[0], x.c[1], x.c[2]);
      }
      case RT_TEX_FBM: {                                                                   // fbm.rs:19-22 ; IdentityMapping3D (mod.rs:102-110)
        Transform w2t(Matrix4::from(t.w2t.m), Matrix4::from(t.w2t.m_inv));
        V3 dpdx = w2t.vector(si.dpdx), dpdy = w2t.vector(si.dpdy), p = w2t.point(si.hit.p);
        return Spectrum(fbm(p, dpdx, dpdy, t.omega, (uint32_t)t.octaves));
      }
      default: return Spectrum(0.0f);
    }
  }
  // material/mod.rs:50-92 (dndu = dndv = 0, see orc_shapes.hpp)
  void bump(int row, SurfaceInteraction& si) const {
    const V3 kZeroN(0, 0, 0);
    SurfaceInteraction si_eval = si;
    float du = 0.5f * (std::fabs(si.dudx) + std::fabs(si.dudy));
    if (du == 0.0f) du = 0.0005f;
    si_eval.hit.p = si.hit.p + du * si.shading.dpdu;
    si_eval.uv = P2(si.uv.x + du, si.uv.y + 0.0f);
    si_eval.hit.n = normalize(cross(si.shading.dpdu, si.shading.dpdv) + du * kZeroN);
    float u_displace = eval(row, si_eval).r;
    float dv = 0.5f * (std::fabs(si.dvdx) + std::fabs(si.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    si_eval.hit.p = si.hit.p + dv * si.shading.dpdv;
    si_eval.uv = P2(si.uv.x + 0.0f, si.uv.y + dv);
    si_eval.hit.n = normalize(cross(si.shading.dpdu, si.shading.dpdv) + dv * kZeroN);
    float v_displace = eval(row, si_eval).r;
    float displace = eval(row, si).r;
    V3 dpdu = si.shading.dpdu + (u_displace - displace) / du * si.shading.n + displace * kZeroN;
    V3 dpdv = si.shading.dpdv + (v_displace - displace) / dv * si.shading.n + displace * kZeroN;
    si.set_shading_geometry(dpdu, dpdv, false);
  }
  // The material row with every textured parameter evaluated at `si` (the `.evaluate(si)` calls of material/*.rs); applies
  // the bump map first when the material has one.
  rt_material resolve(const rt_material& in, SurfaceInteraction& si) const {
    rt_material m = in;
    if (m.type != RT_MAT_MIX && m.tex[RT_TS_BUMP]) bump(m.tex[RT_TS_BUMP] - 1, si);
    auto S = [&](int slot, float* d) { if (m.tex[slot]) { Spectrum v = eval(m.tex[slot] - 1, si); d[0] = v.r; d[1] = v.g; d[2] = v.b; } };
    auto F = [&](int slot, float& d) { if (m.tex[slot]) d = eval(m.tex[slot] - 1, si).r; };
    S(RT_TS_KD, m.kd); S(RT_TS_KS, m.ks); S(RT_TS_KR, m.kr); S(RT_TS_KT, m.kt); S(RT_TS_ETA_RGB, m.eta_rgb); S(RT_TS_K_RGB, m.k_rgb);
    S(RT_TS_OPACITY, m.opacity); S(RT_TS_REFLECT, m.reflect); S(RT_TS_TRANSMIT, m.transmit); S(RT_TS_AMOUNT, m.amount);
    F(RT_TS_SIGMA, m.sigma); F(RT_TS_ROUGHNESS, m.roughness); F(RT_TS_UROUGHNESS, m.uroughness); F(RT_TS_VROUGHNESS, m.vroughness); F(RT_TS_ETA, m.eta);
    return m;
  }
};

}  // namespace orc
