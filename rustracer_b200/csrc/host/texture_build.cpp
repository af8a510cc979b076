// Texture rows for the device (host, product code): parameter copy + MIP pyramids.
// Restates MIPMap::new (rustracer-core/src/mipmap.rs:65-180): Lanczos resampling of non-power-of-two images to the next
// powers of two, then the box-filtered pyramid through the wrap-mode texel lookup.  The filtering lookups themselves
// (trilinear / EWA) run on the device (csrc/device/texture.cuh).
#include <algorithm>
#include <cmath>
#include <limits>
#include <cstring>
#include <stdexcept>
#include "scene_build.hpp"

namespace rth {
namespace {

inline bool is_pow2(int32_t v) { return v != 0 && (v & (v - 1)) == 0; }                  // lib.rs:209-212
inline int32_t round_up_pow2(int32_t v) { v -= 1; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1; }   // lib.rs:214-224
inline int64_t modulo(int64_t a, int64_t b) { int64_t r = a % b; return r < 0 ? r + b : r; }   // mipmap.rs:455-462
inline int64_t wrap_index(int wrap, int64_t i, int64_t n) {
  if (wrap == RT_WRAP_REPEAT) return modulo(i, n);
  if (wrap == RT_WRAP_CLAMP) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  return i;
}
float lanczos(float f) {                                                                  // mipmap.rs:413-425
  const float tau = 2.0f, pi = 3.14159265358979323846f;
  float x = std::fabs(f);
  if (x < 1e-5f) return 1.0f;
  if (x > 1.0f) return 0.0f;
  x *= pi;
  const float s = std::sin(x * tau) / (x * tau);
  const float l = std::sin(x) / x;
  return s * l;
}
struct ResampleWeight { int32_t first_texel; float w[4]; };
std::vector<ResampleWeight> resample_weights(size_t old_res, size_t new_res) {           // mipmap.rs:383-411
  std::vector<ResampleWeight> wt(new_res);
  const float filter_width = 2.0f;
  for (size_t i = 0; i < new_res; i++) {
    const float center = ((float)i + 0.5f) * (float)old_res / (float)new_res;
    const float first = std::floor((center - filter_width) + 0.5f);
    for (int j = 0; j < 4; j++) { const float pos = first + (float)j + 0.5f; wt[i].w[j] = lanczos((pos - center) / filter_width); }
    const float inv = 1.0f / (wt[i].w[0] + wt[i].w[1] + wt[i].w[2] + wt[i].w[3]);
    for (int j = 0; j < 4; j++) wt[i].w[j] *= inv;
    wt[i].first_texel = (int32_t)first;
  }
  return wt;
}

}  // namespace

// MIPMap::new (mipmap.rs:65-180): level 0 is the image itself, or its Lanczos resampling to the next powers of two; every further
// level box-filters the previous one through the wrap-mode texel lookup.  `texels`: rx * ry * nc floats, row-major.
std::vector<MipLevel> build_mip_pyramid(int rx, int ry, int nc, int wrap, const float* texels) {
  std::vector<float> l0; int u0, v0;
  if (!is_pow2(rx) || !is_pow2(ry)) {                                 // mipmap.rs:73-139
    const int px = round_up_pow2(rx), py = round_up_pow2(ry);
    l0.assign((size_t)px * py * nc, 0.0f);
    const std::vector<ResampleWeight> sw = resample_weights((size_t)rx, (size_t)px);
    for (int tt = 0; tt < ry; tt++)
      for (int s = 0; s < px; s++)
        for (int j = 0; j < 4; j++) {
          const int64_t os = wrap_index(wrap, (int64_t)sw[s].first_texel + j, rx);
          if (os >= 0 && os < rx)
            for (int c = 0; c < nc; c++) l0[((size_t)tt * px + s) * nc + c] += texels[((size_t)tt * rx + (size_t)os) * nc + c] * sw[s].w[j];
        }
    const std::vector<ResampleWeight> tw = resample_weights((size_t)ry, (size_t)py);
    std::vector<float> work((size_t)py * nc);
    for (int s = 0; s < px; s++) {
      std::fill(work.begin(), work.end(), 0.0f);
      for (int tt = 0; tt < py; tt++)
        for (int j = 0; j < 4; j++) {
          const int64_t off = wrap_index(wrap, (int64_t)tw[tt].first_texel + j, ry);
          if (off >= 0 && off < ry)
            for (int c = 0; c < nc; c++) work[(size_t)tt * nc + c] += l0[((size_t)off * px + s) * nc + c] * tw[tt].w[j];
        }
      for (int tt = 0; tt < py; tt++)
        for (int c = 0; c < nc; c++) { const float v = work[(size_t)tt * nc + c]; l0[((size_t)tt * px + s) * nc + c] = clampf(v, 0.0f, std::numeric_limits<float>::infinity()); }
    }
    u0 = px; v0 = py;
  } else { l0.assign(texels, texels + (size_t)rx * ry * nc); u0 = rx; v0 = ry; }
  const int n_levels = 1 + (int)std::log2((float)std::max(u0, v0));   // mipmap.rs:150
  std::vector<MipLevel> out;
  std::vector<float> prev = std::move(l0); int pu = u0, pv = v0;
  for (int lv = 0; lv < n_levels; lv++) {
    if (lv > 0) {                                                     // mipmap.rs:157-175
      const int su = std::max(1, pu / 2), sv = std::max(1, pv / 2);
      std::vector<float> cur((size_t)su * sv * nc);
      auto texel = [&](int64_t s, int64_t tt, int c) -> float {       // MIPMap::texel (mipmap.rs:194-210)
        if (wrap == RT_WRAP_REPEAT) { s = modulo(s, pu); tt = modulo(tt, pv); }
        else if (wrap == RT_WRAP_CLAMP) { s = s < 0 ? 0 : (s > pu - 1 ? pu - 1 : s); tt = tt < 0 ? 0 : (tt > pv - 1 ? pv - 1 : tt); }
        else if (s < 0 || s >= pu || tt < 0 || tt >= pv) return 0.0f;
        return prev[((size_t)tt * pu + (size_t)s) * nc + c];
      };
      for (int tt = 0; tt < sv; tt++) for (int s = 0; s < su; s++) for (int c = 0; c < nc; c++)
        cur[((size_t)tt * su + s) * nc + c] = (((texel(2 * s, 2 * tt, c) + texel(2 * s + 1, 2 * tt, c)) + texel(2 * s, 2 * tt + 1, c)) + texel(2 * s + 1, 2 * tt + 1, c)) * 0.25f;
      prev = std::move(cur); pu = su; pv = sv;
    }
    MipLevel L; L.u = pu; L.v = pv; L.d = prev;
    out.push_back(std::move(L));
  }
  return out;
}

void build_textures(const rt_scene& in, std::vector<rtgpu_texture>& rows, std::vector<float>& pool) {
  rows.clear(); pool.clear();
  pool.resize(128);                                                   // EWA weight table (mipmap.rs:35-45)
  for (int i = 0; i < 128; i++) {
    const float alpha = 2.0f;
    const float r2 = (float)i / (128.0f - 1.0f);
    pool[i] = std::exp(-alpha * r2) - std::exp(-alpha);
  }
  std::vector<int> depth(in.n_textures, 1);                           // nesting of combinator textures (device evaluator: kMaxTexDepth = 4)
  for (uint32_t ti = 0; ti < in.n_textures; ti++) {
    const rt_texture& t = in.textures[ti];
    rtgpu_texture o; std::memset(&o, 0, sizeof(o));
    o.kind = t.kind; o.is_float = t.is_float;
    for (int k = 0; k < 3; k++) { o.value[k] = t.value[k]; o.vs[k] = t.vs[k]; o.vt[k] = t.vt[k]; }
    o.tex1 = t.tex1; o.tex2 = t.tex2; o.amount = t.amount;
    o.mapping = t.mapping; o.su = t.su; o.sv = t.sv; o.du = t.du; o.dv = t.dv; o.aa_none = t.aa_none;
    std::memcpy(o.w2t, t.w2t.m, sizeof(o.w2t));
    o.omega = t.omega; o.octaves = t.octaves; o.wrap = t.wrap; o.trilinear = t.trilinear; o.max_aniso = t.max_aniso;
    auto child_ok = [&](int32_t c) { return c >= 0 && (uint32_t)c < ti; };   // children are created before their parent: no cycles
    if ((t.kind == RT_TEX_SCALE || t.kind == RT_TEX_MIX || t.kind == RT_TEX_CHECKERBOARD) && (!child_ok(t.tex1) || !child_ok(t.tex2)))
      throw std::runtime_error("texture: child row out of range");
    if (t.kind == RT_TEX_MIX && !child_ok(t.amount)) throw std::runtime_error("texture: amount row out of range");
    if (t.kind == RT_TEX_SCALE || t.kind == RT_TEX_MIX || t.kind == RT_TEX_CHECKERBOARD) {
      depth[ti] = 1 + std::max(depth[(size_t)t.tex1], depth[(size_t)t.tex2]);
      if (t.kind == RT_TEX_MIX) depth[ti] = std::max(depth[ti], 1 + depth[(size_t)t.amount]);
      if (depth[ti] > 4) throw std::runtime_error("texture graph nested deeper than 4 levels (scale / mix / checkerboard of textures) is not supported on the device");
    }
    if (t.kind == RT_TEX_IMAGEMAP) {
      const int nc = t.is_float ? 1 : 3;
      if (t.img_w <= 0 || t.img_h <= 0 || !t.texels) throw std::runtime_error("imagemap texture without texels");
      o.channels = nc;
      std::vector<MipLevel> levels = build_mip_pyramid(t.img_w, t.img_h, nc, t.wrap, t.texels);
      if ((int)levels.size() > RTGPU_MAX_MIP_LEVELS) throw std::runtime_error("imagemap texture larger than 32768 texels on a side");
      o.n_levels = (int)levels.size();
      for (size_t lv = 0; lv < levels.size(); lv++) {
        if (pool.size() + levels[lv].d.size() > 0xffffffffull) throw std::runtime_error("texture pool exceeds 4 Gi floats");
        o.level_offset[lv] = (uint32_t)pool.size(); o.level_u[lv] = levels[lv].u; o.level_v[lv] = levels[lv].v;
        pool.insert(pool.end(), levels[lv].d.begin(), levels[lv].d.end());
      }
    }
    rows.push_back(o);
  }
}

}  // namespace rth
