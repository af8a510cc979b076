// C ABI of the host side (include/rthost.h).
#include "../../../include/rthost.h"
#include "pbrt_frontend.hpp"
#include "scene_build.hpp"
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

using namespace rth;

struct rth_scene {
  std::unique_ptr<ParsedScene> parsed;
  FlatScene flat;
  bool flattened = false;
};

static thread_local std::string g_error;

template <class F> static int guarded(F f) {
  try { f(); g_error.clear(); return 0; }
  catch (const std::exception& e) { g_error = e.what(); return -1; }
  catch (...) { g_error = "unknown error"; return -1; }
}

extern "C" {

const char* rth_last_error(void) { return g_error.c_str(); }

int rth_parse_file(const char* path, rth_scene** out) {
  *out = nullptr;
  return guarded([&] { auto s = new rth_scene(); try { s->parsed = parse_scene_file(path, FrontendOptions()); } catch (...) { delete s; throw; } *out = s; });
}
int rth_parse_string(const char* text, const char* search_dir, rth_scene** out) {
  *out = nullptr;
  return guarded([&] {
    FrontendOptions opt; if (search_dir) opt.search_dir = search_dir;
    auto s = new rth_scene(); try { s->parsed = parse_scene_text(text, opt); } catch (...) { delete s; throw; } *out = s;
  });
}
void rth_scene_free(rth_scene* s) { delete s; }
rt_scene* rth_scene_ir(rth_scene* s) { return &s->parsed->store.view; }
int rth_n_warnings(rth_scene* s) { return (int)s->parsed->store.warnings.size(); }
const char* rth_warning(rth_scene* s, int i) { return s->parsed->store.warnings.at((size_t)i).c_str(); }
const char* rth_film_filename(rth_scene* s) { return s->parsed->film_filename.c_str(); }
const char* rth_integrator_name(rth_scene* s) { return s->parsed->integrator_name.c_str(); }

int rth_flatten(rth_scene* s, int threads) {
  return guarded([&] { s->flat = FlatScene(); flatten_scene(s->parsed->store.view, threads, s->flat); s->flattened = true; });
}
int rth_flatten_with_builder(rth_scene* s, int threads, rth_bvh_builder builder, void* user) {
  return guarded([&] { s->flat = FlatScene(); flatten_scene(s->parsed->store.view, threads, s->flat, builder, user); s->flattened = true; });
}
const rtgpu_scene_desc* rth_scene_desc(rth_scene* s) { return s->flattened ? &s->flat.desc : nullptr; }
double rth_bvh_build_seconds(rth_scene* s) { return s->flat.bvh.build_seconds; }
uint64_t rth_n_triangles(rth_scene* s) { return s->flat.n_triangles; }
const uint32_t* rth_slot_of_prim(rth_scene* s) { return s->flattened ? s->flat.slot_of_prim.data() : nullptr; }
int rth_render_desc(rth_scene* s, rtgpu_render_desc* out) { return guarded([&] { make_render_desc(s->parsed->store.view, *out); }); }

int rth_tokenize(const char* text, char* out, size_t out_len) {
  int count = -1;
  int rc = guarded([&] {
    std::vector<Token> toks = tokenize(text);
    std::string s;
    for (const Token& t : toks) {
      if (t.kind == Tok::STR) s += "STR:" + t.str;
      else if (t.kind == Tok::NUMBER) { char b[64]; std::snprintf(b, sizeof b, "NUMBER:%.9g", (double)t.num); s += b; }
      else s += tok_name(t.kind);
      s += "\n";
    }
    if (out && out_len) std::snprintf(out, out_len, "%s", s.c_str());
    count = (int)toks.size();
  });
  return rc == 0 ? count : -1;
}
int rth_param_header(const char* s, int* type_out, char* name_out, size_t name_len) {
  ParamType t; std::string n;
  if (!parse_param_header(s, t, n)) return -1;
  *type_out = (int)t;
  std::snprintf(name_out, name_len, "%s", n.c_str());
  return 0;
}

// ---- image output -------------------------------------------------------------------------------
static uint32_t crc32_of(const unsigned char* d, size_t n, uint32_t crc = 0) {
  static uint32_t table[256]; static bool init = false;
  if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
  crc = ~crc;
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ d[i]) & 0xff] ^ (crc >> 8);
  return ~crc;
}
static void png_chunk(FILE* f, const char* type, const std::vector<unsigned char>& data) {
  unsigned char len[4] = {(unsigned char)(data.size() >> 24), (unsigned char)(data.size() >> 16), (unsigned char)(data.size() >> 8), (unsigned char)data.size()};
  std::fwrite(len, 1, 4, f);
  std::vector<unsigned char> buf(type, type + 4);
  buf.insert(buf.end(), data.begin(), data.end());
  std::fwrite(buf.data(), 1, buf.size(), f);
  uint32_t c = crc32_of(buf.data(), buf.size());
  unsigned char cb[4] = {(unsigned char)(c >> 24), (unsigned char)(c >> 16), (unsigned char)(c >> 8), (unsigned char)c};
  std::fwrite(cb, 1, 4, f);
}
static float gamma_correct(float v) { return v <= 0.0031308f ? 12.92f * v : 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f; }   // spectrum.rs:387-393

// SURVEY 8d C4 ray batches, the multi-threaded twin of rustracer_b200/scenes.py ray_batch (which defines them): ray i draws from the
// PCG32 stream (rng.rs:5-52) `seed * 2^32 + i` — origin uniform in the world bounds grown 5 %, then either a uniform direction
// (uniform_sample_sphere, sampling/mod.rs:14-20; t_max = inf) or the segment to a second uniform point (d = p1 - p0, t_max = 1 - 1e-4).
int rth_ray_batch(uint64_t n, const float* world_lo, const float* world_hi, uint64_t seed, int any_hit, uint64_t first, rtgpu_ray* out) {
  return guarded([&] {
    float c[3], h[3];
    for (int k = 0; k < 3; k++) { c[k] = (world_lo[k] + world_hi[k]) * 0.5f; h[k] = (world_hi[k] - world_lo[k]) * (float)(0.5 * 1.05); }
    auto body = [&](uint64_t a, uint64_t b) {
      for (uint64_t i = a; i < b; i++) {
        const uint64_t idx = first + i;
        uint64_t state = 0; const uint64_t inc = (((seed << 32) + idx) << 1) | 1u;        // RNG::set_sequence
        auto u32 = [&]() { const uint64_t o = state; state = o * 0x5851F42D4C957F2Dull + inc;
                           const uint32_t x = (uint32_t)(((o >> 18) ^ o) >> 27), r = (uint32_t)(o >> 59); return (x >> r) | (x << ((~r + 1u) & 31u)); };
        u32(); state += 0x853C49E6748FEA9Bull; u32();
        auto f32 = [&]() { const float v = (float)u32() * 2.3283064365386963e-10f; return v < 0.99999994f ? v : 0.99999994f; };
        rtgpu_ray r;
        float o[3];
        for (int k = 0; k < 3; k++) o[k] = c[k] + (f32() * 2.0f - 1.0f) * h[k];
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        if (any_hit) {
          float d[3];
          for (int k = 0; k < 3; k++) { const float p1 = c[k] + (f32() * 2.0f - 1.0f) * h[k]; d[k] = p1 - o[k]; }
          r.dx = d[0]; r.dy = d[1]; r.dz = d[2]; r.tmax = (float)(1.0 - 1e-4);
        } else {
          const float u0 = f32(), u1 = f32();
          const float z = 1.0f - 2.0f * u0, rr = std::sqrt(std::max(1.0f - z * z, 0.0f)), phi = (float)(2.0 * 3.14159265358979323846) * u1;
          r.dx = rr * std::cos(phi); r.dy = rr * std::sin(phi); r.dz = z; r.tmax = INFINITY;
        }
        r.tag = (uint32_t)(idx & 0xffffffffull);
        out[i] = r;
      }
    };
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    if (n < (1u << 16)) threads = 1;
    const uint64_t chunk = (n + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (uint64_t a = 0; a < n; a += chunk) pool.emplace_back(body, a, std::min(n, a + chunk));
    for (auto& t : pool) t.join();
  });
}

int rth_write_image(const char* path, const float* rgb, int width, int height) {
  return guarded([&] {
    std::string p(path);
    auto ends = [&](const char* e) { size_t L = std::strlen(e); return p.size() >= L && p.compare(p.size() - L, L, e) == 0; };
    FILE* f = std::fopen(path, "wb");
    if (!f) throw std::runtime_error("Failed to save image file " + p);
    if (ends(".pfm")) {
      std::fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
      for (int y = height - 1; y >= 0; y--) std::fwrite(rgb + (size_t)y * width * 3, sizeof(float), (size_t)width * 3, f);
    } else if (ends(".png")) {                                        // imageio.rs:52-73
      static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
      std::fwrite(sig, 1, 8, f);
      std::vector<unsigned char> ihdr = {(unsigned char)(width >> 24), (unsigned char)(width >> 16), (unsigned char)(width >> 8), (unsigned char)width,
                                         (unsigned char)(height >> 24), (unsigned char)(height >> 16), (unsigned char)(height >> 8), (unsigned char)height, 8, 2, 0, 0, 0};
      png_chunk(f, "IHDR", ihdr);
      std::vector<unsigned char> raw; raw.reserve((size_t)height * ((size_t)width * 3 + 1));
      for (int y = 0; y < height; y++) {
        raw.push_back(0);
        for (int x = 0; x < width * 3; x++) {
          float v = 255.0f * gamma_correct(rgb[(size_t)y * width * 3 + x]) + 0.5f;
          v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
          raw.push_back((v == v) ? (unsigned char)v : 0);
        }
      }
      std::vector<unsigned char> z = {0x78, 0x01};                    // zlib header, stored (uncompressed) deflate blocks
      uint32_t a = 1, b = 0;
      for (unsigned char c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
      size_t pos = 0;
      while (pos < raw.size() || raw.empty()) {
        size_t n = std::min<size_t>(65535, raw.size() - pos);
        bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0); z.push_back((unsigned char)n); z.push_back((unsigned char)(n >> 8)); z.push_back((unsigned char)~n); z.push_back((unsigned char)(~n >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
        if (last) break;
      }
      uint32_t ad = (b << 16) | a;
      z.push_back((unsigned char)(ad >> 24)); z.push_back((unsigned char)(ad >> 16)); z.push_back((unsigned char)(ad >> 8)); z.push_back((unsigned char)ad);
      png_chunk(f, "IDAT", z);
      png_chunk(f, "IEND", {});
    } else if (ends(".exr")) {                                        // imageio.rs:75-92: 32-bit float R, G, B scan lines (here ZIP blocks of 16 lines)
      std::vector<unsigned char> hd;
      auto put32 = [](std::vector<unsigned char>& v, uint32_t x) { for (int k = 0; k < 4; k++) v.push_back((unsigned char)(x >> (8 * k))); };
      auto put64 = [](std::vector<unsigned char>& v, uint64_t x) { for (int k = 0; k < 8; k++) v.push_back((unsigned char)(x >> (8 * k))); };
      auto putf = [&](std::vector<unsigned char>& v, float x) { uint32_t u; std::memcpy(&u, &x, 4); put32(v, u); };
      auto puts0 = [](std::vector<unsigned char>& v, const char* t) { while (*t) v.push_back((unsigned char)*t++); v.push_back(0); };
      auto attr = [&](const char* name, const char* type, const std::vector<unsigned char>& data) { puts0(hd, name); puts0(hd, type); put32(hd, (uint32_t)data.size()); hd.insert(hd.end(), data.begin(), data.end()); };
      put32(hd, 20000630u); put32(hd, 2u);
      std::vector<unsigned char> a;
      for (const char* c : {"B", "G", "R"}) { puts0(a, c); put32(a, 2u); put32(a, 0u); put32(a, 1u); put32(a, 1u); }   // FLOAT, pLinear 0, sampling 1 x 1; channels in alphabetical order
      a.push_back(0);
      attr("channels", "chlist", a);
      attr("compression", "compression", {3});
      a.clear(); put32(a, 0u); put32(a, 0u); put32(a, (uint32_t)(width - 1)); put32(a, (uint32_t)(height - 1));
      attr("dataWindow", "box2i", a); attr("displayWindow", "box2i", a);
      attr("lineOrder", "lineOrder", {0});
      a.clear(); putf(a, 1.0f); attr("pixelAspectRatio", "float", a);
      a.clear(); putf(a, 0.0f); putf(a, 0.0f); attr("screenWindowCenter", "v2f", a);
      a.clear(); putf(a, 1.0f); attr("screenWindowWidth", "float", a);
      hd.push_back(0);
      const int lines_per_block = 16;
      const size_t n_blocks = ((size_t)std::max(height, 0) + lines_per_block - 1) / lines_per_block, line_bytes = (size_t)width * 12;
      std::vector<std::vector<unsigned char>> blocks(n_blocks);
      std::vector<unsigned char> raw, tmp;
      for (size_t b = 0; b < n_blocks; b++) {
        const int y0 = (int)b * lines_per_block, lines = std::min(lines_per_block, height - y0);
        const size_t n = line_bytes * (size_t)lines;
        raw.resize(n); tmp.resize(n);
        for (int l = 0; l < lines; l++)
          for (int c = 0; c < 3; c++)                                   // B, G, R planes of the line
            for (int x = 0; x < width; x++) std::memcpy(&raw[(size_t)l * line_bytes + ((size_t)c * width + x) * 4], &rgb[((size_t)(y0 + l) * width + x) * 3 + (2 - c)], 4);
        size_t e = 0, o = (n + 1) / 2;                                  // even bytes first, odd bytes second, then the byte-difference predictor
        for (size_t i = 0; i < n; i++) { if (i & 1) tmp[o++] = raw[i]; else tmp[e++] = raw[i]; }
        unsigned char prev = n ? tmp[0] : 0;
        for (size_t i = 1; i < n; i++) { const unsigned char cur = tmp[i]; tmp[i] = (unsigned char)(cur - prev + 128); prev = cur; }
        uLongf len = compressBound((uLong)n);
        std::vector<unsigned char>& out = blocks[b];
        out.resize(len);
        if (compress2(out.data(), &len, tmp.data(), (uLong)n, 6) != Z_OK || len >= n) out = raw;   // a block that does not shrink is stored as it is
        else out.resize(len);
      }
      uint64_t off = hd.size() + 8 * n_blocks;
      for (size_t b = 0; b < n_blocks; b++) { put64(hd, off); off += 8 + blocks[b].size(); }
      std::fwrite(hd.data(), 1, hd.size(), f);
      for (size_t b = 0; b < n_blocks; b++) {
        std::vector<unsigned char> bh; put32(bh, (uint32_t)(b * lines_per_block)); put32(bh, (uint32_t)blocks[b].size());
        std::fwrite(bh.data(), 1, 8, f);
        std::fwrite(blocks[b].data(), 1, blocks[b].size(), f);
      }
    } else { std::fclose(f); throw std::runtime_error("Unsupported file format"); }   // imageio.rs:47-49
    std::fclose(f);
  });
}

}  // extern "C"
