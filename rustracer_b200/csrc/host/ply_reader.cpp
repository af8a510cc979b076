// PLY mesh reader for `Shape "plymesh"` (host, product code).
// Behaviour follows rustracer-core/src/shapes/plymesh.rs:18-178 (which delegates the container format to
// the `ply-rs` 0.1.3 crate): ascii / binary_little_endian / binary_big_endian; vertex properties must be
// float32 `x y z [nx ny nz] [u v | s t | texture_u texture_v | texture_s texture_t]`; the face property must
// be a list named `vertex_indices` with int32/uint32 items; quads are split (0,1,2)+(3,0,2); faces with any
// other vertex count are ignored; only `vertex` and `face` elements are allowed.
#include "pbrt_frontend.hpp"
#include <thread>
#include <atomic>
#include <cstdio>
#include <sstream>

namespace rth {
namespace {

enum Scalar { S_I8, S_U8, S_I16, S_U16, S_I32, S_U32, S_F32, S_F64, S_BAD };
Scalar scalar_of(const std::string& s) {
  if (s == "char" || s == "int8") return S_I8;
  if (s == "uchar" || s == "uint8") return S_U8;
  if (s == "short" || s == "int16") return S_I16;
  if (s == "ushort" || s == "uint16") return S_U16;
  if (s == "int" || s == "int32") return S_I32;
  if (s == "uint" || s == "uint32") return S_U32;
  if (s == "float" || s == "float32") return S_F32;
  if (s == "double" || s == "float64") return S_F64;
  return S_BAD;
}
size_t scalar_size(Scalar s) { static const size_t z[] = {1, 1, 2, 2, 4, 4, 4, 8, 0}; return z[s]; }

struct Prop { std::string name; bool is_list = false; Scalar count_type = S_BAD, type = S_BAD; };
struct Elem { std::string name; size_t count = 0; std::vector<Prop> props; };

struct Reader {
  const unsigned char* p; const unsigned char* end; int fmt;   // 0 ascii, 1 le, 2 be
  std::string ascii_token() {
    while (p < end && std::isspace(*p)) p++;
    const unsigned char* s = p;
    while (p < end && !std::isspace(*p)) p++;
    if (s == p) throw ParseError("PLY: unexpected end of data");
    return std::string((const char*)s, (const char*)p);
  }
  double scalar(Scalar t) {
    if (fmt == 0) return std::strtod(ascii_token().c_str(), nullptr);
    size_t n = scalar_size(t);
    if ((size_t)(end - p) < n) throw ParseError("PLY: unexpected end of data");
    unsigned char b[8];
    for (size_t i = 0; i < n; i++) b[i] = (fmt == 1) ? p[i] : p[n - 1 - i];
    p += n;
    switch (t) {
      case S_I8: return (double)(int8_t)b[0];
      case S_U8: return (double)b[0];
      case S_I16: { int16_t v; std::memcpy(&v, b, 2); return v; }
      case S_U16: { uint16_t v; std::memcpy(&v, b, 2); return v; }
      case S_I32: { int32_t v; std::memcpy(&v, b, 4); return v; }
      case S_U32: { uint32_t v; std::memcpy(&v, b, 4); return v; }
      case S_F32: { float v; std::memcpy(&v, b, 4); return v; }
      default: { double v; std::memcpy(&v, b, 8); return v; }
    }
  }
};

}  // namespace

// static split of [0, n) over the host threads
template <class F> static void parallel_chunks(size_t n, F body) {
  const size_t threads = std::max<size_t>(1, std::thread::hardware_concurrency());
  if (threads == 1 || n < ((size_t)1 << 16)) { body((size_t)0, n); return; }
  const size_t chunk = (n + threads - 1) / threads;
  std::vector<std::thread> pool;
  for (size_t b = chunk; b < n; b += chunk) pool.emplace_back([=, &body]() { body(b, std::min(n, b + chunk)); });
  body((size_t)0, std::min(n, chunk));
  for (auto& t : pool) t.join();
}

void read_ply(const std::string& filename, PlyMesh& out) {
  FILE* f = std::fopen(filename.c_str(), "rb");
  if (!f) throw ParseError("PLY file \"" + filename + "\": cannot open");   // plymesh.rs:26 `File::open(..).unwrap()`
  uvec<unsigned char> data;                                         // uninitialised: fread fills it (hmath.hpp)
  std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  data.resize((size_t)std::max(0L, sz));
  if (sz > 0 && std::fread(data.data(), 1, (size_t)sz, f) != (size_t)sz) { std::fclose(f); throw ParseError("PLY: short read"); }
  std::fclose(f);
  // header
  size_t pos = 0; int fmt = -1; std::vector<Elem> elems; bool ended = false;
  auto next_line = [&](std::string& line) {
    if (pos >= data.size()) return false;
    size_t e = pos;
    while (e < data.size() && data[e] != '\n') e++;
    line.assign((const char*)&data[pos], e - pos);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    pos = std::min(data.size(), e + 1);
    return true;
  };
  std::string line;
  if (!next_line(line) || line != "ply") throw ParseError("PLY file \"" + filename + "\": missing magic");
  while (next_line(line)) {
    std::istringstream ss(line); std::string kw; ss >> kw;
    if (kw == "format") { std::string fm; ss >> fm; fmt = fm == "ascii" ? 0 : (fm == "binary_little_endian" ? 1 : (fm == "binary_big_endian" ? 2 : -1)); }
    else if (kw == "comment" || kw == "obj_info" || kw.empty()) continue;
    else if (kw == "element") { Elem e; ss >> e.name >> e.count; elems.push_back(e); }
    else if (kw == "property") {
      if (elems.empty()) throw ParseError("PLY: property before element");
      Prop p; std::string t; ss >> t;
      if (t == "list") { std::string ct, it; ss >> ct >> it >> p.name; p.is_list = true; p.count_type = scalar_of(ct); p.type = scalar_of(it); }
      else { p.type = scalar_of(t); ss >> p.name; }
      if (p.type == S_BAD || (p.is_list && p.count_type == S_BAD)) throw ParseError("PLY: unknown property type in \"" + line + "\"");
      elems.back().props.push_back(p);
    } else if (kw == "end_header") { ended = true; break; }
    else throw ParseError("PLY: unknown header line \"" + line + "\"");
  }
  if (!ended || fmt < 0) throw ParseError("PLY file \"" + filename + "\": bad header");
  // plymesh.rs:39-87: header checks
  size_t vertex_count = 0, face_count = 0; bool has_normals = false, has_texture = false;
  for (const Elem& e : elems) {
    auto has = [&](const char* n) { for (const Prop& p : e.props) if (p.name == n) return true; return false; };
    if (e.name == "vertex") {
      vertex_count = e.count;
      if (!has("x") || !has("y") || !has("z")) return;               // "Vertex coordinate property not found" -> no shapes
      has_normals = has("nx") && has("ny") && has("nz");
      has_texture = (has("u") && has("v")) || (has("s") && has("t")) || (has("texture_u") && has("texture_v")) || (has("texture_s") && has("texture_t"));
    } else if (e.name == "face") face_count = e.count;
  }
  if (vertex_count == 0 || face_count == 0) return;                  // caller reports "invalid"
  Reader r{data.data() + pos, data.data() + data.size(), fmt};
  // every vertex / face occupies at least one byte of the body (ASCII: a digit; binary: far more): a header count beyond the
  // remaining file size is corrupt, and must not size an allocation
  for (const Elem& e : elems)
    if (e.count > data.size() - pos) throw ParseError("PLY file \"" + filename + "\": element count exceeds the file size");
  std::vector<int32_t> face;
  for (const Elem& e : elems) {
    if (e.name == "vertex") {
      out.P.assign(e.count * 3, 0.0f);
      if (has_normals) out.N.assign(e.count * 3, 0.0f);
      if (has_texture) out.uv.assign(e.count * 2, 0.0f);
      // where each property goes, decided once per element instead of once per vertex: 0-2 P, 3-5 N, 6-7 uv, -1 read and dropped
      std::vector<int> dest(e.props.size(), -1);
      for (size_t k = 0; k < e.props.size(); k++) {
        const Prop& p = e.props[k];
        if (p.is_list || p.type != S_F32) continue;                  // only ply::Property::Float is consumed (plymesh.rs:202-221)
        const std::string& n = p.name;
        if (n == "x") dest[k] = 0; else if (n == "y") dest[k] = 1; else if (n == "z") dest[k] = 2;
        else if (has_normals && n == "nx") dest[k] = 3; else if (has_normals && n == "ny") dest[k] = 4; else if (has_normals && n == "nz") dest[k] = 5;
        else if (has_texture && (n == "u" || n == "texture_u" || n == "s" || n == "texture_s")) dest[k] = 6;
        else if (has_texture && (n == "v" || n == "t" || n == "texture_v" || n == "texture_t")) dest[k] = 7;
      }
      bool plain_f32 = r.fmt == 1;                                   // fast path: little-endian records of float32 scalars only
      for (const Prop& p : e.props) plain_f32 = plain_f32 && !p.is_list && p.type == S_F32;
      if (plain_f32) {
        const size_t np = e.props.size(), stride = 4 * np;
        if ((size_t)(r.end - r.p) < stride * e.count) throw ParseError("PLY: unexpected end of data");
        for (size_t i = 0; i < e.count; i++, r.p += stride)
          for (size_t k = 0; k < np; k++) {
            const int d = dest[k];
            if (d < 0) continue;
            float fv; std::memcpy(&fv, r.p + 4 * k, 4);
            if (d < 3) out.P[3 * i + d] = fv; else if (d < 6) out.N[3 * i + d - 3] = fv; else out.uv[2 * i + d - 6] = fv;
          }
        continue;
      }
      for (size_t i = 0; i < e.count; i++) {
        for (size_t k = 0; k < e.props.size(); k++) {
          const Prop& p = e.props[k];
          if (p.is_list) { size_t n = (size_t)r.scalar(p.count_type); for (size_t j = 0; j < n; j++) r.scalar(p.type); continue; }
          const double v = r.scalar(p.type);
          const int d = dest[k];
          if (d < 0) continue;
          const float fv = (float)v;
          if (d < 3) out.P[3 * i + d] = fv; else if (d < 6) out.N[3 * i + d - 3] = fv; else out.uv[2 * i + d - 6] = fv;
        }
      }
    } else if (e.name == "face") {
      out.indices.reserve(e.count * 3);
      std::vector<char> takes(e.props.size(), 0);
      for (size_t k = 0; k < e.props.size(); k++)
        takes[k] = e.props[k].is_list && e.props[k].name == "vertex_indices" && (e.props[k].type == S_I32 || e.props[k].type == S_U32);   // ListInt / ListUInt only (plymesh.rs:233-240)
      if (r.fmt == 1 && e.props.size() == 1 && takes[0] && e.props[0].count_type == S_U8) {   // fast path: `list uchar int vertex_indices`, little-endian
        // all triangles (13-byte records): checked and copied by all host threads
        if ((size_t)(r.end - r.p) >= 13 * e.count && e.count >= (1u << 16)) {
          const unsigned char* base = r.p;
          std::atomic<int> other{0};
          parallel_chunks(e.count, [&](size_t i0, size_t i1) { for (size_t i = i0; i < i1; i++) if (base[13 * i] != 3) { other = 1; return; } });
          if (!other) {
            const size_t first = out.indices.size();
            out.indices.resize(first + 3 * e.count);
            int32_t* dst = out.indices.data() + first;
            parallel_chunks(e.count, [&](size_t i0, size_t i1) { for (size_t i = i0; i < i1; i++) std::memcpy(dst + 3 * i, base + 13 * i + 1, 12); });
            r.p += 13 * e.count;
            continue;
          }
        }
        for (size_t i = 0; i < e.count; i++) {
          if (r.p >= r.end) throw ParseError("PLY: unexpected end of data");
          const size_t n = *r.p++;
          if ((size_t)(r.end - r.p) < 4 * n) throw ParseError("PLY: unexpected end of data");
          if (n == 3 || n == 4) {
            int32_t v[4]; std::memcpy(v, r.p, 4 * n);
            out.indices.push_back(v[0]); out.indices.push_back(v[1]); out.indices.push_back(v[2]);
            if (n == 4) { out.indices.push_back(v[3]); out.indices.push_back(v[0]); out.indices.push_back(v[2]); }
          }
          r.p += 4 * n;
        }
        continue;
      }
      for (size_t i = 0; i < e.count; i++) {
        face.clear();
        for (size_t pk = 0; pk < e.props.size(); pk++) {
          const Prop& p = e.props[pk];
          if (!p.is_list) { r.scalar(p.type); continue; }
          size_t n = (size_t)r.scalar(p.count_type);
          const bool take = takes[pk] != 0;
          if (take) face.clear();
          for (size_t k = 0; k < n; k++) { double v = r.scalar(p.type); if (take) face.push_back((int32_t)(uint32_t)(int64_t)v); }
        }
        if (face.size() == 3 || face.size() == 4) {                  // plymesh.rs:107-127
          out.indices.push_back(face[0]); out.indices.push_back(face[1]); out.indices.push_back(face[2]);
          if (face.size() == 4) { out.indices.push_back(face[3]); out.indices.push_back(face[0]); out.indices.push_back(face[2]); }
        }
      }
    } else throw ParseError("PLY: Unexpected element \"" + e.name + "\"");   // plymesh.rs:101 panic
  }
}

}  // namespace rth
