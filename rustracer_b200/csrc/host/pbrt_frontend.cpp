// See pbrt_frontend.hpp.  Reference line numbers are relative to /root/reference/rustracer-core/src/.
#include "pbrt_frontend.hpp"
#include <zlib.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <unistd.h>
#include <climits>

namespace rth {

// tests/golden/copper_rgb.json (hex floats are exact)
const float kCopperEtaRgb[3] = {0x1.9994b8p-3f, 0x1.d81b7ap-1f, 0x1.199178p+0f};
const float kCopperKRgb[3] = {0x1.f3cb18p+1f, 0x1.394c0cp+1f, 0x1.119e9ap+1f};

// ------------------------------------------------------------------------------------------------
// Lexer
static const struct { const char* word; Tok tok; } kKeywords[] = {   // pbrt/lexer.rs:201-242
    {"Accelerator", Tok::ACCELERATOR}, {"ActiveTransform", Tok::ACTIVETRANSFORM}, {"All", Tok::ALL}, {"AreaLightSource", Tok::AREALIGHTSOURCE},
    {"AttributeBegin", Tok::ATTRIBUTEBEGIN}, {"AttributeEnd", Tok::ATTRIBUTEEND}, {"Camera", Tok::CAMERA}, {"ConcatTransform", Tok::CONCATTRANSFORM},
    {"CoordinateSystem", Tok::COORDINATESYSTEM}, {"CoordSysTransform", Tok::COORDSYSTRANSFORM}, {"EndTime", Tok::ENDTIME}, {"Film", Tok::FILM},
    {"Identity", Tok::IDENTITY}, {"Include", Tok::INCLUDE}, {"LightSource", Tok::LIGHTSOURCE}, {"LookAt", Tok::LOOKAT},
    {"MakeNamedMedium", Tok::MAKENAMEDMEDIUM}, {"MakeNamedMaterial", Tok::MAKENAMEDMATERIAL}, {"Material", Tok::MATERIAL},
    {"MediumInterface", Tok::MEDIUMINTERFACE}, {"NamedMaterial", Tok::NAMEDMATERIAL}, {"ObjectBegin", Tok::OBJECTBEGIN}, {"ObjectEnd", Tok::OBJECTEND},
    {"ObjectInstance", Tok::OBJECTINSTANCE}, {"PixelFilter", Tok::PIXELFILTER}, {"ReverseOrientation", Tok::REVERSEORIENTATION}, {"Rotate", Tok::ROTATE},
    {"Sampler", Tok::SAMPLER}, {"Scale", Tok::SCALE}, {"Shape", Tok::SHAPE}, {"StartTime", Tok::STARTTIME}, {"Integrator", Tok::INTEGRATOR},
    {"Texture", Tok::TEXTURE}, {"TransformBegin", Tok::TRANSFORMBEGIN}, {"TransformEnd", Tok::TRANSFORMEND}, {"TransformTimes", Tok::TRANSFORMTIMES},
    {"Transform", Tok::TRANSFORM}, {"Translate", Tok::TRANSLATE}, {"WorldBegin", Tok::WORLDBEGIN}, {"WorldEnd", Tok::WORLDEND}};

const char* tok_name(Tok t) {
  for (const auto& k : kKeywords) if (k.tok == t) return k.word;
  switch (t) { case Tok::STR: return "STR"; case Tok::NUMBER: return "NUMBER"; case Tok::LBRACK: return "["; case Tok::RBRACK: return "]"; default: return "COMMENT"; }
}

// nom `float` (number/complete): [+-]? (digits [. digits?]? | . digits) ([eE][+-]?digits)? | nan | inf | infinity
static size_t recognize_float(const std::string& s, size_t p) {
  size_t i = p, n = s.size();
  if (i < n && (s[i] == '+' || s[i] == '-')) i++;
  auto ci = [&](const char* w) {
    size_t k = 0; while (w[k]) { if (i + k >= n || std::tolower((unsigned char)s[i + k]) != w[k]) return (size_t)0; k++; } return k;
  };
  if (size_t k = ci("infinity")) return i + k;
  if (size_t k = ci("inf")) return i + k;
  if (size_t k = ci("nan")) return i + k;
  size_t d0 = i;
  while (i < n && std::isdigit((unsigned char)s[i])) i++;
  size_t int_digits = i - d0, frac_digits = 0;
  if (i < n && s[i] == '.') {
    size_t j = i + 1;
    while (j < n && std::isdigit((unsigned char)s[j])) j++;
    frac_digits = j - (i + 1);
    if (int_digits > 0 || frac_digits > 0) i = j;
  }
  if (int_digits == 0 && frac_digits == 0) return p;
  if (i < n && (s[i] == 'e' || s[i] == 'E')) {
    size_t j = i + 1;
    if (j < n && (s[j] == '+' || s[j] == '-')) j++;
    size_t e0 = j;
    while (j < n && std::isdigit((unsigned char)s[j])) j++;
    if (j > e0) i = j;
  }
  return i;
}

std::vector<Token> tokenize(const std::string& in) {                 // pbrt/lexer.rs:185-263
  std::vector<Token> out;
  size_t i = 0, n = in.size();
  auto skip_ws = [&]() { while (i < n && (in[i] == ' ' || in[i] == '\t' || in[i] == '\r' || in[i] == '\n')) i++; };
  while (true) {
    skip_ws();
    if (i >= n) break;
    // 1. keyword: alphanumeric1 | "[" | "]"
    size_t j = i;
    while (j < n && std::isalnum((unsigned char)in[j])) j++;
    bool matched = false;
    if (j > i) {
      std::string w = in.substr(i, j - i);
      for (const auto& k : kKeywords) if (w == k.word) { out.push_back(Token{k.tok, "", 0}); i = j; matched = true; break; }
    } else if (in[i] == '[') { out.push_back(Token{Tok::LBRACK, "", 0}); i++; matched = true; }
    else if (in[i] == ']') { out.push_back(Token{Tok::RBRACK, "", 0}); i++; matched = true; }
    if (matched) continue;
    // 2. float
    size_t e = recognize_float(in, i);
    if (e > i) { Token t{Tok::NUMBER, "", 0}; t.num = std::strtof(in.substr(i, e - i).c_str(), nullptr); out.push_back(t); i = e; continue; }
    // 3. string: '"' none_of('"')* '"'
    if (in[i] == '"') {
      size_t q = in.find('"', i + 1);
      if (q == std::string::npos) throw ParseError("Failed to tokenize scene file: unterminated string");
      out.push_back(Token{Tok::STR, in.substr(i + 1, q - i - 1), 0}); i = q + 1; continue;
    }
    // 4. comment: '#' not_line_ending line_ending   (a comment on the last line without newline fails — App. B)
    if (in[i] == '#') {
      size_t q = i + 1;
      while (q < n && in[q] != '\n' && in[q] != '\r') q++;
      if (q >= n) throw ParseError("Failed to tokenize scene file: comment not terminated by a newline");
      if (in[q] == '\r') { if (q + 1 < n && in[q + 1] == '\n') q += 2; else throw ParseError("Failed to tokenize scene file: bare CR"); }
      else q += 1;
      out.push_back(Token{Tok::COMMENT, "", 0}); i = q; continue;
    }
    throw ParseError("Failed to tokenize scene file: unexpected input near '" + in.substr(i, 24) + "'");
  }
  if (out.empty()) throw ParseError("Failed to tokenize scene file: no tokens");   // many1
  return out;
}

// ------------------------------------------------------------------------------------------------
// ParamSet
bool parse_param_header(const std::string& s, ParamType& type, std::string& name) {   // pbrt/parser.rs:198-231
  static const struct { const char* tag; ParamType t; } kTypes[] = {
      {"integer", ParamType::Int}, {"bool", ParamType::Bool}, {"float", ParamType::Float}, {"point2", ParamType::Point2}, {"vector2", ParamType::Vector2},
      {"point3", ParamType::Point3}, {"vector3", ParamType::Vector3}, {"point", ParamType::Point3}, {"vector", ParamType::Vector3},
      {"normal", ParamType::Normal}, {"color", ParamType::Rgb}, {"rgb", ParamType::Rgb}, {"xyz", ParamType::Xyz}, {"blackbody", ParamType::Blackbody},
      {"spectrum", ParamType::Spectrum}, {"string", ParamType::String}, {"texture", ParamType::Texture}};
  for (const auto& k : kTypes) {
    size_t L = std::strlen(k.tag);
    if (s.compare(0, L, k.tag) == 0) {
      // first matching tag wins (nom alt); it must be followed by space1
      size_t p = L;
      while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) p++;
      if (p == L) return false;
      type = k.t; name = s.substr(p);
      return true;
    }
  }
  return false;
}

void ParamSet::add(ParamType t, const std::string& name, const std::vector<float>& nums, const std::vector<std::string>& strs) {   // paramset.rs:60-165
  auto triples = [&](std::vector<ParamItem<Vec3>>& dst) {
    ParamItem<Vec3> it; it.name = name;
    for (size_t i = 0; i + 3 <= nums.size(); i += 3) it.values.push_back(v3(nums[i], nums[i + 1], nums[i + 2]));
    dst.push_back(it);
  };
  switch (t) {
    case ParamType::Bool: { ParamItem<bool> it; it.name = name; for (auto& s : strs) it.values.push_back(s == "true"); bools.push_back(it); break; }
    case ParamType::Int: { ParamItem<int32_t> it; it.name = name; for (float f : nums) it.values.push_back((int32_t)(std::isnan(f) ? 0 : (f >= 2147483648.0f ? INT32_MAX : (f <= -2147483648.0f ? INT32_MIN : (int32_t)f)))); ints.push_back(it); break; }
    case ParamType::Float: { ParamItem<float> it; it.name = name; it.values = nums; floats.push_back(it); break; }
    case ParamType::String: { ParamItem<std::string> it; it.name = name; it.values = strs; strings.push_back(it); break; }
    case ParamType::Texture: { ParamItem<std::string> it; it.name = name; it.values = strs; textures.push_back(it); break; }
    case ParamType::Rgb: { ParamItem<Rgb> it; it.name = name; for (size_t i = 0; i + 3 <= nums.size(); i += 3) it.values.push_back(Rgb{nums[i], nums[i + 1], nums[i + 2]}); spectra.push_back(it); break; }
    case ParamType::Point2: { ParamItem<std::pair<float, float>> it; it.name = name; for (size_t i = 0; i + 2 <= nums.size(); i += 2) it.values.push_back({nums[i], nums[i + 1]}); point2s.push_back(it); break; }
    case ParamType::Point3: triples(point3s); break;
    case ParamType::Vector3: triples(vector3s); break;
    case ParamType::Normal: triples(normal3s); break;
    case ParamType::Vector2: case ParamType::Xyz: notes.push_back("Parameter type of \"" + name + "\" is not implemented yet!"); break;   // paramset.rs:153-160
    case ParamType::Spectrum: case ParamType::Blackbody:
      // paramset.rs:142-151,254-297: SPD files / blackbody need the CIE tables (reference data, not copied) — outside this path's scope
      notes.push_back("unsupported:" + name); break;
  }
}

// ------------------------------------------------------------------------------------------------
// File helpers (fileutil.rs)
static std::string dir_of(const std::string& path) {
  char buf[PATH_MAX];
  std::string p = path;
  if (realpath(path.c_str(), buf)) p = buf;
  size_t s = p.find_last_of('/');
  return s == std::string::npos ? std::string(".") : p.substr(0, s);
}
static std::string resolve_filename(const std::string& fn, const std::string& search_dir) {   // fileutil.rs:35-50
  if (search_dir.empty() || fn.empty() || fn[0] == '/') return fn;
  std::string joined = search_dir + "/" + fn;
  char buf[PATH_MAX];
  if (realpath(joined.c_str(), buf)) return std::string(buf);
  return fn;
}
static std::string read_file(const std::string& fn) {
  std::ifstream f(fn, std::ios::binary);
  if (!f) throw ParseError("Failed to open scene file: " + fn);
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}
static std::vector<Token> tokenize_file(const std::string& fn, const std::string& search_dir) {   // pbrt/mod.rs:27-44
  std::vector<Token> toks = tokenize(read_file(resolve_filename(fn, search_dir)));
  std::vector<Token> out;
  for (auto& t : toks) if (t.kind != Tok::COMMENT) out.push_back(std::move(t));
  return out;
}

// Minimal PFM reader for environment maps ("PF", w h, scale<0 = little endian; rows bottom-to-top).
static bool read_pfm(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  char magic[3] = {0, 0, 0}; float scale = 0;
  if (std::fscanf(f, "%2s %d %d %f", magic, &w, &h, &scale) != 4 || (magic[1] != 'F' && magic[1] != 'f') || w <= 0 || h <= 0 || w > 65536 || h > 65536) { std::fclose(f); return false; }
  std::fgetc(f);
  int nc = magic[1] == 'F' ? 3 : 1;
  std::vector<float> raw((size_t)w * h * nc);
  size_t got = std::fread(raw.data(), sizeof(float), raw.size(), f);
  std::fclose(f);
  if (got != raw.size()) return false;
  if (scale > 0) for (float& v : raw) { uint32_t u; std::memcpy(&u, &v, 4); u = __builtin_bswap32(u); std::memcpy(&v, &u, 4); }
  if (std::fabs(scale) != 1.0f) for (float& v : raw) v *= std::fabs(scale);   // imageio.rs:231-233
  rgb.resize((size_t)w * h * 3);
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) for (int c = 0; c < 3; c++)
    rgb[((size_t)y * w + x) * 3 + c] = raw[((size_t)(h - 1 - y) * w + x) * nc + (nc == 3 ? c : 0)];
  return true;
}

// 8-bit PNG reader (imageio.rs:94-113: `image::open(..).to_rgb8()`, channel / 255): colour types 0, 2, 3, 4, 6 at 8 or 16 bits,
// non-interlaced; zlib inflates the IDAT stream.  Rows top-to-bottom.
static bool read_png(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> file;
  unsigned char buf[65536]; size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
  std::fclose(f);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return false;
  auto be32 = [&](size_t o) { return ((uint32_t)file[o] << 24) | ((uint32_t)file[o + 1] << 16) | ((uint32_t)file[o + 2] << 8) | (uint32_t)file[o + 3]; };
  int depth = 0, ctype = 0, interlace = 0; bool have_ihdr = false;
  w = h = 0;
  std::vector<unsigned char> idat, plte;
  for (size_t o = 8; o + 12 <= file.size();) {
    const uint32_t len = be32(o);
    if (o + 12 + len > file.size()) return false;
    const std::string type((const char*)&file[o + 4], 4);
    const unsigned char* d = &file[o + 8];
    if (type == "IHDR") {
      if (len != 13 || o != 8) return false;                          // IHDR is 13 bytes and comes first
      const uint32_t uw = be32(o + 8), uh = be32(o + 12);
      if (uw == 0 || uh == 0 || uw > 65536u || uh > 65536u) return false;   // keeps every size product below 2^35
      w = (int)uw; h = (int)uh; depth = d[8]; ctype = d[9]; interlace = d[12]; have_ihdr = true;
    }
    else if (!have_ihdr) return false;
    else if (type == "PLTE") plte.assign(d, d + len);
    else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
    else if (type == "IEND") break;
    o += 12 + len;
  }
  if (!have_ihdr || w <= 0 || h <= 0 || interlace != 0 || (depth != 8 && depth != 16) || (ctype == 3 && depth != 8)) return false;
  const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!channels) return false;
  const size_t bpp = (size_t)channels * depth / 8, stride = (size_t)w * bpp;
  std::vector<unsigned char> raw((stride + 1) * (size_t)h);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return false;
  std::vector<unsigned char> img(stride * (size_t)h);
  for (int y = 0; y < h; y++) {                                      // undo the per-row filters
    const unsigned char ft = raw[(stride + 1) * y];
    const unsigned char* in = &raw[(stride + 1) * y + 1];
    unsigned char* cur = &img[stride * y];
    const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
      int pred = 0;
      if (ft == 1) pred = a; else if (ft == 2) pred = b; else if (ft == 3) pred = (a + b) / 2;
      else if (ft == 4) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
      else if (ft != 0) return false;
      cur[i] = (unsigned char)(in[i] + pred);
    }
  }
  rgb.resize((size_t)w * h * 3);
  const size_t step = depth / 8;                                     // 16-bit samples: the high byte (to_rgb8 scales 65535 -> 255)
  for (size_t i = 0; i < (size_t)w * h; i++) {
    const unsigned char* px = &img[i * bpp];
    unsigned char r, g, b;
    if (ctype == 0 || ctype == 4) r = g = b = px[0];
    else if (ctype == 3) { const size_t k = (size_t)px[0] * 3; if (k + 2 >= plte.size()) return false; r = plte[k]; g = plte[k + 1]; b = plte[k + 2]; }
    else { r = px[0]; g = px[step]; b = px[2 * step]; }
    rgb[i * 3] = (float)r / 255.0f; rgb[i * 3 + 1] = (float)g / 255.0f; rgb[i * 3 + 2] = (float)b / 255.0f;
  }
  return true;
}
// TGA reader (imageio.rs:94-113 handles .tga through the same `image::open(..).to_rgb8()`): true-colour (types 2 / 10, 24 or 32 bits)
// and greyscale (types 3 / 11, 8 bits), raw or run-length encoded, either row order.  Rows top-to-bottom on return.
static bool read_tga(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> d;
  unsigned char buf[65536]; size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  if (d.size() < 18) return false;
  const int id_len = d[0], cmap_type = d[1], type = d[2], bpp = d[16], desc = d[17];
  w = d[12] | (d[13] << 8); h = d[14] | (d[15] << 8);
  const bool grey = type == 3 || type == 11, rle = type == 10 || type == 11;
  if (cmap_type != 0 || w <= 0 || h <= 0 || !(type == 2 || type == 3 || type == 10 || type == 11)) return false;
  if ((grey && bpp != 8) || (!grey && bpp != 24 && bpp != 32)) return false;
  const size_t px = (size_t)bpp / 8, total = (size_t)w * h;
  std::vector<unsigned char> img(total * px);
  size_t pos = 18 + (size_t)id_len, out = 0;
  if (!rle) {
    if (pos + total * px > d.size()) return false;
    std::memcpy(img.data(), &d[pos], total * px);
  } else {
    while (out < total) {
      if (pos >= d.size()) return false;
      const int hdr = d[pos++], cnt = (hdr & 127) + 1;
      if (out + (size_t)cnt > total) return false;
      if (hdr & 128) {
        if (pos + px > d.size()) return false;
        for (int k = 0; k < cnt; k++) std::memcpy(&img[(out + k) * px], &d[pos], px);
        pos += px;
      } else {
        if (pos + (size_t)cnt * px > d.size()) return false;
        std::memcpy(&img[out * px], &d[pos], (size_t)cnt * px);
        pos += (size_t)cnt * px;
      }
      out += (size_t)cnt;
    }
  }
  const bool top_down = (desc & 0x20) != 0, right_left = (desc & 0x10) != 0;
  rgb.resize(total * 3);
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    const unsigned char* p = &img[((size_t)(top_down ? y : h - 1 - y) * w + (size_t)(right_left ? w - 1 - x : x)) * px];
    float* o = &rgb[((size_t)y * w + x) * 3];
    if (grey) o[0] = o[1] = o[2] = (float)p[0] / 255.0f;
    else { o[0] = (float)p[2] / 255.0f; o[1] = (float)p[1] / 255.0f; o[2] = (float)p[0] / 255.0f; }   // stored B, G, R
  }
  return true;
}
// Radiance .hdr reader (imageio.rs:115-132: `HdrDecoder::with_strictness(reader, false)` of the image crate, 0.24): "#?RADIANCE" /
// "#?RGBE" signature, header lines up to an empty line, "-Y h +X w", then per scanline either new-style run-length encoding
// (2 2 hi lo, four component planes) or flat RGBE quads (old-style runs with a 1 1 1 marker included).  A pixel (r, g, b, e) decodes to
// c * 2^(e - 136), black when e == 0.  Rows top-to-bottom.
static bool read_hdr(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> d;
  unsigned char buf[65536]; size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  size_t pos = 0;
  auto line = [&](std::string& out) { out.clear(); while (pos < d.size() && d[pos] != '\n') out.push_back((char)d[pos++]); if (pos >= d.size()) return false; pos++; return true; };
  std::string ln;
  if (!line(ln) || (ln.compare(0, 10, "#?RADIANCE") != 0 && ln.compare(0, 6, "#?RGBE") != 0)) return false;
  while (true) { if (!line(ln)) return false; if (ln.empty()) break; }
  if (!line(ln)) return false;
  char sy = 0, sx = 0, ay = 0, ax = 0; int hh = 0, ww = 0;
  if (std::sscanf(ln.c_str(), "%c%c %d %c%c %d", &sy, &ay, &hh, &sx, &ax, &ww) != 6 || sy != '-' || ay != 'Y' || sx != '+' || ax != 'X') return false;
  if (ww <= 0 || hh <= 0 || ww > 65536 || hh > 65536) return false;
  w = ww; h = hh;
  std::vector<unsigned char> row((size_t)w * 4);
  rgb.assign((size_t)w * h * 3, 0.0f);
  for (int y = 0; y < h; y++) {
    if (pos + 4 > d.size()) return false;
    if (w >= 8 && w < 32768 && d[pos] == 2 && d[pos + 1] == 2 && (((int)d[pos + 2] << 8) | d[pos + 3]) == w) {
      pos += 4;
      for (int c = 0; c < 4; c++) {
        int x = 0;
        while (x < w) {
          if (pos >= d.size()) return false;
          int cnt = d[pos++];
          if (cnt > 128) {                                            // run
            cnt -= 128;
            if (cnt == 0 || x + cnt > w || pos >= d.size()) return false;
            const unsigned char v = d[pos++];
            for (int k = 0; k < cnt; k++) row[(size_t)(x++) * 4 + c] = v;
          } else {                                                    // literal
            if (cnt == 0 || x + cnt > w || pos + (size_t)cnt > d.size()) return false;
            for (int k = 0; k < cnt; k++) row[(size_t)(x++) * 4 + c] = d[pos++];
          }
        }
      }
    } else {                                                          // flat, with old-style run markers
      int x = 0, shift = 0;
      while (x < w) {
        if (pos + 4 > d.size()) return false;
        const unsigned char* q = &d[pos]; pos += 4;
        if (q[0] == 1 && q[1] == 1 && q[2] == 1 && x > 0) {
          const int cnt = (int)q[3] << shift;
          if (x + cnt > w) return false;
          for (int k = 0; k < cnt; k++) { std::memcpy(&row[(size_t)x * 4], &row[(size_t)(x - 1) * 4], 4); x++; }
          shift += 8;
        } else { std::memcpy(&row[(size_t)x * 4], q, 4); x++; shift = 0; }
      }
    }
    for (int x = 0; x < w; x++) {
      const unsigned char* q = &row[(size_t)x * 4];
      float* o = &rgb[((size_t)y * w + x) * 3];
      if (q[3] == 0) { o[0] = o[1] = o[2] = 0.0f; continue; }
      const float e = std::exp2((float)q[3] - (128.0f + 8.0f));
      o[0] = e * (float)q[0]; o[1] = e * (float)q[1]; o[2] = e * (float)q[2];
    }
  }
  return true;
}
// OpenEXR reader (imageio.rs:134-160: `exr::prelude::read_first_rgba_layer_from_file`): single-part scan-line files with R, G, B channels
// (A and any other channel ignored) of type HALF, FLOAT or UINT, compression NONE / RLE / ZIPS / ZIP, data window == display window,
// increasing or decreasing line order.  Tiled, multi-part and deep files and the PIZ / PXR24 / B44 / DWA codecs are not read (the
// caller then takes the reference's "not found" path).  Rows top-to-bottom.
static float half_to_float(uint16_t hbits) {
  const uint32_t sign = (uint32_t)(hbits >> 15) << 31, exp = (hbits >> 10) & 31u, man = hbits & 1023u;
  uint32_t out;
  if (exp == 0) {
    if (man == 0) out = sign;
    else { int e = -1; uint32_t m = man; do { e++; m <<= 1; } while (!(m & 1024u)); out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 1023u) << 13); }
  } else if (exp == 31) out = sign | 0x7f800000u | (man << 13);
  else out = sign | ((exp + 127 - 15) << 23) | (man << 13);
  float f; std::memcpy(&f, &out, 4); return f;
}
static bool read_exr(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> d;
  unsigned char buf[65536]; size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  auto u32 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24); };
  auto u64 = [&](size_t o) { return (uint64_t)u32(o) | ((uint64_t)u32(o + 4) << 32); };
  if (d.size() < 8 || u32(0) != 20000630u) return false;
  const uint32_t version = u32(4);
  if ((version & 0xffu) != 2 || (version & 0x1a00u)) return false;     // tiled, deep or multi-part
  size_t pos = 8;
  struct Chan { std::string name; int type; };
  std::vector<Chan> chans;
  int compression = -1, line_order = 0; int32_t dw[4] = {0, 0, -1, -1}, disp[4] = {0, 0, -1, -1};
  while (true) {
    std::string name, type;
    while (pos < d.size() && d[pos]) name.push_back((char)d[pos++]);
    if (pos >= d.size()) return false;
    pos++;
    if (name.empty()) break;
    while (pos < d.size() && d[pos]) type.push_back((char)d[pos++]);
    pos++;
    if (pos + 4 > d.size()) return false;
    const uint32_t size = u32(pos); pos += 4;
    if (pos + size > d.size()) return false;
    if (name == "channels" && type == "chlist") {
      size_t q = pos;
      while (q < pos + size && d[q]) {
        Chan c;
        while (q < pos + size && d[q]) c.name.push_back((char)d[q++]);
        q++;
        if (q + 16 > pos + size) return false;
        c.type = (int)u32(q);
        if (u32(q + 8) != 1 || u32(q + 12) != 1) return false;          // sub-sampled channels
        q += 16;
        chans.push_back(c);
      }
    } else if (name == "compression" && size == 1) compression = d[pos];
    else if (name == "lineOrder" && size == 1) line_order = d[pos];
    else if (name == "dataWindow" && size == 16) for (int k = 0; k < 4; k++) dw[k] = (int32_t)u32(pos + 4 * k);
    else if (name == "displayWindow" && size == 16) for (int k = 0; k < 4; k++) disp[k] = (int32_t)u32(pos + 4 * k);
    pos += size;
  }
  if (compression < 0 || compression > 3 || line_order > 1) return false;
  for (int k = 0; k < 4; k++) if (dw[k] != disp[k]) return false;
  const int64_t W = (int64_t)dw[2] - dw[0] + 1, H = (int64_t)dw[3] - dw[1] + 1;
  if (W <= 0 || H <= 0 || W > 65536 || H > 65536) return false;
  int ci[3] = {-1, -1, -1};
  std::vector<size_t> chan_off(chans.size()); size_t line_bytes = 0;
  for (size_t k = 0; k < chans.size(); k++) {
    if (chans[k].type < 0 || chans[k].type > 2) return false;
    chan_off[k] = line_bytes; line_bytes += (size_t)W * (chans[k].type == 1 ? 2 : 4);
    if (chans[k].name == "R") ci[0] = (int)k; else if (chans[k].name == "G") ci[1] = (int)k; else if (chans[k].name == "B") ci[2] = (int)k;
  }
  if (ci[0] < 0 || ci[1] < 0 || ci[2] < 0) return false;
  const int lines_per_block = compression == 3 ? 16 : 1;
  const size_t n_blocks = (size_t)((H + lines_per_block - 1) / lines_per_block);
  if (pos + 8 * n_blocks > d.size()) return false;
  w = (int)W; h = (int)H;
  rgb.assign((size_t)W * H * 3, 0.0f);
  std::vector<unsigned char> raw, tmp;
  for (size_t b = 0; b < n_blocks; b++) {
    const uint64_t off = u64(pos + 8 * b);
    if (off + 8 > d.size()) return false;
    const int32_t y0 = (int32_t)u32(off); const uint32_t csize = u32(off + 4);
    if (off + 8 + csize > d.size() || y0 < dw[1] || y0 > dw[3]) return false;
    const int lines = (int)std::min<int64_t>(lines_per_block, (int64_t)dw[3] - y0 + 1);
    const size_t usize = line_bytes * (size_t)lines;
    const unsigned char* src = &d[off + 8];
    raw.resize(usize);
    if (compression == 0 || csize >= usize) { if (csize != usize) return false; std::memcpy(raw.data(), src, usize); }
    else {
      tmp.resize(usize);
      if (compression == 1) {                                         // RLE: n < 0 -> -n literal bytes; n >= 0 -> the next byte n + 1 times
        size_t i = 0, o = 0;
        while (i < csize) {
          const int c = (signed char)src[i++];
          if (c < 0) { const size_t m = (size_t)(-c); if (i + m > csize || o + m > usize) return false; std::memcpy(&tmp[o], &src[i], m); i += m; o += m; }
          else { const size_t m = (size_t)c + 1; if (i >= csize || o + m > usize) return false; std::memset(&tmp[o], src[i++], m); o += m; }
        }
        if (o != usize) return false;
      } else {
        uLongf len = (uLongf)usize;
        if (uncompress(tmp.data(), &len, src, (uLong)csize) != Z_OK || len != usize) return false;
      }
      for (size_t i = 1; i < usize; i++) tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128);   // predictor
      const size_t half = (usize + 1) / 2;                            // de-interleave: first half = even bytes, second half = odd bytes
      for (size_t i = 0; i < usize; i++) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
    }
    for (int l = 0; l < lines; l++) {
      const int64_t y = (int64_t)y0 - dw[1] + l;
      for (int c = 0; c < 3; c++) {
        const unsigned char* q = &raw[(size_t)l * line_bytes + chan_off[(size_t)ci[c]]];
        const int type = chans[(size_t)ci[c]].type;
        for (int64_t x = 0; x < W; x++) {
          float v;
          if (type == 1) { uint16_t hb; std::memcpy(&hb, q + 2 * x, 2); v = half_to_float(hb); }
          else if (type == 2) std::memcpy(&v, q + 4 * x, 4);
          else { uint32_t u; std::memcpy(&u, q + 4 * x, 4); v = (float)u; }
          rgb[((size_t)y * W + (size_t)x) * 3 + c] = v;
        }
      }
    }
  }
  return true;
}
// imageio.rs:77-92 `read_image`: by extension.  PFM, PNG, TGA, Radiance HDR and scan-line OpenEXR (NONE / RLE / ZIPS / ZIP) are read.
static bool read_image_rgb(const std::string& fn, int& w, int& h, std::vector<float>& rgb) {
  auto ends = [&](const char* e) { size_t n = std::strlen(e); return fn.size() >= n && fn.compare(fn.size() - n, n, e) == 0; };
  if (ends(".pfm")) return read_pfm(fn, w, h, rgb);
  if (ends(".png")) return read_png(fn, w, h, rgb);
  if (ends(".tga")) return read_tga(fn, w, h, rgb);
  if (ends(".hdr")) return read_hdr(fn, w, h, rgb);
  if (ends(".exr")) return read_exr(fn, w, h, rgb);
  return false;
}

// ------------------------------------------------------------------------------------------------
// API state machine (api.rs)
namespace {

// A material / texture parameter after lookup: a constant, or a row of the scene's texture table.
struct TexRef { Rgb c{0, 0, 0}; float f = 0; int row = -1; };
struct TextureParams {                                               // paramset.rs:349-467
  const ParamSet& geom; const ParamSet& mat;
  const std::map<std::string, int>& ftex; const std::map<std::string, int>& stex;   // name -> row of `table`
  std::vector<rt_texture>* table;
  std::vector<std::string>* warn;
  float find_float(const std::string& n, float d) const { return geom.find_one_float(n, mat.find_one_float(n, d)); }
  int32_t find_int(const std::string& n, int32_t d) const { return geom.find_one_int(n, mat.find_one_int(n, d)); }
  bool find_bool(const std::string& n, bool d) const { return geom.find_one_bool(n, mat.find_one_bool(n, d)); }
  std::string find_string(const std::string& n, const std::string& d) const { return geom.find_one_string(n, mat.find_one_string(n, d)); }
  Rgb find_spectrum(const std::string& n, Rgb d) const { return geom.find_one_spectrum(n, mat.find_one_spectrum(n, d)); }
  Vec3 find_vector3(const std::string& n, Vec3 d) const { return geom.find_one_vector3(n, mat.find_one_vector3(n, d)); }
  std::string tex_name(const std::string& n) const { std::string nm = geom.find_texture(n); if (nm.empty()) nm = mat.find_texture(n); return nm; }
  // a named texture: folded to its value when it is a ConstantTexture, else referenced by row
  TexRef ref_of(int row) const {
    const rt_texture& t = (*table)[(size_t)row];
    TexRef r;
    if (t.kind == RT_TEX_CONSTANT) { r.c = Rgb{t.value[0], t.value[1], t.value[2]}; r.f = t.value[0]; }
    else r.row = row;
    return r;
  }
  TexRef spectrum_texture(const std::string& n, Rgb def) const {     // :406-424
    std::string nm = tex_name(n);
    if (!nm.empty()) {
      auto it = stex.find(nm);
      if (it != stex.end()) return ref_of(it->second);
      warn->push_back("Couldn't find spectrum texture " + nm + " for parameter " + n);
    }
    TexRef r; r.c = geom.find_one_spectrum(n, mat.find_one_spectrum(n, def));
    return r;
  }
  TexRef float_texture(const std::string& n, float def) const {      // :426-442
    std::string nm = tex_name(n);
    if (!nm.empty()) {
      auto it = ftex.find(nm);
      if (it != ftex.end()) return ref_of(it->second);
      warn->push_back("Couldn't find float texture " + nm + " for parameter " + n);
    }
    TexRef r; r.f = geom.find_one_float(n, mat.find_one_float(n, def));
    return r;
  }
  bool float_texture_or_none(const std::string& n, TexRef& out) const {   // :444-466
    std::string nm = tex_name(n);
    if (!nm.empty()) {
      auto it = ftex.find(nm);
      if (it != ftex.end()) { out = ref_of(it->second); return true; }
      warn->push_back("Couldn't find float texture " + nm + " for parameter " + n);
      return false;
    }
    if (const auto* v = geom.find_float(n)) { out = TexRef(); out.f = v->at(0); return true; }
    if (const auto* v = mat.find_float(n)) { out = TexRef(); out.f = v->at(0); return true; }
    return false;
  }
  // children of scale / mix / checkerboard textures are always rows: a constant gets its own ConstantTexture row
  int child_row(const TexRef& r, bool is_float) const {
    if (r.row >= 0) return r.row;
    rt_texture t; std::memset(&t, 0, sizeof(t));
    t.kind = RT_TEX_CONSTANT; t.is_float = is_float; t.tex1 = t.tex2 = t.amount = -1;
    if (is_float) t.value[0] = t.value[1] = t.value[2] = r.f; else { t.value[0] = r.c.r; t.value[1] = r.c.g; t.value[2] = r.c.b; }
    table->push_back(t);
    return (int)table->size() - 1;
  }
};

struct GraphicsState {                                               // api.rs:313-371
  std::map<std::string, int> float_textures;                         // name -> row of SceneStore::textures
  std::map<std::string, int> spectrum_textures;
  ParamSet material_param; std::string material = "matte";
  std::map<std::string, int> named_material;                         // name -> material row
  std::string current_named_material;
  ParamSet area_light_params; std::string area_light;
  bool reverse_orientation = false;
};

struct Api {
  enum { Uninit, Options, World } state = Uninit;
  FrontendOptions opt;
  ParsedScene* out;
  Xform ctm = Xform::identity();
  std::map<std::string, Xform> named_cs;
  std::vector<Xform> pushed_transforms;
  GraphicsState gs; std::vector<GraphicsState> pushed_gs;
  // object instancing (api.rs:1019-1090): RenderOptions::instances / current_instance
  std::map<std::string, int> object_ids;      // name -> object definition id (a re-definition gets a fresh id, like the re-inserted Vec)
  std::vector<int> object_sizes;              // primitives per definition
  int current_object = -1;
  // RenderOptions defaults (api.rs:278-302)
  std::string film_name = "image", filter_name = "box", sampler_name = "halton", accel_name = "bvh", integrator_name = "path", camera_name = "perspective";
  ParamSet film_params, filter_params, sampler_params, accel_params, integrator_params, camera_params;
  Xform camera_to_world = Xform::identity();

  std::vector<std::string>& warn() { return out->store.warnings; }
  void need_init() { if (state == Uninit) throw ParseError("API not initialized"); }
  void need_options(const char* what) { if (state != Options) throw ParseError(std::string(what) + ": options block only (scene description must be inside the options block)"); }
  void need_world(const char* what) { if (state != World) throw ParseError(std::string(what) + ": world block only"); }
  void check_notes(const ParamSet& ps, const char* where) {
    for (const auto& n : ps.notes) {
      if (n.compare(0, 12, "unsupported:") == 0) throw ParseError(std::string(where) + ": parameter \"" + n.substr(12) + "\" uses a spectrum/blackbody type, which this front end does not support");
      warn().push_back(n);
    }
  }

  static Mat4 mat_from_column_major(const std::vector<float>& t) {   // api.rs:589-592: from_elements(t0,t4,t8,t12, t1,...)
    Mat4 m;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m.at(r, c) = t[c * 4 + r];
    return m;
  }

  int make_material(const std::string& name_in, const TextureParams& mp) {   // api.rs:1141-1183 + material/*.rs create()
    std::string name = name_in;
    rt_material m; std::memset(&m, 0, sizeof(m));
    // a parameter is a constant (folded into its field) or a texture row (tex[slot] = row + 1)
    auto putS = [&m](float* d, int slot, const TexRef& r) { d[0] = r.c.r; d[1] = r.c.g; d[2] = r.c.b; m.tex[slot] = r.row + 1; };
    auto putF = [&m](float& d, int slot, const TexRef& r) { d = r.f; m.tex[slot] = r.row + 1; };
    auto opt_rough = [&](const char* pname, float& d, int slot) { TexRef r; if (!mp.float_texture_or_none(pname, r)) return 0; putF(d, slot, r); return 1; };
    auto eta_or_index = [&]() { TexRef r; if (!mp.float_texture_or_none("eta", r)) r = mp.float_texture("index", 1.5f); putF(m.eta, RT_TS_ETA, r); };
    auto bumpmap = [&]() { TexRef r; if (mp.float_texture_or_none("bumpmap", r)) m.tex[RT_TS_BUMP] = mp.child_row(r, true) + 1; };   // material/*.rs create(): get_float_texture_or_none("bumpmap")
    if (name == "disney" || name == "fourier")
      throw ParseError("Material \"" + name + "\" is not on the GPU hot path yet (SURVEY §8f)");
    if (name != "matte" && name != "plastic" && name != "glass" && name != "mirror" && name != "metal" && name != "uber" && name != "substrate" &&
        name != "translucent" && name != "mix") {
      warn().push_back("Unknown material " + name + ". Using matte.");
      name = "matte";
    }
    m.remap_roughness = 1;
    if (name == "matte") {                                           // matte.rs:20-33
      m.type = RT_MAT_MATTE; putS(m.kd, RT_TS_KD, mp.spectrum_texture("Kd", Rgb{0.5f, 0.5f, 0.5f})); putF(m.sigma, RT_TS_SIGMA, mp.float_texture("sigma", 0.0f));
      bumpmap();
    } else if (name == "plastic") {                                  // plastic.rs:26-41
      m.type = RT_MAT_PLASTIC; putS(m.kd, RT_TS_KD, mp.spectrum_texture("Kd", Rgb{0.25f, 0.25f, 0.25f})); putS(m.ks, RT_TS_KS, mp.spectrum_texture("Ks", Rgb{0.25f, 0.25f, 0.25f}));
      putF(m.roughness, RT_TS_ROUGHNESS, mp.float_texture("roughness", 0.1f)); m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    } else if (name == "glass") {                                    // glass.rs:28-49
      m.type = RT_MAT_GLASS; putS(m.kr, RT_TS_KR, mp.spectrum_texture("Kr", Rgb{1, 1, 1})); putS(m.kt, RT_TS_KT, mp.spectrum_texture("Kt", Rgb{1, 1, 1}));
      eta_or_index();
      putF(m.uroughness, RT_TS_UROUGHNESS, mp.float_texture("uroughness", 0.0f)); putF(m.vroughness, RT_TS_VROUGHNESS, mp.float_texture("vroughness", 0.0f));
      m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    } else if (name == "uber") {                                     // uber.rs:32-57
      m.type = RT_MAT_UBER;
      putS(m.kd, RT_TS_KD, mp.spectrum_texture("Kd", Rgb{0.25f, 0.25f, 0.25f})); putS(m.ks, RT_TS_KS, mp.spectrum_texture("Ks", Rgb{0.25f, 0.25f, 0.25f}));
      putS(m.kr, RT_TS_KR, mp.spectrum_texture("Kr", Rgb{0, 0, 0})); putS(m.kt, RT_TS_KT, mp.spectrum_texture("Kt", Rgb{0, 0, 0}));
      putF(m.roughness, RT_TS_ROUGHNESS, mp.float_texture("roughness", 0.1f));
      m.has_uroughness = opt_rough("uroughness", m.uroughness, RT_TS_UROUGHNESS);
      m.has_vroughness = opt_rough("vroughness", m.vroughness, RT_TS_VROUGHNESS);
      eta_or_index();
      putS(m.opacity, RT_TS_OPACITY, mp.spectrum_texture("opacity", Rgb{1, 1, 1}));
      m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    } else if (name == "substrate") {                                // substrate.rs:23-38
      m.type = RT_MAT_SUBSTRATE;
      putS(m.kd, RT_TS_KD, mp.spectrum_texture("Kd", Rgb{0.5f, 0.5f, 0.5f})); putS(m.ks, RT_TS_KS, mp.spectrum_texture("Ks", Rgb{0.5f, 0.5f, 0.5f}));
      putF(m.uroughness, RT_TS_UROUGHNESS, mp.float_texture("uroughness", 0.1f)); putF(m.vroughness, RT_TS_VROUGHNESS, mp.float_texture("vroughness", 0.1f));
      m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    } else if (name == "translucent") {                              // translucent.rs:26-44
      m.type = RT_MAT_TRANSLUCENT;
      putS(m.kd, RT_TS_KD, mp.spectrum_texture("Kd", Rgb{0.25f, 0.25f, 0.25f})); putS(m.ks, RT_TS_KS, mp.spectrum_texture("Ks", Rgb{0.25f, 0.25f, 0.25f}));
      putS(m.reflect, RT_TS_REFLECT, mp.spectrum_texture("reflect", Rgb{0.5f, 0.5f, 0.5f})); putS(m.transmit, RT_TS_TRANSMIT, mp.spectrum_texture("transmit", Rgb{0.5f, 0.5f, 0.5f}));
      putF(m.roughness, RT_TS_ROUGHNESS, mp.float_texture("roughness", 0.1f));
      m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    } else if (name == "mix") {                                      // api.rs:1165-1176 + mixmat.rs:20-31
      m.type = RT_MAT_MIX;
      const std::string n1 = mp.find_string("namedmaterial1", ""), n2 = mp.find_string("namedmaterial2", "");
      auto child = [&](const std::string& n) {
        auto it = gs.named_material.find(n);
        if (it != gs.named_material.end()) return it->second;
        warn().push_back("Named material \"" + n + "\" undefined. Using \"matte\"");
        return make_material("matte", mp);
      };
      m.mix_a = child(n1); m.mix_b = child(n2);
      putS(m.amount, RT_TS_AMOUNT, mp.spectrum_texture("amount", Rgb{0.5f, 0.5f, 0.5f}));
    } else if (name == "mirror") {                                   // mirror.rs:20-26
      m.type = RT_MAT_MIRROR; putS(m.kr, RT_TS_KR, mp.spectrum_texture("Kr", Rgb{0.9f, 0.9f, 0.9f}));
      bumpmap();
    } else {                                                         // metal.rs:24-46
      m.type = RT_MAT_METAL;
      putS(m.eta_rgb, RT_TS_ETA_RGB, mp.spectrum_texture("eta", Rgb{kCopperEtaRgb[0], kCopperEtaRgb[1], kCopperEtaRgb[2]}));
      putS(m.k_rgb, RT_TS_K_RGB, mp.spectrum_texture("k", Rgb{kCopperKRgb[0], kCopperKRgb[1], kCopperKRgb[2]}));
      putF(m.roughness, RT_TS_ROUGHNESS, mp.float_texture("roughness", 0.01f));
      m.has_uroughness = opt_rough("uroughness", m.uroughness, RT_TS_UROUGHNESS);
      m.has_vroughness = opt_rough("vroughness", m.vroughness, RT_TS_VROUGHNESS);
      m.remap_roughness = mp.find_bool("remaproughness", true);
      bumpmap();
    }
    for (int i = 0; i < RT_TS_COUNT; i++) if (m.tex[i] != 0) m.textured = 1;
    if (m.type == RT_MAT_MIX && (out->store.materials[(size_t)m.mix_a].textured || out->store.materials[(size_t)m.mix_b].textured)) m.textured = 1;
    out->store.materials.push_back(m);
    return (int)out->store.materials.size() - 1;
  }
  int create_material(const ParamSet& shape_params) {                // api.rs:318-340
    TextureParams mp{shape_params, gs.material_param, gs.float_textures, gs.spectrum_textures, &out->store.textures, &warn()};
    if (!gs.current_named_material.empty()) {
      auto it = gs.named_material.find(gs.current_named_material);
      if (it != gs.named_material.end()) return it->second;
      warn().push_back("No material named \"" + gs.current_named_material + "\". Using matte material instead.");
      return make_material("matte", mp);
    }
    return make_material(gs.material, mp);
  }

  // ---- directives ----
  void d_camera(const std::string& name, const ParamSet& ps) {       // api.rs:720-730
    need_options("Camera");
    camera_name = name; camera_params = ps; camera_to_world = ctm.inverse();
    named_cs["camera"] = camera_to_world;
  }
  void d_world_begin() {                                             // api.rs:732-744
    need_options("WorldBegin");
    state = World; named_cs["world"] = ctm; ctm = Xform::identity();
  }
  void d_attribute_begin() { need_world("AttributeBegin"); pushed_gs.push_back(gs); pushed_transforms.push_back(ctm); }
  void d_attribute_end() {                                           // api.rs:756-768
    need_world("AttributeEnd");
    if (pushed_gs.empty()) { warn().push_back("Unmatched AttributeEnd encountered. Ignoring it."); return; }
    gs = pushed_gs.back(); pushed_gs.pop_back();
    ctm = pushed_transforms.back(); pushed_transforms.pop_back();
  }
  void d_transform_begin() { need_world("TransformBegin"); pushed_transforms.push_back(ctm); }
  void d_transform_end() {
    need_world("TransformEnd");
    if (pushed_transforms.empty()) { warn().push_back("Unmatched TransformEnd encountered. Ignoring it."); return; }
    ctm = pushed_transforms.back(); pushed_transforms.pop_back();
  }
  // UVMapping2D / PlanarMapping2D parameters (checkerboard.rs:55-79, uv.rs:24-43, imagemap.rs:99-110)
  void texture_mapping(const TextureParams& tp, rt_texture& t, bool planar_ok, const char* cls) {
    const std::string typ = tp.find_string("mapping", "uv");
    t.mapping = RT_TEXMAP_UV; t.su = t.sv = 1.0f; t.du = t.dv = 0.0f;
    if (typ == "uv") {
      t.su = tp.find_float("uscale", 1.0f); t.sv = tp.find_float("vscale", 1.0f); t.du = tp.find_float("udelta", 0.0f); t.dv = tp.find_float("vdelta", 0.0f);
    } else if (typ == "planar" && planar_ok) {
      t.mapping = RT_TEXMAP_PLANAR;
      Vec3 vs = tp.find_vector3("v1", v3(1, 0, 0)), vt = tp.find_vector3("v2", v3(0, 1, 0));
      t.vs[0] = vs.x; t.vs[1] = vs.y; t.vs[2] = vs.z; t.vt[0] = vt.x; t.vt[1] = vt.y; t.vt[2] = vt.z;
      t.du = tp.find_float("udelta", 0.0f); t.dv = tp.find_float("vdelta", 0.0f);
    } else if (typ == "spherical" || typ == "cylindrical" || typ == "planar") {
      throw ParseError(std::string("Texture \"") + cls + "\": mapping \"" + typ + "\" is unimplemented!() in the reference");
    } else warn().push_back("2D texture mapping \"" + typ + "\" unknown.");
  }
  void d_texture(const std::string& name, const std::string& typ, const std::string& cls, const ParamSet& ps) {   // api.rs:791-846, 1201-1259
    need_world("Texture");
    check_notes(ps, "Texture");
    ParamSet empty;
    TextureParams tp{ps, empty, gs.float_textures, gs.spectrum_textures, &out->store.textures, &warn()};
    const bool is_float = typ == "float";
    if (!is_float && typ != "color" && typ != "spectrum") { warn().push_back("Texture type \"" + typ + "\" unknown."); return; }
    rt_texture t; std::memset(&t, 0, sizeof(t));
    t.is_float = is_float; t.tex1 = t.tex2 = t.amount = -1;
    auto childS = [&](const char* n, Rgb d) { return tp.child_row(tp.spectrum_texture(n, d), false); };
    auto childF = [&](const char* n, float d) { return tp.child_row(tp.float_texture(n, d), true); };
    if (cls == "constant") {                                         // texture/constant.rs:20-33
      t.kind = RT_TEX_CONSTANT;
      if (is_float) t.value[0] = t.value[1] = t.value[2] = tp.find_float("value", 1.0f);
      else { Rgb v = tp.find_spectrum("value", Rgb{1, 1, 1}); t.value[0] = v.r; t.value[1] = v.g; t.value[2] = v.b; }
    } else if (cls == "scale") {                                     // texture/scale.rs:29-43
      t.kind = RT_TEX_SCALE;
      if (is_float) { t.tex1 = childF("tex1", 1.0f); t.tex2 = childF("tex2", 1.0f); }
      else { t.tex1 = childS("tex1", Rgb{1, 1, 1}); t.tex2 = childS("tex2", Rgb{1, 1, 1}); }
    } else if (cls == "mix") {                                       // texture/mix.rs:33-51
      t.kind = RT_TEX_MIX;
      if (is_float) { t.tex1 = childF("tex1", 0.0f); t.tex2 = childF("tex2", 1.0f); }
      else { t.tex1 = childS("tex1", Rgb{0, 0, 0}); t.tex2 = childS("tex2", Rgb{1, 1, 1}); }
      t.amount = childF("amount", 0.5f);
    } else if (cls == "imagemap") {                                  // texture/imagemap.rs:97-139, 160-205
      t.kind = RT_TEX_IMAGEMAP;
      texture_mapping(tp, t, false, "imagemap");
      t.max_aniso = tp.find_float("maxanisotropy", 8.0f);
      t.trilinear = tp.find_bool("trilinear", false);
      const std::string wrap = tp.find_string("wrap", "repeat");
      t.wrap = wrap == "black" ? RT_WRAP_BLACK : (wrap == "clamp" ? RT_WRAP_CLAMP : RT_WRAP_REPEAT);
      const float scale = tp.find_float("scale", 1.0f);
      std::string fn = tp.find_string("filename", "");
      fn = fn.empty() ? fn : resolve_filename(fn, opt.search_dir);
      auto ends = [&](const char* e) { size_t n = std::strlen(e); return fn.size() >= n && fn.compare(fn.size() - n, n, e) == 0; };
      const bool gamma = tp.find_bool("gamma", ends(".tga") || ends(".png"));
      int w = 0, h = 0; std::vector<float> rgb;
      if (!read_image_rgb(fn, w, h, rgb)) {                          // imagemap.rs:62-69
        warn().push_back("Could not open texture file " + fn + ". Using grey texture instead");
        w = h = 1; rgb.assign(3, 0.18f);
      }
      for (int y = 0; y < h / 2; y++)                                // :51-58 flip in y
        for (int x = 0; x < w * 3; x++) std::swap(rgb[(size_t)y * w * 3 + x], rgb[(size_t)(h - 1 - y) * w * 3 + x]);
      auto inv_gamma = [](float v) { return v <= 0.04045f ? v / 12.92f : std::pow((v + 0.055f) * 1.0f / 1.055f, 2.4f); };   // spectrum.rs:379-385
      std::vector<float> texels((size_t)w * h * (is_float ? 1 : 3));
      for (size_t i = 0; i < (size_t)w * h; i++) {                   // :72-83
        float c[3];
        for (int k = 0; k < 3; k++) c[k] = scale * (gamma ? inv_gamma(rgb[i * 3 + k]) : rgb[i * 3 + k]);
        if (is_float) texels[i] = 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2];   // Spectrum::y (spectrum.rs:149-152)
        else { texels[i * 3] = c[0]; texels[i * 3 + 1] = c[1]; texels[i * 3 + 2] = c[2]; }
      }
      t.img_w = w; t.img_h = h; t.texels = out->store.keep(std::move(texels));
    } else if (cls == "fbm") {                                       // texture/fbm.rs:26-37
      t.kind = RT_TEX_FBM; t.w2t = to_ir(ctm);
      t.omega = tp.find_float("omega", 0.5f); t.octaves = tp.find_int("octaves", 8);
    } else if (cls == "uv" && !is_float) {                           // texture/uv.rs:22-46
      t.kind = RT_TEX_UV;
      texture_mapping(tp, t, false, "uv");
    } else if (cls == "checkerboard" && !is_float) {                 // texture/checkerboard.rs:44-97
      t.kind = RT_TEX_CHECKERBOARD;
      const int dim = tp.find_int("dimension", 2);
      if (dim != 2) throw ParseError(dim == 3 ? "3 dimensional checkerboard texture is unimplemented!() in the reference" : "checkerboard texture: unsupported dimension");
      t.tex1 = childS("tex1", Rgb{1, 1, 1}); t.tex2 = childS("tex2", Rgb{0, 0, 0});
      texture_mapping(tp, t, true, "checkerboard");
      const std::string aa = tp.find_string("aamode", "closedform");
      if (aa == "none") t.aa_none = 1;
      else if (aa != "closedform") warn().push_back("Unknown aamethod \"" + aa + "\" found for CheckerboardTexture. Using closedform instead");
    } else if (!is_float && (cls == "bilerp" || cls == "dots" || cls == "wrinkled" || cls == "marble" || cls == "windy" || cls == "ptex")) {
      throw ParseError("Texture class \"" + cls + "\" is unimplemented!() in the reference (api.rs:1234-1253)");
    } else {                                                         // api.rs:1217,1255: Err -> the texture is not registered
      warn().push_back("Failed to create texture " + name + ": Unkown texture type " + cls);
      return;
    }
    out->store.textures.push_back(t);
    auto& reg = is_float ? gs.float_textures : gs.spectrum_textures;
    if (reg.count(name)) warn().push_back("Texture \"" + name + "\" being redefined.");
    reg[name] = (int)out->store.textures.size() - 1;
  }
  void d_make_named_material(const std::string& name, const ParamSet& ps) {   // api.rs:855-881
    check_notes(ps, "MakeNamedMaterial");
    ParamSet empty;
    TextureParams mp{ps, empty, gs.float_textures, gs.spectrum_textures, &out->store.textures, &warn()};
    std::string type = mp.find_string("type", "");
    if (type.empty()) throw ParseError("No parameter string \"type\" found in named_material");
    gs.named_material[name] = make_material(type, mp);
  }
  void d_light(const std::string& name, const ParamSet& ps) {        // api.rs:494-513,905-911
    need_world("LightSource");
    check_notes(ps, "LightSource");
    rt_light l; std::memset(&l, 0, sizeof(l)); l.shape = -1;
    auto mul = [](Rgb a, Rgb b) { return Rgb{a.r * b.r, a.g * b.g, a.b * b.b}; };
    Rgb sc = ps.find_one_spectrum("scale", Rgb{1, 1, 1});
    if (name == "point") {                                           // point.rs:28-35
      Rgb I = mul(ps.find_one_spectrum("I", Rgb{1, 1, 1}), sc);
      Vec3 from = ps.find_one_point3("from", v3(0, 0, 0));
      Xform t = compose(translate(from), ctm);
      Vec3 p = xf_point(t.m, v3(0, 0, 0));
      l.kind = RT_LIGHT_POINT; l.pos[0] = p.x; l.pos[1] = p.y; l.pos[2] = p.z; l.I[0] = I.r; l.I[1] = I.g; l.I[2] = I.b;
    } else if (name == "distant") {                                  // distant.rs:35-42
      Rgb L = mul(ps.find_one_spectrum("L", Rgb{1, 1, 1}), sc);
      Vec3 from = ps.find_one_point3("from", v3(0, 0, 0)), to = ps.find_one_point3("to", v3(0, 0, 1));
      Vec3 d = xf_vector(ctm.m, sub(from, to));
      l.kind = RT_LIGHT_DISTANT; l.dir[0] = d.x; l.dir[1] = d.y; l.dir[2] = d.z; l.I[0] = L.r; l.I[1] = L.g; l.I[2] = L.b;
    } else if (name == "infinite") {                                 // infinite.rs:115-127
      Rgb L = mul(ps.find_one_spectrum("L", Rgb{1, 1, 1}), sc);
      l.kind = RT_LIGHT_INFINITE; l.I[0] = L.r; l.I[1] = L.g; l.I[2] = L.b; l.l2w = to_ir(ctm);
      l.n_samples = ps.find_one_int("samples", 1);
      std::string mapname = ps.find_one_string("mapname", "");
      if (!mapname.empty()) {
        std::string fn = resolve_filename(mapname, opt.search_dir);
        int w, h; std::vector<float> rgb;
        if (read_image_rgb(fn, w, h, rgb)) {                          // infinite.rs:54-61: any format read_image knows, any size
          l.env_w = w; l.env_h = h; l.env_rgb = out->store.keep(std::move(rgb));
        } else warn().push_back("Environment map " + fn + " for infinite light not found! Using constant texture instead.");   // infinite.rs:62-69
      }
    } else throw ParseError("Unsupported light type " + name);       // api.rs:509-512 -> Err -> parse failure
    out->store.lights.push_back(l);
  }
  void d_shape(const std::string& name, const ParamSet& ps) {        // api.rs:913-966, 1093-1139
    need_world("Shape");
    check_notes(ps, "Shape");
    rt_shape s; std::memset(&s, 0, sizeof(s));
    s.o2w = to_ir(ctm); s.reverse_orientation = gs.reverse_orientation; s.material = -1; s.area_light = -1;
    s.object_def = current_object; s.instance_of = -1;
    bool have = false;
    if (name == "sphere") {                                          // sphere.rs:53-67
      s.kind = RT_SHAPE_SPHERE; s.radius = ps.find_one_float("radius", 1.0f);
      s.zmin = ps.find_one_float("zmin", -s.radius); s.zmax = ps.find_one_float("zmax", s.radius); s.phimax = ps.find_one_float("phimax", 360.0f);
      have = true;
    } else if (name == "cylinder") {                                 // cylinder.rs:26-46 (z_min / z_max / phi_max — Q13)
      s.kind = RT_SHAPE_CYLINDER; s.radius = ps.find_one_float("radius", 1.0f);
      s.zmin = ps.find_one_float("z_min", -1.0f); s.zmax = ps.find_one_float("z_max", 1.0f); s.phimax = ps.find_one_float("phi_max", 360.0f);
      have = true;
    } else if (name == "disk") {                                     // disk.rs:47-61
      s.kind = RT_SHAPE_DISK; s.height = ps.find_one_float("height", 0.0f); s.radius = ps.find_one_float("radius", 1.0f);
      s.inner_radius = ps.find_one_float("innerradius", 0.0f); s.phimax = ps.find_one_float("phimax", 360.0f);
      if (!(s.radius > 0.0f && s.inner_radius >= 0.0f && s.phimax > 0.0f)) throw ParseError("disk: radius > 0 && innerradius >= 0 && phimax > 0 required (disk.rs:33)");
      have = true;
    } else if (name == "cone" || name == "paraboloid" || name == "hyperboloid" || name == "curve") {
      throw ParseError("Shape \"" + name + "\" is unimplemented in the reference (api.rs:1109-1115)");
    } else if (name == "trianglemesh" || name == "plymesh") {
      std::vector<int32_t> idx; std::vector<float> P, N, S, uv;
      if (name == "trianglemesh") {                                  // mesh.rs:76-171
        if (const auto* vi = ps.find_int("indices")) idx = *vi;
        if (const auto* pp = ParamSet::lookup(ps.point3s, "P")) for (Vec3 p : pp->values) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
        const auto* uvp = ParamSet::lookup(ps.point2s, "uv"); if (!uvp) uvp = ParamSet::lookup(ps.point2s, "st");
        if (uvp) for (auto& q : uvp->values) { uv.push_back(q.first); uv.push_back(q.second); }
        else {
          const std::vector<float>* fuv = ps.find_float("uv"); if (!fuv) fuv = ps.find_float("st");
          if (fuv) { for (size_t i = 0; i + 2 <= fuv->size(); i += 2) { uv.push_back((*fuv)[i]); uv.push_back((*fuv)[i + 1]); } if (fuv->size() % 2) throw ParseError("trianglemesh: odd number of uv floats"); }
        }
        if (idx.empty()) { warn().push_back("Vertex indices \"indices\" not provided with triangle mesh shape"); return; }
        if (P.empty()) { warn().push_back("Vertex positions \"P\" not provided with triangle mesh shape"); return; }
        if (const auto* sp = ParamSet::lookup(ps.vector3s, "S")) { if (sp->values.size() * 3 != P.size()) warn().push_back("Number of \"S\"s for mesh triangle must match \"P\"s"); else for (Vec3 v : sp->values) { S.push_back(v.x); S.push_back(v.y); S.push_back(v.z); } }
        if (const auto* np = ParamSet::lookup(ps.normal3s, "N")) { if (np->values.size() * 3 != P.size()) warn().push_back("Number of \"N\"s for mesh triangle must match \"P\"s"); else for (Vec3 v : np->values) { N.push_back(v.x); N.push_back(v.y); N.push_back(v.z); } }
      } else {                                                       // plymesh.rs:18-178
        std::string fn = ps.find_one_string("filename", "");
        fn = fn.empty() ? fn : resolve_filename(fn, opt.search_dir);
        PlyMesh pm;
        read_ply(fn, pm);
        idx = std::move(pm.indices); P = std::move(pm.P); N = std::move(pm.N); uv = std::move(pm.uv);
        if (P.empty() || idx.empty()) { warn().push_back("PLY file \"" + fn + "\" is invalid! No face/vertex elements found!"); return; }
      }
      if (!ps.find_texture("alpha").empty() || !ps.find_texture("shadowalpha").empty()) throw ParseError("alpha-mask textures are outside the GPU path's scope");
      if (ps.find_one_float("alpha", 1.0f) == 0.0f || ps.find_one_float("shadowalpha", 1.0f) == 0.0f) throw ParseError("constant-zero alpha masks are outside the GPU path's scope");
      size_t nv = P.size() / 3;
      for (int32_t i : idx) if (i < 0 || (size_t)i >= nv) throw ParseError("triangle mesh index out of range");
      if (!uv.empty() && uv.size() / 2 < nv) throw ParseError("triangle mesh: fewer uv than P");
      s.kind = RT_SHAPE_TRIMESH; s.n_indices = (uint32_t)(idx.size() / 3 * 3); s.n_vertices = (uint32_t)nv;
      s.indices = out->store.keep(std::move(idx)); s.P = out->store.keep(std::move(P));
      s.N = N.empty() ? nullptr : out->store.keep(std::move(N));
      s.S = S.empty() ? nullptr : out->store.keep(std::move(S));
      s.uv = uv.empty() ? nullptr : out->store.keep(std::move(uv));
      have = s.n_indices > 0;
    } else { warn().push_back("Unknown shape " + name); return; }
    if (!have) return;
    s.material = create_material(ps);
    if (current_object >= 0) {                                       // api.rs:951-957: the primitives go to the instance, not to the scene
      if (!gs.area_light.empty())
        throw ParseError("AreaLightSource inside ObjectBegin/ObjectEnd: the reference drops such lights from the light list but still lets the surface emit "
                         "(api.rs:951-960); not supported on the GPU path");
      object_sizes[current_object] += s.kind == RT_SHAPE_TRIMESH ? (int)(s.n_indices / 3) : 1;
      out->store.shapes.push_back(s);
      return;
    }
    if (!gs.area_light.empty()) {                                    // api.rs:934-946,1185-1199 ; diffuse.rs:39-51
      if (gs.area_light != "area" && gs.area_light != "diffuse") throw ParseError("Area light " + gs.area_light + " unknown");
      const ParamSet& ap = gs.area_light_params;
      Rgb L = ap.find_one_spectrum("L", Rgb{1, 1, 1}), sc = ap.find_one_spectrum("scale", Rgb{1, 1, 1});
      rt_area_light al; al.L[0] = L.r * sc.r; al.L[1] = L.g * sc.g; al.L[2] = L.b * sc.b;
      al.n_samples = ap.find_one_int("samples", ap.find_one_int("nsamples", 1));
      al.two_sided = ap.find_one_bool("twosided", false);
      out->store.area_lights.push_back(al);
      s.area_light = (int)out->store.area_lights.size() - 1;
    }
    out->store.shapes.push_back(s);
    if (s.area_light >= 0) {                                         // area lights appended after the shape's primitives (api.rs:963)
      rt_light l; std::memset(&l, 0, sizeof(l)); l.kind = RT_LIGHT_AREA; l.shape = (int)out->store.shapes.size() - 1;
      out->store.lights.push_back(l);
    }
  }

  void d_object_begin(const std::string& name) {                     // api.rs:1019-1034
    d_attribute_begin();
    if (current_object >= 0) throw ParseError("ObjectBegin called inside of instance definition");
    current_object = (int)object_sizes.size();
    object_sizes.push_back(0);
    object_ids[name] = current_object;
  }
  void d_object_end() {                                              // api.rs:1036-1050
    need_world("ObjectEnd");
    if (current_object < 0) throw ParseError("ObjectEnd called outside of instance definition ");
    current_object = -1;
    d_attribute_end();
  }
  void d_object_instance(const std::string& name) {                  // api.rs:1052-1090
    need_world("ObjectInstance");
    if (current_object >= 0) throw ParseError("ObjectInstance called inside of instance definition");
    auto it = object_ids.find(name);
    if (it == object_ids.end()) throw ParseError("Unable to find instance named " + name);
    if (object_sizes[it->second] == 0) return;                       // empty definition: nothing is added (api.rs:1067-1069)
    rt_shape s; std::memset(&s, 0, sizeof(s));
    s.kind = RT_SHAPE_INSTANCE; s.o2w = to_ir(ctm); s.material = -1; s.area_light = -1; s.object_def = -1; s.instance_of = it->second;
    out->store.shapes.push_back(s);
  }

  void d_world_end() {                                               // api.rs:977-1017 (+ make_* :181-276)
    need_world("WorldEnd");
    while (!pushed_gs.empty()) { warn().push_back("Missing AttributeEnd"); pushed_gs.pop_back(); if (!pushed_transforms.empty()) pushed_transforms.pop_back(); }
    while (!pushed_transforms.empty()) { warn().push_back("Missing TransformEnd!"); pushed_transforms.pop_back(); }
    rt_scene& v = out->store.view;
    // make_filter
    rt_film& f = v.film;
    check_notes(filter_params, "PixelFilter"); check_notes(film_params, "Film"); check_notes(camera_params, "Camera");
    check_notes(sampler_params, "Sampler"); check_notes(integrator_params, "Integrator"); check_notes(accel_params, "Accelerator");
    if (filter_name == "box") { f.filter = RT_FILTER_BOX; f.filter_xw = filter_params.find_one_float("xwidth", 0.5f); f.filter_yw = filter_params.find_one_float("ywidth", 0.5f); }
    else if (filter_name == "mitchell") { f.filter = RT_FILTER_MITCHELL; f.filter_xw = filter_params.find_one_float("xwidth", 2.0f); f.filter_yw = filter_params.find_one_float("ywidth", 2.0f); f.filter_a = filter_params.find_one_float("B", 1.0f / 3.0f); f.filter_b = filter_params.find_one_float("C", 1.0f / 3.0f); }
    else if (filter_name == "gaussian") { f.filter = RT_FILTER_GAUSSIAN; f.filter_xw = filter_params.find_one_float("xwidth", 2.0f); f.filter_yw = filter_params.find_one_float("ywidth", 2.0f); f.filter_a = filter_params.find_one_float("alpha", 2.0f); }
    else if (filter_name == "triangle") { f.filter = RT_FILTER_TRIANGLE; f.filter_xw = filter_params.find_one_float("xwidth", 2.0f); f.filter_yw = filter_params.find_one_float("ywidth", 2.0f); }
    else throw ParseError("Filter \"" + filter_name + "\" unknown.");
    // make_film (film.rs:117-150)
    if (film_name != "image") throw ParseError("Film \"" + film_name + "\" unknown.");
    std::string fname = film_params.find_one_string("filename", "");
    out->film_filename = fname.empty() ? "image.png" : "rt-" + fname;
    f.xres = film_params.find_one_int("xresolution", 1280); f.yres = film_params.find_one_int("yresolution", 720);
    f.crop[0] = 0; f.crop[1] = 1; f.crop[2] = 0; f.crop[3] = 1;
    if (const auto* cr = film_params.find_float("cropwindow")) {
      if (cr->size() == 4) {
        f.crop[0] = clampf(std::fmin((*cr)[0], (*cr)[1]), 0.0f, 1.0f); f.crop[1] = clampf(std::fmax((*cr)[0], (*cr)[1]), 0.0f, 1.0f);
        f.crop[2] = clampf(std::fmin((*cr)[2], (*cr)[3]), 0.0f, 1.0f); f.crop[3] = clampf(std::fmax((*cr)[2], (*cr)[3]), 0.0f, 1.0f);
      } else warn().push_back("\"cropwindow\" expected 4 values");
    }
    f.scale = film_params.find_one_float("scale", 1.0f);
    f.max_sample_luminance = film_params.find_one_float("maxsampleluminance", std::numeric_limits<float>::infinity());
    // make_camera (camera.rs:74-123)
    if (camera_name != "perspective") throw ParseError("Camera \"" + camera_name + "\" unknown.");
    rt_camera& c = v.camera;
    c.c2w = to_ir(camera_to_world);
    c.lens_radius = camera_params.find_one_float("lensradius", 0.0f); c.focal_distance = camera_params.find_one_float("focaldistance", 1e6f);
    float frame = camera_params.find_one_float("frameaspectratio", (float)f.xres / (float)f.yres);
    if (frame > 1.0f) { c.screen_window[0] = -frame; c.screen_window[1] = frame; c.screen_window[2] = -1.0f; c.screen_window[3] = 1.0f; }
    else { c.screen_window[0] = -1.0f; c.screen_window[1] = 1.0f; c.screen_window[2] = -1.0f / frame; c.screen_window[3] = 1.0f / frame; }
    if (const auto* sw = camera_params.find_float("screenwindow")) {
      if (sw->size() == 4) for (int i = 0; i < 4; i++) c.screen_window[i] = (*sw)[i];
      else warn().push_back("\"screenwindow\" should have 4 values");
    }
    c.fov = camera_params.find_one_float("fov", 90.0f);
    float halffov = camera_params.find_one_float("halffov", -1.0f);
    if (halffov > 0.0f) c.fov = halffov * 2.0f;
    // make_integrator (api.rs:231-246)
    rt_integrator& in = v.integrator; std::memset(&in, 0, sizeof(in));
    std::string iname = integrator_name;
    if (opt.gpu_integrator_names) {
      if (iname == "gpupath") iname = "path"; else if (iname == "gpuwhitted") iname = "whitted"; else if (iname == "gpudirectlighting") iname = "directlighting";
      else if (iname == "gpunormal") iname = "normal"; else if (iname == "gpuao") iname = "ambientocclusion";
    }
    out->integrator_name = iname;
    in.max_depth = integrator_params.find_one_int("maxdepth", 5);
    in.rr_threshold = 1.0f; in.light_strategy = RT_LIGHTSTRATEGY_SPATIAL; in.ao_samples = 64;
    if (iname == "whitted") in.type = RT_INTEGRATOR_WHITTED;
    else if (iname == "directlighting") {                            // directlighting.rs:46-62
      in.type = RT_INTEGRATOR_DIRECT;
      std::string st = integrator_params.find_one_string("strategy", "all");
      if (st == "one") in.direct_strategy = RT_DIRECT_ONE;
      else { if (st != "all") warn().push_back("Strategy \"" + st + "\" for directlighting unknown. Using \"all\"."); in.direct_strategy = RT_DIRECT_ALL; }
    } else if (iname == "path") {                                    // path.rs:49-78
      in.type = RT_INTEGRATOR_PATH;
      in.rr_threshold = integrator_params.find_one_float("rrthreshold", 1.0f);
      in.light_strategy = integrator_params.find_one_string("lightsamplestrategy", "spatial") == "uniform" ? RT_LIGHTSTRATEGY_UNIFORM : RT_LIGHTSTRATEGY_SPATIAL;
      if (const auto* pb = integrator_params.find_int("pixelbounds")) {
        if (pb->size() != 4) warn().push_back("Expected 4 values for \"pixelbounds\" parameter.");
        else { in.has_pixel_bounds = 1; for (int i = 0; i < 4; i++) in.pixel_bounds[i] = (*pb)[i]; }
      }
    } else if (iname == "normal") in.type = RT_INTEGRATOR_NORMAL;
    else if (iname == "ambientocclusion" && opt.gpu_integrator_names) { in.type = RT_INTEGRATOR_AO; in.ao_samples = integrator_params.find_one_int("nsamples", 64); }
    else throw ParseError("Integrator \"" + integrator_name + "\" unknown.");
    // make_sampler (api.rs:205-215)
    if (sampler_name != "lowdiscrepancy" && sampler_name != "02sequence") throw ParseError("Sampler \"" + sampler_name + "\" unknown.");
    v.sampler.spp = sampler_params.find_one_int("pixelsamples", 16); v.sampler.dimensions = sampler_params.find_one_int("dimensions", 4);
    // make_accelerator (api.rs:263-276 ; bvh/mod.rs:63-78)
    if (accel_name == "kdtree") throw ParseError("Accelerator \"kdtree\" is unimplemented in the reference (api.rs:269)");
    if (accel_name != "bvh") warn().push_back("Accelerator \"" + accel_name + "\" unknown.");
    std::string sm = accel_params.find_one_string("splitmethod", "sah");
    if (sm == "sah") v.accel.split_method = RT_SPLIT_SAH;
    else if (sm == "middle") v.accel.split_method = RT_SPLIT_MIDDLE;
    else { warn().push_back("Unknown (or unimplemented) BVH split method " + sm + ".  Using \"sah\""); v.accel.split_method = RT_SPLIT_SAH; }
    v.accel.max_node_prims = accel_params.find_one_int("maxnodeprims", 4);
    out->store.finish();
    out->world_ended = true;
  }
};

struct Parser {                                                      // pbrt/parser.rs:20-188
  const std::vector<Token>& t; size_t p = 0; Api& api;
  Parser(const std::vector<Token>& toks, Api& a) : t(toks), api(a) {}
  [[noreturn]] void fail(const std::string& why) { throw ParseError("Failed to parse scene file: " + why + " (token " + std::to_string(p) + ")"); }
  float num() { if (p >= t.size() || t[p].kind != Tok::NUMBER) fail("expected a number"); return t[p++].num; }
  std::string str() { if (p >= t.size() || t[p].kind != Tok::STR) fail("expected a string"); return t[p++].str; }
  std::vector<float> num_array() {                                   // :259-264
    std::vector<float> v;
    if (p < t.size() && t[p].kind == Tok::LBRACK) {
      p++;
      while (p < t.size() && t[p].kind == Tok::NUMBER) v.push_back(t[p++].num);
      if (v.empty() || p >= t.size() || t[p].kind != Tok::RBRACK) fail("malformed number array");
      p++;
    } else v.push_back(num());
    return v;
  }
  ParamSet param_list() {                                            // :190-196, :233-257
    ParamSet ps;
    while (p < t.size() && t[p].kind == Tok::STR) {
      ParamType type; std::string name;
      if (!parse_param_header(t[p].str, type, name)) break;          // many0 stops; the directive loop then fails on the stray string
      p++;
      std::vector<float> nums; std::vector<std::string> strs;
      if (p < t.size() && t[p].kind == Tok::LBRACK && p + 1 < t.size() && t[p + 1].kind == Tok::STR) {
        p++;
        while (p < t.size() && t[p].kind == Tok::STR) strs.push_back(t[p++].str);
        if (p >= t.size() || t[p].kind != Tok::RBRACK) fail("malformed string array");
        p++;
      } else if (p < t.size() && t[p].kind == Tok::STR) strs.push_back(t[p++].str);
      else nums = num_array();
      ps.add(type, name, nums, strs);
    }
    return ps;
  }
  void run(int include_depth = 0) {
    if (t.empty()) fail("empty token stream");
    while (p < t.size()) {
      if (api.out->world_ended) return;                              // world_end renders in the reference; nothing after it matters here
      Tok k = t[p].kind;
      p++;
      switch (k) {
        case Tok::ACCELERATOR: { std::string n = str(); ParamSet ps = param_list(); api.need_options("Accelerator"); api.accel_name = n; api.accel_params = ps; break; }
        case Tok::ATTRIBUTEBEGIN: api.d_attribute_begin(); break;
        case Tok::ATTRIBUTEEND: api.d_attribute_end(); break;
        case Tok::TRANSFORMBEGIN: api.d_transform_begin(); break;
        case Tok::TRANSFORMEND: api.d_transform_end(); break;
        case Tok::OBJECTBEGIN: { std::string n = str(); api.d_object_begin(n); break; }
        case Tok::OBJECTEND: api.d_object_end(); break;
        case Tok::OBJECTINSTANCE: { std::string n = str(); api.d_object_instance(n); break; }
        case Tok::WORLDBEGIN: api.d_world_begin(); break;
        case Tok::WORLDEND: api.d_world_end(); break;
        case Tok::LOOKAT: {                                          // api.rs:656-680
          float a[9]; for (float& x : a) x = num();
          api.need_init();
          Xform la;
          if (!look_at(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), v3(a[6], a[7], a[8]), la)) api.warn().push_back("\"up\" vector and viewing direction passed to LookAt are pointing in the same direction.  Using the identity transformation.");
          api.ctm = compose(api.ctm, la);
          break;
        }
        case Tok::COORDINATESYSTEM: { std::string n = str(); api.need_init(); api.named_cs[n] = api.ctm; break; }
        case Tok::COORDSYSTRANSFORM: {
          std::string n = str(); api.need_init();
          auto it = api.named_cs.find(n);
          if (it != api.named_cs.end()) api.ctm = it->second; else api.warn().push_back("Couldn't find named coordinate system \"" + n + "\"");
          break;
        }
        case Tok::CAMERA: { std::string n = str(); ParamSet ps = param_list(); api.d_camera(n, ps); break; }
        case Tok::FILM: { std::string n = str(); ParamSet ps = param_list(); api.need_options("Film"); api.film_name = n; api.film_params = ps; break; }
        case Tok::INCLUDE: {                                         // parser.rs:73-81
          std::string n = str();
          if (include_depth > 16) fail("Include nesting too deep");
          std::vector<Token> inc = tokenize_file(n, api.opt.search_dir);
          Parser sub(inc, api);
          sub.run(include_depth + 1);
          break;
        }
        case Tok::INTEGRATOR: { std::string n = str(); ParamSet ps = param_list(); api.need_options("Integrator"); api.integrator_name = n; api.integrator_params = ps; break; }
        case Tok::AREALIGHTSOURCE: { std::string n = str(); ParamSet ps = param_list(); api.check_notes(ps, "AreaLightSource"); api.gs.area_light = n; api.gs.area_light_params = ps; break; }
        case Tok::LIGHTSOURCE: { std::string n = str(); ParamSet ps = param_list(); api.d_light(n, ps); break; }
        case Tok::MATERIAL: { std::string n = str(); ParamSet ps = param_list(); api.check_notes(ps, "Material"); api.gs.material = n; api.gs.material_param = ps; api.gs.current_named_material.clear(); break; }
        case Tok::MAKENAMEDMATERIAL: { std::string n = str(); ParamSet ps = param_list(); api.d_make_named_material(n, ps); break; }
        case Tok::NAMEDMATERIAL: { std::string n = str(); api.need_world("NamedMaterial"); api.gs.current_named_material = n; break; }
        case Tok::SAMPLER: { std::string n = str(); ParamSet ps = param_list(); api.need_options("Sampler"); api.sampler_name = n; api.sampler_params = ps; break; }
        case Tok::SHAPE: { std::string n = str(); ParamSet ps = param_list(); api.d_shape(n, ps); break; }
        case Tok::REVERSEORIENTATION: api.need_world("ReverseOrientation"); api.gs.reverse_orientation = !api.gs.reverse_orientation; break;
        case Tok::PIXELFILTER: { std::string n = str(); ParamSet ps = param_list(); api.need_options("PixelFilter"); api.filter_name = n; api.filter_params = ps; break; }
        case Tok::SCALE: { float x = num(), y = num(), z = num(); api.need_init(); api.ctm = compose(api.ctm, scaling(x, y, z)); break; }
        case Tok::ROTATE: { float a = num(), x = num(), y = num(), z = num(); api.need_init(); api.ctm = compose(api.ctm, rotate(a, v3(x, y, z))); break; }
        case Tok::TEXTURE: { std::string n = str(), ty = str(), cl = str(); ParamSet ps = param_list(); api.d_texture(n, ty, cl, ps); break; }
        case Tok::CONCATTRANSFORM: {                                 // api.rs:566-598
          std::vector<float> v = num_array(); if (v.size() < 16) fail("ConcatTransform needs 16 numbers");
          api.need_init(); Mat4 m = Api::mat_from_column_major(v); api.ctm = compose(api.ctm, Xform{m, invert(m)}); break;
        }
        case Tok::TRANSFORM: {                                       // api.rs:600-632
          std::vector<float> v = num_array(); if (v.size() < 16) fail("Transform needs 16 numbers");
          api.need_init(); Mat4 m = Api::mat_from_column_major(v); api.ctm = Xform{m, invert(m)}; break;
        }
        case Tok::TRANSLATE: { float x = num(), y = num(), z = num(); api.need_init(); api.ctm = compose(api.ctm, translate(v3(x, y, z))); break; }
        default:
          // tokenised but not parsed by the reference (Identity, ActiveTransform, ...), or a stray value
          p--; fail(std::string("unexpected token ") + tok_name(k));
      }
    }
  }
};

}  // namespace

std::unique_ptr<ParsedScene> parse_scene_text(const std::string& text, const FrontendOptions& opt) {
  std::vector<Token> raw = tokenize(text), toks;
  for (auto& t : raw) if (t.kind != Tok::COMMENT) toks.push_back(std::move(t));
  std::unique_ptr<ParsedScene> ps(new ParsedScene());
  Api api; api.opt = opt; api.out = ps.get(); api.state = Api::Options;   // api.init() (api.rs:517-524)
  Parser parser(toks, api);
  parser.run();
  if (!ps->world_ended) throw ParseError("scene has no WorldEnd");
  return ps;
}
std::unique_ptr<ParsedScene> parse_scene_file(const std::string& filename, FrontendOptions opt) {   // pbrt/mod.rs:15-25
  std::string text = read_file(filename);
  if (opt.search_dir.empty()) opt.search_dir = dir_of(filename);
  return parse_scene_text(text, opt);
}

}  // namespace rth
