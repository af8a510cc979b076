// PBRT scene-file front end (host, product code): tokenizer, directive parser, ParamSet and the API
// state machine, restating the subset rustracer accepts:
//   rustracer-core/src/pbrt/lexer.rs, pbrt/parser.rs, pbrt/mod.rs, paramset.rs, api.rs, fileutil.rs.
// Output: a SceneStore (rt_scene) — the state at `RealApi::world_end` (api.rs:977-1010).
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "scene_store.hpp"

namespace rth {

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- lexer (pbrt/lexer.rs:21-67,185-263) ----
enum class Tok {
  ACCELERATOR, ACTIVETRANSFORM, ALL, AREALIGHTSOURCE, ATTRIBUTEBEGIN, ATTRIBUTEEND, CAMERA, CONCATTRANSFORM, COORDINATESYSTEM,
  COORDSYSTRANSFORM, ENDTIME, FILM, IDENTITY, INCLUDE, LIGHTSOURCE, LOOKAT, MAKENAMEDMEDIUM, MAKENAMEDMATERIAL, MATERIAL,
  MEDIUMINTERFACE, NAMEDMATERIAL, OBJECTBEGIN, OBJECTEND, OBJECTINSTANCE, PIXELFILTER, REVERSEORIENTATION, ROTATE, SAMPLER, SCALE,
  SHAPE, STARTTIME, INTEGRATOR, TEXTURE, TRANSFORMBEGIN, TRANSFORMEND, TRANSFORMTIMES, TRANSFORM, TRANSLATE, WORLDBEGIN, WORLDEND,
  STR, NUMBER, LBRACK, RBRACK, COMMENT
};
struct Token { Tok kind; std::string str; float num = 0; };
const char* tok_name(Tok t);
// Whole-input tokenizer; comments are kept as COMMENT tokens (stripped by tokenize_file, pbrt/mod.rs:38-42).
std::vector<Token> tokenize(const std::string& input);

// ---- ParamSet (paramset.rs) ----
enum class ParamType { Int, Bool, Float, Point2, Vector2, Point3, Vector3, Normal, Rgb, Xyz, Blackbody, Spectrum, String, Texture };
struct Rgb { float r, g, b; };
template <class T> struct ParamItem { std::string name; std::vector<T> values; };
struct ParamSet {
  std::vector<ParamItem<bool>> bools;
  std::vector<ParamItem<int32_t>> ints;
  std::vector<ParamItem<float>> floats;
  std::vector<ParamItem<std::string>> strings;
  std::vector<ParamItem<Rgb>> spectra;
  std::vector<ParamItem<std::pair<float, float>>> point2s;
  std::vector<ParamItem<Vec3>> point3s, vector3s, normal3s;
  std::vector<ParamItem<std::string>> textures;
  std::vector<std::string> notes;   // unimplemented / unsupported parameter types met while building

  void add(ParamType t, const std::string& name, const std::vector<float>& nums, const std::vector<std::string>& strs);
  template <class T> static const ParamItem<T>* lookup(const std::vector<ParamItem<T>>& v, const std::string& n) {
    for (const auto& e : v) if (e.name == n) return &e;   // first match (paramset.rs:20-31)
    return nullptr;
  }
  bool find_one_bool(const std::string& n, bool d) const { auto* e = lookup(bools, n); return e ? e->values.at(0) : d; }
  int32_t find_one_int(const std::string& n, int32_t d) const { auto* e = lookup(ints, n); return e ? e->values.at(0) : d; }
  float find_one_float(const std::string& n, float d) const { auto* e = lookup(floats, n); return e ? e->values.at(0) : d; }
  std::string find_one_string(const std::string& n, const std::string& d) const { auto* e = lookup(strings, n); return e ? e->values.at(0) : d; }
  Rgb find_one_spectrum(const std::string& n, Rgb d) const { auto* e = lookup(spectra, n); return e ? e->values.at(0) : d; }
  Vec3 find_one_point3(const std::string& n, Vec3 d) const { auto* e = lookup(point3s, n); return e ? e->values.at(0) : d; }
  Vec3 find_one_vector3(const std::string& n, Vec3 d) const { auto* e = lookup(vector3s, n); return e ? e->values.at(0) : d; }
  std::string find_texture(const std::string& n) const { auto* e = lookup(textures, n); return e ? e->values.at(0) : std::string(); }
  const std::vector<float>* find_float(const std::string& n) const { auto* e = lookup(floats, n); return e ? &e->values : nullptr; }
  const std::vector<int32_t>* find_int(const std::string& n) const { auto* e = lookup(ints, n); return e ? &e->values : nullptr; }
};
// "integer indices" -> (Int, "indices")  (pbrt/parser.rs:198-231)
bool parse_param_header(const std::string& s, ParamType& type, std::string& name);

// ---- parser + API ----
struct FrontendOptions {
  std::string search_dir;            // fileutil.rs:11-16: directory of the top-level scene file
  bool gpu_integrator_names = true;  // accept "gpupath", "gpuwhitted", "gpudirectlighting", "gpuao", "gpunormal" and "ambientocclusion" 
};
struct ParsedScene {
  SceneStore store;
  std::string integrator_name, film_filename;
  bool world_ended = false;
};
// Parse a token stream / text / file into a scene.  Throws ParseError with the reference's failure modes
// (unknown keyword, tokenised-but-unparsed directives, wrong block, unknown sampler/film/camera/filter names...).
std::unique_ptr<ParsedScene> parse_scene_text(const std::string& text, const FrontendOptions& opt);
std::unique_ptr<ParsedScene> parse_scene_file(const std::string& filename, FrontendOptions opt);

// ---- PLY (shapes/plymesh.rs:18-178): ascii / binary_little_endian / binary_big_endian ----
struct PlyMesh { std::vector<float> P, N, uv; std::vector<int32_t> indices; };
void read_ply(const std::string& filename, PlyMesh& out);

// Copper eta / k as RGB: tests/golden/copper_rgb.json (generated from the reference's SPD + CIE tables).
extern const float kCopperEtaRgb[3];
extern const float kCopperKRgb[3];

}  // namespace rth
