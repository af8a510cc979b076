// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Lights, light distributions, scene,
// camera, film, integrators and the tile renderer.  Follows /root/reference/rustracer-core/src/
// light/*.rs, lightdistrib.rs, scene.rs, camera.rs, film.rs, filter/*.rs, integrator/*.rs, renderer.rs.
#pragma once
#include "orc_bvh.hpp"
#include "orc_bsdf.hpp"
#include "orc_sampler.hpp"
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace orc {

// sampling/distribution1d.rs
struct Distribution1D {
  std::vector<float> func, cdf; float func_int = 0;
  Distribution1D() {}
  explicit Distribution1D(const std::vector<float>& f) { init(f.data(), f.size()); }
  void init(const float* f, size_t n) {                          // :11-45
    func.assign(f, f + n); cdf.assign(n + 1, 0.0f);
    for (size_t i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] / (float)n;
    func_int = cdf[n];
    if (func_int == 0.0f) for (size_t i = 1; i < n + 1; i++) cdf[i] = (float)i / (float)n;
    else for (size_t i = 1; i < n + 1; i++) cdf[i] /= func_int;
  }
  size_t count() const { return func.size(); }
  float sample_continuous(float u, float& pdf, size_t& offset) const {   // :51-68
    offset = find_interval(cdf.size(), [&](size_t i) { return cdf[i] <= u; });
    float du = u - cdf[offset];
    if (cdf[offset + 1] - cdf[offset] > 0.0f) du /= cdf[offset + 1] - cdf[offset];
    pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
    return ((float)offset + du) / (float)count();
  }
  size_t sample_discrete(float u, float& pdf) const {            // :70-79
    size_t offset = find_interval(cdf.size(), [&](size_t i) { return cdf[i] <= u; });
    pdf = func_int > 0.0f ? func[offset] / (func_int * (float)count()) : 0.0f;
    return offset;
  }
};
// sampling/distribution2d.rs
struct Distribution2D {
  std::vector<Distribution1D> cond; Distribution1D marginal;
  void init(const float* func, size_t nu, size_t nv) {           // :11-28
    cond.resize(nv);
    std::vector<float> mf(nv);
    for (size_t v = 0; v < nv; v++) { cond[v].init(func + v * nu, nu); mf[v] = cond[v].func_int; }
    marginal.init(mf.data(), nv);
  }
  P2 sample_continuous(P2 u, float& pdf) const {                 // :30-35
    float p1, p0; size_t v, dummy;
    float d1 = marginal.sample_continuous(u.y, p1, v);
    float d0 = cond[v].sample_continuous(u.x, p0, dummy);
    pdf = p0 * p1;
    return P2(d0, d1);
  }
  float pdf(P2 p) const {                                        // :37-49
    int64_t iu = clampv<int64_t>(f2usize(p.x * (float)cond[0].count()), 0, (int64_t)cond[0].count() - 1);
    int64_t iv = clampv<int64_t>(f2usize(p.y * (float)marginal.count()), 0, (int64_t)marginal.count() - 1);
    return cond[iv].func[iu] / marginal.func_int;
  }
};

struct Scene;
struct Light {                                                   // light/mod.rs:66-97
  int kind; int id;
  // point / distant
  V3 pos, dir; Spectrum I;
  // area
  const Shape* shape = nullptr; Spectrum l_emit; bool two_sided = false; float area = 0;
  int n_samples = 1;
  // infinite
  Transform l2w, w2l; int env_w = 1, env_h = 1; std::vector<Spectrum> texels; Distribution2D distribution;
  MIPMap l_map;                                                  // infinite.rs:71-77: MIPMap::new(resolution, texels, false, 0.0, Repeat)
  // preprocess
  V3 w_center; float w_radius = 0;

  bool is_delta() const { return kind == RT_LIGHT_POINT || kind == RT_LIGHT_DISTANT; }   // light/mod.rs:38-40
  Spectrum L(const Interaction& si, V3 w) const {                // diffuse.rs:91-97
    return (two_sided || dot(si.n, w) > 0.0f) ? l_emit : Spectrum(0.0f);
  }
  // MIPMap::lookup(st, width) (mipmap.rs:227-245) on the light's own pyramid: non-power-of-two maps were Lanczos-resampled
  // by MIPMap::new (mipmap.rs:73-139), and width > 0 can select a coarser level (aspect ratios of 4:1 and beyond)
  Spectrum map_lookup(P2 st, float width) const {
    Texel t = l_map.lookup(st, width);
    return Spectrum(t.c[0], t.c[1], t.c[2]);
  }
  void init_infinite() {                                         // infinite.rs:46-113
    std::vector<float> rgb((size_t)env_w * env_h * 3);
    for (size_t i = 0; i < texels.size(); i++) { rgb[3 * i] = texels[i].r; rgb[3 * i + 1] = texels[i].g; rgb[3 * i + 2] = texels[i].b; }
    l_map.build(env_w, env_h, rgb.data(), 3, false, 0.0f, RT_WRAP_REPEAT);
    int width = 2 * l_map.res_x, height = 2 * l_map.res_y;       // :80 (l_map.width() is the resampled resolution)
    const float filter = 0.5f / fmin_((float)width, (float)height);
    std::vector<float> img((size_t)width * height);
    for (int v = 0; v < height; v++) {
      float vp = ((float)v + 0.5f) / (float)height;
      float sin_theta = std::sin(PI * ((float)v + 0.5f) / (float)height);
      for (int u = 0; u < width; u++) {
        float up = ((float)u + 0.5f) / (float)width;
        img[(size_t)v * width + u] = map_lookup(P2(up, vp), filter).y() * sin_theta;
      }
    }
    distribution.init(img.data(), width, height);
  }
  Spectrum le(const Ray& ray) const {                            // infinite.rs:210-219 ; light/mod.rs:92-94
    if (kind != RT_LIGHT_INFINITE) return Spectrum(0.0f);
    V3 w = normalize(w2l.vector(ray.d));
    P2 st(spherical_phi(w) * INV_PI * 0.5f, spherical_theta(w) * INV_PI);
    return map_lookup(st, 0.0f);
  }
  // returns Li; wi, pdf; p0/p1 of the VisibilityTester
  Spectrum sample_li(const Interaction& isect, P2 u, V3& wi, float& pdf, Interaction& p1) const {
    switch (kind) {
      case RT_LIGHT_POINT: {                                     // point.rs:43-54
        V3 w = pos - isect.p;
        float r2 = length_squared(w);
        Spectrum li = I / (4.0f * PI * r2);
        p1 = Interaction::from_point(pos);
        wi = normalize(w); pdf = 1.0f;
        return li;
      }
      case RT_LIGHT_DISTANT: {                                   // distant.rs:58-71
        V3 p_outside = isect.p + dir * (2.0f * w_radius);
        p1 = Interaction::from_point(p_outside);
        wi = dir; pdf = 1.0f;
        return I;
      }
      case RT_LIGHT_AREA: {                                      // diffuse.rs:59-70
        shape->sample_si(isect, u, p1, pdf);
        wi = normalize(p1.p - isect.p);
        return L(p1, -wi);
      }
      default: {                                                 // infinite.rs:143-183
        float map_pdf;
        P2 uv = distribution.sample_continuous(u, map_pdf);
        if (map_pdf == 0.0f) { wi = V3(0, 0, 0); pdf = 0.0f; p1 = Interaction::from_point(V3(0, 0, 0)); return Spectrum(0.0f); }
        float theta = uv.y * PI, phi = uv.x * 2.0f * PI;
        float cos_t = std::cos(theta), sin_t = std::sin(theta), cos_p = std::cos(phi), sin_p = std::sin(phi);
        wi = l2w.vector(V3(sin_t * cos_p, sin_t * sin_p, cos_t));
        pdf = sin_t == 0.0f ? 0.0f : map_pdf / (2.0f * PI * PI * sin_t);
        V3 target = isect.p + wi * (2.0f * w_radius);
        p1 = Interaction::from_point(target);
        return map_lookup(uv, 0.0f);
      }
    }
  }
  float pdf_li(const Interaction& si, V3 w) const {
    if (kind == RT_LIGHT_AREA) return shape->pdf_wi(si, w);      // diffuse.rs:72-74
    if (kind == RT_LIGHT_INFINITE) {                             // infinite.rs:185-198
      V3 wi = w2l.vector(w);
      float theta = spherical_theta(wi), phi = spherical_phi(wi);
      float sin_t = std::sin(theta);
      if (sin_t == 0.0f) return 0.0f;
      return distribution.pdf(P2(phi * INV_PI * 0.5f, theta * INV_PI)) / (2.0f * PI * PI * sin_t);
    }
    return 0.0f;
  }
};

struct Scene {                                                   // scene.rs:22-64
  std::vector<Primitive> prims;
  std::vector<Light> lights;
  std::vector<int> infinite_lights;
  std::vector<rt_material> materials;
  TextureSet textures;
  BVH bvh;
  int intersect(Ray& ray, SurfaceInteraction& si) const { tls_counters().regular_rays++; return bvh.intersect(ray, si); }
  bool intersect_p(const Ray& ray) const { tls_counters().shadow_rays++; return bvh.intersect_p(ray); }
  Bounds3 world_bounds() const { return bvh.world_bounds(); }
  bool unoccluded(const Interaction& p0, const Interaction& p1) const {   // light/mod.rs:52-55
    Ray r = p0.spawn_ray_to_interaction(p1);
    return !intersect_p(r);
  }
  int material_of(const SurfaceInteraction& si) const { return si.material_override != -2 ? si.material_override : prims[si.prim].material; }
  Spectrum le(const SurfaceInteraction& si, V3 w) const {        // interaction.rs:149-154
    int al = prims[si.prim].area_light;
    return al >= 0 ? lights[al].L(si.hit, w) : Spectrum(0.0f);
  }
};

// ---------------------------------------------------------------------------------------
// lightdistrib.rs
struct LightDistribution { virtual ~LightDistribution() {} virtual const Distribution1D* lookup(V3 p) = 0; };
struct UniformLightDistribution : LightDistribution {            // :37-54
  Distribution1D d;
  explicit UniformLightDistribution(const Scene& s) { d = Distribution1D(std::vector<float>(s.lights.size(), 1.0f)); }
  const Distribution1D* lookup(V3) override { return &d; }
};
struct SpatialLightDistribution : LightDistribution {            // :59-296
  // The reference's table is a lock-free hash (atomics, lightdistrib.rs:201-296); here one atomic pointer per voxel (at most 64^3), filled on first use:
  // same values per voxel, and like the reference no lock on the read path (a mutex around every lookup made the 16-thread CPU arm contend on it:
  // samples/s varied 2x between processes with the kernel's share of the CPU time, profiles/r03i_bench.json).
  const Scene* scene; uint32_t n_voxels[3];
  std::vector<std::atomic<const Distribution1D*>> table;
  ~SpatialLightDistribution() override { for (auto& e : table) delete e.load(); }
  SpatialLightDistribution(const Scene* s, uint32_t max_voxels) : scene(s) {   // :67-99
    Bounds3 b = s->world_bounds();
    V3 diag = b.diagonal();
    float b_max = diag[b.maximum_extent()];
    for (int i = 0; i < 3; i++) n_voxels[i] = std::max<uint32_t>(1u, f2u32(std::round(diag[i] / b_max * (float)max_voxels)));
    table = std::vector<std::atomic<const Distribution1D*>>((size_t)n_voxels[0] * n_voxels[1] * n_voxels[2]);
    for (auto& e : table) e.store(nullptr, std::memory_order_relaxed);
  }
  Distribution1D compute_distribution(const int32_t pi[3]) const {   // :101-179
    Bounds3 wb = scene->world_bounds();
    V3 p0((float)pi[0] / (float)n_voxels[0], (float)pi[1] / (float)n_voxels[1], (float)pi[2] / (float)n_voxels[2]);
    V3 p1(((float)pi[0] + 1.0f) / (float)n_voxels[0], ((float)pi[1] + 1.0f) / (float)n_voxels[1], ((float)pi[2] + 1.0f) / (float)n_voxels[2]);
    Bounds3 vb = Bounds3::from_points(wb.lerp(p0), wb.lerp(p1));
    const uint64_t n_samples = 128;
    std::vector<float> contrib(scene->lights.size(), 0.0f);
    for (uint64_t i = 0; i < n_samples; i++) {
      V3 po = vb.lerp(V3(radical_inverse(0, i), radical_inverse(1, i), radical_inverse(2, i)));
      Interaction intr = Interaction::make(po, V3(0, 0, 0), V3(1, 0, 0), V3(0, 0, 0));
      P2 u(radical_inverse(3, i), radical_inverse(4, i));
      for (size_t j = 0; j < scene->lights.size(); j++) {
        V3 wi; float pdf; Interaction p1i;
        Spectrum li = scene->lights[j].sample_li(intr, u, wi, pdf, p1i);
        if (pdf > 0.0f) contrib[j] += li.y() / pdf;
      }
    }
    float sum = 0.0f; for (float c : contrib) sum += c;
    float avg = sum / (float)(n_samples * (uint64_t)contrib.size());
    float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (float& c : contrib) c = fmax_(c, min_contrib);
    return Distribution1D(contrib);
  }
  void voxel_of(V3 p, int32_t pi[3]) const {                     // :185-198
    V3 offset = scene->world_bounds().offset(p);
    for (int i = 0; i < 3; i++) pi[i] = clampv<int32_t>(f2i32(offset[i] * (float)n_voxels[i]), 0, (int32_t)n_voxels[i] - 1);
  }
  const Distribution1D* lookup(V3 p) override {
    int32_t pi[3]; voxel_of(p, pi);
    std::atomic<const Distribution1D*>& slot = table[((size_t)pi[0] * n_voxels[1] + (size_t)pi[1]) * n_voxels[2] + (size_t)pi[2]];
    const Distribution1D* d = slot.load(std::memory_order_acquire);
    if (d) return d;
    const Distribution1D* mine = new Distribution1D(compute_distribution(pi));   // two threads may both compute a new voxel: same values, one copy kept
    const Distribution1D* expected = nullptr;
    if (slot.compare_exchange_strong(expected, mine, std::memory_order_acq_rel, std::memory_order_acquire)) return mine;
    delete mine;
    return expected;
  }
};

// ---------------------------------------------------------------------------------------
// filter/*.rs + film.rs
struct Bounds2i { int x0, y0, x1, y1; int area() const { return (x1 - x0) * (y1 - y0); } };   // bounds.rs Bounds2i (area of degenerate: product as-is)
inline float filter_eval(const rt_film& f, float x, float y) {
  switch (f.filter) {
    case RT_FILTER_BOX: return 1.0f;                                                          // boxfilter.rs:26-28
    case RT_FILTER_GAUSSIAN: {                                                                // gaussian.rs:15-42
      float alpha = f.filter_a;
      float expx = std::exp(-alpha * f.filter_xw * f.filter_xw), expy = std::exp(-alpha * f.filter_yw * f.filter_yw);
      auto g = [&](float d, float e) { return fmax_(std::exp(-alpha * d * d) - e, 0.0f); };
      return g(x, expx) * g(y, expy);
    }
    case RT_FILTER_TRIANGLE: return fmax_(0.0f, f.filter_xw - std::fabs(x)) * fmax_(0.0f, f.filter_yw - std::fabs(y));   // triangle.rs:27-29
    default: {                                                                                // mitchell.rs:25-62
      float B = f.filter_a, C = f.filter_b;
      auto m1 = [&](float v) {
        float fx = std::fabs(v) * 2.0f;
        if (fx < 1.0f) return ((12.0f - 9.0f * B - 6.0f * C) * fx * fx * fx + (-18.0f + 12.0f * B + 6.0f * C) * fx * fx + (6.0f - 2.0f * B)) * (1.0f / 6.0f);
        if (fx < 2.0f) return ((-B - 6.0f * C) * fx * fx * fx + (6.0f * B + 30.0f * C) * fx * fx + (-12.0f * B - 48.0f * C) * fx + (8.0f * B + 24.0f * C)) * (1.0f / 6.0f);
        return 0.0f;
      };
      return m1(x * (1.0f / f.filter_xw)) * m1(y * (1.0f / f.filter_yw));
    }
  }
}

struct Film {
  rt_film desc; int xres, yres; Bounds2i cropped;
  float table[256]; float rx, ry;
  std::vector<float> pixels;       // X,Y,Z,weight per cropped pixel (film.rs:38-43)
  std::mutex mu;
  void init(const rt_film& f) {                                   // film.rs:57-115
    desc = f; xres = f.xres; yres = f.yres;
    int ax = f2i32(std::ceil((float)xres * f.crop[0])), ay = f2i32(std::ceil((float)yres * f.crop[2]));
    int bx = f2i32(std::ceil((float)xres * f.crop[1])), by = f2i32(std::ceil((float)yres * f.crop[3]));
    cropped = Bounds2i{std::min(ax, bx), std::min(ay, by), std::max(ax, bx), std::max(ay, by)};
    pixels.assign((size_t)std::max(0, cropped.area()) * 4, 0.0f);
    rx = f.filter_xw; ry = f.filter_yw;
    for (int y = 0; y < 16; y++) {
      float fy = ((float)y + 0.5f) * (ry / 16.0f);
      for (int x = 0; x < 16; x++) { float fx = ((float)x + 0.5f) * (rx / 16.0f); table[y * 16 + x] = filter_eval(f, fx, fy); }
    }
  }
  Bounds2i get_sample_bounds() const {                            // film.rs:249-257
    Bounds2i b;
    b.x0 = f2i32(std::floor((float)cropped.x0 + 0.5f - rx)); b.y0 = f2i32(std::floor((float)cropped.y0 + 0.5f - ry));
    b.x1 = f2i32(std::ceil((float)cropped.x1 - 0.5f + rx)); b.y1 = f2i32(std::ceil((float)cropped.y1 - 0.5f + ry));
    return b;
  }
};

struct FilmTile {                                                 // film.rs:269-376
  Bounds2i pb; float rx, ry, irx, iry; const float* table; float max_lum;
  std::vector<float> px;          // r,g,b,weight
  FilmTile(const Film& film, const Bounds2i& sb) {                // get_film_tile film.rs:152-175
    float p0x = std::ceil((float)sb.x0 - 0.5f - film.rx), p0y = std::ceil((float)sb.y0 - 0.5f - film.ry);
    float p1x = std::floor((float)sb.x1 - 0.5f + film.rx + 1.0f), p1y = std::floor((float)sb.y1 - 0.5f + film.ry + 1.0f);
    // Bounds2f::from_points + intersect with cropped bounds
    float ax = pmin(p0x, p1x), bx = pmax(p0x, p1x), ay = pmin(p0y, p1y), by = pmax(p0y, p1y);
    pb.x0 = f2i32(pmax(ax, (float)film.cropped.x0)); pb.y0 = f2i32(pmax(ay, (float)film.cropped.y0));
    pb.x1 = f2i32(pmin(bx, (float)film.cropped.x1)); pb.y1 = f2i32(pmin(by, (float)film.cropped.y1));
    rx = film.rx; ry = film.ry; irx = 1.0f / rx; iry = 1.0f / ry; table = film.table; max_lum = film.desc.max_sample_luminance;
    px.assign((size_t)std::max(0, pb.area()) * 4, 0.0f);
  }
  void add_sample(P2 p_film, Spectrum colour) {                   // film.rs:298-361
    if (colour.has_nan()) return;
    Spectrum L = colour.y() > max_lum ? colour * max_lum / colour.y() : colour;
    float dx = p_film.x - 0.5f, dy = p_film.y - 0.5f;
    float p0x = std::ceil(dx - rx), p0y = std::ceil(dy - ry);
    float p1x = std::floor(dx + rx + 1.0f), p1y = std::floor(dy + ry + 1.0f);
    float ax = pmin(p0x, p1x), bx = pmax(p0x, p1x), ay = pmin(p0y, p1y), by = pmax(p0y, p1y);
    int x0 = f2i32(pmax(ax, (float)pb.x0)), y0 = f2i32(pmax(ay, (float)pb.y0));
    int x1 = f2i32(pmin(bx, (float)pb.x1)), y1 = f2i32(pmin(by, (float)pb.y1));
    int w = pb.x1 - pb.x0;
    for (int y = y0; y < y1; y++) {
      float fy = std::fabs(((float)y - dy) * iry * 16.0f);
      int iy = (int)f2usize(fmin_(std::floor(fy), 15.0f));
      for (int x = x0; x < x1; x++) {
        float fx = std::fabs(((float)x - dx) * irx * 16.0f);
        int ix = (int)f2usize(fmin_(std::floor(fx), 15.0f));
        float wgt = table[iy * 16 + ix];
        float* p = &px[((size_t)(y - pb.y0) * w + (x - pb.x0)) * 4];
        Spectrum c = L * wgt;
        p[0] += c.r; p[1] += c.g; p[2] += c.b; p[3] += wgt;
      }
    }
  }
};
inline void merge_film_tile(Film& film, const FilmTile& t) {      // film.rs:177-194
  std::lock_guard<std::mutex> g(film.mu);
  int w = film.cropped.x1 - film.cropped.x0, tw = t.pb.x1 - t.pb.x0;
  for (int y = t.pb.y0; y < t.pb.y1; y++) for (int x = t.pb.x0; x < t.pb.x1; x++) {
    const float* tp = &t.px[((size_t)(y - t.pb.y0) * tw + (x - t.pb.x0)) * 4];
    float xyz[3]; to_xyz(Spectrum(tp[0], tp[1], tp[2]), xyz);
    float* fp = &film.pixels[((size_t)(y - film.cropped.y0) * w + (x - film.cropped.x0)) * 4];
    fp[0] += xyz[0]; fp[1] += xyz[1]; fp[2] += xyz[2]; fp[3] += tp[3];
  }
}
// film.rs:196-234: XYZ -> RGB, / weight, clamp >= 0, * scale.  rgb: 3 floats per cropped pixel.
inline void film_resolve(const Film& film, float* rgb) {
  size_t n = (size_t)std::max(0, film.cropped.area());
  for (size_t i = 0; i < n; i++) {
    const float* p = &film.pixels[i * 4];
    Spectrum c = from_xyz(p);
    float w = p[3];
    if (w != 0.0f) { float inv = 1.0f / w; c.r = fmax_(0.0f, c.r * inv); c.g = fmax_(0.0f, c.g * inv); c.b = fmax_(0.0f, c.b * inv); }
    // splat contribution is zero on this path (:222-230: += 1.0 * from_xyz(0))
    float z[3] = {0, 0, 0}; Spectrum sp = from_xyz(z);
    c.r += 1.0f * sp.r; c.g += 1.0f * sp.g; c.b += 1.0f * sp.b;
    rgb[i * 3 + 0] = c.r * film.desc.scale; rgb[i * 3 + 1] = c.g * film.desc.scale; rgb[i * 3 + 2] = c.b * film.desc.scale;
  }
}

// ---------------------------------------------------------------------------------------
// camera.rs
struct Camera {
  Transform c2w, r2c; float lens_radius, focal_distance;
  void init(const rt_camera& c, int xres, int yres) {             // :29-72
    c2w = Transform(Matrix4::from(c.c2w.m), Matrix4::from(c.c2w.m_inv));
    lens_radius = c.lens_radius; focal_distance = c.focal_distance;
    Transform c2s = Transform::perspective(c.fov, 1e-2f, 1000.0f);
    const float* sw = c.screen_window;   // xmin, xmax, ymin, ymax
    Transform s2r = Transform::mulT(Transform::mulT(Transform::scale((float)xres, (float)yres, 1.0f),
                                                    Transform::scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f)),
                                    Transform::translate(V3(-sw[0], -sw[3], 0.0f)));
    Transform r2s = s2r.inverse();
    r2c = Transform::mulT(c2s.inverse(), r2s);
    init_differentials();
  }
  Ray generate_ray(const CameraSample& s) const {                 // :131-148 == differential-free part of :150-202
    V3 p_camera = r2c.point(V3(s.p_film.x, s.p_film.y, 0.0f));
    Ray ray(V3(0, 0, 0), normalize(p_camera));
    if (lens_radius > 0.0f) {
      P2 d = concentric_sample_disk(s.p_lens);
      P2 p_lens(lens_radius * d.x, lens_radius * d.y);
      float ft = focal_distance / ray.d.z;
      V3 p_focus = ray.at(ft);
      ray.o = V3(p_lens.x, p_lens.y, 0.0f);
      ray.d = normalize(p_focus - ray.o);
    }
    V3 oe, de;
    return ray_transform(ray, c2w, oe, de);
  }
  V3 dx_camera, dy_camera;                                        // :62-65
  void init_differentials() {
    dx_camera = r2c.point(V3(1, 0, 0)) - r2c.point(V3(0, 0, 0));
    dy_camera = r2c.point(V3(0, 1, 0)) - r2c.point(V3(0, 0, 0));
  }
  Ray generate_ray_differential(const CameraSample& s) const {    // :150-202
    V3 p_camera = r2c.point(V3(s.p_film.x, s.p_film.y, 0.0f));
    Ray ray(V3(0, 0, 0), normalize(p_camera));
    if (lens_radius > 0.0f) {
      P2 d = concentric_sample_disk(s.p_lens);
      P2 p_lens(lens_radius * d.x, lens_radius * d.y);
      float ft = focal_distance / ray.d.z;
      V3 p_focus = ray.at(ft);
      ray.o = V3(p_lens.x, p_lens.y, 0.0f);
      ray.d = normalize(p_focus - ray.o);
    }
    ray.has_diff = true;
    if (lens_radius > 0.0f) {
      P2 d = concentric_sample_disk(s.p_lens);
      P2 p_lens(lens_radius * d.x, lens_radius * d.y);
      V3 origin(p_lens.x, p_lens.y, 0.0f);
      V3 dx = normalize(p_camera + dx_camera);
      float ft_x = focal_distance / dx.z;
      V3 p_focus_x = ft_x * dx;
      V3 dy = normalize(p_camera + dy_camera);
      float ft_y = focal_distance / dy.z;
      V3 p_focus_y = ft_y * dy;
      ray.rx_o = origin; ray.ry_o = origin;
      ray.rx_d = normalize(p_focus_x - origin); ray.ry_d = normalize(p_focus_y - origin);
    } else {
      ray.rx_o = ray.o; ray.ry_o = ray.o;
      ray.rx_d = normalize(p_camera + dx_camera); ray.ry_d = normalize(p_camera + dy_camera);
    }
    V3 oe, de;
    return ray_transform(ray, c2w, oe, de);
  }
};

// ---------------------------------------------------------------------------------------
// integrator/*.rs
struct Integrator {
  rt_integrator desc;
  std::unique_ptr<LightDistribution> light_distribution;
  std::vector<size_t> n_light_samples;

  void preprocess(const Scene& scene, Sampler& sampler) {
    if (desc.type == RT_INTEGRATOR_PATH) {                        // path.rs:86-94
      if (desc.light_strategy == RT_LIGHTSTRATEGY_UNIFORM || scene.lights.size() == 1) light_distribution.reset(new UniformLightDistribution(scene));
      else light_distribution.reset(new SpatialLightDistribution(&scene, 64));
    } else if (desc.type == RT_INTEGRATOR_DIRECT && desc.direct_strategy == RT_DIRECT_ALL) {   // directlighting.rs:70-87
      for (const Light& l : scene.lights) n_light_samples.push_back(sampler.round_count((size_t)l.n_samples));
      for (int i = 0; i < (int)(uint8_t)desc.max_depth; i++)
        for (size_t j = 0; j < scene.lights.size(); j++) { sampler.request_2d_array(n_light_samples[j]); sampler.request_2d_array(n_light_samples[j]); }
    }
  }

  // integrator/mod.rs:222-318
  static Spectrum estimate_direct(const SurfaceInteraction& it, const Bsdf& bsdf, P2 u_scattering, const Light& light, P2 u_light, const Scene& scene) {
    const uint32_t bsdf_flags = BSDF_ALL & ~BSDF_SPECULAR;
    Spectrum ld(0.0f);
    V3 wi; float light_pdf; Interaction p1;
    Spectrum li = light.sample_li(it.hit, u_light, wi, light_pdf, p1);
    if (light_pdf > 0.0f && !li.is_black()) {
      Spectrum f = bsdf.f(it.hit.wo, wi, bsdf_flags) * std::fabs(dot(wi, it.shading.n));
      float scattering_pdf = bsdf.pdf(it.hit.wo, wi, bsdf_flags);
      if (!f.is_black()) {
        if (!scene.unoccluded(it.hit, p1)) li = Spectrum(0.0f);
        if (!li.is_black()) {
          if (light.is_delta()) ld += f * li / light_pdf;
          else { float weight = power_heuristic(1, light_pdf, 1, scattering_pdf); ld += f * li * weight / light_pdf; }
        }
      }
    }
    if (!light.is_delta()) {
      Spectrum f; V3 wi2; float scattering_pdf; uint32_t sampled;
      bsdf.sample_f(it.hit.wo, u_scattering, bsdf_flags, f, wi2, scattering_pdf, sampled);
      f = f * std::fabs(dot(wi2, it.shading.n));
      bool sampled_specular = (sampled & BSDF_SPECULAR) != 0;
      if (!f.is_black() && scattering_pdf > 0.0f) {
        float weight = 1.0f;
        if (!sampled_specular) {
          float lp = light.pdf_li(it.hit, wi2);
          if (lp == 0.0f) return ld;
          weight = power_heuristic(1, scattering_pdf, 1, lp);
        }
        Ray ray = it.hit.spawn_ray(wi2);
        SurfaceInteraction li_isect;
        Spectrum li2;
        if (scene.intersect(ray, li_isect) >= 0) {
          int al = scene.prims[li_isect.prim].area_light;
          li2 = (al >= 0 && scene.lights[al].id == light.id) ? scene.le(li_isect, -wi2) : Spectrum(0.0f);
        } else li2 = light.le(ray);
        if (!li2.is_black()) ld += f * li2 * weight / scattering_pdf;
      }
    }
    return ld;
  }
  // integrator/mod.rs:186-220
  static Spectrum uniform_sample_one_light(const SurfaceInteraction& it, const Bsdf& bsdf, const Scene& scene, Sampler& sampler, const Distribution1D* distrib) {
    size_t n_lights = scene.lights.size();
    if (n_lights == 0) return Spectrum(0.0f);
    float s = sampler.get_1d();
    size_t light_num; float light_pdf;
    if (distrib) light_num = distrib->sample_discrete(s, light_pdf);
    else { light_num = (size_t)pmin<int64_t>((int64_t)n_lights - 1, f2usize(s * (float)n_lights)); light_pdf = 1.0f / (float)n_lights; }
    if (light_pdf == 0.0f) return Spectrum(0.0f);
    P2 u_light = sampler.get_2d();
    P2 u_scattering = sampler.get_2d();
    return estimate_direct(it, bsdf, u_scattering, scene.lights[light_num], u_light, scene) / light_pdf;
  }
  // integrator/mod.rs:145-184
  Spectrum uniform_sample_all_light(const SurfaceInteraction& it, const Bsdf& bsdf, const Scene& scene, Sampler& sampler) const {
    Spectrum L(0.0f);
    for (size_t j = 0; j < scene.lights.size(); j++) {
      size_t n = n_light_samples[j];
      std::vector<P2> ula, usa; bool have_l = false, have_s = false;
      if (const P2* a = sampler.get_2d_array(n)) { ula.assign(a, a + n); have_l = true; }
      if (const P2* a = sampler.get_2d_array(n)) { usa.assign(a, a + n); have_s = true; }
      if (have_l && have_s) {
        Spectrum Ld(0.0f);
        for (size_t k = 0; k < n; k++) Ld += estimate_direct(it, bsdf, usa[k], scene.lights[j], ula[k], scene);
        L += Ld / (float)n;
      } else {
        P2 u_light = sampler.get_2d();
        P2 u_scattering = sampler.get_2d();
        L += estimate_direct(it, bsdf, u_scattering, scene.lights[j], u_light, scene);
      }
    }
    return L;
  }

  Spectrum li(const Scene& scene, Ray ray, Sampler& sampler, uint32_t depth, uint32_t node = 1) const {
    switch (desc.type) {
      case RT_INTEGRATOR_PATH: return li_path(scene, ray, sampler);
      case RT_INTEGRATOR_WHITTED: case RT_INTEGRATOR_DIRECT: return li_recursive(scene, ray, sampler, depth, node);
      case RT_INTEGRATOR_AO: return li_ao(scene, ray, sampler);
      default: return li_normal(scene, ray);
    }
  }

  // path.rs:96-215
  Spectrum li_path(const Scene& scene, Ray ray, Sampler& sampler) const {
    Spectrum l(0.0f), beta(1.0f);
    bool specular_bounce = false;
    uint8_t bounces = 0;
    const uint8_t max_ray_depth = (uint8_t)desc.max_depth;      // `max_ray_depth as u8` (:42)
    float eta_scale = 1.0f;
    while (true) {
      SurfaceInteraction isect;
      bool found = scene.intersect(ray, isect) >= 0;
      if (bounces == 0 || specular_bounce) {
        if (found) l += beta * scene.le(isect, -ray.d);
        else for (int li : scene.infinite_lights) l += beta * scene.lights[li].le(ray);
      }
      if (!found || bounces >= max_ray_depth) break;
      Bsdf bsdf;
      int mat = scene.material_of(isect);
      isect.compute_differential(ray);                           // interaction.rs:199 (compute_scattering_functions starts with it)
      if (mat < 0 || !compute_scattering_functions(scene.materials.data(), scene.materials[mat], isect, true, bsdf, &scene.textures)) {
        ray = isect.hit.spawn_ray(ray.d);
        bounces -= 1;                                            // u8 wrap (Q23)
        continue;
      }
      const Distribution1D* distrib = light_distribution->lookup(isect.hit.p);
      if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
        Spectrum ld = beta * uniform_sample_one_light(isect, bsdf, scene, sampler, distrib);
        l += ld;
      }
      V3 wo = -ray.d;
      Spectrum f; V3 wi; float pdf; uint32_t flags;
      bsdf.sample_f(wo, sampler.get_2d(), BSDF_ALL, f, wi, pdf, flags);
      if (f.is_black() || pdf <= 0.0f) break;
      beta = beta * f * std::fabs(dot(wi, isect.shading.n)) / pdf;
      specular_bounce = (flags & BSDF_SPECULAR) != 0;
      if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
        float eta = bsdf.eta;
        eta_scale *= dot(wo, isect.hit.n) > 0.0f ? eta * eta : 1.0f / (eta * eta);
      }
      ray = isect.hit.spawn_ray(wi);
      Spectrum rr_beta = beta * eta_scale;
      if (rr_beta.max_component_value() < desc.rr_threshold && bounces > 3) {
        float q = fmax_(1.0f - rr_beta.max_component_value(), 0.05f);
        if (sampler.get_1d() < q) break;
        beta = beta / (1.0f - q);
      }
      bounces += 1;
    }
    return l;
  }

  // whitted.rs:41-99, directlighting.rs:89-143, integrator/mod.rs:49-142 (ray differentials dropped)
  Spectrum li_recursive(const Scene& scene, Ray ray, Sampler& sampler, uint32_t depth, uint32_t node) const {
    Spectrum colour(0.0f);
    SurfaceInteraction isect;
    if (scene.intersect(ray, isect) >= 0) {
      V3 n = isect.shading.n, wo = isect.hit.wo;
      Bsdf bsdf;
      int mat = scene.material_of(isect);
      isect.compute_differential(ray);
      if (mat < 0 || !compute_scattering_functions(scene.materials.data(), scene.materials[mat], isect, false, bsdf, &scene.textures)) {
        Ray r = isect.hit.spawn_ray(ray.d);
        return li_recursive(scene, r, sampler, depth, node);
      }
      Sampler::Saved saved = sampler.enter_node(node);
      colour += scene.le(isect, wo);
      if (desc.type == RT_INTEGRATOR_WHITTED) {
        for (const Light& light : scene.lights) {
          V3 wi; float pdf; Interaction p1;
          Spectrum li = light.sample_li(isect.hit, sampler.get_2d(), wi, pdf, p1);
          if (li.is_black() || pdf == 0.0f) continue;
          Spectrum f = bsdf.f(wo, wi, BSDF_ALL);
          if (!f.is_black() && scene.unoccluded(isect.hit, p1)) colour += f * li * std::fabs(dot(wi, n)) / pdf;
        }
      } else if (!scene.lights.empty()) {
        if (desc.direct_strategy == RT_DIRECT_ALL) colour += uniform_sample_all_light(isect, bsdf, scene, sampler);
        else colour += uniform_sample_one_light(isect, bsdf, scene, sampler, nullptr);
      }
      if (depth + 1 < (uint32_t)(uint8_t)desc.max_depth) {
        for (int pass = 0; pass < 2; pass++) {                   // specular_reflection then specular_transmission
          uint32_t flags = (pass == 0 ? BSDF_REFLECTION : BSDF_TRANSMISSION) | BSDF_SPECULAR;
          Spectrum f; V3 wi; float pdf; uint32_t st;
          bsdf.sample_f(isect.hit.wo, sampler.get_2d(), flags, f, wi, pdf, st);
          V3 ns = isect.shading.n;
          if (pdf > 0.0f && !f.is_black() && std::fabs(dot(wi, ns)) != 0.0f) {
            Ray r = isect.hit.spawn_ray(wi);
            if (ray.has_diff) {                                  // integrator/mod.rs:64-83 (reflection), :107-136 (transmission)
              const V3 kZeroN(0, 0, 0);                          // shading.dndu / dndv / isect.dndv (orc_shapes.hpp)
              r.has_diff = true;
              r.rx_o = isect.hit.p + isect.dpdx; r.ry_o = isect.hit.p + isect.dpdy;
              V3 dndx = kZeroN * isect.dudx + kZeroN * isect.dvdx, dndy = kZeroN * isect.dudy + kZeroN * isect.dvdy;
              V3 dwodx = -ray.rx_d - isect.hit.wo, dwody = -ray.ry_d - isect.hit.wo;
              float dDNdx = dot(dwodx, ns) + dot(isect.hit.wo, dndx), dDNdy = dot(dwody, ns) + dot(isect.hit.wo, dndy);
              if (pass == 0) {
                r.rx_d = wi - dwodx + 2.0f * (dot(isect.hit.wo, ns) * dndx + dDNdx * ns);
                r.ry_d = wi - dwody + 2.0f * (dot(isect.hit.wo, ns) * dndy + dDNdy * ns);
              } else {
                float eta = bsdf.eta;
                V3 w = -isect.hit.wo;
                if (dot(isect.hit.wo, ns) < 0.0f) eta = 1.0f / eta;
                float mu = eta * dot(w, ns) - dot(wi, ns);
                r.rx_d = wi + eta * dwodx - (mu * dndx + dDNdx * ns);
                r.ry_d = wi + eta * dwody - (mu * dndy + dDNdy * ns);
              }
            }
            Spectrum sub = li_recursive(scene, r, sampler, depth + 1, node * 2 + (uint32_t)pass);
            colour += f * sub * std::fabs(dot(wi, ns)) / pdf;
          }
        }
      }
      sampler.leave_node(saved);
    } else {
      for (const Light& l : scene.lights) colour = colour + l.le(ray);
    }
    return colour;
  }

  // ao.rs:32-58
  Spectrum li_ao(const Scene& scene, Ray ray, Sampler& sampler) const {
    size_t n_clear = 0;
    SurfaceInteraction isect;
    if (scene.intersect(ray, isect) >= 0) {
      V3 n = isect.hit.n;
      for (int i = 0; i < desc.ao_samples; i++) {
        P2 s = sampler.get_2d();
        V3 w = uniform_sample_sphere(s);
        if (dot(w, n) < 0.0f) w = -w;
        Ray ao_ray = isect.hit.spawn_ray(w);
        if (!scene.intersect_p(ao_ray)) n_clear += 1;
      }
    }
    return Spectrum((float)n_clear / (float)desc.ao_samples);
  }
  // normal.rs:20-34
  Spectrum li_normal(const Scene& scene, Ray ray) const {
    SurfaceInteraction isect;
    if (scene.intersect(ray, isect) >= 0) return Spectrum(std::fabs(dot(ray.d, isect.hit.n)));
    return Spectrum(0.0f);
  }
};

// ---------------------------------------------------------------------------------------
// renderer.rs:22-143
struct RenderStats { Counters counters; double seconds_tiles = 0, seconds_total = 0; int threads = 0; };

inline Bounds2i integrator_pixel_bounds(const rt_integrator& d, const Film& film) {
  Bounds2i sb = film.get_sample_bounds();
  if (d.type == RT_INTEGRATOR_PATH) {                            // path.rs:53-70
    if (d.has_pixel_bounds) {
      Bounds2i pb{d.pixel_bounds[0], d.pixel_bounds[2], d.pixel_bounds[1], d.pixel_bounds[3]};
      sb = Bounds2i{std::max(sb.x0, pb.x0), std::max(sb.y0, pb.y0), std::min(sb.x1, pb.x1), std::min(sb.y1, pb.y1)};
    }
    return sb;
  }
  if (d.reference_empty_pixel_bounds) return Bounds2i{INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};   // bounds.rs:242-249 (F3)
  return sb;
}

// sampler_kind: 0 = ZeroTwoSequence (reference), 1 = CounterSampler (device twin).
// tile_stride / tile_offset select every k-th tile (bounded CPU-baseline sample; rank partition in the multi-process tests).
inline void render(const Scene& scene, Integrator& integ, const Camera& camera, Film& film, const rt_sampler& sd, int sampler_kind, uint64_t seed,
                   int num_threads, int tile_stride, RenderStats* stats, int tile_offset = 0) {
  auto t_start = std::chrono::steady_clock::now();
  std::unique_ptr<Sampler> proto;
  if (sampler_kind == 0) proto.reset(new ZeroTwoSequence((size_t)sd.spp, (size_t)sd.dimensions));
  else proto.reset(new CounterSampler((size_t)sd.spp, (size_t)sd.dimensions, seed));
  integ.preprocess(scene, *proto);
  Bounds2i sb = film.get_sample_bounds();
  Bounds2i pixel_bounds = integrator_pixel_bounds(integ.desc, film);
  const int bs = 16;
  int ntx = ((sb.x1 - sb.x0) + bs - 1) / bs, nty = ((sb.y1 - sb.y0) + bs - 1) / bs;
  std::atomic<int> next_tile(0);
  int n_tiles = std::max(0, ntx) * std::max(0, nty);
  if (num_threads <= 0) num_threads = (int)std::thread::hardware_concurrency();
  if (num_threads <= 0) num_threads = 1;
  std::mutex stats_mu; Counters total;
  auto t_tiles = std::chrono::steady_clock::now();
  auto worker = [&]() {
    tls_counters() = Counters();
    std::unique_ptr<Sampler> sampler = proto->clone();
    while (true) {
      int ti = next_tile.fetch_add(1);                           // row-major tile order, as the Bounds2i iterator (bounds.rs:387-406)
      if (ti >= n_tiles) break;
      if (tile_stride > 1 && (ti % tile_stride) != tile_offset) continue;
      int tx = ti % ntx, ty = ti / ntx;
      sampler->reseed((uint64_t)(ty * ntx + tx));
      int x0 = sb.x0 + tx * bs, x1 = std::min(x0 + bs, sb.x1), y0 = sb.y0 + ty * bs, y1 = std::min(y0 + bs, sb.y1);
      FilmTile tile(film, Bounds2i{x0, y0, x1, y1});
      for (int y = y0; y < y1; y++) for (int x = x0; x < x1; x++) {
        sampler->start_pixel(x, y);
        if (!(x >= pixel_bounds.x0 && x < pixel_bounds.x1 && y >= pixel_bounds.y0 && y < pixel_bounds.y1)) continue;
        while (true) {
          CameraSample s = sampler->get_camera_sample(x, y);
          Ray ray = camera.generate_ray_differential(s);         // renderer.rs:110-111
          ray.scale_differentials(1.0f / std::sqrt((float)sampler->spp));
          tls_counters().camera_rays++;
          Spectrum c = integ.li(scene, ray, *sampler, 0);
          if (c.has_nan()) c = Spectrum(0.0f);
          if (c.y() < -1e-5f) c = Spectrum(0.0f);
          if (std::isinf(c.y())) c = Spectrum(0.0f);
          tile.add_sample(s.p_film, c);
          if (!sampler->start_next_sample()) break;
        }
      }
      merge_film_tile(film, tile);
    }
    std::lock_guard<std::mutex> g(stats_mu);
    total.add(tls_counters());
  };
  std::vector<std::thread> th;
  for (int i = 0; i < num_threads; i++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  auto t_end = std::chrono::steady_clock::now();
  if (stats) {
    stats->counters = total; stats->threads = num_threads;
    stats->seconds_tiles = std::chrono::duration<double>(t_end - t_tiles).count();
    stats->seconds_total = std::chrono::duration<double>(t_end - t_start).count();
  }
}

}  // namespace orc
