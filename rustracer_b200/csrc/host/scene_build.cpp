// See scene_build.hpp.  Reference line numbers are relative to /root/reference/rustracer-core/src/.
#include "scene_build.hpp"
#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include "../common/material_lobes.hpp"
#include <stdexcept>
#include <thread>

namespace rth {
namespace {

inline void put_bits(float* dst, uint32_t v) { std::memcpy(dst, &v, 4); }
inline int32_t sat_i32(float f) { if (!(f == f)) return 0; if (f <= -2147483648.0f) return INT32_MIN; if (f >= 2147483648.0f) return INT32_MAX; return (int32_t)f; }
inline size_t next_pow2(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }

float roughness_to_alpha(float roughness) {                           // bsdf/microfacet.rs:485-493 (host libm `ln`, as the reference)
  roughness = std::fmax(roughness, 1e-3f);
  float x = std::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

rtgpu_material prep_material(const rt_material& m) {
  rtgpu_material o; std::memset(&o, 0, sizeof(o));
  o.type = (uint32_t)m.type;
  auto clamp0 = [](float v) { return clampf(v, 0.0f, std::numeric_limits<float>::infinity()); };   // Spectrum::clamp (spectrum.rs:156-162)
  for (int i = 0; i < 3; i++) { o.kd[i] = m.kd[i]; o.ks[i] = m.ks[i]; o.kr[i] = m.kr[i]; o.kt[i] = m.kt[i]; o.eta_rgb[i] = m.eta_rgb[i]; o.k_rgb[i] = m.k_rgb[i]; }
  o.eta = m.eta;
  switch (m.type) {
    case RT_MAT_MATTE: {                                              // matte.rs:50-58 ; oren_nayar.rs:17-26 (sigma in degrees, Q32)
      for (int i = 0; i < 3; i++) o.kd[i] = clamp0(m.kd[i]);
      float sigma = clampf(m.sigma, 0.0f, 1.0f);
      o.use_oren_nayar = sigma != 0.0f;
      if (o.use_oren_nayar) {
        float sr = radians(sigma), s2 = sr * sr;
        o.oren_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
        o.oren_b = 0.45f * s2 / (s2 + 0.09f);
      }
      break;
    }
    case RT_MAT_PLASTIC: {                                            // plastic.rs:62-68
      float r = m.roughness;
      if (m.remap_roughness) r = roughness_to_alpha(r);
      o.alpha_u = o.alpha_v = r;
      break;
    }
    case RT_MAT_METAL: {                                              // metal.rs:61-67
      float ur = m.has_uroughness ? m.uroughness : m.roughness, vr = m.has_vroughness ? m.vroughness : m.roughness;
      if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
      o.alpha_u = ur; o.alpha_v = vr;
      break;
    }
    case RT_MAT_GLASS: {                                              // glass.rs:68-76
      float ur = m.uroughness, vr = m.vroughness;
      o.glass_specular = (ur == 0.0f && vr == 0.0f);
      if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
      o.alpha_u = ur; o.alpha_v = vr;
      break;
    }
    case RT_MAT_MIRROR: for (int i = 0; i < 3; i++) o.kr[i] = clamp0(m.kr[i]); break;   // mirror.rs:41
    default: o.type = RTGPU_MAT_NONE; break;
  }
  return o;
}

// ---- lobe lists of UberMaterial / SubstrateMaterial / TranslucentMaterial / MixMaterial ------------------------------
// With constant textures `compute_scattering_functions` adds the same BxDFs at every hit, so the host lists them once
// per material (and per allow_multiple_lobes, which only a glass child of a mix looks at).  The listing itself is
// common/material_lobes.hpp, shared with the device (textured materials are listed there per hit).
float list_lobes(const rt_scene& in, int row, bool allow_multiple_lobes, std::vector<rtgpu_lobe>& out) {
  if (row < 0 || (uint32_t)row >= in.n_materials) throw std::runtime_error("material row out of range");
  rtgpu_lobe buf[rtml::kMaxLobes];
  rtml::LobeList L{buf, 0, rtml::kOk};
  auto children = [&in](int r, bool) -> rt_material {
    if (r < 0 || (uint32_t)r >= in.n_materials) throw std::runtime_error("material row out of range");
    return in.materials[r];
  };
  const float eta = rtml::list_lobes<0>(in.materials[row], allow_multiple_lobes, children, L);
  if (L.error == rtml::kTooManyLobes) throw std::runtime_error("a material builds more than 8 BxDFs: the reference's BxDFHolder panics there (bsdf/mod.rs:41-52)");
  if (L.error == rtml::kMixTooDeep) throw std::runtime_error("MixMaterial nested deeper than two levels is not supported");
  if (L.error == rtml::kNoBsdf) throw std::runtime_error("MixMaterial: a child material has no BSDF");
  out.insert(out.end(), buf, buf + L.n);
  return eta;
}
// Upper bound of the BxDFs a (textured) material can add at a hit, whatever its textures evaluate to.
int max_lobes(const rt_scene& in, int row, int depth = 0) {
  if (row < 0 || (uint32_t)row >= in.n_materials) throw std::runtime_error("material row out of range");
  const rt_material& m = in.materials[row];
  switch (m.type) {
    case RT_MAT_PLASTIC: case RT_MAT_GLASS: return 2;
    case RT_MAT_UBER: return 5;
    case RT_MAT_TRANSLUCENT: return 4;
    case RT_MAT_MIX:
      if (depth >= rtml::kMaxMixDepth) throw std::runtime_error("MixMaterial nested deeper than two levels is not supported");
      return max_lobes(in, m.mix_a, depth + 1) + max_lobes(in, m.mix_b, depth + 1);
    default: return 1;
  }
}

// Static split of [0, n) over the host threads for the per-primitive loops of the flattener (each index writes its own outputs).
template <class F> void parallel_for(size_t n, int threads, F body) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 1 || n < (size_t)1 << 16) { body((size_t)0, n); return; }
  const size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
  std::vector<std::thread> pool;
  for (size_t b = 0; b < n; b += chunk) pool.emplace_back([=, &body]() { body(b, std::min(n, b + chunk)); });
  for (auto& t : pool) t.join();
}

// sampling/distribution1d.rs:11-45
void distribution1d(const float* func, size_t n, float* cdf, float& func_int) {
  cdf[0] = 0.0f;
  for (size_t i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] / (float)n;
  func_int = cdf[n];
  if (func_int == 0.0f) for (size_t i = 1; i < n + 1; i++) cdf[i] = (float)i / (float)n;
  else for (size_t i = 1; i < n + 1; i++) cdf[i] /= func_int;
}

}  // namespace

void make_render_desc(const rt_scene& in, rtgpu_render_desc& rd) {
  std::memset(&rd, 0, sizeof(rd));
  const rt_film& f = in.film; const rt_camera& c = in.camera; const rt_integrator& ig = in.integrator;
  rd.integrator = ig.type; rd.max_depth = ig.max_depth; rd.rr_threshold = ig.rr_threshold;
  rd.light_strategy = ig.light_strategy; rd.direct_strategy = ig.direct_strategy; rd.ao_samples = ig.ao_samples;
  rd.xres = f.xres; rd.yres = f.yres;
  // Film::new (film.rs:66-75)
  int ax = sat_i32(std::ceil((float)f.xres * f.crop[0])), ay = sat_i32(std::ceil((float)f.yres * f.crop[2]));
  int bx = sat_i32(std::ceil((float)f.xres * f.crop[1])), by = sat_i32(std::ceil((float)f.yres * f.crop[3]));
  rd.cropped[0] = std::min(ax, bx); rd.cropped[1] = std::min(ay, by); rd.cropped[2] = std::max(ax, bx); rd.cropped[3] = std::max(ay, by);
  rd.filter_radius[0] = f.filter_xw; rd.filter_radius[1] = f.filter_yw;
  // filter table (film.rs:92-102 ; filter/*.rs)
  auto eval = [&](float x, float y) -> float {
    switch (f.filter) {
      case RT_FILTER_BOX: return 1.0f;
      case RT_FILTER_GAUSSIAN: {
        float a = f.filter_a, ex = std::exp(-a * f.filter_xw * f.filter_xw), ey = std::exp(-a * f.filter_yw * f.filter_yw);
        return std::fmax(std::exp(-a * x * x) - ex, 0.0f) * std::fmax(std::exp(-a * y * y) - ey, 0.0f);
      }
      case RT_FILTER_TRIANGLE: return std::fmax(0.0f, f.filter_xw - std::fabs(x)) * std::fmax(0.0f, f.filter_yw - std::fabs(y));
      default: {
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
  };
  for (int y = 0; y < 16; y++) {
    float fy = ((float)y + 0.5f) * (f.filter_yw / 16.0f);
    for (int x = 0; x < 16; x++) { float fx = ((float)x + 0.5f) * (f.filter_xw / 16.0f); rd.filter_table[y * 16 + x] = eval(fx, fy); }
  }
  // Film::get_sample_bounds (film.rs:249-257)
  rd.sample_bounds[0] = sat_i32(std::floor((float)rd.cropped[0] + 0.5f - f.filter_xw)); rd.sample_bounds[1] = sat_i32(std::floor((float)rd.cropped[1] + 0.5f - f.filter_yw));
  rd.sample_bounds[2] = sat_i32(std::ceil((float)rd.cropped[2] - 0.5f + f.filter_xw)); rd.sample_bounds[3] = sat_i32(std::ceil((float)rd.cropped[3] - 0.5f + f.filter_yw));
  // SamplerIntegrator::pixel_bounds: path.rs:53-70; the other integrators use the sample bounds here
  for (int i = 0; i < 4; i++) rd.pixel_bounds[i] = rd.sample_bounds[i];
  if (ig.type == RT_INTEGRATOR_PATH) {
    if (ig.has_pixel_bounds) {
      rd.pixel_bounds[0] = std::max(rd.sample_bounds[0], ig.pixel_bounds[0]); rd.pixel_bounds[1] = std::max(rd.sample_bounds[1], ig.pixel_bounds[2]);
      rd.pixel_bounds[2] = std::min(rd.sample_bounds[2], ig.pixel_bounds[1]); rd.pixel_bounds[3] = std::min(rd.sample_bounds[3], ig.pixel_bounds[3]);
    }
  } else if (ig.reference_empty_pixel_bounds) {                       // Bounds2i::new (bounds.rs:242-249): nothing is inside
    rd.pixel_bounds[0] = rd.pixel_bounds[1] = INT32_MAX; rd.pixel_bounds[2] = rd.pixel_bounds[3] = INT32_MIN;
  }
  rd.spp = (int32_t)next_pow2((size_t)std::max(1, in.sampler.spp));   // zerotwosequence.rs:32
  rd.sampler_dims = in.sampler.dimensions;
  // PerspectiveCamera::new (camera.rs:29-72)
  Xform c2s = perspective(c.fov, 1e-2f, 1000.0f);
  const float* sw = c.screen_window;
  Xform s2r = compose(compose(scaling((float)f.xres, (float)f.yres, 1.0f), scaling(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f)),
                      translate(v3(-sw[0], -sw[3], 0.0f)));
  Xform r2c = compose(c2s.inverse(), s2r.inverse());
  std::memcpy(rd.raster_to_camera, r2c.m.m, 64);
  std::memcpy(rd.camera_to_world, c.c2w.m, 64);
  rd.lens_radius = c.lens_radius; rd.focal_distance = c.focal_distance;
  rd.max_sample_luminance = f.max_sample_luminance; rd.scale = f.scale;
  rd.tile_rank = 0; rd.tile_world = 1; rd.sample_begin = 0; rd.sample_end = rd.spp;
  rd.seed = 0; rd.clear_film = 1; rd.wave_paths = 0;
}

void flatten_scene(const rt_scene& in, int threads, FlatScene& out, ExternalBvhBuilder external_builder, void* builder_user) {
  // RT_UPLOAD_TIMING=1: wall time of each step on stderr (tools/upload_probe.py)
  const bool timing = std::getenv("RT_UPLOAD_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[flatten] %-27s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  // 1. primitives in Shape-directive order with world-space geometry and bounds
  // Shapes of an object definition (api.rs:951-957) form their own primitive list; an ObjectInstance is one primitive of
  // the scene's list, at the directive's place.
  struct Prim { uint32_t shape; uint32_t local; };                    // local: triangle number | 0
  std::vector<Prim> prims;
  std::vector<Box3> bounds;
  std::vector<std::vector<Prim>> dprims(in.n_objects);
  std::vector<std::vector<Box3>> dbounds(in.n_objects);
  std::vector<std::vector<Vec3>> world_p(in.n_shapes);
  std::vector<int> quadric_of_shape(in.n_shapes, -1);
  bool any_n = false, any_s = false, any_uv = false;
  for (uint32_t si = 0; si < in.n_shapes; si++) {
    const rt_shape& s = in.shapes[si];
    Xform o2w = from_ir(s.o2w);
    if (s.kind == RT_SHAPE_INSTANCE) { prims.push_back(Prim{si, 0}); bounds.push_back(Box3()); continue; }   // bounds: below, once the definitions are built
    std::vector<Prim>& P_ = s.object_def >= 0 ? dprims[s.object_def] : prims;
    std::vector<Box3>& B_ = s.object_def >= 0 ? dbounds[s.object_def] : bounds;
    if (s.kind == RT_SHAPE_TRIMESH) {
      std::vector<Vec3>& wp = world_p[si];
      wp.resize(s.n_vertices);
      parallel_for(s.n_vertices, threads, [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; i++) wp[i] = xf_point(o2w.m, v3(s.P[3 * i], s.P[3 * i + 1], s.P[3 * i + 2]));   // mesh.rs:61
      });
      any_n |= s.N != nullptr; any_s |= s.S != nullptr; any_uv |= s.uv != nullptr;
      const size_t first = P_.size(), nt = s.n_indices / 3;
      P_.resize(first + nt); B_.resize(first + nt);
      parallel_for(nt, threads, [&](size_t t0, size_t t1) {
        for (size_t t = t0; t < t1; t++) {
          Vec3 p0 = wp[s.indices[3 * t]], p1 = wp[s.indices[3 * t + 1]], p2 = wp[s.indices[3 * t + 2]];
          Box3 b = box_of_points(p0, p1); b.grow(p2);                 // mesh.rs:603-608
          P_[first + t] = Prim{si, (uint32_t)t}; B_[first + t] = b;
        }
      });
      out.n_triangles += s.n_indices / 3;
    } else {
      rtgpu_quadric q; std::memset(&q, 0, sizeof(q));
      std::memcpy(q.o2w, o2w.m.m, 64); std::memcpy(q.w2o, o2w.inv.m, 64);
      q.flags = ((s.reverse_orientation != 0) ^ swaps_handedness(o2w.m) ? RTGPU_PRIMFLAG_FLIP : 0) | (s.reverse_orientation ? RTGPU_PRIMFLAG_REVERSE : 0);
      Box3 b;
      if (s.kind == RT_SHAPE_SPHERE) {                                // sphere.rs:30-51
        float r = s.radius;
        q.kind = RTGPU_PRIM_SPHERE; q.radius = r;
        q.z_min = clampf(std::fmin(s.zmin, s.zmax), -r, r); q.z_max = clampf(std::fmax(s.zmin, s.zmax), -r, r);
        q.theta_min = std::acos(clampf(std::fmin(s.zmin, s.zmax) / r, -1.0f, 1.0f));
        q.theta_max = std::acos(clampf(std::fmax(s.zmin, s.zmax) / r, -1.0f, 1.0f));
        q.phi_max = radians(clampf(s.phimax, 0.0f, 360.0f));
        q.area = q.phi_max * r * (q.z_max - q.z_min);                 // :336-338
        for (int k = 0; k < 8; k++)                                   // :205-225
          b.grow(xf_point(o2w.m, v3((k & 1) ? r : -r, (k & 2) ? r : -r, (k & 4) ? q.z_max : q.z_min)));
      } else if (s.kind == RT_SHAPE_DISK) {                           // disk.rs:25-45
        q.kind = RTGPU_PRIM_DISK; q.radius = s.radius; q.height = s.height; q.inner_radius = s.inner_radius;
        q.phi_max = radians(clampf(s.phimax, 0.0f, 360.0f));
        q.area = q.phi_max * 0.5f * (q.radius * q.radius - q.inner_radius * q.inner_radius);   // :156-158
        Vec3 p1 = xf_point(o2w.m, v3(-q.radius, -q.radius, q.height)), p2 = xf_point(o2w.m, v3(q.radius, q.radius, q.height));   // :129-136 (Q14)
        b = box_of_points(v3(std::fmin(p1.x, p2.x), std::fmin(p1.y, p2.y), std::fmin(p1.z, p2.z)), v3(std::fmax(p1.x, p2.x), std::fmax(p1.y, p2.y), std::fmax(p1.z, p2.z)));
      } else {                                                        // cylinder.rs:26-59
        q.kind = RTGPU_PRIM_CYLINDER; q.radius = s.radius; q.z_min = s.zmin; q.z_max = s.zmax;
        q.phi_max = radians(clampf(s.phimax, 0.0f, 360.0f));
        q.area = (q.z_max - q.z_min) * q.radius * q.phi_max;          // :251-253
        Box3 ob = box_of_points(v3(-q.radius, -q.radius, q.z_min), v3(q.radius, q.radius, q.z_max));
        for (int k = 0; k < 8; k++) b.grow(xf_point(o2w.m, v3((k & 1) ? ob.hi.x : ob.lo.x, (k & 2) ? ob.hi.y : ob.lo.y, (k & 4) ? ob.hi.z : ob.lo.z)));
      }
      quadric_of_shape[si] = (int)out.quadrics.size();
      out.quadrics.push_back(q);
      P_.push_back(Prim{si, 0}); B_.push_back(b);
    }
  }
  lap("world vertices + bounds");
  // 2a. object definitions: the aggregate `ObjectInstance` builds when a definition holds more than one primitive
  //     (api.rs:1071-1080, same accelerator parameters), else the primitive itself
  std::vector<FlatBvh> dbvh(in.n_objects);
  std::vector<Box3> def_bounds(in.n_objects);
  for (uint32_t d = 0; d < in.n_objects; d++) {
    if (dprims[d].empty()) continue;
    if (dprims[d].size() > 1) {
      build_bvh(dbounds[d], in.accel.max_node_prims, in.accel.split_method, threads, dbvh[d]);
      def_bounds[d].lo = v3(dbvh[d].node_lo[0], dbvh[d].node_lo[1], dbvh[d].node_lo[2]);
      def_bounds[d].hi = v3(dbvh[d].node_hi[0], dbvh[d].node_hi[1], dbvh[d].node_hi[2]);
    } else def_bounds[d] = dbounds[d][0];
  }
  for (size_t pn = 0; pn < prims.size(); pn++) {                      // TransformedPrimitive::world_bounds (primitive.rs:86-88, transform.rs:342-389)
    const rt_shape& s = in.shapes[prims[pn].shape];
    if (s.kind != RT_SHAPE_INSTANCE) continue;
    if (s.instance_of < 0 || (uint32_t)s.instance_of >= in.n_objects || dprims[s.instance_of].empty()) throw std::runtime_error("ObjectInstance of an empty or unknown object");
    const float* m = s.o2w.m;
    if (m[12] != 0.0f || m[13] != 0.0f || m[14] != 0.0f || m[15] != 1.0f) throw std::runtime_error("ObjectInstance under a projective transform is not supported");
    Xform p2w = from_ir(s.o2w);
    const Box3& ob = def_bounds[s.instance_of];
    Box3 b;
    for (int k = 0; k < 8; k++) b.grow(xf_point(p2w.m, v3((k & 1) ? ob.hi.x : ob.lo.x, (k & 2) ? ob.hi.y : ob.lo.y, (k & 4) ? ob.hi.z : ob.lo.z)));
    bounds[pn] = b;
  }
  // 2b. same SAH BVH as the reference over the scene's primitive list
  if (external_builder && in.accel.split_method == RT_SPLIT_SAH && !bounds.empty()) {
    const size_t nb = bounds.size();
    uvec<float> pb(nb * 6);
    parallel_for(nb, threads, [&](size_t i0, size_t i1) {
      for (size_t i = i0; i < i1; i++) { pb[i * 6] = bounds[i].lo.x; pb[i * 6 + 1] = bounds[i].lo.y; pb[i * 6 + 2] = bounds[i].lo.z; pb[i * 6 + 3] = bounds[i].hi.x; pb[i * 6 + 4] = bounds[i].hi.y; pb[i * 6 + 5] = bounds[i].hi.z; }
    });
    lap("primitive bounds packed");
    out.bvh = FlatBvh();
    out.bvh.node_lo.resize(nb * 8); out.bvh.node_hi.resize(nb * 8); out.bvh.ordered.resize(nb);
    // first touch of the builder's output pages by all threads: a device-to-host copy into untouched pageable memory faults them in one by one
    // (317 ms for the 680 MB of a 10 M-triangle tree against ~70 ms into resident pages, profiles/r02v_upload_probe_c4.log)
    parallel_for(nb, threads, [&](size_t i0, size_t i1) {
      std::memset(&out.bvh.node_lo[i0 * 8], 0, (i1 - i0) * 8 * sizeof(float)); std::memset(&out.bvh.node_hi[i0 * 8], 0, (i1 - i0) * 8 * sizeof(float));
      std::memset(&out.bvh.ordered[i0], 0, (i1 - i0) * sizeof(uint32_t));
    });
    lap("builder output pages touched");
    uint32_t n_nodes = 0; float ms = 0.0f;
    const int rc = external_builder(builder_user, pb.data(), nb, std::max(0, (int)in.accel.max_node_prims), out.bvh.node_lo.data(), out.bvh.node_hi.data(), out.bvh.ordered.data(), &n_nodes, &ms);
    lap("external builder call");
    if (rc != 0) throw std::runtime_error("external BVH builder failed (code " + std::to_string(rc) + ")");
    out.bvh.n_nodes = n_nodes; out.bvh.node_lo.resize((size_t)n_nodes * 4); out.bvh.node_hi.resize((size_t)n_nodes * 4);
    out.bvh.build_seconds = ms * 1e-3;
    std::mutex tally;
    parallel_for(n_nodes, threads, [&](size_t i0, size_t i1) {
      uint32_t leaves = 0, widest = 0;
      for (size_t i = i0; i < i1; i++) {
        uint32_t meta; std::memcpy(&meta, &out.bvh.node_hi[i * 4 + 3], 4);
        if (meta >> 2) { leaves++; widest = std::max(widest, meta >> 2); }
      }
      std::lock_guard<std::mutex> g(tally);
      out.bvh.n_leaves += leaves; out.bvh.max_leaf_prims = std::max(out.bvh.max_leaf_prims, widest);
    });
  } else build_bvh(bounds, in.accel.max_node_prims, in.accel.split_method, threads, out.bvh);
  lap("BVH build (incl. builder I/O)");
  const size_t N0 = prims.size();
  // 2c. append the definitions' trees and slots to the same arrays with absolute indices
  std::vector<uint32_t> def_root_node(in.n_objects, 0xffffffffu), def_first_slot(in.n_objects, 0), def_first_pn(in.n_objects, 0);
  {
    auto bits = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    uint32_t next_pn = (uint32_t)N0;
    for (uint32_t d = 0; d < in.n_objects; d++) {
      if (dprims[d].empty()) continue;
      const uint32_t base_slot = (uint32_t)out.bvh.ordered.size(), base_node = out.bvh.n_nodes;
      def_first_slot[d] = base_slot; def_first_pn[d] = next_pn;
      if (dprims[d].size() > 1) {
        const FlatBvh& b = dbvh[d];
        def_root_node[d] = base_node;
        for (uint32_t i = 0; i < b.n_nodes; i++) {
          const uint32_t n_prims = bits(b.node_hi[i * 4 + 3]) >> 2;
          for (int k = 0; k < 3; k++) { out.bvh.node_lo.push_back(b.node_lo[i * 4 + k]); out.bvh.node_hi.push_back(b.node_hi[i * 4 + k]); }
          float w; put_bits(&w, bits(b.node_lo[i * 4 + 3]) + (n_prims > 0 ? base_slot : base_node));
          out.bvh.node_lo.push_back(w); out.bvh.node_hi.push_back(b.node_hi[i * 4 + 3]);
        }
        out.bvh.n_nodes += b.n_nodes;
        for (uint32_t local : b.ordered) out.bvh.ordered.push_back(next_pn + local);
      } else out.bvh.ordered.push_back(next_pn);
      next_pn += (uint32_t)dprims[d].size();
    }
  }
  const size_t N = out.bvh.ordered.size();
  // global primitive number -> (list, index): the scene's list first, then the definitions' lists
  auto prim_of = [&](uint32_t pn) -> const Prim& {
    if (pn < N0) return prims[pn];
    for (uint32_t d = in.n_objects; d-- > 0;) if (!dprims[d].empty() && pn >= def_first_pn[d]) return dprims[d][pn - def_first_pn[d]];
    throw std::runtime_error("primitive number out of range");
  };
  // 3. lights: Scene::lights order, area lights one per primitive (api.rs:934-946,963)
  std::vector<uint32_t> first_prim(in.n_shapes + 1, 0);
  for (uint32_t si = 0; si < in.n_shapes; si++)
    first_prim[si + 1] = first_prim[si] + (in.shapes[si].object_def >= 0 ? 0 : (in.shapes[si].kind == RT_SHAPE_TRIMESH ? in.shapes[si].n_indices / 3 : 1));
  out.slot_of_prim.resize(N);                                          // `ordered` is a permutation: every entry is written
  parallel_for(N, threads, [&](size_t s0, size_t s1) { for (size_t slot = s0; slot < s1; slot++) out.slot_of_prim[out.bvh.ordered[slot]] = (uint32_t)slot; });
  lap("definitions + slot_of_prim");
  std::vector<int32_t> light_of_prim(N, -1);
  Vec3 wc = v3(0, 0, 0); float world_radius = 0.0f;
  if (out.bvh.n_nodes > 0) {                                          // Bounds3::bounding_sphere (bounds.rs:197-211)
    Vec3 lo = v3(out.bvh.node_lo[0], out.bvh.node_lo[1], out.bvh.node_lo[2]), hi = v3(out.bvh.node_hi[0], out.bvh.node_hi[1], out.bvh.node_hi[2]);
    wc = v3((lo.x + hi.x) / 2.0f, (lo.y + hi.y) / 2.0f, (lo.z + hi.z) / 2.0f);
    bool inside = wc.x >= lo.x && wc.x <= hi.x && wc.y >= lo.y && wc.y <= hi.y && wc.z >= lo.z && wc.z <= hi.z;
    world_radius = inside ? len(sub(hi, wc)) : 0.0f;
    for (int i = 0; i < 3; i++) { out.desc.world_lo[i] = lo[i]; out.desc.world_hi[i] = hi[i]; }
  }
  for (uint32_t li = 0; li < in.n_lights; li++) {
    const rt_light& l = in.lights[li];
    rtgpu_light g; std::memset(&g, 0, sizeof(g));
    g.kind = (uint32_t)l.kind; g.n_samples = 1; g.world_radius = world_radius;
    for (int i = 0; i < 3; i++) g.I[i] = l.I[i];
    if (l.kind == RT_LIGHT_AREA) {
      const rt_shape& s = in.shapes[l.shape];
      const rt_area_light& al = in.area_lights[s.area_light];
      for (uint32_t k = first_prim[l.shape]; k < first_prim[l.shape + 1]; k++) {
        rtgpu_light a = g;
        for (int i = 0; i < 3; i++) a.I[i] = al.L[i];
        a.two_sided = al.two_sided != 0; a.n_samples = (uint32_t)al.n_samples; a.prim_slot = out.slot_of_prim[k];
        if (s.kind == RT_SHAPE_TRIMESH) {                             // mesh.rs:588-594
          uint32_t t = k - first_prim[l.shape];
          const std::vector<Vec3>& wp = world_p[l.shape];
          Vec3 p0 = wp[s.indices[3 * t]], p1 = wp[s.indices[3 * t + 1]], p2 = wp[s.indices[3 * t + 2]];
          a.area = 0.5f * len(cross(sub(p1, p0), sub(p2, p0)));
        } else a.area = out.quadrics[quadric_of_shape[l.shape]].area;
        light_of_prim[k] = (int32_t)out.lights.size();
        out.lights.push_back(a);
      }
      continue;
    }
    if (l.kind == RT_LIGHT_POINT) for (int i = 0; i < 3; i++) g.pos[i] = l.pos[i];
    else if (l.kind == RT_LIGHT_DISTANT) { Vec3 d = unit(v3(l.dir[0], l.dir[1], l.dir[2])); g.dir[0] = d.x; g.dir[1] = d.y; g.dir[2] = d.z; }   // distant.rs:24-33
    else {                                                            // infinite.rs:46-113
      g.n_samples = (uint32_t)l.n_samples;
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { g.l2w[r * 3 + c] = l.l2w.m[r * 4 + c]; g.w2l[r * 3 + c] = l.l2w.m_inv[r * 4 + c]; }
      // texels * power (infinite.rs:61 / :69), then MIPMap::new(resolution, texels, false, 0.0, Repeat) (:71-77): a map whose sides
      // are not powers of two is Lanczos-resampled, and the pyramid's coarser levels serve the sampling image of very wide maps
      const bool has_map = l.env_w > 0 && l.env_h > 0 && l.env_rgb;
      const int rw = has_map ? l.env_w : 1, rh = has_map ? l.env_h : 1;
      std::vector<float> scaled((size_t)rw * rh * 3);
      for (size_t i = 0; i < (size_t)rw * rh; i++) for (int c = 0; c < 3; c++) scaled[3 * i + c] = (has_map ? l.env_rgb[3 * i + c] : 1.0f) * l.I[c];
      const std::vector<MipLevel> pyr = build_mip_pyramid(rw, rh, 3, RT_WRAP_REPEAT, scaled.data());
      const int w = pyr[0].u, h = pyr[0].v;
      g.env_w = (uint32_t)w; g.env_h = (uint32_t)h;
      std::vector<float>& E = out.env_data;
      g.env_texels = (uint32_t)E.size();
      E.insert(E.end(), pyr[0].d.begin(), pyr[0].d.end());            // the device only ever looks up level 0 (width 0: le, sample_li)
      auto triangle = [&](size_t level, float sx, float sy, float rgb[3]) {   // MIPMap::triangle (mipmap.rs:285-309), Repeat wrap (:201-225)
        level = std::min(level, pyr.size() - 1);
        const MipLevel& L = pyr[level];
        auto texel = [&](int64_t s, int64_t t, int c) {
          int64_t ss = s % L.u; if (ss < 0) ss += L.u;
          int64_t tt = t % L.v; if (tt < 0) tt += L.v;
          return L.d[(size_t)(tt * L.u + ss) * 3 + c];
        };
        const float s = sx * (float)L.u - 0.5f, t = sy * (float)L.v - 0.5f;
        const float fs = std::floor(s), ft = std::floor(t);
        const int64_t s0 = (int64_t)fs, t0 = (int64_t)ft;
        const float ds = s - fs, dt = t - ft;
        for (int c = 0; c < 3; c++)
          rgb[c] = texel(s0, t0, c) * (1.0f - ds) * (1.0f - dt) + texel(s0, t0 + 1, c) * (1.0f - ds) * dt + texel(s0 + 1, t0, c) * ds * (1.0f - dt) + texel(s0 + 1, t0 + 1, c) * ds * dt;
      };
      const int W2 = 2 * w, H2 = 2 * h;
      const float filter = 0.5f / std::min((float)W2, (float)H2);      // infinite.rs:81
      const float n_levels_f = (float)pyr.size();
      std::vector<float> img((size_t)W2 * H2);
      for (int v = 0; v < H2; v++) {
        float vp = ((float)v + 0.5f) / (float)H2;
        float sin_theta = std::sin(kPi * ((float)v + 0.5f) / (float)H2);
        for (int u = 0; u < W2; u++) {
          float up = ((float)u + 0.5f) / (float)W2;
          float rgb[3];
          const float level = n_levels_f - 1.0f + std::log2(std::max(filter, 1e-8f));   // MIPMap::lookup (mipmap.rs:227-245)
          if (level < 0.0f) triangle(0, up, vp, rgb);
          else if (level >= n_levels_f - 1.0f) { for (int c = 0; c < 3; c++) rgb[c] = pyr.back().d[c]; }
          else {
            const float il = std::floor(level), delta = level - il;
            float a[3], b2[3];
            triangle((size_t)il, up, vp, a); triangle((size_t)il + 1, up, vp, b2);
            for (int c = 0; c < 3; c++) rgb[c] = a[c] * (1.0f - delta) + b2[c] * delta;   // lerp (lib.rs:107-117)
          }
          float y = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
          img[(size_t)v * W2 + u] = y * sin_theta;
        }
      }
      // Distribution2D::new (sampling/distribution2d.rs:11-28)
      g.env_func = (uint32_t)E.size(); E.insert(E.end(), img.begin(), img.end());
      g.env_cdf = (uint32_t)E.size(); E.resize(E.size() + (size_t)(W2 + 1) * H2);
      g.env_func_int = (uint32_t)E.size(); E.resize(E.size() + H2);
      std::vector<float> mfunc(H2);
      for (int v = 0; v < H2; v++) {
        float fi; distribution1d(&img[(size_t)v * W2], W2, &E[g.env_cdf + (size_t)v * (W2 + 1)], fi);
        E[g.env_func_int + v] = fi; mfunc[v] = fi;
      }
      g.env_mfunc = (uint32_t)E.size(); E.insert(E.end(), mfunc.begin(), mfunc.end());
      g.env_mcdf = (uint32_t)E.size(); E.resize(E.size() + H2 + 1);
      distribution1d(mfunc.data(), H2, &E[g.env_mcdf], g.env_mfunc_int);
    }
    out.lights.push_back(g);
  }
  // 4. per-slot arrays in ordered_prims order
  // uninitialised (hmath.hpp uvec): the loop below writes every word of every row it owns, starting from zeros
  out.prim_geom.resize(N * 12);
  out.prim_info.resize(N * 4);
  if (any_n) out.tri_n.resize(N * 9);
  if (any_s) out.tri_s.resize(N * 9);
  if (any_uv) out.tri_uv.resize(N * 6);
  lap("lights + output allocation");
  std::vector<uint32_t> mesh_flags(in.n_shapes, 0);                   // per mesh: orientation flags (interaction.rs:119-122)
  bool any_instance = false;
  for (uint32_t si = 0; si < in.n_shapes; si++) {
    const rt_shape& s = in.shapes[si];
    any_instance |= s.kind == RT_SHAPE_INSTANCE;
    if (s.kind != RT_SHAPE_TRIMESH) continue;
    Xform o2w = from_ir(s.o2w);
    if ((s.reverse_orientation != 0) ^ swaps_handedness(o2w.m)) mesh_flags[si] |= RTGPU_PRIMFLAG_FLIP;
    if (s.reverse_orientation) mesh_flags[si] |= RTGPU_PRIMFLAG_REVERSE;
  }
  // every slot writes its own rows; only instance rows are numbered in slot order, so scenes with instances stay on one thread
  parallel_for(N, any_instance ? 1 : threads, [&](size_t slot0, size_t slot1) {
  for (size_t slot = slot0; slot < slot1; slot++) {
    uint32_t pn = out.bvh.ordered[slot];
    const Prim& pr = prim_of(pn);
    const rt_shape& s = in.shapes[pr.shape];
    float* g = &out.prim_geom[slot * 12];
    std::memset(g, 0, 12 * sizeof(float));
    if (any_n) std::memset(&out.tri_n[slot * 9], 0, 9 * sizeof(float));
    if (any_s) std::memset(&out.tri_s[slot * 9], 0, 9 * sizeof(float));
    if (any_uv) std::memset(&out.tri_uv[slot * 6], 0, 6 * sizeof(float));
    uint32_t flags = 0;
    if (s.kind == RT_SHAPE_INSTANCE) {                                // a degenerate triangle for walkers that do not know instances
      const uint32_t d = (uint32_t)s.instance_of;
      rtgpu_instance row; std::memset(&row, 0, sizeof(row));
      for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) { row.o2w[r * 4 + c] = s.o2w.m[r * 4 + c]; row.w2o[r * 4 + c] = s.o2w.m_inv[r * 4 + c]; }
      row.root_node = def_root_node[d]; row.first_slot = def_first_slot[d];
      for (int k = 0; k < 3; k++) { row.lo[k] = def_bounds[d].lo[k]; row.hi[k] = def_bounds[d].hi[k]; }
      row.prim_number = pn;
      put_bits(&g[3], RTGPU_PRIM_TRIANGLE); put_bits(&g[7], 2u); put_bits(&g[11], (uint32_t)out.instances.size());
      out.instances.push_back(row);
    } else if (s.kind == RT_SHAPE_TRIMESH) {
      const int32_t* ix = &s.indices[3 * pr.local];
      for (int v = 0; v < 3; v++) { Vec3 p = world_p[pr.shape][ix[v]]; g[4 * v] = p.x; g[4 * v + 1] = p.y; g[4 * v + 2] = p.z; }
      put_bits(&g[3], RTGPU_PRIM_TRIANGLE);
      flags = mesh_flags[pr.shape];
      if (s.N) { flags |= RTGPU_PRIMFLAG_HAS_N; for (int v = 0; v < 3; v++) for (int c = 0; c < 3; c++) out.tri_n[slot * 9 + v * 3 + c] = s.N[3 * ix[v] + c]; }
      if (s.S) { flags |= RTGPU_PRIMFLAG_HAS_S; for (int v = 0; v < 3; v++) for (int c = 0; c < 3; c++) out.tri_s[slot * 9 + v * 3 + c] = s.S[3 * ix[v] + c]; }
      if (s.uv) { flags |= RTGPU_PRIMFLAG_HAS_UV; for (int v = 0; v < 3; v++) for (int c = 0; c < 2; c++) out.tri_uv[slot * 6 + v * 2 + c] = s.uv[2 * ix[v] + c]; }
    } else {
      const rtgpu_quadric& q = out.quadrics[quadric_of_shape[pr.shape]];
      put_bits(&g[3], q.kind | ((uint32_t)quadric_of_shape[pr.shape] << 2));
      flags = q.flags;
    }
    uint32_t* info = &out.prim_info[slot * 4];
    info[0] = pn; info[1] = s.material >= 0 ? (uint32_t)s.material : 0xffffffffu; info[2] = (uint32_t)light_of_prim[pn]; info[3] = flags;
  }
  });
  lap("geometry + info rows");
  for (uint32_t d = 0; d < in.n_objects; d++)                         // a one-primitive definition has no leaf node to mark its last slot
    if (dprims[d].size() == 1) { uint32_t u; std::memcpy(&u, &out.prim_geom[(size_t)def_first_slot[d] * 12 + 7], 4); put_bits(&out.prim_geom[(size_t)def_first_slot[d] * 12 + 7], u | 1u); }
  bool any_textured = false;
  for (uint32_t i = 0; i < in.n_materials; i++) {
    rtgpu_material pm = prep_material(in.materials[i]);
    const int ty = in.materials[i].type;
    if (in.materials[i].textured) {                                   // evaluated per hit on the device (texture.cuh)
      if (max_lobes(in, (int)i) > rtml::kMaxLobes)
        throw std::runtime_error("a textured material can build more than 8 BxDFs: the reference's BxDFHolder panics there (bsdf/mod.rs:41-52)");
      std::memset(&pm, 0, sizeof(pm));
      pm.type = RTGPU_MAT_TEXTURED;
      any_textured = true;
    } else if (ty == RT_MAT_UBER || ty == RT_MAT_SUBSTRATE || ty == RT_MAT_TRANSLUCENT || ty == RT_MAT_MIX) {
      pm.type = RTGPU_MAT_LOBES;
      for (int allow = 0; allow < 2; allow++) {
        pm.lobe_first[allow] = (uint32_t)out.lobes.size();
        pm.bsdf_eta = list_lobes(in, (int)i, allow != 0, out.lobes);
        pm.lobe_count[allow] = (uint32_t)out.lobes.size() - pm.lobe_first[allow];
        if (pm.lobe_count[allow] > 8)
          throw std::runtime_error("a material builds more than 8 BxDFs: the reference's BxDFHolder panics there (bsdf/mod.rs:41-52)");
      }
    }
    out.materials.push_back(pm);
  }
  if (any_textured) build_textures(in, out.textures, out.tex_data);
  // 5. descriptor views
  rtgpu_scene_desc& d = out.desc;
  d.n_nodes = out.bvh.n_nodes; d.node_lo = out.bvh.node_lo.data(); d.node_hi = out.bvh.node_hi.data();
  d.n_prims = (uint32_t)N; d.prim_geom = out.prim_geom.data(); d.prim_info = out.prim_info.data();
  d.tri_n = any_n ? out.tri_n.data() : nullptr; d.tri_s = any_s ? out.tri_s.data() : nullptr; d.tri_uv = any_uv ? out.tri_uv.data() : nullptr;
  d.n_quadrics = (uint32_t)out.quadrics.size(); d.quadrics = out.quadrics.data();
  d.n_materials = (uint32_t)out.materials.size(); d.materials = out.materials.data();
  d.n_lobes = (uint32_t)out.lobes.size(); d.lobes = out.lobes.empty() ? nullptr : out.lobes.data();
  d.n_texmats = any_textured ? in.n_materials : 0; d.texmats = any_textured ? in.materials : nullptr;
  d.n_textures = (uint32_t)out.textures.size(); d.textures = out.textures.empty() ? nullptr : out.textures.data();
  d.n_tex_floats = (uint32_t)out.tex_data.size(); d.tex_data = out.tex_data.empty() ? nullptr : out.tex_data.data();
  d.n_instances = (uint32_t)out.instances.size(); d.instances = out.instances.empty() ? nullptr : out.instances.data();
  d.n_lights = (uint32_t)out.lights.size(); d.lights = out.lights.data();
  d.n_env_floats = (uint32_t)out.env_data.size(); d.env_data = out.env_data.data();
  make_render_desc(in, out.render);
  lap("materials + descriptor");
}

}  // namespace rth
