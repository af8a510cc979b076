// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement, for tests/ (ctypes),
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
// (rustracer_b200/) never links or loads this library.
#include "orc_render.hpp"
#include <chrono>
#include <functional>
#include <cstdio>
#include <random>
#include <string>

using namespace orc;

struct orc_scene {
  Scene scene;
  Camera camera;
  rt_film film_desc;
  rt_sampler sampler;
  rt_integrator integrator;
  std::vector<std::shared_ptr<TriangleMesh>> meshes;
  std::vector<size_t> shape_first_prim, shape_n_prims;
  double build_seconds = 0;
  std::string error;
  std::vector<std::shared_ptr<ObjectDef>> defs;
};

static Transform to_transform(const rt_transform& t) { return Transform(Matrix4::from(t.m), Matrix4::from(t.m_inv)); }

extern "C" {

orc_scene* orc_scene_create(const rt_scene* in) {
  orc_scene* o = new orc_scene();
  Scene& sc = o->scene;
  o->film_desc = in->film; o->sampler = in->sampler; o->integrator = in->integrator;
  sc.materials.assign(in->materials, in->materials + in->n_materials);
  sc.textures.init(in->textures, in->n_textures);
  // shapes -> primitives in directive order (api.rs:913-966, mesh.rs:636-679); shapes of an object definition go to the
  // definition's own list (api.rs:951-957), an ObjectInstance adds one TransformedPrimitive (api.rs:1081-1085)
  std::vector<std::shared_ptr<ObjectDef>> defs(in->n_objects);
  for (auto& d : defs) d = std::make_shared<ObjectDef>();
  for (uint32_t si = 0; si < in->n_shapes; si++) {
    const rt_shape& s = in->shapes[si];
    Transform o2w = to_transform(s.o2w);
    o->shape_first_prim.push_back(sc.prims.size());
    if (s.kind == RT_SHAPE_INSTANCE) {
      Primitive p; p.material = -1;
      p.shape = std::make_shared<InstanceShape>(defs[s.instance_of], o2w);
      sc.prims.push_back(p);
      o->shape_n_prims.push_back(1);
      continue;
    }
    std::vector<Primitive>& dst = s.object_def >= 0 ? defs[s.object_def]->prims : sc.prims;
    const size_t dst_before = dst.size();
    if (s.kind == RT_SHAPE_TRIMESH) {
      auto mesh = std::make_shared<TriangleMesh>();
      mesh->o2w = o2w;
      mesh->vi.resize(s.n_indices);
      for (uint32_t i = 0; i < s.n_indices; i++) mesh->vi[i] = (int64_t)s.indices[i];
      mesh->p.resize(s.n_vertices);
      for (uint32_t i = 0; i < s.n_vertices; i++) mesh->p[i] = o2w.point(V3(s.P[3 * i], s.P[3 * i + 1], s.P[3 * i + 2]));   // mesh.rs:61
      if (s.N) { mesh->has_n = true; mesh->n.resize(s.n_vertices); for (uint32_t i = 0; i < s.n_vertices; i++) mesh->n[i] = V3(s.N[3 * i], s.N[3 * i + 1], s.N[3 * i + 2]); }
      if (s.S) { mesh->has_s = true; mesh->s.resize(s.n_vertices); for (uint32_t i = 0; i < s.n_vertices; i++) mesh->s[i] = V3(s.S[3 * i], s.S[3 * i + 1], s.S[3 * i + 2]); }
      if (s.uv) { mesh->has_uv = true; mesh->uv.resize(s.n_vertices); for (uint32_t i = 0; i < s.n_vertices; i++) mesh->uv[i] = P2(s.uv[2 * i], s.uv[2 * i + 1]); }
      o->meshes.push_back(mesh);
      size_t ntri = s.n_indices / 3;
      for (size_t t = 0; t < ntri; t++) {
        Primitive p; p.shape = std::make_shared<Triangle>(mesh, t, s.reverse_orientation != 0); p.material = s.material;
        dst.push_back(p);
      }
    } else {
      Primitive p; p.material = s.material;
      if (s.kind == RT_SHAPE_SPHERE) p.shape = std::make_shared<Sphere>(o2w, s.radius, s.zmin, s.zmax, s.phimax, s.reverse_orientation != 0);
      else if (s.kind == RT_SHAPE_DISK) p.shape = std::make_shared<Disk>(s.height, s.radius, s.inner_radius, s.phimax, o2w, s.reverse_orientation != 0);
      else p.shape = std::make_shared<Cylinder>(o2w, s.radius, s.zmin, s.zmax, s.phimax, s.reverse_orientation != 0);
      dst.push_back(p);
    }
    o->shape_n_prims.push_back(s.object_def >= 0 ? 0 : dst.size() - dst_before);
  }
  for (auto& d : defs) if (!d->prims.empty()) d->finish(in->accel.max_node_prims, in->accel.split_method);
  o->defs = defs;
  // lights in creation order; ids = get_next_id order (light/mod.rs:58-64)
  for (uint32_t li = 0; li < in->n_lights; li++) {
    const rt_light& l = in->lights[li];
    if (l.kind == RT_LIGHT_AREA) {
      const rt_shape& s = in->shapes[l.shape];
      const rt_area_light& al = in->area_lights[s.area_light];
      for (size_t k = 0; k < o->shape_n_prims[l.shape]; k++) {
        size_t pn = o->shape_first_prim[l.shape] + k;
        Light L; L.kind = RT_LIGHT_AREA; L.id = (int)sc.lights.size();
        L.shape = sc.prims[pn].shape.get(); L.l_emit = Spectrum(al.L[0], al.L[1], al.L[2]);
        L.two_sided = al.two_sided != 0; L.n_samples = al.n_samples; L.area = L.shape->area();
        sc.prims[pn].area_light = L.id;
        sc.lights.push_back(L);
      }
    } else {
      Light L; L.kind = l.kind; L.id = (int)sc.lights.size();
      L.I = Spectrum(l.I[0], l.I[1], l.I[2]);
      if (l.kind == RT_LIGHT_POINT) L.pos = V3(l.pos[0], l.pos[1], l.pos[2]);
      else if (l.kind == RT_LIGHT_DISTANT) L.dir = normalize(V3(l.dir[0], l.dir[1], l.dir[2]));   // distant.rs:24-33
      else {
        L.l2w = to_transform(l.l2w); L.w2l = L.l2w.inverse(); L.n_samples = l.n_samples;
        if (l.env_w > 0 && l.env_h > 0 && l.env_rgb) {
          L.env_w = l.env_w; L.env_h = l.env_h; L.texels.resize((size_t)l.env_w * l.env_h);
          for (size_t i = 0; i < L.texels.size(); i++) L.texels[i] = Spectrum(l.env_rgb[3 * i], l.env_rgb[3 * i + 1], l.env_rgb[3 * i + 2]) * L.I;   // infinite.rs:61
        } else { L.env_w = 1; L.env_h = 1; L.texels.assign(1, L.I); }
        L.init_infinite();
      }
      sc.lights.push_back(L);
    }
  }
  auto t0 = std::chrono::steady_clock::now();
  sc.bvh.build(sc.prims, in->accel.max_node_prims, in->accel.split_method);
  o->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // Scene::new: light.preprocess + infinite list (scene.rs:29-49)
  if (!sc.bvh.nodes.empty()) {
    for (Light& L : sc.lights) if (L.kind == RT_LIGHT_DISTANT || L.kind == RT_LIGHT_INFINITE) sc.world_bounds().bounding_sphere(L.w_center, L.w_radius);
  }
  for (size_t i = 0; i < sc.lights.size(); i++) if (sc.lights[i].kind == RT_LIGHT_INFINITE) sc.infinite_lights.push_back((int)i);
  o->camera.init(in->camera, in->film.xres, in->film.yres);
  return o;
}
void orc_scene_destroy(orc_scene* s) { delete s; }
const char* orc_scene_error(orc_scene* s) { return s->error.empty() ? nullptr : s->error.c_str(); }
double orc_scene_build_seconds(orc_scene* s) { return s->build_seconds; }

void orc_bvh_info(orc_scene* s, uint64_t* n_nodes, uint64_t* n_prims, uint64_t* n_lights) {
  *n_nodes = s->scene.bvh.nodes.size(); *n_prims = s->scene.prims.size(); *n_lights = s->scene.lights.size();
}
// bounds: 6 floats per node (min xyz, max xyz); meta: 3 ints per node (n_prims [0 = interior], axis, offset); ordered: prim_number per slot
void orc_bvh_export(orc_scene* s, float* bounds, int64_t* meta, int32_t* ordered) {
  const BVH& b = s->scene.bvh;
  for (size_t i = 0; i < b.nodes.size(); i++) {
    const LinearNode& n = b.nodes[i];
    bounds[6 * i + 0] = n.bounds.p_min.x; bounds[6 * i + 1] = n.bounds.p_min.y; bounds[6 * i + 2] = n.bounds.p_min.z;
    bounds[6 * i + 3] = n.bounds.p_max.x; bounds[6 * i + 4] = n.bounds.p_max.y; bounds[6 * i + 5] = n.bounds.p_max.z;
    meta[3 * i + 0] = n.leaf ? (int64_t)n.n_prims : 0; meta[3 * i + 1] = n.axis; meta[3 * i + 2] = (int64_t)n.offset;
  }
  for (size_t i = 0; i < b.ordered.size(); i++) ordered[i] = b.ordered[i];
}
// world-space triangle vertices of prim_number order (9 floats per prim, zeros for quadrics) — to check the host flattening
void orc_prim_world_vertices(orc_scene* s, float* out) {
  for (size_t i = 0; i < s->scene.prims.size(); i++) {
    const Triangle* t = dynamic_cast<const Triangle*>(s->scene.prims[i].shape.get());
    for (int k = 0; k < 9; k++) out[9 * i + k] = 0.0f;
    if (t) for (int v = 0; v < 3; v++) { V3 p = t->mesh->p[t->v(v)]; out[9 * i + 3 * v] = p.x; out[9 * i + 3 * v + 1] = p.y; out[9 * i + 3 * v + 2] = p.z; }
  }
}

// rays: 8 floats each {ox,oy,oz,tmax, dx,dy,dz, tag}.  hits: 4 floats {t, prim (as int bits), b1, b2}.
// Optional per-ray outputs (may be NULL): nodes visited, prims tested, edge proximity of the final hit.
static void run_parallel(size_t n, int threads, const std::function<void(size_t, size_t)>& f) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::vector<std::thread> th;
  size_t chunk = (n + threads - 1) / threads;
  for (int i = 0; i < threads; i++) {
    size_t a = std::min(n, (size_t)i * chunk), b = std::min(n, a + chunk);
    if (a < b) th.emplace_back(f, a, b);
  }
  for (auto& t : th) t.join();
}
void orc_intersect(orc_scene* s, const float* rays, uint64_t n, float* hits, uint32_t* nodes_visited, uint32_t* prims_tested, float* edge_prox, int threads) {
  run_parallel(n, threads, [&](size_t a, size_t b) {
    Counters& c = tls_counters();
    for (size_t i = a; i < b; i++) {
      const float* r = rays + 8 * i;
      Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
      Ray orig = ray;
      uint64_t nv0 = c.nodes_visited, pt0 = c.prims_tested;
      SurfaceInteraction si;
      int prim = s->scene.intersect(ray, si);
      float* h = hits + 4 * i;
      h[0] = prim >= 0 ? ray.t_max : INF;
      int32_t pi = prim; std::memcpy(&h[1], &pi, 4);
      h[2] = 0; h[3] = 0;
      if (nodes_visited) nodes_visited[i] = (uint32_t)(c.nodes_visited - nv0);
      if (prims_tested) prims_tested[i] = (uint32_t)(c.prims_tested - pt0);
      if (edge_prox) edge_prox[i] = INF;
      if (prim >= 0) {
        if (const Triangle* t = dynamic_cast<const Triangle*>(s->scene.prims[prim].shape.get())) {
          float b0, b1, b2, tt, ep; orig.t_max = INF;
          t->hit_test(orig, b0, b1, b2, tt, &ep);
          h[2] = b1; h[3] = b2;
          if (edge_prox) edge_prox[i] = ep;
        } else { h[2] = si.uv.x; h[3] = si.uv.y; }
      }
    }
  });
}
void orc_occluded(orc_scene* s, const float* rays, uint64_t n, uint8_t* out, uint32_t* nodes_visited, uint32_t* prims_tested, int threads) {
  run_parallel(n, threads, [&](size_t a, size_t b) {
    Counters& c = tls_counters();
    for (size_t i = a; i < b; i++) {
      const float* r = rays + 8 * i;
      Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
      uint64_t nv0 = c.nodes_visited, pt0 = c.prims_tested;
      out[i] = s->scene.intersect_p(ray) ? 1 : 0;
      if (nodes_visited) nodes_visited[i] = (uint32_t)(c.nodes_visited - nv0);
      if (prims_tested) prims_tested[i] = (uint32_t)(c.prims_tested - pt0);
    }
  });
}
// Full surface interaction of the closest hit, for shading-input parity: out = 24 floats per ray
// {t, prim, p.xyz, p_err.xyz, n.xyz, ns.xyz, ss(dpdu_s).xyz, ts(dpdv_s).xyz, wo.xyz, uv.xy}
void orc_intersect_full(orc_scene* s, const float* rays, uint64_t n, float* out) {
  for (size_t i = 0; i < n; i++) {
    const float* r = rays + 8 * i;
    Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
    SurfaceInteraction si;
    int prim = s->scene.intersect(ray, si);
    float* o = out + 24 * i;
    for (int k = 0; k < 24; k++) o[k] = 0;
    o[0] = prim >= 0 ? ray.t_max : INF; o[1] = (float)prim;
    if (prim < 0) continue;
    V3 v[7] = {si.hit.p, si.hit.p_error, si.hit.n, si.shading.n, si.shading.dpdu, si.shading.dpdv, si.hit.wo};
    for (int k = 0; k < 7; k++) { o[2 + 3 * k] = v[k].x; o[3 + 3 * k] = v[k].y; o[4 + 3 * k] = v[k].z; }
    o[23] = si.uv.x;
  }
}
// samples: 4 floats {p_film.x, p_film.y, p_lens.x, p_lens.y}; rays out: 8 floats
void orc_camera_rays(orc_scene* s, const float* samples, uint64_t n, float* rays) {
  for (size_t i = 0; i < n; i++) {
    CameraSample cs; cs.p_film = P2(samples[4 * i], samples[4 * i + 1]); cs.p_lens = P2(samples[4 * i + 2], samples[4 * i + 3]); cs.time = 0;
    Ray r = s->camera.generate_ray(cs);
    float* o = rays + 8 * i;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.t_max; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = 0;
  }
}
void orc_film_bounds(orc_scene* s, int32_t* cropped4, int32_t* sample4) {
  Film f; f.init(s->film_desc);
  cropped4[0] = f.cropped.x0; cropped4[1] = f.cropped.y0; cropped4[2] = f.cropped.x1; cropped4[3] = f.cropped.y1;
  Bounds2i sb = f.get_sample_bounds();
  sample4[0] = sb.x0; sample4[1] = sb.y0; sample4[2] = sb.x1; sample4[3] = sb.y1;
}

struct orc_stats {
  uint64_t camera_rays, regular_rays, shadow_rays, tri_tests, tri_hits, nodes_visited, prims_tested;
  double seconds_tiles, seconds_total; int32_t threads;
};
// film_xyzw: 4 floats per cropped pixel (may be NULL); rgb: 3 floats per cropped pixel (may be NULL).
// integrator_override: NULL = the scene's.
int orc_render(orc_scene* s, const rt_integrator* integrator_override, const rt_sampler* sampler_override, int sampler_kind, uint64_t seed,
               int threads, int tile_stride, float* film_xyzw, float* rgb, orc_stats* st, int tile_offset) {
  Film film; film.init(s->film_desc);
  Integrator integ; integ.desc = integrator_override ? *integrator_override : s->integrator;
  rt_sampler sd = sampler_override ? *sampler_override : s->sampler;
  RenderStats rs;
  render(s->scene, integ, s->camera, film, sd, sampler_kind, seed, threads, tile_stride, &rs, tile_offset);
  if (film_xyzw) std::memcpy(film_xyzw, film.pixels.data(), film.pixels.size() * sizeof(float));
  if (rgb) film_resolve(film, rgb);
  if (st) {
    st->camera_rays = rs.counters.camera_rays; st->regular_rays = rs.counters.regular_rays; st->shadow_rays = rs.counters.shadow_rays;
    st->tri_tests = rs.counters.tri_tests; st->tri_hits = rs.counters.tri_hits; st->nodes_visited = rs.counters.nodes_visited;
    st->prims_tested = rs.counters.prims_tested; st->seconds_tiles = rs.seconds_tiles; st->seconds_total = rs.seconds_total; st->threads = rs.threads;
  }
  return 0;
}
// Radiance of single samples with the counter sampler: pix = {x, y, sample_index} triples; out = 3 floats each.
void orc_li_samples(orc_scene* s, const rt_integrator* integrator_override, const rt_sampler* sampler_override, uint64_t seed, const int32_t* pix,
                    uint64_t n, float* out, float* p_film_out) {
  Integrator integ; integ.desc = integrator_override ? *integrator_override : s->integrator;
  rt_sampler sd = sampler_override ? *sampler_override : s->sampler;
  CounterSampler sampler((size_t)sd.spp, (size_t)sd.dimensions, seed);
  integ.preprocess(s->scene, sampler);
  for (size_t i = 0; i < n; i++) {
    sampler.start_pixel(pix[3 * i], pix[3 * i + 1]);
    sampler.s = (uint32_t)pix[3 * i + 2];
    CameraSample cs = sampler.get_camera_sample(pix[3 * i], pix[3 * i + 1]);
    Ray ray = s->camera.generate_ray_differential(cs);          // renderer.rs:110-111
    ray.scale_differentials(1.0f / std::sqrt((float)sampler.spp));
    Spectrum c = integ.li(s->scene, ray, sampler, 0);
    out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b;
    if (p_film_out) { p_film_out[2 * i] = cs.p_film.x; p_film_out[2 * i + 1] = cs.p_film.y; }
  }
}

// ---- small exports for restated reference KATs (tests/test_oracle_kats.py) ----
uint64_t orc_find_interval_le(const float* arr, uint64_t size, float x) { return find_interval(size, [&](size_t i) { return arr[i] <= x; }); }
uint64_t orc_distribution1d_sample_discrete(const float* f, uint64_t n, float u, float* pdf) { Distribution1D d; d.init(f, n); return d.sample_discrete(u, *pdf); }
float orc_distribution1d_sample_continuous(const float* f, uint64_t n, float u, float* pdf, uint64_t* off) {
  Distribution1D d; d.init(f, n); size_t o; float r = d.sample_continuous(u, *pdf, o); *off = o; return r;
}
int32_t orc_is_power_of_2(int32_t v) { return is_power_of_2(v); }
int32_t orc_round_up_pow_2(int32_t v) { return round_up_pow_2(v); }
float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
float orc_gamma(uint32_t n) { return gamma_f(n); }
float orc_radical_inverse(uint32_t base_index, uint64_t a) { return radical_inverse(base_index, a); }
void orc_pcg32_sequence(uint64_t seed, uint32_t* out, uint64_t n) { RNG r; r.set_sequence(seed); for (uint64_t i = 0; i < n; i++) out[i] = r.uniform_u32(); }
void orc_matrix_inverse(const float* m, float* out) { Matrix4 r = Matrix4::from(m).inverse(); std::memcpy(out, r.m, 64); }
// ZeroTwoSequence: fill `out` (spp * 5 floats: 2D#0.x, 2D#0.y, 1D#0, 2D#1.x, 2D#1.y) for the first pixel of tile `seed`
void orc_zerotwo_camera_samples(uint64_t spp, uint64_t dims, uint64_t seed, float* out) {
  ZeroTwoSequence z(spp, dims); z.reseed(seed); z.start_pixel(0, 0);
  size_t i = 0;
  do { P2 a = z.get_2d(); float t = z.get_1d(); P2 b = z.get_2d(); float* o = out + 5 * i++; o[0] = a.x; o[1] = a.y; o[2] = t; o[3] = b.x; o[4] = b.y; } while (z.start_next_sample());
}
// Counter sampler draws (for bit-parity against the device copy): out[0]=1D(counter), out[1..2]=2D(counter)
void orc_counter_draws(int32_t x, int32_t y, uint64_t seed, uint32_t s, uint32_t spp, uint32_t dims, uint32_t counter, float* out) {
  uint32_t ph = CounterSampler::pixel_hash(x, y, seed);
  out[0] = CounterSampler::draw_1d(ph, s, spp, dims, counter);
  P2 p = CounterSampler::draw_2d(ph, s, spp, dims, counter);
  out[1] = p.x; out[2] = p.y;
}
// fr_dielectric / TR microfacet helpers for BSDF unit parity
float orc_fr_dielectric(float c, float ei, float et) { return fr_dielectric(c, ei, et); }

// FresnelBlend::pdf for TrowbridgeReitz(ax, ay) (the reference's own property test: bsdf/fresnel.rs:427-436)
float orc_fresnel_blend_pdf(const float* wo, const float* wi, float ax, float ay) {
  Lobe l; l.kind = LOBE_FRESNEL_BLEND; l.r = Spectrum(1.0f); l.t = Spectrum(1.0f); l.dist.ax = ax; l.dist.ay = ay;
  return l.pdf(V3(wo[0], wo[1], wo[2]), V3(wi[0], wi[1], wi[2]));
}

// The Bsdf a material row builds on a canonical surface (n = +z, dpdu = +x): f, pdf and sample_f for world directions.
// out = { f.rgb, pdf, sampled f.rgb, sampled wi.xyz, sampled pdf, sampled flags, n_lobes, eta }  (14 floats)
int orc_material_bsdf(orc_scene* s, int row, int allow_multiple_lobes, const float* wo, const float* wi, const float* u, uint32_t flags, float* out) {
  if (row < 0 || (size_t)row >= s->scene.materials.size()) return -1;
  SurfaceInteraction si;
  si.hit = Interaction::make(V3(0, 0, 0), V3(0, 0, 0), V3(wo[0], wo[1], wo[2]), V3(0, 0, 1));
  si.dpdu = V3(1, 0, 0); si.dpdv = V3(0, 1, 0);
  si.shading.n = V3(0, 0, 1); si.shading.dpdu = V3(1, 0, 0); si.shading.dpdv = V3(0, 1, 0);
  Bsdf b;
  try {
    if (!compute_scattering_functions(s->scene.materials.data(), s->scene.materials[row], si, allow_multiple_lobes != 0, b)) return -2;
  } catch (const std::exception& e) { s->error = e.what(); return -3; }
  V3 o(wo[0], wo[1], wo[2]), i(wi[0], wi[1], wi[2]);
  Spectrum f = b.f(o, i, flags);
  out[0] = f.r; out[1] = f.g; out[2] = f.b; out[3] = b.pdf(o, i, flags);
  Spectrum sf; V3 swi; float spdf; uint32_t sampled;
  b.sample_f(o, P2(u[0], u[1]), flags, sf, swi, spdf, sampled);
  out[4] = sf.r; out[5] = sf.g; out[6] = sf.b; out[7] = swi.x; out[8] = swi.y; out[9] = swi.z; out[10] = spdf; out[11] = (float)sampled;
  out[12] = (float)b.n; out[13] = b.eta;
  return 0;
}
// Batched orc_material_bsdf: n triples (wo, wi, u) -> 14 floats each (same layout).  The independent pins (tests/test_pins_*.py:
// pdf normalisation, chi-square of sample_f against pdf, white furnace, reciprocity) run on this and on the device's twin
// rtgpu_bsdf_probe.
int orc_bsdf_probe(orc_scene* s, int row, int allow_multiple_lobes, uint64_t n, const float* wo, const float* wi, const float* u, uint32_t flags, float* out) {
  for (uint64_t i = 0; i < n; i++) {
    int rc = orc_material_bsdf(s, row, allow_multiple_lobes, wo + 3 * i, wi + 3 * i, u + 2 * i, flags, out + 14 * i);
    if (rc) return rc;
  }
  return 0;
}
// Light `light` (row of Scene::lights) probed from n reference points: ref = 6 floats each {p.xyz, n.xyz} (p_error = 0), u = 2 floats,
// w = 3 floats (a direction for pdf_li / le).  out = 16 floats each:
//   { Li.rgb, wi.xyz, pdf, p1.xyz (far end of the VisibilityTester), pdf_li(ref, w), le(w).rgb, pdf_li(ref, wi), is_delta }
int orc_light_probe(orc_scene* s, int light, uint64_t n, const float* ref, const float* u, const float* w, float* out) {
  if (light < 0 || (size_t)light >= s->scene.lights.size()) return -1;
  const Light& L = s->scene.lights[(size_t)light];
  for (uint64_t i = 0; i < n; i++) {
    const float* r = ref + 6 * i;
    Interaction it = Interaction::make(V3(r[0], r[1], r[2]), V3(0, 0, 0), V3(0, 0, 0), V3(r[3], r[4], r[5]));
    V3 wi; float pdf; Interaction p1;
    Spectrum li = L.sample_li(it, P2(u[2 * i], u[2 * i + 1]), wi, pdf, p1);
    float* o = out + 16 * i;
    o[0] = li.r; o[1] = li.g; o[2] = li.b; o[3] = wi.x; o[4] = wi.y; o[5] = wi.z; o[6] = pdf; o[7] = p1.p.x; o[8] = p1.p.y; o[9] = p1.p.z;
    V3 ww(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
    o[10] = L.pdf_li(it, ww);
    Spectrum le = L.le(Ray(it.p, ww, INF));
    o[11] = le.r; o[12] = le.g; o[13] = le.b;
    o[14] = pdf > 0.0f ? L.pdf_li(it, wi) : 0.0f;
    o[15] = L.is_delta() ? 1.0f : 0.0f;
  }
  return 0;
}
// The environment map of infinite light `light` as InfiniteAreaLight::new leaves it: level 0 of the MIP pyramid (w x h RGB texels,
// already times L * scale) and the scalar image the Distribution2D is built from ((2w) x (2h)).  Sizes first (texels == NULL).
int orc_light_env(orc_scene* s, int light, int32_t* w, int32_t* h, float* texels, float* func) {
  if (light < 0 || (size_t)light >= s->scene.lights.size() || s->scene.lights[(size_t)light].kind != RT_LIGHT_INFINITE) return -1;
  const Light& L = s->scene.lights[(size_t)light];
  *w = L.l_map.res_x; *h = L.l_map.res_y;
  if (texels) std::memcpy(texels, L.l_map.pyramid[0].d.data(), (size_t)L.l_map.res_x * L.l_map.res_y * 3 * sizeof(float));
  if (func) for (size_t v = 0; v < L.distribution.cond.size(); v++)
    std::memcpy(func + v * L.distribution.cond[v].func.size(), L.distribution.cond[v].func.data(), L.distribution.cond[v].func.size() * sizeof(float));
  return 0;
}
float orc_roughness_to_alpha(float r) { return TrowbridgeReitz::roughness_to_alpha(r); }

// ---- textures (tests/test_oracle_textures.py, tests/test_gpu_textures.py) ----
// Texture row `row` evaluated at n explicit surface points: in = 15 floats each {uv.xy, p.xyz, dpdx.xyz, dpdy.xyz, dudx, dvdx, dudy, dvdy};
// out = 3 floats each (a float texture fills all three).
int orc_texture_eval(orc_scene* s, int row, const float* in, uint64_t n, float* out) {
  if (row < 0 || (size_t)row >= s->scene.textures.rows.size()) return -1;
  for (size_t i = 0; i < n; i++) {
    const float* a = in + 15 * i;
    SurfaceInteraction si;
    si.uv = P2(a[0], a[1]); si.hit.p = V3(a[2], a[3], a[4]); si.dpdx = V3(a[5], a[6], a[7]); si.dpdy = V3(a[8], a[9], a[10]);
    si.dudx = a[11]; si.dvdx = a[12]; si.dudy = a[13]; si.dvdy = a[14];
    Spectrum v = s->scene.textures.eval(row, si);
    out[3 * i] = v.r; out[3 * i + 1] = v.g; out[3 * i + 2] = v.b;
  }
  return 0;
}
// MIP pyramid of imagemap row `row`: number of levels (level < 0), or the level's size (u, v) and, when out != NULL, its texels.
int orc_texture_mip_level(orc_scene* s, int row, int level, int32_t* u, int32_t* v, int32_t* channels, float* out) {
  if (row < 0 || (size_t)row >= s->scene.textures.rows.size() || !s->scene.textures.mips[(size_t)row]) return -1;
  const MIPMap& m = *s->scene.textures.mips[(size_t)row];
  if (level < 0) return (int)m.levels();
  if ((size_t)level >= m.levels()) return -1;
  const MIPMap::Level& l = m.pyramid[(size_t)level];
  *u = l.u; *v = l.v; *channels = m.nc;
  if (out) std::memcpy(out, l.d.data(), l.d.size() * sizeof(float));
  return 0;
}
float orc_noise(float x, float y, float z) { return noise(x, y, z); }
float orc_fbm(const float* p, const float* dpdx, const float* dpdy, float omega, uint32_t octaves) {
  return fbm(V3(p[0], p[1], p[2]), V3(dpdx[0], dpdx[1], dpdx[2]), V3(dpdy[0], dpdy[1], dpdy[2]), omega, octaves);
}
// Camera ray with its differentials, scaled by `scale` (renderer.rs:110-111): out = 8 + 12 floats {ray, rx_o, ry_o, rx_d, ry_d}
void orc_camera_rays_diff(orc_scene* s, const float* samples, uint64_t n, float scale, float* rays) {
  for (size_t i = 0; i < n; i++) {
    CameraSample cs; cs.p_film = P2(samples[4 * i], samples[4 * i + 1]); cs.p_lens = P2(samples[4 * i + 2], samples[4 * i + 3]); cs.time = 0;
    Ray r = s->camera.generate_ray_differential(cs);
    r.scale_differentials(scale);
    float* o = rays + 20 * i;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.t_max; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = 0;
    const V3 v[4] = {r.rx_o, r.ry_o, r.rx_d, r.ry_d};
    for (int k = 0; k < 4; k++) { o[8 + 3 * k] = v[k].x; o[9 + 3 * k] = v[k].y; o[10 + 3 * k] = v[k].z; }
  }
}

// Restated reference property tests (rustracer-core/tests/efloat.rs:52-154, tests/shapes.rs:16-54,
// bsdf/fresnel.rs:427-436 is covered in python).  Returns the number of violations.
int orc_selftest(uint64_t seed, char* msg, int msg_len) {
  std::mt19937_64 rng(seed);
  auto unif = [&](double a, double b) { return a + (b - a) * (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
  int fails = 0; std::string log;
  auto get_float = [&](float min_exp, float max_exp) {              // tests/efloat.rs:10-20
    float logu = (float)unif(min_exp, max_exp);
    float sign = unif(0, 1) < 0.5 ? -1.0f : 1.0f;
    return sign * std::pow(10.0f, logu);
  };
  auto get_efloat = [&](float min_exp, float max_exp) {             // tests/efloat.rs:22-50
    float val = get_float(min_exp, max_exp);
    float err = 0;
    int k = (int)(rng() % 4);
    if (k == 1) { uint32_t ulp = (uint32_t)(rng() % 1024); float offset = u2f(f2u(std::fabs(val)) + ulp); err = std::fabs(offset - val); }
    else if (k == 2) { uint32_t ulp = (uint32_t)(rng() % (1024 * 1024)); float offset = u2f(f2u(std::fabs(val)) + ulp); err = std::fabs(offset - val); }
    else if (k == 3) { err = (float)(4 * unif(0, 1)) * std::fabs(val); }
    return EFloat(val, err);
  };
  auto get_precise = [&](const EFloat& e) -> double {               // tests/efloat.rs:42-50: uniformly inside the interval
    if (e.low == e.high) return (double)e.v;
    double t = unif(0, 1);
    double p = (double)e.low * (1.0 - t) + (double)e.high * t;
    if (p < e.low) p = e.low; if (p > e.high) p = e.high;
    return p;
  };
  const int N = 10000;
  for (int trial = 0; trial < N; trial++) {
    EFloat a = get_efloat(-4, 4), b = get_efloat(-4, 4);
    double pa = get_precise(a), pb = get_precise(b);
    struct { const char* nm; EFloat r; double p; } ops[] = {
        {"abs", ef_abs(a), std::fabs(pa)}, {"add", a + b, pa + pb}, {"sub", a - b, pa - pb}, {"mul", a * b, pa * pb}};
    for (auto& o : ops) if (!((double)o.r.low <= o.p && o.p <= (double)o.r.high)) { fails++; if (log.size() < 400) log += std::string("efloat ") + o.nm + " containment; "; }
    if (!(b.low < 0 && b.high > 0)) {
      EFloat q = a / b; double pq = pa / pb;
      if (!((double)q.low <= pq && pq <= (double)q.high)) { fails++; if (log.size() < 400) log += "efloat div containment; "; }
    }
    EFloat aa = ef_abs(a);
    EFloat sq = ef_sqrt(aa); double psq = std::sqrt(std::fabs(pa));
    if (!((double)sq.low <= psq && psq <= (double)sq.high)) { fails++; if (log.size() < 400) log += "efloat sqrt containment; "; }
  }
  if (std::isnan(next_float_up(-0.0f)) || std::isnan(next_float_down(0.0f))) { fails++; log += "next_float of zero is NaN; "; }   // tests/efloat.rs:150-154
  // tests/shapes.rs:16-54: a ray leaving a full sphere's surface into the outer hemisphere must not re-intersect
  for (int i = 0; i < 200; i++) {
    auto p_exp = [&](float e) { float logu = (float)unif(-e, e); return std::pow(10.0f, logu); };
    V3 origin((float)unif(-1, 1) * p_exp(8), (float)unif(-1, 1) * p_exp(8), (float)unif(-1, 1) * p_exp(8));
    Transform o2w = Transform::translate(origin);
    float radius = p_exp(4);
    Sphere sphere(o2w, radius, -radius, radius, 360.0f, false);
    // ray from outside towards the centre
    P2 u((float)unif(0, 1), (float)unif(0, 1));
    V3 dir = uniform_sample_sphere(u);
    Ray r(origin + dir * (radius * 3.0f), -dir);
    SurfaceInteraction si; float t;
    if (!sphere.intersect(r, si, t)) continue;
    for (int j = 0; j < 200; j++) {
      P2 u2((float)unif(0, 1), (float)unif(0, 1));
      V3 w = uniform_sample_sphere(u2);
      if (dot(w, si.hit.n) < 0.0f) w = -w;
      Ray r2 = si.hit.spawn_ray(w);
      SurfaceInteraction s2; float t2;
      if (sphere.intersect_p(r2) || sphere.intersect(r2, s2, t2)) { fails++; if (log.size() < 400) log += "sphere re-intersection; "; }
    }
  }
  if (msg && msg_len > 0) { std::snprintf(msg, (size_t)msg_len, "%s", log.c_str()); }
  return fails;
}

}  // extern "C"
