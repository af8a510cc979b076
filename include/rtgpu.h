/* rtgpu.h — C ABI of the B200 wavefront renderer (librtgpu.so).
 *
 * This is the drop-in boundary of SURVEY.md §8(b): what a Rust `extern "C"` block + build.rs/nvcc in
 * rustracer-core would bind to replace the body of `renderer::render` (rustracer-core/src/renderer.rs:22-143)
 * and the `Primitive` aggregate calls `BVH::intersect` / `BVH::intersect_p`
 * (rustracer-core/src/bvh/mod.rs:366-501).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Call order (mirrors `RealApi::world_end`, rustracer-core/src/api.rs:992-1010):
 *   rtgpu_create -> rtgpu_upload_scene -> rtgpu_render -> rtgpu_read_film | rtgpu_resolve_film -> rtgpu_destroy
 * Every function returns 0 on success or a negative rtgpu_status; rtgpu_last_error() gives the text.
 * One context per GPU; calls on one context are serialised by the caller and block until finished.
 *
 * All scene arrays are HOST pointers owned by the caller; rtgpu_upload_scene copies them.
 * The host side that fills them (same SAH BVH as the reference, flattened) lives in
 * rustracer_b200/csrc/host and is exported through include/rthost.h.
 */
#ifndef RTGPU_H
#define RTGPU_H
#include <stddef.h>
#include <stdint.h>
#include "rt_scene.h"   /* rt_material / RT_TEX_* for the textured-material rows */
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtgpu_ctx rtgpu_ctx;

enum rtgpu_status {
  RTGPU_OK = 0, RTGPU_ERR_CUDA = -1, RTGPU_ERR_ARG = -2, RTGPU_ERR_NO_SCENE = -3, RTGPU_ERR_OOM = -4,
  RTGPU_ERR_QUEUE_OVERFLOW = -5, RTGPU_ERR_UNSUPPORTED = -6
};

/* Ray as `Ray {o, d, t_max}` (rustracer-core/src/ray.rs:9-15); d need not be unit (Q6). 32 B. */
typedef struct rtgpu_ray { float ox, oy, oz, tmax, dx, dy, dz; uint32_t tag; } rtgpu_ray;
/* Closest hit: t, prim = prim_number in Shape-directive order (bvh/mod.rs:92) or -1, and for triangles the
 * barycentrics b1,b2 of mesh.rs:296-299 (quadrics: u,v). 16 B. */
typedef struct rtgpu_hit { float t; int32_t prim; float b1, b2; } rtgpu_hit;

enum { RTGPU_PRIM_TRIANGLE = 0, RTGPU_PRIM_SPHERE = 1, RTGPU_PRIM_DISK = 2, RTGPU_PRIM_CYLINDER = 3 };
enum { RTGPU_PRIMFLAG_FLIP = 1 /* reverse_orientation ^ swaps_handedness */, RTGPU_PRIMFLAG_HAS_N = 2, RTGPU_PRIMFLAG_HAS_S = 4,
       RTGPU_PRIMFLAG_HAS_UV = 8, RTGPU_PRIMFLAG_REVERSE = 16 /* reverse_orientation alone (sphere.rs:300-302) */ };

/* Sphere / Disk / Cylinder with the derived constants the reference computes at construction
 * (shapes/sphere.rs:30-51, disk.rs:25-45, cylinder.rs:26-46).  176 B. */
typedef struct rtgpu_quadric {
  float o2w[16], w2o[16];      /* object_to_world.m and .m_inv */
  float radius, z_min, z_max, theta_min, theta_max, phi_max;
  float height, inner_radius;
  float area;
  uint32_t kind, flags, pad;
} rtgpu_quadric;

/* One `ObjectInstance` (TransformedPrimitive, primitive.rs:79-118).  Its slot in the top-level BVH has bit 1 of the second
 * geometry float4's w set and the row number in the third float4's w.  The definition's own BVH nodes and primitive slots live
 * in the same node / primitive arrays (absolute indices). */
typedef struct rtgpu_instance {
  float w2o[12], o2w[12];      /* rows 0..2 of primitive_to_world.m_inv and .m (affine only) */
  uint32_t root_node;          /* node index of the definition's BVH root, or 0xffffffff when it holds one primitive (api.rs:1071) */
  uint32_t first_slot;         /* that single primitive's slot (root_node == 0xffffffff) */
  float lo[3], hi[3];          /* bounds of the definition's BVH root (object space) */
  uint32_t prim_number;        /* number of the TransformedPrimitive in the scene's primitive list */
  uint32_t root_ref;           /* filled by rtgpu_upload_scene: traversal-engine reference of the root */
} rtgpu_instance;

enum { RTGPU_MAT_MATTE = 0, RTGPU_MAT_PLASTIC = 1, RTGPU_MAT_METAL = 2, RTGPU_MAT_GLASS = 3, RTGPU_MAT_MIRROR = 4, RTGPU_MAT_NONE = 5,
       RTGPU_MAT_LOBES = 6,     /* uber / substrate / translucent / mix: the host lists the lobes (rtgpu_lobe rows) */
       RTGPU_MAT_TEXTURED = 7 };/* some parameter is a non-constant texture, or there is a bump map: the device evaluates
                                   rtgpu_scene_desc.texmats[row] at every hit (same row index as the material) and lists the lobes there */

/* Texture row on the device (texture/ and mipmap.rs of the reference): rt_texture with the imagemap's MIP pyramid (MIPMap::new,
 * mipmap.rs:65-180, built by the host) addressed inside rtgpu_scene_desc.tex_data.  tex_data[0..128) is the EWA weight
 * table (mipmap.rs:35-45). */
#define RTGPU_MAX_MIP_LEVELS 16
typedef struct rtgpu_texture {
  int32_t kind, is_float;      /* RT_TEX_* */
  float value[3];
  int32_t tex1, tex2, amount;  /* child rows */
  int32_t mapping;             /* RT_TEXMAP_* */
  float su, sv, du, dv; float vs[3], vt[3];
  int32_t aa_none;
  float w2t[16];               /* fbm: world_to_texture.m */
  float omega; int32_t octaves;
  int32_t wrap, trilinear; float max_aniso;
  int32_t channels;            /* imagemap: 3 (spectrum) or 1 (float) floats per texel */
  int32_t n_levels;
  uint32_t level_offset[RTGPU_MAX_MIP_LEVELS];   /* first float of each level in tex_data, row-major u x v texels */
  int32_t level_u[RTGPU_MAX_MIP_LEVELS], level_v[RTGPU_MAX_MIP_LEVELS];
} rtgpu_texture;

/* One BxDF of a material whose lobe list the host builds (constant textures make it a per-material constant):
 * material/{uber,substrate,translucent,mixmat}.rs.  kind = RTGPU_LOBE_*. */
enum { RTGPU_LOBE_LAMBERT_R = 0, RTGPU_LOBE_OREN_NAYAR = 1, RTGPU_LOBE_SPEC_REFL = 2, RTGPU_LOBE_SPEC_TRANS = 3, RTGPU_LOBE_FRESNEL_SPEC = 4,
       RTGPU_LOBE_MICRO_REFL = 5, RTGPU_LOBE_MICRO_TRANS = 6, RTGPU_LOBE_LAMBERT_T = 7, RTGPU_LOBE_FRESNEL_BLEND = 8 };
typedef struct rtgpu_lobe {
  uint32_t kind;
  uint32_t n_scales;           /* ScaledBxDF wrappers around the lobe (bsdf/bxdf.rs:48-71), innermost first; 0..2 */
  float scale[2][3];
  float r[3], t[3];            /* reflectance / transmittance ; FresnelBlend: rs in r, rd in t */
  float on_a, on_b;            /* OrenNayar A, B */
  uint32_t fr_kind;            /* 0 no-op, 1 dielectric, 2 conductor */
  float fr_eta_i, fr_eta_t, c_eta_t[3], c_k[3];
  float ax, ay;                /* TrowbridgeReitz alpha (after the optional remap) */
  float eta_a, eta_b;
} rtgpu_lobe;                  /* 112 B */
/* Material constants after texture evaluation and after the host-side scalar prep the reference does with
 * libm at shading time (roughness_to_alpha: microfacet.rs:485-493; OrenNayar A/B: oren_nayar.rs:17-26). 116 B. */
typedef struct rtgpu_material {
  uint32_t type;
  float kd[3], ks[3], kr[3], kt[3], eta_rgb[3], k_rgb[3];
  float oren_a, oren_b; uint32_t use_oren_nayar;
  float alpha_u, alpha_v;      /* after optional remap */
  float eta;                   /* glass index */
  uint32_t glass_specular;     /* uroughness == 0 && vroughness == 0 (glass.rs:68) */
  /* RTGPU_MAT_LOBES: rows [lobe_first[a], +lobe_count[a]) of rtgpu_scene_desc.lobes, a = allow_multiple_lobes
   * (path: 1, whitted / directlighting: 0 — only a glass child of a mix tells them apart); bsdf_eta = Bsdf::eta */
  uint32_t lobe_first[2], lobe_count[2];
  float bsdf_eta;
} rtgpu_material;

enum { RTGPU_LIGHT_POINT = 0, RTGPU_LIGHT_DISTANT = 1, RTGPU_LIGHT_INFINITE = 2, RTGPU_LIGHT_AREA = 3 };
/* Light table row in `Scene::lights` order (api.rs:905-911,963); the row index is the light id. */
typedef struct rtgpu_light {
  uint32_t kind;
  float pos[3];                /* point */
  float dir[3];                /* distant, normalised (distant.rs:24-33) */
  float I[3];                  /* point I ; distant L ; area L_emit */
  uint32_t prim_slot;          /* area: index into the ordered primitive arrays */
  uint32_t two_sided, n_samples;
  float area;                  /* area: Shape::area() */
  float world_radius;          /* distant / infinite: Scene bounding-sphere radius (scene.rs:36-41) */
  float l2w[9], w2l[9];        /* infinite: upper 3x3 of light_to_world.m and .m_inv (vectors only) */
  uint32_t env_w, env_h;       /* infinite: map size (1x1 for a constant) */
  uint32_t env_texels;         /* offset (in floats) into env_data: 3*w*h RGB texels already * L*scale */
  uint32_t env_func, env_cdf, env_func_int;      /* Distribution2D conditional rows: (2w)x(2h), (2w+1)x(2h), 2h */
  uint32_t env_mfunc, env_mcdf; float env_mfunc_int; /* marginal: 2h, 2h+1 */
} rtgpu_light;

/* Flattened scene (SoA, host pointers). */
typedef struct rtgpu_scene_desc {
  /* LinearBVHNode (bvh/mod.rs:582-598) as two float4 per node:
   *   lo = {min.x, min.y, min.z, bits(offset)}   offset = primitives_offset (leaf) | second_child_offset (interior)
   *   hi = {max.x, max.y, max.z, bits((n_prims << 2) | axis)}   n_prims == 0 for interior nodes */
  uint32_t n_nodes; const float* node_lo; const float* node_hi;
  /* primitives in `ordered_prims` order (bvh/mod.rs:120).  prim_geom: 3 float4 per slot —
   *   triangle: {v0.xyz, bits(0)}, {v1.xyz, 0}, {v2.xyz, 0}  world space (mesh.rs:61)
   *   quadric : {0,0,0, bits(kind | quadric_index << 2)}, 0, 0 */
  uint32_t n_prims; const float* prim_geom;
  const uint32_t* prim_info;   /* 4 per slot: prim_number, material row, light row or 0xffffffff, RTGPU_PRIMFLAG_* */
  const float* tri_n;          /* 9 per slot (n0,n1,n2) or NULL when no mesh has normals */
  const float* tri_s;          /* 9 per slot or NULL */
  const float* tri_uv;         /* 6 per slot or NULL */
  uint32_t n_quadrics;  const rtgpu_quadric* quadrics;
  uint32_t n_materials; const rtgpu_material* materials;
  uint32_t n_lobes;     const rtgpu_lobe* lobes;       /* lobe lists of the RTGPU_MAT_LOBES materials (may be 0 / NULL) */
  /* textured materials (may be 0 / NULL): the neutral material rows (all of them, same indices as `materials`, so mix
   * children resolve), the texture rows and their float pool */
  uint32_t n_texmats;   const rt_material* texmats;
  uint32_t n_textures;  const rtgpu_texture* textures;
  uint32_t n_tex_floats; const float* tex_data;
  uint32_t n_instances; const rtgpu_instance* instances; /* object instances (may be 0 / NULL) */
  uint32_t n_lights;    const rtgpu_light* lights;
  uint32_t n_env_floats; const float* env_data;
  float world_lo[3], world_hi[3];  /* nodes[0].bounds */
} rtgpu_scene_desc;

enum { RTGPU_INTEGRATOR_PATH = 0, RTGPU_INTEGRATOR_WHITTED = 1, RTGPU_INTEGRATOR_DIRECT = 2, RTGPU_INTEGRATOR_AO = 3, RTGPU_INTEGRATOR_NORMAL = 4 };

/* Everything `renderer::render` reads from integrator, camera, film and sampler. */
typedef struct rtgpu_render_desc {
  int32_t integrator;          /* RTGPU_INTEGRATOR_* */
  int32_t max_depth;           /* as u8 in the reference (path.rs:42) */
  float rr_threshold;
  int32_t light_strategy;      /* 0 uniform, 1 spatial (path.rs:86-94; uniform is forced when n_lights == 1) */
  int32_t direct_strategy;     /* 0 all, 1 one */
  int32_t ao_samples;
  int32_t xres, yres;
  int32_t cropped[4];          /* Film::cropped_pixel_bounds x0,y0,x1,y1 (film.rs:66-75) */
  int32_t sample_bounds[4];    /* Film::get_sample_bounds (film.rs:249-257) */
  int32_t pixel_bounds[4];     /* SamplerIntegrator::pixel_bounds (integrator/mod.rs:35) */
  int32_t spp;                 /* already rounded up to a power of two (zerotwosequence.rs:32) */
  int32_t sampler_dims;
  float raster_to_camera[16], camera_to_world[16];   /* .m of each (camera.rs:38-60) */
  float lens_radius, focal_distance;
  float filter_radius[2]; float filter_table[256];    /* film.rs:92-102 */
  float max_sample_luminance, scale;
  /* work partition (multi-GPU): 16x16 tiles t with t % tile_world == tile_rank, samples [sample_begin, sample_end).  Tile t is the tile of
     row t / tiles_x whose column is (t % tiles_x + row) % tiles_x, tiles_x = tiles per row of the sample bounds: every tile belongs to
     exactly one rank, and a rank's share runs diagonally through the frame whatever tiles_x % tile_world is */
  int32_t tile_rank, tile_world;
  int32_t sample_begin, sample_end;
  uint64_t seed;
  int32_t clear_film;          /* 1 = zero the film first */
  int32_t wave_paths;          /* 0 = default; paths in flight per wave */
} rtgpu_render_desc;

/* The counters the reference reports plus device timings (CUDA events, milliseconds). */
typedef struct rtgpu_stats {
  uint64_t camera_rays, regular_rays, shadow_rays;
  uint64_t waves, kernel_launches;
  float ms_total, ms_closest, ms_anyhit, ms_shade, ms_other;
  uint64_t closest_launches, anyhit_launches;
  /* with option "count_traversal": BVH nodes visited / primitives tested by the closest-hit (incl. MIS) and any-hit rays */
  uint64_t nodes_closest, prims_closest, nodes_anyhit, prims_anyhit;
  /* rays walked by the closest-hit kernels / by the any-hit kernels (MIS rays towards infinite lights are "regular" rays
   * for the reference's counters but are answered by an any-hit walk: they count here under anyhit_rays) */
  uint64_t closest_rays, anyhit_rays;
  /* items handed to the shade kernels: path vertices incl. escaped rays (path), items of every level (whitted / directlighting), camera hits (ao) */
  uint64_t shaded_items;
  uint64_t lightgrid_rows;     /* sparse spatial light distribution: voxels that hold a distribution so far (0 in dense mode) */
} rtgpu_stats;

int rtgpu_create(int device, rtgpu_ctx** out);
int rtgpu_destroy(rtgpu_ctx* ctx);
const char* rtgpu_last_error(rtgpu_ctx* ctx);

int rtgpu_upload_scene(rtgpu_ctx* ctx, const rtgpu_scene_desc* scene);

/* == BVH::intersect / BVH::intersect_p over a batch; host buffers, H2D + kernel + D2H inside, pipelined in 4 Mi-ray chunks (copy-in, traversal
 * and copy-out of successive chunks overlap).  Buffers from rtgpu_host_alloc (page-locked) move at PCIe rate; pageable memory is accepted. */
int rtgpu_intersect(rtgpu_ctx* ctx, const rtgpu_ray* rays, size_t n, rtgpu_hit* hits);
int rtgpu_occluded(rtgpu_ctx* ctx, const rtgpu_ray* rays, size_t n, uint8_t* occluded);
/* Same, device-resident buffers (device pointers), asynchronous on the context's stream; elapsed_ms (may be NULL)
 * is the CUDA-event time of the kernel alone and forces a synchronise. */
int rtgpu_intersect_device(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, rtgpu_hit* d_hits, float* elapsed_ms);
int rtgpu_occluded_device(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint8_t* d_occluded, float* elapsed_ms);
/* Same walks, additionally writing per ray {BVH nodes visited, primitives tested} (2 x uint32 per ray, device
 * pointer): the N and T of the roofline's algorithmic bytes per ray; equal to the oracle's counts. */
int rtgpu_intersect_device_stats(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, rtgpu_hit* d_hits, uint32_t* d_stats);
int rtgpu_occluded_device_stats(rtgpu_ctx* ctx, const rtgpu_ray* d_rays, size_t n, uint8_t* d_occluded, uint32_t* d_stats);
/* == BVH::new with the SAH split method (bvh/mod.rs:80-135; recursive_build :137-312; flatten_bvh :314-358) on the device:
 * the same tree, node for node and slot for slot, as the reference builds on one CPU thread.  prim_bounds: 6 floats per
 * primitive {min.xyz, max.xyz} in prim_number order (host pointer).  Outputs (host pointers): node_lo / node_hi with room for
 * 2 * n_prims float4 each (the layout of rtgpu_scene_desc), ordered[n_prims] = slot -> prim_number, *n_nodes; build_ms (may be
 * NULL) = device time from the first to the last kernel.  Needs no scene.  The signature is rthost.h's rth_bvh_builder. */
int rtgpu_build_bvh(rtgpu_ctx* ctx, const float* prim_bounds, uint64_t n_prims, int max_prims_per_node, float* node_lo, float* node_hi, uint32_t* ordered,
                    uint32_t* n_nodes, float* build_ms);
/* Tunables: "sort_rays" (1 = bin batch rays by origin cell + direction octant before traversal; default 1),
 * "sort_min_rays" (batches smaller than this skip the binning; default 32768),
 * "lightgrid_dense_mib" / "lightgrid_sparse_mib" (SpatialLightDistribution, lightdistrib.rs:59-296: the per-voxel tables are built for every voxel up
 *   front while they fit the first budget [2048 MiB], else on demand for the voxels path vertices fall into, within the second budget [8192 MiB]),
 * "sort_items" (1 = rtgpu_render sorts the listed-lobes shade queue / the recursive integrators' items by material; default 1),
 * "overlap_bounces" (path integrator: 0 = every launch on one stream; 1 = the shadow / MIS traces of bounce b on a second stream beside
 *   the closest-hit launch of bounce b + 1; 2 = also the closest-hit MIS rays beside the any-hit MIS rays; default 2; results identical),
 * "waves_in_flight" (path integrator: 2 = two waves at a time on two stream groups with two sets of wave buffers, so the first bounces of one
 *   wave run under the short late-bounce launches of the other; 1 = one after the other; default 2; samples identical),
 * "engine_carveout" (shared memory the traversal-engine kernels ask for, in per cent of the SM maximum; the rest of the 256 KB is L1; -1 = the
 *   driver's choice; default 33 = the 100 KB configuration that holds their 8 x 9 KB of traversal stacks),
 * "profile" (1 = rtgpu_render times every launch with CUDA events and fills rtgpu_stats.ms_closest/anyhit/shade/other),
 * "count_traversal" (1 = rtgpu_render also fills rtgpu_stats.nodes_* / prims_*). */
int rtgpu_set_option(rtgpu_ctx* ctx, const char* name, int value);

/* == PerspectiveCamera::generate_ray_differential's ray (camera.rs:150-202) for explicit camera samples
 * {p_film.x, p_film.y, p_lens.x, p_lens.y}; host buffers.  Used by the parity tests. */
int rtgpu_generate_rays(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, const float* samples, size_t n, rtgpu_ray* rays);

/* == renderer::render: accumulate into the device film. */
int rtgpu_render(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, rtgpu_stats* stats);
/* Radiance `li()` of individual samples {x, y, sample_index} (no film); host buffers; rgb = 3 floats each. */
int rtgpu_li_samples(rtgpu_ctx* ctx, const rtgpu_render_desc* desc, const int32_t* pixels, size_t n, float* rgb);

/* Probes of the shading code the render kernels run (host buffers; tests/test_gpu_pins.py).
 * rtgpu_bsdf_probe: the Bsdf material `material_row` builds (Material::compute_scattering_functions, material/ *.rs) on a canonical surface
 *   (p = 0, n = +z, dpdu = +x), evaluated for n world-space triples wo[3], wi[3], u[2] with BxDFType `flags` (bsdf/bxdf.rs:8-16):
 *   out[14] = { Bsdf::f rgb, Bsdf::pdf, then Bsdf::sample_f(wo, u): f rgb, wi xyz, pdf, sampled type, lobe count, Bsdf::eta }
 *   (bsdf/mod.rs:94-251).  Textured materials are rejected (their lobes depend on the hit point).
 * rtgpu_light_probe: light `light_row` seen from n reference points ref[6] = {p, n} with u[2] and a direction w[3]:
 *   out[16] = { Light::sample_li: Li rgb, wi xyz, pdf, far end of the VisibilityTester xyz; Light::pdf_li(ref, w); Light::le(w) rgb;
 *   Light::pdf_li(ref, wi); is_delta } (light/{point,distant,diffuse,infinite}.rs). */
int rtgpu_bsdf_probe(rtgpu_ctx* ctx, uint32_t material_row, int allow_multiple_lobes, size_t n, const float* wo, const float* wi, const float* u, uint32_t flags,
                     float* out);
int rtgpu_light_probe(rtgpu_ctx* ctx, uint32_t light_row, size_t n, const float* ref, const float* u, const float* w, float* out);

/* Film accumulators X,Y,Z,weight per cropped pixel (film.rs:38-43), row-major; host buffer of 4*W*H floats. */
int rtgpu_read_film(rtgpu_ctx* ctx, float* xyzw);
/* == Film::write_image arithmetic (film.rs:196-234): RGB = max(0, XYZ->RGB / weight) * scale; 3*W*H floats. */
int rtgpu_resolve_film(rtgpu_ctx* ctx, float* rgb);
/* Device pointer + length (floats) of the raw accumulator (R,G,B,weight sums) so the host can run the film
 * reduce (NCCL via its own communicator, or rtgpu_reduce_film below). */
int rtgpu_film_device_ptr(rtgpu_ctx* ctx, void** d_ptr, size_t* n_floats);
/* Single-process multi-GPU: sum the films of ctxs[0..n) into ctxs[root] with peer copies + an add kernel. */
int rtgpu_reduce_film(rtgpu_ctx** ctxs, int n, int root);

/* Multi-process multi-GPU (one process and one context per GPU — SURVEY 8e): the film sum is ONE ncclReduce over NVLink / NVSwitch.
 *   rank 0:      rtgpu_comm_unique_id(id)  -> ship the RTGPU_COMM_ID_BYTES to the other ranks by any means (file, socket, MPI, torch store)
 *   every rank:  rtgpu_comm_init(ctx, id, rank, world);  rtgpu_render(... tile_rank = rank, tile_world = world ...);
 *                rtgpu_reduce_film_nccl(ctx, root, NULL); then on the root: rtgpu_read_film / rtgpu_resolve_film
 *   every rank:  rtgpu_comm_destroy(ctx) (or rtgpu_destroy, which calls it) at the SAME point of the job: ncclCommDestroy may wait for the peers, so a rank
 *                must not block on another rank's exit (waitpid, join) while its own communicator is still open
 * NCCL is opened at run time (dlopen "libnccl.so.2"), so the library has no link-time dependency on it; RTGPU_ERR_UNSUPPORTED when absent. */
#define RTGPU_COMM_ID_BYTES 128
int rtgpu_comm_unique_id(void* id_out);
int rtgpu_comm_init(rtgpu_ctx* ctx, const void* unique_id, int rank, int world);
int rtgpu_reduce_film_nccl(rtgpu_ctx* ctx, int root, float* elapsed_ms /* may be NULL: CUDA-event time of the collective on this rank */);
int rtgpu_comm_destroy(rtgpu_ctx* ctx);

/* device memory helpers so FFI callers need no CUDA runtime binding */
int rtgpu_malloc(rtgpu_ctx* ctx, size_t bytes, void** d_ptr);
int rtgpu_free(rtgpu_ctx* ctx, void* d_ptr);
int rtgpu_memcpy_h2d(rtgpu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int rtgpu_memcpy_d2h(rtgpu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int rtgpu_synchronize(rtgpu_ctx* ctx);
/* page-locked host memory for the host-buffer entry points and the film read-back */
int rtgpu_host_alloc(rtgpu_ctx* ctx, size_t bytes, void** h_ptr);
int rtgpu_host_free(rtgpu_ctx* ctx, void* h_ptr);
/* number of kernels launched by this context since creation (the bench's gpu_launches claim) */
uint64_t rtgpu_launch_count(rtgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
