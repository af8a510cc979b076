/* rthost.h — C ABI of the host side (librthost.so): PBRT scene front end, SAH BVH build, flattening.
 *
 * It restates, in C++, what the Rust host does before the hot path starts (the north star keeps that part
 * on the host): `pbrt::parse_scene` (rustracer-core/src/pbrt/mod.rs:15-25), the `Api` state machine
 * (api.rs:481-1091), `BVH::new` (bvh/mod.rs:80-135) and the object constructors, and produces the arrays
 * include/rtgpu.h consumes.  A Rust host would not need this library: it would fill rtgpu_scene_desc from
 * its own `BVH` / `Scene`.  No CUDA here.
 */
#ifndef RTHOST_H
#define RTHOST_H
#include "rt_scene.h"
#include "rtgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rth_scene rth_scene;

/* == pbrt::parse_scene up to (not including) renderer::render.  0 on success; on failure *out is NULL and
 * rth_last_error() holds the message (thread-local). */
int rth_parse_file(const char* path, rth_scene** out);
int rth_parse_string(const char* text, const char* search_dir, rth_scene** out);
const char* rth_last_error(void);
void rth_scene_free(rth_scene* s);

/* The parsed, pre-acceleration scene (state at RealApi::world_end). Mutable so callers can switch integrator /
 * sampler / film settings without re-parsing; call rth_flatten / rth_render_desc afterwards. */
rt_scene* rth_scene_ir(rth_scene* s);
int rth_n_warnings(rth_scene* s);
const char* rth_warning(rth_scene* s, int i);
const char* rth_film_filename(rth_scene* s);      /* "image.png" or "rt-<name>" (film.rs:118-125) */
const char* rth_integrator_name(rth_scene* s);

/* Build the reference-identical SAH BVH and flatten everything for the device. threads <= 0: all cores. */
int rth_flatten(rth_scene* s, int threads);
/* Same, with the scene's top-level SAH BVH built by `builder` (e.g. rtgpu_build_bvh with its context as `user`) instead of the
 * host builder; object definitions' trees and `splitmethod "middle"` stay on the host.  The tree must be the reference's. */
typedef int (*rth_bvh_builder)(void* user, const float* prim_bounds, uint64_t n_prims, int max_prims_per_node, float* node_lo, float* node_hi,
                               uint32_t* ordered, uint32_t* n_nodes, float* build_ms);
int rth_flatten_with_builder(rth_scene* s, int threads, rth_bvh_builder builder, void* user);
const rtgpu_scene_desc* rth_scene_desc(rth_scene* s);
double rth_bvh_build_seconds(rth_scene* s);
uint64_t rth_n_triangles(rth_scene* s);
const uint32_t* rth_slot_of_prim(rth_scene* s);   /* prim_number -> ordered slot */
/* Film/camera/integrator/sampler descriptor from the current IR (no geometry needed). */
int rth_render_desc(rth_scene* s, rtgpu_render_desc* out);

/* Lexer / parser probes for the restated reference KATs (pbrt/lexer.rs:269-336, parser.rs:311-363).
 * rth_tokenize writes one token per line ("Shape", "STR:abc", "NUMBER:1.5", "[", "]", "COMMENT"). Returns the
 * token count or -1. */
int rth_tokenize(const char* text, char* out, size_t out_len);
int rth_param_header(const char* s, int* type_out, char* name_out, size_t name_len);

/* The synthetic ray batches of the ray-batch microbenchmark, multi-threaded: ray i = PCG32 stream seed * 2^32 + first + i
 * (rng.rs:5-52): origin uniform in the world bounds grown 5 %, direction uniform on the sphere with t_max = inf (closest-hit batch) or the
 * segment to a second uniform point with t_max = 1 - 1e-4 (any_hit != 0).  rustracer_b200/scenes.py ray_batch is the definition. */
int rth_ray_batch(uint64_t n, const float* world_lo, const float* world_hi, uint64_t seed, int any_hit, uint64_t first, rtgpu_ray* out);

/* == imageio::write_image (imageio.rs:35-92): .png (8-bit sRGB, spectrum.rs:52-66) and .exr (32-bit float R, G, B scan lines, ZIP blocks);
 * for parity work also .pfm (raw float).  Any other extension fails with the reference's "Unsupported file format". */
int rth_write_image(const char* path, const float* rgb, int width, int height);

#ifdef __cplusplus
}
#endif
#endif
