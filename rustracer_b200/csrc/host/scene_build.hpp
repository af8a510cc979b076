// rt_scene -> rtgpu_scene_desc + rtgpu_render_desc (host, product code).
// This is the "flatten LinearBVHNode, triangle, sphere, disc and cylinder data into SoA device arrays" step of
// the north star, plus the host-side constants the reference computes at object construction
// (Sphere::new, Disk::new, Film::new, PerspectiveCamera::new, InfiniteAreaLight::new, Scene::new ...).
#pragma once
#include <string>
#include <vector>
#include "bvh_builder.hpp"
#include "../../../include/rtgpu.h"

namespace rth {

struct FlatScene {
  FlatBvh bvh;
  uvec<float> prim_geom;                 // 12 floats per slot
  uvec<uint32_t> prim_info;              // 4 per slot
  uvec<float> tri_n, tri_s, tri_uv;
  std::vector<rtgpu_quadric> quadrics;
  std::vector<rtgpu_material> materials;
  std::vector<rtgpu_lobe> lobes;         // lobe lists of the RTGPU_MAT_LOBES materials
  std::vector<rtgpu_instance> instances; // object instances (TransformedPrimitive rows)
  std::vector<rtgpu_texture> textures;   // texture rows + MIP pyramids (only when some material is textured)
  std::vector<float> tex_data;
  std::vector<rtgpu_light> lights;
  std::vector<float> env_data;
  uvec<uint32_t> slot_of_prim;           // prim_number -> slot
  rtgpu_scene_desc desc{};
  rtgpu_render_desc render{};
  size_t n_triangles = 0;
  std::string error;
};

// threads: BVH build threads (<= 0: all).  Throws std::runtime_error on unsupported input.
// external_builder (may be null): builds the top-level SAH tree instead of build_bvh (include/rthost.h rth_bvh_builder).
typedef int (*ExternalBvhBuilder)(void* user, const float* prim_bounds, uint64_t n_prims, int max_prims_per_node, float* node_lo, float* node_hi,
                                  uint32_t* ordered, uint32_t* n_nodes, float* build_ms);
void flatten_scene(const rt_scene& in, int threads, FlatScene& out, ExternalBvhBuilder external_builder = nullptr, void* builder_user = nullptr);
// Film / camera / integrator / sampler part only (no geometry): fills out.render from `in`.
void make_render_desc(const rt_scene& in, rtgpu_render_desc& rd);
// MIPMap::new (mipmap.rs:65-180): Lanczos resampling of non-power-of-two images + box-filtered levels through the wrap mode.
struct MipLevel { int u = 0, v = 0; std::vector<float> d; };
std::vector<MipLevel> build_mip_pyramid(int rx, int ry, int nc, int wrap, const float* texels);
// Texture rows for the device: copies the parameters and builds each imagemap's MIP pyramid (MIPMap::new) into `pool`.
void build_textures(const rt_scene& in, std::vector<rtgpu_texture>& rows, std::vector<float>& pool);

}  // namespace rth
