// Owner of an rt_scene (include/rt_scene.h): keeps every array alive and re-points the POD view.
#pragma once
#include <deque>
#include <string>
#include <vector>
#include "hmath.hpp"

namespace rth {

struct SceneStore {
  std::vector<rt_shape> shapes;
  std::vector<rt_area_light> area_lights;
  std::vector<rt_light> lights;
  std::vector<rt_material> materials;
  std::vector<rt_texture> textures;
  // stable storage for per-shape arrays
  std::deque<std::vector<int32_t>> index_arrays;
  std::deque<std::vector<float>> float_arrays;
  rt_scene view{};
  std::vector<std::string> warnings;

  const int32_t* keep(std::vector<int32_t>&& v) { index_arrays.push_back(std::move(v)); return index_arrays.back().data(); }
  const float* keep(std::vector<float>&& v) { float_arrays.push_back(std::move(v)); return float_arrays.back().data(); }

  const rt_scene* finish() {
    view.n_shapes = (uint32_t)shapes.size(); view.shapes = shapes.data();
    int32_t n_obj = 0;
    for (const rt_shape& s : shapes) { if (s.object_def >= n_obj) n_obj = s.object_def + 1; if (s.instance_of >= n_obj) n_obj = s.instance_of + 1; }
    view.n_objects = (uint32_t)n_obj;
    view.n_area_lights = (uint32_t)area_lights.size(); view.area_lights = area_lights.data();
    view.n_lights = (uint32_t)lights.size(); view.lights = lights.data();
    view.n_materials = (uint32_t)materials.size(); view.materials = materials.data();
    view.n_textures = (uint32_t)textures.size(); view.textures = textures.data();
    return &view;
  }
  size_t n_primitives() const {
    size_t n = 0;
    for (const rt_shape& s : shapes) if (s.object_def < 0) n += (s.kind == RT_SHAPE_TRIMESH) ? s.n_indices / 3 : 1;   // top-level primitives
    return n;
  }
};

}  // namespace rth
