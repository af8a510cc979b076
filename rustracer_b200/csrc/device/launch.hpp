// Launch wrappers: each kernel group is its own translation unit so nvcc can build them in parallel.
#pragma once
#include "wave.cuh"

namespace rt {

enum { TRACE_ENGINE = 0, TRACE_SIMPLE = 1, TRACE_COUNTING = 2 };

void launch_generate_rays(const RenderParams& p, const float4* samples, uint32_t n, float4* rays, cudaStream_t s);
void launch_raygen(const RenderParams& p, cudaStream_t s);
void launch_trace_closest(int mode, const RenderParams& p, const float4* ray_o, const float4* ray_d, const uint32_t* list, int count_idx,
                          HitRec* hits, unsigned blocks, cudaStream_t s);
void launch_classify(const RenderParams& p, const uint32_t* list, int count_idx, const HitRec* hits, bool from_hit_class, unsigned blocks, cudaStream_t s);
void launch_trace_shadow(bool atomic, int mode, const RenderParams& p, int q, unsigned blocks, cudaStream_t s);
void launch_trace_mis(bool atomic, int mode, const RenderParams& p, unsigned blocks, cudaStream_t s);
// shared-memory carve-out (per cent of the SM's maximum) of the traversal-engine kernels on the current device; < 0 leaves the driver's choice
int configure_trace_engines(int carveout_percent);
void launch_material_sort(const RenderParams& p, const uint32_t* list, int count_idx, uint32_t* hist, uint32_t n_bins, uint32_t* out, unsigned blocks, cudaStream_t s);
uint32_t ray_sort_bins();
void launch_ray_sort(const RenderParams& p, const uint32_t* list, int count_idx, uint32_t* keys, uint32_t* hist, uint32_t* out, unsigned blocks, cudaStream_t s);
void launch_shade_miss(const RenderParams& p, unsigned blocks, cudaStream_t s);
void launch_next_bounce(const RenderParams& p, int live_idx, int count_camera, int part, cudaStream_t s);
void launch_lightgrid(const DScene& sc, int nvx, int nvy, int nvz, float* table, cudaStream_t s);
void launch_lightgrid_bounce(const RenderParams& p, const uint32_t* list, int count_idx, float* table, unsigned blocks, cudaStream_t s);
void launch_film_add(const FilmParams& f, const float4* L, const float2* pfilm, uint32_t n, cudaStream_t s);
void launch_li_out(const float4* L, float ao_div, uint32_t n, float* out, cudaStream_t s);
void launch_film_xyz(const float4* film, size_t n, float4* out, cudaStream_t s);
void launch_film_resolve(const float4* film, size_t n, float scale, float* rgb, cudaStream_t s);
void launch_film_accumulate(float4* dst, const float4* src, size_t n, cudaStream_t s);
// material-sorted path shading: one translation unit per material class (tu_path.cu with -DRT_PATH_MAT=n)
void launch_shade_path_0(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_1(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_2(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_3(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_4(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_5(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_path_6(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);   // Q_LOBES
void launch_eval_textured(const RenderParams& p, const uint32_t* list, unsigned blocks, cudaStream_t s);   // texture pass before launch_shade_path_6
void launch_shade_recursive(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s);
void launch_shade_ao(const RenderParams& p, unsigned blocks, cudaStream_t s);

}  // namespace rt
