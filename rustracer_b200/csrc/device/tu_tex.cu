// Translation unit: the path integrator's texture pass (k_eval_textured).
#include "kernels_tex.cuh"
#include "launch.hpp"

namespace rt {
void launch_eval_textured(const RenderParams& p, const uint32_t* list, unsigned blocks, cudaStream_t s) { k_eval_textured<<<blocks * 128 / RT_TEX_THREADS, RT_TEX_THREADS, 0, s>>>(p, list); }
}  // namespace rt
