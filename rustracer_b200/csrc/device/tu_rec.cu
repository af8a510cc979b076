// Translation unit: Whitted / DirectLighting / AmbientOcclusion / Normal shading kernels.
#define RT_QUADRIC_INLINE 1   // shapes.cuh: quadric tests inlined (hot on sphere / disk / cylinder scenes)
#include "kernels_rec.cuh"
#include "launch.hpp"

namespace rt {
void launch_shade_recursive(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s) {
  if (p.sc.texmats) k_shade_recursive<true><<<blocks * 128 / RT_REC_THREADS, RT_REC_THREADS, 0, s>>>(p, parity);
  else k_shade_recursive<false><<<blocks * 128 / RT_REC_THREADS, RT_REC_THREADS, 0, s>>>(p, parity);
}
void launch_shade_ao(const RenderParams& p, unsigned blocks, cudaStream_t s) { k_shade_ao<<<blocks, 128, 0, s>>>(p); }
}  // namespace rt
