// Translation unit: k_shade_path<RT_PATH_MAT> — built once per material class (-DRT_PATH_MAT=0..5).
#include "kernels_path.cuh"
#include "launch.hpp"

#ifndef RT_PATH_MAT
#error "compile with -DRT_PATH_MAT=<material class>"
#endif
#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)

namespace rt {
void RT_CAT(launch_shade_path_, RT_PATH_MAT)(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s) {
  constexpr unsigned threads = RT_PATH_MAT == 6 ? RT_LOBES_THREADS : RT_SHADE_THREADS;
  k_shade_path<RT_PATH_MAT><<<blocks * 128 / threads, threads, 0, s>>>(p, parity);
}
}  // namespace rt
