// Translation unit: k_shade_path<RT_PATH_MAT> — built once per material class (-DRT_PATH_MAT=0..5).
#ifndef RT_PATH_MAT
#error "compile with -DRT_PATH_MAT=<material class>"
#endif
// The lobes Material::compute_scattering_functions can list for this class with allow_multiple_lobes = true (bsdf.cuh make_bsdf; kinds
// are the LOBE_* numbers): the kernel's lobe dispatch is compiled for these alone.
#if RT_PATH_MAT == 0      /* matte.rs:37-62: LambertianReflection | OrenNayar */
#define RT_LOBE_KINDS ((1u << 0) | (1u << 1))
#define RT_MAX_BUILT_LOBES 1
#elif RT_PATH_MAT == 1    /* plastic.rs:45-74: LambertianReflection + MicrofacetReflection */
#define RT_LOBE_KINDS ((1u << 0) | (1u << 5))
#define RT_MAX_BUILT_LOBES 2
#elif RT_PATH_MAT == 2    /* metal.rs:50-81: MicrofacetReflection */
#define RT_LOBE_KINDS (1u << 5)
#define RT_MAX_BUILT_LOBES 1
#elif RT_PATH_MAT == 3    /* glass.rs:53-106: FresnelSpecular, or SpecularReflection / MicrofacetReflection + SpecularTransmission / MicrofacetTransmission */
#define RT_LOBE_KINDS ((1u << 2) | (1u << 3) | (1u << 4) | (1u << 5) | (1u << 6))
#define RT_MAX_BUILT_LOBES 2
#elif RT_PATH_MAT == 4    /* mirror.rs:30-48: SpecularReflection */
#define RT_LOBE_KINDS (1u << 2)
#define RT_MAX_BUILT_LOBES 1
#elif RT_PATH_MAT == 5    /* no material: no lobes */
#define RT_LOBE_KINDS (1u << 0)
#define RT_MAX_BUILT_LOBES 1
#endif
#include "kernels_path.cuh"
#include "launch.hpp"

#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)

namespace rt {
void RT_CAT(launch_shade_path_, RT_PATH_MAT)(const RenderParams& p, int parity, unsigned blocks, cudaStream_t s) {
  constexpr unsigned threads = RT_PATH_MAT == 6 ? RT_LOBES_THREADS : RT_SHADE_THREADS;
  k_shade_path<RT_PATH_MAT><<<blocks * 128 / threads, threads, 0, s>>>(p, parity);
}
}  // namespace rt
