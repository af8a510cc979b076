// Translation unit: BSDF and light probes (include/rtgpu.h rtgpu_bsdf_probe / rtgpu_light_probe).  They expose the device's
// Bsdf::{f, pdf, sample_f} (bsdf/mod.rs:94-251) and Light::{sample_li, pdf_li, le} (light/*.rs) on explicit inputs, so that the
// properties no reading of the Rust can get wrong — pdfs integrate to one, sample_f follows pdf (chi-square), energy is conserved,
// symmetric lobes are reciprocal, sample_li agrees with pdf_li — are checked on the code the render kernels run (tests/test_gpu_pins.py),
// next to the bit-level comparison with the oracle's twin probes.
#include "context.hpp"
#include "shade_common.cuh"
#include <functional>

namespace rt {

// canonical surface of oracle/orc_api.cpp orc_material_bsdf: p = 0, n = ns = +z, dpdu = +x
__global__ void __launch_bounds__(128) k_bsdf_probe(DScene sc, uint32_t row, int allow_multiple_lobes, uint32_t n, const float* __restrict__ wo_in,
                                                     const float* __restrict__ wi_in, const float* __restrict__ u_in, uint32_t flags, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 wo = v3(wo_in[3 * i], wo_in[3 * i + 1], wo_in[3 * i + 2]), wi = v3(wi_in[3 * i], wi_in[3 * i + 1], wi_in[3 * i + 2]);
  SurfHit si; si.p = v3(0, 0, 0); si.p_error = v3(0, 0, 0); si.n = v3(0, 0, 1); si.wo = wo; si.ns = v3(0, 0, 1); si.dpdu_s = v3(1, 0, 0);
  float* o = out + 14 * (size_t)i;
  for (int k = 0; k < 14; k++) o[k] = 0.0f;
  const rtgpu_material& mt = sc.materials[row];
  Bsdf bsdf;
  if (material_queue(mt.type) == Q_NONE || !make_bsdf(mt.type, mt, sc.lobes, si, allow_multiple_lobes != 0, bsdf)) return;
  const Spec f = bsdf_f(bsdf, wo, wi, flags);
  o[0] = f.r; o[1] = f.g; o[2] = f.b; o[3] = bsdf_pdf(bsdf, wo, wi, flags);
  Spec sf; V3 swi; float spdf; uint32_t sampled;
  bsdf_sample_f(bsdf, wo, mk2(u_in[2 * i], u_in[2 * i + 1]), flags, sf, swi, spdf, sampled);
  o[4] = sf.r; o[5] = sf.g; o[6] = sf.b; o[7] = swi.x; o[8] = swi.y; o[9] = swi.z; o[10] = spdf; o[11] = (float)sampled;
  o[12] = (float)bsdf.n; o[13] = bsdf.eta;
}

__global__ void __launch_bounds__(128) k_light_probe(DScene sc, uint32_t row, uint32_t n, const float* __restrict__ ref, const float* __restrict__ u_in,
                                                      const float* __restrict__ w_in, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const rtgpu_light& L = sc.lights[row];
  Inter it; it.p = v3(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]); it.p_error = v3(0, 0, 0); it.n = v3(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
  V3 wi; float pdf; Inter p1;
  const Spec li = light_sample_li(sc, L, it, mk2(u_in[2 * i], u_in[2 * i + 1]), wi, pdf, p1);
  float* o = out + 16 * (size_t)i;
  o[0] = li.r; o[1] = li.g; o[2] = li.b; o[3] = wi.x; o[4] = wi.y; o[5] = wi.z; o[6] = pdf; o[7] = p1.p.x; o[8] = p1.p.y; o[9] = p1.p.z;
  const V3 w = v3(w_in[3 * i], w_in[3 * i + 1], w_in[3 * i + 2]);
  o[10] = light_pdf_li(sc, L, it, w);
  const Spec le = light_le(sc, L, w);
  o[11] = le.r; o[12] = le.g; o[13] = le.b;
  o[14] = pdf > 0.0f ? light_pdf_li(sc, L, it, wi) : 0.0f;
  o[15] = light_is_delta(L) ? 1.0f : 0.0f;
}

// host buffers in, host buffers out
static int run_probe(rtgpu_ctx* ctx, const void* const* in, const size_t* in_floats, int n_in, size_t out_floats, float* out,
                     const std::function<void(float* const*, float*)>& launch) {
  cudaSetDevice(ctx->device);
  float* d_in[4] = {nullptr, nullptr, nullptr, nullptr}; float* d_out = nullptr;
  int rc = 0;
  for (int k = 0; k < n_in && !rc; k++) {
    rc = check_cuda(ctx, cudaMalloc((void**)&d_in[k], in_floats[k] * sizeof(float)), "cudaMalloc");
    if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(d_in[k], in[k], in_floats[k] * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "h2d");
  }
  if (!rc) rc = check_cuda(ctx, cudaMalloc((void**)&d_out, out_floats * sizeof(float)), "cudaMalloc");
  if (!rc) {
    launch(d_in, d_out);
    ctx->launches++;
    rc = check_cuda(ctx, cudaGetLastError(), "probe launch");
  }
  if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(out, d_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "d2h");
  if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync");
  for (int k = 0; k < n_in; k++) if (d_in[k]) cudaFree(d_in[k]);
  if (d_out) cudaFree(d_out);
  return rc;
}

}  // namespace rt

using namespace rt;

extern "C" {

int rtgpu_bsdf_probe(rtgpu_ctx* ctx, uint32_t material_row, int allow_multiple_lobes, size_t n, const float* wo, const float* wi, const float* u, uint32_t flags,
                     float* out) {
  if (!ctx || (n && (!wo || !wi || !u || !out))) return RTGPU_ERR_ARG;
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  if (material_row >= ctx->scene.n_materials) return fail(ctx, RTGPU_ERR_ARG, "material row out of range");
  if (ctx->h_materials[material_row].type == RTGPU_MAT_TEXTURED) return fail(ctx, RTGPU_ERR_UNSUPPORTED, "bsdf probe: textured materials depend on the hit point");
  if (n == 0) return RTGPU_OK;
  if (n > 0x7fffffffull) return fail(ctx, RTGPU_ERR_ARG, "probe batch too large");
  const void* in[3] = {wo, wi, u}; const size_t nf[3] = {3 * n, 3 * n, 2 * n};
  return run_probe(ctx, in, nf, 3, 14 * n, out, [&](float* const* d, float* d_out) {
    k_bsdf_probe<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->scene, material_row, allow_multiple_lobes, (uint32_t)n, d[0], d[1], d[2], flags, d_out);
  });
}

int rtgpu_light_probe(rtgpu_ctx* ctx, uint32_t light_row, size_t n, const float* ref, const float* u, const float* w, float* out) {
  if (!ctx || (n && (!ref || !u || !w || !out))) return RTGPU_ERR_ARG;
  if (!ctx->has_scene) return fail(ctx, RTGPU_ERR_NO_SCENE, "no scene uploaded");
  if (light_row >= ctx->scene.n_lights) return fail(ctx, RTGPU_ERR_ARG, "light row out of range");
  if (n == 0) return RTGPU_OK;
  if (n > 0x7fffffffull) return fail(ctx, RTGPU_ERR_ARG, "probe batch too large");
  const void* in[3] = {ref, u, w}; const size_t nf[3] = {6 * n, 2 * n, 3 * n};
  return run_probe(ctx, in, nf, 3, 16 * n, out, [&](float* const* d, float* d_out) {
    k_light_probe<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->scene, light_row, (uint32_t)n, d[0], d[1], d[2], d_out);
  });
}

}  // extern "C"
