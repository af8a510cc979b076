#include "context.hpp"
namespace rt {
void free_wave_buffers(rtgpu_ctx*) {}
void free_lightgrid(rtgpu_ctx*) {}
}
extern "C" {
int rtgpu_generate_rays(rtgpu_ctx* ctx, const rtgpu_render_desc*, const float*, size_t, rtgpu_ray*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_render(rtgpu_ctx* ctx, const rtgpu_render_desc*, rtgpu_stats*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_li_samples(rtgpu_ctx* ctx, const rtgpu_render_desc*, const int32_t*, size_t, float*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_read_film(rtgpu_ctx* ctx, float*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_resolve_film(rtgpu_ctx* ctx, float*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_film_device_ptr(rtgpu_ctx* ctx, void**, size_t*) { return rt::fail(ctx, RTGPU_ERR_UNSUPPORTED, "todo"); }
int rtgpu_reduce_film(rtgpu_ctx**, int, int) { return RTGPU_ERR_UNSUPPORTED; }
}
