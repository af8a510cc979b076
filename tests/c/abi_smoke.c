/* abi_smoke.c — the drop-in boundary exercised from plain C, no Python, no torch.
 *
 *   gcc -std=c11 -O1 -Iinclude tests/c/abi_smoke.c -Lrustracer_b200/lib -lrthost -lrtgpu -Wl,-rpath,$PWD/rustracer_b200/lib -lm -o build/abi_smoke
 *   build/abi_smoke            one process: parse -> flatten (device BVH build) -> create -> upload -> render -> read_film / resolve_film
 *   build/abi_smoke 2          two processes, one GPU each: tile shares, rtgpu_comm_init from an id passed through a pipe,
 *                              rtgpu_reduce_film_nccl into rank 0, compared with rank 0's own single-GPU film
 *
 * The call order is `RealApi::world_end`'s (rustracer-core/src/api.rs:992-1010): scene + camera + integrator + sampler, then
 * renderer::render (renderer.rs:22-143), then Film::write_image (film.rs:196-247). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include "rthost.h"

static const char* kScene =
    "LookAt 0 1.5 -6  0 0.8 0  0 1 0\nCamera \"perspective\" \"float fov\" [40]\n"
    "Film \"image\" \"integer xresolution\" [96] \"integer yresolution\" [64]\nSampler \"02sequence\" \"integer pixelsamples\" [8]\nPixelFilter \"box\"\n"
    "Integrator \"path\" \"integer maxdepth\" [5]\nWorldBegin\n"
    "LightSource \"infinite\" \"rgb L\" [0.3 0.35 0.45]\n"
    "AttributeBegin\nAreaLightSource \"diffuse\" \"rgb L\" [12 11 10]\nTranslate 0 4 0\nRotate 90 1 0 0\nShape \"disk\" \"float radius\" [1]\nAttributeEnd\n"
    "Material \"matte\" \"rgb Kd\" [0.6 0.6 0.55]\n"
    "Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point P\" [-8 0 -8  8 0 -8  8 0 8  -8 0 8]\n"
    "AttributeBegin\nMaterial \"plastic\" \"rgb Kd\" [0.7 0.2 0.2]\nTranslate -1 0.8 0\nShape \"sphere\" \"float radius\" [0.8]\nAttributeEnd\n"
    "AttributeBegin\nMaterial \"glass\"\nTranslate 1 0.7 -0.5\nShape \"sphere\" \"float radius\" [0.7]\nAttributeEnd\n"
    "WorldEnd\n";

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, ctx ? rtgpu_last_error(ctx) : rth_last_error()); return 10; } } while (0)

/* one rank's job: returns 0 and fills film (4 floats per pixel) */
static int render_share(int device, int rank, int world, const unsigned char* comm_id, float** film_out, size_t* n_pix_out, rtgpu_ctx** ctx_out) {
  rtgpu_ctx* ctx = NULL;
  rth_scene* sc = NULL;
  if (rth_parse_string(kScene, NULL, &sc) != 0) { fprintf(stderr, "parse: %s\n", rth_last_error()); return 11; }
  CHECK(rtgpu_create(device, &ctx));
  /* the SAH tree is built on the device, node for node the reference's (bvh/mod.rs:137-358) */
  if (rth_flatten_with_builder(sc, 0, (rth_bvh_builder)rtgpu_build_bvh, ctx) != 0) { fprintf(stderr, "flatten: %s\n", rth_last_error()); return 12; }
  CHECK(rtgpu_upload_scene(ctx, rth_scene_desc(sc)));
  rtgpu_render_desc rd;
  if (rth_render_desc(sc, &rd) != 0) return 13;
  rd.seed = 3; rd.tile_rank = rank; rd.tile_world = world; rd.clear_film = 1;
  if (comm_id) CHECK(rtgpu_comm_init(ctx, comm_id, rank, world));
  rtgpu_stats st;
  CHECK(rtgpu_render(ctx, &rd, &st));
  if (comm_id) CHECK(rtgpu_reduce_film_nccl(ctx, 0, NULL));
  const size_t n_pix = (size_t)(rd.cropped[2] - rd.cropped[0]) * (size_t)(rd.cropped[3] - rd.cropped[1]);
  float* film = (float*)malloc(n_pix * 4 * sizeof(float));
  CHECK(rtgpu_read_film(ctx, film));
  printf("rank %d/%d: %llu camera rays, %llu regular, %llu shadow, %llu launches, %.2f ms\n", rank, world, (unsigned long long)st.camera_rays,
         (unsigned long long)st.regular_rays, (unsigned long long)st.shadow_rays, (unsigned long long)st.kernel_launches, st.ms_total);
  *film_out = film; *n_pix_out = n_pix; *ctx_out = ctx;
  rth_scene_free(sc);
  return 0;
}

static int film_sane(const float* film, size_t n_pix, int expect_spp) {
  double y = 0;
  for (size_t i = 0; i < n_pix; i++) {
    /* box filter of radius 0.5: a sample weighs 1 in its own pixel (and also in the neighbour when it falls exactly on a pixel edge, film.rs:311-318) */
    if (fabsf(film[4 * i + 3] - (float)expect_spp) > 2.0f) { fprintf(stderr, "pixel %zu: weight %g, expected %d\n", i, film[4 * i + 3], expect_spp); return 0; }
    for (int c = 0; c < 3; c++) if (!(film[4 * i + c] >= 0.0f) || !isfinite(film[4 * i + c])) { fprintf(stderr, "pixel %zu: bad value\n", i); return 0; }
    y += film[4 * i + 1] / film[4 * i + 3];
  }
  y /= (double)n_pix;
  printf("mean luminance Y = %.5f\n", y);
  return y > 0.05 && y < 5.0;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 1;
  alarm(240);   /* a wedged collective must not hold the test box: SIGALRM ends the process (alarms are not inherited across fork: the child arms its own) */
  rtgpu_ctx* ctx = NULL;
  float* film = NULL; size_t n_pix = 0;
  if (world <= 1) {
    int rc = render_share(0, 0, 1, NULL, &film, &n_pix, &ctx);
    if (rc) return rc;
    if (!film_sane(film, n_pix, 8)) return 20;
    float* rgb = (float*)malloc(n_pix * 3 * sizeof(float));
    CHECK(rtgpu_resolve_film(ctx, rgb));
    if (rth_write_image("build/abi_smoke.png", rgb, 96, 64) != 0) fprintf(stderr, "write_image: %s\n", rth_last_error());
    /* batched BVH::intersect through the same context */
    rtgpu_ray ray = {0.0f, 1.5f, -6.0f, INFINITY, -1.0f, -0.7f, 6.0f, 0};   /* camera position towards the plastic sphere at (-1, 0.8, 0) */
    rtgpu_hit hit;
    CHECK(rtgpu_intersect(ctx, &ray, 1, &hit));
    printf("probe ray: prim %d at t = %g\n", hit.prim, hit.t);
    if (hit.prim < 0) return 21;
    CHECK(rtgpu_destroy(ctx));
    printf("abi_smoke ok\n");
    return 0;
  }
  /* two processes (fork before any CUDA call), NCCL id through a pipe */
  int fd[2];
  if (pipe(fd) != 0) return 30;
  pid_t child = fork();
  if (child < 0) return 31;
  unsigned char id[RTGPU_COMM_ID_BYTES];
  if (child == 0) {
    alarm(240);
    close(fd[1]);
    if (read(fd[0], id, sizeof(id)) != (ssize_t)sizeof(id)) return 32;
    int rc = render_share(1, 1, 2, id, &film, &n_pix, &ctx);
    if (!rc) rtgpu_destroy(ctx);
    return rc;
  }
  close(fd[0]);
  if (rtgpu_comm_unique_id(id) != 0) { fprintf(stderr, "no NCCL\n"); return 33; }
  if (write(fd[1], id, sizeof(id)) != (ssize_t)sizeof(id)) return 34;
  int rc = render_share(0, 0, 2, id, &film, &n_pix, &ctx);
  /* both ranks tear the communicator down at the same point of the job (ncclCommDestroy may wait for its peers): the child does it in
     rtgpu_destroy right after its render_share, so the parent must not sit in waitpid with its own communicator still open */
  if (!rc) rtgpu_comm_destroy(ctx);
  int status = 0;
  waitpid(child, &status, 0);
  if (rc || !WIFEXITED(status) || WEXITSTATUS(status) != 0) { fprintf(stderr, "rank failed: %d / %d\n", rc, status); return 35; }
  if (!film_sane(film, n_pix, 8)) return 36;
  /* against the same job on one GPU */
  rtgpu_ctx* solo = NULL; float* film1 = NULL; size_t n1 = 0;
  rc = render_share(0, 0, 1, NULL, &film1, &n1, &solo);
  if (rc) return rc;
  double worst = 0;
  for (size_t i = 0; i < n_pix * 4; i++) { double d = fabs(film[i] - film1[i]) / fmax(fabs(film1[i]), 1e-3); if (d > worst) worst = d; }
  printf("reduced film vs single-GPU film: max relative difference %.3g\n", worst);
  if (worst > 1e-5) return 37;
  rtgpu_destroy(solo); rtgpu_destroy(ctx);
  printf("abi_smoke (2 ranks, NCCL) ok\n");
  return 0;
}
