/* rt_scene.h — neutral, pre-acceleration scene description (plain C, POD only).
 *
 * This is the state rustracer holds at the moment `RealApi::world_end` is entered
 * (rustracer-core/src/api.rs:977-1010): the list of shapes in `Shape`-directive order with
 * their bound material / area light, the light list in creation order, and the render
 * options (camera, film, filter, sampler, integrator, accelerator).  No BVH, no flattening.
 *
 * Producers: the host front end (rustracer_b200/csrc/host, PBRT parser + API state machine)
 * and test code.  Consumers: the host BVH builder/flattener that feeds include/rtgpu.h, and
 * — independently — the CPU oracle under oracle/ (which builds its own BVH from it).
 * It is a data format, not code: neither consumer calls the other.
 *
 * Conventions: matrices are row-major float[16] (m[r*4+c]) exactly like
 * `Matrix4x4.m[r][c]` (rustracer-core/src/geometry/matrix.rs:7-9).  A transform carries both
 * `m` and `m_inv` because the reference never re-derives one from the other after
 * composition (rustracer-core/src/transform.rs:332-351).
 */
#ifndef RT_SCENE_H
#define RT_SCENE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_transform { float m[16]; float m_inv[16]; } rt_transform;

enum { RT_SHAPE_TRIMESH = 0, RT_SHAPE_SPHERE = 1, RT_SHAPE_DISK = 2, RT_SHAPE_CYLINDER = 3,
       RT_SHAPE_INSTANCE = 4 };   /* an `ObjectInstance` directive: one TransformedPrimitive (api.rs:1052-1090, primitive.rs:79-118) */

/* One `Shape` directive (api.rs:913-966).  A trimesh expands to n_indices/3 primitives, a
 * quadric to one.  Primitive numbering = shape order, triangles in face order. */
typedef struct rt_shape {
  int32_t kind;
  rt_transform o2w;            /* CTM at the directive */
  int32_t reverse_orientation; /* graphics state flag (api.rs:968-975) */
  int32_t material;            /* index into rt_scene.materials (one row per Shape directive) */
  int32_t area_light;          /* index into rt_scene.area_lights, or -1 */
  /* RT_SHAPE_TRIMESH (shapes/mesh.rs:76-171, plymesh.rs:18-178): object-space arrays */
  uint32_t n_indices;  const int32_t* indices;
  uint32_t n_vertices; const float* P;   /* 3*n_vertices, object space (transformed by o2w at mesh creation, mesh.rs:61) */
  const float* N;                        /* 3*n_vertices or NULL — NOT transformed (mesh.rs:67) */
  const float* S;                        /* 3*n_vertices or NULL — NOT transformed (mesh.rs:68) */
  const float* uv;                       /* 2*n_vertices or NULL */
  /* quadrics: raw parameter values as given in the file (degrees for phimax) */
  float radius;                /* sphere, disk, cylinder */
  float zmin, zmax;            /* sphere: zmin/zmax ; cylinder: z_min/z_max */
  float phimax;                /* degrees */
  float height, inner_radius;  /* disk */
  /* Object instancing (api.rs:1019-1090).  A shape declared between ObjectBegin/ObjectEnd belongs to object definition
   * `object_def` (>= 0) and is NOT a primitive of the scene; its o2w is the CTM inside the definition.  An
   * RT_SHAPE_INSTANCE entry stands at the place of an `ObjectInstance "name"` directive in the primitive list: it is the
   * TransformedPrimitive over definition `instance_of` with primitive_to_world = o2w (the CTM at the directive).
   * Top-level shapes have object_def = -1. */
  int32_t object_def;
  int32_t instance_of;
} rt_shape;

/* `AreaLightSource "diffuse"` bound to a shape (light/diffuse.rs:39-51): one DiffuseAreaLight
 * per primitive of the shape, appended to the light list after the shape's primitives. */
typedef struct rt_area_light {
  float L[3];                  /* L * scale, already multiplied */
  int32_t n_samples;
  int32_t two_sided;
} rt_area_light;

enum { RT_LIGHT_POINT = 0, RT_LIGHT_DISTANT = 1, RT_LIGHT_INFINITE = 2, RT_LIGHT_AREA = 3 };

/* Light list entry in creation order (api.rs:905-911 and :963).  An RT_LIGHT_AREA entry
 * stands for ALL primitives of shape `shape` (one light each, in primitive order). */
typedef struct rt_light {
  int32_t kind;
  float pos[3];                /* point: world position (light/point.rs:28-35) */
  float dir[3];                /* distant: l2w * (from - to), NOT yet normalised (distant.rs:35-42) */
  float I[3];                  /* point: I*scale ; distant/infinite: L*scale */
  rt_transform l2w;            /* infinite */
  int32_t n_samples;           /* infinite: "samples" */
  int32_t env_w, env_h;        /* infinite: 0,0 = no map (constant 1x1 = I) */
  const float* env_rgb;        /* 3*env_w*env_h linear RGB, row-major, NOT yet scaled by I */
  int32_t shape;               /* RT_LIGHT_AREA: index into rt_scene.shapes */
} rt_light;

/* ---- textures (texture/ and mipmap.rs of the reference) -------------------------------------------------------------
 * One row per texture object the front end creates: every `Texture` directive, plus one RT_TEX_CONSTANT row for each literal
 * (or defaulted) tex1 / tex2 / amount parameter of a scale / mix / checkerboard texture (paramset.rs:406-443 wraps those in a
 * ConstantTexture), so children are always rows.  `is_float` tells Texture<f32> from Texture<Spectrum>. */
enum { RT_TEX_CONSTANT = 0, RT_TEX_SCALE = 1, RT_TEX_MIX = 2, RT_TEX_CHECKERBOARD = 3, RT_TEX_UV = 4, RT_TEX_IMAGEMAP = 5, RT_TEX_FBM = 6 };
enum { RT_TEXMAP_UV = 0, RT_TEXMAP_PLANAR = 1 };     /* texture/mod.rs:32-83 (spherical / cylindrical are unimplemented!() there) */
enum { RT_WRAP_REPEAT = 0, RT_WRAP_BLACK = 1, RT_WRAP_CLAMP = 2 };   /* mipmap.rs:27-32 */
typedef struct rt_texture {
  int32_t kind, is_float;
  float value[3];              /* constant (float textures use value[0]) */
  int32_t tex1, tex2, amount;  /* scale / mix / checkerboard children: rows of rt_scene.textures (amount: a float texture) */
  int32_t mapping;             /* checkerboard / uv / imagemap: RT_TEXMAP_* */
  float su, sv, du, dv;        /* UVMapping2D ; planar uses du, dv as ds, dt (checkerboard.rs:71-72) */
  float vs[3], vt[3];          /* PlanarMapping2D v1, v2 */
  int32_t aa_none;             /* checkerboard aamode "none" (default closedform) */
  rt_transform w2t;            /* fbm: IdentityMapping3D.world_to_texture = the CTM at the directive (fbm.rs:30, mod.rs:96-100) */
  float omega; int32_t octaves;/* fbm */
  /* imagemap (imagemap.rs:40-94): texels as MIPMap::new receives them — y-flipped, scaled, inverse-gamma'd, converted
   * (RGB triples for spectrum textures, luminance for float textures); row-major, img_w * img_h texels */
  int32_t img_w, img_h; const float* texels;
  int32_t wrap, trilinear; float max_aniso;
} rt_texture;

/* texture slot of a material parameter: rt_material.tex[slot] = 1 + row of rt_scene.textures, or 0 = the constant field
 * (so a zero-initialised material has no textures) */
enum { RT_TS_KD = 0, RT_TS_KS, RT_TS_KR, RT_TS_KT, RT_TS_ETA_RGB, RT_TS_K_RGB, RT_TS_SIGMA, RT_TS_ROUGHNESS, RT_TS_UROUGHNESS, RT_TS_VROUGHNESS,
       RT_TS_ETA, RT_TS_OPACITY, RT_TS_REFLECT, RT_TS_TRANSMIT, RT_TS_AMOUNT, RT_TS_BUMP, RT_TS_COUNT };

enum { RT_MAT_MATTE = 0, RT_MAT_PLASTIC = 1, RT_MAT_METAL = 2, RT_MAT_GLASS = 3, RT_MAT_MIRROR = 4, RT_MAT_NONE = 5,
       RT_MAT_UBER = 6, RT_MAT_SUBSTRATE = 7, RT_MAT_TRANSLUCENT = 8, RT_MAT_MIX = 9 };

/* Material.  A parameter bound to a constant texture is folded into its field (texture/constant.rs:10-39;
 * paramset.rs:406-443); a parameter bound to any other texture has its row in tex[] and the field is unused.  Field use per
 * type follows material/{matte,plastic,metal,glass,mirror,uber,substrate,
 * translucent,mixmat}.rs. */
typedef struct rt_material {
  int32_t type;
  float kd[3];                 /* matte Kd (0.5) ; plastic Kd (0.25) */
  float ks[3];                 /* plastic Ks (0.25) */
  float kr[3];                 /* glass Kr (1) ; mirror Kr (0.9) */
  float kt[3];                 /* glass Kt (1) */
  float eta_rgb[3];            /* metal eta (copper) */
  float k_rgb[3];              /* metal k (copper) */
  float sigma;                 /* matte */
  float roughness;             /* plastic (0.1) ; metal (0.01) */
  float uroughness, vroughness;/* glass (0,0) ; metal optional */
  int32_t has_uroughness, has_vroughness; /* metal: whether "uroughness"/"vroughness" were given */
  float eta;                   /* glass index (1.5) ; uber "eta" else "index" (1.5) */
  int32_t remap_roughness;
  /* uber.rs:32-57: Kd Ks (0.25), Kr Kt (0), roughness (0.1), optional u/vroughness, eta, opacity (1)
   * substrate.rs:23-38: Kd Ks (0.5), uroughness vroughness (0.1)
   * translucent.rs:26-44: Kd Ks (0.25), reflect transmit (0.5), roughness (0.1)
   * mixmat.rs:20-31: amount (0.5) and the two named materials (api.rs:1165-1176) as rows of rt_scene.materials */
  float opacity[3];
  float reflect[3], transmit[3];
  float amount[3];
  int32_t mix_a, mix_b;
  int32_t tex[16];             /* RT_TS_* -> 1 + texture row, or 0 ; tex[RT_TS_BUMP] = the "bumpmap" float texture (material/mod.rs:50-92) */
  int32_t textured;            /* 1 = some tex[] != 0 here or in a mix child: evaluated per hit on the device */
} rt_material;

enum { RT_FILTER_BOX = 0, RT_FILTER_GAUSSIAN = 1, RT_FILTER_TRIANGLE = 2, RT_FILTER_MITCHELL = 3 };
enum { RT_INTEGRATOR_PATH = 0, RT_INTEGRATOR_WHITTED = 1, RT_INTEGRATOR_DIRECT = 2, RT_INTEGRATOR_AO = 3, RT_INTEGRATOR_NORMAL = 4 };
enum { RT_LIGHTSTRATEGY_UNIFORM = 0, RT_LIGHTSTRATEGY_SPATIAL = 1 };
enum { RT_DIRECT_ALL = 0, RT_DIRECT_ONE = 1 };
enum { RT_SPLIT_SAH = 0, RT_SPLIT_MIDDLE = 1 };

typedef struct rt_camera {        /* camera.rs:74-123 */
  rt_transform c2w;               /* CTM.inverse() at the Camera directive (api.rs:720-730) */
  float fov;                      /* degrees, after the halffov override */
  float lens_radius, focal_distance;
  float screen_window[4];         /* xmin, xmax, ymin, ymax */
} rt_camera;

typedef struct rt_film {          /* film.rs:117-150 */
  int32_t xres, yres;
  float crop[4];                  /* xmin, xmax, ymin, ymax — already clamped/sorted */
  float scale, max_sample_luminance;
  int32_t filter;                 /* RT_FILTER_* (api.rs:181-192) */
  float filter_xw, filter_yw;     /* radius */
  float filter_a, filter_b;       /* gaussian alpha ; mitchell B, C */
} rt_film;

typedef struct rt_sampler { int32_t spp, dimensions; } rt_sampler;   /* zerotwosequence.rs:58-63 (spp NOT yet rounded) */

typedef struct rt_integrator {
  int32_t type;
  int32_t max_depth;              /* path.rs:50, whitted.rs:30, directlighting.rs:47 */
  float rr_threshold;             /* path.rs:51 */
  int32_t light_strategy;         /* path.rs:52 */
  int32_t direct_strategy;        /* directlighting.rs:48-58 */
  int32_t ao_samples;             /* ao.rs:19-24 */
  int32_t has_pixel_bounds; int32_t pixel_bounds[4]; /* path.rs:53-70: x0,x1,y0,y1 as given */
  int32_t reference_empty_pixel_bounds; /* 1 = reproduce the reference's empty pixel_bounds for
                                           whitted/directlighting/ao/normal */
} rt_integrator;

typedef struct rt_accel { int32_t split_method; int32_t max_node_prims; } rt_accel; /* bvh/mod.rs:63-78 */

typedef struct rt_scene {
  uint32_t n_objects;     /* object definitions (ObjectBegin blocks), numbered in file order */
  uint32_t n_shapes;      const rt_shape* shapes;
  uint32_t n_area_lights; const rt_area_light* area_lights;
  uint32_t n_lights;      const rt_light* lights;
  uint32_t n_materials;   const rt_material* materials;
  uint32_t n_textures;    const rt_texture* textures;
  rt_camera camera; rt_film film; rt_sampler sampler; rt_integrator integrator; rt_accel accel;
} rt_scene;

#ifdef __cplusplus
}
#endif
#endif
