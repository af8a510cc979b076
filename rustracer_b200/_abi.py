"""ctypes mirrors of include/rt_scene.h and include/rtgpu.h (plain-C structs, same field order)."""
import ctypes as C

c_f = C.c_float
c_i32 = C.c_int32
c_u32 = C.c_uint32
c_u64 = C.c_uint64
PF = C.POINTER(C.c_float)
PI32 = C.POINTER(C.c_int32)
PU32 = C.POINTER(C.c_uint32)


class rt_transform(C.Structure):
    _fields_ = [("m", c_f * 16), ("m_inv", c_f * 16)]


class rt_shape(C.Structure):
    _fields_ = [("kind", c_i32), ("o2w", rt_transform), ("reverse_orientation", c_i32), ("material", c_i32), ("area_light", c_i32),
                ("n_indices", c_u32), ("indices", PI32), ("n_vertices", c_u32), ("P", PF), ("N", PF), ("S", PF), ("uv", PF),
                ("radius", c_f), ("zmin", c_f), ("zmax", c_f), ("phimax", c_f), ("height", c_f), ("inner_radius", c_f),
                ("object_def", c_i32), ("instance_of", c_i32)]


class rt_area_light(C.Structure):
    _fields_ = [("L", c_f * 3), ("n_samples", c_i32), ("two_sided", c_i32)]


class rt_light(C.Structure):
    _fields_ = [("kind", c_i32), ("pos", c_f * 3), ("dir", c_f * 3), ("I", c_f * 3), ("l2w", rt_transform), ("n_samples", c_i32),
                ("env_w", c_i32), ("env_h", c_i32), ("env_rgb", PF), ("shape", c_i32)]


class rt_material(C.Structure):
    _fields_ = [("type", c_i32), ("kd", c_f * 3), ("ks", c_f * 3), ("kr", c_f * 3), ("kt", c_f * 3), ("eta_rgb", c_f * 3), ("k_rgb", c_f * 3),
                ("sigma", c_f), ("roughness", c_f), ("uroughness", c_f), ("vroughness", c_f), ("has_uroughness", c_i32), ("has_vroughness", c_i32),
                ("eta", c_f), ("remap_roughness", c_i32),
                ("opacity", c_f * 3), ("reflect", c_f * 3), ("transmit", c_f * 3), ("amount", c_f * 3), ("mix_a", c_i32), ("mix_b", c_i32),
                ("tex", c_i32 * 16), ("textured", c_i32)]


class rt_texture(C.Structure):
    _fields_ = [("kind", c_i32), ("is_float", c_i32), ("value", c_f * 3), ("tex1", c_i32), ("tex2", c_i32), ("amount", c_i32), ("mapping", c_i32),
                ("su", c_f), ("sv", c_f), ("du", c_f), ("dv", c_f), ("vs", c_f * 3), ("vt", c_f * 3), ("aa_none", c_i32), ("w2t", rt_transform),
                ("omega", c_f), ("octaves", c_i32), ("img_w", c_i32), ("img_h", c_i32), ("texels", PF), ("wrap", c_i32), ("trilinear", c_i32),
                ("max_aniso", c_f)]


class rt_camera(C.Structure):
    _fields_ = [("c2w", rt_transform), ("fov", c_f), ("lens_radius", c_f), ("focal_distance", c_f), ("screen_window", c_f * 4)]


class rt_film(C.Structure):
    _fields_ = [("xres", c_i32), ("yres", c_i32), ("crop", c_f * 4), ("scale", c_f), ("max_sample_luminance", c_f), ("filter", c_i32),
                ("filter_xw", c_f), ("filter_yw", c_f), ("filter_a", c_f), ("filter_b", c_f)]


class rt_sampler(C.Structure):
    _fields_ = [("spp", c_i32), ("dimensions", c_i32)]


class rt_integrator(C.Structure):
    _fields_ = [("type", c_i32), ("max_depth", c_i32), ("rr_threshold", c_f), ("light_strategy", c_i32), ("direct_strategy", c_i32),
                ("ao_samples", c_i32), ("has_pixel_bounds", c_i32), ("pixel_bounds", c_i32 * 4), ("reference_empty_pixel_bounds", c_i32)]


class rt_accel(C.Structure):
    _fields_ = [("split_method", c_i32), ("max_node_prims", c_i32)]


class rt_scene(C.Structure):
    _fields_ = [("n_objects", c_u32), ("n_shapes", c_u32), ("shapes", C.POINTER(rt_shape)), ("n_area_lights", c_u32), ("area_lights", C.POINTER(rt_area_light)),
                ("n_lights", c_u32), ("lights", C.POINTER(rt_light)), ("n_materials", c_u32), ("materials", C.POINTER(rt_material)),
                ("n_textures", c_u32), ("textures", C.POINTER(rt_texture)),
                ("camera", rt_camera), ("film", rt_film), ("sampler", rt_sampler), ("integrator", rt_integrator), ("accel", rt_accel)]


# enums (include/rt_scene.h)
RT_INTEGRATOR_PATH, RT_INTEGRATOR_WHITTED, RT_INTEGRATOR_DIRECT, RT_INTEGRATOR_AO, RT_INTEGRATOR_NORMAL = range(5)
RT_LIGHTSTRATEGY_UNIFORM, RT_LIGHTSTRATEGY_SPATIAL = 0, 1
RT_DIRECT_ALL, RT_DIRECT_ONE = 0, 1


class rtgpu_ray(C.Structure):
    _fields_ = [("ox", c_f), ("oy", c_f), ("oz", c_f), ("tmax", c_f), ("dx", c_f), ("dy", c_f), ("dz", c_f), ("tag", c_u32)]


class rtgpu_hit(C.Structure):
    _fields_ = [("t", c_f), ("prim", c_i32), ("b1", c_f), ("b2", c_f)]


class rtgpu_quadric(C.Structure):
    _fields_ = [("o2w", c_f * 16), ("w2o", c_f * 16), ("radius", c_f), ("z_min", c_f), ("z_max", c_f), ("theta_min", c_f), ("theta_max", c_f),
                ("phi_max", c_f), ("height", c_f), ("inner_radius", c_f), ("area", c_f), ("kind", c_u32), ("flags", c_u32), ("pad", c_u32)]


class rtgpu_material(C.Structure):
    _fields_ = [("type", c_u32), ("kd", c_f * 3), ("ks", c_f * 3), ("kr", c_f * 3), ("kt", c_f * 3), ("eta_rgb", c_f * 3), ("k_rgb", c_f * 3),
                ("oren_a", c_f), ("oren_b", c_f), ("use_oren_nayar", c_u32), ("alpha_u", c_f), ("alpha_v", c_f), ("eta", c_f), ("glass_specular", c_u32),
                ("lobe_first", c_u32 * 2), ("lobe_count", c_u32 * 2), ("bsdf_eta", c_f)]


class rtgpu_lobe(C.Structure):
    _fields_ = [("kind", c_u32), ("n_scales", c_u32), ("scale", (c_f * 3) * 2), ("r", c_f * 3), ("t", c_f * 3), ("on_a", c_f), ("on_b", c_f),
                ("fr_kind", c_u32), ("fr_eta_i", c_f), ("fr_eta_t", c_f), ("c_eta_t", c_f * 3), ("c_k", c_f * 3), ("ax", c_f), ("ay", c_f),
                ("eta_a", c_f), ("eta_b", c_f)]


class rtgpu_texture(C.Structure):
    _fields_ = [("kind", c_i32), ("is_float", c_i32), ("value", c_f * 3), ("tex1", c_i32), ("tex2", c_i32), ("amount", c_i32), ("mapping", c_i32),
                ("su", c_f), ("sv", c_f), ("du", c_f), ("dv", c_f), ("vs", c_f * 3), ("vt", c_f * 3), ("aa_none", c_i32), ("w2t", c_f * 16),
                ("omega", c_f), ("octaves", c_i32), ("wrap", c_i32), ("trilinear", c_i32), ("max_aniso", c_f), ("channels", c_i32), ("n_levels", c_i32),
                ("level_offset", c_u32 * 16), ("level_u", c_i32 * 16), ("level_v", c_i32 * 16)]


class rtgpu_instance(C.Structure):
    _fields_ = [("w2o", c_f * 12), ("o2w", c_f * 12), ("root_node", c_u32), ("first_slot", c_u32), ("lo", c_f * 3), ("hi", c_f * 3),
                ("prim_number", c_u32), ("root_ref", c_u32)]


class rtgpu_light(C.Structure):
    _fields_ = [("kind", c_u32), ("pos", c_f * 3), ("dir", c_f * 3), ("I", c_f * 3), ("prim_slot", c_u32), ("two_sided", c_u32), ("n_samples", c_u32),
                ("area", c_f), ("world_radius", c_f), ("l2w", c_f * 9), ("w2l", c_f * 9), ("env_w", c_u32), ("env_h", c_u32), ("env_texels", c_u32),
                ("env_func", c_u32), ("env_cdf", c_u32), ("env_func_int", c_u32), ("env_mfunc", c_u32), ("env_mcdf", c_u32), ("env_mfunc_int", c_f)]


class rtgpu_scene_desc(C.Structure):
    _fields_ = [("n_nodes", c_u32), ("node_lo", PF), ("node_hi", PF), ("n_prims", c_u32), ("prim_geom", PF), ("prim_info", PU32),
                ("tri_n", PF), ("tri_s", PF), ("tri_uv", PF), ("n_quadrics", c_u32), ("quadrics", C.POINTER(rtgpu_quadric)),
                ("n_materials", c_u32), ("materials", C.POINTER(rtgpu_material)), ("n_lobes", c_u32), ("lobes", C.POINTER(rtgpu_lobe)),
                ("n_texmats", c_u32), ("texmats", C.POINTER(rt_material)), ("n_textures", c_u32), ("textures", C.POINTER(rtgpu_texture)),
                ("n_tex_floats", c_u32), ("tex_data", PF), ("n_instances", c_u32), ("instances", C.POINTER(rtgpu_instance)),
                ("n_lights", c_u32), ("lights", C.POINTER(rtgpu_light)),
                ("n_env_floats", c_u32), ("env_data", PF), ("world_lo", c_f * 3), ("world_hi", c_f * 3)]


class rtgpu_render_desc(C.Structure):
    _fields_ = [("integrator", c_i32), ("max_depth", c_i32), ("rr_threshold", c_f), ("light_strategy", c_i32), ("direct_strategy", c_i32),
                ("ao_samples", c_i32), ("xres", c_i32), ("yres", c_i32), ("cropped", c_i32 * 4), ("sample_bounds", c_i32 * 4),
                ("pixel_bounds", c_i32 * 4), ("spp", c_i32), ("sampler_dims", c_i32), ("raster_to_camera", c_f * 16), ("camera_to_world", c_f * 16),
                ("lens_radius", c_f), ("focal_distance", c_f), ("filter_radius", c_f * 2), ("filter_table", c_f * 256),
                ("max_sample_luminance", c_f), ("scale", c_f), ("tile_rank", c_i32), ("tile_world", c_i32), ("sample_begin", c_i32),
                ("sample_end", c_i32), ("seed", c_u64), ("clear_film", c_i32), ("wave_paths", c_i32)]


class rtgpu_stats(C.Structure):
    _fields_ = [("camera_rays", c_u64), ("regular_rays", c_u64), ("shadow_rays", c_u64), ("waves", c_u64), ("kernel_launches", c_u64),
                ("ms_total", c_f), ("ms_closest", c_f), ("ms_anyhit", c_f), ("ms_shade", c_f), ("ms_other", c_f),
                ("closest_launches", c_u64), ("anyhit_launches", c_u64),
                ("nodes_closest", c_u64), ("prims_closest", c_u64), ("nodes_anyhit", c_u64), ("prims_anyhit", c_u64),
                ("closest_rays", c_u64), ("anyhit_rays", c_u64), ("shaded_items", c_u64), ("lightgrid_rows", c_u64)]
