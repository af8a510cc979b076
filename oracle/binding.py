"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (rustracer_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class orc_stats(C.Structure):
    _fields_ = [("camera_rays", C.c_uint64), ("regular_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("tri_tests", C.c_uint64),
                ("tri_hits", C.c_uint64), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64), ("seconds_tiles", C.c_double),
                ("seconds_total", C.c_double), ("threads", C.c_int32)]


def _cpu_stamp():
    """Model name + feature flags of this host: what `-march=native` specialises the native build for."""
    import hashlib
    try:
        info = open("/proc/cpuinfo").read().split("\n\n")[0]
        keep = [l for l in info.splitlines() if l.split(":")[0].strip() in ("model name", "flags")]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(native=False, quiet=True):
    """Compile the oracle with g++ (no FMA contraction).  native=True adds -march=native (GPU-box baseline); a native library that was built
    on another kind of CPU (it travels with the repository snapshot) is rebuilt for the host it is about to run on."""
    target = "native" if native else "all"
    cmd = ["make", "-C", _HERE, target]
    stamp = os.path.join(_HERE, "_build", "native.cpu")
    if native:
        have = open(stamp).read().strip() if os.path.exists(stamp) else None
        if have != _cpu_stamp():
            cmd.insert(1, "-B")
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)
    if native:
        with open(stamp, "w") as f:
            f.write(_cpu_stamp() + "\n")
    return os.path.join(_HERE, "_build", "liboracle_native.so" if native else "liboracle.so")


def lib(native=False):
    global _lib
    if _lib is not None and not native:
        return _lib
    path = os.path.join(_HERE, "_build", "liboracle_native.so" if native else "liboracle.so")
    if native or not os.path.exists(path):
        build(native=native)                 # native: no-op when up to date and built for this CPU
    l = C.CDLL(path)
    PF = C.POINTER(C.c_float)
    l.orc_scene_create.restype = C.c_void_p
    l.orc_scene_create.argtypes = [C.c_void_p]
    l.orc_scene_destroy.argtypes = [C.c_void_p]
    l.orc_scene_error.argtypes = [C.c_void_p]
    l.orc_scene_error.restype = C.c_char_p
    l.orc_scene_build_seconds.argtypes = [C.c_void_p]
    l.orc_scene_build_seconds.restype = C.c_double
    l.orc_bvh_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    l.orc_bvh_export.argtypes = [C.c_void_p, PF, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    l.orc_prim_world_vertices.argtypes = [C.c_void_p, PF]
    l.orc_intersect.argtypes = [C.c_void_p, PF, C.c_uint64, PF, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), PF, C.c_int]
    l.orc_occluded.argtypes = [C.c_void_p, PF, C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
    l.orc_intersect_full.argtypes = [C.c_void_p, PF, C.c_uint64, PF]
    l.orc_camera_rays.argtypes = [C.c_void_p, PF, C.c_uint64, PF]
    l.orc_film_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    l.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, PF, PF, C.POINTER(orc_stats), C.c_int]
    l.orc_li_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.c_uint64, PF, PF]
    l.orc_find_interval_le.argtypes = [PF, C.c_uint64, C.c_float]
    l.orc_find_interval_le.restype = C.c_uint64
    l.orc_distribution1d_sample_discrete.argtypes = [PF, C.c_uint64, C.c_float, PF]
    l.orc_distribution1d_sample_discrete.restype = C.c_uint64
    l.orc_distribution1d_sample_continuous.argtypes = [PF, C.c_uint64, C.c_float, PF, C.POINTER(C.c_uint64)]
    l.orc_distribution1d_sample_continuous.restype = C.c_float
    l.orc_next_float_up.argtypes = [C.c_float]
    l.orc_next_float_up.restype = C.c_float
    l.orc_next_float_down.argtypes = [C.c_float]
    l.orc_next_float_down.restype = C.c_float
    l.orc_gamma.argtypes = [C.c_uint32]
    l.orc_gamma.restype = C.c_float
    l.orc_radical_inverse.argtypes = [C.c_uint32, C.c_uint64]
    l.orc_radical_inverse.restype = C.c_float
    l.orc_pcg32_sequence.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.c_uint64]
    l.orc_matrix_inverse.argtypes = [PF, PF]
    l.orc_zerotwo_camera_samples.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, PF]
    l.orc_counter_draws.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, PF]
    l.orc_fr_dielectric.argtypes = [C.c_float] * 3
    l.orc_fr_dielectric.restype = C.c_float
    l.orc_roughness_to_alpha.argtypes = [C.c_float]
    l.orc_roughness_to_alpha.restype = C.c_float
    l.orc_fresnel_blend_pdf.argtypes = [PF, PF, C.c_float, C.c_float]
    l.orc_fresnel_blend_pdf.restype = C.c_float
    l.orc_material_bsdf.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF, PF, C.c_uint32, PF]
    l.orc_bsdf_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, PF, PF, PF, C.c_uint32, PF]
    l.orc_light_probe.argtypes = [C.c_void_p, C.c_int, C.c_uint64, PF, PF, PF, PF]
    l.orc_light_env.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), PF, PF]
    l.orc_selftest.argtypes = [C.c_uint64, C.c_char_p, C.c_int]
    l.orc_texture_eval.argtypes = [C.c_void_p, C.c_int, PF, C.c_uint64, PF]
    l.orc_texture_mip_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), PF]
    l.orc_noise.argtypes = [C.c_float] * 3
    l.orc_noise.restype = C.c_float
    l.orc_fbm.argtypes = [PF, PF, PF, C.c_float, C.c_uint32]
    l.orc_fbm.restype = C.c_float
    l.orc_camera_rays_diff.argtypes = [C.c_void_p, PF, C.c_uint64, C.c_float, PF]
    if not native:
        _lib = l
    return l


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class OracleScene:
    """CPU restatement of rustracer's Scene + BVH + integrators built from an `rt_scene` pointer."""

    def __init__(self, rt_scene_ptr, native=False):
        self._l = lib(native)
        self._h = C.c_void_p(self._l.orc_scene_create(C.cast(rt_scene_ptr, C.c_void_p)))
        err = self._l.orc_scene_error(self._h)
        if err:
            raise RuntimeError(err.decode())
        n, p, li = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._l.orc_bvh_info(self._h, C.byref(n), C.byref(p), C.byref(li))
        self.n_nodes, self.n_prims, self.n_lights = n.value, p.value, li.value

    def __del__(self):
        try:
            if self._h:
                self._l.orc_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def build_seconds(self):
        return self._l.orc_scene_build_seconds(self._h)

    def material_bsdf(self, row, wo, wi, u, allow_multiple_lobes=True, flags=31):
        """The Bsdf material `row` builds on a canonical surface (n = +z, dpdu = +x), evaluated for world directions.
        Returns dict(f, pdf, sf, swi, spdf, sflags, n_lobes, eta)."""
        wo, wi, u = (np.ascontiguousarray(a, np.float32) for a in (wo, wi, u))
        out = np.zeros(14, np.float32)
        rc = self._l.orc_material_bsdf(self._h, row, 1 if allow_multiple_lobes else 0, _pf(wo), _pf(wi), _pf(u), flags, _pf(out))
        if rc:
            raise RuntimeError(f"orc_material_bsdf: {rc} {self._l.orc_scene_error(self._h)}")
        return dict(f=out[0:3].copy(), pdf=float(out[3]), sf=out[4:7].copy(), swi=out[7:10].copy(), spdf=float(out[10]), sflags=int(out[11]),
                    n_lobes=int(out[12]), eta=float(out[13]))

    def bsdf_probe(self, row, wo, wi, u, allow_multiple_lobes=True, flags=31):
        """Batched material_bsdf: (n, 3) wo, wi and (n, 2) u -> dict of arrays (f, pdf, sf, swi, spdf, sflags, n_lobes, eta)."""
        wo, wi, u = (np.ascontiguousarray(a, np.float32) for a in (wo, wi, u))
        n = len(wo)
        out = np.zeros((n, 14), np.float32)
        rc = self._l.orc_bsdf_probe(self._h, row, 1 if allow_multiple_lobes else 0, n, _pf(wo), _pf(wi), _pf(u), flags, _pf(out))
        if rc:
            raise RuntimeError(f"orc_bsdf_probe: {rc} {self._l.orc_scene_error(self._h)}")
        return dict(f=out[:, 0:3], pdf=out[:, 3], sf=out[:, 4:7], swi=out[:, 7:10], spdf=out[:, 10], sflags=out[:, 11].astype(np.int32),
                    n_lobes=out[:, 12].astype(np.int32), eta=out[:, 13])

    def light_probe(self, light, ref, u, w):
        """Light row probed from reference points: ref (n, 6) {p, n}, u (n, 2), w (n, 3) -> dict(li, wi, pdf, p1, pdf_w, le_w, pdf_wi, delta)."""
        ref, u, w = (np.ascontiguousarray(a, np.float32) for a in (ref, u, w))
        n = len(ref)
        out = np.zeros((n, 16), np.float32)
        if self._l.orc_light_probe(self._h, light, n, _pf(ref), _pf(u), _pf(w), _pf(out)):
            raise RuntimeError("orc_light_probe: bad light row")
        return dict(li=out[:, 0:3], wi=out[:, 3:6], pdf=out[:, 6], p1=out[:, 7:10], pdf_w=out[:, 10], le_w=out[:, 11:14], pdf_wi=out[:, 14], delta=out[:, 15])

    def light_env(self, light):
        """Environment map of an infinite light after MIPMap::new: (texels (h, w, 3), distribution image (2h, 2w))."""
        w, h = C.c_int32(), C.c_int32()
        if self._l.orc_light_env(self._h, light, C.byref(w), C.byref(h), None, None):
            raise RuntimeError("orc_light_env: not an infinite light")
        tex = np.zeros((h.value, w.value, 3), np.float32)
        func = np.zeros((2 * h.value, 2 * w.value), np.float32)
        self._l.orc_light_env(self._h, light, C.byref(w), C.byref(h), _pf(tex), _pf(func))
        return tex, func

    def texture_eval(self, row, points):
        """Texture row evaluated at explicit surface points: (n, 15) {uv, p, dpdx, dpdy, dudx, dvdx, dudy, dvdy} -> (n, 3)."""
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 15)
        out = np.zeros((len(pts), 3), np.float32)
        if self._l.orc_texture_eval(self._h, row, _pf(pts), len(pts), _pf(out)):
            raise RuntimeError("orc_texture_eval: bad texture row")
        return out

    def texture_mip(self, row):
        """MIP pyramid of an imagemap row: list of (v, u, channels) arrays, finest first."""
        n = self._l.orc_texture_mip_level(self._h, row, -1, None, None, None, None)
        if n < 0:
            raise RuntimeError("orc_texture_mip_level: not an imagemap row")
        levels = []
        for i in range(n):
            u, v, c = C.c_int32(), C.c_int32(), C.c_int32()
            self._l.orc_texture_mip_level(self._h, row, i, C.byref(u), C.byref(v), C.byref(c), None)
            a = np.zeros((v.value, u.value, c.value), np.float32)
            self._l.orc_texture_mip_level(self._h, row, i, C.byref(u), C.byref(v), C.byref(c), _pf(a))
            levels.append(a)
        return levels

    def camera_rays_diff(self, samples, scale=1.0):
        s = np.ascontiguousarray(samples, np.float32).reshape(-1, 4)
        out = np.zeros((len(s), 20), np.float32)
        self._l.orc_camera_rays_diff(self._h, _pf(s), len(s), scale, _pf(out))
        return out

    def bvh(self):
        bounds = np.zeros((self.n_nodes, 6), np.float32)
        meta = np.zeros((self.n_nodes, 3), np.int64)
        ordered = np.zeros(self.n_prims, np.int32)
        self._l.orc_bvh_export(self._h, _pf(bounds), meta.ctypes.data_as(C.POINTER(C.c_int64)), ordered.ctypes.data_as(C.POINTER(C.c_int32)))
        return bounds, meta, ordered

    def prim_world_vertices(self):
        out = np.zeros((self.n_prims, 9), np.float32)
        self._l.orc_prim_world_vertices(self._h, _pf(out))
        return out

    def intersect(self, rays, threads=0, stats=True):
        """rays: (n, 8) float32 {o, tmax, d, tag}. Returns dict(t, prim, b1, b2, nodes, prims, edge)."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        hits = np.zeros((n, 4), np.float32)
        nodes = np.zeros(n, np.uint32)
        prims = np.zeros(n, np.uint32)
        edge = np.zeros(n, np.float32)
        self._l.orc_intersect(self._h, _pf(rays), n, _pf(hits), nodes.ctypes.data_as(C.POINTER(C.c_uint32)),
                              prims.ctypes.data_as(C.POINTER(C.c_uint32)), _pf(edge), threads)
        return dict(t=hits[:, 0].copy(), prim=hits[:, 1].copy().view(np.int32), b1=hits[:, 2].copy(), b2=hits[:, 3].copy(),
                    nodes=nodes, prims=prims, edge=edge)

    def occluded(self, rays, threads=0):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        out = np.zeros(n, np.uint8)
        nodes = np.zeros(n, np.uint32)
        prims = np.zeros(n, np.uint32)
        self._l.orc_occluded(self._h, _pf(rays), n, out.ctypes.data_as(C.POINTER(C.c_uint8)), nodes.ctypes.data_as(C.POINTER(C.c_uint32)),
                             prims.ctypes.data_as(C.POINTER(C.c_uint32)), threads)
        return dict(occluded=out, nodes=nodes, prims=prims)

    def intersect_full(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros((rays.shape[0], 24), np.float32)
        self._l.orc_intersect_full(self._h, _pf(rays), rays.shape[0], _pf(out))
        return out

    def camera_rays(self, samples):
        samples = np.ascontiguousarray(samples, np.float32)
        out = np.zeros((samples.shape[0], 8), np.float32)
        self._l.orc_camera_rays(self._h, _pf(samples), samples.shape[0], _pf(out))
        return out

    def film_bounds(self):
        c = (C.c_int32 * 4)()
        s = (C.c_int32 * 4)()
        self._l.orc_film_bounds(self._h, c, s)
        return list(c), list(s)

    def render(self, integrator=None, sampler=None, sampler_kind=0, seed=0, threads=0, tile_stride=1, tile_offset=0):
        """Returns (film_xyzw (H,W,4), rgb (H,W,3), stats). sampler_kind 0 = ZeroTwoSequence, 1 = counter sampler."""
        c, _ = self.film_bounds()
        w, h = c[2] - c[0], c[3] - c[1]
        film = np.zeros((h, w, 4), np.float32)
        rgb = np.zeros((h, w, 3), np.float32)
        st = orc_stats()
        ip = C.cast(C.byref(integrator), C.c_void_p) if integrator is not None else None
        sp = C.cast(C.byref(sampler), C.c_void_p) if sampler is not None else None
        self._l.orc_render(self._h, ip, sp, sampler_kind, seed, threads, tile_stride, _pf(film), _pf(rgb), C.byref(st), tile_offset)
        return film, rgb, st

    def li_samples(self, pixels, integrator=None, sampler=None, seed=0):
        pixels = np.ascontiguousarray(pixels, np.int32)
        n = pixels.shape[0]
        out = np.zeros((n, 3), np.float32)
        pf = np.zeros((n, 2), np.float32)
        ip = C.cast(C.byref(integrator), C.c_void_p) if integrator is not None else None
        sp = C.cast(C.byref(sampler), C.c_void_p) if sampler is not None else None
        self._l.orc_li_samples(self._h, ip, sp, seed, pixels.ctypes.data_as(C.POINTER(C.c_int32)), n, _pf(out), _pf(pf))
        return out, pf
