"""Host side: PBRT scene front end + SAH BVH + flattening (librthost.so, include/rthost.h)."""
import ctypes as C
import numpy as np

from . import _abi as A
from ._native import load


class SceneError(RuntimeError):
    """Scene file rejected (same failure modes as rustracer's `pbrt::parse_scene`)."""


def _lib():
    lib = load("rthost")
    if getattr(lib, "_rth_ready", False):
        return lib
    lib.rth_parse_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.rth_parse_string.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.rth_last_error.restype = C.c_char_p
    lib.rth_scene_free.argtypes = [C.c_void_p]
    lib.rth_scene_ir.argtypes = [C.c_void_p]
    lib.rth_scene_ir.restype = C.POINTER(A.rt_scene)
    lib.rth_n_warnings.argtypes = [C.c_void_p]
    lib.rth_warning.argtypes = [C.c_void_p, C.c_int]
    lib.rth_warning.restype = C.c_char_p
    lib.rth_film_filename.argtypes = [C.c_void_p]
    lib.rth_film_filename.restype = C.c_char_p
    lib.rth_integrator_name.argtypes = [C.c_void_p]
    lib.rth_integrator_name.restype = C.c_char_p
    lib.rth_flatten.argtypes = [C.c_void_p, C.c_int]
    lib.rth_flatten_with_builder.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.rth_scene_desc.argtypes = [C.c_void_p]
    lib.rth_scene_desc.restype = C.POINTER(A.rtgpu_scene_desc)
    lib.rth_bvh_build_seconds.argtypes = [C.c_void_p]
    lib.rth_bvh_build_seconds.restype = C.c_double
    lib.rth_n_triangles.argtypes = [C.c_void_p]
    lib.rth_n_triangles.restype = C.c_uint64
    lib.rth_slot_of_prim.argtypes = [C.c_void_p]
    lib.rth_slot_of_prim.restype = C.POINTER(C.c_uint32)
    lib.rth_render_desc.argtypes = [C.c_void_p, C.POINTER(A.rtgpu_render_desc)]
    lib.rth_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    lib.rth_param_header.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    lib.rth_ray_batch.argtypes = [C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
    lib.rth_write_image.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_int, C.c_int]
    lib._rth_ready = True
    return lib


class Scene:
    """A parsed scene: the state rustracer holds on entry to `RealApi::world_end` (api.rs:977-1010)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self._flat = False

    @classmethod
    def from_file(cls, path):
        lib = _lib()
        h = C.c_void_p()
        if lib.rth_parse_file(str(path).encode(), C.byref(h)) != 0:
            raise SceneError(lib.rth_last_error().decode())
        return cls(h.value)

    @classmethod
    def from_string(cls, text, search_dir=None):
        lib = _lib()
        h = C.c_void_p()
        sd = None if search_dir is None else str(search_dir).encode()
        if lib.rth_parse_string(text.encode(), sd, C.byref(h)) != 0:
            raise SceneError(lib.rth_last_error().decode())
        return cls(h.value)

    def __del__(self):
        try:
            if self._h:
                _lib().rth_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def ir(self):
        """Mutable `rt_scene` view (ctypes)."""
        return _lib().rth_scene_ir(self._h).contents

    @property
    def ir_ptr(self):
        return _lib().rth_scene_ir(self._h)

    @property
    def warnings(self):
        lib = _lib()
        return [lib.rth_warning(self._h, i).decode() for i in range(lib.rth_n_warnings(self._h))]

    @property
    def film_filename(self):
        return _lib().rth_film_filename(self._h).decode()

    @property
    def integrator_name(self):
        return _lib().rth_integrator_name(self._h).decode()

    def flatten(self, threads=0, device=None):
        """Build the SAH BVH and flatten.  device: a `Device` whose rtgpu_build_bvh builds the top-level tree (same tree)."""
        lib = _lib()
        if device is not None:
            fn, user = device.bvh_builder()
            rc = lib.rth_flatten_with_builder(self._h, threads, fn, user)
        else:
            rc = lib.rth_flatten(self._h, threads)
        if rc != 0:
            raise SceneError(lib.rth_last_error().decode())
        self._flat = True
        return self

    @property
    def desc(self):
        if not self._flat:
            self.flatten()
        return _lib().rth_scene_desc(self._h)

    @property
    def bvh_build_seconds(self):
        return _lib().rth_bvh_build_seconds(self._h)

    @property
    def n_triangles(self):
        return int(_lib().rth_n_triangles(self._h))

    def render_desc(self):
        rd = A.rtgpu_render_desc()
        lib = _lib()
        if lib.rth_render_desc(self._h, C.byref(rd)) != 0:
            raise SceneError(lib.rth_last_error().decode())
        return rd

    # numpy views of the flattened arrays (for tests / tools)
    def nodes(self):
        d = self.desc.contents
        n = d.n_nodes
        lo = np.ctypeslib.as_array(d.node_lo, shape=(n, 4)).copy()
        hi = np.ctypeslib.as_array(d.node_hi, shape=(n, 4)).copy()
        return lo, hi

    def prim_geom(self):
        d = self.desc.contents
        return np.ctypeslib.as_array(d.prim_geom, shape=(d.n_prims, 12)).copy()

    def prim_info(self):
        d = self.desc.contents
        return np.ctypeslib.as_array(d.prim_info, shape=(d.n_prims, 4)).copy()

    def slot_of_prim(self):
        d = self.desc.contents
        return np.ctypeslib.as_array(_lib().rth_slot_of_prim(self._h), shape=(d.n_prims,)).copy()


def tokenize(text):
    """Token list as strings, for the restated lexer KATs (pbrt/lexer.rs:269-336)."""
    lib = _lib()
    buf = C.create_string_buffer(max(4096, 16 * len(text) + 64))
    n = lib.rth_tokenize(text.encode(), buf, len(buf))
    if n < 0:
        raise SceneError(lib.rth_last_error().decode())
    return [t for t in buf.value.decode().split("\n") if t != ""]


def param_header(s):
    lib = _lib()
    t = C.c_int()
    name = C.create_string_buffer(256)
    if lib.rth_param_header(s.encode(), C.byref(t), name, 256) != 0:
        return None
    return t.value, name.value.decode()


def write_image(path, rgb):
    """rgb: (H, W, 3) float32 linear; .png (8-bit sRGB), .exr (32-bit float scan lines, ZIP) as imageio.rs:35-92 writes, or .pfm."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    lib = _lib()
    if lib.rth_write_image(str(path).encode(), rgb.ctypes.data_as(C.POINTER(C.c_float)), w, h) != 0:
        raise SceneError(lib.rth_last_error().decode())


def ray_batch(n, world_lo, world_hi, seed=5, any_hit=False, first=0, out=None):
    """rth_ray_batch: the multi-threaded twin of scenes.ray_batch.  Returns (n, 8) float32 {o, tmax, d, tag};
    `out` may be a preallocated (n, 8) float32 array (e.g. the numpy view of a pinned buffer)."""
    lo = np.ascontiguousarray(world_lo, np.float32)
    hi = np.ascontiguousarray(world_hi, np.float32)
    if out is None:
        out = np.empty((n, 8), np.float32)
    assert out.shape == (n, 8) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
    lib = _lib()
    if lib.rth_ray_batch(n, lo.ctypes.data_as(C.POINTER(C.c_float)), hi.ctypes.data_as(C.POINTER(C.c_float)), seed, 1 if any_hit else 0, first, out.ctypes.data) != 0:
        raise SceneError(lib.rth_last_error().decode())
    return out
