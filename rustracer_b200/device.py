"""Device binding: ctypes over librtgpu.so (include/rtgpu.h).  There is no CPU fallback: a missing library or a
missing GPU raises."""
import ctypes as C

import numpy as np

from . import _abi as A
from ._native import load


class DeviceError(RuntimeError):
    pass


_STATUS = {-1: "CUDA error", -2: "bad argument", -3: "no scene uploaded", -4: "out of device memory", -5: "queue overflow", -6: "unsupported"}


def _lib():
    lib = load("rtgpu")
    if getattr(lib, "_rtgpu_ready", False):
        return lib
    vp, sz = C.c_void_p, C.c_size_t
    lib.rtgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.rtgpu_destroy.argtypes = [vp]
    lib.rtgpu_last_error.argtypes = [vp]
    lib.rtgpu_last_error.restype = C.c_char_p
    lib.rtgpu_upload_scene.argtypes = [vp, C.POINTER(A.rtgpu_scene_desc)]
    lib.rtgpu_intersect.argtypes = [vp, vp, sz, vp]
    lib.rtgpu_occluded.argtypes = [vp, vp, sz, vp]
    lib.rtgpu_intersect_device.argtypes = [vp, vp, sz, vp, C.POINTER(C.c_float)]
    lib.rtgpu_occluded_device.argtypes = [vp, vp, sz, vp, C.POINTER(C.c_float)]
    lib.rtgpu_intersect_device_stats.argtypes = [vp, vp, sz, vp, vp]
    lib.rtgpu_occluded_device_stats.argtypes = [vp, vp, sz, vp, vp]
    lib.rtgpu_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.rtgpu_generate_rays.argtypes = [vp, C.POINTER(A.rtgpu_render_desc), vp, sz, vp]
    lib.rtgpu_render.argtypes = [vp, C.POINTER(A.rtgpu_render_desc), C.POINTER(A.rtgpu_stats)]
    lib.rtgpu_li_samples.argtypes = [vp, C.POINTER(A.rtgpu_render_desc), vp, sz, vp]
    lib.rtgpu_read_film.argtypes = [vp, vp]
    lib.rtgpu_resolve_film.argtypes = [vp, vp]
    lib.rtgpu_film_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    lib.rtgpu_reduce_film.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.rtgpu_malloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.rtgpu_free.argtypes = [vp, vp]
    lib.rtgpu_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    lib.rtgpu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    lib.rtgpu_synchronize.argtypes = [vp]
    lib.rtgpu_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.rtgpu_host_free.argtypes = [vp, vp]
    lib.rtgpu_build_bvh.argtypes = [vp, vp, C.c_uint64, C.c_int, vp, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    lib.rtgpu_bsdf_probe.argtypes = [vp, C.c_uint32, C.c_int, sz, vp, vp, vp, C.c_uint32, vp]
    lib.rtgpu_light_probe.argtypes = [vp, C.c_uint32, sz, vp, vp, vp, vp]
    lib.rtgpu_comm_unique_id.argtypes = [vp]
    lib.rtgpu_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.rtgpu_reduce_film_nccl.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.rtgpu_comm_destroy.argtypes = [vp]
    lib.rtgpu_launch_count.argtypes = [vp]
    lib.rtgpu_launch_count.restype = C.c_uint64
    lib._rtgpu_ready = True
    return lib


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can alias a device buffer owned by the context."""

    def __init__(self, ptr, n_floats, owner):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class Device:
    """One rtgpu context == one GPU (rtgpu_create .. rtgpu_destroy)."""

    def __init__(self, index=0):
        self._lib = _lib()
        h = C.c_void_p()
        rc = self._lib.rtgpu_create(index, C.byref(h))
        if rc != 0:
            raise DeviceError(f"rtgpu_create(device={index}) failed: {_STATUS.get(rc, rc)} — a CUDA GPU is required, there is no CPU fallback")
        self._h = h
        self.index = index
        self.scene = None

    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._lib.rtgpu_host_free(self._h, p)
            self._pinned = []
            self._lib.rtgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise DeviceError(f"{_STATUS.get(rc, rc)}: {self._lib.rtgpu_last_error(self._h).decode()}")

    # ---- BVH construction (rtgpu_build_bvh) ------------------------------------------------------------
    def bvh_builder(self):
        """(function pointer, user pointer) for rth_flatten_with_builder: this context's rtgpu_build_bvh."""
        return C.cast(self._lib.rtgpu_build_bvh, C.c_void_p), self._h

    def build_bvh(self, prim_bounds, max_prims_per_node=4):
        """== BVH::new (SAH) on the device.  prim_bounds: (n, 6) float32 {min, max}.  Returns dict(node_lo (m,4), node_hi (m,4), ordered (n,), ms)."""
        b = np.ascontiguousarray(prim_bounds, np.float32).reshape(-1, 6)
        n = b.shape[0]
        lo, hi = np.zeros((2 * n, 4), np.float32), np.zeros((2 * n, 4), np.float32)
        ordered = np.zeros(n, np.uint32)
        m, ms = C.c_uint32(), C.c_float()
        self._check(self._lib.rtgpu_build_bvh(self._h, b.ctypes.data, n, int(max_prims_per_node), lo.ctypes.data, hi.ctypes.data, ordered.ctypes.data, C.byref(m), C.byref(ms)))
        return dict(node_lo=lo[: m.value].copy(), node_hi=hi[: m.value].copy(), ordered=ordered, ms=ms.value)

    # ---- scene -------------------------------------------------------------------------------------
    def upload(self, scene):
        """scene: rustracer_b200.host.Scene.  Flattened on demand, the SAH BVH built on this device (rtgpu_build_bvh: the same
        tree as the host builder's); a scene flattened beforehand is taken as it is."""
        if not scene._flat:
            scene.flatten(device=self)
        self._check(self._lib.rtgpu_upload_scene(self._h, scene.desc))
        self.scene = scene
        return self

    def set_option(self, name, value):
        self._check(self._lib.rtgpu_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self):
        return int(self._lib.rtgpu_launch_count(self._h))

    # ---- batched BVH::intersect / intersect_p, host buffers -------------------------------------------
    def intersect(self, rays, out=None):
        """== BVH::intersect over a host batch.  out: optional preallocated (n, 4) float32 array (then returned raw: {t, prim bits, b1, b2})."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        if out is not None:
            self._check(self._lib.rtgpu_intersect(self._h, rays.ctypes.data, n, out.ctypes.data))
            return out
        hits = np.zeros((n, 4), np.float32)
        self._check(self._lib.rtgpu_intersect(self._h, rays.ctypes.data, n, hits.ctypes.data))
        return dict(t=hits[:, 0].copy(), prim=hits[:, 1].copy().view(np.int32), b1=hits[:, 2].copy(), b2=hits[:, 3].copy())

    def occluded(self, rays, out=None):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        if out is None:
            out = np.zeros(n, np.uint8)
        self._check(self._lib.rtgpu_occluded(self._h, rays.ctypes.data, n, out.ctypes.data))
        return out

    # ---- device memory + device-resident batches ------------------------------------------------------
    def malloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._lib.rtgpu_malloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        self._check(self._lib.rtgpu_free(self._h, ptr))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self._lib.rtgpu_memcpy_h2d(self._h, dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, arr, dptr):
        assert arr.flags["C_CONTIGUOUS"]
        self._check(self._lib.rtgpu_memcpy_d2h(self._h, arr.ctypes.data, dptr, arr.nbytes))

    def pinned_empty(self, shape, dtype=np.float32):
        """numpy array over page-locked host memory (rtgpu_host_alloc): host-buffer batches and film read-backs move at PCIe rate.
        The memory lives until the Device is closed."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(self._lib.rtgpu_host_alloc(self._h, n, C.byref(p)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p.value)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def synchronize(self):
        self._check(self._lib.rtgpu_synchronize(self._h))

    def intersect_device(self, d_rays, n, d_hits, timed=True):
        ms = C.c_float(0)
        self._check(self._lib.rtgpu_intersect_device(self._h, d_rays, n, d_hits, C.byref(ms) if timed else None))
        return ms.value

    def occluded_device(self, d_rays, n, d_occ, timed=True):
        ms = C.c_float(0)
        self._check(self._lib.rtgpu_occluded_device(self._h, d_rays, n, d_occ, C.byref(ms) if timed else None))
        return ms.value

    def intersect_stats(self, rays):
        """Closest-hit batch with per-ray (nodes visited, primitives tested)."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        d_r, d_h, d_s = self.malloc(max(1, rays.nbytes)), self.malloc(max(1, 16 * n)), self.malloc(max(1, 8 * n))
        try:
            self.h2d(d_r, rays)
            self._check(self._lib.rtgpu_intersect_device_stats(self._h, d_r, n, d_h, d_s))
            hits = np.zeros((n, 4), np.float32)
            st = np.zeros((n, 2), np.uint32)
            self.d2h(hits, d_h)
            self.d2h(st, d_s)
        finally:
            self.free(d_r), self.free(d_h), self.free(d_s)
        return dict(t=hits[:, 0].copy(), prim=hits[:, 1].copy().view(np.int32), b1=hits[:, 2].copy(), b2=hits[:, 3].copy(), nodes=st[:, 0].copy(), prims=st[:, 1].copy())

    def occluded_stats(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        d_r, d_o, d_s = self.malloc(max(1, rays.nbytes)), self.malloc(max(1, n)), self.malloc(max(1, 8 * n))
        try:
            self.h2d(d_r, rays)
            self._check(self._lib.rtgpu_occluded_device_stats(self._h, d_r, n, d_o, d_s))
            occ = np.zeros(n, np.uint8)
            st = np.zeros((n, 2), np.uint32)
            self.d2h(occ, d_o)
            self.d2h(st, d_s)
        finally:
            self.free(d_r), self.free(d_o), self.free(d_s)
        return dict(occluded=occ, nodes=st[:, 0].copy(), prims=st[:, 1].copy())

    # ---- probes of the shading code (rtgpu_bsdf_probe / rtgpu_light_probe) ---------------------------------
    def bsdf_probe(self, row, wo, wi, u, allow_multiple_lobes=True, flags=31):
        """The Bsdf material `row` builds on a canonical surface (n = +z, dpdu = +x): (n, 3) wo, wi and (n, 2) u ->
        dict of arrays (f, pdf, sf, swi, spdf, sflags, n_lobes, eta), the layout of the oracle's probe."""
        wo, wi, u = (np.ascontiguousarray(a, np.float32) for a in (wo, wi, u))
        n = len(wo)
        out = np.zeros((n, 14), np.float32)
        self._check(self._lib.rtgpu_bsdf_probe(self._h, row, 1 if allow_multiple_lobes else 0, n, wo.ctypes.data, wi.ctypes.data, u.ctypes.data, flags, out.ctypes.data))
        return dict(f=out[:, 0:3], pdf=out[:, 3], sf=out[:, 4:7], swi=out[:, 7:10], spdf=out[:, 10], sflags=out[:, 11].astype(np.int32),
                    n_lobes=out[:, 12].astype(np.int32), eta=out[:, 13])

    def light_probe(self, light, ref, u, w):
        """Light row probed from reference points: ref (n, 6) {p, n}, u (n, 2), w (n, 3) -> dict(li, wi, pdf, p1, pdf_w, le_w, pdf_wi, delta)."""
        ref, u, w = (np.ascontiguousarray(a, np.float32) for a in (ref, u, w))
        n = len(ref)
        out = np.zeros((n, 16), np.float32)
        self._check(self._lib.rtgpu_light_probe(self._h, light, n, ref.ctypes.data, u.ctypes.data, w.ctypes.data, out.ctypes.data))
        return dict(li=out[:, 0:3], wi=out[:, 3:6], pdf=out[:, 6], p1=out[:, 7:10], pdf_w=out[:, 10], le_w=out[:, 11:14], pdf_wi=out[:, 14], delta=out[:, 15])

    # ---- camera / render / film -----------------------------------------------------------------------
    def generate_rays(self, rd, samples):
        """samples: (n, 4) {p_film.x, p_film.y, p_lens.x, p_lens.y} -> (n, 8) rays (camera.rs:150-202)."""
        samples = np.ascontiguousarray(samples, np.float32)
        n = samples.shape[0]
        rays = np.zeros((n, 8), np.float32)
        self._check(self._lib.rtgpu_generate_rays(self._h, C.byref(rd), samples.ctypes.data, n, rays.ctypes.data))
        return rays

    def render(self, rd):
        """== renderer::render (renderer.rs:22-143): accumulates into the device film; returns rtgpu_stats."""
        st = A.rtgpu_stats()
        self._check(self._lib.rtgpu_render(self._h, C.byref(rd), C.byref(st)))
        self._film_shape = (rd.cropped[3] - rd.cropped[1], rd.cropped[2] - rd.cropped[0])
        return st

    def li_samples(self, rd, pixels):
        """Radiance of individual (x, y, sample_index) camera samples -> (n, 3) RGB."""
        pixels = np.ascontiguousarray(pixels, np.int32)
        n = pixels.shape[0]
        out = np.zeros((n, 3), np.float32)
        self._check(self._lib.rtgpu_li_samples(self._h, C.byref(rd), pixels.ctypes.data, n, out.ctypes.data))
        return out

    def read_film(self, out=None):
        """X, Y, Z, weight per cropped pixel.  `out`: optional preallocated (H, W, 4) float32 buffer (numpy array or a
        pinned torch tensor) written in place."""
        h, w = self._film_shape
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._check(self._lib.rtgpu_read_film(self._h, ptr))
        return out

    def resolve_film(self):
        h, w = self._film_shape
        out = np.zeros((h, w, 3), np.float32)
        self._check(self._lib.rtgpu_resolve_film(self._h, out.ctypes.data))
        return out

    # ---- multi-process film reduce (rtgpu_comm_* / rtgpu_reduce_film_nccl) ---------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks)."""
        buf = C.create_string_buffer(128)
        rc = _lib().rtgpu_comm_unique_id(buf)
        if rc != 0:
            raise DeviceError(f"rtgpu_comm_unique_id: {_STATUS.get(rc, rc)}")
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        self._check(self._lib.rtgpu_comm_init(self._h, C.c_char_p(bytes(unique_id)), rank, world))

    def reduce_film(self, root=0):
        """One ncclReduce(sum) of the raw film accumulators into `root` (in place).  Returns the CUDA-event time of the collective (ms)."""
        ms = C.c_float(0)
        self._check(self._lib.rtgpu_reduce_film_nccl(self._h, root, C.byref(ms)))
        return ms.value

    def film_device_array(self):
        """The raw film accumulator as a __cuda_array_interface__ object (for torch.as_tensor + NCCL reduce)."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.rtgpu_film_device_ptr(self._h, C.byref(p), C.byref(n)))
        return _CudaArray(p.value, n.value, self)
