"""Host-side mirror of rustracer's integrator plugin interface for the GPU path.

rustracer selects an integrator by name in `make_integrator` (rustracer-core/src/api.rs:231-246) and drives it through
`trait SamplerIntegrator { pixel_bounds(); preprocess(scene, sampler); li(scene, ray, sampler, arena, depth) }`
(rustracer-core/src/integrator/mod.rs:34-47) from `renderer::render` (renderer.rs:22-143).  Per-ray `li()` is the wrong
granularity for a GPU, so the GPU integrators keep the trait's names and meaning but do their work per render:
`preprocess` uploads the flattened scene (and builds the light distribution lazily), `render` replaces the tile loop,
`li` evaluates individual camera samples (what `li()` returns for the sample's camera ray).

Multi-GPU: one process per GPU, scene replicated, 16x16 tiles dealt round-robin by rank, one
`torch.distributed` reduce of the film (NCCL over NVLink; gloo in the CPU tests of the plumbing).
"""
import numpy as np

from . import _abi as A
from .device import Device

_NAMES = {"path": A.RT_INTEGRATOR_PATH, "whitted": A.RT_INTEGRATOR_WHITTED, "directlighting": A.RT_INTEGRATOR_DIRECT,
          "ambientocclusion": A.RT_INTEGRATOR_AO, "normal": A.RT_INTEGRATOR_NORMAL}
# names accepted in `Integrator "<name>"` (the existing ones map onto the GPU when this backend is selected)
ALIASES = {"gpupath": "path", "gpuwhitted": "whitted", "gpudirectlighting": "directlighting", "gpuao": "ambientocclusion", "gpunormal": "normal"}


def tile_partition(sample_bounds, rank, world):
    """16x16 tiles of the sample bounds (renderer.rs:38-47) owned by `rank`: tile t (row-major) with t % world == rank."""
    x0, y0, x1, y1 = sample_bounds
    ntx, nty = max(0, (x1 - x0 + 15) // 16), max(0, (y1 - y0 + 15) // 16)
    return [t for t in range(ntx * nty) if t % world == rank], ntx, nty


def sample_partition(spp, rank, world):
    """Contiguous share of the sample indices [0, spp) for `rank` (the alternative partition for very high spp)."""
    per = (spp + world - 1) // world
    return min(spp, rank * per), min(spp, (rank + 1) * per)


def reduce_film(film, dst=0):
    """Sum the per-rank film accumulators into rank `dst` — the path's single collective.  `film` is a torch tensor
    (CUDA with the nccl backend, CPU with gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


class GpuSamplerIntegrator:
    """SamplerIntegrator on a B200.  scene: rustracer_b200.host.Scene (a parsed .pbrt); the integrator named in the scene
    file is used unless `name` overrides it."""

    def __init__(self, scene, name=None, device=0, rank=0, world=1, seed=0, wave_paths=0):
        self.scene = scene
        self.rank, self.world = rank, world
        if name is not None:
            name = ALIASES.get(name, name)
            if name not in _NAMES:
                raise ValueError(f"unknown integrator {name!r} (api.rs:231-246 accepts whitted, directlighting, path, normal; this backend adds ambientocclusion)")
            scene.ir.integrator.type = _NAMES[name]
        self.device = device if isinstance(device, Device) else Device(device)
        self.seed, self.wave_paths = seed, wave_paths
        self._ready = False

    def pixel_bounds(self):
        rd = self.scene.render_desc()
        return tuple(rd.pixel_bounds)

    def preprocess(self):
        """== SamplerIntegrator::preprocess + the scene hand-over: BVH build (host), flatten, upload."""
        if not self._ready:
            self.device.upload(self.scene)
            self._ready = True
        return self

    def render_desc(self):
        rd = self.scene.render_desc()
        rd.seed = self.seed
        rd.tile_rank, rd.tile_world = self.rank, self.world
        rd.wave_paths = self.wave_paths
        return rd

    def render(self, sample_range=None, clear=True):
        """== renderer::render for this rank's tiles.  Returns rtgpu_stats (the reference's ray counters)."""
        self.preprocess()
        rd = self.render_desc()
        if sample_range is not None:
            rd.sample_begin, rd.sample_end = sample_range
        rd.clear_film = 1 if clear else 0
        self.stats = self.device.render(rd)
        return self.stats

    def li(self, pixels):
        """Radiance of camera samples {x, y, sample index} (== `li()` of their camera rays)."""
        self.preprocess()
        return self.device.li_samples(self.render_desc(), pixels)

    def film(self):
        """X, Y, Z, weight per pixel (film.rs:38-43) of this rank."""
        return self.device.read_film()

    def image(self):
        """== Film::write_image arithmetic: linear RGB (H, W, 3).  With world > 1 call after `reduce()`."""
        return self.device.resolve_film()

    def reduce(self, dst=0):
        """The film reduce over NVLink (NCCL): in-place on the device film via torch.distributed."""
        import torch
        t = torch.as_tensor(self.device.film_device_array(), device=f"cuda:{self.device.index}")
        reduce_film(t, dst)
        torch.cuda.synchronize()


def render_file(path, device=0, threads=0, out=None, integrator=None):
    """`rustracer scene.pbrt`: parse, build, render on the GPU, write `rt-<name>` (film.rs:118-125).  Returns the path."""
    import os
    import time
    from .host import Scene, write_image
    sc = Scene.from_file(path)
    sc.flatten(threads)
    integ = GpuSamplerIntegrator(sc, name=integrator, device=device)
    t0 = time.time()
    st = integ.render()
    img = integ.image()
    name = out or os.path.join(os.path.dirname(os.path.abspath(path)), sc.film_filename)
    write_image(name, img)
    print(f"Render time: {time.time() - t0:.3f} s  camera rays {st.camera_rays}  regular {st.regular_rays}  shadow {st.shadow_rays}")
    return name
