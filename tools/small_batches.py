"""Latency of small ray batches through the host-buffer entry points (rtgpu_intersect / rtgpu_occluded, pinned host memory) on the
C3 scene, with and without the ray binning: python tools/small_batches.py  ->  one line per batch size (run on the GPU box).
The default break-even `sort_min_rays` (api.cu) comes from this table."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, host, scenes
from rustracer_b200.device import Device


def main():
    tmp = tempfile.mkdtemp()
    sc = Scene.from_string(scenes.c3_scene(tmp), search_dir=tmp)
    dev = Device(0).upload(sc)
    lo, hi = sc.nodes()
    print("rays | closest us (binned / unbinned) | any-hit us (binned / unbinned)")
    for n in (1 << 10, 1 << 12, 20000, 1 << 16, 1 << 18, 1 << 20, 1 << 22):
        row = [f"{n:8d}"]
        for any_hit in (False, True):
            src = host.ray_batch(n, lo[0, :3], hi[0, :3], seed=5, any_hit=any_hit)
            rays = dev.pinned_empty(src.shape, src.dtype)
            rays[...] = src
            out = dev.pinned_empty((n,), np.uint8) if any_hit else dev.pinned_empty((n, 4), np.float32)
            cell = []
            for sort_min in (0, 1 << 30):
                dev.set_option("sort_min_rays", sort_min)
                f = (lambda: dev.occluded(rays, out)) if any_hit else (lambda: dev.intersect(rays, out))
                for _ in range(5):
                    f()
                reps = 30 if n <= 1 << 18 else 8
                t = time.perf_counter()
                for _ in range(reps):
                    f()
                cell.append((time.perf_counter() - t) / reps * 1e6)
            row.append(f"{cell[0]:9.1f} / {cell[1]:9.1f}")
        print(" | ".join(row), flush=True)


if __name__ == "__main__":
    main()
