"""C4 ray batches for profiling (run under ncu on the GPU box): python tools/c4_probe.py [--rays N] [--level L]
Builds the 10,014,720-triangle field, then launches the closest-hit and the any-hit batch `--reps` times each (device-resident)."""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, host, scenes
from rustracer_b200.device import Device


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    sc = Scene.from_string(scenes.c4_scene(tmp, level=a.level), search_dir=tmp)
    dev = Device(0).upload(sc)
    lo, hi = sc.nodes()
    n = a.rays
    d_r, d_o = dev.malloc(n * 32), dev.malloc(n * 16)
    for any_hit in (False, True):
        rays = host.ray_batch(n, lo[0, :3], hi[0, :3], seed=5, any_hit=any_hit)
        dev.h2d(d_r, rays)
        for _ in range(a.reps):
            ms = dev.occluded_device(d_r, n, d_o) if any_hit else dev.intersect_device(d_r, n, d_o)
        print("any-hit" if any_hit else "closest", f"{n / ms / 1e3:.1f} Mrays/s ({ms:.2f} ms)", flush=True)


if __name__ == "__main__":
    main()
