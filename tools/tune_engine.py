"""Sweep the traversal engine's scheduling knobs on the GPU box: python tools/tune_engine.py [--big]"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--nt", default="8,12,16,20")
    ap.add_argument("--rt", default="4,8,12,16")
    a = ap.parse_args()
    NT = [int(x) for x in a.nt.split(",")]
    RT = [int(x) for x in a.rt.split(",")]
    print("variant:", os.environ.get("RT_LIB_VARIANT", "(default)"))
    dev = Device(0)
    tmp = tempfile.mkdtemp()
    sc = Scene.from_string(scenes.c3_scene(tmp, level=5), search_dir=tmp)
    sc.flatten()
    dev.upload(sc)
    rd = sc.render_desc()
    rd.sample_begin, rd.sample_end = 0, 4
    lo, hi = sc.nodes()
    n = 1 << 22
    rays = scenes.ray_batch(n, lo[0, :3], hi[0, :3])
    d_r, d_o = dev.malloc(rays.nbytes), dev.malloc(16 * n)
    dev.h2d(d_r, rays)
    dev.set_option("sort_rays", 0)
    print("node_thr refill | c3 random closest Mrays/s | c3 path 4spp ms (closest/any/shade)")
    for nt in NT:
        for rt in RT:
            dev.set_option("node_threshold", nt)
            dev.set_option("refill_threshold", rt)
            dev.intersect_device(d_r, n, d_o)
            ms = min(dev.intersect_device(d_r, n, d_o) for _ in range(3))
            for sb in ((0, 1) if os.environ.get("RT_TRY_SORT_BOUNCE") else (0,)):
                dev.set_option("sort_bounce_rays", sb)
                dev.render(rd)
                dev.set_option("profile", 1)
                st = dev.render(rd)
                dev.set_option("profile", 0)
                st2 = dev.render(rd)
                print(f"{nt:8d} {rt:6d} | {n / ms / 1e3:9.1f} | {st.ms_total:7.2f} ({st.ms_closest:.2f}/{st.ms_anyhit:.2f}/{st.ms_shade:.2f})  sort_bounce_rays={sb} unprofiled {st2.ms_total:.2f} ms", flush=True)
            dev.set_option("sort_bounce_rays", 0)
    if a.big:
        sc = Scene.from_string(scenes.c4_scene(tmp), search_dir=tmp)
        sc.flatten()
        dev.upload(sc)
        lo, hi = sc.nodes()
        n = 1 << 24
        rays = scenes.ray_batch(n, lo[0, :3], hi[0, :3])
        dev.free(d_r), dev.free(d_o)
        d_r, d_o = dev.malloc(rays.nbytes), dev.malloc(16 * n)
        dev.h2d(d_r, rays)
        for sort in (0, 1):
            dev.set_option("sort_rays", sort)
            for nt in NT:
                for rt in (8, 16):
                    dev.set_option("node_threshold", nt)
                    dev.set_option("refill_threshold", rt)
                    dev.intersect_device(d_r, n, d_o)
                    ms = min(dev.intersect_device(d_r, n, d_o) for _ in range(3))
                    print(f"c4 sort={sort} node_thr={nt} refill={rt}: {n / ms / 1e3:.1f} Mrays/s", flush=True)


if __name__ == "__main__":
    main()
