"""Traversal parity + throughput probe (run on the GPU box): python tools/trav_bench.py [--big]"""
import argparse
import sys
import os
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device
from oracle import binding as ob


def parity(dev, sc, n=200000, label=""):
    lo, hi = sc.nodes()
    o = ob.OracleScene(sc.ir_ptr)
    for any_hit in (False, True):
        rays = scenes.ray_batch(n, lo[0, :3], hi[0, :3], any_hit=any_hit)
        if not any_hit:
            ref = o.intersect(rays)
            got = dev.intersect_stats(rays)
            idm = (ref["prim"] == got["prim"])
            tm = (ref["t"] == got["t"]) | ~idm
            print(f"[{label}] closest: n={n} hit%={100*(ref['prim']>=0).mean():.1f} id match={idm.mean()*100:.5f}% t bit-equal={tm.mean()*100:.5f}% "
                  f"b1/b2 equal={(np.array_equal(ref['b1'][idm], got['b1'][idm]) and np.array_equal(ref['b2'][idm], got['b2'][idm]))} "
                  f"nodes equal={np.array_equal(ref['nodes'], got['nodes'])} prims equal={np.array_equal(ref['prims'], got['prims'])} "
                  f"N={ref['nodes'].mean():.2f} T={ref['prims'].mean():.2f}")
            got2 = dev.intersect(rays)
            assert np.array_equal(got2["prim"], got["prim"]) and np.array_equal(got2["t"], got["t"])
        else:
            ref = o.occluded(rays)
            got = dev.occluded_stats(rays)
            print(f"[{label}] any-hit: occluded%={100*ref['occluded'].mean():.1f} match={(ref['occluded']==got['occluded']).mean()*100:.5f}% "
                  f"nodes equal={np.array_equal(ref['nodes'], got['nodes'])} N={ref['nodes'].mean():.2f} T={ref['prims'].mean():.2f}")


def throughput(dev, sc, n, label, reps=3):
    lo, hi = sc.nodes()
    for any_hit in (False, True):
        rays = scenes.ray_batch(n, lo[0, :3], hi[0, :3], any_hit=any_hit)
        d_r = dev.malloc(rays.nbytes)
        d_o = dev.malloc(16 * n)
        dev.h2d(d_r, rays)
        for simple in (1, 0):
            dev.set_option("simple_traversal", simple)
            for sort in (0, 1):
                dev.set_option("sort_rays", sort)
                f = dev.occluded_device if any_hit else dev.intersect_device
                f(d_r, n, d_o)
                ms = min(f(d_r, n, d_o) for _ in range(reps))
                print(f"[{label}] {'any' if any_hit else 'closest'} {'simple' if simple else 'engine'} sort={sort}: {n/ms/1e3:.1f} Mrays/s ({ms:.2f} ms for {n} rays)")
        dev.free(d_r), dev.free(d_o)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--level", type=int, default=5)
    a = ap.parse_args()
    dev = Device(0)
    tmp = tempfile.mkdtemp()
    for name, txt in (("cornell", scenes.cornell_box()), ("balls", scenes.balls())):
        sc = Scene.from_string(txt)
        dev.upload(sc)
        parity(dev, sc, label=name)
    t = time.time()
    sc = Scene.from_string(scenes.c3_scene(tmp, level=a.level), search_dir=tmp)
    sc.flatten()
    print(f"c3 level {a.level}: {sc.n_triangles} tris, parse+flatten {time.time()-t:.1f}s, bvh {sc.bvh_build_seconds:.2f}s")
    dev.upload(sc)
    parity(dev, sc, label="c3")
    throughput(dev, sc, 1 << 22, "c3")
    if a.big:
        t = time.time()
        sc = Scene.from_string(scenes.c4_scene(tmp), search_dir=tmp)
        sc.flatten()
        print(f"c4: {sc.n_triangles} tris, parse+flatten {time.time()-t:.1f}s, bvh {sc.bvh_build_seconds:.2f}s")
        dev.upload(sc)
        parity(dev, sc, n=1 << 20, label="c4")
        throughput(dev, sc, 1 << 24, "c4")


if __name__ == "__main__":
    main()
