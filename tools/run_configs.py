"""BASELINE.json configs C1..C5 on the GPU box, each next to the CPU oracle on a bounded sample.

    python tools/run_configs.py [--configs c1,c2,c3,c4,c5] [--out gpurun_out/configs.json]

For every config: the full-size GPU run (all spp unless noted), device time from rtgpu_render's CUDA events, the
reference's ray counters, and the oracle (C++ restatement, all host threads, reference sampler) on every k-th tile.
C4 is the 64 M-ray microbench against the 10 M-triangle BVH with the bit-exact id check on a 1 M-ray subset."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob
from rustracer_b200 import Scene, scenes, _abi as A
from rustracer_b200.device import Device


def cpu_rate(sc, spp, seconds=8.0, native=True):
    ob.build(native=native)
    o = ob.OracleScene(sc.ir_ptr, native=native)
    samp = A.rt_sampler(spp=spp, dimensions=4)
    _, _, st = o.render(sampler=samp, sampler_kind=0, tile_stride=97)
    rate = st.camera_rays / max(st.seconds_tiles, 1e-6)
    rd = sc.render_desc()
    total = (rd.sample_bounds[2] - rd.sample_bounds[0]) * (rd.sample_bounds[3] - rd.sample_bounds[1]) * spp
    stride = max(1, int(np.ceil(total / max(rate * seconds, 1.0))))
    _, _, st = o.render(sampler=samp, sampler_kind=0, tile_stride=stride)
    return dict(samples_per_s=st.camera_rays / st.seconds_tiles, mrays_per_s=(st.regular_rays + st.shadow_rays) / st.seconds_tiles / 1e6,
                threads=int(st.threads), tile_stride=stride, spp=spp, seconds=st.seconds_tiles, camera_rays=int(st.camera_rays))


def gpu_render(dev, sc, spp_limit=None, label=""):
    rd = sc.render_desc()
    if spp_limit is not None:
        rd.sample_end = min(rd.spp, spp_limit)
    dev.render(rd)                                   # warm-up (allocations, light grid)
    st = dev.render(rd)
    s = st.ms_total * 1e-3
    gpu_render.last_stats = st
    return dict(label=label, ms=st.ms_total, spp_rendered=int(rd.sample_end - rd.sample_begin), spp_config=int(rd.spp), camera=int(st.camera_rays),
                regular=int(st.regular_rays), shadow=int(st.shadow_rays), samples_per_s=st.camera_rays / s,
                mrays_per_s=(st.regular_rays + st.shadow_rays) / s / 1e6, waves=int(st.waves), launches=int(st.kernel_launches))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--c4-rays", type=int, default=1 << 26)
    a = ap.parse_args()
    want = a.configs.split(",")
    dev = Device(0)
    for kv in filter(None, os.environ.get("RT_OPTIONS", "").split(",")):        # A/B runs: RT_OPTIONS=overlap_bounces=0
        dev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    tmp = tempfile.mkdtemp()
    res = {}

    def record(name, value):
        res[name] = value
        print(name, json.dumps(value), flush=True)
        json.dump(res, open(a.out, "w"), indent=1)

    if "c1" in want:      # Cornell box, path uniform, 512x512, 16 spp, maxdepth 5
        sc = Scene.from_string(scenes.cornell_box())
        dev.upload(sc)
        g = gpu_render(dev, sc, label="C1 cornell path uniform 512x512 16spp")
        c = cpu_rate(sc, 16)
        record("c1", dict(gpu=g, cpu=c, speedup=g["samples_per_s"] / c["samples_per_s"]))
    if "c2" in want:      # Balls: 64 spheres + disk, Whitted and DirectLighting all, 1024x768, 64 spp
        for name, integ in (("whitted", 'Integrator "whitted" "integer maxdepth" [5]'), ("direct_all", 'Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]')):
            sc = Scene.from_string(scenes.balls(integrator=integ))
            dev.upload(sc)
            g = gpu_render(dev, sc, label=f"C2 balls {name} 1024x768 64spp")
            c = cpu_rate(sc, 64)
            record(f"c2_{name}", dict(gpu=g, cpu=c, speedup=g["samples_per_s"] / c["samples_per_s"]))
    if "tex" in want:     # the textured balls scene (every texture class, bump maps), path and Whitted, 1024x768, 64 spp
        for name, integ in (("path", None), ("whitted", 'Integrator "whitted" "integer maxdepth" [5]')):
            sc = Scene.from_string(scenes.balls_textured(tmp, integrator=integ), search_dir=tmp)
            dev.upload(sc)
            dev.set_option("profile", 1)
            g = gpu_render(dev, sc, label=f"textured balls {name} 1024x768 64spp")
            st = gpu_render.last_stats
            g.update(ms_closest=st.ms_closest, ms_anyhit=st.ms_anyhit, ms_shade=st.ms_shade, ms_other=st.ms_other)
            dev.set_option("profile", 0)
            g["unprofiled"] = gpu_render(dev, sc)
            c = cpu_rate(sc, 64)
            record(f"tex_{name}", dict(gpu=g, cpu=c, speedup=g["unprofiled"]["samples_per_s"] / c["samples_per_s"]))
    if "c3" in want:      # 1M-triangle field, AO (64 samples) and path spatial, 1920x1080, 256 spp
        for name, integ, spp_lim in (("path_spatial", None, 32), ("ao64", 'Integrator "ambientocclusion" "integer nsamples" [64]', 8)):
            sc = Scene.from_string(scenes.c3_scene(tmp, integrator=integ), search_dir=tmp)
            dev.upload(sc)
            g = gpu_render(dev, sc, spp_limit=spp_lim, label=f"C3 1M tris {name} 1920x1080 ({spp_lim} of 256 spp timed)")
            c = cpu_rate(sc, 4)
            record(f"c3_{name}", dict(gpu=g, cpu=c, speedup=g["samples_per_s"] / c["samples_per_s"]))
    if "c4" in want:      # 64M incoherent rays vs 10M-triangle SAH BVH
        t0 = time.time()
        sc = Scene.from_string(scenes.c4_scene(tmp), search_dir=tmp)
        sc.flatten()
        dev.upload(sc)
        lo, hi = sc.nodes()
        n = a.c4_rays
        out = dict(triangles=sc.n_triangles, bvh_build_s=sc.bvh_build_seconds, nodes=int(lo.shape[0]), rays=n)
        o = ob.OracleScene(sc.ir_ptr, native=True)
        for any_hit in (False, True):
            kind = "any" if any_hit else "closest"
            chunk = 1 << 24
            d_r, d_o = dev.malloc(32 * chunk), dev.malloc(16 * chunk)
            ms_dev, ms_e2e = 0.0, 0.0
            first_rays, first_res = None, None
            for first in range(0, n, chunk):
                m = min(chunk, n - first)
                rays = scenes.ray_batch(m, lo[0, :3], hi[0, :3], any_hit=any_hit, first=first)
                dev.h2d(d_r, rays)
                f = dev.occluded_device if any_hit else dev.intersect_device
                if first == 0:
                    f(d_r, m, d_o)
                ms_dev += f(d_r, m, d_o)
                t1 = time.perf_counter()
                r = dev.occluded(rays) if any_hit else dev.intersect(rays)       # host buffers: H2D + kernels + D2H
                ms_e2e += (time.perf_counter() - t1) * 1e3
                if first == 0:
                    first_rays, first_res = rays[: 1 << 20], r
            dev.free(d_r), dev.free(d_o)
            sub = first_rays
            t1 = time.time()
            ref = o.occluded(sub) if any_hit else o.intersect(sub)
            cpu_s = time.time() - t1
            if any_hit:
                match = float((ref["occluded"] == first_res[: len(sub)]).mean())
                Nn, Tt = float(ref["nodes"].mean()), float(ref["prims"].mean())
                bytes_ray = 33 + 32 * Nn + 48 * Tt
            else:
                match = float((ref["prim"] == first_res["prim"][: len(sub)]).mean())
                tmatch = float((ref["t"] == first_res["t"][: len(sub)]).mean())
                out["closest_t_bit_equal"] = tmatch
                Nn, Tt = float(ref["nodes"].mean()), float(ref["prims"].mean())
                bytes_ray = 48 + 32 * Nn + 48 * Tt
            out[kind] = dict(mrays_per_s_device=n / ms_dev / 1e3, ms_device=ms_dev, mrays_per_s_host_buffers=n / ms_e2e / 1e3, id_match_1M_subset=match,
                             nodes_per_ray=Nn, prims_per_ray=Tt, algorithmic_bytes_per_ray=bytes_ray, algorithmic_gbs=n / (ms_dev * 1e-3) * bytes_ray / 1e9,
                             cpu_mrays_per_s=len(sub) / cpu_s / 1e6)
        out["total_s"] = time.time() - t0
        record("c4", out)
    if "c5" in want:      # 4K, 5M triangles mixed materials, path spatial, 1024 spp (8 spp timed on one GPU here)
        sc = Scene.from_string(scenes.c5_scene(tmp), search_dir=tmp)
        dev.upload(sc)
        g = gpu_render(dev, sc, spp_limit=8, label="C5 5M tris mixed materials path spatial 3840x2160 (8 of 1024 spp timed)")
        c = cpu_rate(sc, 4, seconds=10.0)
        record("c5", dict(gpu=g, cpu=c, speedup=g["samples_per_s"] / c["samples_per_s"], triangles=sc.n_triangles))


if __name__ == "__main__":
    main()
