#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wavefront path tracer (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: rustracer's algorithm on the host cores

Headline workload, the same for every N (config.workload = "c5_path"): SURVEY 8d config C5 — 4,997,120-triangle icosphere field with the
five materials cycling, four area lights and an infinite light, PathIntegrator maxdepth 5, lightsamplestrategy "spatial", 3840x2160,
1024 spp job.  One step = `--spp-per-step` (32) sample indices of every pixel of the 4K frame = 265.4 M camera paths, a FIXED amount of
work whatever N is: with N GPUs the scene is replicated and the 16x16 tiles are dealt round-robin over the ranks (tile_rank / tile_world),
i.e. strong scaling.  The job's single collective is one NCCL reduce of the film into rank 0 (rtgpu_reduce_film_nccl).

  value  = camera path samples of the K steps / device time (CUDA events inside rtgpu_render, max over ranks, + the film reduce)
  e2e    = the same job through the C ABI with host buffers: every step ends with the film reduced to rank 0 and read back
           (X, Y, Z, weight per pixel, 132.7 MB) into pinned host memory; wall clock between barriers
  N > 1  : the reduced film is compared with the same job rendered by rank 0 alone (`film_check`)

At N = 1 the line also carries `legs` — the other halves of BASELINE.json's metric on their own configs: c3_path (the 1 M-triangle
scene the >= 100x north-star target is quoted on, with its own CPU arm), c3_ao (AmbientOcclusion, 64 samples) and c4 (64 M incoherent
rays against 10,014,720 triangles, closest and any hit, device-resident and host-buffer rates, every id / t compared with the oracle).
"""
import argparse
import ctypes as C
import json
import os
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRICS = {"c5_path": "path samples/s (C5: 4K frame of the 5M-triangle mixed-material field, PathIntegrator spatial, tiles over 1/2/4/8 B200; Mrays/s closest-hit+shadow alongside)",
           "c3_path": "path samples/s (C3: 1M-triangle field, PathIntegrator spatial, 1920x1080; Mrays/s closest-hit+shadow alongside)"}
METRIC = METRICS["c5_path"]
UNIT = "samples/s"

_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Native libraries write to fd 1 behind Python's back (NCCL prints its version
    banner there whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole run and the line goes to a private
    duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_path", choices=["c5_path", "c3_path"], help="c3_path: the leg's config as the main workload (used for its CPU arm)")
    ap.add_argument("--spp-per-step", type=int, default=0, help="0 = the workload's default (c5_path 32, c3_path 8)")
    ap.add_argument("--level", type=int, default=5, help="icosphere subdivision level (5 = the named configs)")
    ap.add_argument("--xres", type=int, default=0)
    ap.add_argument("--yres", type=int, default=0)
    ap.add_argument("--legs", default="auto", help="auto (all legs at N = 1), none, or a comma list of c3_path,c3_ao,c4")
    ap.add_argument("--c4-rays", type=int, default=64 << 20)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of a cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    dflt = {"c5_path": (3840, 2160, 32), "c3_path": (1920, 1080, 8)}[a.workload]
    a.xres, a.yres = a.xres or dflt[0], a.yres or dflt[1]
    a.spp_per_step = a.spp_per_step or dflt[2]
    global METRIC
    METRIC = METRICS[a.workload]
    return a


def build_scene(a, tmp, workload=None):
    from rustracer_b200 import Scene, scenes
    workload = workload or a.workload
    if workload == "c5_path":
        txt = scenes.c5_scene(tmp, level=a.level, xres=a.xres, yres=a.yres, spp=1024)
    else:
        txt = scenes.c3_scene(tmp, level=a.level, xres=a.xres, yres=a.yres, spp=256)
    return Scene.from_string(txt, search_dir=tmp)


def workload_config(a, sc=None, world=1):
    if a.workload == "c5_path":
        cfg = {"workload": "c5_path", "scene": f"244 icospheres level {a.level}, matte / plastic / metal / glass / mirror cycling, ground, 4 area lights + infinite light",
               "job_spp": 1024}
    else:
        cfg = {"workload": "c3_path", "scene": f"49 icospheres level {a.level} + ground + area light + infinite light", "job_spp": 256}
    cfg.update({"integrator": "path maxdepth 5 lightsamplestrategy spatial", "resolution": [a.xres, a.yres], "spp_per_step": a.spp_per_step,
                "camera_paths_per_step": a.xres * a.yres * a.spp_per_step,
                "partition": f"16x16 tiles dealt round-robin over {world} rank(s); scene replicated; one NCCL reduce of the film",
                "sampler": "02sequence (counter-based (0,2) twin on the device)",
                "l2_policy": "inputs larger than L2: every step streams > 10 GB of wavefront queues through HBM besides the scene; no explicit flush"})
    if sc is not None:
        cfg["triangles"] = count_triangles(sc)
    return cfg


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                reasons = set()
                for r in rows:
                    for k, nm in enumerate(names):
                        if "Active" == r[5 + k].strip() and "Not" not in r[5 + k]:
                            reasons.add(nm)
                out["reasons"] = sorted(reasons)
                out["samples"] = len(sm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ---------------------------------------------------------------------------------------------------------------------------------------
# reference arm: the C++ restatement of rustracer's renderer on the host cores

def count_triangles(sc):
    ir = sc.ir
    return int(sum(ir.shapes[i].n_indices // 3 for i in range(ir.n_shapes)))


def oracle_rate(o, a, seconds=15.0, threads=0):
    """The oracle renderer (ZeroTwoSequence sampler, 16x16 tiles from a shared counter, all host threads) on a bounded sample of the
    workload: every `stride`-th tile of the frame, each rendered the way renderer::render does it (renderer.rs:83-130) — ALL of the job's
    samples of a pixel before the next pixel, one film merge under the mutex per tile.
    o: the OracleScene (its SAH BVH is built once)."""
    from rustracer_b200 import _abi as A
    job_spp = workload_config(a)["job_spp"]
    samp = A.rt_sampler(spp=job_spp, dimensions=4)
    n_tiles = ((a.xres + 15) // 16) * ((a.yres + 15) // 16)
    # calibrate on ~8 tiles per thread at an eighth of the samples, then size the real sample for ~`seconds` (at least 8 tiles per thread)
    cal = A.rt_sampler(spp=max(1, job_spp // 8), dimensions=4)
    _, _, st = o.render(sampler=cal, sampler_kind=0, threads=threads, tile_stride=max(1, n_tiles // (8 * max(1, os.cpu_count() or 1))))
    rate = st.camera_rays / max(st.seconds_tiles, 1e-6)
    want_tiles = max(8 * int(st.threads), int(rate * seconds / (256.0 * job_spp)))
    stride = max(1, n_tiles // want_tiles)
    ru0 = resource.getrusage(resource.RUSAGE_SELF)
    _, _, st = o.render(sampler=samp, sampler_kind=0, threads=threads, tile_stride=stride)
    ru1 = resource.getrusage(resource.RUSAGE_SELF)
    cpu_user, cpu_sys = ru1.ru_utime - ru0.ru_utime, ru1.ru_stime - ru0.ru_stime
    return {"value": st.camera_rays / st.seconds_tiles, "unit": UNIT, "cores": int(st.threads), "kind": "port",
            "sys_frac": cpu_sys / max(cpu_user + cpu_sys, 1e-9),
            "sample": f"every {stride}-th 16x16 tile of the {a.xres}x{a.yres} frame at the job's {job_spp} spp (the reference's loop order: all samples of a tile, then the "
                      f"next tile): {st.camera_rays} camera paths in {st.seconds_tiles:.2f} s (C++ restatement of rustracer's renderer, -O3 -march=native, ZeroTwoSequence "
                      f"sampler; the Rust binary cannot be built here)",
            "mrays_per_s": (st.regular_rays + st.shadow_rays) / st.seconds_tiles / 1e6, "seconds": st.seconds_tiles, "camera_rays": int(st.camera_rays)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    tmp = tempfile.mkdtemp(prefix="rtb200_")
    sc = build_scene(a, tmp)
    ob.build(native=True)
    t0 = time.perf_counter()
    o = ob.OracleScene(sc.ir_ptr, native=True)                      # rustracer's single-threaded SAH build (bvh/mod.rs:137-312), outside the render timer as in api.rs:995-1002
    t_build = time.perf_counter() - t0
    # each step = one bounded sample; K steps sized to finish within minutes
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    vals = []
    last = None
    for i in range(a.warmup + a.steps):
        r = oracle_rate(o, a, seconds=per_step)
        if i >= a.warmup:
            vals.append(r)
        last = r
    rates = [r["value"] for r in vals] if vals else [last["value"]]
    # value = the BEST of the K bounded samples (spread alongside): the CPU arm is the denominator of the speed-up, so its fastest run is the
    # conservative one.  Until round 2 the restated SpatialLightDistribution took a mutex on every lookup where the reference's hash is lock-free;
    # with 16 threads that lock made the same sample run at 2.3 M or at 5.5 M samples/s from one process to the next, the slow ones with most of
    # their CPU time in the kernel (profiles/r03i_bench.json).  The lookup is lock-free now (oracle/orc_render.hpp); as a safety net a measurement
    # whose kernel share of the CPU time is above 15 % is still repeated once in a fresh process and the better one kept.
    v = float(np.max(rates))
    best = (vals or [last])[int(np.argmax(rates))]
    secs = best["seconds"]
    sys_frac = float(np.median([r["sys_frac"] for r in (vals or [last])]))
    retried = None
    if sys_frac > 0.15 and not os.environ.get("RT_CPU_ARM_RETRY"):
        cmd = [sys.executable, os.path.abspath(__file__)] + sys.argv[1:]
        env = dict(os.environ, RT_CPU_ARM_RETRY="1")
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env).stdout.strip().splitlines()[-1]
            retried = json.loads(out)
        except Exception as e:
            log("retry of the disturbed CPU arm failed:", e)
        if retried and retried.get("value", 0.0) > v:
            retried["cpu_baseline"]["first_attempt"] = {"value": v, "sys_frac": sys_frac}
            emit(retried)
            return
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, sc),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": best["sample"] + "; value = best of the K samples",
                             "spread": {"min": float(np.min(rates)), "max": float(np.max(rates)), "mean": float(np.mean(rates)), "n": len(rates)},
                             "sys_frac": sys_frac, "retried": retried is not None},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mrays_per_s": best["mrays_per_s"], "gpu_launches": 0,
            "scene_build_seconds": t_build}
    emit(line)


def cpu_baseline_subprocess(a, workload, timeout=900):
    """The CPU arm in a fresh process: measured inside this one (CUDA context, torch thread pools, the clock sampler) the same CPU sample
    ran 2x slower (profiles/r01h).  One untimed pass first: a cold first pass was 2x slower once (profiles/r01zc_bench.json)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1", "--level", str(a.level),
           "--cpu-seconds", str(a.cpu_seconds)]
    if workload == a.workload:
        cmd += ["--xres", str(a.xres), "--yres", str(a.yres), "--spp-per-step", str(a.spp_per_step)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env).stdout.strip().splitlines()[-1]
    return json.loads(out)["cpu_baseline"]


# ---------------------------------------------------------------------------------------------------------------------------------------
# roofline of the traversal kernels, measured live

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return float(d.get("hbm_gbs", 6650.0)), ("MEASURED_PEAKS.json hbm_gbs" if d else "fallback 6650 GB/s (B200_PROFILING.md)")


def traffic_entry(key):
    """Measured DRAM / L2 bytes per launch of a kernel class from the committed ncu --set full captures (profiles/roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return {}
    return json.load(open(p)).get(key, {})


# algorithmic bytes of one shaded path vertex: read hit 16 + ray 32 + throughput 16 + state 16 + geometry 48 + primitive
# info 16 + sampler key 8 + radiance 16; written radiance 16 + next ray 32 + throughput 16 + state 16 + list entry 4 + shadow-queue entry 48 +
# MIS-queue entry 48
SHADE_BYTES_PER_VERTEX = 16 + 32 + 16 + 16 + 48 + 16 + 8 + 16 + 16 + 32 + 16 + 16 + 4 + 48 + 48


def render_roofline(dev, step_desc, reps, traffic_key, full_defaults):
    """Counting pass (exact N and T of the step's rays) + `reps` profiled passes (CUDA events around every launch, per kernel class)."""
    peak, peak_src = peaks()
    dev.set_option("profile", 1)
    dev.set_option("count_traversal", 1)
    stc = dev.render(step_desc())                                   # not timed
    dev.set_option("count_traversal", 0)
    acc = {"closest": 0.0, "anyhit": 0.0, "shade": 0.0, "other": 0.0, "total": 0.0, "launches": 0}
    for _ in range(reps):
        stp = dev.render(step_desc())
        acc["closest"] += stp.ms_closest
        acc["anyhit"] += stp.ms_anyhit
        acc["shade"] += stp.ms_shade
        acc["other"] += stp.ms_other
        acc["total"] += stp.ms_total
        acc["launches"] += stp.closest_launches
    dev.set_option("profile", 0)
    n_rays = stc.closest_rays
    bytes_step = 48.0 * n_rays + 32.0 * stc.nodes_closest + 48.0 * stc.prims_closest          # 32 B ray + 16 B hit + 32 N + 48 T
    bytes_any = 33.0 * stc.anyhit_rays + 32.0 * stc.nodes_anyhit + 48.0 * stc.prims_anyhit     # 32 B ray + 1 B flag + 32 N + 48 T
    t_closest, t_any, t_shade = acc["closest"] / reps * 1e-3, acc["anyhit"] / reps * 1e-3, acc["shade"] / reps * 1e-3
    achieved = bytes_step / t_closest / 1e9
    tr = traffic_entry(traffic_key) if full_defaults else {}
    launches_per_step = acc["launches"] / reps
    roof = {"bound": "hbm", "limited_by": tr.get("limited_by"), "kernel": "k_trace_closest_engine + k_trace_mis_engine (closest-hit BVH traversal)", "achieved": achieved, "peak": peak,
            "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr.get("dram_bytes_per_launch"), "l2_bytes_per_launch": tr.get("l2_bytes_per_launch"), "l2_frac": tr.get("l2_frac"),
            "l1_frac": tr.get("l1_frac"), "dram_frac": tr.get("dram_frac"), "traffic_source": tr.get("source"),
            "algorithmic_bytes_per_launch": bytes_step / max(1.0, launches_per_step), "algorithmic_bytes_per_step": bytes_step, "rays_per_step": int(n_rays),
            "nodes_per_ray": stc.nodes_closest / max(1, n_rays), "prims_per_ray": stc.prims_closest / max(1, n_rays),
            "launches_per_step": launches_per_step, "avg_launch_ms": acc["closest"] / max(1, acc["launches"]),
            "share_of_step": {k: acc[k] / acc["total"] for k in ("closest", "anyhit", "shade", "other")},
            "note": tr.get("note", "achieved counts algorithmic bytes (SURVEY 8d); the tree is partly L2-resident, so it can exceed what DRAM delivers")}
    roof_any = {"bound": "hbm", "kernel": "k_trace_shadow_engine (any-hit BVH traversal: NEE shadow rays + MIS rays towards infinite lights)",
                "achieved": bytes_any / max(t_any, 1e-9) / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_any / max(t_any, 1e-9) / 1e9 / peak,
                "algorithmic_bytes_per_step": bytes_any, "rays_per_step": int(stc.anyhit_rays),
                "nodes_per_ray": stc.nodes_anyhit / max(1, stc.anyhit_rays), "prims_per_ray": stc.prims_anyhit / max(1, stc.anyhit_rays)}
    vertices = int(stc.shaded_items)                              # every entry of a bounce's live list ends in one shade (or miss) visit
    shade_bytes = SHADE_BYTES_PER_VERTEX * float(vertices)
    roof_shade = {"bound": "hbm", "kernel": "k_classify + k_shade_path<material> + k_shade_miss (+ k_eval_textured, k_matsort_*) — everything between two traversals",
                  "achieved": shade_bytes / max(t_shade, 1e-9) / 1e9, "peak": peak, "unit": "GB/s", "frac": shade_bytes / max(t_shade, 1e-9) / 1e9 / peak,
                  "algorithmic_bytes_per_vertex": SHADE_BYTES_PER_VERTEX, "vertices_per_step": vertices, "ms_per_step": t_shade * 1e3}
    return roof, roof_any, roof_shade


# ---------------------------------------------------------------------------------------------------------------------------------------
# legs (N = 1): the other configs of the metric

def leg_c3_path(dev, a, tmp):
    """SURVEY 8d C3 / north star: the 1M-triangle scene, 1920x1080, 8 spp per step; device rate, end-to-end rate with the film read back to
    pinned host memory every step, roofline, and the CPU arm on the same config."""
    import torch
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.c3_scene(tmp, level=a.level, xres=1920, yres=1080, spp=256), search_dir=tmp)
    dev.upload(sc)
    rd = sc.render_desc()
    rd.seed = 1
    spp, steps, warm = 8, 8, 3

    def desc(k, clear=False):
        rd.sample_begin, rd.sample_end = (k * spp) % rd.spp, (k * spp) % rd.spp + spp
        rd.clear_film = 1 if clear else 0
        return rd
    film_host = torch.empty((1080, 1920, 4), dtype=torch.float32, pin_memory=True)
    for k in range(warm):
        dev.render(desc(k, k == 0))
        dev.read_film(out=film_host)
    ms = cam = reg = sh = 0
    for k in range(steps):
        st = dev.render(desc(warm + k))
        ms += st.ms_total
        cam += st.camera_rays
        reg += st.regular_rays
        sh += st.shadow_rays
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        dev.render(desc(warm + k))
        dev.read_film(out=film_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    out = {"workload": "c3_path", "triangles": count_triangles(sc), "resolution": [1920, 1080], "spp_per_step": spp, "steps": steps,
           "value": cam / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "mrays_per_s": (reg + sh) / (ms * 1e-3) / 1e6,
           "e2e": {"value": cam / e2e_s, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(type(rd)), "d2h_bytes_per_step": int(film_host.numel() * 4)}}
    roof, roof_any, roof_shade = render_roofline(dev, lambda: desc(warm), 4, "c3_path_closest", a.level == 5)
    out["roofline"], out["roofline_anyhit"], out["roofline_shade"] = roof, roof_any, roof_shade
    if not a.no_cpu_baseline:
        try:
            cb = cpu_baseline_subprocess(a, "c3_path")
            out["cpu_baseline"] = cb
            out["vs_cpu"] = {"device": out["value"] / cb["value"], "e2e": out["e2e"]["value"] / cb["value"], "north_star_target": 100.0}
        except Exception as e:
            out["cpu_baseline"] = {"error": str(e)}
    return out


def leg_c3_ao(dev, a, tmp):
    """SURVEY 8d C3, AmbientOcclusion with 64 samples (ao.rs:32-58): 1 camera ray + 64 any-hit rays per hit sample, 1 spp per step."""
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.c3_scene(tmp, level=a.level, xres=1920, yres=1080, spp=256, integrator='Integrator "ambientocclusion" "integer nsamples" [64]'),
                           search_dir=tmp)
    dev.upload(sc)
    rd = sc.render_desc()
    rd.seed = 1
    steps, warm = 6, 2

    def desc(k, clear=False):
        rd.sample_begin, rd.sample_end = k % rd.spp, k % rd.spp + 1
        rd.clear_film = 1 if clear else 0
        return rd
    for k in range(warm):
        dev.render(desc(k, k == 0))
    ms = cam = reg = sh = 0
    for k in range(steps):
        st = dev.render(desc(warm + k))
        ms += st.ms_total
        cam += st.camera_rays
        reg += st.regular_rays
        sh += st.shadow_rays
    return {"workload": "c3_ao", "triangles": count_triangles(sc), "resolution": [1920, 1080], "ao_samples": 64, "spp_per_step": 1, "steps": steps,
            "value": cam / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "mrays_per_s": (reg + sh) / (ms * 1e-3) / 1e6,
            "rays_per_step": {"camera": cam / steps, "shadow": sh / steps}}


def leg_c4(dev, a, tmp):
    """64 M incoherent rays against the 10,014,720-triangle field, closest and any hit.  Device-resident rate (CUDA events
    around sort + traversal), host-buffer rate through rtgpu_intersect / rtgpu_occluded (pinned buffers, copies inside), roofline from
    exact per-ray node / primitive counts on a 4 M-ray subset, and EVERY ray's id / t / occlusion flag compared with the oracle."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, host, scenes
    n = int(a.c4_rays)
    t0 = time.perf_counter()
    sc = Scene.from_string(scenes.c4_scene(tmp, level=a.level), search_dir=tmp)
    dev.upload(sc)
    t_scene = time.perf_counter() - t0
    lo, hi = sc.nodes()
    wlo, whi = lo[0, :3], hi[0, :3]
    out = {"workload": "c4_ray_batch", "triangles": count_triangles(sc), "rays": n, "scene_seconds": t_scene}
    rays = dev.pinned_empty((n, 8), np.float32)
    hits = dev.pinned_empty((n, 4), np.float32)
    occ = dev.pinned_empty((n,), np.uint8)
    peak, _ = peaks()
    chunk = 1 << 24
    o = None
    for kind in ("closest", "anyhit"):
        any_hit = kind == "anyhit"
        host.ray_batch(n, wlo, whi, seed=5, any_hit=any_hit, out=rays)
        # device-resident: 16 Mi-ray launches (sort + traversal inside the event pair), after one warm-up launch
        d_r, d_o = dev.malloc(chunk * 32), dev.malloc(chunk * 16)
        ms = 0.0
        for rep in range(2):
            ms = 0.0
            for first in range(0, n, chunk):
                m = min(chunk, n - first)
                dev.h2d(d_r, rays[first:first + m])
                ms += dev.occluded_device(d_r, m, d_o) if any_hit else dev.intersect_device(d_r, m, d_o)
        # N and T on the first 4 Mi rays
        m = min(n, 1 << 22)
        dev.h2d(d_r, rays[:m])
        d_s = dev.malloc(8 * m)
        lib = dev._lib
        if any_hit:
            dev._check(lib.rtgpu_occluded_device_stats(dev._h, d_r, m, d_o, d_s))
        else:
            dev._check(lib.rtgpu_intersect_device_stats(dev._h, d_r, m, d_o, d_s))
        stt = np.zeros((m, 2), np.uint32)
        dev.d2h(stt, d_s)
        for p in (d_r, d_o, d_s):
            dev.free(p)
        nodes, prims = float(stt[:, 0].mean()), float(stt[:, 1].mean())
        bytes_ray = (33.0 if any_hit else 48.0) + 32.0 * nodes + 48.0 * prims
        # host buffers through the plugin call: twice, the second timed
        for rep in range(2):
            t0 = time.perf_counter()
            if any_hit:
                dev.occluded(rays, out=occ)
            else:
                dev.intersect(rays, out=hits)
            t_host = time.perf_counter() - t0
        tr = traffic_entry("c4_" + kind) if a.level == 5 else {}
        res = {"mrays_per_s_device": n / (ms * 1e-3) / 1e6, "ms_device": ms, "mrays_per_s_host_buffers": n / t_host / 1e6, "seconds_host_buffers": t_host,
               "h2d_bytes": n * 32, "d2h_bytes": n * (1 if any_hit else 16), "host_memory": "pinned (rtgpu_host_alloc)",
               "nodes_per_ray": nodes, "prims_per_ray": prims, "algorithmic_bytes_per_ray": bytes_ray,
               "roofline": {"bound": "hbm", "limited_by": tr.get("limited_by"), "kernel": "k_anyhit_batch_engine" if any_hit else "k_closest_batch_engine",
                            "achieved": n * bytes_ray / (ms * 1e-3) / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": n * bytes_ray / (ms * 1e-3) / 1e9 / peak, "traffic": tr.get("dram_bytes_per_launch"),
                            "algorithmic_bytes_per_launch": (1 << 24) * bytes_ray, "l2_bytes_per_launch": tr.get("l2_bytes_per_launch"), "l2_frac": tr.get("l2_frac"),
                            "l1_frac": tr.get("l1_frac"), "dram_frac": tr.get("dram_frac"), "traffic_source": tr.get("source"),
                            "note": "includes the ray binning (k_sort_*) inside the timed launches"}}
        # every ray against the oracle (all host threads)
        if not a.no_cpu_baseline:
            if o is None:
                t0 = time.perf_counter()
                o = ob.OracleScene(sc.ir_ptr, native=True)
                out["oracle_build_seconds"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            if any_hit:
                ref = o.occluded(rays)
                equal = int((ref["occluded"] == occ).sum())
                res["check"] = {"rays": n, "equal": equal, "pct_equal": 100.0 * equal / n}
            else:
                ref = o.intersect(rays)
                prim = hits[:, 1].view(np.int32)
                same_id = ref["prim"] == prim
                t_ref, t_got = ref["t"], hits[:, 0]
                hit = ref["prim"] >= 0
                with np.errstate(invalid="ignore"):
                    t_ok = np.where(hit, np.abs(t_got - t_ref) <= 1e-5 * np.abs(t_ref), prim < 0)
                res["check"] = {"rays": n, "ids_equal": int(same_id.sum()), "pct_ids_equal": 100.0 * float(same_id.mean()),
                                "t_bit_equal": int((t_ref[hit] == t_got[hit]).sum()), "pct_t_within_1e-5": 100.0 * float(t_ok.mean()), "hits": int(hit.sum())}
            res["check"]["oracle_seconds"] = time.perf_counter() - t0
            res["check"]["oracle_mrays_per_s"] = n / res["check"]["oracle_seconds"] / 1e6
        out[kind] = res
    return out


# ---------------------------------------------------------------------------------------------------------------------------------------

def run_ours(a):
    import torch
    from rustracer_b200.device import Device
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GPU path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line (NCCL prints its version banner to stdout)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix=f"rtb200_{rank}_")
    t0 = time.perf_counter()
    sc = build_scene(a, tmp)
    dev = Device(local).upload(sc)                                   # flatten with the SAH BVH built on this device
    t_scene = time.perf_counter() - t0
    for kv in filter(None, os.environ.get("RT_OPTIONS", "").split(",")):        # A/B runs: RT_OPTIONS=overlap_bounces=0,node_threshold=12
        dev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if world > 1:
        # the film reduce runs inside the C ABI (rtgpu_reduce_film_nccl); torch.distributed only ships the NCCL id and keeps the barrier
        ident = [Device.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        dev.comm_init(ident[0], rank, world)
    rd = sc.render_desc()
    rd.seed = 1
    if os.environ.get("RT_WAVE_PATHS"):                                 # A/B runs: paths per wave (0 = the library's default)
        rd.wave_paths = int(os.environ["RT_WAVE_PATHS"])
    spp, job_spp = a.spp_per_step, rd.spp
    rd.tile_rank, rd.tile_world = rank, world

    def step_desc(k, clear, solo=False):
        # step k -> sample indices [k * spp, +spp) of the job (wrapping past job_spp keeps per-step work fixed); my tiles of the frame
        s0 = (k * spp) % job_spp
        rd.sample_begin, rd.sample_end = s0, min(job_spp, s0 + spp)
        rd.clear_film = 1 if clear else 0
        rd.tile_rank, rd.tile_world = (0, 1) if solo else (rank, world)
        return rd

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fh, fw = rd.cropped[3] - rd.cropped[1], rd.cropped[2] - rd.cropped[0]
    film_host = torch.empty((fh, fw, 4), dtype=torch.float32, pin_memory=True) if rank == 0 else None

    # ---- device-resident timing: K steps, CUDA events inside rtgpu_render (stats.ms_total), one film reduce; max over ranks -------
    for k in range(a.warmup):
        dev.render(step_desc(k, k == 0))
        if world > 1:
            dev.reduce_film(0)
        if rank == 0:
            dev.read_film(out=film_host)             # also warms the read-back path (staging buffer, lazily loaded kernel)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = dev.launch_count
    t_wall = time.perf_counter()
    ms, cam, reg, sh, ms_reduce = 0.0, 0, 0, 0, 0.0
    for k in range(a.steps):
        st = dev.render(step_desc(a.warmup + k, k == 0))
        ms += st.ms_total
        cam += st.camera_rays
        reg += st.regular_rays
        sh += st.shadow_rays
    if world > 1:                            # the job's single film reduce over NVLink, inside the timed region
        ms_reduce = dev.reduce_film(0)
        ms += ms_reduce
    barrier()
    wall = time.perf_counter() - t_wall
    launches = dev.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None

    film_check = None
    if world > 1:
        # reduced film of the K-step job against the same job rendered by rank 0 alone (not timed)
        if rank == 0:
            reduced = dev.read_film().copy()
            t0 = time.perf_counter()
            ms1 = 0.0
            for k in range(a.steps):
                ms1 += dev.render(step_desc(a.warmup + k, k == 0, solo=True)).ms_total
            solo = dev.read_film()
            w_equal = bool(np.array_equal(solo[..., 3], reduced[..., 3]))
            den = np.maximum(np.abs(solo[..., :3]), 1e-3 * float(np.abs(solo[..., :3]).mean()))
            rel = np.abs(reduced[..., :3] - solo[..., :3]) / den
            film_check = {"weights_equal": w_equal, "max_rel_err": float(rel.max()), "mean_rel_err": float(rel.mean()),
                          "rel_l1": float(np.abs(reduced[..., :3] - solo[..., :3]).sum() / np.abs(solo[..., :3]).sum()),
                          "tolerance": 1e-5, "ok": bool(w_equal and rel.max() <= 1e-5),
                          "single_gpu_ms_same_box": ms1,
                          "note": "film of the N-rank job after the reduce vs the same K steps rendered by rank 0 alone; a pixel's samples are summed with float atomics, "
                                  "so the comparison tolerates fp32 summation order (relative, floored at 0.1 % of the mean value)"}
        barrier()

    # ---- end-to-end through the C ABI with host buffers: every step renders, reduces the film into rank 0 and reads it back ----
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        # rank 0 keeps accumulating; the other ranks start every step from an empty film, so the per-step reduce adds each sample once
        dev.render(step_desc(a.warmup + k, (k == 0) if rank == 0 else True))
        if world > 1:
            dev.reduce_film(0)
        if rank == 0:
            dev.read_film(out=film_host)                 # X,Y,Z,weight into pinned host memory (rtgpu_read_film)
    barrier()
    e2e_s = time.perf_counter() - t0

    stats = torch.tensor([ms, float(cam), float(reg), float(sh), e2e_s, float(launches)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[4])
        cam, reg, sh, launches = float(sm[1]), float(sm[2]), float(sm[3]), float(sm[5])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = cam / (ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, sc, world), "mrays_per_s": (reg + sh) / (ms * 1e-3) / 1e6,
            "rays": {"camera": cam, "regular": reg, "shadow": sh}, "wall_s_timed_region": wall, "scene_seconds": t_scene,
            "e2e": {"value": cam / e2e_s, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(type(rd)), "d2h_bytes_per_step": int(fh * fw * 16),
                    "what": "per step: rtgpu_render of the rank's tiles, rtgpu_reduce_film_nccl into rank 0 (N > 1), rtgpu_read_film into pinned host memory on rank 0"},
            "gpu_launches": int(launches), "clocks": clk}
    if world > 1:
        line["reduce_ms"] = ms_reduce
        if film_check is not None:
            film_check["speedup_vs_single_gpu_same_box"] = film_check["single_gpu_ms_same_box"] / ms
        line["film_check"] = film_check

    # ---- roofline of the dominant kernel class (closest-hit traversal), measured live with CUDA events on the context's stream ----
    if world == 1:
        try:
            full = a.level == 5 and (a.xres, a.yres) == ((3840, 2160) if a.workload == "c5_path" else (1920, 1080))
            roof, roof_any, roof_shade = render_roofline(dev, lambda: step_desc(a.warmup, False), min(a.steps, 2), a.workload + "_closest", full)
            line["roofline"], line["roofline_anyhit"], line["roofline_shade"] = roof, roof_any, roof_shade
        except Exception as e:  # the headline number must not depend on the diagnostics
            line["roofline"] = {"error": str(e)}
        if not a.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_subprocess(a, a.workload)
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)}
        legs = [] if a.legs == "none" else (["c3_path", "c3_ao", "c4"] if a.legs == "auto" else a.legs.split(","))
        if a.workload != "c5_path":
            legs = [x for x in legs if x != "c3_path"] if a.legs == "auto" else legs
        line["legs"] = {}
        for name in legs:
            t0 = time.perf_counter()
            try:
                fn = {"c3_path": leg_c3_path, "c3_ao": leg_c3_ao, "c4": leg_c4}[name]
                line["legs"][name] = fn(dev, a, tmp)
            except Exception as e:
                line["legs"][name] = {"error": repr(e)}
            line["legs"][name]["leg_seconds"] = time.perf_counter() - t0
            log(name, "done in", round(time.perf_counter() - t0, 1), "s")
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse_args()
    claim_stdout()
    import __graft_entry__ as g
    g.ensure_built()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
