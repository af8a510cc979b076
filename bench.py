#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wavefront path tracer (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: rustracer's algorithm on the host cores

Workload (config.workload = "c3_path"): SURVEY 8d config C3 — synthetic 1,003,520-triangle subdivided-icosphere PLY
field + ground + 2-triangle area light + constant infinite light, PathIntegrator maxdepth 5, lightsamplestrategy
"spatial", 1920x1080, Sampler "02sequence" 256 spp.  One step = `--spp-per-step` sample indices (default 8) of every
pixel: 16.6 M camera paths.  Successive steps take successive sample-index ranges of the 256-spp job; with N GPUs rank r
takes the r-th range of each step (scene replicated, sample indices partitioned) and the films are summed with one NCCL
reduce at the end of the job.  metric = camera path samples per second (the reference's "Camera rays traced" / s).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "path samples/s (C3: 1M-triangle field, PathIntegrator spatial, 1920x1080; Mrays/s closest-hit+shadow alongside)"
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=8)
    ap.add_argument("--level", type=int, default=5, help="icosphere subdivision level (5 = the 1M-triangle config)")
    ap.add_argument("--xres", type=int, default=1920)
    ap.add_argument("--yres", type=int, default=1080)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_scene(a, tmp):
    from rustracer_b200 import Scene, scenes
    txt = scenes.c3_scene(tmp, level=a.level, xres=a.xres, yres=a.yres, spp=256)
    sc = Scene.from_string(txt, search_dir=tmp)
    return sc


def workload_config(a, sc=None):
    cfg = {"workload": "c3_path", "scene": f"49 icospheres level {a.level} + ground + area light + infinite light",
           "integrator": "path maxdepth 5 lightsamplestrategy spatial", "resolution": [a.xres, a.yres], "job_spp": 256,
           "spp_per_step": a.spp_per_step, "sampler": "02sequence (counter-based (0,2) twin on the device)",
           "l2_policy": "inputs larger than L2: each step streams ~2 GB of wavefront queues plus the 80 MB scene; no explicit flush"}
    if sc is not None:
        cfg["triangles"] = sc.n_triangles or (49 * 20 * 4 ** a.level + 4)
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


def oracle_rate(sc, a, native=True, seconds=15.0, threads=0):
    """Reference-arm measurement: the C++ restatement of rustracer's renderer (ZeroTwoSequence sampler, 16x16 tiles from a
    shared counter, all host threads) on a bounded sample of the workload: every `stride`-th tile at spp_per_step spp."""
    from oracle import binding as ob
    from rustracer_b200 import _abi as A
    ob.build(native=native)
    o = ob.OracleScene(sc.ir_ptr, native=native)
    samp = A.rt_sampler(spp=a.spp_per_step, dimensions=4)
    # calibrate on a sparse subset, then size the real sample for ~`seconds`
    _, _, st = o.render(sampler=samp, sampler_kind=0, threads=threads, tile_stride=64)
    rate = st.camera_rays / max(st.seconds_tiles, 1e-6)
    total = a.xres * a.yres * a.spp_per_step
    stride = max(1, int(np.ceil(total / max(rate * seconds, 1.0))))
    _, _, st = o.render(sampler=samp, sampler_kind=0, threads=threads, tile_stride=stride)
    return {"value": st.camera_rays / st.seconds_tiles, "unit": UNIT, "cores": int(st.threads), "kind": "port",
            "sample": f"every {stride}-th 16x16 tile of the {a.xres}x{a.yres} frame at {a.spp_per_step} spp: {st.camera_rays} camera paths in {st.seconds_tiles:.2f} s "
                      f"(C++ restatement of rustracer's renderer, -O3 -march=native, ZeroTwoSequence sampler; the Rust binary cannot be built here)",
            "mrays_per_s": (st.regular_rays + st.shadow_rays) / st.seconds_tiles / 1e6, "seconds": st.seconds_tiles, "camera_rays": int(st.camera_rays)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="rtb200_")
    sc = build_scene(a, tmp)
    # each step = one bounded sample; K steps sized to finish within minutes
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    vals = []
    last = None
    for i in range(a.warmup + a.steps):
        r = oracle_rate(sc, a, seconds=per_step)
        if i >= a.warmup:
            vals.append(r)
        last = r
    v = float(np.mean([r["value"] for r in vals])) if vals else last["value"]
    secs = float(np.mean([r["seconds"] for r in vals])) if vals else last["seconds"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, sc),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mrays_per_s": float(np.mean([r["mrays_per_s"] for r in vals])) if vals else last["mrays_per_s"], "gpu_launches": 0}
    emit(line)


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
    sc = build_scene(a, tmp)
    sc.flatten()
    dev = Device(local).upload(sc)
    for kv in filter(None, os.environ.get("RT_OPTIONS", "").split(",")):        # A/B runs: RT_OPTIONS=overlap_bounces=0,node_threshold=12
        dev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    rd = sc.render_desc()
    rd.seed = 1
    spp = a.spp_per_step
    job_spp = rd.spp
    n_pix = (rd.pixel_bounds[2] - rd.pixel_bounds[0]) * (rd.pixel_bounds[3] - rd.pixel_bounds[1])

    def step_desc(k, clear):
        # step k, rank r -> sample indices [(k*world + r) * spp, +spp) of the job (wrapping past job_spp keeps per-step work fixed)
        s0 = ((k * world + rank) * spp) % job_spp
        rd.sample_begin, rd.sample_end = s0, min(job_spp, s0 + spp)
        rd.clear_film = 1 if clear else 0
        return rd

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    film_host = torch.empty((rd.cropped[3] - rd.cropped[1], rd.cropped[2] - rd.cropped[0], 4), dtype=torch.float32, pin_memory=True)

    # ---- device-resident timing: K steps, CUDA events inside rtgpu_render (stats.ms_total), max over ranks -------
    for k in range(a.warmup):
        dev.render(step_desc(k, k == 0))
        dev.read_film(out=film_host)             # also warms the read-back path (staging buffer, lazily loaded kernel)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = dev.launch_count
    t_wall = time.perf_counter()
    ms, cam, reg, sh = 0.0, 0, 0, 0
    for k in range(a.steps):
        st = dev.render(step_desc(a.warmup + k, False))
        ms += st.ms_total
        cam += st.camera_rays
        reg += st.regular_rays
        sh += st.shadow_rays
    film_t = None
    if dist is not None:                     # the job's single film reduce over NVLink, inside the timed region
        film_t = torch.as_tensor(dev.film_device_array(), device=f"cuda:{local}")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(film_t, dst=0, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    barrier()
    wall = time.perf_counter() - t_wall
    launches = dev.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None

    # ---- end-to-end through the C ABI with host buffers: render step + film read-back to host memory every step ----
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        dev.render(step_desc(a.warmup + k, False))
        dev.read_film(out=film_host)                 # X,Y,Z,weight into pinned host memory (rtgpu_read_film)
    torch.cuda.synchronize()
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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, sc), "mrays_per_s": (reg + sh) / (ms * 1e-3) / 1e6,
            "rays": {"camera": cam, "regular": reg, "shadow": sh}, "wall_s_timed_region": wall,
            "e2e": {"value": cam / e2e_s, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(type(rd)), "d2h_bytes_per_step": int(film_host.numel() * 4)},
            "gpu_launches": int(launches), "clocks": clk}

    # ---- roofline of the dominant kernel class (closest-hit traversal), measured live with CUDA events on the context's stream ----
    try:
        dev.set_option("profile", 1)
        dev.set_option("count_traversal", 1)
        stc = dev.render(step_desc(a.warmup, False))            # counting pass: N and T of this step's rays (not timed)
        dev.set_option("count_traversal", 0)
        acc = {"closest": 0.0, "anyhit": 0.0, "shade": 0.0, "other": 0.0, "total": 0.0, "launches": 0}
        for k in range(min(a.steps, 4)):
            stp = dev.render(step_desc(a.warmup, False))
            acc["closest"] += stp.ms_closest
            acc["anyhit"] += stp.ms_anyhit
            acc["shade"] += stp.ms_shade
            acc["other"] += stp.ms_other
            acc["total"] += stp.ms_total
            acc["launches"] += stp.closest_launches
        dev.set_option("profile", 0)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = float(peaks.get("hbm_gbs", 6650.0))
        n_rays = stc.closest_rays                                                            # rays walked by the closest-hit kernels
        bytes_step = 48.0 * n_rays + 32.0 * stc.nodes_closest + 48.0 * stc.prims_closest     # 32 B ray + 16 B hit + 32 N + 48 T
        bytes_any = 33.0 * stc.anyhit_rays + 32.0 * stc.nodes_anyhit + 48.0 * stc.prims_anyhit   # 32 B ray + 1 B flag + 32 N + 48 T
        reps = min(a.steps, 4)
        traffic = None                       # measured DRAM bytes per launch of the same kernels, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and a.spp_per_step == 8 and a.level == 5 and (a.xres, a.yres) == (1920, 1080):
            traffic = float(json.load(open(tpath))["dram_bytes_per_launch"])
        t_closest = acc["closest"] / reps * 1e-3
        achieved = bytes_step / t_closest / 1e9
        t_any = acc["anyhit"] / reps * 1e-3
        line["roofline_anyhit"] = {"bound": "hbm", "kernel": "k_trace_shadow_engine (any-hit BVH traversal: NEE shadow rays + MIS rays towards infinite lights)",
                                   "achieved": bytes_any / max(t_any, 1e-9) / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_any / max(t_any, 1e-9) / 1e9 / peak,
                                   "algorithmic_bytes_per_step": bytes_any, "rays_per_step": int(stc.anyhit_rays),
                                   "nodes_per_ray": stc.nodes_anyhit / max(1, stc.anyhit_rays), "prims_per_ray": stc.prims_anyhit / max(1, stc.anyhit_rays)}
        line["roofline"] = {"bound": "hbm", "kernel": "k_trace_closest_engine + k_trace_mis_engine (closest-hit BVH traversal)", "achieved": achieved, "peak": peak,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s", "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "traffic_source": "profiles/roofline_traffic.json (ncu --set full, DRAM read+write per launch)" if traffic else None,
                            "algorithmic_bytes_per_launch": bytes_step / max(1.0, acc["launches"] / reps), "algorithmic_bytes_per_step": bytes_step, "rays_per_step": int(n_rays),
                            "nodes_per_ray": stc.nodes_closest / max(1, n_rays), "prims_per_ray": stc.prims_closest / max(1, n_rays),
                            "launches_per_step": acc["launches"] / reps, "avg_launch_ms": acc["closest"] / max(1, acc["launches"]),
                            "share_of_step": {k: acc[k] / acc["total"] for k in ("closest", "anyhit", "shade", "other")},
                            "note": "scene (80 MB) is L2-resident on B200: achieved counts algorithmic bytes, so it can exceed what DRAM delivers"}
    except Exception as e:  # the headline number must not depend on the diagnostics
        line["roofline"] = {"error": str(e)}

    if not a.no_cpu_baseline:
        # In a fresh process: measured inside this one (CUDA context, torch thread pools, the clock sampler) the same CPU
        # sample ran 2x slower than through `--impl reference` (profiles/r01h), which would flatter the GPU/CPU ratio.
        # One untimed pass first: a cold first pass was 2x slower once (profiles/r01zc_bench.json, 2.66 M against 4.90 M samples/s).
        try:
            import subprocess
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1", "--spp-per-step", str(a.spp_per_step),
                   "--level", str(a.level), "--xres", str(a.xres), "--yres", str(a.yres), "--cpu-seconds", str(a.cpu_seconds)]
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env).stdout.strip().splitlines()[-1]
            line["cpu_baseline"] = json.loads(out)["cpu_baseline"]
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}
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
