"""SURVEY 8d C5 check under torchrun: scene replicated, tiles (and then sample ranges) partitioned over the ranks, one NCCL
reduce of the film, and the reduced image compared with the image rank 0 renders alone (<= 1e-5 relative: fp32 sum order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py [--level 3]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--spheres", type=int, default=64)
    ap.add_argument("--xres", type=int, default=960)
    ap.add_argument("--yres", type=int, default=540)
    ap.add_argument("--spp", type=int, default=16)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rustracer_b200 import Scene, scenes
    from rustracer_b200.integrator import GpuSamplerIntegrator, sample_partition
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix=f"mgc_{rank}_")
    sc = Scene.from_string(scenes.c5_scene(tmp, level=a.level, xres=a.xres, yres=a.yres, spp=a.spp, n_spheres=a.spheres), search_dir=tmp)
    out = {"world": world, "triangles": None}
    # reference image: rank 0 alone
    solo = GpuSamplerIntegrator(sc, device=local, rank=0, world=1, seed=5)
    solo.render()
    ref = solo.film() if rank == 0 else None
    out["triangles"] = sc.n_triangles
    # (1) tiles dealt round-robin
    part = GpuSamplerIntegrator(sc, device=solo.device, rank=rank, world=world, seed=5)
    part._ready = True
    st = part.render()
    part.reduce(dst=0)
    if rank == 0:
        got = part.film()
        out["tiles"] = {"max_rel": float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3))), "equal": bool(np.array_equal(got, ref)),
                        "camera_rays_rank0": int(st.camera_rays)}
    # (2) sample-index ranges
    s0, s1 = sample_partition(a.spp, rank, world)
    allt = GpuSamplerIntegrator(sc, device=solo.device, rank=0, world=1, seed=5)
    allt._ready = True
    st = allt.render(sample_range=(s0, s1))
    allt.reduce(dst=0)
    if rank == 0:
        got = allt.film()
        out["samples"] = {"max_rel": float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3))), "range_rank0": [s0, s1], "camera_rays_rank0": int(st.camera_rays)}
        print(json.dumps(out))
        assert out["tiles"]["max_rel"] <= 1e-5 and out["samples"]["max_rel"] <= 1e-5, out
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
