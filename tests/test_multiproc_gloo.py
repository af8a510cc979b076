"""N > 1 host logic on CPU: world_size-2 `gloo` processes exercise the tile partition and the film reduce
(rustracer_b200/integrator.py).  The per-rank films come from the CPU oracle restricted to the rank's tiles — the oracle
is the checker here; the reduce and partition code under test is the product's."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    from rustracer_b200.integrator import reduce_film, tile_partition
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = Scene.from_string(scenes.cornell_box(xres=48, yres=40, spp=4))
    rd = sc.render_desc()
    mine, ntx, nty = tile_partition(list(rd.sample_bounds), rank, world)
    assert (ntx, nty) == (3, 3) and mine == [t for t in range(9) if t % world == rank]
    o = ob.OracleScene(sc.ir_ptr)
    film, _, st = o.render(sampler_kind=1, seed=3, threads=2, tile_stride=world, tile_offset=rank)
    t = torch.from_numpy(film.copy())
    reduce_film(t, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"cam_{rank}.npy"), np.array([st.camera_rays]))
    dist.barrier()
    dist.destroy_process_group()


def test_tile_partition_and_film_reduce_world2(native_libs, tmp_path):
    import torch.multiprocessing as mp
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    sc = Scene.from_string(scenes.cornell_box(xres=48, yres=40, spp=4))
    o = ob.OracleScene(sc.ir_ptr)
    full, _, st = o.render(sampler_kind=1, seed=3, threads=2)
    # disjoint tiles with the box filter: every pixel is written by exactly one rank, so the sum is exact
    assert np.array_equal(reduced, full)
    cams = sum(int(np.load(tmp_path / f"cam_{r}.npy")[0]) for r in range(2))
    assert cams == st.camera_rays == 48 * 40 * 4


def test_partitions_cover_the_work_exactly():
    from rustracer_b200.integrator import sample_partition, tile_partition
    for world in (1, 2, 3, 4, 8):
        tiles = []
        for r in range(world):
            t, ntx, nty = tile_partition([0, 0, 1920, 1080], r, world)
            tiles += t
        assert sorted(tiles) == list(range(120 * 68))
        spans = [sample_partition(1024, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 1024 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
