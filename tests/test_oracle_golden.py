"""The oracle against its committed vectors (tests/golden/oracle_vectors.npz, made by tests/golden/gen_oracle_vectors.py).
These pin the restatement against accidental change; the device is checked against the same vectors in test_gpu_golden.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_oracle_vectors as gen  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"))


@pytest.mark.parametrize("name", list(gen.CASES))
def test_oracle_reproduces_golden_vectors(native_libs, gold, name):
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(gen.CASES[name]())
    sc.flatten()
    o = ob.OracleScene(sc.ir_ptr)
    lo, hi = sc.nodes()
    r = o.intersect(scenes.ray_batch(gen.N_RAYS, lo[0, :3], hi[0, :3]))
    assert np.array_equal(r["prim"], gold[f"{name}/prim"])
    assert np.array_equal(r["t"], gold[f"{name}/t"])
    assert np.array_equal(r["nodes"], gold[f"{name}/nodes"]) and np.array_equal(r["prims"], gold[f"{name}/prims_tested"])
    occ = o.occluded(scenes.ray_batch(gen.N_RAYS, lo[0, :3], hi[0, :3], any_hit=True))["occluded"]
    assert np.array_equal(occ, gold[f"{name}/occluded"])
    pix = gen.pixel_samples(sc.render_desc(), gen.N_LI)
    li, pfilm = o.li_samples(pix, seed=gen.SEED)
    assert np.array_equal(pfilm, gold[f"{name}/pfilm"])
    assert np.allclose(li, gold[f"{name}/li"], rtol=1e-5, atol=1e-7)


def test_counter_sampler_is_a_stratified_net(native_libs):
    """The device's counter sampler (twin in the oracle): per pixel, each of the first `dimensions` 1-D / 2-D draws visits
    every stratum exactly once over the spp samples, like ZeroTwoSequence (zerotwosequence.rs:67-108)."""
    import ctypes as C
    from oracle import binding as ob
    l = ob.lib()
    spp = 16
    vals = np.zeros((spp, 3), np.float32)
    for counter in range(4):
        for s in range(spp):
            l.orc_counter_draws(3, 5, 42, s, spp, 4, counter, vals[s].ctypes.data_as(C.POINTER(C.c_float)))
        assert sorted(np.floor(vals[:, 0] * spp).astype(int).tolist()) == list(range(spp))
        for lx in range(5):
            nx, ny = 1 << lx, 1 << (4 - lx)
            cells = np.floor(vals[:, 1] * nx).astype(int) * ny + np.floor(vals[:, 2] * ny).astype(int)
            assert len(set(cells.tolist())) == spp
    # beyond `dimensions`: plain random numbers in [0, 1)
    for s in range(spp):
        l.orc_counter_draws(3, 5, 42, s, spp, 4, 9, vals[s].ctypes.data_as(C.POINTER(C.c_float)))
    assert (vals >= 0).all() and (vals < 1).all() and len(np.unique(vals[:, 0])) == spp


def test_counter_and_reference_samplers_agree_statistically(native_libs):
    """Image parity between the two samplers is statistical: same mean image within Monte-Carlo noise."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.cornell_box(xres=32, yres=32, spp=64))
    o = ob.OracleScene(sc.ir_ptr)
    _, a, _ = o.render(sampler_kind=0)
    _, b, _ = o.render(sampler_kind=1, seed=9)
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02
    # block means (8x8 pixel tiles) agree to a few percent
    ta = a.reshape(4, 8, 4, 8, 3).mean((1, 3, 4))
    tb = b.reshape(4, 8, 4, 8, 3).mean((1, 3, 4))
    assert np.abs(ta - tb).max() / ta.mean() < 0.15


def test_spatial_light_distribution_is_thread_count_independent(native_libs, tmp_path):
    """SpatialLightDistribution::lookup (lightdistrib.rs:183-296) fills its voxel table on first use from whichever thread gets there first; the
    oracle's lock-free table (one atomic pointer per voxel, a losing thread drops its copy) must give every voxel the same distribution whoever
    builds it: the film rendered by 1 thread and by 8 threads contending on fresh tables is identical, sample for sample (tiles own their pixels
    under the box filter, so the merge order does not matter either)."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.c3_scene(str(tmp_path), level=2, xres=96, yres=64, spp=4), search_dir=tmp_path)
    films = []
    for threads in (1, 8, 8):
        o = ob.OracleScene(sc.ir_ptr)                      # a fresh scene = empty voxel tables
        film, _, st = o.render(sampler_kind=0, threads=threads)
        assert st.threads == threads
        films.append(film)
    assert np.array_equal(films[0], films[1]) and np.array_equal(films[1], films[2])
