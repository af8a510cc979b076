"""rtgpu_build_bvh (SAH BVH construction on the device) builds the reference's tree: node for node and slot for slot equal to
the host builder (rustracer_b200/csrc/host/bvh_builder.cpp), which tests/test_host_bvh.py pins against the oracle's independent
restatement of `BVH::recursive_build` + `flatten_bvh` (rustracer-core/src/bvh/mod.rs:137-358)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(native_libs):
    from rustracer_b200.device import Device
    return Device(0)


def _host_and_device(dev, sc):
    sc.flatten()
    lo_h, hi_h = (a.copy() for a in sc.nodes())
    slot_h = sc.slot_of_prim().copy()
    sc.flatten(device=dev)
    lo_d, hi_d = (a.copy() for a in sc.nodes())
    return (lo_h, hi_h, slot_h), (lo_d, hi_d, sc.slot_of_prim().copy())


def _assert_same_tree(h, d):
    assert h[0].shape == d[0].shape
    assert np.array_equal(h[0][:, :3], d[0][:, :3]) and np.array_equal(h[1][:, :3], d[1][:, :3])            # bounds
    assert np.array_equal(h[0][:, 3].view(np.uint32), d[0][:, 3].view(np.uint32))                            # primitives_offset | second_child_offset
    assert np.array_equal(h[1][:, 3].view(np.uint32), d[1][:, 3].view(np.uint32))                            # n_prims << 2 | axis
    assert np.array_equal(h[2], d[2])                                                                        # ordered_prims


@pytest.mark.parametrize("name", ["cornell", "balls", "field2", "field4", "maxprims1", "maxprims64"])
def test_device_builder_reproduces_the_host_tree(dev, tmp_path, name):
    from rustracer_b200 import Scene, scenes
    txt = {"cornell": lambda: scenes.cornell_box(xres=32, yres=32, spp=1),
           "balls": lambda: scenes.balls(xres=32, yres=32, spp=1),
           "field2": lambda: scenes.c3_scene(str(tmp_path), level=2, xres=32, yres=32, spp=1),
           "field4": lambda: scenes.c3_scene(str(tmp_path), level=4, xres=32, yres=32, spp=1),
           "maxprims1": lambda: scenes.c3_scene(str(tmp_path), level=3, xres=32, yres=32, spp=1).replace("WorldBegin", 'Accelerator "bvh" "integer maxnodeprims" [1]\nWorldBegin'),
           "maxprims64": lambda: scenes.c3_scene(str(tmp_path), level=3, xres=32, yres=32, spp=1).replace("WorldBegin", 'Accelerator "bvh" "integer maxnodeprims" [64]\nWorldBegin')}[name]()
    sc = Scene.from_string(txt, search_dir=str(tmp_path))
    h, d = _host_and_device(dev, sc)
    _assert_same_tree(h, d)


def test_degenerate_inputs(dev):
    """Coincident centroids (leaf whatever the count), duplicated boxes, a single primitive, two primitives in both orders."""
    rng = np.random.default_rng(2)

    for n in (1, 2, 3, 5, 33, 200):
        lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        b = np.concatenate([lo, lo + rng.uniform(0.01, 0.3, (n, 3)).astype(np.float32)], 1)
        r = dev.build_bvh(b)
        assert sorted(r["ordered"].tolist()) == list(range(n))
        meta = r["node_hi"][:, 3].view(np.uint32)
        assert (meta >> 2).sum() == n                                                      # every primitive in exactly one leaf
        assert np.array_equal(r["node_lo"][0, :3], b[:, :3].min(0)) and np.array_equal(r["node_hi"][0, :3], b[:, 3:].max(0))
    same = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (100, 1))                   # all centroids equal -> one leaf of 100 (bvh/mod.rs:175-180)
    r = dev.build_bvh(same)
    assert r["node_lo"].shape[0] == 1 and (r["node_hi"][0, 3:].view(np.uint32)[0] >> 2) == 100
    assert np.array_equal(r["ordered"], np.arange(100, dtype=np.uint32))


def test_renders_identically_with_the_device_built_tree(dev, tmp_path):
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.c3_scene(str(tmp_path), level=3, xres=64, yres=48, spp=4), search_dir=str(tmp_path))
    sc.flatten()
    dev.upload(sc)
    rd = sc.render_desc()
    dev.render(rd)
    a = dev.read_film().copy()
    sc.flatten(device=dev)
    dev.upload(sc)
    dev.render(rd)
    b = dev.read_film()
    assert np.array_equal(a[..., 3], b[..., 3]) and np.allclose(a, b, rtol=1e-5, atol=1e-6)   # same tree; float atomics reorder the sums
