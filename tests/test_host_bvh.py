"""Host SAH BVH builder + flattening (product, C++) against the oracle's independent restatement of
BVH::recursive_build / flatten_bvh (rustracer-core/src/bvh/mod.rs:137-358): same tree, node for node, slot for slot."""
import numpy as np
import pytest

from rustracer_b200 import Scene, scenes


def _compare(sc):
    from oracle import binding as ob
    sc.flatten()
    lo, hi = sc.nodes()
    o = ob.OracleScene(sc.ir_ptr)
    bounds, meta, ordered = o.bvh()
    assert lo.shape[0] == o.n_nodes
    assert np.array_equal(lo[:, :3], bounds[:, :3]) and np.array_equal(hi[:, :3], bounds[:, 3:])
    off, m = lo[:, 3].copy().view(np.uint32), hi[:, 3].copy().view(np.uint32)
    assert np.array_equal(off, meta[:, 2].astype(np.uint32))                 # primitives_offset / second_child_offset
    assert np.array_equal(m >> 2, meta[:, 0].astype(np.uint32))              # n_prims (0 = interior)
    interior = meta[:, 0] == 0
    assert np.array_equal((m & 3)[interior], meta[interior, 1].astype(np.uint32))   # split axis
    assert np.array_equal(sc.prim_info()[:, 0], ordered.astype(np.uint32))   # ordered_prims
    # world-space triangle vertices (mesh.rs:61) in slot order
    wv = o.prim_world_vertices()
    geom = sc.prim_geom()
    tri = (geom[:, 3].copy().view(np.uint32) & 3) == 0
    got = geom[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]]
    assert np.array_equal(got[tri], wv[ordered][tri])
    return o


def test_cornell_and_balls_trees_are_identical(native_libs):
    _compare(Scene.from_string(scenes.cornell_box(xres=16, yres=16, spp=1)))
    _compare(Scene.from_string(scenes.balls(xres=16, yres=16, spp=1)))


def test_icosphere_field_tree_is_identical_and_thread_invariant(native_libs, tmp_path):
    txt = scenes.c3_scene(str(tmp_path), level=3, xres=16, yres=16, spp=1)
    sc = Scene.from_string(txt, search_dir=tmp_path)
    _compare(sc)
    lo1, hi1 = sc.nodes()
    info1 = sc.prim_info()
    sc.flatten(threads=1)
    lo2, hi2 = sc.nodes()
    assert np.array_equal(lo1.view(np.uint32), lo2.view(np.uint32)) and np.array_equal(hi1.view(np.uint32), hi2.view(np.uint32))
    assert np.array_equal(info1, sc.prim_info())


def test_maxnodeprims_and_leaf_structure(native_libs):
    """bvh/mod.rs:63-78,264-285: leaves hold at most max(maxnodeprims, ...) primitives unless the SAH says otherwise;
    every primitive appears in exactly one leaf."""
    txt = scenes.cornell_box(xres=16, yres=16, spp=1).replace("WorldBegin", 'Accelerator "bvh" "integer maxnodeprims" [1]\nWorldBegin')
    sc = Scene.from_string(txt)
    _compare(sc)
    lo, hi = sc.nodes()
    m = hi[:, 3].copy().view(np.uint32)
    n_prims = m >> 2
    assert n_prims.sum() == 36
    leaves = n_prims > 0
    starts = lo[:, 3].copy().view(np.uint32)[leaves]
    assert sorted(starts.tolist()) == np.concatenate([[0], np.cumsum(n_prims[leaves][np.argsort(starts)])[:-1]]).tolist()


def test_degenerate_scenes(native_libs):
    """Empty world and single-primitive world (bvh/mod.rs:84-90,146-152)."""
    from oracle import binding as ob
    sc = Scene.from_string('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nWorldEnd\n')
    sc.flatten()
    assert sc.desc.contents.n_nodes == 0 and sc.desc.contents.n_prims == 0
    sc = Scene.from_string('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    o = _compare(sc)
    assert o.n_nodes == 1


def test_middle_split_matches_or_rejects(native_libs):
    """SplitMethod::Middle computes `start + partition + start` (bvh/mod.rs:186-190): for start > 0 the index can
    leave the range and the reference panics; the host builder either builds the same tree as the oracle or refuses."""
    from rustracer_b200 import SceneError
    txt = scenes.cornell_box(xres=16, yres=16, spp=1).replace("WorldBegin", 'Accelerator "bvh" "string splitmethod" "middle"\nWorldBegin')
    sc = Scene.from_string(txt)
    try:
        sc.flatten()
    except SceneError as e:
        assert "middle" in str(e)
        return
    _compare(sc)


def test_itertools_partition_has_the_closed_form_the_device_builder_uses():
    """csrc/device/bvh_build.cu replaces itertools::partition (0.10.3: front scan for a failing element, back scan for a passing one,
    swap — bvh/mod.rs:187,267 call it) by: with m passing elements, the k-th failing element of [0, m) in ascending order is exchanged
    with the k-th passing element of [m, n) in descending order.  Checked here against the sequential algorithm on random inputs."""
    rng = np.random.default_rng(9)

    def sequential(items, pred):
        a = list(items)
        front, back, passed = 0, len(a), 0
        while front < back:
            if not pred(a[front]):
                swapped = False
                while front + 1 < back:
                    back -= 1
                    if pred(a[back]):
                        a[front], a[back] = a[back], a[front]
                        swapped = True
                        break
                if not swapped:
                    return a, passed
            passed += 1
            front += 1
        return a, passed

    def closed_form(items, pred):
        a = list(items)
        ok = [pred(x) for x in a]
        m = sum(ok)
        fails = [i for i in range(m) if not ok[i]]                     # ascending
        passes = [i for i in range(len(a) - 1, m - 1, -1) if ok[i]]    # descending
        assert len(fails) == len(passes)
        for i, j in zip(fails, passes):
            a[i], a[j] = a[j], a[i]
        return a, m

    for n in list(range(0, 12)) + [33, 100, 257]:
        for _ in range(40):
            keys = rng.integers(0, 12, n)
            thr = int(rng.integers(0, 12))
            items = list(zip(keys.tolist(), range(n)))                 # (bucket, identity)
            pred = lambda x: x[0] <= thr
            assert sequential(items, pred) == closed_form(items, pred)


def test_external_builder_hook_round_trips_the_host_tree(native_libs, tmp_path):
    """rth_flatten_with_builder (the hook rtgpu_build_bvh plugs into): a builder that hands back the host builder's own tree must give
    the same flattened scene; the primitive bounds it receives are the world bounds of the primitive list; a failing builder is an error."""
    import ctypes as C
    from rustracer_b200 import Scene, scenes
    from rustracer_b200.host import _lib, SceneError
    sc = Scene.from_string(scenes.c3_scene(str(tmp_path), level=1, xres=16, yres=16, spp=1), search_dir=str(tmp_path))
    sc.flatten()
    lo, hi = sc.nodes()
    slot_of_prim = sc.slot_of_prim()
    geom = sc.prim_geom().copy()
    ordered = np.empty_like(slot_of_prim)
    ordered[slot_of_prim] = np.arange(len(slot_of_prim), dtype=slot_of_prim.dtype)
    seen = {}
    BUILDER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                          C.POINTER(C.c_uint32), C.POINTER(C.c_float))

    def builder(user, bounds, n, max_prims, node_lo, node_hi, out_ordered, n_nodes, ms):
        seen["n"], seen["max_prims"] = n, max_prims
        seen["bounds"] = np.ctypeslib.as_array(bounds, shape=(n, 6)).copy()
        np.ctypeslib.as_array(node_lo, shape=(lo.shape[0], 4))[:] = lo
        np.ctypeslib.as_array(node_hi, shape=(hi.shape[0], 4))[:] = hi
        np.ctypeslib.as_array(out_ordered, shape=(n,))[:] = ordered
        n_nodes[0] = lo.shape[0]
        ms[0] = 1.5
        return 0

    cb = BUILDER(builder)
    assert _lib().rth_flatten_with_builder(sc._h, 0, C.cast(cb, C.c_void_p), None) == 0
    lo2, hi2 = sc.nodes()
    assert np.array_equal(lo.view(np.uint32), lo2.view(np.uint32)) and np.array_equal(hi.view(np.uint32), hi2.view(np.uint32))
    assert np.array_equal(sc.slot_of_prim(), slot_of_prim) and np.array_equal(sc.prim_geom(), geom)
    assert seen["n"] == len(slot_of_prim) and seen["max_prims"] == 4 and abs(sc.bvh_build_seconds - 1.5e-3) < 1e-9
    tri = geom[slot_of_prim].reshape(-1, 3, 4)[:, :, :3]                         # world-space vertices of primitive pn
    assert np.array_equal(seen["bounds"][:, :3], tri.min(1)) and np.array_equal(seen["bounds"][:, 3:], tri.max(1))
    fail = BUILDER(lambda *a: -4)
    assert _lib().rth_flatten_with_builder(sc._h, 0, C.cast(fail, C.c_void_p), None) != 0
    assert "external BVH builder failed" in _lib().rth_last_error().decode()


def test_flattened_arrays_are_fully_written_and_thread_invariant(native_libs, tmp_path):
    """The per-slot / per-node arrays of the flattener are allocated uninitialised and filled by parallel loops (hmath.hpp uvec): every word must be
    written.  Scenes with every kind of slot (triangles with and without normals / uvs, quadrics, instances, area lights) flattened with 1 and with
    8 threads, twice each, give byte-identical geometry, primitive info, node and slot tables — stale heap contents would differ between runs."""
    import ctypes as C
    from rustracer_b200 import Scene, scenes
    cases = {"lights_zoo": lambda: scenes.lights_zoo(str(tmp_path), xres=16, yres=16, spp=1), "instanced": lambda: scenes.instanced_scene(xres=16, yres=16, spp=1),
             "textured": lambda: scenes.balls_textured(str(tmp_path), xres=16, yres=16, spp=1), "field": lambda: scenes.c3_scene(str(tmp_path), level=3, xres=16, yres=16, spp=1)}

    def snapshot(txt, threads):
        # churn the heap first so that freshly malloc'ed blocks are not pristine zero pages
        junk = [np.full(1 << 18, 0x7F, np.uint8) for _ in range(8)]
        del junk
        sc = Scene.from_string(txt, search_dir=tmp_path)
        sc.flatten(threads=threads)
        d = sc.desc.contents
        out = [sc.prim_geom().view(np.uint32), sc.prim_info(), sc.slot_of_prim()] + [a.view(np.uint32) for a in sc.nodes()]
        for name, width in (("tri_n", 9), ("tri_s", 9), ("tri_uv", 6)):
            ptr = getattr(d, name)
            if ptr:
                out.append(np.ctypeslib.as_array(ptr, shape=(d.n_prims, width)).copy().view(np.uint32))
        return out

    for name, make in cases.items():
        txt = make()
        ref = snapshot(txt, 1)
        for threads in (1, 8, 8):
            got = snapshot(txt, threads)
            assert len(got) == len(ref), name
            for a, b in zip(ref, got):
                assert np.array_equal(a, b), (name, threads)
