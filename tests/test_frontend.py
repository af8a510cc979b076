"""Host front end (C++ restatement of pbrt/lexer.rs, pbrt/parser.rs, api.rs, paramset.rs): the reference's lexer / parser
KATs, the defaults of SURVEY App. D and the reference's failure modes."""
import numpy as np
import pytest

from rustracer_b200 import Scene, SceneError, host, scenes

LEXER_SCENE = r'''
LookAt 0 0 5 0 0 0 0 1 0
Camera "perspective" "float fov" [50]


Film "image" "integer xresolution" [800] "integer yresolution" [600]
    "string filename" "test-whitted.tga"

Integrator "whitted"

WorldBegin
  LightSource "distant" "point from" [0 1 5] "point to" [0 0 0]

  #Material "matte" "rgb Kd" [1.0 0.0 0.0] "float sigma" [20]
  AttributeBegin
    #Material "matte" "rgb Kd" [1.0 0.0 0.0]
    Material "plastic" "rgb Kd" [1.0 0.0 0.0] "rgb Ks" [1.0 1.0 1.0]

    Shape "sphere"
  AttributeEnd

  AttributeBegin
    Rotate -90 1 0 0
    Material "matte" "rgb Kd" [1.0 1.0 1.0]
    Shape "disk" "float radius" [20] "float height" [-1]
  AttributeEnd

  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [2.0 2.0 2.0]
    Rotate 90 1 0 0
    Shape "disk" "float height" [-2] "float radius" [0.5]
  AttributeEnd
WorldEnd
        '''


def test_tokenize_sample_scene(native_libs):
    """pbrt/lexer.rs:269-309 `test_tokenize`: the whole sample scene tokenises with nothing left over."""
    toks = host.tokenize(LEXER_SCENE)
    assert toks[0] == "LookAt" and toks[-1] == "WorldEnd"
    assert toks.count("COMMENT") == 2
    assert "STR:test-whitted.tga" in toks and "NUMBER:-90" in toks


def test_token_kats(native_libs):
    """pbrt/lexer.rs:311-336: float, string, keyword and comment tokens."""
    assert host.tokenize("-1.23e2") == ["NUMBER:-123"]
    assert host.tokenize('"this is a string"') == ["STR:this is a string"]
    assert host.tokenize("Accelerator") == ["Accelerator"]
    assert host.tokenize("[") == ["["]
    assert host.tokenize("#foo\n") == ["COMMENT"]
    with pytest.raises(SceneError):
        host.tokenize("#comment without newline")       # lexer.rs:258-263: a comment must end with a newline


def test_param_header_kats(native_libs):
    """pbrt/parser.rs:345-363: "float fov" -> (Float, "fov")."""
    assert host.param_header("float fov") == (2, "fov")
    assert host.param_header("integer indices")[1] == "indices"
    assert host.param_header("rgb Kd")[1] == "Kd"
    assert host.param_header("bogus x") is None


def test_defaults_app_d(native_libs):
    """what an empty parameter list means (path.rs:49-78, film.rs:117-150, zerotwosequence.rs:58-63, ...)."""
    sc = Scene.from_string('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    ir = sc.ir
    assert (ir.film.xres, ir.film.yres) == (1280, 720) and ir.film.scale == 1.0
    assert ir.film.filter == 0 and (ir.film.filter_xw, ir.film.filter_yw) == (0.5, 0.5)      # box 0.5 (api.rs:285)
    assert ir.sampler.spp == 16 and ir.sampler.dimensions == 4
    assert ir.integrator.type == 0 and ir.integrator.max_depth == 5 and ir.integrator.rr_threshold == 1.0 and ir.integrator.light_strategy == 1
    assert ir.accel.split_method == 0 and ir.accel.max_node_prims == 4
    assert ir.camera.fov == 90.0 and ir.camera.lens_radius == 0.0 and ir.camera.focal_distance == 1e6
    s = ir.shapes[0]
    assert s.kind == 1 and s.radius == 1.0 and s.zmin == -1.0 and s.zmax == 1.0 and s.phimax == 360.0
    m = ir.materials[s.material]
    assert m.type == 0 and list(m.kd) == [0.5, 0.5, 0.5] and m.sigma == 0.0                   # default material: matte (api.rs:349)
    assert sc.film_filename == "image.png"                                                   # film.rs:118-125
    sc2 = Scene.from_string('Camera "perspective"\nFilm "image" "string filename" "out.png"\nSampler "02sequence"\nWorldBegin\nWorldEnd\n')
    assert sc2.film_filename == "rt-out.png"


def test_material_defaults(native_libs):
    txt = ('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nMaterial "plastic"\nShape "sphere"\nMaterial "glass"\nShape "sphere"\n'
           'Material "mirror"\nShape "sphere"\nMaterial "metal"\nShape "sphere"\nWorldEnd\n')
    sc = Scene.from_string(txt)          # keep the owner alive: `ir` is a view into it
    ir = sc.ir
    pl, gl, mi, me = (ir.materials[ir.shapes[i].material] for i in range(4))
    assert pl.type == 1 and list(pl.kd) == [0.25] * 3 and list(pl.ks) == [0.25] * 3 and abs(pl.roughness - 0.1) < 1e-7 and pl.remap_roughness == 1
    assert gl.type == 3 and list(gl.kr) == [1.0] * 3 and list(gl.kt) == [1.0] * 3 and gl.eta == 1.5 and gl.uroughness == 0.0
    assert mi.type == 4 and all(abs(v - 0.9) < 1e-7 for v in mi.kr)
    assert me.type == 2 and abs(me.roughness - 0.01) < 1e-8 and me.has_uroughness == 0


def test_f1_material_defaults_and_lobe_lists(native_libs):
    """uber / substrate / translucent / mix: parameter defaults of material/{uber,substrate,translucent,
    mixmat}.rs `create`, and the host's lobe lists (scene_build.cpp) against the oracle's independent restatement."""
    import ctypes as C
    from oracle import binding as ob
    from rustracer_b200 import _abi as A
    txt = ('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nMaterial "uber"\nShape "sphere"\nMaterial "substrate"\nShape "sphere"\n'
           'Material "translucent"\nShape "sphere"\nMakeNamedMaterial "a" "string type" "glass"\nMakeNamedMaterial "b" "string type" "uber" "rgb opacity" [0.5 0.5 0.5] "rgb Kr" [0.1 0.1 0.1]\n'
           'Material "mix" "string namedmaterial1" "a" "string namedmaterial2" "b"\nShape "sphere"\n'
           'Material "mix" "string namedmaterial1" "nope" "string namedmaterial2" "a" "rgb Kd" [0.1 0.2 0.3]\nShape "sphere"\nWorldEnd\n')
    sc = Scene.from_string(txt)
    ir = sc.ir
    ub, su, tr, mx, mx2 = (ir.materials[ir.shapes[i].material] for i in range(5))
    assert ub.type == 6 and list(ub.kd) == [0.25] * 3 and list(ub.ks) == [0.25] * 3 and list(ub.kr) == [0.0] * 3 and list(ub.kt) == [0.0] * 3
    assert list(ub.opacity) == [1.0] * 3 and ub.eta == 1.5 and abs(ub.roughness - 0.1) < 1e-7 and ub.has_uroughness == 0 and ub.remap_roughness == 1
    assert su.type == 7 and list(su.kd) == [0.5] * 3 and list(su.ks) == [0.5] * 3 and abs(su.uroughness - 0.1) < 1e-7 and abs(su.vroughness - 0.1) < 1e-7
    assert tr.type == 8 and list(tr.reflect) == [0.5] * 3 and list(tr.transmit) == [0.5] * 3 and list(tr.kd) == [0.25] * 3
    assert mx.type == 9 and list(mx.amount) == [0.5] * 3 and ir.materials[mx.mix_a].type == 3 and ir.materials[mx.mix_b].type == 6
    # an undefined named material falls back to matte built from the mix's own parameters (api.rs:1168-1171)
    assert ir.materials[mx2.mix_a].type == 0 and np.allclose(list(ir.materials[mx2.mix_a].kd), [0.1, 0.2, 0.3]) and any("undefined" in w for w in sc.warnings)
    d = sc.desc.contents
    o = ob.OracleScene(sc.ir_ptr)
    wo = np.array([0.0, 0.6, 0.8], np.float32)
    for i in range(5):
        row = ir.shapes[i].material
        m = d.materials[row]
        assert m.type == 6                                                    # RTGPU_MAT_LOBES
        for allow in (0, 1):
            r = o.material_bsdf(row, wo, wo, [0.5, 0.5], allow_multiple_lobes=bool(allow))
            assert m.lobe_count[allow] == r["n_lobes"] and m.bsdf_eta == r["eta"], (i, allow)
            kinds = [d.lobes[m.lobe_first[allow] + k].kind for k in range(m.lobe_count[allow])]
            if i == 3:    # glass + uber(opacity 0.5, Kr): FresnelSpecular | SpecRefl + SpecTrans, then pass-through, Lambert, microfacet, SpecRefl
                assert kinds == ([4] if allow else [2, 3]) + [3, 0, 5, 2]
                assert all(d.lobes[m.lobe_first[allow] + k].n_scales == 1 for k in range(m.lobe_count[allow]))
    # nine BxDFs: the reference's BxDFHolder holds eight (bsdf/mod.rs:41-52, index out of bounds)
    many = ('Camera "perspective"\nSampler "02sequence"\nWorldBegin\n'
            'MakeNamedMaterial "u" "string type" "uber" "rgb opacity" [0.5 0.5 0.5] "rgb Kr" [0.1 0.1 0.1] "rgb Kt" [0.1 0.1 0.1]\n'
            'MakeNamedMaterial "t" "string type" "translucent"\nMaterial "mix" "string namedmaterial1" "u" "string namedmaterial2" "t"\nShape "sphere"\nWorldEnd\n')
    with pytest.raises(SceneError):
        Scene.from_string(many).flatten()


def test_reference_failure_modes(native_libs):
    """The same inputs the reference rejects are rejected."""
    ok = 'Camera "perspective"\nSampler "02sequence"\nWorldBegin\nWorldEnd\n'
    Scene.from_string(ok)
    with pytest.raises(SceneError):      # default sampler "halton" is not constructible (api.rs:205-215,287)
        Scene.from_string('Camera "perspective"\nWorldBegin\nWorldEnd\n')
    with pytest.raises(SceneError):      # tokenised but never parsed (lexer.rs:201-240 vs parser.rs:150-185)
        Scene.from_string('Identity\n' + ok)
    with pytest.raises(SceneError):
        Scene.from_string(ok.replace('"perspective"', '"orthographic"'))
    with pytest.raises(SceneError):      # api.rs:1109-1115 unimplemented! shapes
        Scene.from_string('Camera "perspective"\nSampler "02sequence"\nWorldBegin\nShape "cone"\nWorldEnd\n')
    with pytest.raises(SceneError):      # Shape outside the world block
        Scene.from_string('Camera "perspective"\nSampler "02sequence"\nShape "sphere"\nWorldBegin\nWorldEnd\n')
    with pytest.raises(SceneError):
        Scene.from_string('Camera "perspective"\nSampler "02sequence"\nIntegrator "bidir"\nWorldBegin\nWorldEnd\n')


def test_integrator_names(native_libs):
    """make_integrator's names (api.rs:231-246) plus the GPU aliases and the harness-added ambient occlusion."""
    base = 'Camera "perspective"\nSampler "02sequence"\nIntegrator "%s"\nWorldBegin\nWorldEnd\n'
    for name, typ in (("path", 0), ("whitted", 1), ("directlighting", 2), ("ambientocclusion", 3), ("normal", 4),
                      ("gpupath", 0), ("gpuwhitted", 1), ("gpudirectlighting", 2), ("gpuao", 3), ("gpunormal", 4)):
        sc = Scene.from_string(base % name)
        assert sc.ir.integrator.type == typ
    sc = Scene.from_string('Camera "perspective"\nSampler "02sequence"\nIntegrator "directlighting" "string strategy" "one"\nWorldBegin\nWorldEnd\n')
    assert sc.ir.integrator.direct_strategy == 1


def test_primitive_and_light_numbering(native_libs):
    """App. B: primitives numbered in Shape order, triangles in face order; one area light per triangle appended after the
    shape's primitives (api.rs:933-963)."""
    sc = Scene.from_string(scenes.cornell_box(xres=32, yres=32, spp=1))
    sc.flatten()
    d = sc.desc.contents
    assert d.n_prims == 36 and d.n_lights == 2
    info = sc.prim_info()
    slot = sc.slot_of_prim()
    assert sorted(info[:, 0].tolist()) == list(range(36))
    assert info[slot[34], 2] == 0 and info[slot[35], 2] == 1            # the light quad is the last Shape: prims 34, 35
    assert (info[[slot[i] for i in range(34)], 2] == 0xFFFFFFFF).all()
    lights = [d.lights[i] for i in range(2)]
    assert [l.prim_slot for l in lights] == [slot[34], slot[35]] and all(l.kind == 3 for l in lights)
    assert np.allclose([l.area for l in lights], 130 * 105 / 2)


def test_render_desc_camera_and_bounds(native_libs):
    """Film::new / get_sample_bounds (film.rs:66-75,249-257) with a crop window and the spp power-of-two rounding
    (zerotwosequence.rs:32)."""
    sc = Scene.from_string(scenes.cornell_box(xres=100, yres=60, spp=12, crop=[0.25, 0.75, 0.5, 1.0]))
    rd = sc.render_desc()
    assert list(rd.cropped) == [25, 30, 75, 60] and list(rd.sample_bounds) == [25, 30, 75, 60] and list(rd.pixel_bounds) == [25, 30, 75, 60]
    assert rd.spp == 16
    sc = Scene.from_string(scenes.cornell_box(xres=64, yres=64, spp=4).replace('PixelFilter "box"', 'PixelFilter "gaussian"'))
    rd = sc.render_desc()
    assert list(rd.sample_bounds) == [-2, -2, 66, 66] and rd.filter_radius[0] == 2.0
    assert rd.filter_table[0] > rd.filter_table[255] >= 0.0


def test_ply_round_trip(native_libs, tmp_path):
    """plymesh.rs:18-178: binary little-endian PLY with `list uchar int vertex_indices`."""
    v, f = scenes.icosphere(1)
    scenes.write_ply(tmp_path / "m.ply", v, f)
    txt = scenes.header(16, 16, 1, 'Integrator "path"', 40, ([0, 0, -5], [0, 0, 0], [0, 1, 0])) + 'WorldBegin\nShape "plymesh" "string filename" "m.ply"\nWorldEnd\n'
    sc = Scene.from_string(txt, search_dir=tmp_path)
    s = sc.ir.shapes[0]
    assert s.n_indices == 3 * len(f) and s.n_vertices == len(v)
    got = np.ctypeslib.as_array(s.P, shape=(s.n_vertices, 3))
    assert np.array_equal(got, v.astype(np.float32))
    assert np.array_equal(np.ctypeslib.as_array(s.indices, shape=(len(f), 3)), f)


def test_ply_large_mesh_fast_paths(native_libs, tmp_path):
    """The all-triangles record copy of large face lists (ply_reader.cpp, taken from 65,536 faces on) against numpy, and the same file with its
    last face turned into a quad, which must take the per-face loop and fan into two triangles (plymesh.rs:107-127)."""
    v, f = scenes.icosphere(6)                                           # 81,920 faces
    assert len(f) >= 1 << 16
    scenes.write_ply(tmp_path / "big.ply", v, f)
    txt = scenes.header(16, 16, 1, 'Integrator "path"', 40, ([0, 0, -5], [0, 0, 0], [0, 1, 0])) + 'WorldBegin\nShape "plymesh" "string filename" "big.ply"\nWorldEnd\n'
    sc = Scene.from_string(txt, search_dir=tmp_path)
    s = sc.ir.shapes[0]
    assert s.n_indices == 3 * len(f) and s.n_vertices == len(v)
    assert np.array_equal(np.ctypeslib.as_array(s.indices, shape=(len(f), 3)), f)
    assert np.array_equal(np.ctypeslib.as_array(s.P, shape=(s.n_vertices, 3)), v.astype(np.float32))
    # same mesh, last record a quad (a, b, c, d) -> triangles (a, b, c) and (d, a, c)
    raw = open(tmp_path / "big.ply", "rb").read()
    quad = np.array([f[-1][0], f[-1][1], f[-1][2], f[0][0]], "<i4")
    open(tmp_path / "quad.ply", "wb").write(raw[:-13] + bytes([4]) + quad.tobytes())
    sc2 = Scene.from_string(txt.replace("big.ply", "quad.ply"), search_dir=tmp_path)
    s2 = sc2.ir.shapes[0]
    assert s2.n_indices == 3 * (len(f) + 1)
    got = np.ctypeslib.as_array(s2.indices, shape=(len(f) + 1, 3))
    assert np.array_equal(got[:-2], f[:-1]) and list(got[-2]) == list(quad[:3]) and list(got[-1]) == [quad[3], quad[0], quad[2]]
