"""Textures, MIP maps, noise, bump mapping and ray differentials: self-tests of the oracle's restatement
of rustracer-core/src/{texture/*.rs, mipmap.rs, noise.rs, camera.rs:150-202}, the front end's `Texture` directive and the
host's MIP pyramids (which must equal the oracle's, texel for texel).

The reference holds no test or fixture for any of this (its mipmap.rs tests check ndarray indexing only): parity unpinned.
What is checked here are closed-form values of the algorithms as written in the reference."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from rustracer_b200 import Scene, scenes

HEAD = ('Camera "perspective" "float fov" [40]\nFilm "image" "integer xresolution" [32] "integer yresolution" [32]\n'
        'Sampler "02sequence" "integer pixelsamples" [4]\nPixelFilter "box"\nIntegrator "path"\nWorldBegin\n')


def _scene(body, search_dir=None):
    return Scene.from_string(HEAD + body + 'Shape "sphere"\nWorldEnd\n', search_dir=search_dir)


def _pts(uv, p=(0, 0, 0), dpdx=(0, 0, 0), dpdy=(0, 0, 0), duv=(0, 0, 0, 0)):
    uv = np.atleast_2d(np.asarray(uv, np.float32))
    a = np.zeros((len(uv), 15), np.float32)
    a[:, 0:2] = uv
    a[:, 2:5] = p
    a[:, 5:8] = dpdx
    a[:, 8:11] = dpdy
    a[:, 11:15] = duv
    return a


def _row(sc, kind):
    ir = sc.ir
    return [i for i in range(ir.n_textures) if ir.textures[i].kind == kind][-1]


def test_checkerboard_uv_scale_mix_values():
    sc = _scene('Texture "c" "spectrum" "checkerboard" "float uscale" [4] "float vscale" [2] "rgb tex1" [1 0 0] "rgb tex2" [0 0 1]\n'
                'Texture "n" "spectrum" "checkerboard" "string aamode" "none" "float uscale" [4] "float vscale" [2]\n'
                'Texture "u" "spectrum" "uv" "float uscale" [3] "float udelta" [0.5]\n'
                'Texture "s" "spectrum" "scale" "texture tex1" "c" "rgb tex2" [0.5 0.5 0.25]\n'
                'Texture "m" "float" "mix" "float tex1" [2] "float tex2" [4] "float amount" [0.25]\n')
    o = ob.OracleScene(sc.ir_ptr)
    ir = sc.ir
    rows = {k: v for k, v in zip("cnusm", [i for i in range(ir.n_textures) if ir.textures[i].kind != 0])}
    uv = [[0.1, 0.1], [0.3, 0.1], [0.3, 0.6], [0.1, 0.6]]                       # st = (0.4,0.2) (1.2,0.2) (1.2,1.2) (0.4,1.2)
    red, blue = [1, 0, 0], [0, 0, 1]
    got = o.texture_eval(rows["c"], _pts(uv))                                    # no differentials: the point-sampled branch
    assert np.array_equal(got, np.array([red, blue, red, blue], np.float32))
    assert np.array_equal(o.texture_eval(rows["n"], _pts(uv)), np.array([[1] * 3, [0] * 3, [1] * 3, [0] * 3], np.float32))   # defaults: white / black
    # closed-form box filter: a footprint wider than a check (ds > 1) gives the 50 % blend (checkerboard.rs:137-139)
    wide = o.texture_eval(rows["c"], _pts([[0.1, 0.1]], duv=(0.3, 0, 0, 0)))
    assert np.allclose(wide, [[0.5, 0, 0.5]])
    # a footprint straddling one edge in s: area2 from the integrated bump function
    st_s, ds = 4 * 0.24, 4 * 0.02                                                # s in [0.88, 1.04]
    bump_int = lambda x: np.floor(x / 2) + 2 * max(x / 2 - np.floor(x / 2) - 0.5, 0)
    sint = (bump_int(st_s + ds) - bump_int(st_s - ds)) / (2 * ds)
    edge = o.texture_eval(rows["c"], _pts([[0.24, 0.1]], duv=(0.02, 0, 0, 0.01)))   # (dt = 0 would give 0/0 = NaN, as in the reference)
    assert np.allclose(edge, [[1 - sint, 0, sint]], atol=1e-5) and 0.1 < sint < 0.4
    assert np.allclose(o.texture_eval(rows["u"], _pts([[0.3, 0.7]])), [[0.4, 0.7, 0.0]], atol=1e-6)   # fract(3*0.3 + 0.5), fract(0.7)
    assert np.allclose(o.texture_eval(rows["s"], _pts(uv[:2])), [[0.5, 0, 0], [0, 0, 0.25]])
    assert np.allclose(o.texture_eval(rows["m"], _pts(uv[:1])), [[2.5] * 3])     # 2 * 0.75 + 4 * 0.25


def test_planar_mapping_uses_world_position():
    sc = _scene('Texture "p" "spectrum" "checkerboard" "string mapping" "planar" "vector v1" [1 0 0] "vector v2" [0 0 1] "float udelta" [0.5]\n')
    o = ob.OracleScene(sc.ir_ptr)
    r = _row(sc, 3)
    got = o.texture_eval(r, np.concatenate([_pts([[9, 9]], p=(0.2, 5, 0.2)), _pts([[9, 9]], p=(0.7, 5, 0.2)), _pts([[9, 9]], p=(-0.7, 5, 1.2))]))
    # s = x + 0.5, t = z: floor(0.7) + floor(0.2) = 0 -> tex1 ; 1 + 0 -> tex2 ; floor(-0.2) + floor(1.2) = -1 + 1 = 0 -> tex1 (i32 sum, checkerboard.rs:125)
    assert np.array_equal(got[:, 0], np.array([1, 0, 1], np.float32))


def test_noise_and_fbm():
    l = ob.lib()
    for p in [(0, 0, 0), (3, -2, 7), (255, 256, -256)]:
        assert l.orc_noise(*map(float, p)) == 0.0                                # gradient noise vanishes on the lattice
    vals = np.array([l.orc_noise(0.37 * i, 1.3 + 0.11 * i, -0.5 + 0.23 * i) for i in range(200)])
    assert np.abs(vals).max() <= 1.0 + 1e-6 and np.abs(vals).max() > 0.2 and abs(vals.mean()) < 0.15
    assert l.orc_noise(0.25, 0.375, 0.5) == l.orc_noise(256.25, 0.375, 256.5) != 0    # period 256
    import ctypes
    P = lambda a: np.asarray(a, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    p = np.array([0.3, 0.4, 0.5], np.float32)
    z = np.zeros(3, np.float32)
    # no differentials: log2(0) = -inf -> n = max octaves, all of them full weight and a zero-weight partial one (noise.rs:44-59)
    want = sum((0.5 ** i) * l.orc_noise(*(np.float32(1.99) ** i * p).astype(np.float32).tolist()) for i in range(3))
    got = l.orc_fbm(P(p), P(z), P(z), 0.5, 3)
    assert abs(got - want) < 1e-5
    big = np.array([4, 0, 0], np.float32)                                        # huge footprint: n = clamp(-1 - 0.5*log2(16)) = 0 -> only the partial octave with weight smoothstep(0) = 0
    assert l.orc_fbm(P(p), P(big), P(z), 0.5, 8) == 0.0


def test_mipmap_pyramid_of_power_of_two_image(tmp_path):
    img = np.arange(8 * 4 * 3, dtype=np.float32).reshape(4, 8, 3) / 10
    scenes.write_pfm(str(tmp_path / "a.pfm"), img)
    sc = _scene('Texture "i" "spectrum" "imagemap" "string filename" "a.pfm"\n', search_dir=str(tmp_path))
    o = ob.OracleScene(sc.ir_ptr)
    lv = o.texture_mip(_row(sc, 5))
    assert [l.shape for l in lv] == [(4, 8, 3), (2, 4, 3), (1, 2, 3), (1, 1, 3)]    # 1 + log2(max(8, 4)) levels (mipmap.rs:150)
    assert np.array_equal(lv[0], img[::-1])                                      # ImageTexture flips the rows (imagemap.rs:51-58)
    want1 = (lv[0][0::2, 0::2] + lv[0][0::2, 1::2] + lv[0][1::2, 0::2] + lv[0][1::2, 1::2]) * np.float32(0.25)
    assert np.array_equal(lv[1], want1)
    # level 3 (1x1) of a 2x1 level with wrap = repeat: texel(.., t = 1) wraps to row 0 (mipmap.rs:161-169, :197)
    assert np.allclose(lv[3][0, 0], (lv[2][0, 0] + lv[2][0, 1]) * 0.5)


def test_mipmap_resampling_keeps_a_constant_image_and_host_pyramid_equals_oracle(tmp_path):
    scenes.write_pfm(str(tmp_path / "c.pfm"), np.full((5, 12, 3), 0.4, np.float32))
    scenes.texture_images(str(tmp_path))
    body = ('Texture "c" "spectrum" "imagemap" "string filename" "c.pfm" "string wrap" "clamp"\n'
            'Texture "a" "spectrum" "imagemap" "string filename" "tex_rgb.pfm"\n'
            'Texture "b" "float" "imagemap" "string filename" "tex_small.png" "string wrap" "black" "float scale" [2]\n'
            'Texture "d" "spectrum" "imagemap" "string filename" "tex_rgb.pfm" "string wrap" "black"\n'
            'Material "matte" "texture Kd" "a" "texture sigma" "b"\n')
    sc = _scene(body, search_dir=str(tmp_path))
    o = ob.OracleScene(sc.ir_ptr)
    ir = sc.ir
    rows = [i for i in range(ir.n_textures) if ir.textures[i].kind == 5]
    lv = o.texture_mip(rows[0])
    assert lv[0].shape == (8, 16, 3)                                             # 12x5 -> 16x8 (mipmap.rs:73-76)
    for l in lv:
        assert np.allclose(l, 0.4, atol=1e-6)                                    # Lanczos weights are normalised (mipmap.rs:394-396)
    # wrap "black": the resampling loses energy at the borders, the interior stays put
    assert o.texture_mip(rows[3])[0][0, 0, 0] < o.texture_mip(rows[1])[0][0, 0, 0] + 1e-6
    # the product's pyramids (csrc/host/texture_build.cpp, uploaded to the device) are the oracle's, bit for bit
    sc.flatten()
    desc = sc.desc.contents
    assert desc.n_textures == ir.n_textures and desc.n_texmats == ir.n_materials
    pool = np.ctypeslib.as_array(desc.tex_data, shape=(desc.n_tex_floats,))
    for r in rows:
        t = desc.textures[r]
        levels = o.texture_mip(r)
        assert t.n_levels == len(levels)
        for i, l in enumerate(levels):
            got = pool[t.level_offset[i]:t.level_offset[i] + l.size].reshape(l.shape)
            assert (t.level_v[i], t.level_u[i], t.channels) == l.shape and np.array_equal(got, l)
    # float imagemap: luminance of the scaled, inverse-gamma'd PNG texel (imagemap.rs:72-83, spectrum.rs:379-385)
    g = (127.5 + 127.5 * np.sin(0 * 0.9) * np.cos(15 * 0.6))                     # texel (x=0, y=15) of tex_small.png = first texel after the flip
    v = np.float32(int(g)) / np.float32(255)
    lin = ((v + np.float32(0.055)) / np.float32(1.055)) ** np.float32(2.4)
    assert np.allclose(o.texture_mip(rows[2])[0][0, 0, 0], 2 * lin * (0.212671 + 0.715160 + 0.072169), rtol=1e-5)


def test_image_lookups_trilinear_and_ewa(tmp_path):
    img = np.zeros((4, 4, 3), np.float32)
    img[:, :, 0] = np.arange(4)[None, :]                                         # r = column
    img[:, :, 1] = np.arange(4)[::-1][:, None]                                   # g = row after the flip
    scenes.write_pfm(str(tmp_path / "g.pfm"), img)
    sc = _scene('Texture "t" "spectrum" "imagemap" "string filename" "g.pfm" "bool trilinear" "true" "string wrap" "clamp"\n'
                'Texture "e" "spectrum" "imagemap" "string filename" "g.pfm" "string wrap" "clamp"\n', search_dir=str(tmp_path))
    o = ob.OracleScene(sc.ir_ptr)
    ir = sc.ir
    t, e = [i for i in range(ir.n_textures) if ir.textures[i].kind == 5]
    # zero footprint -> bilinear "triangle" filter on level 0 (mipmap.rs:215-216, :246-247): texel centres reproduce the texels
    for r in (t, e):
        got = o.texture_eval(r, _pts([[(1 + 0.5) / 4, (2 + 0.5) / 4], [0.5, 0.5]]))
        assert np.allclose(got[0], [1, 2, 0], atol=1e-6) and np.allclose(got[1], [1.5, 1.5, 0], atol=1e-6)
    # a footprint as wide as the image -> the 1x1 top level = the mean (trilinear: level >= n_levels - 1)
    assert np.allclose(o.texture_eval(t, _pts([[0.3, 0.3]], duv=(0.6, 0, 0, 0.6))), [[1.5, 1.5, 0]], atol=1e-6)
    # EWA: isotropic footprint of one texel on a linear ramp returns the ramp value at the centre (symmetric weights)
    got = o.texture_eval(e, _pts([[0.5, 0.5]], duv=(0.25, 0, 0, 0.25)))
    assert np.allclose(got, [[1.5, 1.5, 0]], atol=1e-4)
    # anisotropic footprint: clamped to max_anisotropy, still finite and inside the texel range
    got = o.texture_eval(e, _pts([[0.4, 0.6]], duv=(0.5, 0, 0, 0.001)))
    assert np.isfinite(got).all() and 0 <= got[0, 0] <= 3 and 0 <= got[0, 1] <= 3


def test_png_reader_matches_the_generator(tmp_path):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    scenes.write_png8(str(tmp_path / "r.png"), rgb)
    sc = _scene('Texture "i" "spectrum" "imagemap" "string filename" "r.png" "bool gamma" "false" "string wrap" "clamp"\n', search_dir=str(tmp_path))
    t = sc.ir.textures[_row(sc, 5)]
    assert (t.img_w, t.img_h) == (9, 7)
    got = np.ctypeslib.as_array(t.texels, shape=(7, 9, 3))
    assert np.array_equal(got, (rgb[::-1].astype(np.float32) / np.float32(255)))  # imageio.rs:102-110, flipped by imagemap.rs:51-58
    assert not sc.warnings


def test_tga_reader(tmp_path):
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (5, 6, 3), dtype=np.uint8)
    hdr = bytes([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 0, 5, 0, 24, 0])                     # raw true-colour, bottom-left origin
    (tmp_path / "a.tga").write_bytes(hdr + rgb[::-1, :, ::-1].tobytes())
    rle = bytearray([0, 0, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 0, 5, 0, 24, 0x20])               # RLE, top-left origin
    for y in range(5):
        rle += bytes([0x80 | 1]) + bytes(rgb[y, 0, ::-1]) if (rgb[y, 0] == rgb[y, 1]).all() else bytes([1]) + rgb[y, 0:2, ::-1].tobytes()
        rle += bytes([3]) + rgb[y, 2:6, ::-1].tobytes()
    (tmp_path / "b.tga").write_bytes(bytes(rle))
    for name in ("a.tga", "b.tga"):
        sc = _scene(f'Texture "i" "spectrum" "imagemap" "string filename" "{name}" "bool gamma" "false"\n', search_dir=str(tmp_path))
        t = sc.ir.textures[_row(sc, 5)]
        assert (t.img_w, t.img_h) == (6, 5) and not sc.warnings
        assert np.array_equal(np.ctypeslib.as_array(t.texels, shape=(5, 6, 3)), rgb[::-1].astype(np.float32) / np.float32(255))


def test_front_end_texture_directive_semantics(tmp_path):
    sc = _scene('Texture "k" "spectrum" "constant" "rgb value" [0.1 0.2 0.3]\nTexture "f" "float" "constant" "float value" [0.7]\n'
                'Texture "chk" "spectrum" "checkerboard"\nTexture "missing" "spectrum" "imagemap" "string filename" "nope.png"\n'
                'Texture "odd" "float" "wrinkled"\nTexture "chk" "spectrum" "checkerboard"\n'
                'Material "plastic" "texture Kd" "k" "texture roughness" "f" "texture Ks" "chk" "texture bumpmap" "f"\n', search_dir=str(tmp_path))
    ir = sc.ir
    m = ir.materials[ir.shapes[0].material]
    assert np.allclose(list(m.kd), [0.1, 0.2, 0.3]) and abs(m.roughness - 0.7) < 1e-7   # constant textures fold into the fields
    assert m.tex[0] == 0 and m.tex[7] == 0 and m.tex[1] != 0 and m.tex[15] != 0 and m.textured == 1   # Ks row, bump row (a ConstantTexture row)
    w = " ".join(sc.warnings)
    assert "Could not open texture file" in w and "Unkown texture type wrinkled" in w and "being redefined" in w
    grey = [i for i in range(ir.n_textures) if ir.textures[i].kind == 5][0]
    assert (ir.textures[grey].img_w, ir.textures[grey].img_h) == (1, 1)
    lin = ((np.float32(0.18) + np.float32(0.055)) / np.float32(1.055)) ** np.float32(2.4)   # .png -> gamma defaults to true (imagemap.rs:126-129)
    assert abs(ir.textures[grey].texels[0] - lin) < 1e-6
    with pytest.raises(RuntimeError, match="unimplemented"):
        _scene('Texture "m" "spectrum" "marble"\n')
    with pytest.raises(RuntimeError, match="unimplemented"):
        _scene('Texture "c" "spectrum" "checkerboard" "string mapping" "spherical"\n')


def test_camera_differentials_and_their_scaling():
    sc = Scene.from_string(scenes.cornell_box(xres=64, yres=64, spp=16))
    o = ob.OracleScene(sc.ir_ptr)
    s = np.array([[20.5, 30.5, 0.5, 0.5]], np.float32)
    r1 = o.camera_rays_diff(s, 1.0)[0]
    rx = o.camera_rays_diff(s + np.array([[1, 0, 0, 0]], np.float32), 1.0)[0]
    ry = o.camera_rays_diff(s + np.array([[0, 1, 0, 0]], np.float32), 1.0)[0]
    assert np.allclose(r1[14:17], rx[4:7], atol=1e-6) and np.allclose(r1[17:20], ry[4:7], atol=1e-6)   # rx/ry = the neighbouring pixels' rays
    # pinhole: the camera position; the main ray's origin is additionally shifted along d by its error bound (ray.rs:52-55)
    assert np.allclose(r1[8:11], [278, 273, -800]) and np.array_equal(r1[11:14], r1[8:11]) and np.allclose(r1[0:3], r1[8:11], atol=1e-3)
    q = o.camera_rays_diff(s, 0.25)[0]                                                                   # renderer.rs:111: 1 / sqrt(16)
    assert np.allclose(q[14:17] - q[4:7], 0.25 * (r1[14:17] - r1[4:7]), atol=1e-7)


def test_bump_map_and_textured_materials_change_the_oracle_image_where_expected(tmp_path):
    base = scenes.balls_textured(str(tmp_path), xres=64, yres=48, spp=4)
    sc = Scene.from_string(base, search_dir=str(tmp_path))
    o = ob.OracleScene(sc.ir_ptr)
    _, rgb, _ = o.render(sampler_kind=1, seed=3)
    assert np.isfinite(rgb).all() and rgb.mean() > 0.02
    flat = Scene.from_string(base.replace('"texture bumpmap" "bumpf"', "").replace('"texture bumpmap" "bumpimg"', ""), search_dir=str(tmp_path))
    _, rgb_flat, _ = ob.OracleScene(flat.ir_ptr).render(sampler_kind=1, seed=3)
    d = np.abs(rgb - rgb_flat).sum(-1)
    assert (d > 1e-4).mean() > 0.01 and (d == 0).mean() > 0.3                     # bump maps perturb their own spheres only
    # a constant bump map displaces nothing: d(displace) = 0 in both directions -> same shading frame (material/mod.rs:82-91)
    const = Scene.from_string(base.replace('Texture "bumpf" "float" "fbm" "integer octaves" [4] "float omega" [0.6]', 'Texture "bumpf" "float" "scale" "float tex1" [0.5] "float tex2" [0.5]')
                              .replace('"texture bumpmap" "bumpimg"', ""), search_dir=str(tmp_path))
    nobump = Scene.from_string(base.replace('"texture bumpmap" "bumpf"', "").replace('"texture bumpmap" "bumpimg"', ""), search_dir=str(tmp_path))
    a = ob.OracleScene(const.ir_ptr).render(sampler_kind=1, seed=3)[1]
    b = ob.OracleScene(nobump.ir_ptr).render(sampler_kind=1, seed=3)[1]
    assert np.abs(a - b).sum() / np.abs(b).sum() < 2e-3
