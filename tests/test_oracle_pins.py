"""Oracle-independent pins, run on the CPU oracle (tests/pins_common.py explains each property).  The same checks run on the device in
tests/test_gpu_pins.py.  Reference quirks the pins surfaced are asserted as quirks, with the Rust they come from:
  * ScaledBxDF::pdf is the default cosine pdf whatever it wraps (bsdf/bxdf.rs:48-71) -> a mix material's pdf is not its sampling density;
  * MicrofacetTransmission::pdf / f are non-zero for unreachable (wo, wi) pairs (bsdf/microfacet.rs:126-170, :215-229) -> its pdf
    integrates to more than one; on the reachable pairs it is the true density."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pins_common as pc  # noqa: E402


@pytest.fixture(scope="module")
def mats(native_libs):
    from oracle import binding as ob
    from rustracer_b200 import Scene
    sc = Scene.from_string(pc.materials_scene())
    sc.flatten()
    return ob.OracleScene(sc.ir_ptr), pc.material_rows(sc), sc


@pytest.fixture(scope="module")
def lights(native_libs, tmp_path_factory):
    from oracle import binding as ob
    from rustracer_b200 import Scene
    d = tmp_path_factory.mktemp("pins")
    txt, rows = pc.lights_scene(str(d))
    sc = Scene.from_string(txt, search_dir=str(d))
    sc.flatten()
    return ob.OracleScene(sc.ir_ptr), rows, sc


NAMES = [m[0] for m in pc.MATERIALS]
PROPS = {m[0]: m[2] for m in pc.MATERIALS}


def run_bsdf_pins(prober, rows, name):
    props, row = PROPS[name], rows[name]
    if not props.get("specular") and not props.get("no_chi2") and not props.get("cosine_pdf"):
        for mass, chi2, dof in pc.check_pdf_and_chi2(prober, row, transmission_index=props.get("transmission_index")):
            assert mass <= 1.01, (name, mass)
            assert chi2 < dof + 6.0 * np.sqrt(2.0 * dof), (name, chi2, dof)
    rel_f, rel_p, n_checked = pc.check_sample_f_consistency(prober, row)
    assert rel_f < 1e-4, (name, rel_f)
    if props.get("cosine_pdf"):
        # the quirk itself: the pdf of every direction is cos / pi, although sample_f draws from the wrapped lobes
        grid, _ = pc.sphere_grid(16, 32)
        wo = np.broadcast_to(pc.WOS[1], grid.shape)
        pdf = prober.bsdf_probe(row, wo, grid, np.zeros((len(grid), 2), np.float32), True, pc.NON_SPECULAR)["pdf"]
        assert np.allclose(pdf, np.where(grid[:, 2] > 0, grid[:, 2] / np.pi, 0.0), rtol=1e-5, atol=1e-7)
    else:
        assert rel_p < 1e-4, (name, rel_p)
    for wo in pc.WOS:                                               # white furnace: never more energy out than in
        a, err = pc.albedo(prober, row, wo)
        assert (a <= 1.0 + 5 * err + 1e-3).all(), (name, a)
        if "kd" in props:
            assert np.allclose(a, props["kd"], atol=5 * err.max() + 1e-4), (name, a)
        if "kr" in props:
            assert np.allclose(a, props["kr"], atol=1e-5), (name, a)
    if "kd" in props:                                               # f = Kd / pi exactly where both directions are above the surface
        z = np.zeros((1, 2), np.float32)
        f = prober.bsdf_probe(row, pc.WOS[:1], pc.WOS[1:2], z)["f"][0]
        assert np.allclose(f, np.float32(props["kd"]) / np.float32(np.pi), rtol=2e-7)
    if props.get("reflect_only"):
        assert pc.check_reciprocity(prober, row) < 1e-4, name


def run_light_pins(prober, rows):
    n = 20_000
    for name, row in rows.items():
        q, valid, r, ref, u = pc.check_light_pdf_consistency(prober, row, n)
        assert valid == 1.0, name
        p = ref[:, :3].astype(np.float64)
        wi = r["wi"].astype(np.float64)
        if name == "point":                                         # point.rs:43-54: I / (4 pi r^2), towards the light
            d = np.float64([1, 5, 2]) - p
            r2 = (d * d).sum(1)
            assert np.allclose(r["li"], np.float64([10, 20, 30])[None] / (4 * np.pi * r2[:, None]), rtol=1e-5)
            assert np.allclose(wi, d / np.sqrt(r2)[:, None], atol=1e-6) and (r["pdf"] == 1).all() and (r["delta"] == 1).all()
            assert np.allclose(r["p1"], [1, 5, 2]) and (r["pdf_w"] == 0).all()
        elif name == "distant":                                     # distant.rs:49-87: L from the normalised (from - to), pdf 1
            d = np.float64([0, 10, 0]) - np.float64([3, 0, 4])
            d /= np.linalg.norm(d)
            assert np.allclose(wi, d[None], atol=1e-6) and np.allclose(r["li"], [[2, 3, 4]]) and (r["pdf"] == 1).all() and (r["delta"] == 1).all()
            t = ((r["p1"].astype(np.float64) - p) * d).sum(1)       # the far end lies 2 x world radius along wi
            assert np.allclose(t, t[0], rtol=1e-4) and t[0] > 40.0
        elif name in ("triangle", "triangle_two_sided", "disk"):   # shapes/mod.rs:39-53: area pdf -> solid angle d^2 / (|cos| A)
            assert q < 1e-4, (name, q)
            q1 = r["p1"].astype(np.float64)
            d2 = ((q1 - p) ** 2).sum(1)
            cos_l = np.abs(wi[:, 1])                                # all three lights lie in planes y = const
            area = {"triangle": 2.0, "triangle_two_sided": 2.0, "disk": np.pi * 0.8 ** 2}[name]
            assert np.allclose(r["pdf"], d2 / (cos_l * area), rtol=2e-4), name
            y = {"triangle": 4.0, "triangle_two_sided": 3.0, "disk": 4.0}[name]
            assert np.allclose(q1[:, 1], y, atol=1e-5)
            # emission: both face down (-y) towards the reference points; the two-sided one also emits upwards
            assert (np.abs(r["li"]).sum(1) > 0).all(), name
        elif name == "sphere":                                      # sphere.rs:245-308: uniform cone, 1 / (2 pi (1 - cos theta_max))
            assert q < 1e-4
            dc2 = ((np.float64([0, 5, 6]) - p) ** 2).sum(1)
            cos_max = np.sqrt(np.maximum(0, 1 - 0.49 / dc2))
            assert np.allclose(r["pdf"], 1 / (2 * np.pi * (1 - cos_max)), rtol=2e-4)
            assert np.allclose(np.linalg.norm(r["p1"].astype(np.float64) - [0, 5, 6], axis=1), 0.7, atol=1e-4)
        elif name == "cylinder":
            # Cylinder::sample draws from the whole surface (cylinder.rs:259-278) while pdf_wi finds the nearest hit along wi
            # (shapes/mod.rs:59-68): the two agree where the sampled point faces the reference point — exactly the samples that
            # carry radiance (one-sided emission, diffuse.rs:91-97)
            lit = np.abs(r["li"]).sum(1) > 0
            assert 0.3 < lit.mean() < 0.7
            rel = np.abs(r["pdf_wi"][lit] - r["pdf"][lit]) / r["pdf"][lit]
            assert np.quantile(rel, 0.9) < 1e-3 and np.quantile(rel, 0.99) < 0.05      # |cos| -> 0 at the silhouette amplifies float32 noise
    # two-sided emission seen from above the light (diffuse.rs:91-97): the one-sided triangle is black from there
    above = np.tile(np.float32([0, 8, 0, 0, -1, 0]), (64, 1))
    above[:, 0] = np.linspace(-0.5, 6.5, 64)
    u = np.random.default_rng(0).random((64, 2)).astype(np.float32)
    w = np.tile(np.float32([0, -1, 0]), (64, 1))
    assert (prober.light_probe(rows["triangle"], above, u, w)["li"] == 0).all()
    assert np.allclose(prober.light_probe(rows["triangle_two_sided"], above, u, w)["li"], 4.0)
    pdf_mass, le_int, mc, err = pc.env_checks(prober, rows["infinite"])
    assert abs(pdf_mass - 1.0) < 0.01, pdf_mass                     # infinite.rs:185-198 with the 1 / (2 pi^2 sin theta) Jacobian
    assert abs(mc - le_int) < 5 * err + 0.005 * le_int, (mc, le_int, err)   # sample_li (:143-183) against le (:210-219)


@pytest.mark.parametrize("name", NAMES)
def test_bsdf_pins_on_oracle(mats, name):
    o, rows, _ = mats
    run_bsdf_pins(o, rows, name)


def test_microfacet_transmission_pdf_quirk_is_the_references(mats):
    """Without the reachability mask the pdf of a rough dielectric integrates to well over one: the reference's behaviour
    (bsdf/microfacet.rs:215-229), kept by the oracle; the pin above shows the excess sits entirely on unreachable pairs."""
    o, rows, _ = mats
    masses = [m for m, _, _ in pc.check_pdf_and_chi2(o, rows["glass_rough"], n_samples=1000)]
    assert max(masses) > 1.05


def test_light_pins_on_oracle(lights):
    o, rows, _ = lights
    run_light_pins(o, rows)


def test_distribution2d_against_numpy_inversion(lights):
    """Distribution2D::sample_continuous / pdf (sampling/distribution2d.rs:11-49) of the environment light against an inversion of
    the same scalar image written in float64 numpy: for random u, the (u, v) the light samples — recovered from the direction it
    returns through world_to_light — and its pdf."""
    o, rows, sc = lights
    tex, func = o.light_env(rows["infinite"])
    h2, w2 = func.shape
    f = func.astype(np.float64)
    row_int = f.sum(1) / w2
    cond_cdf = np.concatenate([np.zeros((h2, 1)), np.cumsum(f, 1) / w2], 1) / row_int[:, None]
    marg_int = row_int.sum() / h2
    marg_cdf = np.concatenate([[0.0], np.cumsum(row_int) / h2]) / marg_int
    rng = np.random.default_rng(2)
    n = 5000
    u = rng.random((n, 2)).astype(np.float32)
    iv = np.clip(np.searchsorted(marg_cdf, u[:, 1], side="right") - 1, 0, h2 - 1)
    dv = (u[:, 1] - marg_cdf[iv]) / (marg_cdf[iv + 1] - marg_cdf[iv])
    iu = np.array([np.clip(np.searchsorted(cond_cdf[v], x, side="right") - 1, 0, w2 - 1) for v, x in zip(iv, u[:, 0])])
    du = (u[:, 0] - cond_cdf[iv, iu]) / (cond_cdf[iv, iu + 1] - cond_cdf[iv, iu])
    uv = np.stack([(iu + du) / w2, (iv + dv) / h2], 1)
    map_pdf = f[iv, iu] / marg_int
    theta, phi = uv[:, 1] * np.pi, uv[:, 0] * 2 * np.pi
    local = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], 1)
    l2w = np.array(sc.ir.lights[2].l2w.m, np.float64).reshape(4, 4)[:3, :3]
    want_wi = local @ l2w.T
    want_pdf = map_pdf / (2 * np.pi * np.pi * np.sin(theta))
    ref = np.tile(np.float32([0, 0, 0, 0, 1, 0]), (n, 1))
    r = o.light_probe(rows["infinite"], ref, u, np.tile(np.float32([0, 1, 0]), (n, 1)))
    # a u that falls within float32 rounding of a CDF step lands in the neighbouring texel: allow a few
    close = np.abs(r["wi"] - want_wi).max(1) < 2e-4
    assert close.mean() > 0.995, close.mean()
    assert np.allclose(r["pdf"][close], want_pdf[close], rtol=2e-3)


def test_host_environment_tables_equal_the_oracles(native_libs, tmp_path):
    """The flattened infinite light (csrc/host/scene_build.cpp: MIPMap::new incl. the Lanczos resampling of non-power-of-two maps,
    the sampling image of infinite.rs:79-96 and Distribution2D::new) is bit-equal to the oracle's, for PFM / HDR maps of power-of-two,
    odd and very wide sizes."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    for kw in (dict(), dict(env_size=(24, 10)), dict(env_size=(40, 12), env_name="env.hdr"), dict(env_size=(64, 8)), dict(env_size=(128, 16))):
        sc = Scene.from_string(scenes.lights_zoo(str(tmp_path), xres=16, yres=16, spp=1, **kw), search_dir=str(tmp_path))
        assert sc.warnings == []
        sc.flatten()
        d = sc.desc.contents
        rows = [i for i in range(d.n_lights) if d.lights[i].kind == 2]
        assert len(rows) == 1
        L = d.lights[rows[0]]
        env = np.ctypeslib.as_array(d.env_data, shape=(d.n_env_floats,))
        tex, func = ob.OracleScene(sc.ir_ptr).light_env(rows[0])
        assert (L.env_h, L.env_w) == tex.shape[:2]
        assert np.array_equal(env[L.env_texels:L.env_texels + tex.size].reshape(tex.shape), tex)
        assert np.array_equal(env[L.env_func:L.env_func + func.size].reshape(func.shape), func)


def test_hdr_reader_decodes_rgbe(native_libs, tmp_path):
    """read_image_hdr (imageio.rs:115-132) restated: run-length and flat scanlines decode to c * 2^(e - 136)."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    img = scenes.env_map_image(32, 16)
    for rle in (True, False):
        want = scenes.write_hdr(str(tmp_path / "env.hdr"), img, rle=rle)
        txt = scenes.lights_zoo(str(tmp_path), xres=16, yres=16, spp=1, env_name="unused.pfm").replace("unused.pfm", "env.hdr")
        txt = txt.replace('"rgb L" [0.9 1.0 1.1] "rgb scale" [1.2 1.2 1.2]', '"rgb L" [1 1 1]')
        sc = Scene.from_string(txt, search_dir=str(tmp_path))
        sc.flatten()
        tex, _ = ob.OracleScene(sc.ir_ptr).light_env(0)
        assert np.array_equal(tex, want)
        assert np.abs(want - img).max() / img.max() < 1 / 128


@pytest.mark.parametrize("pixel_type,compression,decreasing", [("half", "zip", False), ("float", "zip", True), ("half", "none", False), ("float", "zips", False),
                                                               ("half", "rle", True)])
def test_exr_reader(native_libs, tmp_path, pixel_type, compression, decreasing):
    """read_image_exr (imageio.rs:134-160) restated for scan-line files: HALF / FLOAT channels, NONE / RLE / ZIPS / ZIP blocks, both line
    orders; 40 x 21 pixels so the last ZIP block is short.  Checked through the environment light, which keeps the texels unfiltered
    when the size is a power of two — so a 64 x 32 copy is compared bit for bit and the odd size through its resampled mean."""
    from oracle import binding as ob
    from rustracer_b200 import Scene, scenes
    for size in ((64, 32), (40, 21)):
        img = scenes.env_map_image(*size)
        want = scenes.write_exr(str(tmp_path / "env.exr"), img, pixel_type=pixel_type, compression=compression, decreasing_y=decreasing)
        txt = scenes.lights_zoo(str(tmp_path), xres=16, yres=16, spp=1, env_name="unused.pfm").replace("unused.pfm", "env.exr")
        txt = txt.replace('"rgb L" [0.9 1.0 1.1] "rgb scale" [1.2 1.2 1.2]', '"rgb L" [1 1 1]')
        sc = Scene.from_string(txt, search_dir=str(tmp_path))
        assert sc.warnings == []
        sc.flatten()
        tex, _ = ob.OracleScene(sc.ir_ptr).light_env(0)
        if size == (64, 32):
            assert np.array_equal(tex, want)
        else:
            assert tex.shape == (32, 64, 3) and abs(tex.mean() - want.mean()) / want.mean() < 0.05


def test_exr_writer_round_trip(native_libs, tmp_path):
    """write_image_exr (imageio.rs:75-92: f32 R, G, B): rth_write_image("*.exr") writes scan-line ZIP blocks; read back through the front end's
    reader as an environment map (64 x 32 stays unfiltered), bit for bit — incompressible noise exercises the stored-block path, 21 lines the
    short last block."""
    import zlib
    from oracle import binding as ob
    from rustracer_b200 import Scene, host, scenes
    rng = np.random.default_rng(3)
    for img in (scenes.env_map_image(64, 32).astype(np.float32), rng.random((32, 64, 3)).astype(np.float32)):
        host.write_image(str(tmp_path / "env.exr"), img)
        txt = scenes.lights_zoo(str(tmp_path), xres=16, yres=16, spp=1, env_name="unused.pfm").replace("unused.pfm", "env.exr")
        txt = txt.replace('"rgb L" [0.9 1.0 1.1] "rgb scale" [1.2 1.2 1.2]', '"rgb L" [1 1 1]')
        sc = Scene.from_string(txt, search_dir=str(tmp_path))
        assert sc.warnings == []
        sc.flatten()
        tex, _ = ob.OracleScene(sc.ir_ptr).light_env(0)
        assert np.array_equal(tex, img)
    # header as the format defines it: magic, version 2 single-part scan lines, channels B G R of type FLOAT, ZIP, one offset per 16 lines
    img = scenes.env_map_image(40, 21).astype(np.float32)
    host.write_image(str(tmp_path / "odd.exr"), img)
    d = open(tmp_path / "odd.exr", "rb").read()
    assert d[:8] == (20000630).to_bytes(4, "little") + (2).to_bytes(4, "little")
    assert b"channels\0chlist\0" in d and b"B\0\x02\0\0\0" in d and b"compression\0compression\0\x01\0\0\0\x03" in d
    end = d.index(b"screenWindowWidth\0float\0") + len(b"screenWindowWidth\0float\0") + 8 + 1
    offs = np.frombuffer(d[end:end + 16], "<u8")
    assert offs[0] == end + 16
    y0, size = np.frombuffer(d[int(offs[1]):int(offs[1]) + 8], "<i4")
    assert y0 == 16 and int(offs[1]) + 8 + size == len(d)
    assert len(zlib.decompress(d[int(offs[1]) + 8:])) == 5 * 40 * 12          # the short last block: 21 - 16 lines of three float planes
