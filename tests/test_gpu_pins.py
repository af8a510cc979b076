"""The oracle-independent pins of tests/pins_common.py run on the DEVICE through rtgpu_bsdf_probe / rtgpu_light_probe (the code the
render kernels call: bsdf.cuh, lights.cuh), plus the direct comparison of both probes with the oracle's twins."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pins_common as pc  # noqa: E402
from test_oracle_pins import NAMES, run_bsdf_pins, run_light_pins  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mats(native_libs):
    from oracle import binding as ob
    from rustracer_b200 import Scene
    from rustracer_b200.device import Device
    sc = Scene.from_string(pc.materials_scene())
    dev = Device(0).upload(sc)
    return dev, ob.OracleScene(sc.ir_ptr), pc.material_rows(sc)


@pytest.fixture(scope="module")
def lights(native_libs, tmp_path_factory):
    from oracle import binding as ob
    from rustracer_b200 import Scene
    from rustracer_b200.device import Device
    d = tmp_path_factory.mktemp("pins")
    txt, rows = pc.lights_scene(str(d))
    sc = Scene.from_string(txt, search_dir=str(d))
    dev = Device(0).upload(sc)
    return dev, ob.OracleScene(sc.ir_ptr), rows


@pytest.mark.parametrize("name", NAMES)
def test_bsdf_pins_on_device(mats, name):
    dev, _, rows = mats
    run_bsdf_pins(dev, rows, name)


def test_light_pins_on_device(lights):
    dev, _, rows = lights
    run_light_pins(dev, rows)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("allow", [True, False])
def test_bsdf_probe_matches_oracle(mats, name, allow):
    """Bsdf::f / pdf / sample_f (bsdf/mod.rs:94-251) for every material class, allow_multiple_lobes on (path) and off (whitted / direct):
    device vs oracle on random directions.  libm differences (sin / cos / atan in the microfacet sampling) stay below 1e-4 relative;
    a different lobe pick on a rare sample (u within an ulp of a lobe boundary) is tolerated."""
    dev, o, rows = mats
    rng = np.random.default_rng(12)
    n = 20_000
    wo = pc._dir(rng.uniform(-1, 1, n), rng.uniform(0, 2 * np.pi, n))
    wi = pc._dir(rng.uniform(-1, 1, n), rng.uniform(0, 2 * np.pi, n))
    u = rng.random((n, 2)).astype(np.float32)
    for flags in (pc.BSDF_ALL, pc.NON_SPECULAR):
        a = dev.bsdf_probe(rows[name], wo, wi, u, allow, flags)
        b = o.bsdf_probe(rows[name], wo, wi, u, allow, flags)
        assert np.array_equal(a["n_lobes"], b["n_lobes"]) and np.array_equal(a["eta"], b["eta"])
        tol = lambda x: 1e-4 * np.maximum(np.abs(x), 1e-3) + 1e-6
        assert (np.abs(a["f"] - b["f"]) <= tol(b["f"])).all()
        assert (np.abs(a["pdf"] - b["pdf"]) <= tol(b["pdf"])).all()
        bad = (a["sflags"] != b["sflags"]) | (np.abs(a["spdf"] - b["spdf"]) > tol(b["spdf"])) | (np.abs(a["sf"] - b["sf"]) > tol(b["sf"])).any(1) | \
              (np.abs(a["swi"] - b["swi"]) > 1e-4).any(1)
        assert bad.mean() < 2e-3, (name, flags, bad.sum())


def test_light_probe_matches_oracle(lights):
    dev, o, rows = lights
    rng = np.random.default_rng(13)
    n = 20_000
    ref = pc.ref_points(n, 14)
    u = rng.random((n, 2)).astype(np.float32)
    w = pc._dir(rng.uniform(-1, 1, n), rng.uniform(0, 2 * np.pi, n))
    for name, row in rows.items():
        a, b = dev.light_probe(row, ref, u, w), o.light_probe(row, ref, u, w)
        for k in ("li", "wi", "p1", "le_w"):
            bad = (np.abs(a[k] - b[k]) > 1e-4 * np.maximum(np.abs(b[k]), 1e-2) + 1e-6).any(1)
            assert bad.mean() < 2e-3, (name, k, bad.sum())
        for k in ("pdf", "pdf_w", "pdf_wi"):
            # pdf_li of a cylinder / triangle re-intersects the shape: at a silhouette the hit can fall on either side of an edge
            bad = np.abs(a[k] - b[k]) > 2e-4 * np.maximum(np.abs(b[k]), 1e-3) + 1e-6
            assert bad.mean() < 5e-3, (name, k, bad.sum())
