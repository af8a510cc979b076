#!/usr/bin/env python3
"""Generate tests/golden/oracle_vectors.npz: seeded inputs + outputs of the CPU oracle (oracle/) on the harness scenes.

The reference (Rust) cannot be built or run here, so these vectors pin the ORACLE, not the reference: they
are a regression anchor for the restatement and a fixed target for the device (tests/test_gpu_golden.py).  Scenes and
rays are regenerated from seeds by rustracer_b200.scenes, so only the expected outputs are stored.

    python tests/golden/gen_oracle_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402
from rustracer_b200 import Scene, scenes  # noqa: E402

import tempfile  # noqa: E402
_TEX_DIR = os.path.join(tempfile.gettempdir(), "rustracer_b200_golden_textures")

CASES = {
    "cornell_path": lambda: scenes.cornell_box(xres=64, yres=64, spp=8),
    "balls_path": lambda: scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "path" "integer maxdepth" [5]'),
    "balls_whitted": lambda: scenes.balls(xres=96, yres=72, spp=8),
    "balls_direct_all": lambda: scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'),
    "balls_direct_one": lambda: scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "directlighting" "string strategy" "one" "integer maxdepth" [5]'),
    "balls_ao": lambda: scenes.balls(xres=96, yres=72, spp=4, integrator='Integrator "ambientocclusion" "integer nsamples" [16]'),
    # uber / substrate / translucent / mix
    "balls_ext_path": lambda: scenes.balls_ext(xres=96, yres=72, spp=8, integrator='Integrator "path" "integer maxdepth" [5]'),
    "balls_ext_whitted": lambda: scenes.balls_ext(xres=96, yres=72, spp=8),
    "balls_ext_direct_all": lambda: scenes.balls_ext(xres=96, yres=72, spp=8, integrator='Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'),
    # textures + bump mapping (the texture images are regenerated next to the system temp directory)
    "textured_path": lambda: scenes.balls_textured(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True),
    "textured_whitted": lambda: scenes.balls_textured(_TEX_DIR, xres=96, yres=72, spp=8, integrator='Integrator "whitted" "integer maxdepth" [4]', absolute_paths=True),
    # image-mapped infinite light (rotated), distant light, disk / cylinder / two-sided triangle area lights
    "lights_path": lambda: scenes.lights_zoo(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True),
    "lights_path_uniform": lambda: scenes.lights_zoo(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True,
                                                     integrator='Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "uniform"'),
    "lights_direct_all": lambda: scenes.lights_zoo(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True,
                                                   integrator='Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'),
    "lights_direct_one": lambda: scenes.lights_zoo(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True,
                                                   integrator='Integrator "directlighting" "string strategy" "one" "integer maxdepth" [5]'),
    "lights_whitted": lambda: scenes.lights_zoo(_TEX_DIR, xres=96, yres=72, spp=8, absolute_paths=True, integrator='Integrator "whitted" "integer maxdepth" [5]'),
}
N_RAYS, N_LI, SEED = 2000, 400, 11


def pixel_samples(rd, n):
    rng = np.random.default_rng(5)
    sb = list(rd.sample_bounds)
    return np.stack([rng.integers(sb[0], sb[2], n), rng.integers(sb[1], sb[3], n), rng.integers(0, rd.spp, n)], 1).astype(np.int32)


def main():
    out = {}
    for name, make in CASES.items():
        sc = Scene.from_string(make())
        sc.flatten()
        o = ob.OracleScene(sc.ir_ptr)
        lo, hi = sc.nodes()
        rays = scenes.ray_batch(N_RAYS, lo[0, :3], hi[0, :3])
        r = o.intersect(rays)
        out[f"{name}/prim"], out[f"{name}/t"] = r["prim"], r["t"]
        out[f"{name}/nodes"], out[f"{name}/prims_tested"] = r["nodes"], r["prims"]
        seg = scenes.ray_batch(N_RAYS, lo[0, :3], hi[0, :3], any_hit=True)
        out[f"{name}/occluded"] = o.occluded(seg)["occluded"]
        pix = pixel_samples(sc.render_desc(), N_LI)
        li, pfilm = o.li_samples(pix, seed=SEED)
        out[f"{name}/li"], out[f"{name}/pfilm"] = li, pfilm
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
