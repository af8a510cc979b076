"""li()/image parity probe (run on the GPU box): python tools/render_check.py"""
import copy
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes, _abi as A
from rustracer_b200.device import Device
from oracle import binding as ob


def cmp_li(name, dev, sc, o, n=4000, seed=7):
    rd = sc.render_desc()
    rd.seed = seed
    sb = list(rd.sample_bounds)
    rng = np.random.default_rng(1)
    pix = np.stack([rng.integers(sb[0], sb[2], n), rng.integers(sb[1], sb[3], n), rng.integers(0, rd.spp, n)], 1).astype(np.int32)
    t = time.time()
    ref, _ = o.li_samples(pix, seed=seed)
    t_ref = time.time() - t
    got = dev.li_samples(rd, pix)
    err = np.abs(got - ref)
    tol = 1e-4 * np.maximum(np.abs(ref), 1e-3) + 1e-6
    bad = (err > tol).any(1)
    exact = (got == ref).all(1).mean()
    print(f"[{name}] li: n={n} exact={exact*100:.2f}% within 1e-4 rel={100-100*bad.mean():.3f}% mean ref={ref.mean():.5f} mean got={got.mean():.5f} "
          f"max abs err={err.max():.3g} (oracle {t_ref:.2f}s)")
    if bad.any():
        idx = np.where(bad)[0][:5]
        for i in idx:
            print("   mismatch", pix[i], ref[i], got[i])
    return bad.mean()


def cmp_image(name, dev, sc, o, seed=3):
    rd = sc.render_desc()
    rd.seed = seed
    st = dev.render(rd)
    film = dev.read_film()
    rgb = dev.resolve_film()
    film_ref, rgb_ref, ost = o.render(sampler_kind=1, seed=seed)
    d = np.abs(rgb - rgb_ref)
    rel = d.sum() / max(1e-12, np.abs(rgb_ref).sum())
    print(f"[{name}] image {rgb.shape}: mean ref={rgb_ref.mean():.5f} got={rgb.mean():.5f} L1 rel diff={rel:.3g} max abs={d.max():.3g} weight equal={np.array_equal(film[..., 3], film_ref[..., 3])} "
          f"| gpu {st.ms_total:.1f} ms, camera={st.camera_rays} regular={st.regular_rays} shadow={st.shadow_rays} launches={st.kernel_launches} waves={st.waves} "
          f"| oracle camera={ost.camera_rays} regular={ost.regular_rays} shadow={ost.shadow_rays} {ost.seconds_tiles:.2f}s x{ost.threads}thr")
    return rel


def main():
    dev = Device(0)
    tmp = tempfile.mkdtemp()
    # camera rays
    sc = Scene.from_string(scenes.cornell_box(xres=64, yres=64, spp=4))
    dev.upload(sc)
    o = ob.OracleScene(sc.ir_ptr)
    rng = np.random.default_rng(0)
    samples = np.concatenate([rng.uniform(0, 64, (5000, 2)), rng.uniform(0, 1, (5000, 2))], 1).astype(np.float32)
    rd = sc.render_desc()
    print("camera rays bit-equal:", np.array_equal(dev.generate_rays(rd, samples)[:, :7], o.camera_rays(samples)[:, :7]))

    cases = [
        ("cornell path uniform", scenes.cornell_box(xres=64, yres=64, spp=8), None),
        ("cornell path spatial", scenes.cornell_box(xres=64, yres=64, spp=8, integrator='Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "spatial"'), None),
        ("balls path", scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "path" "integer maxdepth" [5]'), None),
        ("balls whitted", scenes.balls(xres=96, yres=72, spp=8), None),
        ("balls direct all", scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'), None),
        ("balls direct one", scenes.balls(xres=96, yres=72, spp=8, integrator='Integrator "directlighting" "string strategy" "one" "integer maxdepth" [5]'), None),
        ("balls ao", scenes.balls(xres=96, yres=72, spp=4, integrator='Integrator "ambientocclusion" "integer nsamples" [16]'), None),
        ("balls normal", scenes.balls(xres=96, yres=72, spp=4, integrator='Integrator "normal"'), None),
        ("c3-l2 path spatial", scenes.c3_scene(tmp, level=2, xres=96, yres=54, spp=8), tmp),
        ("c3-l2 ao", scenes.c3_scene(tmp, level=2, xres=96, yres=54, spp=4, integrator='Integrator "ambientocclusion" "integer nsamples" [16]'), tmp),
    ]
    for name, txt, sd in cases:
        sc = Scene.from_string(txt, search_dir=sd)
        dev.upload(sc)
        o = ob.OracleScene(sc.ir_ptr)
        try:
            cmp_li(name, dev, sc, o)
            cmp_image(name, dev, sc, o)
        except Exception as e:
            print(f"[{name}] FAILED: {e}")


if __name__ == "__main__":
    main()
