"""A/B of rtgpu option overlap_bounces (0 = one stream, 2 = default) on the small configs: ms per render and film equality.
   python tools/ab_overlap_configs.py > gpurun_out/ab_overlap_configs.log"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device

dev = Device(0)
tmp = tempfile.mkdtemp()
W = 'Integrator "whitted" "integer maxdepth" [5]'
D = 'Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'
cases = [("C1 cornell path", lambda: Scene.from_string(scenes.cornell_box()), None),
         ("C2 balls whitted", lambda: Scene.from_string(scenes.balls(integrator=W)), None),
         ("C2 balls direct all", lambda: Scene.from_string(scenes.balls(integrator=D)), None),
         ("textured balls path", lambda: Scene.from_string(scenes.balls_textured(tmp, integrator=None), search_dir=tmp), None),
         ("textured balls whitted", lambda: Scene.from_string(scenes.balls_textured(tmp, integrator=W), search_dir=tmp), None),
         ("C5 4K 5M tris path (8 spp)", lambda: Scene.from_string(scenes.c5_scene(tmp), search_dir=tmp), 8)]
for label, make, spp_limit in cases:
    try:
        sc = make()
    except Exception as e:
        print(label, "skipped:", e)
        continue
    dev.upload(sc)
    rd = sc.render_desc()
    if spp_limit is not None:
        rd.sample_end = min(rd.spp, spp_limit)
    films, line = {}, []
    for ov in (0, 2, 0, 2):
        dev.set_option("overlap_bounces", ov)
        if ov not in films:
            dev.render(rd)                       # warm-up
        rd.clear_film = 1
        st = dev.render(rd)
        line.append(f"ov={ov} {st.ms_total:8.2f} ms {st.camera_rays / st.ms_total / 1e3:7.1f} M/s")
        films.setdefault(ov, ((st.regular_rays, st.shadow_rays), dev.read_film().copy()))
    f0, f2 = films[0][1], films[2][1]
    rel = float(np.abs(f0 - f2).max() / max(1e-30, np.abs(f0).max()))
    print(f"{label:28s}", " | ".join(line), "| ray counters equal:", films[0][0] == films[2][0], "| film bit-equal:", bool(np.array_equal(f0, f2)),
          f"max abs diff / max {rel:.2e}", flush=True)
