"""Wave-size sweep (L2 residency of the wavefront queues): python tools/wave_sweep.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device

dev = Device(0)
tmp = tempfile.mkdtemp()
cases = [("c3 path spatial 8spp", scenes.c3_scene(tmp), tmp, 8),
         ("c1 cornell path 16spp", scenes.cornell_box(), None, 16),
         ("c2 whitted 16spp", scenes.balls(), None, 16),
         ("c2 direct all 16spp", scenes.balls(integrator='Integrator "directlighting" "string strategy" "all" "integer maxdepth" [5]'), None, 16)]
for name, txt, sd, spp in cases:
    sc = Scene.from_string(txt, search_dir=sd)
    dev.upload(sc)
    for wp in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
        rd = sc.render_desc()
        rd.sample_end = min(rd.spp, spp)
        rd.wave_paths = wp
        dev.render(rd)
        st = dev.render(rd)
        print(f"{name:24s} wave_paths={wp:9d}: {st.ms_total:8.2f} ms  {st.camera_rays / st.ms_total / 1e3:8.1f} Msamples/s  waves={st.waves} launches={st.kernel_launches}", flush=True)
