import os, sys, tempfile
sys.path.insert(0, os.getcwd())
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device
dev = Device(0); tmp = tempfile.mkdtemp()
for name, txt, spp in (("c3 path spatial 32spp", scenes.c3_scene(tmp), 32), ("c5 4K path 8spp", scenes.c5_scene(tmp), 8)):
    sc = Scene.from_string(txt, search_dir=tmp); dev.upload(sc)
    for wp in (1 << 23, 1 << 24, 1 << 25, 1 << 26):
        rd = sc.render_desc(); rd.sample_end = min(rd.spp, spp); rd.wave_paths = wp
        try:
            dev.render(rd); st = dev.render(rd)
            print(f"{name:24s} wave_paths={wp:9d}: {st.ms_total:8.2f} ms  {st.camera_rays / st.ms_total / 1e3:8.1f} Msamples/s  waves={st.waves}", flush=True)
        except Exception as e:
            print(name, wp, "failed:", e, flush=True)
