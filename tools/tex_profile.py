"""One render of the textured balls scene (path, 1024x768, SPP spp) for ncu captures.   python tools/tex_profile.py [spp] [integrator]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
integ = None if len(sys.argv) < 3 or sys.argv[2] == "path" else 'Integrator "whitted" "integer maxdepth" [5]'
tmp = tempfile.mkdtemp()
which = sys.argv[3] if len(sys.argv) > 3 else "textured"       # textured | ext (uber / substrate / translucent / mix, constant) | plain
if integ is None:
    integ = 'Integrator "path" "integer maxdepth" [5]'
if which == "textured":
    sc = Scene.from_string(scenes.balls_textured(tmp, spp=spp, integrator=integ), search_dir=tmp)
elif which == "ext":
    sc = Scene.from_string(scenes.balls_ext(spp=spp, integrator=integ))
else:
    sc = Scene.from_string(scenes.balls(spp=spp, integrator=integ))
dev = Device(0).upload(sc)
rd = sc.render_desc()
dev.set_option("profile", 1)
if os.environ.get("RT_SORT_ITEMS"):
    dev.set_option("sort_items", int(os.environ["RT_SORT_ITEMS"]))
st = dev.render(rd)
st = dev.render(rd)
print(which, "ms", st.ms_total, "closest", st.ms_closest, "anyhit", st.ms_anyhit, "shade", st.ms_shade, "camera", st.camera_rays)
