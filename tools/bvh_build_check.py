"""Device SAH builder vs host builder: equality and timings on growing scenes.   python tools/bvh_build_check.py [levels]"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device

dev = Device(0)
tmp = tempfile.mkdtemp()
cases = [("cornell", lambda: scenes.cornell_box(xres=32, yres=32, spp=1)), ("balls", lambda: scenes.balls(xres=32, yres=32, spp=1))]
for lvl in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2,4,5".split(","))]:
    cases.append((f"field level {lvl}", lambda lvl=lvl: scenes.c3_scene(tmp, level=lvl, xres=32, yres=32, spp=1)))
if "--c4" in sys.argv:
    cases.append(("c4 10M triangles", lambda: scenes.c4_scene(tmp)))
for name, make in cases:
    sc = Scene.from_string(make(), search_dir=tmp)
    t0 = time.time(); sc.flatten(); t_host_flat = time.time() - t0
    host_s = sc.bvh_build_seconds
    lo_h, hi_h = (a.copy() for a in sc.nodes()); slot_h = sc.slot_of_prim().copy()
    sc.flatten(device=dev)                                  # warm-up (allocations, module load)
    t0 = time.time(); sc.flatten(device=dev); t_dev_flat = time.time() - t0
    dev_s = sc.bvh_build_seconds
    lo_d, hi_d = sc.nodes(); slot_d = sc.slot_of_prim()
    same = lo_h.shape == lo_d.shape and np.array_equal(lo_h.view(np.uint32) & 0xffffffff, lo_d.view(np.uint32)) and np.array_equal(hi_h.view(np.uint32), hi_d.view(np.uint32)) and np.array_equal(slot_h, slot_d)
    if not same and lo_h.shape == lo_d.shape:
        bad = np.where((lo_h.view(np.uint32) != lo_d.view(np.uint32)).any(1) | (hi_h.view(np.uint32) != hi_d.view(np.uint32)).any(1))[0]
        print("   first differing nodes", bad[:5], lo_h[bad[:2]], lo_d[bad[:2]], hi_h[bad[:2]].view(np.uint32), hi_d[bad[:2]].view(np.uint32))
    t0 = time.time(); dev.upload(sc); t_up = time.time() - t0
    print(f"   rtgpu_upload_scene wall {t_up:.2f} s")
    print(f"{name}: prims {len(slot_h)} nodes host {lo_h.shape[0]} device {lo_d.shape[0]} identical={same}  host build {host_s*1e3:.1f} ms  device build (kernels) {dev_s*1e3:.2f} ms  "
          f"flatten wall host {t_host_flat:.2f} s device {t_dev_flat:.2f} s", flush=True)
