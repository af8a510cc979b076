"""Where the scene set-up time goes on the GPU box: parse, flatten (device BVH build inside), upload with per-step laps (RT_UPLOAD_TIMING).
python tools/upload_probe.py [c4|c5|c3]"""
import os
import sys
import tempfile
import time

os.environ["RT_UPLOAD_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    tmp = tempfile.mkdtemp()
    txt = {"c4": scenes.c4_scene, "c5": scenes.c5_scene, "c3": scenes.c3_scene}[which](tmp)
    dev = Device(0)
    for rep in range(2):
        t0 = time.perf_counter()
        sc = Scene.from_string(txt, search_dir=tmp)
        t1 = time.perf_counter()
        sc.flatten(device=dev)
        t2 = time.perf_counter()
        dev.upload(sc)
        t3 = time.perf_counter()
        print(f"{which} rep {rep}: parse {t1 - t0:.3f} s, flatten (device BVH build {sc.bvh_build_seconds * 1e3:.1f} ms inside) {t2 - t1:.3f} s, upload {t3 - t2:.3f} s", flush=True)


if __name__ == "__main__":
    main()
