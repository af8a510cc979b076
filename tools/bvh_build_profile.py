"""One device BVH build of the level-L field (default 5 = 1M triangles) for ncu launch lists.  python tools/bvh_build_profile.py [L]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 5
tmp = tempfile.mkdtemp()
sc = Scene.from_string(scenes.c3_scene(tmp, level=lvl, xres=32, yres=32, spp=1), search_dir=tmp)
dev = Device(0)
sc.flatten(device=dev)
print("device build ms", sc.bvh_build_seconds * 1e3, "launches", dev.launch_count)
