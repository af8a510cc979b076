"""Where does the end-to-end step time go?  python tools/e2e_probe.py   (GPU box)"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device

tmp = tempfile.mkdtemp()
sc = Scene.from_string(scenes.c3_scene(tmp, level=5), search_dir=tmp)
sc.flatten()
dev = Device(0).upload(sc)
rd = sc.render_desc()
rd.sample_begin, rd.sample_end = 0, 8
dev.render(rd)
h, w = rd.cropped[3] - rd.cropped[1], rd.cropped[2] - rd.cropped[0]
pinned = torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True)
pageable = np.zeros((h, w, 4), np.float32)
for name, buf in (("pinned", pinned), ("pageable", pageable)):
    dev.read_film(out=buf)
    t = time.perf_counter()
    for _ in range(5):
        dev.read_film(out=buf)
    dt = (time.perf_counter() - t) / 5
    print(f"read_film {name}: {dt * 1e3:.2f} ms  ({h * w * 16 / dt / 1e9:.1f} GB/s)")
g = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
for _ in range(2):
    pinned.copy_(g, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    pinned.copy_(g, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print(f"torch D2H pinned: {dt * 1e3:.2f} ms  ({h * w * 16 / dt / 1e9:.1f} GB/s)")
t = time.perf_counter()
for _ in range(5):
    st = dev.render(rd)
dt = (time.perf_counter() - t) / 5
print(f"render wall {dt * 1e3:.2f} ms, device {st.ms_total:.2f} ms")
# replica of bench.py's e2e loop, timing each part
for rep in range(2):
    tr = tf = 0.0
    t0 = time.perf_counter()
    for k in range(8):
        rd.sample_begin, rd.sample_end = (k * 8) % 256, (k * 8) % 256 + 8
        a = time.perf_counter(); st = dev.render(rd); b = time.perf_counter(); dev.read_film(out=pinned); c = time.perf_counter()
        tr += b - a; tf += c - b
    torch.cuda.synchronize()
    print(f"e2e loop: {(time.perf_counter() - t0) / 8 * 1e3:.2f} ms/step (render {tr / 8 * 1e3:.2f}, read_film {tf / 8 * 1e3:.2f}); last device ms {st.ms_total:.2f}")
