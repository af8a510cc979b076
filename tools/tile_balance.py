"""Work balance of the tile dealing on ONE GPU: renders every rank's share of the C5 frame in turn (32 sample indices) and prints max / mean of the
device times — what a strong-scaled job loses to the slowest rank.  python tools/tile_balance.py [worlds, default 2,4,8]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device


def main():
    worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2,4,8").split(",")]
    tmp = tempfile.mkdtemp()
    sc = Scene.from_string(scenes.c5_scene(tmp), search_dir=tmp)
    dev = Device(0).upload(sc)
    rd = sc.render_desc()
    rd.sample_begin, rd.sample_end = 0, 32
    dev.render(rd)                                   # warm-up (light grid, buffers)
    for w in worlds:
        ms = []
        for r in range(w):
            rd.tile_rank, rd.tile_world, rd.clear_film = r, w, 1
            dev.render(rd)
            ms.append(dev.render(rd).ms_total)
        mean = sum(ms) / len(ms)
        print(f"world {w}: ms per rank {[round(m, 2) for m in ms]}  max / mean = {max(ms) / mean:.4f}", flush=True)


if __name__ == "__main__":
    main()
