"""Preview renders of the BASELINE scenes through the product path (for eyeballing): python tools/render_preview.py OUT_DIR
C5 at 960x540 / 64 spp and the lights-zoo scene, written as PNG by rth_write_image (and one EXR)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, host, scenes
from rustracer_b200.device import Device


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp()
    dev = Device(0)
    for name, txt in (("c5_preview", scenes.c5_scene(tmp, xres=960, yres=540, spp=64)), ("c2_preview", scenes.balls(xres=512, yres=384, spp=64, integrator='Integrator "path" "integer maxdepth" [5]')),
                      ("textured_preview", scenes.balls_textured(tmp, xres=512, yres=384, spp=32))):
        sc = Scene.from_string(txt, search_dir=tmp)
        dev.upload(sc)
        st = dev.render(sc.render_desc())
        rgb = dev.resolve_film()
        host.write_image(os.path.join(out, name + ".png"), rgb)
        print(name, rgb.shape, f"{st.ms_total:.1f} ms", float(rgb.mean()))
    host.write_image(os.path.join(out, "textured_preview.exr"), rgb)


if __name__ == "__main__":
    main()
