"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:rtgpu|rth)_\w+)\s*\(", txt)))


def test_rtgpu_exports_every_declared_symbol(native_libs):
    lib = C.CDLL(os.path.join(ROOT, "rustracer_b200", "lib", "librtgpu.so"))
    names = declared_functions("rtgpu.h")
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_rthost_exports_every_declared_symbol(native_libs):
    lib = C.CDLL(os.path.join(ROOT, "rustracer_b200", "lib", "librthost.so"))
    names = declared_functions("rthost.h")
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_ctypes_mirror(native_libs):
    """The ctypes mirrors in rustracer_b200/_abi.py must have the C layout (checked through sizes known from the headers)."""
    from rustracer_b200 import _abi as A
    assert C.sizeof(A.rtgpu_ray) == 32
    assert C.sizeof(A.rtgpu_hit) == 16
    assert C.sizeof(A.rtgpu_quadric) == 176
    assert C.sizeof(A.rt_transform) == 128
    # descriptor produced by the C++ host must be readable through the mirror
    from rustracer_b200 import Scene, scenes
    sc = Scene.from_string(scenes.cornell_box(xres=32, yres=32, spp=4))
    rd = sc.render_desc()
    assert (rd.xres, rd.yres, rd.spp, rd.max_depth) == (32, 32, 4, 5)
    assert list(rd.cropped) == [0, 0, 32, 32] and list(rd.sample_bounds) == [0, 0, 32, 32]
    assert rd.filter_table[0] == 1.0 and rd.filter_table[255] == 1.0
    assert rd.tile_world == 1 and rd.sample_end == 4


def test_every_ctypes_mirror_has_the_size_gcc_gives_the_header_struct(tmp_path):
    """sizeof of every struct of include/rt_scene.h and include/rtgpu.h, from a C probe compiled here, against rustracer_b200/_abi.py."""
    import subprocess
    from rustracer_b200 import _abi as A
    names = ["rt_transform", "rt_shape", "rt_area_light", "rt_light", "rt_texture", "rt_material", "rt_camera", "rt_film", "rt_sampler", "rt_integrator",
             "rt_accel", "rt_scene", "rtgpu_ray", "rtgpu_hit", "rtgpu_quadric", "rtgpu_instance", "rtgpu_lobe", "rtgpu_texture", "rtgpu_material", "rtgpu_light",
             "rtgpu_scene_desc", "rtgpu_render_desc", "rtgpu_stats"]
    src = '#include <stdio.h>\n#include "rtgpu.h"\nint main(void) {\n' + "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n"
    (tmp_path / "probe.c").write_text(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "probe"), str(tmp_path / "probe.c")], check=True)
    out = subprocess.run([str(tmp_path / "probe")], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    wrong = {n: (int(sizes[n]), C.sizeof(getattr(A, n))) for n in names if int(sizes[n]) != C.sizeof(getattr(A, n))}
    assert not wrong, wrong


def test_no_gpu_means_loud_failure(native_libs):
    """The product path has no CPU fallback: without a CUDA device rtgpu_create fails and the binding raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rustracer_b200.device import Device, DeviceError
    with pytest.raises(DeviceError):
        Device(0)


def test_product_does_not_import_oracle():
    """Nothing under rustracer_b200/ may import, link or call the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rustracer_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                includes = re.findall(r'#include\s+"([^"]+)"', txt)
                assert not [i for i in includes if "oracle" in i or i.startswith("orc_")], os.path.join(dirpath, f)
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt and "orc_scene_create" not in txt, os.path.join(dirpath, f)
