"""Locate and load the in-tree native libraries.  No fallbacks: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
REPO_ROOT = os.path.dirname(_HERE)

_cache = {}


class NativeLibraryMissing(RuntimeError):
    pass


def load(name):
    """name: 'rthost' | 'rtgpu'.  Raises NativeLibraryMissing with the build hint if the .so is absent."""
    if name in _cache:
        return _cache[name]
    # RT_LIB_VARIANT selects a differently-tuned build of the same sources (tools/engine_sweep.sh); never a fallback
    variant = os.environ.get("RT_LIB_VARIANT", "") if name == "rtgpu" else ""
    path = os.path.join(LIB_DIR, f"lib{name}{variant}.so")
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `python -m rustracer_b200.build`). There is no CPU fallback for the GPU path.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    _cache[name] = lib
    return lib
