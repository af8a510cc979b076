"""The boundary from plain C (tests/c/abi_smoke.c): compiled with gcc against include/*.h and linked with the two shared libraries —
no Python, no torch on the call path.  CPU: it compiles, links (every symbol the C program uses resolves) and fails cleanly without a
GPU.  GPU: create -> upload -> render -> read_film / resolve_film; with two GPUs also the multi-process NCCL film reduce."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "abi_smoke")


def _compile():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    lib = os.path.join(ROOT, "rustracer_b200", "lib")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
           "-L" + lib, "-lrthost", "-lrtgpu", "-Wl,-rpath," + lib, "-lm", "-o", EXE]
    subprocess.run(cmd, check=True, cwd=ROOT)


def test_c_program_compiles_links_and_needs_a_gpu(native_libs):
    _compile()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 10 and "rtgpu_create" in r.stderr          # no CPU fallback: the product path fails loudly


@pytest.mark.gpu
def test_c_program_renders(native_libs):
    _compile()
    r = subprocess.run([EXE], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout


@pytest.mark.gpu
def test_c_program_two_ranks_nccl(native_libs):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _compile()
    # the program arms alarm(240) in both ranks, so a wedged collective ends it with SIGALRM (-14) well inside this timeout
    r = subprocess.run([EXE, "2"], cwd=ROOT, capture_output=True, text=True, timeout=300, env=dict(os.environ, NCCL_DEBUG="WARN"))
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    assert "abi_smoke (2 ranks, NCCL) ok" in r.stdout
