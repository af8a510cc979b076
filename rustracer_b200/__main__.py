"""`python -m rustracer_b200 [-t N] INPUT.pbrt` — the rustracer CLI (rustracer-cli/src/main.rs:11-44) on the GPU backend."""
import argparse

from .integrator import render_file


def main():
    ap = argparse.ArgumentParser(prog="rustracer_b200")
    ap.add_argument("input", metavar="INPUT")
    ap.add_argument("-t", "--nthreads", type=int, default=0, help="host threads for the BVH build")
    ap.add_argument("-o", "--output", default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--integrator", default=None, help="override the scene's integrator (path, whitted, directlighting, ambientocclusion, normal)")
    a = ap.parse_args()
    print(render_file(a.input, device=a.device, threads=a.nthreads, out=a.output, integrator=a.integrator))


if __name__ == "__main__":
    main()
