#!/usr/bin/env python3
"""Generate tests/golden/copper_rgb.json from the reference sources (run in the build container only).

Metal's default `eta` / `k` are copper SPDs converted to RGB at material creation
(rustracer-core/src/material/metal.rs:24-30,84-159) through `Spectrum::from_sampled`
(rustracer-core/src/spectrum.rs:110-126, 196-212) and the CIE tables in rustracer-core/src/cie.rs.
The tables themselves are reference data we do not copy; this script evaluates the conversion in
float32, step by step as the Rust does, and commits only the six resulting floats.
"""
import json, re, sys
import numpy as np

REF = "/root/reference/rustracer-core/src"
f32 = np.float32


def arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"(?:pub )?const (\w+): \[f32; \w+\] = \[(.*?)\];", txt, re.S):
        body = re.sub(r"//[^\n]*", "", m.group(2))
        out[m.group(1)] = [f32(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", body)]
    for m in re.finditer(r"(?:pub )?const (\w+): f32 = ([-+\d.eE]+);", txt):
        out[m.group(1)] = f32(m.group(2))
    return out


def find_interval(size, pred):  # lib.rs:171-189
    first, length = 0, size
    while length > 0:
        half = length >> 1
        middle = first + half
        if pred(middle):
            first = middle + 1
            length -= half + 1
        else:
            length = half
    return min(max(first - 1, 0), size - 2)


def interp(lam, vals, n, l):  # spectrum.rs:196-212
    if l <= lam[0]:
        return vals[0]
    if l >= lam[n - 1]:
        return vals[n - 1]
    off = find_interval(n, lambda i: lam[i] <= l)
    t = f32(f32(l - lam[off]) / f32(lam[off + 1] - lam[off]))
    return f32(f32(vals[off] * f32(f32(1.0) - t)) + f32(vals[off + 1] * t))  # lerp: a*(1-t) + b*t


def from_sampled(cie, lam, v, n):  # spectrum.rs:110-126
    X, Y, Z, L = cie["CIE_X"], cie["CIE_Y"], cie["CIE_Z"], cie["CIE_LAMBDA"]
    N = len(L)
    xyz = [f32(0), f32(0), f32(0)]
    for i in range(N):
        val = interp(lam, v, n, L[i])
        xyz[0] = f32(xyz[0] + f32(val * X[i]))
        xyz[1] = f32(xyz[1] + f32(val * Y[i]))
        xyz[2] = f32(xyz[2] + f32(val * Z[i]))
    scale = f32(f32(L[N - 1] - L[0]) / f32(cie["CIE_Y_INTEGRAL"] * f32(N)))
    xyz = [f32(c * scale) for c in xyz]
    r = f32(f32(f32(f32(3.240479) * xyz[0]) - f32(f32(1.537150) * xyz[1])) - f32(f32(0.498535) * xyz[2]))
    g = f32(f32(f32(f32(-0.969256) * xyz[0]) + f32(f32(1.875991) * xyz[1])) + f32(f32(0.041556) * xyz[2]))
    b = f32(f32(f32(f32(0.055648) * xyz[0]) - f32(f32(0.204043) * xyz[1])) + f32(f32(1.057311) * xyz[2]))
    return [float(r), float(g), float(b)]


if __name__ == "__main__":
    cie = arrays(f"{REF}/cie.rs")
    met = arrays(f"{REF}/material/metal.rs")
    n = len(met["COPPER_WAVELENGTHS"])
    assert n == 56 and len(cie["CIE_LAMBDA"]) == 471, (n, len(cie["CIE_LAMBDA"]))
    out = {
        "source": "rustracer-core/src/material/metal.rs:24-30,84-159 via spectrum.rs:110-126 and cie.rs",
        "copper_eta_rgb": from_sampled(cie, met["COPPER_WAVELENGTHS"], met["COPPER_N"], n),
        "copper_k_rgb": from_sampled(cie, met["COPPER_WAVELENGTHS"], met["COPPER_K"], n),
    }
    out["copper_eta_rgb_hex"] = [float(f32(x)).hex() for x in out["copper_eta_rgb"]]
    out["copper_k_rgb_hex"] = [float(f32(x)).hex() for x in out["copper_k_rgb"]]
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/copper_rgb.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
