"""Synthetic scenes for the benchmark configs of SURVEY.md §8(d) (C1..C5), written as PBRT text + binary PLY
that obey the subset rustracer's front end accepts: `Sampler "02sequence"`, box filter,
`trianglemesh` / `plymesh` / `sphere` / `disk`, matte / plastic / metal / glass / mirror.

Everything is seeded with PCG32 (rustracer-core/src/rng.rs) so the same files are produced everywhere.
"""
import os

import numpy as np

# ------------------------------------------------------------------------------------------------
# PCG32 (rng.rs:5-52), vectorised over independent streams


class PCG32:
    MULT = np.uint64(0x5851F42D4C957F2D)

    def __init__(self, seeds):
        """One generator per entry of `seeds` (== RNG::set_sequence(seed))."""
        old = np.seterr(over="ignore")
        try:
            seeds = np.atleast_1d(np.asarray(seeds, dtype=np.uint64))
            self.state = np.zeros_like(seeds)
            self.inc = (seeds << np.uint64(1)) | np.uint64(1)
            self.u32()
            self.state = self.state + np.uint64(0x853C49E6748FEA9B)
            self.u32()
        finally:
            np.seterr(**old)

    def u32(self):
        old = np.seterr(over="ignore")
        try:
            s = self.state
            self.state = s * self.MULT + self.inc
            xorshifted = (((s >> np.uint64(18)) ^ s) >> np.uint64(27)).astype(np.uint32)
            rot = (s >> np.uint64(59)).astype(np.uint32)
            return (xorshifted >> rot) | (xorshifted << ((~rot + np.uint32(1)) & np.uint32(31)))
        finally:
            np.seterr(**old)

    def f32(self):
        return np.minimum(self.u32().astype(np.float32) * np.float32(2.3283064365386963e-10), np.float32(0.99999994))


def uniform_sample_sphere(u0, u1):
    """sampling/mod.rs:14-20 in float32."""
    z = np.float32(1.0) - np.float32(2.0) * u0
    r = np.sqrt(np.maximum(np.float32(1.0) - z * z, np.float32(0.0)))
    phi = np.float32(2.0 * np.pi) * u1
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# geometry


def icosphere(level):
    """Unit icosphere: (V,3) float64 vertices, (F,3) int32 faces; 20*4**level faces."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e_sorted = np.sort(e, axis=1)
        key = e_sorted[:, 0] * (len(v) + 1) + e_sorted[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        first = np.zeros(len(uniq), dtype=np.int64)
        first[inv[::-1]] = np.arange(len(key))[::-1]
        mid = v[e_sorted[first, 0]] + v[e_sorted[first, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        mid_idx = len(v) + inv
        v = np.concatenate([v, mid], axis=0)
        n = len(f)
        a, b, c = mid_idx[:n], mid_idx[n:2 * n], mid_idx[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], a, c], 1), np.stack([f[:, 1], b, a], 1), np.stack([f[:, 2], c, b], 1), np.stack([a, b, c], 1)], axis=0)
    return v, f.astype(np.int32)


def write_ply(path, verts, faces):
    """binary_little_endian PLY: float x y z, `list uchar int vertex_indices` (what plymesh.rs accepts)."""
    verts = np.ascontiguousarray(verts, dtype="<f4")
    faces = np.ascontiguousarray(faces, dtype="<i4")
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment rustracer_b200 synthetic\n"
           f"element vertex {len(verts)}\nproperty float x\nproperty float y\nproperty float z\n"
           f"element face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n")
    rec = np.zeros(len(faces), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    rec["n"] = 3
    rec["i"] = faces
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(verts.tobytes())
        f.write(rec.tobytes())


def sphere_field(n_spheres, level, seed, extent, rmin, rmax, grid=None):
    """n_spheres icospheres: centres jittered on a grid (grid=(nx,ny): XZ plane, resting on y=0) or uniform in a
    cube of half-size `extent`.  Returns verts (float32), faces (int32), centres, radii."""
    rng = PCG32([seed])
    base_v, base_f = icosphere(level)
    centres = np.zeros((n_spheres, 3), np.float32)
    radii = np.zeros(n_spheres, np.float32)
    for i in range(n_spheres):
        r = np.float32(rmin) + np.float32(rmax - rmin) * rng.f32()[0]
        if grid is not None:
            nx, ny = grid
            gx, gz = i % nx, i // nx
            cell = np.float32(2.0 * extent / max(nx, ny))
            jx = (rng.f32()[0] - np.float32(0.5)) * cell * np.float32(0.4)
            jz = (rng.f32()[0] - np.float32(0.5)) * cell * np.float32(0.4)
            centres[i] = [(gx + 0.5) * cell - extent + jx, r, (gz + 0.5) * cell - extent + jz]
        else:
            centres[i] = [(rng.f32()[0] * 2 - 1) * extent, (rng.f32()[0] * 2 - 1) * extent, (rng.f32()[0] * 2 - 1) * extent]
        radii[i] = r
    nv = len(base_v)
    verts = (base_v[None, :, :] * radii[:, None, None].astype(np.float64) + centres[:, None, :].astype(np.float64)).reshape(-1, 3).astype(np.float32)
    faces = (base_f[None, :, :].astype(np.int64) + (np.arange(n_spheres, dtype=np.int64) * nv)[:, None, None]).reshape(-1, 3).astype(np.int32)
    return verts, faces, centres, radii


# ------------------------------------------------------------------------------------------------
# pbrt text helpers


def _fmt(a):
    return " ".join(repr(float(np.float32(x))) for x in np.asarray(a).ravel())


def _mesh(P, idx):
    return f'Shape "trianglemesh" "integer indices" [{" ".join(str(int(i)) for i in np.asarray(idx).ravel())}] "point P" [{_fmt(P)}]\n'


def _quad(p0, p1, p2, p3):
    return _mesh([p0, p1, p2, p3], [0, 1, 2, 0, 2, 3])


def _box(lo, hi, rot_y_deg=0.0, centre=None):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
    if rot_y_deg:
        ctr = (lo + hi) / 2 if centre is None else np.asarray(centre, float)
        a = np.deg2rad(rot_y_deg)
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        c = (c - ctr) @ R.T + ctr
    idx = [0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 2, 3, 7, 2, 7, 6, 1, 2, 6, 1, 6, 5, 3, 0, 4, 3, 4, 7]
    return _mesh(c, idx)


def header(xres, yres, spp, integrator, fov, look, crop=None, extra_film=""):
    eye, at, up = look
    cropw = "" if crop is None else f' "float cropwindow" [{_fmt(crop)}]'
    return (f"LookAt {_fmt(eye)}  {_fmt(at)}  {_fmt(up)}\n"
            f'Camera "perspective" "float fov" [{fov}]\n'
            f'Film "image" "integer xresolution" [{xres}] "integer yresolution" [{yres}] "string filename" "out.png"{cropw}{extra_film}\n'
            f'Sampler "02sequence" "integer pixelsamples" [{spp}]\n'
            'PixelFilter "box"\n'
            f"{integrator}\n")


# ------------------------------------------------------------------------------------------------
# C1: Cornell box


def cornell_box(xres=512, yres=512, spp=16, integrator=None, crop=None):
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "uniform"'
    s = header(xres, yres, spp, integrator, 39, ([278, 273, -800], [278, 273, 0], [0, 1, 0]), crop)
    s += "WorldBegin\n"
    white, red, green = "0.73 0.73 0.73", "0.65 0.05 0.05", "0.12 0.45 0.15"
    s += f'Material "matte" "rgb Kd" [{white}]\n'
    s += _quad([0, 0, 0], [556, 0, 0], [556, 0, 559.2], [0, 0, 559.2])              # floor
    s += _quad([0, 548.8, 0], [0, 548.8, 559.2], [556, 548.8, 559.2], [556, 548.8, 0])  # ceiling
    s += _quad([0, 0, 559.2], [556, 0, 559.2], [556, 548.8, 559.2], [0, 548.8, 559.2])  # back
    s += f'Material "matte" "rgb Kd" [{green}]\n'
    s += _quad([0, 0, 0], [0, 0, 559.2], [0, 548.8, 559.2], [0, 548.8, 0])             # right (x=0)
    s += f'Material "matte" "rgb Kd" [{red}]\n'
    s += _quad([556, 0, 0], [556, 548.8, 0], [556, 548.8, 559.2], [556, 0, 559.2])     # left (x=556)
    s += f'Material "matte" "rgb Kd" [{white}]\n'
    s += _box([130, 0, 65], [295, 165, 230], rot_y_deg=-18)                           # short box
    s += _box([265, 0, 295], [430, 330, 460], rot_y_deg=15)                           # tall box
    s += "AttributeBegin\n"
    s += 'AreaLightSource "diffuse" "rgb L" [17 12 4]\n'
    s += 'Material "matte" "rgb Kd" [0 0 0]\n'
    # ceiling light, facing down (normal = -y)
    s += _quad([213, 548.7, 227], [343, 548.7, 227], [343, 548.7, 332], [213, 548.7, 332])
    s += "AttributeEnd\nWorldEnd\n"
    return s


# ------------------------------------------------------------------------------------------------
# C2: balls (64 spheres + ground disk)

_MATERIALS = [
    'Material "matte" "rgb Kd" [{c}]',
    'Material "plastic" "rgb Kd" [{c}] "rgb Ks" [0.3 0.3 0.3] "float roughness" [0.05]',
    'Material "metal" "float roughness" [0.02]',
    'Material "glass" "float index" [1.5]',
    'Material "mirror" "rgb Kr" [0.9 0.9 0.9]',
]


# the materials that reuse the path's lobes (uber, substrate, translucent, mix incl. a nested mix and a
# glass child, which is the one case where allow_multiple_lobes changes the lobe list)
_MATERIALS_EXT = [
    'Material "uber" "rgb Kd" [{c}] "rgb Ks" [0.3 0.3 0.3] "rgb Kr" [0.1 0.1 0.1] "float roughness" [0.08]',
    'Material "uber" "rgb Kd" [{c}] "rgb opacity" [0.6 0.6 0.6] "rgb Kt" [0.2 0.2 0.2] "float index" [1.3]',
    'Material "substrate" "rgb Kd" [{c}] "rgb Ks" [0.2 0.2 0.2] "float uroughness" [0.05] "float vroughness" [0.2]',
    'Material "translucent" "rgb Kd" [{c}] "rgb Ks" [0.2 0.2 0.2] "rgb reflect" [0.6 0.6 0.6] "rgb transmit" [0.4 0.4 0.4]',
    'NamedMaterial "mix_pm"',
    'NamedMaterial "mix_gs"',
    'NamedMaterial "mix_nested"',
    'Material "matte" "rgb Kd" [{c}] "float sigma" [20]',
]
_MATERIALS_EXT_PREAMBLE = """MakeNamedMaterial "pl" "string type" "plastic" "rgb Kd" [0.3 0.5 0.2] "float roughness" [0.1]
MakeNamedMaterial "mi" "string type" "mirror"
MakeNamedMaterial "gl" "string type" "glass"
MakeNamedMaterial "su" "string type" "substrate" "rgb Kd" [0.5 0.2 0.2]
MakeNamedMaterial "mt" "string type" "matte" "rgb Kd" [0.2 0.3 0.7] "float sigma" [30]
MakeNamedMaterial "mix_pm" "string type" "mix" "string namedmaterial1" "pl" "string namedmaterial2" "mi" "rgb amount" [0.7 0.7 0.7]
MakeNamedMaterial "mix_gs" "string type" "mix" "string namedmaterial1" "gl" "string namedmaterial2" "su"
MakeNamedMaterial "mix_nested" "string type" "mix" "string namedmaterial1" "mix_pm" "string namedmaterial2" "mt" "rgb amount" [0.3 0.5 0.7]
"""


def balls_ext(**kw):
    """The balls scene with the SURVEY 8f rank-1 materials on the spheres."""
    return balls(materials=_MATERIALS_EXT, preamble=_MATERIALS_EXT_PREAMBLE, **kw)


def balls(xres=1024, yres=768, spp=64, integrator=None, n_side=8, crop=None, materials=None, preamble=""):
    materials = _MATERIALS if materials is None else materials
    if integrator is None:
        integrator = 'Integrator "whitted" "integer maxdepth" [5]'
    s = header(xres, yres, spp, integrator, 40, ([0, 7.5, -13], [0, 0.3, 0], [0, 1, 0]), crop)
    s += "WorldBegin\n" + preamble
    s += 'LightSource "point" "rgb I" [220 220 220] "point from" [-6 9 -6]\n'
    s += 'LightSource "point" "rgb I" [120 110 100] "point from" [7 6 -3]\n'
    s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [12 12 12]\nTranslate 0 6 2\nMaterial "matte" "rgb Kd" [0 0 0]\nShape "sphere" "float radius" [0.6]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.55 0.55 0.5]\nRotate -90 1 0 0\nShape "disk" "float radius" [20]\nAttributeEnd\n'
    rng = PCG32([2])
    n = n_side * n_side
    for i in range(n):
        r = np.float32(0.3) + np.float32(0.2) * rng.f32()[0]
        gx, gz = i % n_side, i // n_side
        x = (gx - (n_side - 1) / 2) * 1.25 + float(rng.f32()[0] - 0.5) * 0.3
        z = (gz - (n_side - 1) / 2) * 1.25 + float(rng.f32()[0] - 0.5) * 0.3
        col = f"{0.2 + 0.7 * float(rng.f32()[0]):.4f} {0.2 + 0.7 * float(rng.f32()[0]):.4f} {0.2 + 0.7 * float(rng.f32()[0]):.4f}"
        s += "AttributeBegin\n" + materials[i % len(materials)].format(c=col) + f"\nTranslate {x:.5f} {float(r):.5f} {z:.5f}\n"
        s += f'Shape "sphere" "float radius" [{float(r):.5f}]\nAttributeEnd\n'
    s += "WorldEnd\n"
    return s


# ------------------------------------------------------------------------------------------------
# textures and bump mapping

def write_pfm(path, img):
    """img: (H, W, 3) or (H, W) float32, row 0 at the top (PFM stores rows bottom-to-top)."""
    img = np.asarray(img, np.float32)
    with open(path, "wb") as f:
        f.write((b"PF" if img.ndim == 3 else b"Pf") + b"\n%d %d\n-1.0\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img[::-1]).tobytes())


def write_png8(path, img):
    """img: (H, W, 3) or (H, W) uint8; deflate-compressed, filter type 0 on even rows and 'sub' on odd rows."""
    import struct
    import zlib
    img = np.asarray(img, np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else 3
    rows = img.reshape(h, w * ch)
    raw = bytearray()
    for y in range(h):
        if y % 2 == 0:
            raw += b"\x00" + rows[y].tobytes()
        else:
            d = rows[y].astype(np.int16)
            d[ch:] -= rows[y, :-ch].astype(np.int16)
            raw += b"\x01" + (d & 0xff).astype(np.uint8).tobytes()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if ch == 1 else 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(bytes(raw), 6)) + chunk(b"IEND", b""))


def write_hdr(path, img, rle=True):
    """Radiance RGBE (.hdr), rows top-to-bottom ("-Y h +X w"); rle=True writes new-style run-length scanlines (needs 8 <= w < 32768).
    Returns the image as a decoder sees it (RGBE quantisation applied), (H, W, 3) float32."""
    img = np.asarray(img, np.float32)
    h, w, _ = img.shape
    m = img.max(-1)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0).astype(np.int32)       # m in [2^(e-1), 2^e)
    scale = np.where(m > 1e-32, np.exp2((8 - e).astype(np.float32)), 0).astype(np.float32)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(np.floor(img * scale[..., None]), 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w))
        for y in range(h):
            if not rle:
                f.write(rgbe[y].tobytes())
                continue
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for c in range(4):
                row = rgbe[y, :, c]
                x = 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]:
                        run += 1
                    if run >= 4:
                        f.write(bytes([128 + run, int(row[x])]))
                        x += run
                    else:
                        lit = 1
                        while x + lit < w and lit < 128 and not (x + lit + 3 < w and row[x + lit] == row[x + lit + 1] == row[x + lit + 2] == row[x + lit + 3]):
                            lit += 1
                        f.write(bytes([lit]) + row[x:x + lit].tobytes())
                        x += lit
    dec = rgbe[..., :3].astype(np.float32) * np.exp2(rgbe[..., 3:4].astype(np.float32) - 136.0)
    dec[rgbe[..., 3] == 0] = 0
    return dec.astype(np.float32)


def write_exr(path, img, pixel_type="half", compression="zip", decreasing_y=False):
    """Single-part scan-line OpenEXR with channels B, G, R (alphabetical, as the format requires): pixel_type "half" | "float",
    compression "none" | "rle" | "zips" | "zip".  Returns the image as a reader sees it (half rounding applied), (H, W, 3) float32."""
    import struct
    import zlib
    img = np.asarray(img, np.float32)
    h, w, _ = img.shape
    dt = np.float16 if pixel_type == "half" else np.float32
    code = {"none": 0, "rle": 1, "zips": 2, "zip": 3}[compression]
    lines_per_block = 16 if code == 3 else 1

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBxxxii", 1 if pixel_type == "half" else 2, 0, 1, 1) for n in ("B", "G", "R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = (struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([code])) + attr("dataWindow", "box2i", box) +
           attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", bytes([1 if decreasing_y else 0])) + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) +
           attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")

    def rle(b):
        out, i = bytearray(), 0
        while i < len(b):
            run = 1
            while i + run < len(b) and run < 127 and b[i + run] == b[i]:
                run += 1
            if run >= 3:
                out += bytes([run - 1, b[i]])
                i += run
            else:
                j = i
                while j < len(b) and j - i < 127 and not (j + 2 < len(b) and b[j] == b[j + 1] == b[j + 2]):
                    j += 1
                out += bytes([(256 - (j - i)) & 255]) + bytes(b[i:j])
                i = j
        return bytes(out)
    blocks = []
    for y0 in range(0, h, lines_per_block):
        rows = img[y0:y0 + lines_per_block]
        raw = b"".join(rows[l, :, c].astype(dt).tobytes() for l in range(rows.shape[0]) for c in (2, 1, 0))
        data = raw
        if code:
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
            t[1:] = (t[1:] - t[:-1] + 128 + 256) & 255
            pre = t.astype(np.uint8).tobytes()
            comp = rle(pre) if code == 1 else zlib.compress(pre, 6)
            data = comp if len(comp) < len(raw) else raw
        blocks.append((y0, data))
    if decreasing_y:
        blocks = blocks[::-1]
    table_pos = len(hdr)
    pos = table_pos + 8 * len(blocks)
    offsets = {}
    body = bytearray()
    for y0, data in blocks:
        offsets[y0] = pos + len(body)
        body += struct.pack("<ii", y0, len(data)) + data
    with open(path, "wb") as f:
        f.write(hdr + b"".join(struct.pack("<Q", offsets[y0]) for y0 in sorted(offsets)) + bytes(body))
    return img.astype(dt).astype(np.float32)


def texture_images(out_dir):
    """The two harness textures: a 48x20 (non power-of-two) RGB PFM and a 16x16 greyscale PNG."""
    os.makedirs(out_dir, exist_ok=True)
    y, x = np.mgrid[0:20, 0:48].astype(np.float32)
    rgb = np.stack([0.5 + 0.5 * np.sin(x * 0.4), 0.5 + 0.5 * np.cos(y * 0.7 + x * 0.1), ((x.astype(int) // 6 + y.astype(int) // 5) % 2) * 0.8 + 0.1], -1)
    write_pfm(os.path.join(out_dir, "tex_rgb.pfm"), rgb.astype(np.float32))
    y, x = np.mgrid[0:16, 0:16]
    g = (127.5 + 127.5 * np.sin(x * 0.9) * np.cos(y * 0.6)).astype(np.uint8)
    write_png8(os.path.join(out_dir, "tex_small.png"), g)


_TEXTURES_PREAMBLE = """
Texture "checks" "spectrum" "checkerboard" "float uscale" [8] "float vscale" [8] "rgb tex1" [0.8 0.8 0.8] "rgb tex2" [0.1 0.1 0.3]
Texture "img" "spectrum" "imagemap" "string filename" "tex_rgb.pfm" "float uscale" [2] "float vscale" [2]
Texture "imgtri" "spectrum" "imagemap" "string filename" "tex_rgb.pfm" "bool trilinear" "true" "string wrap" "clamp" "float scale" [0.9]
Texture "png" "spectrum" "imagemap" "string filename" "tex_small.png" "string wrap" "black" "float maxanisotropy" [4]
Texture "bumpf" "float" "fbm" "integer octaves" [4] "float omega" [0.6]
Texture "bumpimg" "float" "imagemap" "string filename" "tex_small.png" "float scale" [0.05] "float uscale" [3] "float vscale" [3]
Texture "fbms" "spectrum" "fbm" "integer octaves" [3]
Texture "rough" "float" "mix" "float tex1" [0.02] "float tex2" [0.3] "texture amount" "bumpimg"
Texture "uvt" "spectrum" "uv" "float uscale" [3] "float vscale" [2] "float udelta" [0.25]
Texture "sc" "spectrum" "scale" "texture tex1" "checks" "texture tex2" "uvt"
Texture "mx" "spectrum" "mix" "texture tex1" "img" "rgb tex2" [0.9 0.2 0.2] "float amount" [0.3]
Texture "planar" "spectrum" "checkerboard" "string mapping" "planar" "vector v1" [0.5 0 0] "vector v2" [0 0 0.5] "rgb tex1" [0.6 0.6 0.55] "texture tex2" "fbms"
Texture "ground_aa" "spectrum" "checkerboard" "float uscale" [24] "float vscale" [16] "string aamode" "none" "texture tex1" "png" "rgb tex2" [0.3 0.5 0.3]
Texture "kconst" "spectrum" "constant" "rgb value" [0.4 0.5 0.6]
Texture "fconst" "float" "constant" "float value" [0.2]
MakeNamedMaterial "t_matte" "string type" "matte" "texture Kd" "checks"
MakeNamedMaterial "t_plastic" "string type" "plastic" "texture Kd" "img" "texture roughness" "rough"
MakeNamedMaterial "t_mix" "string type" "mix" "string namedmaterial1" "t_matte" "string namedmaterial2" "t_plastic" "texture amount" "checks"
"""

_MATERIALS_TEXTURED = [
    'Material "matte" "texture Kd" "checks"',
    'Material "plastic" "texture Kd" "img" "texture roughness" "rough" "rgb Ks" [0.3 0.3 0.3]',
    'Material "matte" "texture Kd" "mx" "texture bumpmap" "bumpf" "float sigma" [20]',
    'Material "glass" "texture bumpmap" "bumpimg"',
    'Material "mirror" "texture Kr" "sc"',
    'Material "uber" "texture Kd" "imgtri" "texture opacity" "checks" "rgb Kr" [0.1 0.1 0.1]',
    'Material "metal" "texture roughness" "rough" "texture bumpmap" "bumpf"',
    'Material "substrate" "texture Kd" "png" "texture uroughness" "fconst"',
    'Material "translucent" "texture Kd" "uvt" "texture reflect" "kconst"',
    'NamedMaterial "t_mix"',
    'Material "matte" "rgb Kd" [{c}]',
    'Material "plastic" "texture Kd" "fbms" "texture bumpmap" "bumpimg"',
]


def balls_textured(out_dir, xres=1024, yres=768, spp=64, integrator=None, n_side=5, crop=None, lens=False, absolute_paths=False):
    """The balls scene with textured materials (every texture class, both mappings, bump maps), a planar-mapped checkerboard
    ground disk and a uv-mapped, image-textured triangle quad seen at a grazing angle (anisotropic EWA lookups).  Writes the
    texture images into out_dir; parse with search_dir=out_dir."""
    texture_images(out_dir)
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5]'
    s = header(xres, yres, spp, integrator, 40, ([0, 5.5, -9.5], [0, 0.3, 0], [0, 1, 0]), crop)
    if lens:
        s = s.replace('"float fov" [40]', '"float fov" [40] "float lensradius" [0.05] "float focaldistance" [10]')
    pre = _TEXTURES_PREAMBLE
    if absolute_paths:                                                # parse without a search directory
        pre = pre.replace('"tex_rgb.pfm"', f'"{os.path.join(os.path.abspath(out_dir), "tex_rgb.pfm")}"').replace(
            '"tex_small.png"', f'"{os.path.join(os.path.abspath(out_dir), "tex_small.png")}"')
    s += "WorldBegin\n" + pre
    s += 'LightSource "point" "rgb I" [220 220 220] "point from" [-6 9 -6]\n'
    s += 'LightSource "infinite" "rgb L" [0.25 0.3 0.4]\n'
    s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [12 12 12]\nTranslate 0 6 2\nMaterial "matte" "rgb Kd" [0 0 0]\nShape "sphere" "float radius" [0.6]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "matte" "texture Kd" "planar"\nRotate -90 1 0 0\nShape "disk" "float radius" [20]\nAttributeEnd\n'
    s += ('AttributeBegin\nMaterial "matte" "texture Kd" "ground_aa"\n'
          'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-6 0.01 2  6 0.01 2  6 0.01 14  -6 0.01 14] "float uv" [0 0 1 0 1 1 0 1]\nAttributeEnd\n')
    s += ('AttributeBegin\nMaterial "plastic" "texture Kd" "img" "texture bumpmap" "bumpimg"\nTranslate -4.5 1.2 4\nRotate 30 0 1 0\n'
          'Shape "cylinder" "float radius" [0.7] "float zmin" [-1] "float zmax" [1]\nAttributeEnd\n')
    rng = PCG32([7])
    n = n_side * n_side
    for i in range(n):
        r = np.float32(0.4) + np.float32(0.2) * rng.f32()[0]
        gx, gz = i % n_side, i // n_side
        x = (gx - (n_side - 1) / 2) * 1.6 + float(rng.f32()[0] - 0.5) * 0.3
        z = (gz - (n_side - 1) / 2) * 1.6 + float(rng.f32()[0] - 0.5) * 0.3
        col = f"{0.2 + 0.7 * float(rng.f32()[0]):.4f} {0.2 + 0.7 * float(rng.f32()[0]):.4f} {0.2 + 0.7 * float(rng.f32()[0]):.4f}"
        s += "AttributeBegin\n" + _MATERIALS_TEXTURED[i % len(_MATERIALS_TEXTURED)].format(c=col) + f"\nTranslate {x:.5f} {float(r):.5f} {z:.5f}\nRotate {37 * i} 0.3 1 0.2\n"
        s += f'Shape "sphere" "float radius" [{float(r):.5f}]\nAttributeEnd\n'
    s += "WorldEnd\n"
    return s


# ------------------------------------------------------------------------------------------------
# every light class and every Shape::sample the lights call

def env_map_image(w=32, h=16):
    """Lat-long environment with strong contrast: a dim sky gradient, a warm horizon band and a small very bright 'sun' patch
    (1000x the sky) — the importance sampling of infinite.rs:143-183 has to find it.  (h, w, 3) float32, row 0 = theta 0."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    sky = np.stack([0.05 + 0.02 * y / h, 0.08 + 0.03 * y / h, 0.15 + 0.1 * (1 - y / h)], -1)
    band = np.exp(-((y - 0.55 * h) / (0.08 * h)) ** 2)[..., None] * np.array([0.6, 0.35, 0.15], np.float32)
    img = (sky + band).astype(np.float32)
    sy, sx = int(0.3 * h), int(0.7 * w)
    img[sy:sy + max(1, h // 16), sx:sx + max(1, w // 16)] = np.array([60.0, 55.0, 45.0], np.float32)
    return img


def lights_zoo(out_dir, xres=96, yres=72, spp=8, integrator=None, env_size=(32, 16), env_name="env.pfm", absolute_paths=False, crop=None):
    """Image-mapped infinite light under a rotated light_to_world (infinite.rs:46-219, distribution2d.rs, mipmap.rs:227-245),
    a distant light (distant.rs:49-87), disk and cylinder area lights (disk.rs:138-154, cylinder.rs:259-278), a two-sided
    and a one-sided triangle area light seen from both sides (diffuse.rs:91-97, mesh.rs:610-634), over matte / plastic / glass /
    mirror / metal receivers.  Writes the environment map into out_dir; parse with search_dir=out_dir."""
    os.makedirs(out_dir, exist_ok=True)
    if env_name.endswith(".hdr"):
        write_hdr(os.path.join(out_dir, env_name), env_map_image(*env_size))
    elif env_name.endswith(".exr"):
        write_exr(os.path.join(out_dir, env_name), env_map_image(*env_size))
    else:
        write_pfm(os.path.join(out_dir, env_name), env_map_image(*env_size))
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5]'
    s = header(xres, yres, spp, integrator, 42, ([0.5, 5.0, -10.5], [0, 1.0, 0], [0, 1, 0]), crop)
    fn = os.path.join(os.path.abspath(out_dir), env_name) if absolute_paths else env_name
    s += "WorldBegin\n"
    s += ('AttributeBegin\nRotate -90 1 0 0\nRotate 35 0 0 1\nLightSource "infinite" "rgb L" [0.9 1.0 1.1] "rgb scale" [1.2 1.2 1.2] '
          f'"string mapname" "{fn}" "integer samples" [2]\nAttributeEnd\n')
    s += 'LightSource "distant" "rgb L" [1.6 1.5 1.3] "point from" [4 9 -3] "point to" [0 0 0]\n'
    # disk light facing down, phimax-clipped annulus
    s += ('AttributeBegin\nAreaLightSource "diffuse" "rgb L" [14 13 11] "integer samples" [2]\nMaterial "matte" "rgb Kd" [0 0 0]\nTranslate -3 4.5 0.5\nRotate 90 1 0 0\n'
          'Shape "disk" "float radius" [0.9] "float innerradius" [0.25] "float phimax" [300]\nAttributeEnd\n')
    # cylinder light lying along x (emits outwards), partial sweep
    s += ('AttributeBegin\nAreaLightSource "diffuse" "rgb L" [9 10 12]\nMaterial "matte" "rgb Kd" [0 0 0]\nTranslate 3.2 2.6 1.5\nRotate 90 0 1 0\n'
          'Shape "cylinder" "float radius" [0.3] "float z_min" [-1.1] "float z_max" [1.1] "float phi_max" [270]\nAttributeEnd\n')
    # two-sided vertical triangle quad between two receivers; one-sided quad facing away from the camera (emits towards +z only)
    s += ('AttributeBegin\nAreaLightSource "diffuse" "rgb L" [8 3 3] "bool twosided" "true"\nMaterial "matte" "rgb Kd" [0 0 0]\n'
          + _quad([-0.6, 0.4, 1.0], [0.6, 0.4, 1.0], [0.6, 1.8, 1.0], [-0.6, 1.8, 1.0]) + 'AttributeEnd\n')
    s += ('AttributeBegin\nAreaLightSource "diffuse" "rgb L" [3 8 3]\nMaterial "matte" "rgb Kd" [0 0 0]\n'
          + _quad([1.6, 0.3, -1.5], [1.6, 1.5, -1.5], [2.8, 1.5, -1.5], [2.8, 0.3, -1.5]) + 'AttributeEnd\n')
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.55 0.55 0.5]\nRotate -90 1 0 0\nShape "disk" "float radius" [14]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.6 0.6 0.65]\n' + _quad([-7, 0, 5], [7, 0, 5], [7, 6, 5], [-7, 6, 5]) + 'AttributeEnd\n'
    receivers = [('Material "matte" "rgb Kd" [0.7 0.3 0.25]', (-2.2, 0.7, -1.0), 0.7), ('Material "plastic" "rgb Kd" [0.2 0.4 0.7] "float roughness" [0.08]', (-0.3, 0.6, -2.2), 0.6),
                 ('Material "glass" "float index" [1.5]', (1.2, 0.65, -3.0), 0.65), ('Material "mirror"', (2.6, 0.8, 3.0), 0.8),
                 ('Material "metal" "float roughness" [0.1]', (-2.4, 0.75, 2.6), 0.75), ('Material "matte" "rgb Kd" [0.5 0.5 0.5] "float sigma" [25]', (0.0, 0.5, 2.8), 0.5)]
    for mat, c, r in receivers:
        s += f'AttributeBegin\n{mat}\nTranslate {c[0]} {c[1]} {c[2]}\nShape "sphere" "float radius" [{r}]\nAttributeEnd\n'
    s += "WorldEnd\n"
    return s


def emissive_mesh_scene(out_dir, level=5, xres=64, yres=48, spp=4, integrator=None):
    """Many-light scene: an emissive icosphere mesh (20 * 4^level triangle lights; level 5 = 20,480) above a ground quad with a matte,
    a plastic and a mirror ball — SpatialLightDistribution (lightdistrib.rs:59-296) over thousands of lights, where only the voxels
    that path vertices fall into can hold a distribution."""
    os.makedirs(out_dir, exist_ok=True)
    v, f = icosphere(level)
    write_ply(os.path.join(out_dir, f"emitter_{level}.ply"), v * 0.8 + np.array([0.0, 3.2, 0.5]), f)
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [4] "string lightsamplestrategy" "spatial"'
    s = header(xres, yres, spp, integrator, 40, ([0, 3.0, -9.0], [0, 1.2, 0], [0, 1, 0]))
    s += "WorldBegin\n"
    s += f'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [6 5.5 5]\nMaterial "matte" "rgb Kd" [0 0 0]\nShape "plymesh" "string filename" "emitter_{level}.ply"\nAttributeEnd\n'
    s += 'Material "matte" "rgb Kd" [0.55 0.55 0.5]\n' + _quad([-8, 0, -8], [8, 0, -8], [8, 0, 8], [-8, 0, 8])
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.7 0.3 0.25]\nTranslate -2 0.8 0\nShape "sphere" "float radius" [0.8]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "plastic" "rgb Kd" [0.2 0.4 0.7] "float roughness" [0.1]\nTranslate 0 0.7 -1.5\nShape "sphere" "float radius" [0.7]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "mirror"\nTranslate 2 0.9 0.5\nShape "sphere" "float radius" [0.9]\nAttributeEnd\n'
    s += "WorldEnd\n"
    return s


# ------------------------------------------------------------------------------------------------
# object instancing (ObjectBegin / ObjectEnd / ObjectInstance)

def _inline_icosphere(level, radius=1.0):
    v, f = icosphere(level)
    return _mesh(np.asarray(v) * radius, f)


def instanced_scene(xres=96, yres=72, spp=8, integrator=None, n_side=4, baked=False, level=1):
    """A "plant" object (icosphere mesh + sphere + cylinder + a mirrored-scale part, mixed materials) and a one-primitive object,
    instanced on a jittered grid under translate / rotate / non-uniform scale / negative scale, over a ground disk with a point
    light, an area light and a constant environment.  baked=True declares the same shapes directly under the composed transforms
    (no instancing): the two scenes render the same image up to fp32 noise."""
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5]'
    s = header(xres, yres, spp, integrator, 40, ([0, 6.5, -12], [0, 0.6, 0], [0, 1, 0]))
    s += "WorldBegin\n"
    s += 'LightSource "point" "rgb I" [160 160 150] "point from" [-5 9 -6]\n'
    s += 'LightSource "infinite" "rgb L" [0.25 0.3 0.4]\n'
    s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [10 10 9]\nTranslate 0 7 1\nMaterial "matte" "rgb Kd" [0 0 0]\nShape "sphere" "float radius" [0.5]\nAttributeEnd\n'
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.5 0.5 0.45]\nRotate -90 1 0 0\nShape "disk" "float radius" [20]\nAttributeEnd\n'
    plant = ('Material "matte" "rgb Kd" [0.2 0.6 0.25]\nAttributeBegin\nTranslate 0 1.1 0\nScale 0.55 0.55 0.55\n' + _inline_icosphere(level) + 'AttributeEnd\n'
             'Material "plastic" "rgb Kd" [0.5 0.3 0.15] "rgb Ks" [0.2 0.2 0.2] "float roughness" [0.1]\n'
             'AttributeBegin\nRotate -90 1 0 0\nShape "cylinder" "float radius" [0.12] "float z_min" [0] "float z_max" [0.8]\nAttributeEnd\n'
             'Material "glass"\nAttributeBegin\nTranslate 0.55 0.35 0\nShape "sphere" "float radius" [0.22]\nAttributeEnd\n'
             'Material "metal" "float roughness" [0.05]\nAttributeBegin\nTranslate -0.5 0.3 0.1\nScale -0.3 0.3 0.3\n' + _inline_icosphere(0) + 'AttributeEnd\n')
    pebble = 'Material "mirror"\nShape "sphere" "float radius" [0.25]\n'
    if not baked:
        s += 'ObjectBegin "plant"\n' + plant + 'ObjectEnd\nObjectBegin "pebble"\n' + pebble + 'ObjectEnd\nObjectBegin "empty"\nObjectEnd\n'
    rng = PCG32([8])
    for i in range(n_side * n_side):
        gx, gz = i % n_side, i // n_side
        x = (gx - (n_side - 1) / 2) * 2.2 + float(rng.f32()[0] - 0.5) * 0.5
        z = (gz - (n_side - 1) / 2) * 2.2 + float(rng.f32()[0] - 0.5) * 0.5
        rot = 360.0 * float(rng.f32()[0])
        sx, sy, sz = (0.7 + 0.8 * float(rng.f32()[0]) for _ in range(3))
        if i % 5 == 3:
            sx = -sx                                         # mirrored instance: swaps handedness
        xf = f"Translate {x:.5f} 0 {z:.5f}\nRotate {rot:.4f} 0 1 0\nScale {sx:.5f} {sy:.5f} {sz:.5f}\n"
        s += "AttributeBegin\n" + xf + (plant if baked else 'ObjectInstance "plant"\n') + "AttributeEnd\n"
        s += f"AttributeBegin\nTranslate {x + 0.9:.5f} 0.25 {z - 0.7:.5f}\n" + (pebble if baked else 'ObjectInstance "pebble"\nObjectInstance "empty"\n') + "AttributeEnd\n"
    s += "WorldEnd\n"
    return s


# ------------------------------------------------------------------------------------------------
# C3 / C5: icosphere fields written as PLY; C4: ray-batch field


def sphere_field_scene(out_dir, name, n_spheres, level, seed, xres, yres, spp, integrator, grid, materials_cycle=False, n_area_lights=1,
                       env=True, crop=None):
    """Writes <out_dir>/<name>*.ply and returns the pbrt text (the scene must be parsed with search_dir=out_dir)."""
    os.makedirs(out_dir, exist_ok=True)
    extent = 2.2 * max(grid)
    verts, faces, centres, radii = sphere_field(n_spheres, level, seed, extent, 0.7, 1.3, grid=grid)
    s = header(xres, yres, spp, integrator, 38, ([0, 0.9 * extent, -2.1 * extent], [0, 0.8, -0.1 * extent], [0, 1, 0]), crop)
    s += "WorldBegin\n"
    if env:
        s += 'LightSource "infinite" "rgb L" [0.35 0.4 0.5]\n'
    nv = len(verts) // n_spheres
    nf = len(faces) // n_spheres
    groups = 5 if materials_cycle else 1
    for g in range(groups):
        sel = np.arange(g, n_spheres, groups)
        if len(sel) == 0:
            continue
        v = verts.reshape(n_spheres, nv, 3)[sel].reshape(-1, 3)
        f = (faces.reshape(n_spheres, nf, 3)[sel] - (sel * nv)[:, None, None] + (np.arange(len(sel)) * nv)[:, None, None]).reshape(-1, 3)
        fn = f"{name}_{g}.ply"
        write_ply(os.path.join(out_dir, fn), v, f)
        mat = _MATERIALS[g % 5].format(c="0.6 0.55 0.5") if materials_cycle else 'Material "matte" "rgb Kd" [0.6 0.55 0.5]'
        s += f'AttributeBegin\n{mat}\nShape "plymesh" "string filename" "{fn}"\nAttributeEnd\n'
    e = 1.6 * extent
    s += 'Material "matte" "rgb Kd" [0.5 0.5 0.5]\n' + _quad([-e, 0, -e], [e, 0, -e], [e, 0, e], [-e, 0, e])
    for k in range(n_area_lights):
        cx = (k - (n_area_lights - 1) / 2) * extent * 0.8
        h = 2.2 * extent * 0.5 + 3.0
        q = extent * 0.25
        s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [30 28 25]\nMaterial "matte" "rgb Kd" [0 0 0]\n'
        s += _quad([cx - q, h, -q], [cx + q, h, -q], [cx + q, h, q], [cx - q, h, q]) + "AttributeEnd\n"
    s += "WorldEnd\n"
    return s


def c3_scene(out_dir, integrator=None, level=5, xres=1920, yres=1080, spp=256, crop=None):
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "spatial"'
    return sphere_field_scene(out_dir, "c3", 49, level, 3, xres, yres, spp, integrator, grid=(7, 7), crop=crop)


def c5_scene(out_dir, integrator=None, level=5, xres=3840, yres=2160, spp=1024, n_spheres=244, crop=None):
    if integrator is None:
        integrator = 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "spatial"'
    return sphere_field_scene(out_dir, "c5", n_spheres, level, 6, xres, yres, spp, integrator, grid=(16, 16), materials_cycle=True,
                              n_area_lights=4, crop=crop)


def c4_scene(out_dir, n_spheres=489, level=5, seed=4):
    """Ray-batch field: n_spheres level-`level` icospheres uniformly placed in a cube (489 x 20480 = 10 014 720 tris)."""
    os.makedirs(out_dir, exist_ok=True)
    verts, faces, _, _ = sphere_field(n_spheres, level, seed, 30.0, 1.5, 3.5, grid=None)
    write_ply(os.path.join(out_dir, "c4.ply"), verts, faces)
    s = header(64, 64, 1, 'Integrator "path"', 40, ([0, 0, -120], [0, 0, 0], [0, 1, 0]))
    s += 'WorldBegin\nMaterial "matte"\nShape "plymesh" "string filename" "c4.ply"\nWorldEnd\n'
    return s


def ray_batch(n, world_lo, world_hi, seed=5, any_hit=False, first=0):
    """ray i uses PCG32 stream (seed*2^32 + i): origin uniform in the world bounds grown 5 %,
    direction uniform on the sphere (closest-hit, t_max = inf) or the segment to a second uniform point
    (any-hit, d = p1 - p0, t_max = 1 - 1e-4).  Returns (n, 8) float32 {o, tmax, d, tag}."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    rng = PCG32((np.uint64(seed) << np.uint64(32)) + idx)
    lo = np.asarray(world_lo, np.float32)
    hi = np.asarray(world_hi, np.float32)
    c, h = (lo + hi) * np.float32(0.5), (hi - lo) * np.float32(0.5 * 1.05)
    rays = np.zeros((n, 8), np.float32)
    for k in range(3):
        rays[:, k] = c[k] + (rng.f32() * np.float32(2) - np.float32(1)) * h[k]
    if any_hit:
        for k in range(3):
            p1 = c[k] + (rng.f32() * np.float32(2) - np.float32(1)) * h[k]
            rays[:, 4 + k] = p1 - rays[:, k]
        rays[:, 3] = np.float32(1.0 - 1e-4)
    else:
        rays[:, 4:7] = uniform_sample_sphere(rng.f32(), rng.f32())
        rays[:, 3] = np.inf
    rays[:, 7] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
    return rays
