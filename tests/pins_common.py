"""Oracle-independent pins of the shading code ("missing" item 4).

The reference holds no fixture for radiance, so a misreading of the Rust shared by the oracle and the device would pass every
device-vs-oracle comparison.  The checks here do not depend on anyone's reading of the Rust: they are properties every correct
BSDF / light implementation has, plus closed-form values computed in numpy from the scene text:

  BSDF   * the pdf integrates to at most one over the sphere, and sample_f is distributed as pdf says (chi-square on a
           (cos theta, phi) histogram, samples that fail counted against the pdf's missing mass)
         * sample_f returns the f and pdf that f() / pdf() give for the direction it returns
         * energy conservation (white furnace): the albedo E[f |cos| / pdf] is <= 1 for every material whose reflectances are <= 1,
           equals Kd for a Lambertian surface and Kr for a mirror; f = Kd / pi exactly for Lambert
         * reciprocity f(wo, wi) = f(wi, wo) of the reflection lobes
  lights * sample_li's pdf equals pdf_li of the direction it returns
         * closed forms: point I / r^2, distant L along the normalised direction, triangle / disk area pdf d^2 / (|cos| A),
           sphere cone pdf 1 / (2 pi (1 - cos theta_max))
         * the environment light's pdf integrates to one over the sphere, and E[Li / pdf] over sample_li equals the quadrature of
           le over the sphere (wrong Jacobians, transposed maps or a mis-rotated frame fail this)

Every function takes a `prober` — oracle.binding.OracleScene or rustracer_b200.device.Device, which expose the same
bsdf_probe / light_probe — so the CPU suite pins the oracle and the GPU suite pins the device with the same code."""
import numpy as np

BSDF_REFLECTION, BSDF_TRANSMISSION, BSDF_DIFFUSE, BSDF_GLOSSY, BSDF_SPECULAR = 1, 2, 4, 8, 16
BSDF_ALL = 31
NON_SPECULAR = BSDF_ALL & ~BSDF_SPECULAR

# name -> (Material directive, properties).  reflect_only: no transmission lobe (reciprocity applies); conserving: albedo <= 1
MATERIALS = [
    ("lambert", 'Material "matte" "rgb Kd" [0.6 0.5 0.4]', dict(reflect_only=True, kd=(0.6, 0.5, 0.4))),
    ("oren_nayar", 'Material "matte" "rgb Kd" [0.7 0.7 0.7] "float sigma" [30]', dict(reflect_only=True)),
    ("plastic", 'Material "plastic" "rgb Kd" [0.5 0.4 0.3] "rgb Ks" [0.4 0.4 0.4] "float roughness" [0.15]', dict(reflect_only=True)),
    ("metal", 'Material "metal" "float roughness" [0.2]', dict(reflect_only=True)),
    ("metal_aniso", 'Material "metal" "float uroughness" [0.3] "float vroughness" [0.08]', dict(reflect_only=True)),
    ("glass_smooth", 'Material "glass" "float index" [1.5]', dict(specular=True)),
    ("glass_rough", 'Material "glass" "float index" [1.5] "float uroughness" [0.25] "float vroughness" [0.25]', dict(transmission_index=1.5)),
    ("mirror", 'Material "mirror" "rgb Kr" [0.9 0.8 0.7]', dict(specular=True, kr=(0.9, 0.8, 0.7))),
    ("uber", 'Material "uber" "rgb Kd" [0.4 0.4 0.4] "rgb Ks" [0.3 0.3 0.3] "rgb Kr" [0.1 0.1 0.1] "float roughness" [0.2]', dict(reflect_only=True)),
    ("substrate", 'Material "substrate" "rgb Kd" [0.5 0.4 0.3] "rgb Ks" [0.3 0.3 0.3] "float uroughness" [0.2] "float vroughness" [0.3]', dict(reflect_only=True)),
    # diffuse + glossy, reflection + transmission: its pdf averages four lobes, so the microfacet-transmission quirk below cannot be
    # masked from outside (no chi-square); translucent_glossy (Kd = 0) has the two microfacet lobes only
    ("translucent", 'Material "translucent" "rgb Kd" [0.5 0.5 0.5] "rgb Ks" [0.3 0.3 0.3] "rgb reflect" [0.5 0.5 0.5] "rgb transmit" [0.5 0.5 0.5] "float roughness" [0.2]', dict(no_chi2=True)),
    ("translucent_glossy", 'Material "translucent" "rgb Kd" [0 0 0] "rgb Ks" [0.6 0.6 0.6] "float roughness" [0.25]', dict(transmission_index=1.5)),
    # ScaledBxDF::pdf keeps the trait's default cosine pdf whatever it wraps (bsdf/bxdf.rs:48-71): a reference quirk, pinned as such
    ("mix", 'NamedMaterial "mixed"', dict(reflect_only=True, cosine_pdf=True)),
]
_PREAMBLE = ('MakeNamedMaterial "a" "string type" "plastic" "rgb Kd" [0.3 0.5 0.2] "float roughness" [0.2]\n'
             'MakeNamedMaterial "b" "string type" "matte" "rgb Kd" [0.6 0.3 0.3] "float sigma" [20]\n'
             'MakeNamedMaterial "mixed" "string type" "mix" "string namedmaterial1" "a" "string namedmaterial2" "b" "rgb amount" [0.4 0.4 0.4]\n')


def materials_scene():
    """One small sphere per entry of MATERIALS (shape i carries material MATERIALS[i]) and a point light."""
    s = ('LookAt 0 0 -30  0 0 0  0 1 0\nCamera "perspective" "float fov" [30]\nFilm "image" "integer xresolution" [16] "integer yresolution" [16]\n'
         'Sampler "02sequence" "integer pixelsamples" [1]\nPixelFilter "box"\nIntegrator "path"\nWorldBegin\n' + _PREAMBLE +
         'LightSource "point" "rgb I" [1 1 1] "point from" [0 20 0]\n')
    for i, (_, mat, _) in enumerate(MATERIALS):
        s += f'AttributeBegin\n{mat}\nTranslate {2.5 * (i - 6.5)} 0 0\nShape "sphere" "float radius" [1]\nAttributeEnd\n'
    return s + "WorldEnd\n"


def material_rows(scene):
    """Material row of each MATERIALS entry (rt_shape.material of its sphere)."""
    ir = scene.ir
    assert ir.n_shapes == len(MATERIALS)
    return {name: int(ir.shapes[i].material) for i, (name, _, _) in enumerate(MATERIALS)}


def _dir(cos_t, phi):
    s = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
    return np.stack([s * np.cos(phi), s * np.sin(phi), cos_t], -1).astype(np.float32)


def sphere_grid(n_cos, n_phi):
    """Midpoint grid over the sphere in (cos theta, phi): directions (n_cos * n_phi, 3) and the solid angle of one cell."""
    c = (np.arange(n_cos) + 0.5) / n_cos * 2.0 - 1.0
    p = (np.arange(n_phi) + 0.5) / n_phi * 2.0 * np.pi
    cc, pp = np.meshgrid(c, p, indexing="ij")
    return _dir(cc.ravel(), pp.ravel()), 4.0 * np.pi / (n_cos * n_phi)


def transmission_valid(wo, wi, index):
    """The reference's MicrofacetTransmission::pdf / f (bsdf/microfacet.rs:126-170, :215-229 — pbrt-v3's, before its later fixes)
    test nothing but `same_hemisphere`: they are non-zero for every (wo, wi) on opposite sides, including pairs no refraction
    produces — (1) the generalised half vector has wo and wi on the SAME side (a back-facing microfacet), or (2) the microfacet
    normal that refracts wo into wi points below the macro-surface (D is even in wh, so it is evaluated anyway).  sample_f never
    returns such a pair (sample_wh draws visible normals of the upper hemisphere, refract does the rest), hence the integral of the
    pdf over the sphere exceeds one (measured 1.23 / 1.50 / 1.86 for alpha ~ 0.4 at cos theta_o = 0.95 / 0.6 / 0.25).  The oracle and
    the device both keep that behaviour; it biases nothing as long as pdf is only evaluated at sampled directions, but it enters
    the MIS weight of light samples seen through a rough dielectric.  This mask selects the reachable pairs; on them the pdf is the
    true density of sample_f (checked independently in float64 numpy to 1e-6)."""
    wo = np.asarray(wo, np.float64); wi = np.asarray(wi, np.float64)
    eta = np.where(wo[..., 2] > 0, index, 1.0 / index)[..., None]
    wh = wo + wi * eta
    wh /= np.maximum(np.linalg.norm(wh, axis=-1, keepdims=True), 1e-30)
    a, b = (wo * wh).sum(-1), (wi * wh).sum(-1)
    m_z = np.where(a > 0, wh[..., 2], -wh[..., 2])                   # the microfacet normal on wo's side of the facet
    reachable = (a * b < 0) & (m_z * wo[..., 2] > 0)
    return reachable | (wo[..., 2] * wi[..., 2] > 0)                # the reflection hemisphere is not affected


WOS = [_dir(np.float64(c), np.float64(p)) for c, p in ((0.95, 0.3), (0.6, 2.0), (0.25, 4.0))]


def check_pdf_and_chi2(prober, row, allow=True, n_samples=200_000, bins=(16, 32), sub=6, seed=3, transmission_index=None):
    """For each wo: integral of the non-specular pdf over the sphere <= 1 (+ quadrature slack), and the histogram of sample_f's
    directions follows the bin integrals of pdf.  Returns [(pdf_mass, chi2, dof)] per wo."""
    rng = np.random.default_rng(seed)
    n_cos, n_phi = bins
    grid, cell = sphere_grid(n_cos * sub, n_phi * sub)
    out = []
    for wo in WOS:
        wo_b = np.broadcast_to(wo, grid.shape)
        pr = prober.bsdf_probe(row, wo_b, grid, np.zeros((len(grid), 2), np.float32), allow, NON_SPECULAR)
        pdf = pr["pdf"].astype(np.float64)
        if transmission_index is not None:
            pdf = pdf * transmission_valid(wo_b, grid, transmission_index)
        pdf = pdf.reshape(n_cos, sub, n_phi, sub)
        expected_frac = pdf.sum((1, 3)) * cell                     # integral of pdf over each histogram bin
        mass = expected_frac.sum()
        u = rng.random((n_samples, 2)).astype(np.float32)
        wo_s = np.broadcast_to(wo, (n_samples, 3))
        sm = prober.bsdf_probe(row, wo_s, wo_s, u, allow, NON_SPECULAR)
        ok = (sm["spdf"] > 0) & (np.abs(sm["sf"]).sum(1) > 0)
        wi = sm["swi"][ok].astype(np.float64)
        wi /= np.linalg.norm(wi, axis=1, keepdims=True)
        ic = np.clip(((wi[:, 2] + 1.0) * 0.5 * n_cos).astype(int), 0, n_cos - 1)
        ph = np.arctan2(wi[:, 1], wi[:, 0])
        ph[ph < 0] += 2 * np.pi
        ip = np.clip((ph / (2 * np.pi) * n_phi).astype(int), 0, n_phi - 1)
        obs = np.zeros((n_cos, n_phi))
        np.add.at(obs, (ic, ip), 1.0)
        # directions where the pdf is positive but f is black are rejected by callers exactly like failed samples: they sit in the
        # "lost" bin together with the pdf's missing mass
        exp = expected_frac * n_samples
        keep = exp >= 10.0
        lost_obs = n_samples - obs[keep].sum()
        lost_exp = n_samples - exp[keep].sum()
        o = np.append(obs[keep], lost_obs)
        e = np.append(exp[keep], max(lost_exp, 1e-9))
        big = e >= 10.0
        chi2 = float((((o - e) ** 2) / e)[big].sum())
        out.append((float(mass), chi2, int(big.sum()) - 1))
    return out


def check_sample_f_consistency(prober, row, allow=True, n=4000, seed=4, flags=BSDF_ALL):
    """sample_f's (f, pdf) equal f() / pdf() at the direction it returns, for non-specular samples.  Returns the worst relative errors."""
    rng = np.random.default_rng(seed)
    wo = _dir(rng.uniform(-1, 1, n), rng.uniform(0, 2 * np.pi, n))
    u = rng.random((n, 2)).astype(np.float32)
    s = prober.bsdf_probe(row, wo, wo, u, allow, flags)
    ok = (s["spdf"] > 0) & ((s["sflags"] & BSDF_SPECULAR) == 0)
    if not ok.any():
        return 0.0, 0.0, 0
    e = prober.bsdf_probe(row, wo[ok], s["swi"][ok], u[ok], allow, flags)
    # when the chosen lobe is not specular, Bsdf::sample_f re-evaluates f over all matching lobes and averages the pdfs
    # (bsdf/mod.rs:214-247) — the same sums f() and pdf() make
    rel_f = np.abs(e["f"] - s["sf"][ok]).max(1) / np.maximum(np.abs(s["sf"][ok]).max(1), 1e-3)
    rel_p = np.abs(e["pdf"] - s["spdf"][ok]) / np.maximum(s["spdf"][ok], 1e-3)
    return float(np.quantile(rel_f, 0.999)), float(np.quantile(rel_p, 0.999)), int(ok.sum())


def albedo(prober, row, wo, allow=True, n=100_000, seed=5, flags=BSDF_ALL):
    """Monte-Carlo albedo E[f |cos theta_i| / pdf] per channel through sample_f (specular lobes included)."""
    rng = np.random.default_rng(seed)
    u = rng.random((n, 2)).astype(np.float32)
    wo_b = np.broadcast_to(wo, (n, 3))
    s = prober.bsdf_probe(row, wo_b, wo_b, u, allow, flags)
    ok = s["spdf"] > 0
    w = np.zeros((n, 3))
    w[ok] = s["sf"][ok].astype(np.float64) * np.abs(s["swi"][ok, 2:3].astype(np.float64)) / s["spdf"][ok, None].astype(np.float64)
    return w.mean(0), w.std(0) / np.sqrt(n)


def check_reciprocity(prober, row, allow=True, n=4000, seed=6):
    rng = np.random.default_rng(seed)
    wo = _dir(rng.uniform(0.05, 1, n), rng.uniform(0, 2 * np.pi, n))
    wi = _dir(rng.uniform(0.05, 1, n), rng.uniform(0, 2 * np.pi, n))
    z = np.zeros((n, 2), np.float32)
    a = prober.bsdf_probe(row, wo, wi, z, allow, NON_SPECULAR)["f"].astype(np.float64)
    b = prober.bsdf_probe(row, wi, wo, z, allow, NON_SPECULAR)["f"].astype(np.float64)
    return float(np.quantile(np.abs(a - b).max(1) / np.maximum(np.abs(a).max(1), 1e-3), 0.999))


# ---- lights -------------------------------------------------------------------------------------------------------------------------

def lights_scene(out_dir):
    """One light of every class at known places (closed forms below): point, distant, infinite (image, rotated), a triangle,
    a two-sided triangle, a disk, a sphere and a cylinder area light.  Returns (text, dict name -> light row)."""
    import os
    from rustracer_b200 import scenes
    os.makedirs(out_dir, exist_ok=True)
    scenes.write_pfm(os.path.join(out_dir, "env_pins.pfm"), scenes.env_map_image(32, 16))
    s = ('LookAt 0 0 -30  0 0 0  0 1 0\nCamera "perspective" "float fov" [30]\nFilm "image" "integer xresolution" [16] "integer yresolution" [16]\n'
         'Sampler "02sequence" "integer pixelsamples" [1]\nPixelFilter "box"\nIntegrator "path"\nWorldBegin\n'
         'LightSource "point" "rgb I" [10 20 30] "point from" [1 5 2]\n'
         'LightSource "distant" "rgb L" [2 3 4] "point from" [0 10 0] "point to" [3 0 4]\n'
         'AttributeBegin\nRotate -90 1 0 0\nRotate 40 0 0 1\nLightSource "infinite" "rgb L" [1 1 1] "string mapname" "env_pins.pfm"\nAttributeEnd\n'
         'Material "matte"\n'
         'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [5 5 5]\n'
         'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [-1 4 -1  1 4 -1  0 4 1]\nAttributeEnd\n'
         'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [4 4 4] "bool twosided" "true"\n'
         'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [5 3 -1  7 3 -1  6 3 1]\nAttributeEnd\n'
         'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [3 3 3]\nTranslate -5 4 0\nRotate 90 1 0 0\nShape "disk" "float radius" [0.8]\nAttributeEnd\n'
         'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [6 6 6]\nTranslate 0 5 6\nShape "sphere" "float radius" [0.7]\nAttributeEnd\n'
         'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [2 2 2]\nTranslate 0 4 -6\nShape "cylinder" "float radius" [0.4] "float z_min" [-1] "float z_max" [1]\nAttributeEnd\n'
         'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-20 0 -20  20 0 -20  20 0 20  -20 0 20]\n'
         'WorldEnd\n')
    rows = dict(point=0, distant=1, infinite=2, triangle=3, triangle_two_sided=4, disk=5, sphere=6, cylinder=7)
    return s, rows


def ref_points(n, seed=8):
    """Reference points on the ground plane y = 0 (normal +y) in a 6 x 6 patch under the lights."""
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(-3, 3, n), np.zeros(n), rng.uniform(-3, 3, n)], 1)
    return np.concatenate([p, np.tile([0.0, 1.0, 0.0], (n, 1))], 1).astype(np.float32)


def check_light_pdf_consistency(prober, light, n=20_000, seed=9):
    """pdf returned by sample_li vs pdf_li(wi) at the sampled direction: quantile of the relative error, and the valid fraction."""
    rng = np.random.default_rng(seed)
    ref = ref_points(n, seed)
    u = rng.random((n, 2)).astype(np.float32)
    r = prober.light_probe(light, ref, u, np.tile(np.float32([0, 1, 0]), (n, 1)))
    ok = (r["pdf"] > 0) & np.isfinite(r["pdf"])
    rel = np.abs(r["pdf_wi"][ok] - r["pdf"][ok]) / r["pdf"][ok]
    return float(np.quantile(rel, 0.99)), float(ok.mean()), r, ref, u


def env_checks(prober, light, n_theta=256, n_phi=512, n_mc=200_000, seed=10):
    """Environment light: (integral of pdf_li over the sphere, quadrature of le.y over the sphere, E[Li.y / pdf] over sample_li)."""
    th = (np.arange(n_theta) + 0.5) / n_theta * np.pi
    ph = (np.arange(n_phi) + 0.5) / n_phi * 2 * np.pi
    tt, pp = np.meshgrid(th, ph, indexing="ij")
    w = np.stack([np.sin(tt) * np.cos(pp), np.sin(tt) * np.sin(pp), np.cos(tt)], -1).reshape(-1, 3).astype(np.float32)
    dw = (np.sin(tt) * (np.pi / n_theta) * (2 * np.pi / n_phi)).ravel()
    n = len(w)
    ref = np.tile(np.float32([0, 0, 0, 0, 1, 0]), (n, 1))
    r = prober.light_probe(light, ref, np.zeros((n, 2), np.float32), w)
    y = lambda c: 0.212671 * c[:, 0] + 0.715160 * c[:, 1] + 0.072169 * c[:, 2]
    pdf_mass = float((r["pdf_w"].astype(np.float64) * dw).sum())
    le_int = float((y(r["le_w"].astype(np.float64)) * dw).sum())
    rng = np.random.default_rng(seed)
    u = rng.random((n_mc, 2)).astype(np.float32)
    refm = np.tile(np.float32([0, 0, 0, 0, 1, 0]), (n_mc, 1))
    s = prober.light_probe(light, refm, u, np.tile(np.float32([0, 1, 0]), (n_mc, 1)))
    ok = s["pdf"] > 0
    est = np.zeros(n_mc)
    est[ok] = y(s["li"][ok].astype(np.float64)) / s["pdf"][ok]
    return pdf_mass, le_int, float(est.mean()), float(est.std() / np.sqrt(n_mc))
