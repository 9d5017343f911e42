"""Parity cases for branches of the hot path that round 1 shipped without a device-side test ("missing #3"):
a traced AGeoWinstonCone2D, the Schott / Cauchy / mixed refractive-index formulas, QE(theta) on a focal surface, a TH2
reflectance on an AMirror.  Each case returns (export, input rays, opts, extra) and is run by two test modules: the host build
of the device code (tests/test_parity_branches.py, CPU) and the CUDA path through the C ABI (-m gpu)."""
import math

import numpy as np

import helpers as H
import scenes

nm, mm, cm, m = 1e-7, 0.1, 1.0, 100.0


def _isotropic(rng, n, origin, lam=400 * nm, spread=0.0):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    inp = np.zeros((n, 8))
    inp[:, 0:3] = np.asarray(origin)[None, :] + spread * (rng.random((n, 3)) - 0.5)
    inp[:, 4:7] = v
    inp[:, 7] = lam
    return inp


def _towards(rng, n, radius, target_spread, lam=400 * nm):
    """rays from a sphere of the given radius aimed at random points near the origin (all faces of a body get hit from outside)"""
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    src = radius * v
    dst = target_spread * (rng.random((n, 3)) - 0.5)
    d = dst - src
    d /= np.linalg.norm(d, axis=1)[:, None]
    inp = np.zeros((n, 8))
    inp[:, 0:3], inp[:, 4:7], inp[:, 7] = src, d, lam
    return inp


def winston2d_solid(material, n=6000, seed=11):
    mgr, keep = scenes.winston2d("solid", material)
    rng = np.random.default_rng(seed)
    inp = np.vstack([_towards(rng, n // 2, 12., 5.0), _isotropic(rng, n - n // 2, (0.3, -0.2, 0.5), spread=1.0)])
    # limit 20: a ray trapped by total internal reflection amplifies rounding differences bounce after bounce (checked: <= 4e-11 cm
    # up to 40 points, 1e-5 cm after 100)
    return mgr.ExportScene(), inp, H.opts(seed=seed, disable_fresnel=0, limit=20), keep


def winston2d_hex3(theta_deg, n=6000, seed=12):
    from robast_b200 import configs
    mgr, keep = scenes.winston2d("hex3")
    beam = configs.beam(5, theta_deg, n_side=5.0)
    return mgr.ExportScene(), (beam, n), H.opts(seed=seed), keep


def index_formulas(R):
    """the four analytic index models and one mixture, as (name, index object, closed form n(lambda_um))"""
    schott_c = (2.2718929, -1.0108077e-2, 1.0592509e-2, 2.0816965e-4, -7.6472538e-6, 4.9240991e-7)  # BK7, Schott 1992 catalogue
    cauchy_c = (1.4580, 0.00354, 1.2e-5)  # fused silica, micron units
    sell_c = (1.03961212, 0.231792344, 1.01046945, 0.00600069867, 0.0200179144, 103.560653)  # N-BK7 (unittest_robast.py:532-537)
    schott = R.ASchottFormula(*schott_c)
    cauchy = R.ACauchyFormula(*cauchy_c)
    sell = R.ASellmeierFormula(*sell_c)
    mixed = R.AMixedRefractiveIndex(schott, cauchy, 3., 1.)  # fractions are normalised: 0.75 / 0.25

    def n_schott(l):
        return math.sqrt(schott_c[0] + schott_c[1] * l ** 2 + schott_c[2] * l ** -2 + schott_c[3] * l ** -4 + schott_c[4] * l ** -6 + schott_c[5] * l ** -8)

    def n_cauchy(l):
        return cauchy_c[0] + cauchy_c[1] * l ** -2 + cauchy_c[2] * l ** -4

    return [("schott", schott, n_schott), ("cauchy", cauchy, n_cauchy), ("mixed", mixed, lambda l: 0.75 * n_schott(l) + 0.25 * n_cauchy(l)),
            ("mixed_sell", R.AMixedRefractiveIndex(sell, cauchy, 0.4, 0.6), None)], [schott, cauchy, sell, mixed]


def dispersive_lens_case(index, n=4000, seed=13, disable_fresnel=0):
    """polychromatic rays (300-700 nm) through a glass cube at assorted incidence angles: refraction in, refraction or total
    internal reflection out, Fresnel side branches decided by the same Philox streams on both sides"""
    mgr, lens = scenes.dispersive_lens(index)
    rng = np.random.default_rng(seed)
    inp = _towards(rng, n, 150., 60.0)
    inp[:, 7] = (300 + 400 * rng.random(n)) * nm
    return mgr.ExportScene(), inp, H.opts(seed=seed, disable_fresnel=disable_fresnel, limit=30), [lens, index]


def qe_case(R, with_lambda, with_angle, n_side=300):
    """unittest_robast.py:470-522: 45-degree beam on a focal box; QE(lambda = 400 nm) = 0.5, QE(45 deg) = 0.5"""
    ql = qa = None
    if with_lambda:
        ql = R.TGraph()
        ql.SetPoint(0, 300 * nm, 0.0)
        ql.SetPoint(1, 500 * nm, 1.0)
    if with_angle:
        qa = R.TGraph()
        qa.SetPoint(0, 0., 1.)
        qa.SetPoint(1, math.pi / 2, 0.)
    mgr, focal = scenes.focal_box_with_qe(qe_lambda=ql, qe_angle=qa)
    beam = dict(kind=0, nx=n_side, ny=n_side, dx=1 * mm, dy=1 * mm, lambda_min=400 * nm, lambda_max=400 * nm, rot=[1, 0, 0, 0, 1, 0, 0, 0, 1],
                tr=[0, 0, 2 * mm], dir=[math.cos(math.pi / 4), 0, -math.sin(math.pi / 4)], seed=1)
    return mgr.ExportScene(), (beam, n_side * n_side), H.opts(seed=470), [focal, ql, qa]


def th2_mirror_case(n=20000, seed=14):
    mgr, mirror, keep = scenes.th2_mirror()
    rng = np.random.default_rng(seed)
    inp = np.zeros((n, 8))
    inp[:, 2] = 51.
    ang = rng.random(n) * 1.5
    inp[:, 4], inp[:, 6] = np.sin(ang), -np.cos(ang)
    inp[:, 7] = (290 + 220 * rng.random(n)) * nm  # also below / above the histogram's wavelength range (-> 0 there)
    return mgr.ExportScene(), inp, H.opts(seed=seed), keep + [mirror, ang]


def run(oracle, backend, export, inp, o, nthreads=4):
    """trace the same rays with the oracle and with `backend` ('gpu' or the emul library); returns (ref, got, report)"""

    def rays():
        if isinstance(inp, tuple):
            return H.make_rays(oracle, inp[0], 0, inp[1])
        return H.Rays(inp)

    ref = H.trace_with(oracle.orc_trace, export, rays(), o, nthreads=nthreads)
    got = H.trace_gpu(export, rays(), o) if backend == "gpu" else H.trace_with(backend.emul_trace, export, rays(), o)
    return ref, got, H.compare(ref, got)


def clean(rep):
    return rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0
