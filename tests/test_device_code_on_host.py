"""CPU-side check of the CUDA path's per-ray math: robast_b200/csrc/rb_device.cuh compiled for the
host (tests/emul) and compared with the oracle per ray on all five configs.  This is a development
aid for the GPU-less box; the real parity tests (-m gpu) run the kernels through the C ABI."""
import ctypes as C
import math

import numpy as np
import pytest

import helpers as H
import scenes
from robast_b200 import configs

CASES = [(1, 0.0, 101, {}), (1, 1.5, 61, {}), (2, 0.0, 81, {}), (2, 2.5, 81, {}), (3, 0.0, 101, {}), (3, 5.0, 81, {}),
         (4, 0.0, 60, {}), (4, 0.1, 60, {}), (5, 0.0, 50, {}), (5, 25.0, 50, {}), (5, 38.0, 40, {})]


@pytest.mark.parametrize("cfg,theta,nside,kw", CASES)
def test_configs_match_oracle(oracle, emul, cfg, theta, nside, kw):
    mgr, _keep = configs.BUILDERS[cfg](**kw)
    ex = mgr.ExportScene()
    beam = configs.beam(cfg, theta, n_side=nside if cfg <= 3 else None)
    n = nside * nside
    o = H.opts(disable_fresnel=1 if cfg == 2 else 0, seed=99)
    ref = H.trace_with(oracle.orc_trace, ex, H.make_rays(oracle, beam, 0, n), o, nthreads=4)
    got = H.trace_with(emul.emul_trace, ex, H.make_rays(oracle, beam, 0, n), o)
    rep = H.compare(ref, got)
    assert rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0, rep


def test_tmm_device_code_matches_oracle(R, oracle, emul):
    import os
    air = R.ARefractiveIndex(1., 0.)
    sio2 = R.AFilmetrixDotCom(os.path.join(configs.DATA, "SiO2.nk.txt"))
    al = R.AFilmetrixDotCom(os.path.join(configs.DATA, "Al.nk.txt"))
    ml = R.AMultilayer(air, al)
    ml.InsertLayer(sio2, 25.4e-7)
    ex, mid = R.export_multilayer(ml)
    worst = 0
    for lam in np.linspace(250e-7, 950e-7, 15):
        for th in np.linspace(0, 1.55, 12):
            for pol in (0, 1, 2):
                a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_double()
                oracle.orc_tmm(ex.desc_ptr(), mid, pol, th, lam, C.byref(a), C.byref(b))
                emul.emul_tmm(ex.desc_ptr(), mid, pol, th, lam, C.byref(c), C.byref(d))
                worst = max(worst, abs(a.value - c.value), abs(b.value - d.value))
    assert worst < 1e-12


def test_unit_scenes_match_oracle(R, oracle, emul):
    mgr, _k = scenes.snell_slab(1.5)
    th = math.radians(30)
    for fn in (oracle.orc_trace, emul.emul_trace):
        rays = H.Rays([[0, 0, 0.2, 0, math.sin(th), 0, -math.cos(th), 400e-7]])
        H.trace_with(fn, mgr.ExportScene(), rays, H.opts(disable_fresnel=1))
        assert abs(rays.dirs[0, 0] - math.sin(th) / 1.5) < 1e-12 and rays.status[0] == 3
    mgr, _k = scenes.sphere_shell_mirror()
    # 999 bounces between a convex and a concave sphere amplify rounding differences exponentially
    # (dispersing billiard): the long run pins the count, a short run pins the per-ray agreement
    rays = H.Rays([[0, 0, 0, 0, 0.3, 0.1, -1, 400e-7]])
    H.trace_with(emul.emul_trace, mgr.ExportScene(), rays, H.opts(limit=1000))
    assert rays.npoints[0] == 1000 and rays.status[0] == 4
    got = H.trace_with(emul.emul_trace, mgr.ExportScene(), H.Rays([[0, 0, 0, 0, 0.3, 0.1, -1, 400e-7]]), H.opts(limit=12))
    ref = H.trace_with(oracle.orc_trace, mgr.ExportScene(), H.Rays([[0, 0, 0, 0, 0.3, 0.1, -1, 400e-7]]), H.opts(limit=12))
    assert got.npoints[0] == 12 and H.compare(ref, got)["bad"] == 0


def test_tgraph2d_mirror_device_code_matches_oracle(R, oracle, emul):
    g = R.TGraph2D()
    for i, (lam, th, v) in enumerate(((300e-7, 0., 0.0), (300e-7, math.pi / 2, 0.3), (500e-7, 0., 0.7), (500e-7, math.pi / 2, 1.0), (420e-7, 0.7, 0.9))):
        g.SetPoint(i, lam, th, v)
    mgr, mirror, keep = scenes.mirror_box_with_border(reflectance=g)
    rng = np.random.default_rng(5)
    n = 4000
    inp = np.zeros((n, 8))
    inp[:, 2] = 51.
    ang = rng.random(n) * 1.4
    inp[:, 4], inp[:, 6], inp[:, 7] = np.sin(ang), -np.cos(ang), 300e-7 + 200e-7 * rng.random(n)
    ex = mgr.ExportScene()
    ref = H.trace_with(oracle.orc_trace, ex, H.Rays(inp), H.opts(seed=3))
    got = H.trace_with(emul.emul_trace, ex, H.Rays(inp), H.opts(seed=3))
    rep = H.compare(ref, got)
    assert rep["bad"] == 0 and rep["status_mismatch"] == 0, rep
    frac = (got.status == 2).mean()
    assert 0.3 < frac < 0.9


@pytest.mark.parametrize("kind,theta", [("pgon", 0.0), ("pgon", 22.0), ("pcon", 0.0), ("pcon", 18.0), ("pcon", 35.0)])
def test_bezier_cones_match_oracle(oracle, emul, kind, theta):
    """HexOkumuraCone.C mode 1 (AGeoBezierPgon, 100 sections) and its round AGeoBezierPcon counterpart: the device code's
    convex-slab polygon/cone intersection against the oracle's candidate/Contains restatement of TGeoPgon/TGeoPcon"""
    mgr, _keep = configs.okumura_cone(kind)
    ex = mgr.ExportScene()
    beam = configs.beam(5, theta, n_side=6.0)
    n = 3000
    o = H.opts(seed=7)
    ref = H.trace_with(oracle.orc_trace, ex, H.make_rays(oracle, beam, 0, n), o, nthreads=4)
    got = H.trace_with(emul.emul_trace, ex, H.make_rays(oracle, beam, 0, n), o)
    rep = H.compare(ref, got)
    assert rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0, rep
    frac = (got.status == 3).mean()
    assert (frac > 0.05) if theta < 25 else (frac < 0.05)  # inside / outside the ~30 deg acceptance of the guide


@pytest.mark.parametrize("cfg,theta,n,kw", [(4, 0.05, 2500, {}), (5, 18.0, 2500, {"rings": 1}), (2, 1.0, 2500, {})])
def test_polyline_history_matches_oracle(oracle, emul, cfg, theta, n, kw):
    """ARay's full polyline + node history (include/ARay.h:24-68): every AddPoint/AddNode of the device code against the oracle's"""
    mgr, _keep = configs.BUILDERS[cfg](**kw)
    ex = mgr.ExportScene()
    beam = configs.beam(cfg, theta, n_side=50 if cfg <= 3 else (12.0 if cfg == 5 else None))
    o = H.opts(disable_fresnel=1 if cfg == 2 else 0, seed=17)
    ra, rb = H.make_rays(oracle, beam, 0, n), H.make_rays(oracle, beam, 0, n)
    ha = H.trace_history_with(oracle.orc_trace_history, ex, ra, o, 12, nthreads=4)
    hb = H.trace_history_with(emul.emul_trace_history, ex, rb, o, 12)
    assert H.compare(ra, rb)["bad"] == 0 and (ra.npoints == rb.npoints).all()
    assert H.compare_history(ha, hb, ra.npoints) == 0
    # the record is consistent with the plain outputs: point 0 = start, last recorded point = last point, node = last node
    k = np.minimum(rb.npoints, 12) - 1
    idx = np.arange(n)
    full = rb.npoints <= 12
    assert np.allclose(hb.pts[:3, 0, :], rb.inp[:3]) and (hb.node[0] == -1).all()
    assert np.allclose(hb.pts[:3, k, idx][:, full], rb.out[:3][:, full], atol=0) and (hb.node[k, idx][full & (rb.npoints > 1)] == rb.last_node[full & (rb.npoints > 1)]).all()
    assert rb.npoints.max() >= 3


def point_source_rays(oracle, kind, origin, n, seed, theta=0.):
    params = dict(kind=kind, nx=1, ny=1, dx=theta, dy=0., lambda_min=400e-7, lambda_max=400e-7, rot=[1, 0, 0, 0, 1, 0, 0, 0, 1], tr=list(origin), dir=[0, 0, 1], seed=seed)
    return H.make_rays(oracle, params, 0, n)


@pytest.mark.parametrize("kind,phi1,dphi", [("pcon", 0., 360.), ("pcon", 30., 250.), ("pgon", 0., 360.), ("pgon", -40., 200.)])
def test_general_pcon_pgon_match_oracle(oracle, emul, kind, phi1, dphi):
    """hollow / azimuthally segmented TGeoPcon and TGeoPgon (device: ordered-candidate walk; oracle: sorted candidates +
    Contains): isotropic point sources outside the solid, in its bore and inside its material; mirror walls and glass"""
    bounces = 0
    for material, sources in (("mirror", (((40., 10., -5.), 1), ((1., -2., 3.), 2))), ("glass", (((40., 10., -5.), 3), ((1.5, 5.5, 4.), 4)))):
        mgr, _keep = scenes.hollow_poly(kind, phi1, dphi, material)
        ex = mgr.ExportScene()
        for origin, seed in sources:
            n = 3000
            o = H.opts(seed=5, limit=30)
            ref = H.trace_with(oracle.orc_trace, ex, point_source_rays(oracle, 5, origin, n, seed), o, nthreads=4)
            got = H.trace_with(emul.emul_trace, ex, point_source_rays(oracle, 5, origin, n, seed), o)
            rep = H.compare(ref, got)
            assert rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0, (material, origin, rep)
            bounces = max(bounces, int(got.npoints.max()))
    assert bounces > 4  # walls are hit repeatedly (bore reflections, total internal reflection in the glass)


def test_general_tmm_device_code_matches_oracle_and_golden(R, oracle, emul):
    """tmm_coherent_sub (complex angle, reversed) and tmm_incoherent of rb_device.cuh against the oracle and tmm.py's values"""
    import test_oracle_golden as G
    emul.emul_tmm_general.restype = C.c_int
    emul.emul_tmm_general.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    multi, keep, th_0 = G.incoherent_stack(R)
    ex, mid = R.export_multilayer(multi)
    want = {0: (0.3776110935131179, 1.2856977234844612e-05), 1: (0.03199545463016445, 2.0900281396463212e-05)}
    for pol in (0, 1):
        r, t = C.c_double(), C.c_double()
        assert emul.emul_tmm_general(ex.desc_ptr(), mid, 1, pol, 0, th_0.real, th_0.imag, 400., C.byref(r), C.byref(t)) == 0
        assert abs(r.value - want[pol][0]) < 1e-12 and abs(t.value / want[pol][1] - 1) < 1e-10
    rng = np.random.default_rng(4)
    for k in range(200):
        th = complex(rng.random() * 1.5, 0.0)
        lam = 300. + 500. * rng.random()
        for mode, rev in ((0, 0), (0, 1), (1, 0)):
            for pol in (0, 1):
                a, b, c, d_ = C.c_double(), C.c_double(), C.c_double(), C.c_double()
                # an absorbing entrance medium needs the matching complex angle: take it from Snell's law out of vacuum
                thc = np.lib.scimath.arcsin(math.sin(th.real) / (1 + 0.1j)) if rev == 0 else np.lib.scimath.arcsin(math.sin(th.real) / (4 + 0.2j))
                thc = complex(thc)
                assert emul.emul_tmm_general(ex.desc_ptr(), mid, mode, pol, rev, thc.real, thc.imag, lam, C.byref(a), C.byref(b)) == 0
                assert oracle.orc_tmm_general(ex.desc_ptr(), mid, mode, pol, rev, thc.real, thc.imag, lam, C.byref(c), C.byref(d_)) == 0
                assert abs(a.value - c.value) < 1e-11 * max(1, abs(c.value)) and abs(b.value - d_.value) < 1e-11 * max(1e-3, abs(d_.value)), (k, mode, rev, pol)


ARB8_XTRU_KINDS = ["arb8_prism", "arb8_twisted", "arb8_pyramid", "arb8_ccw", "xtru_profile", "xtru_scaled"]


@pytest.mark.parametrize("kind", ARB8_XTRU_KINDS)
@pytest.mark.parametrize("composite", [False, True])
def test_arb8_xtru_match_oracle(oracle, emul, kind, composite):
    """TGeoArb8 (prism, twisted face, pyramid, counter-clockwise input) and TGeoXtru (concave outline, scaled sections with an
    outline jump) — tutorials/AshraOptics.C:264-284,403-441,791-1021 — alone and cut by a sphere: device code (candidate roots
    + Contains at midpoints) against the oracle's face-by-face restatement; isotropic point sources outside and inside"""
    bounces = 0
    inside = (1.5, -1.5, 3.2) if kind != "arb8_twisted" else (14.2, -9.0, 8.6)
    for material, sources in (("mirror", (((26., 8., -3.), 1), ((-17., -12., 18.), 2))), ("glass", (((26., 8., -3.), 3), (inside, 4)))):
        mgr, _keep = scenes.arb8_xtru(kind, material, composite)
        ex = mgr.ExportScene()
        for origin, seed in sources:
            n = 3000
            o = H.opts(seed=5, limit=12)
            ref = H.trace_with(oracle.orc_trace, ex, point_source_rays(oracle, 5, origin, n, seed), o, nthreads=4)
            got = H.trace_with(emul.emul_trace, ex, point_source_rays(oracle, 5, origin, n, seed), o)
            rep = H.compare(ref, got)
            assert rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0, (material, origin, rep)
            bounces = max(bounces, int(got.npoints.max()))
            assert (got.npoints > 2).mean() > 0.005  # the solid is hit
    assert bounces > 3


def overlap_beam(n_side, tilt_deg):
    """parallel beam onto the scenes.overlapping_frame system, starting inside the overlapping holder"""
    th = math.radians(tilt_deg)
    xs = np.linspace(-60., 60., n_side)
    X, Y = np.meshgrid(xs, xs)
    inp = np.zeros((n_side * n_side, 8))
    inp[:, 0], inp[:, 1], inp[:, 2] = X.ravel(), Y.ravel(), 150.
    inp[:, 4], inp[:, 5], inp[:, 6], inp[:, 7] = math.sin(th), 0., -math.cos(th), 400e-7
    return inp


@pytest.mark.parametrize("nested", [False, True])
@pytest.mark.parametrize("tilt", [0.0, 3.0])
def test_overlapping_nodes_match_oracle(oracle, emul, nested, tilt):
    """AddNodeOverlap ("MANY") nodes (tutorials/AshraOptics.C:91,1117-1120): rays start inside an overlapping holder laid over the
    whole system and still meet the ordinary sisters (glass plate, mirror, focal plane) and the holder's own bars — flattened
    device navigation against the oracle's path-stack navigator"""
    mgr, _keep = scenes.overlapping_frame(nested)
    ex = mgr.ExportScene()
    o = H.opts(seed=3, limit=20, disable_fresnel=1)
    ref = H.trace_with(oracle.orc_trace, ex, H.Rays(overlap_beam(60, tilt)), o, nthreads=4)
    got = H.trace_with(emul.emul_trace, ex, H.Rays(overlap_beam(60, tilt)), o)
    rep = H.compare(ref, got)
    assert rep["bad"] == 0 and rep["status_mismatch"] == 0 and rep["npoints_mismatch"] == 0 and rep["node_mismatch"] == 0, rep
    st = np.bincount(got.status, minlength=6)
    assert st[3] > 200 and st[1] > 200, st  # focused via plate + mirror; stopped on the holder's bars or the overlapping stop ring
    assert got.npoints[got.status == 3].max() >= 6  # start, plate in/out, mirror, plate in/out, focal plane


def test_zero_numerator_division_shortcut_is_ieee_exact(emul):
    """rb_div (rb_device.cuh) keeps zero numerators away from the compiler's out-of-line fp64 division; its result has to be the
    IEEE quotient bit for bit, signs of zero, infinities and NaNs included"""
    import struct
    inf, nan = float("inf"), float("nan")
    vals = [0.0, -0.0, 1.0, -1.0, 3.5, -2.25e-300, 1e300, 5e-324, -5e-324, inf, -inf, nan]
    for a in vals:
        for b in vals:
            got = emul.emul_div(a, b)
            with np.errstate(all="ignore"):
                want = float(np.float64(a) / np.float64(b))
            if want != want:
                assert got != got, (a, b, got)
            else:
                assert struct.pack("<d", got) == struct.pack("<d", want), (a, b, got, want)
