"""Pins the CPU oracle against the reference's own golden vectors (tutorials/unittest_robast.py)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import helpers as H
import scenes
from robast_b200 import configs

nm, um, mm, m = 1e-7, 1e-4, 0.1, 100.0
deg = math.pi / 180


def tmm(oracle, ml, pol, th, lam, export_fn=None):
    import robast_b200 as R
    ex, mid = R.export_multilayer(ml)
    r, t = C.c_double(), C.c_double()
    assert oracle.orc_tmm(ex.desc_ptr(), mid, pol, th, lam, C.byref(r), C.byref(t)) == 0
    return r.value, t.value


def basic_stack(R, reverse=False):  # unittest_robast.py:629-636
    med1, med2, med3, med4 = R.ARefractiveIndex(1.), R.ARefractiveIndex(2., 4.), R.ARefractiveIndex(3., .3), R.ARefractiveIndex(1., .1)
    if not reverse:
        multi = R.AMultilayer(med1, med4)
        multi.InsertLayer(med2, 2)
        multi.InsertLayer(med3, 3)
    else:
        multi = R.AMultilayer(med4, med1)
        multi.InsertLayer(med3, 3)
        multi.InsertLayer(med2, 2)
    return multi, [med1, med2, med3, med4]


def test_kat_tmm_basic(R, oracle):  # unittest_robast.py:627-655 (tmm.tests.basic_test)
    multi, _k = basic_stack(R)
    rs, rp = 0.37273208839139516, 0.37016110373044969
    ts, tp = 0.22604491247079261, 0.22824374314132009
    r, t = tmm(oracle, multi, 0, 0.1, 100)
    assert abs(r - rs) < 1e-12 and abs(t - ts) < 1e-12
    r, t = tmm(oracle, multi, 1, 0.1, 100)
    assert abs(r - rp) < 1e-12 and abs(t - tp) < 1e-12
    r, t = tmm(oracle, multi, 2, 0.1, 100)
    assert abs(r - (rs + rp) / 2) < 1e-12 and abs(t - (ts + tp) / 2) < 1e-12


def test_kat_tmm_bare_interface_is_fresnel(R, oracle):
    # two semi-infinite media: TMM must reduce to the Fresnel equations
    a, b = R.ARefractiveIndex(1.), R.ARefractiveIndex(1.5)
    ml = R.AMultilayer(a, b)
    th = 30 * deg
    th2 = math.asin(math.sin(th) / 1.5)
    rs = ((math.cos(th) - 1.5 * math.cos(th2)) / (math.cos(th) + 1.5 * math.cos(th2))) ** 2
    rp = ((1.5 * math.cos(th) - math.cos(th2)) / (1.5 * math.cos(th) + math.cos(th2))) ** 2
    r, t = tmm(oracle, ml, 0, th, 500 * nm)
    assert abs(r - rs) < 1e-14 and abs(r + t - 1) < 1e-14
    r, t = tmm(oracle, ml, 1, th, 500 * nm)
    assert abs(r - rp) < 1e-14 and abs(r + t - 1) < 1e-14


def test_kat_tmm_table_matches_direct_at_bin_centre(R, oracle):  # unittest_robast.py:698-709, table built by the oracle
    multi, _k = basic_stack(R)
    r0, t0 = tmm(oracle, multi, 2, 45 * deg, 600)
    h_r = R.TH2D("", "", 801, 199.5, 1000.5, 90, -0.5 * deg, 89.5 * deg)
    for j in range(1, 91):
        for i in range(1, 802):
            lam, th = 199.5 + (i - 0.5), (-0.5 + (j - 0.5)) * deg
            if abs(lam - 600) < 2 and abs(th - 45 * deg) < 2 * deg:
                h_r.SetBinContent(i, j, tmm(oracle, multi, 2, th, lam)[0])
    ex, hid = R.export_th2(h_r)
    assert abs(oracle.orc_th2_interp(ex.desc_ptr(), hid, 600., 45 * deg) - r0) < 1e-7
    assert abs(h_r.Interpolate(600., 45 * deg) - r0) < 1e-7  # host TH2 mirror agrees with the oracle's restatement


def test_kat_sellmeier_nbk7(R, oracle):  # unittest_robast.py:530-561
    nbk7 = R.ASellmeierFormula(1.03961212, 0.231792344, 1.01046945, 0.00600069867, 0.0200179144, 103.560653)
    ex, iid = R.export_index(nbk7)
    data = ((2325.4, 1.489210), (1970.1, 1.494950), (1529.6, 1.500910), (1060.0, 1.506690), (1014.0, 1.507310), (852.1, 1.509800),
            (706.5, 1.512890), (656.3, 1.514320), (643.8, 1.514720), (632.8, 1.515090), (589.3, 1.516730), (587.6, 1.516800),
            (546.1, 1.518720), (486.1, 1.522380), (480.0, 1.522830), (435.8, 1.526680), (404.7, 1.530240), (365.0, 1.536270),
            (334.1, 1.542720), (312.6, 1.548620))
    for wl, n in data:
        assert abs(oracle.orc_index_n(ex.desc_ptr(), iid, wl * nm) - n) < 5e-5  # 4 decimal places, as the reference asserts
        assert abs(nbk7.GetRefractiveIndex(wl * nm) - n) < 5e-5
    assert abs(nbk7.GetAbbeNumber() - 64.17) < 0.05


def test_kat_agf_nbk7(R, oracle):  # unittest_robast.py:524-528
    schott = R.AGlassCatalog(os.path.join(configs.DATA, "nbk7.agf"))
    r = schott.GetRefractiveIndex("N-BK7")
    ex, iid = R.export_index(r)
    assert abs(oracle.orc_index_n(ex.desc_ptr(), iid, 589.3 * nm) / 1.51680 - 1.) < 5e-5
    assert abs(oracle.orc_index_k(ex.desc_ptr(), iid, 2325 * nm) - 4.2911e-6) < 5e-10
    assert abs(r.GetExtinctionCoefficient(2325 * nm) - 4.2911e-6) < 5e-10


def test_kat_tgraph_eval(R, oracle):  # unittest_robast.py:415-426, src/ARefractiveIndex.cxx:19-27
    g = R.TGraph()
    g.SetPoint(0, 400 * nm, 1.6)
    g.SetPoint(1, 500 * nm, 1.5)
    ex, gid = R.export_graph(g)
    ev = lambda x: oracle.orc_graph_eval(ex.desc_ptr(), gid, x)
    assert abs(ev(450 * nm) - 1.55) < 1e-15 and abs(g.Eval(450 * nm) - 1.55) < 1e-15
    assert abs(ev(600 * nm) - 1.4) < 1e-12 and abs(ev(300 * nm) - 1.7) < 1e-12  # linear extrapolation
    one = R.ARefractiveIndex(1.25, 0.5)
    ex, iid = R.export_index(one)
    assert oracle.orc_index_n(ex.desc_ptr(), iid, 123 * nm) == 1.25 and oracle.orc_index_k(ex.desc_ptr(), iid, 1.) == 0.5


def test_kat_mixed_index(R, oracle):  # unittest_robast.py:612-625
    a, b = R.ARefractiveIndex(1., 1.), R.ARefractiveIndex(2., 2.)
    mixed = R.AMixedRefractiveIndex(a, b, 3, 7)
    ex, iid = R.export_index(mixed)
    assert abs(oracle.orc_index_n(ex.desc_ptr(), iid, 100 * nm) - 1.7) < 1e-15
    assert abs(oracle.orc_index_k(ex.desc_ptr(), iid, 100 * nm) - 1.7) < 1e-15


def test_kat_snell_slab(oracle):  # unittest_robast.py:428-468
    mgr, _k = scenes.snell_slab(1.5)
    th = 30 * deg
    rays = H.Rays([[0, 0, 2 * mm, 0, math.sin(th), 0, -math.cos(th), 400 * nm]])
    H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(disable_fresnel=1))
    assert abs(rays.dirs[0, 0] - math.sin(th) / 1.5) < 1e-12 and abs(rays.dirs[0, 1]) < 1e-15
    assert rays.status[0] == 3  # focused on the box nested inside the lens


def test_kat_limit_suspend(oracle):  # unittest_robast.py:390-413
    mgr, _k = scenes.sphere_shell_mirror()
    rays = H.Rays([[0, 0, 0, 0, 0, 0, -1, 400 * nm]])
    H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(limit=1000))
    assert rays.npoints[0] == 1000 and rays.status[0] == 4


def test_stat_fresnel_n3(R, oracle):  # unittest_robast.py:122-160
    wl, absl, idx = 400 * nm, 1 * um, 3.
    k = R.ARefractiveIndex.AbsorptionLengthToExtinctionCoefficient(absl, wl)
    refidx = R.ARefractiveIndex(idx, k)
    mgr, lens = scenes.lens_box(refidx)
    N = 100000
    rays = H.Rays(np.tile([0, 0, 0.8 * m, 0, 0, 0, -1, wl], (N, 1)))
    H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(seed=7), nthreads=4)
    n = int((rays.status == 2).sum())
    ref = (idx - 1) ** 2 / (idx + 1) ** 2
    # k > 0 -> absorbing-medium Fresnel form; the reference's own 3 sigma window
    assert (n - 3 * n ** 0.5) / N < ref * 1.002 and ref * 0.998 < (n + 3 * n ** 0.5) / N
    assert int((rays.status == 5).sum()) + n == N


def test_stat_absorption_length(R, oracle):  # unittest_robast.py:67-120
    wl, absl = 400 * nm, 1 * mm
    g_n, g_k = R.TGraph(), R.TGraph()
    g_n.SetPoint(0, wl, 1)
    g_k.SetPoint(0, wl, R.ARefractiveIndex.AbsorptionLengthToExtinctionCoefficient(absl, wl))
    refidx = R.ARefractiveIndex()
    refidx.SetRefractiveIndex(g_n)
    refidx.SetExtinctionCoefficient(g_k)
    mgr, lens = scenes.lens_box(refidx)
    N = 20000
    rng = np.random.default_rng(1)
    d = rng.normal(size=(N, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    inp = np.zeros((N, 8))
    inp[:, 4:7] = d
    inp[:, 7] = wl
    rays = H.Rays(inp)
    H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(seed=3), nthreads=4)
    ab = rays.status == 5
    assert ab.sum() == N  # 1 mm absorption length in a 1 m cube: nothing escapes
    dist = np.linalg.norm(rays.pos[ab], axis=1)
    assert abs(dist.mean() / absl - 1) < 3 / math.sqrt(N)


def test_kat_parabola_focus_and_stepback_quirk(oracle):
    mgr, _k = configs.simple_parabolic()
    ex = mgr.ExportScene()
    beam = configs.beam(1, 0.0, n_side=101)
    ideal = H.make_rays(oracle, beam, 0, 101 * 101)
    H.trace_with(oracle.orc_trace, ex, ideal, H.opts(quirks=2))
    f = ideal.status == 3
    assert f.sum() > 2000
    assert np.abs(ideal.pos[f][:, :2]).max() < 1e-11 and np.abs(ideal.pos[f][:, 2] - 300.).max() < 1e-11
    quirk = H.make_rays(oracle, beam, 0, 101 * 101)
    H.trace_with(oracle.orc_trace, ex, quirk, H.opts(quirks=3))
    assert (quirk.status == ideal.status).all()
    # reflection vertex 2e-6 cm short along d1=(0,0,-1): the hit moves by 2e-6 * tan(2 alpha), tan(alpha) = r / 2F
    r = np.hypot(quirk.inp[0][f], quirk.inp[1][f])
    alpha = np.arctan(r / 600.)
    shift = np.hypot(quirk.pos[f][:, 0], quirk.pos[f][:, 1])
    assert np.abs(shift - 2e-6 * np.tan(2 * alpha)).max() < 1e-11
    assert 1.0e-6 < shift.max() < 1.1e-6
    # classes: r < 20.001 cm stops on obs1, 20.001 < r < 150 cm focuses, else exits at z = -10 m
    r_all = np.hypot(quirk.inp[0], quirk.inp[1])
    assert (quirk.status[r_all < 20.0] == 1).all() and (quirk.status[(r_all > 20.01) & (r_all < 149.99)] == 3).all()
    ex_ = quirk.status == 2
    assert (r_all[ex_] > 149.99).all() and np.abs(quirk.pos[ex_][:, 2] + 1000.).max() < 1e-9
    # time of flight: path length / c
    tof = quirk.time[f] * 2.99792458e10
    assert np.abs(tof - (600. - r[...] ** 2 / 1200. + np.hypot(r, 300. - r ** 2 / 1200.))).max() < 1e-5


def test_kat_davies_cotton_facets_aim_at_2f(R):  # SURVEY.md Appendix B sanity check of the Euler convention
    mgr, _k = configs.davies_cotton()
    ex = mgr.ExportScene()
    n = 0
    for i in range(ex.num_nodes()):
        vol, mat, copy, ovl = ex.node(i)
        if mat < 0 or i >= 88:
            continue
        rot, tr = ex.matrix(mat)
        normal = np.array([rot[2], rot[5], rot[8]])  # facet local +z in the world
        to_2f = np.array([0, 0, 3200.]) - np.array(tr)
        to_2f /= np.linalg.norm(to_2f)
        assert np.abs(normal - to_2f).max() < 1e-12
        # no net spin: the local y axis stays in the plane spanned by z and the radial direction's normal
        n += 1
    assert n == 88


def test_kat_shooter_grids(R, oracle):  # src/ARayShooter.cxx:122-183,401-452
    arr = R.ARayShooter.Square(400 * nm, 90., 4)
    c = arr.columns()
    assert arr.GetN() == 16
    assert np.allclose(c["x"][:5], [-45, -45, -45, -45, -15]) and np.allclose(c["y"][:5], [-45, -15, 15, 45, -45])  # x-major order
    assert np.allclose(c["dz"], 1)
    circ = R.ARayShooter.Circle(400 * nm, 100., 3, 6)
    assert circ.GetN() == 1 + 6 + 12 + 18
    # the oracle's Philox shooter reproduces the same grid
    b = dict(kind=0, nx=4, ny=4, dx=90., dy=90., lambda_min=400 * nm, lambda_max=400 * nm, rot=[1, 0, 0, 0, 1, 0, 0, 0, 1], tr=[0, 0, 0], dir=[0, 0, 1], seed=1)
    rays = H.make_rays(oracle, b, 0, 16)
    assert np.allclose(rays.inp[0], c["x"]) and np.allclose(rays.inp[1], c["y"])
    b.update(kind=3, nx=3, ny=6, dx=100.)
    rays = H.make_rays(oracle, b, 0, 37)
    cc = circ.columns()
    assert np.allclose(rays.inp[0], cc["x"], atol=1e-12) and np.allclose(rays.inp[1], cc["y"], atol=1e-12)


def test_point_source_shooters(R, oracle):  # src/ARayShooter.cxx:240-392
    base = dict(nx=1, ny=1, lambda_min=400 * nm, lambda_max=400 * nm, rot=[1, 0, 0, 0, 1, 0, 0, 0, 1], tr=[1., 2., 3.], dir=[0, 0, 1], seed=5)
    n = 40000
    cone = H.make_rays(oracle, dict(base, kind=4, dx=45., dy=10.), 0, n).inp   # RandomCone(r = 45, d = 10)
    sph = H.make_rays(oracle, dict(base, kind=5, dx=0., dy=0.), 0, n).inp      # RandomSphere
    scone = H.make_rays(oracle, dict(base, kind=6, dx=25., dy=0.), 0, n).inp   # RandomSphericalCone(theta = 25 deg)
    for a in (cone, sph, scone):
        assert np.allclose(a[0], 1.) and np.allclose(a[1], 2.) and np.allclose(a[2], 3.) and np.allclose(a[3], 0.)
        assert np.allclose(a[4] ** 2 + a[5] ** 2 + a[6] ** 2, 1., atol=1e-14)
    # RandomCone: the aim points d/dz * (dx, dy) fill the disc of radius r uniformly
    gx, gy = 10. * cone[4] / cone[6], 10. * cone[5] / cone[6]
    rr = np.hypot(gx, gy)
    assert rr.max() <= 45. * (1 + 1e-12) and abs((rr < 45. / math.sqrt(2)).mean() - 0.5) < 3 * 0.5 / math.sqrt(n)
    # RandomSphere: isotropic -> each component uniform in [-1, 1]
    for k in (4, 5, 6):
        assert abs(sph[k].mean()) < 4 / math.sqrt(3 * n) and abs((sph[k] ** 2).mean() - 1. / 3.) < 0.01
    # RandomSphericalCone: cos(theta) uniform in [cos 25 deg, 1], phi uniform
    c0 = math.cos(math.radians(25.))
    assert scone[6].min() >= c0 - 1e-12 and abs(scone[6].mean() - (1 + c0) / 2) < 4 * (1 - c0) / math.sqrt(12 * n)
    assert abs(np.arctan2(scone[5], scone[4]).mean()) < 4 * math.pi / math.sqrt(3 * n)
    # the host-side generators of the mirror classes have the same geometry (they draw from gRandom instead of Philox)
    arr = R.ARayShooter.RandomSphericalCone(400 * nm, 2000, 25.)
    c = arr.columns()
    assert c["dz"].min() >= c0 - 1e-12 and np.allclose(c["dx"] ** 2 + c["dy"] ** 2 + c["dz"] ** 2, 1.)
    arr = R.ARayShooter.RandomCone(400 * nm, 45., 10., 2000)
    c = arr.columns()
    assert (10. * np.hypot(c["dx"], c["dy"]) / c["dz"]).max() <= 45. * (1 + 1e-12)


def test_winston_cone_cutoff(oracle):  # closed-form property of a Winston cone: acceptance asin(R2/R1) = 30 deg
    mgr, _k = configs.hex_winston_cone(rings=0, coating="ideal")
    ex = mgr.ExportScene()
    frac = {}
    for th in (0., 15., 45.):
        rays = H.make_rays(oracle, configs.beam(5, th, n_side=30 * mm), 0, 4000)
        H.trace_with(oracle.orc_trace, ex, rays, H.opts(), nthreads=4)
        frac[th] = (rays.status == 3).mean()
    assert frac[0.] > 0.98 and frac[15.] > 0.9 and frac[45.] == 0.0


def test_schmidt_cassegrain_design_spot(oracle):  # the Zemax sample design focuses d-line light to a few-micron spot
    mgr, _k = configs.schmidt_cassegrain(disable_fresnel=True)
    b = configs.beam(4, 0.0)
    b["lambda_min"] = b["lambda_max"] = 587.6 * nm
    rays = H.make_rays(oracle, b, 0, 5000)
    H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(disable_fresnel=1), nthreads=4)
    f = rays.status == 3
    assert f.sum() > 3000 and rays.pos[f][:, :2].std(0).max() < 5e-4  # < 5 um rms


def test_kat_tgraph2d_interpolate(R, oracle):  # unittest_robast.py:220-229: 0.5 at (400 nm, 45 deg); src/AMirror.cxx:46-47
    deg = math.pi / 180.
    g = R.TGraph2D()
    g.SetPoint(0, 300 * nm, 0 * deg, 0.0)
    g.SetPoint(1, 300 * nm, 90 * deg, 0.3)
    g.SetPoint(2, 500 * nm, 0 * deg, 0.7)
    g.SetPoint(3, 500 * nm, 90 * deg, 1.0)
    assert len(g.GetTriangles()) == 6  # two triangles over the rectangle
    assert abs(g.Interpolate(400 * nm, 45 * deg) - 0.5) < 1e-12
    mirror = R.AMirror("mirror", R.TGeoBBox("mirrorbox", 50., 50., 50.))
    mirror.SetReflectance(g)
    assert abs(mirror.GetReflectance(400 * nm, 45 * deg) - 0.5) < 1e-3  # the reference asserts 3 places
    assert mirror.GetReflectance(800 * nm, 45 * deg) == 0.0  # outside the convex hull: TGraph2D returns 0
    mgr, mirror, keep = scenes.mirror_box_with_border(reflectance=g)
    ex = mgr.ExportScene()
    for lam, th in ((400 * nm, 45 * deg), (310 * nm, 80 * deg), (499 * nm, 1 * deg), (300 * nm, 90 * deg), (600 * nm, 0.2)):
        assert abs(oracle.orc_graph2d_interp(ex.desc_ptr(), 0, lam, th) - g.Interpolate(lam, th)) < 1e-13


def test_tgraph2d_delaunay_properties(R, oracle):
    """scattered points: the triangulation covers the convex hull exactly once, satisfies the empty-circumcircle
    property on the normalised coordinates ROOT triangulates in, and reproduces a plane exactly"""
    rng = np.random.default_rng(3)
    n = 60
    x, y = 3e-5 + 4e-5 * rng.random(n), 1.5 * rng.random(n)  # wavelengths (cm) and angles (rad): very different scales
    f = lambda a, b: 0.2 + 3000. * a + 0.25 * b
    g = R.TGraph2D()
    for i in range(n):
        g.SetPoint(i, x[i], y[i], f(x[i], y[i]))
    tri = np.asarray(g.GetTriangles()).reshape(-1, 3)
    xn, yn = (x - x.min()) / (x.max() - x.min()), (y - y.min()) / (y.max() - y.min())
    from scipy.spatial import ConvexHull
    hull = ConvexHull(np.c_[xn, yn])
    area = 0.5 * np.abs((xn[tri[:, 1]] - xn[tri[:, 0]]) * (yn[tri[:, 2]] - yn[tri[:, 0]]) - (xn[tri[:, 2]] - xn[tri[:, 0]]) * (yn[tri[:, 1]] - yn[tri[:, 0]]))
    assert abs(area.sum() - hull.volume) < 1e-9 and len(tri) == 2 * n - 2 - len(hull.vertices)
    for a, b, c in tri:  # empty circumcircle
        ax, ay, bx, by, cx, cy = xn[a], yn[a], xn[b], yn[b], xn[c], yn[c]
        d = 2 * (ax * (by - cy) + bx * (cy - ay) + cx * (ay - by))
        ux = ((ax * ax + ay * ay) * (by - cy) + (bx * bx + by * by) * (cy - ay) + (cx * cx + cy * cy) * (ay - by)) / d
        uy = ((ax * ax + ay * ay) * (cx - bx) + (bx * bx + by * by) * (ax - cx) + (cx * cx + cy * cy) * (bx - ax)) / d
        r2 = (ax - ux) ** 2 + (ay - uy) ** 2
        others = np.ones(n, bool)
        others[[a, b, c]] = False
        assert ((xn[others] - ux) ** 2 + (yn[others] - uy) ** 2 > r2 * (1 - 1e-9)).all()
    mirror = R.AMirror("mirror", R.TGeoBBox("mirrorbox", 50., 50., 50.))
    mirror.SetReflectance(g)
    mgr = scenes.make_the_world()
    mgr.GetTopVolume().AddNode(mirror, 1)
    mgr.CloseGeometry()
    ex = mgr.ExportScene()
    for k in range(200):
        a, b = 3e-5 + 4e-5 * rng.random(), 1.5 * rng.random()
        v = g.Interpolate(a, b)
        assert abs(oracle.orc_graph2d_interp(ex.desc_ptr(), 0, a, b) - v) < 1e-12
        assert v == 0.0 or abs(v - f(a, b)) < 1e-12


def test_containment_radius_of_gaussian_psf(oracle):  # src/AGeoUtil.cxx:198-308; closed form for a 2-D Gaussian
    rng = np.random.default_rng(1)
    n, sig = 400000, 0.7
    x, y = 1.5 + sig * rng.standard_normal(n), -0.4 + sig * rng.standard_normal(n)
    bins, stats = H.psf_histogram(x, y, 300, -3., 6., 300, -5., 4.)
    for frac in (0.8, 0.5):
        r, cx, cy = H.oracle_containment(oracle, bins, stats, 300, -3., 6., 300, -5., 4., frac)
        want = sig * math.sqrt(-2 * math.log(1 - frac))
        assert abs(r / want - 1) < 0.02 and abs(cx - 1.5) < 0.03 and abs(cy + 0.4) < 0.03
    # the containing circle really holds the requested fraction of the histogram (bin centres)
    r, cx, cy = H.oracle_containment(oracle, bins, stats, 300, -3., 6., 300, -5., 4., 0.8)
    cxs = -3. + (np.arange(300) + 0.5) * 9. / 300
    cys = -5. + (np.arange(300) + 0.5) * 9. / 300
    inside = ((cxs[None, :] - cx) ** 2 + (cys[:, None] - cy) ** 2 <= r * r)
    assert abs(bins.reshape(300, 300)[inside].sum() / bins.sum() - 0.8) < 2e-3


def tmm_general(oracle, ml, mode, pol, th, lam, reverse=0):
    import robast_b200 as R
    ex, mid = R.export_multilayer(ml)
    r, t = C.c_double(), C.c_double()
    assert oracle.orc_tmm_general(ex.desc_ptr(), mid, mode, pol, reverse, complex(th).real, complex(th).imag, lam, C.byref(r), C.byref(t)) == 0
    return r.value, t.value


def incoherent_stack(R):  # unittest_robast.py:711-759
    n0, n1, n2, n3 = R.ARefractiveIndex(1., 0.1), R.ARefractiveIndex(2., 0.2), R.ARefractiveIndex(3., 0.004), R.ARefractiveIndex(4., 0.2)
    d1, d2 = 100, 1000
    multi = R.AMultilayer(n0, n3)
    multi.InsertLayer(n1, d1)
    multi.InsertLayer(n2, d2, False)
    multi.InsertLayer(n1, d1)
    multi.InsertLayer(n2, d1)
    multi.InsertLayer(n3, d1)
    multi.InsertLayer(n1, d2, False)
    multi.InsertLayer(n3, d1)
    multi.InsertLayer(n1, d1)
    th_0 = np.lib.scimath.arcsin(1. / (1 + 0.1j) * math.sin(math.pi / 3.))
    return multi, [n0, n1, n2, n3], th_0


def test_kat_tmm_reversed_stack(R, oracle):  # unittest_robast.py:657-666: the reversed stack evaluated with reverse = True
    rev, keep = basic_stack(R, reverse=True)
    rs, rp = 0.37273208839139516, 0.37016110373044969
    ts, tp = 0.22604491247079261, 0.22824374314132009
    for pol, (r0, t0) in ((0, (rs, ts)), (1, (rp, tp))):
        r, t = tmm_general(oracle, rev, 0, pol, 0.1, 100., reverse=1)
        assert abs(r - r0) < 1e-12 and abs(t - t0) < 1e-12


def test_kat_incoherent_tmm(R, oracle):  # unittest_robast.py:711-781, values from tmm.inc_tmm
    multi, keep, th_0 = incoherent_stack(R)
    rs, ts = 0.3776110935131179, 1.2856977234844612e-05
    rp, tp = 0.03199545463016445, 2.0900281396463212e-05
    r, t = tmm_general(oracle, multi, 1, 0, th_0, 400.)
    assert abs(r - rs) < 1e-12 and abs(t / ts - 1) < 1e-10
    r, t = tmm_general(oracle, multi, 1, 1, th_0, 400.)
    assert abs(r - rp) < 1e-12 and abs(t / tp - 1) < 1e-10
    # an all-coherent stack: the incoherent TMM degenerates to the coherent one
    co, keep2 = basic_stack(R)
    for pol in (0, 1):
        a = tmm_general(oracle, co, 1, pol, 0.1, 100.)
        b = tmm_general(oracle, co, 0, pol, 0.1, 100.)
        assert abs(a[0] - b[0]) < 1e-13 and abs(a[1] - b[1]) < 1e-13


def test_refractiveindex_dot_info_parser(R, oracle, tmp_path):  # src/ARefractiveIndexDotInfo.cxx:24-105
    for sep, eol in ((",", "\n"), (",", "\r\n"), ("\t", "\n"), ("\t", "\r\n")):
        f = tmp_path / "nk.csv"
        rows_n = [(0.30, 1.50), (0.40, 1.47), (0.50, 1.46), (0.70, 1.455)]
        rows_k = [(0.30, 1e-6), (0.70, 3e-6)]
        txt = "wl" + sep + "n" + eol + "".join("%g%s%g%s" % (w, sep, v, eol) for w, v in rows_n)
        txt += "wl" + sep + "k" + eol + "".join("%g%s%g%s" % (w, sep, v, eol) for w, v in rows_k)
        f.write_bytes(txt.encode())
        idx = R.ARefractiveIndexDotInfo(str(f))
        assert abs(idx.GetRefractiveIndex(450 * nm) - 1.465) < 1e-12
        assert abs(idx.GetExtinctionCoefficient(500 * nm) - 2e-6) < 1e-18
        ex, iid = R.export_index(idx)
        assert abs(oracle.orc_index_n(ex.desc_ptr(), iid, 450 * nm) - 1.465) < 1e-12
        assert abs(oracle.orc_index_k(ex.desc_ptr(), iid, 500 * nm) - 2e-6) < 1e-18
    (tmp_path / "n.csv").write_text("wl,n\n0.4,1.5\n0.6,1.4\n")
    only_n = R.ARefractiveIndexDotInfo(str(tmp_path / "n.csv"))
    assert abs(only_n.GetRefractiveIndex(500 * nm) - 1.45) < 1e-12 and only_n.GetExtinctionCoefficient(500 * nm) == 0


# ---------------------------------------------------------------------------------------------- TGeoArb8 / TGeoXtru (oracle pins)
def _shape_of_type(ex, want):
    """id of the first shape of type `want` in an exported scene (rbg_scene_desc: nshapes at int32[2], shapes* at byte 88)"""
    import ctypes as C
    base = ex.desc_ptr()
    nshapes = C.cast(base, C.POINTER(C.c_int32))[2]
    shapes = C.cast(C.cast(base + 88, C.POINTER(C.c_void_p))[0], C.POINTER(C.c_int32))
    for i in range(nshapes):
        if shapes[7 * i] == want:
            return i
    raise KeyError(want)


def _solo(R, shape):
    import scenes
    mgr = scenes.make_the_world()
    comp = R.AMirror("m", shape)
    mgr.GetTopVolume().AddNode(comp, 1)
    mgr.CloseGeometry()
    return mgr, mgr.ExportScene(), comp


def _probe(oracle, ex, sid, pts, dirs):
    import ctypes as C
    out = []
    for p, d in zip(pts, dirs):
        pa, da, na = (C.c_double * 3)(*p), (C.c_double * 3)(*d), (C.c_double * 3)()
        inside = oracle.orc_shape_contains(ex.desc_ptr(), sid, pa)
        dist = oracle.orc_shape_dist(ex.desc_ptr(), sid, pa, da, 1 if inside else 0)
        n = (0., 0., 0.)
        if dist < 1e29:
            q = (C.c_double * 3)(*[p[k] + dist * d[k] for k in range(3)])
            oracle.orc_shape_normal(ex.desc_ptr(), sid, q, da, na)
            n = tuple(na)
        out.append((inside, dist, n))
    return out


def _random_probes(seed, n, box):
    rng = np.random.default_rng(seed)
    pts = (rng.random((n, 3)) * 2 - 1) * box
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    return pts.tolist(), dirs.tolist()


def _same(a, b, tol=1e-9):
    hit = compared = 0
    for (ia, da, na), (ib, db, nb) in zip(a, b):
        assert ia == ib
        if da > 1e29 or db > 1e29:
            assert da > 1e29 and db > 1e29
            continue
        compared += 1
        assert abs(da - db) < tol * max(1., abs(da)), (da, db)
        # normals agree except on edges, where the nearest face is ambiguous
        if abs(sum(x * y for x, y in zip(na, nb)) - 1) < 1e-9:
            hit += 1
    assert compared > 60 and hit > 0.97 * compared


def test_arb8_and_xtru_boxes_equal_tgeobbox(R, oracle):
    """closed form: a TGeoArb8 with equal rectangular faces and a TGeoXtru with a rectangular outline are the TGeoBBox of the same
    half lengths — Contains, DistFromInside/Outside and normals of the oracle's three restatements agree"""
    hx, hy, hz = 4., 2.5, 6.
    rect = [-hx, -hy, -hx, hy, hx, hy, hx, -hy]  # clockwise seen from +z
    _m1, ex_box, _k1 = _solo(R, R.TGeoBBox("b", hx, hy, hz))
    _m2, ex_arb, _k2 = _solo(R, R.TGeoArb8("a", hz, rect + rect))
    xt = R.TGeoXtru(2)
    xt.SetName("x")
    xt.DefinePolygon(rect[0::2], rect[1::2])
    xt.DefineSection(0, -hz)
    xt.DefineSection(1, hz)
    _m3, ex_xtru, _k3 = _solo(R, xt)
    pts, dirs = _random_probes(11, 600, np.array([9., 7., 11.]))
    # shape 0 is the world box in every export; the solid under test is the other TGeoBBox / the Arb8 / the Xtru
    ref = _probe(oracle, ex_box, 1, pts, dirs)
    assert sum(1 for r in ref if r[0]) > 30 and sum(1 for r in ref if not r[0] and r[1] < 1e29) > 30
    _same(ref, _probe(oracle, ex_arb, _shape_of_type(ex_arb, R.RBG_SHAPE_ARB8), pts, dirs))
    _same(ref, _probe(oracle, ex_xtru, _shape_of_type(ex_xtru, R.RBG_SHAPE_XTRU), pts, dirs))


def test_scaled_xtru_equals_arb8_frustum(R, oracle):
    """a TGeoXtru whose upper section is scaled and shifted is the TGeoArb8 with the corresponding upper vertices (planar faces)"""
    lo = [-4., -3., -5., 3., 4., 2., 3., -3.]
    sc, ox, oy, hz = 0.55, 0.8, -0.4, 5.
    up = [v * sc + (ox if i % 2 == 0 else oy) for i, v in enumerate(lo)]
    _m1, ex_arb, _k1 = _solo(R, R.TGeoArb8("a", hz, lo + up))
    xt = R.TGeoXtru(2)
    xt.SetName("x")
    xt.DefinePolygon(lo[0::2], lo[1::2])
    xt.DefineSection(0, -hz, 0., 0., 1.)
    xt.DefineSection(1, hz, ox, oy, sc)
    _m2, ex_xtru, _k2 = _solo(R, xt)
    pts, dirs = _random_probes(12, 600, np.array([8., 7., 9.]))
    ref = _probe(oracle, ex_arb, _shape_of_type(ex_arb, R.RBG_SHAPE_ARB8), pts, dirs)
    assert sum(1 for r in ref if r[0]) > 30
    _same(ref, _probe(oracle, ex_xtru, _shape_of_type(ex_xtru, R.RBG_SHAPE_XTRU), pts, dirs))


def test_twisted_arb8_hits_lie_on_the_ruled_surface(R, oracle):
    """twisted TGeoArb8 (upper face rotated against the lower one): every reported boundary point separates inside from outside,
    lateral hits satisfy the bilinear-patch equation of their face, and the normal is perpendicular to edge and ruling there"""
    import ctypes as C
    lo = [(-4., -4.), (-4., 4.), (4., 4.), (4., -4.)]
    a = math.radians(25.)
    up = [(x * math.cos(a) - y * math.sin(a), x * math.sin(a) + y * math.cos(a)) for x, y in lo]
    hz = 6.
    _m, ex, _k = _solo(R, R.TGeoArb8("tw", hz, [c for v in lo + up for c in v]))
    sid = _shape_of_type(ex, R.RBG_SHAPE_ARB8)
    pts, dirs = _random_probes(13, 800, np.array([9., 9., 10.]))
    res = _probe(oracle, ex, sid, pts, dirs)
    lateral = 0
    for p, d, (inside, dist, n) in zip(pts, dirs, res):
        if dist > 1e29:
            continue
        for sgn, want in ((-1, inside), (1, not inside)):
            q = (C.c_double * 3)(*[p[k] + (dist + sgn * 1e-7) * d[k] for k in range(3)])
            if dist + sgn * 1e-7 > 0:
                assert bool(oracle.orc_shape_contains(ex.desc_ptr(), sid, q)) == bool(want), (p, d, dist, sgn)
        q = [p[k] + dist * d[k] for k in range(3)]
        if abs(abs(q[2]) - hz) < 1e-9:
            assert abs(abs(n[2]) - 1) < 1e-12
            continue
        s = 0.5 * (q[2] + hz) / hz
        quad = [(l[0] + s * (u[0] - l[0]), l[1] + s * (u[1] - l[1])) for l, u in zip(lo, up)]
        # on the carrier line of one edge of the section at this height
        cross = [(q[0] - quad[i][0]) * (quad[(i + 1) % 4][1] - quad[i][1]) - (q[1] - quad[i][1]) * (quad[(i + 1) % 4][0] - quad[i][0]) for i in range(4)]
        i = int(np.argmin(np.abs(cross)))
        assert abs(cross[i]) < 1e-8
        e = (quad[(i + 1) % 4][0] - quad[i][0], quad[(i + 1) % 4][1] - quad[i][1], 0.)
        assert abs(sum(x * y for x, y in zip(n, e))) < 1e-9 * math.hypot(*e[:2])
        assert sum(x * y for x, y in zip(n, d)) >= 0
        lateral += 1
    assert lateral > 100


def test_empty_overlapping_holder_is_transparent(R, oracle):
    """semantic pin of the AddNodeOverlap ("MANY") restatement: a holder without daughters laid over the whole system (the role of
    `comp` in tutorials/AshraOptics.C:883-884,1117-1120) must not change where any ray ends — the ordinary nodes sharing its space
    are found from inside it (ONLY priority, sister candidates) exactly as without it"""
    import helpers as H
    import scenes
    import test_device_code_on_host as T
    res = []
    for holder in (False, True):
        mgr, _keep = scenes.overlapping_frame(nested=False, holder=holder, bars=False, stop_ring=True)
        rays = H.Rays(T.overlap_beam(50, 2.0))
        H.trace_with(oracle.orc_trace, mgr.ExportScene(), rays, H.opts(seed=3, limit=20, disable_fresnel=1), nthreads=4)
        res.append(rays)
    plain, held = res
    assert (plain.status == held.status).all()
    st = np.bincount(plain.status, minlength=6)
    assert st[3] > 200 and st[1] > 100
    done = plain.status != 2  # exiting rays end on the world box in both cases too, but cross the holder's own boundary on the way
    assert np.abs(plain.out[:3] - held.out[:3]).max() < 1e-9
    assert (plain.npoints[done] == held.npoints[done]).all()


def test_corsika_bunch_expansion_closed_form(R, oracle):
    """ACorsikaIACTFile::GetRayArray (src/ACorsikaIACTFile.cxx:71-133): ray count per bunch = iterations of `j < photons`, start
    point projected back along the direction to height z, arrival time shifted by the path at c/n — against a numpy restatement;
    the library's ray count (rbg_bunch_rays, host side) agrees"""
    import ctypes as C
    import helpers as H
    b, a = H.make_bunches(2000, 7)
    counts = np.where(a["photons"] > 0, np.ceil(a["photons"].astype(np.float64)), 0).astype(np.int64)
    total = int(counts.sum())
    nr = C.c_int64()
    R.check(R.rbg_bunch_rays(C.byref(b), C.byref(nr)))
    assert nr.value == total and total > 2000
    out = np.zeros((8, total))
    assert oracle.orc_shoot_bunches(C.byref(b), 0, total, *[out[i].ctypes.data for i in range(8)]) == 0
    idx = np.repeat(np.arange(2000), counts)
    f = {k: v.astype(np.float64)[idx] for k, v in a.items()}
    dist = (b.z - b.telescope_z) * (-1. / f["cz"])
    assert np.allclose(out[0], f["x"] - dist * f["cx"], rtol=0, atol=1e-9)
    assert np.allclose(out[1], f["y"] - dist * f["cy"], rtol=0, atol=1e-9)
    assert (out[2] == b.z).all()
    assert np.allclose(out[3], f["time"] * 1e-9 - dist / (2.99792458e10 / b.refractive_index), rtol=0, atol=1e-18)
    assert (out[4] == f["cx"]).all() and (out[5] == f["cy"]).all() and (out[6] == f["cz"]).all()
    fixed = f["lambda_"] != 0
    assert np.allclose(out[7][fixed], f["lambda_"][fixed] * 1e-7, rtol=1e-15)
    lam = out[7][~fixed] / 1e-7
    assert (~fixed).sum() > 500 and lam.min() >= 300. and lam.max() <= 600.
    # uniform in 1/lambda: the median of 1/lambda sits halfway between the ends
    assert abs(np.median(1. / lam) - 0.5 * (1 / 300. + 1 / 600.)) < 1.5e-4
    # a sub-range reproduces the same rays (streams are keyed by the global ray index)
    part = np.zeros((8, 500))
    assert oracle.orc_shoot_bunches(C.byref(b), 1234, 500, *[part[i].ctypes.data for i in range(8)]) == 0
    assert (part == out[:, 1234:1734]).all()


def test_arb8_and_xtru_from_points_place_the_solid_on_its_corners(R, oracle):
    """AGeoUtil::MakeArb8FromPoints / MakeXtruFromPoints (src/AGeoUtil.cxx:47-125): a prism given by the corners of its top face
    (clockwise seen from the top) and the bottom corner under the first one; the returned solid, placed with the returned
    TGeoCombiTrans, must hold points just inside every given corner and exclude points just outside"""
    import ctypes as C
    import scenes
    rng = np.random.default_rng(21)
    for kind in ("arb8", "xtru"):
        for trial in range(6):
            # an arbitrarily oriented prism: orthonormal frame (e1, e2, axis), top face in the plane spanned by e1, e2
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            if np.linalg.det(q) < 0:
                q[:, 2] = -q[:, 2]
            e1, e2, axis = q[:, 0], q[:, 1], q[:, 2]
            origin = rng.normal(size=3) * 30.
            height = 5. + 10. * rng.random()
            if kind == "arb8":
                outline = [(0., 0.), (0., 6.), (8., 7.), (9., -1.)]  # clockwise seen from +axis
            else:
                outline = [(0., 0.), (0., 6.), (5., 9.), (8., 7.), (9., -1.)]  # five corners, clockwise (concave outlines: test_device_code_on_host)
            top = [origin + u * e1 + v * e2 for u, v in outline]
            bottom0 = top[0] - height * axis
            vec = lambda a: R.TVector3(float(a[0]), float(a[1]), float(a[2]))  # noqa: E731
            if kind == "arb8":
                shape, combi = R.MakeArb8FromPoints("fp%s%d" % (kind, trial), vec(top[0]), vec(top[1]), vec(top[2]), vec(top[3]), vec(bottom0))
            else:
                shape, combi = R.MakeXtruFromPoints("fp%s%d" % (kind, trial), [vec(t) for t in top] + [vec(bottom0)])
            mgr = scenes.make_the_world()
            comp = R.AMirror("m%s%d" % (kind, trial), shape)
            mgr.GetTopVolume().AddNode(comp, 1, combi)
            mgr.CloseGeometry()
            rays = []
            centroid = np.mean(top, axis=0) - 0.5 * height * axis
            corners = [t - s * height * axis for t in top for s in (0., 1.)]
            for c in corners:  # shoot from far outside through a point 1 % inside / 1 % outside the corner (seen from the centroid)
                for f, want in ((0.99, True), (1.01, False)):
                    target = centroid + f * (c - centroid)
                    rays.append((target, want))
            inp = np.zeros((len(rays), 8))
            for i, (target, _w) in enumerate(rays):
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                inp[i, :3] = target - 1e-3 * d  # start right next to the probe point: inside the solid the first step ends within ~20 cm
                inp[i, 4:7] = d
                inp[i, 7] = 400e-7
            batch = H.Rays(inp)
            H.trace_with(oracle.orc_trace, mgr.ExportScene(), batch, H.opts(limit=2), nthreads=1)
            # a ray starting inside the mirror solid is stopped at once (typeCurrent == mirror), one starting outside flies on
            inside_flags = np.array([w for _t, w in rays])
            stopped = batch.status == 1
            assert (stopped[inside_flags]).all() and not (stopped[~inside_flags] & (np.linalg.norm(batch.out[:3].T - inp[:, :3], axis=1)[~inside_flags] < 0.5)).any(), (kind, trial)


# ----------------------------------------------------------------------------- daughter-box trees (round 2)
@pytest.mark.parametrize("cfg,theta,nside,kw", [(1, 1.0, 60, {}), (2, 0.0, 70, {}), (2, 3.0, 70, {}), (3, 2.0, 60, {}), (4, 0.1, 50, {}),
                                                (5, 0.0, 50, dict(rings=2)), (5, 25.0, 60, dict(rings=4))])
def test_voxel_lookup_is_bit_identical_to_the_full_daughter_walk(oracle, cfg, theta, nside, kw):
    """the oracle's daughter-box trees (the role of TGeoVoxelFinder) only skip daughters that could not have changed the step:
    every output bit equals the walk over all daughters (src/AOpticsManager.cxx:363 -> TGeoNavigator::FindNextBoundaryAndStep)"""
    from robast_b200 import configs
    mgr, _keep = configs.BUILDERS[cfg](**kw)
    ex = mgr.ExportScene()
    side = {2: 40.0, 4: 75.0}.get(kw.get("rings"))
    beam = configs.beam(cfg, theta, n_side=nside if cfg <= 3 else side)
    n = nside * nside
    o = H.opts(disable_fresnel=1 if cfg == 2 else 0, seed=31)
    out = []
    try:
        for vox in (0, 1):
            oracle.orc_set_voxels(vox)
            out.append(H.trace_with(oracle.orc_trace, ex, H.make_rays(oracle, beam, 0, n), o, nthreads=2))
    finally:
        oracle.orc_set_voxels(1)
    a, b = out
    assert np.array_equal(a.out, b.out) and np.array_equal(a.iout, b.iout)
    assert len(np.unique(a.status)) >= 2


def test_voxel_lookup_with_overlapping_nodes_and_history(oracle):
    """MANY nodes (cluster look-up goes through next_daughter_holding with `from` > 0) and a scene with 12 sisters"""
    import scenes
    mgr, _keep = scenes.overlapping_frame(nested=True)
    ex = mgr.ExportScene()
    rng = np.random.default_rng(8)
    n = 3000
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    inp = np.zeros((n, 8))
    inp[:, 0:3] = 150 * (rng.random((n, 3)) - 0.5)
    inp[:, 4:7], inp[:, 7] = v, 400e-7
    out = []
    try:
        for vox in (0, 1):
            oracle.orc_set_voxels(vox)
            out.append(H.trace_with(oracle.orc_trace, ex, H.Rays(inp), H.opts(seed=3, limit=20)))
    finally:
        oracle.orc_set_voxels(1)
    assert np.array_equal(out[0].out, out[1].out) and np.array_equal(out[0].iout, out[1].iout)
