"""Drop-in run (-m gpu): the reference's tutorial macros, UNMODIFIED, compiled against include/robast/ and librobast_b200.so and
executed on the GPU.  The reference sources never enter this repository: the executables are built in the build container,
where /root/reference is mounted, by `__graft_entry__.build()` (tests/build_tutorials.py) from the sources where they lie, into
the git-ignored tests/_build/tutorials/, and travel to the GPU box as binaries (the same arrangement as oracle/_ref).
Drawing is stubbed; with ROBAST_DRAW_SUMMARY set the histograms the macros would draw print their statistics, which are
checked against the optics of each telescope."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_build", "tutorials")
pytestmark = pytest.mark.gpu


def run_macro(macro, cwd=None):
    exe = os.path.join(BIN, macro)
    if not os.path.exists(exe):
        pytest.skip("tests/_build/tutorials/%s was not built (needs the reference checkout at build time)" % macro)
    env = dict(os.environ, ROBAST_DRAW_SUMMARY="1")
    out = subprocess.run([exe], capture_output=True, text=True, check=True, env=env, timeout=900, cwd=cwd).stdout
    rows = []
    for ln in out.splitlines():
        if ln.startswith("TH2 "):
            d = dict(re.findall(r'(\w+)=("[^"]*"|[-+\w.]+)', ln))
            rows.append({k: (v.strip('"') if v.startswith('"') else float(v)) for k, v in d.items()})
    return rows, out


def test_simple_parabolic_telescope_macro():
    rows, out = run_macro("SimpleParabolicTelescope")
    hists = [r for r in rows if str(r["name"]).startswith("hist")]
    assert len(hists) >= 3
    on = hists[0]
    assert on["entries"] > 10000 and on["rmsx"] < 1e-5 and on["rmsy"] < 1e-5  # on-axis parabola: a point image
    assert hists[-1]["rmsx"] > hists[1]["rmsx"] > on["rmsx"]  # coma grows with the field angle


def test_davies_cotton_macro():
    rows, out = run_macro("DaviesCotton")
    mir = [r for r in rows if r["name"] == "hMirror"]
    psf = [r for r in rows if r["name"] == ""]
    assert len(mir) == 1 and len(psf) == 8
    assert mir[0]["entries"] > 50000 and abs(mir[0]["meanx"]) < 0.05 and 2.0 < mir[0]["rmsx"] < 4.5  # hit points (m) over the 12 m dish
    assert all(p["entries"] > 40000 for p in psf[:6])
    assert 1.0 < psf[0]["rmsy"] < 10. and psf[7]["rmsy"] > psf[0]["rmsy"]  # Davies-Cotton PSF (mm): a few mm on axis, growing off axis


def test_hess1_and_mst_macros():
    rows, out = run_macro("HESS1")
    assert len(rows) == 6 and all(r["entries"] > 50000 for r in rows)
    assert rows[0]["rmsy"] < rows[5]["rmsy"] and rows[0]["rmsy"] < 15.
    rows, out = run_macro("MST")
    assert len(rows) >= 8 and all(r["entries"] > 10000 for r in rows[:8])


def test_schwarzschild_couder_macro():
    rows, out = run_macro("SchwarzschildCouder")
    assert len(rows) >= 8 and all(r["entries"] > 5000 for r in rows[:8])


def test_ashra_optics_macro():
    """tutorials/AshraOptics.C: Baker-Nunn optics with TGeoArb8 / TGeoXtru frame parts, composites nested six deep and two nodes
    placed with AddNodeOverlap over the whole system; 22 field angles x 30 wavelengths x 400 rays"""
    rows, out = run_macro("AshraOptics")
    spots = {r["name"]: r for r in rows if str(r["name"]).startswith("hist")}
    assert len(spots) >= 22
    on = spots["hist0_1"]
    # 400-ray grids over a 1.2 m square, 1 m aperture, frame obscurations: ~100 of 400 reach the focal sphere per wavelength;
    # the spot (mm on the focal sphere) stays within a small fraction of a millimetre out to 20 degrees (wide-field Baker-Nunn)
    assert on["entries"] > 30 * 60
    assert on["rmsx"] < 0.5 and on["rmsy"] < 0.5
    off = spots["hist20_1"]
    assert off["entries"] > 30 * 30 and off["rmsx"] < 1.0 and off["rmsy"] < 1.0


def graphs(out):
    """TGraph summaries printed by the display stubs: list of [(x, y), ...]"""
    res = []
    for ln in out.splitlines():
        if ln.startswith("TGraph "):
            res.append([tuple(map(float, p.split(":"))) for p in ln.split("points=")[1].strip().split(",") if p])
    return res


def at(graph, x):
    return min(graph, key=lambda p: abs(p[0] - x))[1]


@pytest.fixture
def tutorial_dir(tmp_path):
    """the macros open their data files relative to tutorials/: the n,k tables and the glass catalogue excerpt shipped in
    robast_b200/data under the names the macros use"""
    data = os.path.join(ROOT, "robast_b200", "data")
    tut, misc = tmp_path / "tutorials", tmp_path / "misc"
    tut.mkdir()
    misc.mkdir()
    import shutil
    for f in ("SiO2", "TiO2", "Al", "Si", "Si3N4"):
        shutil.copy(os.path.join(data, f + ".nk.txt"), str(tut / (f + ".txt")))
    shutil.copy(os.path.join(data, "nbk7.agf"), str(misc / "schottzemax-20180601.agf"))
    return str(tut)


def test_hex_winston_cone_macro():
    """tutorials/HexWinstonCone.C: collection efficiency of a hexagonal Winston cone against the angle of incidence — flat up to
    the cut-off asin(Rout/Rin) = 30 deg, gone beyond it"""
    rows, out = run_macro("HexWinstonCone")
    g = graphs(out)
    assert len(g) == 1 and len(g[0]) == 400
    assert all(0.95 < at(g[0], a) < 1.03 for a in (0., 5., 10., 15., 20.))
    assert 0.5 < at(g[0], 30.) < 0.85 and at(g[0], 35.) < 0.02 and at(g[0], 39.9) == 0.


def test_hex_okumura_cone_macro():
    """tutorials/HexOkumuraCone.C: Winston cone vs Bezier-profile (Okumura) cone, AGeoBezierPgon with 100 sections"""
    rows, out = run_macro("HexOkumuraCone")
    g = graphs(out)
    assert len(g) == 3 and len(g[0]) == 400 and len(g[1]) == 400
    win, oku = g[0], g[1]
    assert all(97. < at(c, a) < 102. for c in (win, oku) for a in (0., 10., 20.))
    assert at(oku, 25.) > at(win, 25.) > 85.  # the Okumura profile holds the plateau closer to the cut-off ...
    assert at(oku, 30.) < at(win, 30.) and at(oku, 35.) < at(win, 35.) < 3.  # ... and falls faster behind it


def test_abs_length_macro():
    """tutorials/AbsLengthTest.C: 10 000 photons from a point in a medium of 10 cm absorption length — exponential track lengths"""
    rows, out = run_macro("AbsLengthTest")
    h = [dict(re.findall(r'(\w+)=("[^"]*"|[-+\w.]+)', ln)) for ln in out.splitlines() if ln.startswith("TH1 ")]
    assert len(h) == 1 and float(h[0]["entries"]) == 10000
    assert abs(float(h[0]["mean"]) - 10.) < 0.4 and abs(float(h[0]["rms"]) - 10.) < 0.6


def test_edmund_optics_macro():
    """tutorials/EdmundOptics.C: achromatic doublet of Ohara glasses; prints n_d of S-FSL5 and S-TIH13 (the catalogue values)"""
    rows, out = run_macro("EdmundOptics")
    vals = [float(ln) for ln in out.splitlines() if re.fullmatch(r"\d\.\d+", ln.strip())]
    assert abs(vals[0] - 1.48749) < 2e-5 and abs(vals[1] - 1.74077) < 2e-5
    spots = [r for r in rows if str(r["name"]).startswith("spot")]
    assert len(spots) == 3 and all(r["entries"] > 3000 and r["rmsx"] < 8. for r in spots)  # micrometres
    assert spots[1]["rmsx"] < spots[0]["rmsx"]  # best corrected at the d line


def test_schmidt_cassegrain_macro(tutorial_dir):
    """tutorials/SchmidtCassegrain.C (BASELINE configs[3]): N-BK7 corrector from the AGF catalogue, spot of a few micrometres"""
    rows, out = run_macro("SchmidtCassegrain", cwd=tutorial_dir)
    assert len(rows) == 2 and all(r["entries"] > 500000 for r in rows)
    assert rows[0]["rmsx"] < 8. and rows[0]["rmsy"] < 8. and abs(rows[0]["meanx"]) < 0.1
    assert -1.5 < rows[1]["meany"] < -0.3  # 0.1 deg off axis: the centroid moves off the chief-ray position by under 2 um


def test_multilayer_macro(tutorial_dir):
    """tutorials/multilayer.C: transmittance of a UV-cut and an IR-cut dielectric stack (SiO2 / TiO2 quarter-wave layers) by the
    transfer-matrix method on the GPU, 300-800 nm"""
    rows, out = run_macro("multilayer", cwd=tutorial_dir)
    g = graphs(out)
    assert len(g) == 2 and len(g[0]) == 501
    uv, ir = g
    assert all(0. <= y <= 100.0001 for c in g for _x, y in c)
    assert at(uv, 350.) < 1e-3 and at(uv, 450.) > 95. and at(uv, 700.) > 85.
    assert at(ir, 500.) > 90. and at(ir, 700.) < 1. and at(ir, 800.) < 0.1
