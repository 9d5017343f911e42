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


def run_macro(macro):
    exe = os.path.join(BIN, macro)
    if not os.path.exists(exe):
        pytest.skip("tests/_build/tutorials/%s was not built (needs the reference checkout at build time)" % macro)
    env = dict(os.environ, ROBAST_DRAW_SUMMARY="1")
    out = subprocess.run([exe], capture_output=True, text=True, check=True, env=env, timeout=900).stdout
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
