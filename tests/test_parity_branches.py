"""Branches of the hot path that had no device-side parity test in round 1, per ray against the oracle:
AGeoWinstonCone2D (src/AGeoWinstonCone2D.cxx:120-428, tutorials/HexWinstonCone.C:54-57), ASchottFormula / ACauchyFormula /
AMixedRefractiveIndex (src/ASchottFormula.cxx:43-55, src/ACauchyFormula.cxx:40-46, include/AMixedRefractiveIndex.h:36-45),
QE(theta) (src/AFocalSurface.cxx:35-52, src/AOpticsManager.cxx:495-513), TH2 reflectance (src/AMirror.cxx:39-60).
Every case runs twice: on the host build of the device code (CPU, here) and on the CUDA path through the C ABI (-m gpu)."""
import math

import numpy as np
import pytest

import helpers as H
import parity_cases as P

BACKENDS = [pytest.param("emul", id="emul"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]
nm = 1e-7


@pytest.fixture
def backend(request, emul):
    return "gpu" if request.param == "gpu" else emul


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
@pytest.mark.parametrize("material", ["mirror", "glass"])
def test_winston2d_solid(oracle, backend, material):
    ex, inp, o, _keep = P.winston2d_solid(material)
    ref, got, rep = P.run(oracle, backend, ex, inp, o)
    assert P.clean(rep), rep
    # the beam does exercise the shape: a good share of the rays meets the cone, and glass rays take more than one step through it
    touched = (got.last_node >= 1) | (got.npoints > 2)
    assert touched.mean() > 0.2
    if material == "glass":
        assert (got.npoints >= 4).mean() > 0.05


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
@pytest.mark.parametrize("theta", [0.0, 12.0, 27.0, 38.0])
def test_winston2d_hex_intersection(oracle, backend, theta):
    """HexWinstonCone.C mode 1: the hexagonal cone as the intersection of three 2-D cones (nesting depth 3)"""
    ex, inp, o, _keep = P.winston2d_hex3(theta)
    ref, got, rep = P.run(oracle, backend, ex, inp, o)
    assert P.clean(rep), rep
    frac = (got.status == 3).mean()
    if theta != 27.0:  # acceptance cut-off at asin(rout/rin) = 30 deg; 27 deg sits on its shoulder
        assert (frac > 0.05) if theta < 25 else (frac < 0.05)


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_winston2d_hex_equals_winston_poly(oracle, backend):
    """the same guide built from AGeoWinstonConePoly(6) (mode 0) collects the same rays: the two shapes describe one solid"""
    from robast_b200 import configs
    ex3, inp, o, _k3 = P.winston2d_hex3(10.0, n=4000)
    mgr0, _k0 = configs.hex_winston_cone(rings=0, coating="ideal")

    def trace(ex):
        rays = H.make_rays(oracle, inp[0], 0, inp[1])
        return H.trace_gpu(ex, rays, o) if backend == "gpu" else H.trace_with(backend.emul_trace, ex, rays, o)

    a, b = trace(ex3), trace(mgr0.ExportScene())
    same = a.status == b.status
    assert same.mean() > 0.995  # rays grazing the seams between the three cones may differ
    foc = same & (a.status == 3)
    assert np.abs(a.pos[foc] - b.pos[foc]).max() < 1e-6


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
@pytest.mark.parametrize("which", ["schott", "cauchy", "mixed", "mixed_sell"])
@pytest.mark.parametrize("fresnel_off", [1, 0])
def test_index_formula_lenses(R, oracle, backend, which, fresnel_off):
    models, _keep = P.index_formulas(R)
    name, index, closed = next(m for m in models if m[0] == which)
    if closed is not None:  # the host classes, the oracle and the closed forms agree on n(lambda)
        for lam_um in (0.3, 0.4358, 0.5876, 0.7):
            assert abs(index.GetRefractiveIndex(lam_um * 1e-4) - closed(lam_um)) < 1e-12
    ex, inp, o, _k = P.dispersive_lens_case(index, disable_fresnel=fresnel_off)
    ref, got, rep = P.run(oracle, backend, ex, inp, o)
    assert P.clean(rep), rep
    entered = got.npoints >= 3
    assert entered.mean() > 0.3
    if fresnel_off and closed is not None:
        # Snell at the entrance face, from the rays that went straight through two parallel faces: direction restored exactly
        through = (got.npoints == 4) & (got.status == 2)
        d_in = inp[through, 4:7]
        assert through.sum() > 100 and np.abs(got.dirs[through] - d_in).max() < 1e-9


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
@pytest.mark.parametrize("with_lambda,with_angle,p", [(False, False, 1.0), (True, False, 0.5), (False, True, 0.5), (True, True, 0.25)])
def test_quantum_efficiency_lambda_and_angle(R, oracle, backend, with_lambda, with_angle, p):
    ex, inp, o, _keep = P.qe_case(R, with_lambda, with_angle)
    ref, got, rep = P.run(oracle, backend, ex, inp, o)
    assert P.clean(rep), rep
    n = got.n
    nf, ns = int((got.status == 3).sum()), int((got.status == 1).sum())
    assert nf + ns == n
    if p == 1.0:
        assert nf == n
    else:
        assert abs(nf - n * p) < 3 * math.sqrt(n * p * (1 - p))  # unittest_robast.py:514-520


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_th2_mirror_reflectance(R, oracle, backend):
    ex, inp, o, keep = P.th2_mirror_case()
    ref, got, rep = P.run(oracle, backend, ex, inp, o)
    assert P.clean(rep), rep
    h, mirror, ang = keep[-3], keep[-2], keep[-1]
    lam = inp[:, 7]
    # expected reflected fraction = mean of TH2::Interpolate over the beam (host class == oracle table lookup)
    expect = np.array([min(1., max(0., h.Interpolate(l, a))) for l, a in zip(lam[:3000], ang[:3000])])
    refl = (got.status[:3000] == 2)
    assert abs(refl.mean() - expect.mean()) < 4 * math.sqrt(expect.mean() * (1 - expect.mean()) / 3000)
    out_of_range = (lam < 300 * nm) | (lam >= 500 * nm)
    assert out_of_range.sum() > 500 and (got.status[out_of_range] == 5).all()  # TH2::Interpolate gives 0 outside the axis range
