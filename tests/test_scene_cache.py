"""The flattened scene is kept across TraceNonSequential calls until a mutator ran (round-1 advice: every call re-exported and
re-hashed the whole scene): the geometry epoch moves with every mutator and with nothing else (CPU), and a loop of small calls
as tutorials/Optimize.C / optimize_multilayer.C make them gives the same rays before and after, and new rays after a change
(-m gpu)."""
import math

import numpy as np
import pytest

import helpers as H
import scenes


def test_epoch_moves_with_mutators_only(R):
    e0 = R.RbGeomEpoch()
    box = R.TGeoBBox("b", 1., 1., 1.)
    tr = R.TGeoTranslation("t", 0., 0., 1.)
    g = R.TGraph()
    e1 = R.RbGeomEpoch()
    ix = R.ARefractiveIndex(1.5)
    e1 = R.RbGeomEpoch()
    # read-only calls, shooting rays and building the matrices a beam is placed with leave the epoch alone
    ix.GetRefractiveIndex(400e-7)
    rays = R.ARayShooter.Square(400e-7, 10., 3, R.TGeoRotation("rayrot", 90., 190., 0.), R.TGeoTranslation("raytr", 0., 0., 80.), R.TVector3(0., 0., -1.))
    assert rays.GetRunning().GetLast() == 8
    assert R.RbGeomEpoch() == e1 >= e0
    seen = e1
    for mutate in (lambda: g.SetPoint(0, 300e-7, 0.5), lambda: tr.SetTranslation(1., 2., 3.), lambda: R.TGeoRotation("r", 0., 0., 0.).SetAngles(10., 20., 30.)):
        mutate()
        assert R.RbGeomEpoch() > seen
        seen = R.RbGeomEpoch()
    mgr, mirror, keep = scenes.mirror_box_with_border()
    seen = R.RbGeomEpoch()
    mirror.SetReflectance(0.5)
    assert R.RbGeomEpoch() > seen
    air, al = R.ARefractiveIndex(1., 0.), R.ARefractiveIndex(1.2, 6.0)
    ml = R.AMultilayer(air, al)
    seen = R.RbGeomEpoch()
    ml.InsertLayer(R.ARefractiveIndex(1.46), 25e-7)
    assert R.RbGeomEpoch() > seen
    seen = R.RbGeomEpoch()
    ml.ChangeThickness(1, 30e-7)
    assert R.RbGeomEpoch() > seen


@pytest.mark.gpu
def test_small_call_loop_keeps_the_scene_and_sees_changes(R):
    """1000 TraceNonSequential calls of 1000 rays (the MINUIT-loop regime): same geometry -> same scene handle, same results;
    a changed reflectance / layer thickness / placement is picked up by the next call"""
    import time
    mgr, mirror, keep = scenes.mirror_box_with_border()
    mgr.SetSeed(7)

    def shoot():
        return R.ARayShooter.Square(400e-7, 10., 32, None, R.TGeoTranslation("raytr", 0., 0., 80.), R.TVector3(0., 0., -1.))

    def exited(arr):
        return arr.GetExited().GetLast() + 1

    first = shoot()
    mgr.TraceNonSequential(first)
    assert exited(first) == 1024  # R = 1
    epoch = R.RbGeomEpoch()
    t0 = time.perf_counter()
    calls = 1000
    for _ in range(calls):
        a = R.ARayArray()
        for i in range(0, 1024, 64):  # rays added one by one, as the reference's loops do
            a.Add(R.ARay(i, 400e-7, 0.01 * i, 0., 80., 0., 0., 0., -1.))
        mgr.TraceNonSequential(a)
        assert exited(a) == 16
    dt = (time.perf_counter() - t0) / calls
    assert R.RbGeomEpoch() == epoch  # tracing and building rays does not touch the geometry
    print("small-call loop: %.1f us per TraceNonSequential call" % (dt * 1e6))
    assert dt < 5e-3
    mirror.SetReflectance(0.25)  # mutator -> next call re-exports
    b = shoot()
    mgr.SetSeed(7)
    mgr.TraceNonSequential(b)
    n = exited(b)
    assert abs(n - 256) < 5 * math.sqrt(1024 * 0.25 * 0.75)
    mirror.SetReflectance(1.0)
    c = shoot()
    mgr.TraceNonSequential(c)
    assert exited(c) == 1024


@pytest.mark.gpu
def test_tobjarray_of_rays_is_traced_in_one_batch(R):
    """TraceNonSequential(TObjArray*) (src/AOpticsManager.cxx:335): every running ARay of the array, same results as the batch"""
    mgr, mirror, keep = scenes.mirror_box_with_border()
    arr = R.TObjArray()
    rays = [R.ARay(i, 400e-7, 0.1 * i, 0., 80., 0., 0.05, 0., -1.) for i in range(50)]
    for r in rays:
        arr.Add(r)
    l0 = R.rbg_launch_count()
    mgr.TraceNonSequential(arr)
    assert R.rbg_launch_count() - l0 <= 4  # one batch (trace + the two kernels that size its polyline record), not 50 single-ray traces
    for i, r in enumerate(rays):
        assert r.IsExited()
        d = r.GetDirection()
        assert d[2] > 0.99 and abs(d[0] - 0.05 / math.sqrt(1 + 0.05 ** 2)) < 1e-12
