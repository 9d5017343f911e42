"""Several GPUs behind the C ABI (rbg_multi_*, include/robast_b200.h): the reference fans a batch out over threads inside
TraceNonSequential (src/AOpticsManager.cxx:529-568); here the contiguous chunks go to GPUs.  The results must not depend on the
number of devices: bit-identical ray tables, identical histograms and counters.  The single-device cases run on any GPU box; the
two-device cases skip below two GPUs."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from robast_b200 import configs

pytestmark = pytest.mark.gpu


def _multi(R, ex, ndev):
    m = C.c_void_p()
    R.check(R.rbg_multi_create(ex.desc_ptr(), ndev, None, C.byref(m)))
    assert R.rbg_multi_num_devices(m) == ndev
    return m


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_trace_equals_single_device(R, oracle, ndev):
    if R.rbg_device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    mgr, _k = configs.schmidt_cassegrain()  # stochastic: exercises the global Philox ray ids
    ex = mgr.ExportScene()
    n = 1_200_003  # not divisible by 2: the last device takes the remainder
    beam = configs.beam(4, 0.05)
    o = H.opts(seed=20180601, ray_id_offset=12345)
    ref = H.trace_gpu(ex, H.make_rays(oracle, beam, 0, n), o)
    got = H.make_rays(oracle, beam, 0, n)
    m = _multi(R, ex, ndev)
    try:
        r = got.struct()
        R.check(R.rbg_multi_trace(m, C.byref(o), C.byref(r)))
    finally:
        R.rbg_multi_destroy(m)
    assert np.array_equal(ref.out.view(np.int64), got.out.view(np.int64)) and np.array_equal(ref.iout, got.iout)
    assert len(np.unique(got.status)) >= 4


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_shoot_trace_reduce(R, ndev):
    """generate + trace + reduce on the devices (the 1e9-ray pattern of BASELINE configs[4], here 3e6 rays in batches of 7e5):
    histogram and status counters identical to one device doing the same with the single-GPU calls"""
    import torch
    if R.rbg_device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    mgr, _k = configs.hex_winston_cone(rings=2)
    ex = mgr.ExportScene()
    n, batch = 3_000_001, 700_000
    d = H.shoot_desc(configs.beam(5, 15.0, n_side=20.0))
    o = H.opts(seed=20110306, ray_id_offset=0)
    nx = ny = 64
    # single-GPU reference with the plain calls
    dev = torch.device("cuda:0")
    h = C.c_void_p()
    R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(h)))
    inp = torch.empty((8, n), dtype=torch.float64, device=dev)
    out = torch.empty((7, n), dtype=torch.float64, device=dev)
    io = torch.empty((3, n), dtype=torch.int32, device=dev)
    hist = torch.zeros(nx * ny, dtype=torch.int64, device=dev)
    mom = torch.zeros(8, dtype=torch.float64, device=dev)
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[i].data_ptr() for i in range(8)], 0, None))
    r = R.rbg_rays()
    r.n, r.on_device = n, 1
    for i, k in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
        setattr(r, k, inp[i].data_ptr())
    for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
        setattr(r, k, out[i].data_ptr())
    for i, k in enumerate(["status", "last_node", "npoints"]):
        setattr(r, k, io[i].data_ptr())
    R.check(R.rbg_trace(h, C.byref(o), C.byref(r), None))
    R.check(R.rbg_hist2d(n, out[0].data_ptr(), out[1].data_ptr(), io[0].data_ptr(), R.RBG_FOCUSED, nx, -12., 12., ny, -12., 12., hist.data_ptr(), 0, None))
    R.check(R.rbg_moments(n, out[0].data_ptr(), out[1].data_ptr(), out[3].data_ptr(), io[0].data_ptr(), R.RBG_FOCUSED, mom.data_ptr(), cnt.data_ptr(), 0, None))
    torch.cuda.synchronize()
    R.rbg_scene_destroy(h)
    # the library's multi-device pipeline
    mh = np.zeros(nx * ny, dtype=np.uint64)
    mm = np.zeros(8)
    mc = np.zeros(6, dtype=np.int64)
    m = _multi(R, ex, ndev)
    try:
        R.check(R.rbg_multi_shoot_trace_reduce(m, C.byref(o), C.byref(d), n, batch, R.RBG_FOCUSED, nx, -12., 12., ny, -12., 12., mh.ctypes.data, mm.ctypes.data, mc.ctypes.data))
    finally:
        R.rbg_multi_destroy(m)
    assert np.array_equal(mh.astype(np.int64), hist.cpu().numpy())
    assert np.array_equal(mc, cnt.cpu().numpy()) and mc.sum() == n and mc[3] > 0.2 * n
    assert np.allclose(mm[:7], mom.cpu().numpy()[:7], rtol=1e-10, atol=1e-12)  # sums of doubles: the order of the atomics differs


def test_cpp_manager_fans_out_with_set_max_threads(R):
    """AOpticsManager::SetMaxThreads(n) + SetMultiThread (tutorials/DaviesCotton.C:30-32): n GPUs when the box has them, results
    independent of n"""
    mgr, _k = configs.davies_cotton()
    mgr.SetSeed(1)
    n_side = 800
    a = R.ARayShooter.Square(400e-7, 1400., n_side, R.TGeoRotation("rayrot", 0., 0., 0.), R.TGeoTranslation("raytr", 0., 0., 3200.), R.TVector3(0., 0., -1.))
    mgr.TraceNonSequential(a)
    mgr2, _k2 = configs.davies_cotton()
    mgr2.SetSeed(1)
    mgr2.SetMultiThread(True)
    mgr2.SetMaxThreads(8)
    assert mgr2.GetNumberOfGPUs() == min(8, R.rbg_device_count())
    b = R.ARayShooter.Square(400e-7, 1400., n_side, R.TGeoRotation("rayrot", 0., 0., 0.), R.TGeoTranslation("raytr", 0., 0., 3200.), R.TVector3(0., 0., -1.))
    mgr2.TraceNonSequential(b)
    for get in ("GetFocused", "GetStopped", "GetExited"):
        assert getattr(a, get)().GetLast() == getattr(b, get)().GetLast()
    assert a.GetFocused().GetLast() > 0.3 * n_side * n_side
