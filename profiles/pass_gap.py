# Where does a short TraceNonSequential pass spend its wall time?  Host wall clock of one rbg_trace call against the summed
# CUDA-event time of its kernels, then the same nine passes issued from 1, 2 and 3 host threads (one scene handle, stream and
# output buffer each).  usage: pass_gap.py [n_side]
import sys, os, time, threading, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
import robast_b200 as R
from robast_b200 import configs
import helpers as H
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 3334
n = nside * nside
dev = torch.device('cuda:0')
mgr, keep = configs.davies_cotton()
ex = mgr.ExportScene()
ANG = [0.5 * k for k in range(9)]
inp = torch.empty((9, 8, n), dtype=torch.float64, device=dev)
for k, th in enumerate(ANG):
    d = H.shoot_desc(configs.beam(2, th, n_side=nside))
    R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[k, i].data_ptr() for i in range(8)], 0, None))
torch.cuda.synchronize()
def mk(k, out, iout):
    r = R.rbg_rays(); r.n = n; r.on_device = 1
    for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]): setattr(r, key, inp[k, i].data_ptr())
    for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]): setattr(r, key, out[i].data_ptr())
    for i, key in enumerate(["status", "last_node", "npoints"]): setattr(r, key, iout[i].data_ptr())
    return r
NT = 3
scenes, streams, outs, iouts = [], [], [], []
for t in range(NT):
    h = C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(h))); scenes.append(h)
    streams.append(torch.cuda.Stream())
    outs.append(torch.empty((7, n), dtype=torch.float64, device=dev)); iouts.append(torch.empty((3, n), dtype=torch.int32, device=dev))
op = H.opts(disable_fresnel=1, steps_per_launch=0, seed=5)
# --- one pass: wall vs kernels
for sort in ("auto", "0"):
    r = mk(4, outs[0], iouts[0])
    for _ in range(3): R.check(R.rbg_trace(scenes[0], C.byref(op), C.byref(r), None))
    torch.cuda.synchronize()
    R.rbg_profile_enable(1)
    bm, bn, cm, cn = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm), C.byref(cn))
    t0 = time.perf_counter()
    for _ in range(5): R.check(R.rbg_trace(scenes[0], C.byref(op), C.byref(r), None))
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm), C.byref(cn)); R.rbg_profile_enable(0)
    print("one pass (%d rays): wall %.3f ms, bounce kernels %.3f ms (%d launches), compaction %.3f ms (%d)" % (n, wall * 1e3, bm.value / 5, bn.value / 5, cm.value / 5, cn.value / 5), flush=True)
# --- nine passes from nt threads
def worker(t, nt, ks):
    with torch.cuda.stream(streams[t]):
        for k in ks:
            r = mk(k, outs[t], iouts[t])
            R.check(R.rbg_trace(scenes[t], C.byref(op), C.byref(r), C.c_void_p(streams[t].cuda_stream)))
for nt in (1, 2, 3):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ths = [threading.Thread(target=worker, args=(t, nt, list(range(t, 9, nt)))) for t in range(nt)]
        for th in ths: th.start()
        for th in ths: th.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("nine passes from %d host thread(s): %.2f ms -> %.4g rays/s" % (nt, dt * 1e3, 9 * n / dt), flush=True)
