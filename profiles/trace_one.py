#!/usr/bin/env python
"""Trace one BASELINE config with device-resident rays and print rays/s (profiling driver: ncu wraps this).
usage: trace_one.py CFG THETA_DEG N [reps] [key=value ...]   e.g.  trace_one.py 5 20 10000000 3 rings=10 precalc=1"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers as H
import robast_b200 as R
from robast_b200 import configs


def main():
    cfg, theta, n = int(sys.argv[1]), float(sys.argv[2]), int(float(sys.argv[3]))
    reps = int(sys.argv[4]) if len(sys.argv) > 4 and "=" not in sys.argv[4] else 3
    kw = {}
    spl = 0
    for a in sys.argv[4:]:
        if "=" in a:
            k, v = a.split("=")
            if k == "spl":
                spl = int(v)
            else:
                kw[k] = int(v) if v.lstrip("-").isdigit() else v
    if "precalc" in kw:
        kw["precalc"] = bool(kw["precalc"])
    mgr, keep = configs.BUILDERS[cfg](**kw)
    ex = mgr.ExportScene()
    h = C.c_void_p()
    R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(h)))
    nside = int(round(n ** 0.5)) if cfg <= 3 else None
    if nside:
        n = nside * nside
    p = configs.beam(cfg, theta, n_side=nside)
    if cfg == 5:
        p = configs.beam(5, theta, n_side=(2 * kw.get("rings", 2) + 1) * 4.0)
    d = H.shoot_desc(p)
    dev = torch.device("cuda:0")
    inp = torch.empty((8, n), dtype=torch.float64, device=dev)
    o = torch.empty((7, n), dtype=torch.float64, device=dev)
    io = torch.empty((3, n), dtype=torch.int32, device=dev)
    R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[i].data_ptr() for i in range(8)], 0, None))
    r = R.rbg_rays()
    r.n, r.on_device = n, 1
    for i, k in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
        setattr(r, k, inp[i].data_ptr())
    for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
        setattr(r, k, o[i].data_ptr())
    for i, k in enumerate(["status", "last_node", "npoints"]):
        setattr(r, k, io[i].data_ptr())
    op = H.opts(disable_fresnel=1 if cfg == 2 else 0, steps_per_launch=spl, seed=5)
    R.check(R.rbg_trace(h, C.byref(op), C.byref(r), None))
    torch.cuda.synchronize()
    R.rbg_profile_enable(1)
    l0 = R.rbg_launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        R.check(R.rbg_trace(h, C.byref(op), C.byref(r), None))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    bm, bn, cm, cn = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm), C.byref(cn))
    st = np.bincount(io[0].cpu().numpy(), minlength=6)
    print("cfg%d theta=%.1f %s n=%d variant=%s spl=%d: %.4g rays/s, %.3f ms/trace, %d launches/trace, bounce kernels %.3f ms (%d), compaction %.3f ms; status=%s mean npoints=%.2f"
          % (cfg, theta, kw, n, R.rbg_scene_kernel_variant(h).decode(), spl, n / dt, dt * 1e3, (R.rbg_launch_count() - l0) // reps, bm.value / reps, bn.value // reps,
             cm.value / reps, st.tolist(), io[2].float().mean().item()), flush=True)
    R.rbg_scene_destroy(h)


if __name__ == "__main__":
    main()
