#!/usr/bin/env python
"""What happens to the candidate daughters of a boundary step (host build of the device code with counters, no GPU needed):
listed by the box search, dropped because their box starts beyond the step already found, dropped by the node-frame box test,
evaluated (DistFromOutside), hit.  usage: cand_stats.py CFG THETA_DEG N [key=value ...]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers as H
import robast_b200 as R
from robast_b200 import configs

so = "/tmp/libemul_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DRB_EMUL_STATS", "-o", so, os.path.join(ROOT, "tests/emul/emul.cpp")])
lib = C.CDLL(so)
lib.emul_trace.restype = C.c_int
lib.emul_trace.argtypes = [C.c_void_p, C.POINTER(R.rbg_trace_opts), C.POINTER(R.rbg_rays), C.c_int]
cfg, theta, n = int(sys.argv[1]), float(sys.argv[2]), int(float(sys.argv[3]))
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[4:])}
mgr, keep = configs.BUILDERS[cfg](**kw)
ex = mgr.ExportScene()
oracle = H.load_oracle()
nside = int(round(n ** 0.5))
p = configs.beam(cfg, theta, n_side=nside) if cfg <= 3 else configs.beam(cfg, theta, **({"n_side": (2 * kw.get("rings", 2) + 1) * 4.0} if cfg == 5 else {}))
rays = H.make_rays(oracle, p, 0, nside * nside if cfg <= 3 else n)
H.trace_with(lib.emul_trace, ex, rays, H.opts(disable_fresnel=1 if cfg == 2 else 0, seed=5))
out = (C.c_longlong * 48)()
lib.emul_stats(out, 1)
a = np.array(list(out)).reshape(6, 8)
print("step   rays   listed/ray  far/ray  boxed/ray  evaluated/ray  hits/ray")
for k in range(8):
    if a[0, k]:
        print("%4d %8d %9.2f %9.2f %9.2f %11.2f %10.2f" % ((k, a[0, k]) + tuple(a[j, k] / a[0, k] for j in range(1, 6))))
