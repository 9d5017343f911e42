# throughput of one config under several compiled kernel instantiations (RB_VARIANT), device-resident rays
# usage: sweep_variants.py <cfg> <theta> <n> name1,name2,... [json kwargs of the config builder] [beam side]
# (x* names need a library built with `make EXP=1`)
import sys, os, json, ctypes as C, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
import robast_b200 as R
from robast_b200 import configs
import helpers as H
cfg, theta, n = int(sys.argv[1]), float(sys.argv[2]), int(float(sys.argv[3]))
names = sys.argv[4].split(',')
dev = torch.device('cuda:0')
kw = json.loads(sys.argv[5]) if len(sys.argv) > 5 else {}
mgr, keep = configs.BUILDERS[cfg](**kw)
ex = mgr.ExportScene()
nside = int(round(n ** 0.5)) if cfg <= 3 else (float(sys.argv[6]) if len(sys.argv) > 6 else None)
d = H.shoot_desc(configs.beam(cfg, theta, n_side=nside))
inp = torch.empty((8, n), dtype=torch.float64, device=dev); o = torch.empty((7, n), dtype=torch.float64, device=dev); io = torch.empty((3, n), dtype=torch.int32, device=dev)
R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[i].data_ptr() for i in range(8)], 0, None))
r = R.rbg_rays(); r.n = n; r.on_device = 1
for i, k in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]): setattr(r, k, inp[i].data_ptr())
for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]): setattr(r, k, o[i].data_ptr())
for i, k in enumerate(["status", "last_node", "npoints"]): setattr(r, k, io[i].data_ptr())
ref = None
for name in names:
    if name == 'default': os.environ.pop('RB_VARIANT', None)
    else: os.environ['RB_VARIANT'] = name
    h = C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(h)))
    op = H.opts(disable_fresnel=1 if cfg == 2 else 0, steps_per_launch=0, seed=5)
    for _ in range(2): R.check(R.rbg_trace(h, C.byref(op), C.byref(r), None))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): R.check(R.rbg_trace(h, C.byref(op), C.byref(r), None))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    st = np.bincount(io[0].cpu().numpy(), minlength=6).tolist()
    chk = float(o[0].double().sum().item())
    if ref is None: ref = (st, chk)
    print("cfg%d %-12s -> %-24s %.4g rays/s  same_as_first=%s" % (cfg, name, R.rbg_scene_kernel_variant(h).decode(), n / dt, (st, chk) == ref), flush=True)
    R.rbg_scene_destroy(h)
