"""per-ray difference between the wavefront (k_nav/k_shade) and the single-launch (k_trace) modes on one config"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H, robast_b200 as R
from robast_b200 import configs
cfg, theta, n = int(sys.argv[1]), float(sys.argv[2]), int(float(sys.argv[3]))
kw = dict(rings=10) if cfg == 5 else {}
mgr, keep = configs.BUILDERS[cfg](**kw); ex = mgr.ExportScene()
h = C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(h)))
nside = int(round(n ** 0.5)) if cfg <= 3 else (84.0 if cfg == 5 else None)
if cfg <= 3: n = nside * nside
d = H.shoot_desc(configs.beam(cfg, theta, n_side=nside))
dev = torch.device("cuda:0")
inp = torch.empty((8, n), dtype=torch.float64, device=dev)
R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[i].data_ptr() for i in range(8)], 0, None))
res = []
for spl in (0, -1):
    o = torch.zeros((7, n), dtype=torch.float64, device=dev); io = torch.zeros((3, n), dtype=torch.int32, device=dev)
    r = R.rbg_rays(); r.n, r.on_device = n, 1
    for i, k in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]): setattr(r, k, inp[i].data_ptr())
    for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]): setattr(r, k, o[i].data_ptr())
    for i, k in enumerate(["status", "last_node", "npoints"]): setattr(r, k, io[i].data_ptr())
    op = H.opts(disable_fresnel=1 if cfg == 2 else 0, steps_per_launch=spl, seed=5)
    R.check(R.rbg_trace(h, C.byref(op), C.byref(r), None)); torch.cuda.synchronize()
    res.append((o.cpu().numpy(), io.cpu().numpy()))
(a, ia), (b, ib) = res
neq = (a.view(np.int64) != b.view(np.int64)).any(axis=0) | (ia != ib).any(axis=0)
print("cfg%d n=%d differing rays: %d; status diff %d npoints diff %d node diff %d; max |dpos| %.3g max |ddir| %.3g" % (
    cfg, n, neq.sum(), (ia[0] != ib[0]).sum(), (ia[2] != ib[2]).sum(), (ia[1] != ib[1]).sum(), np.abs(a[:3] - b[:3]).max(), np.abs(a[4:7] - b[4:7]).max()))
idx = np.where(neq)[0][:5]
for i in idx: print(i, ia[:, i], ib[:, i], a[:, i] - b[:, i])
