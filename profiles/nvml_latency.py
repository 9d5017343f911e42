import sys, time, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, pynvml as N
import robast_b200 as R
from robast_b200 import configs
import helpers as H
N.nvmlInit(); h = N.nvmlDeviceGetHandleByIndex(0)
dev = torch.device('cuda:0'); n = 3334 * 3334
mgr, keep = configs.davies_cotton(); ex = mgr.ExportScene()
inp = torch.empty((8, n), dtype=torch.float64, device=dev); out = torch.empty((7, n), dtype=torch.float64, device=dev); io = torch.empty((3, n), dtype=torch.int32, device=dev)
d = H.shoot_desc(configs.beam(2, 2.0, n_side=3334)); R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[i].data_ptr() for i in range(8)], 0, None))
r = R.rbg_rays(); r.n = n; r.on_device = 1
for i, k in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]): setattr(r, k, inp[i].data_ptr())
for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]): setattr(r, k, out[i].data_ptr())
for i, k in enumerate(["status", "last_node", "npoints"]): setattr(r, k, io[i].data_ptr())
sc = C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(sc)))
op = H.opts(disable_fresnel=1, steps_per_launch=0, seed=5)
for _ in range(3): R.check(R.rbg_trace(sc, C.byref(op), C.byref(r), None))
calls = {"clock": lambda: N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), "reasons": lambda: N.nvmlDeviceGetCurrentClocksEventReasons(h), "none": lambda: None,
         "power": lambda: N.nvmlDeviceGetPowerUsage(h)}
for name, fn in calls.items():
    lat, tr = [], []
    for it in range(40):
        t0 = time.perf_counter(); R.check(R.rbg_trace(sc, C.byref(op), C.byref(r), None)); t1 = time.perf_counter(); fn(); t2 = time.perf_counter()
        tr.append((t1 - t0) * 1e3); lat.append((t2 - t1) * 1e3)
    tr.sort(); lat.sort()
    print("%-8s call ms: median %.3f p90 %.3f max %.3f | trace ms: median %.2f p90 %.2f max %.2f" % (name, lat[20], lat[36], lat[-1], tr[20], tr[36], tr[-1]), flush=True)
