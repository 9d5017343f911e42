# Host-buffer path of rbg_trace: nine calls of 1.11e7 rays (the bench's e2e arm) against one call of 1.0e8 rays over the same
# pinned buffers — separates the pipeline's fill/drain per call from its steady state.
import sys, time, ctypes as C
import os; _r = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _r); sys.path.insert(0, os.path.join(_r, 'tests'))
import torch
import robast_b200 as R
from robast_b200 import configs
import helpers as H
nside = 3334; n = nside * nside; nang = 9
mgr, keep = configs.davies_cotton(); ex = mgr.ExportScene()
sc = C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(), 0, C.byref(sc)))
dev = torch.device('cuda:0')
tmp = torch.empty((8, n), dtype=torch.float64, device=dev)
hin = torch.empty((8, nang * n), dtype=torch.float64).pin_memory()
for k in range(nang):
    d = H.shoot_desc(configs.beam(2, 0.5 * k, n_side=nside))
    R.check(R.rbg_shoot(C.byref(d), 0, n, *[tmp[i].data_ptr() for i in range(8)], 0, None))
    torch.cuda.synchronize()
    hin[:, k * n:(k + 1) * n].copy_(tmp.cpu())
hout = torch.empty((7, nang * n), dtype=torch.float64).pin_memory(); hio = torch.empty((3, nang * n), dtype=torch.int32).pin_memory()
def struct(first, count):
    r = R.rbg_rays(); r.n = count; r.on_device = 0
    for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]): setattr(r, key, hin[i].data_ptr() + 8 * first)
    for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]): setattr(r, key, hout[i].data_ptr() + 8 * first)
    for i, key in enumerate(["status", "last_node", "npoints"]): setattr(r, key, hio[i].data_ptr() + 4 * first)
    return r
op = H.opts(disable_fresnel=1, steps_per_launch=int(sys.argv[1]) if len(sys.argv) > 1 else 0, seed=5)
def nine():
    for k in range(nang):
        op.ray_id_offset = k * n
        R.check(R.rbg_trace(sc, C.byref(op), C.byref(struct(k * n, n)), None))
def one():
    op.ray_id_offset = 0
    R.check(R.rbg_trace(sc, C.byref(op), C.byref(struct(0, nang * n)), None))
for name, f in (("nine calls", nine), ("one call", one), ("nine calls", nine), ("one call", one)):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); f(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("%-10s %.1f ms  %.4g rays/s  (%.1f GB/s H2D+D2H)" % (name, dt * 1e3, nang * n / dt, nang * n * 132 / dt / 1e9), flush=True)
