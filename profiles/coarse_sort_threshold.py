# DaviesCotton, 1 deg, device-resident grid beams of different sizes: wavefront rays/s (run with RB_SORT_COARSE_MIN=0 and unset)
import sys, ctypes as C, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
import robast_b200 as R
from robast_b200 import configs
import helpers as H
dev=torch.device('cuda:0')
mgr,keep=configs.BUILDERS[2](); ex=mgr.ExportScene()
h=C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(),0,C.byref(h)))
for nside in (128,256,362,512,724,1000,1448):
    n=nside*nside; d=H.shoot_desc(configs.beam(2,1.0,n_side=nside))
    inp=torch.empty((8,n),dtype=torch.float64,device=dev); o=torch.empty((7,n),dtype=torch.float64,device=dev); io=torch.empty((3,n),dtype=torch.int32,device=dev)
    R.check(R.rbg_shoot(C.byref(d),0,n,*[inp[i].data_ptr() for i in range(8)],0,None))
    r=R.rbg_rays(); r.n=n; r.on_device=1
    for i,k in enumerate(["x","y","z","t","dx","dy","dz","lambda_"]): setattr(r,k,inp[i].data_ptr())
    for i,k in enumerate(["ox","oy","oz","ot","odx","ody","odz"]): setattr(r,k,o[i].data_ptr())
    for i,k in enumerate(["status","last_node","npoints"]): setattr(r,k,io[i].data_ptr())
    op=H.opts(disable_fresnel=1, steps_per_launch=0, seed=5)
    for _ in range(3): R.check(R.rbg_trace(h,C.byref(op),C.byref(r),None))
    torch.cuda.synchronize(); t0=time.perf_counter(); reps=10
    for _ in range(reps): R.check(R.rbg_trace(h,C.byref(op),C.byref(r),None))
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/reps
    print("n=%8d  %.3f ms  %.3g rays/s  focused=%d"%(n,dt*1e3,n/dt,int((io[0]==3).sum())),flush=True)
