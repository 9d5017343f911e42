# throughput of every BASELINE config on one GPU (device-resident rays), quick survey
import sys, ctypes as C, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
import robast_b200 as R
from robast_b200 import configs
import helpers as H
dev=torch.device('cuda:0')
out={}
for cfg,theta,n,kw in ((1,0.0,1000*1000,{}),(1,1.0,3000*3000,{}),(2,1.0,3334*3334,{}),(3,0.0,3000*3000,{}),(3,3.0,3000*3000,{}),(4,0.0,10_000_000,{}),(5,0.0,10_000_000,dict(rings=2)),(5,20.0,10_000_000,dict(rings=10)),(5,20.0,10_000_000,dict(rings=10,precalc=True))):
    mgr,keep=configs.BUILDERS[cfg](**kw)
    ex=mgr.ExportScene()
    h=C.c_void_p(); R.check(R.rbg_scene_create(ex.desc_ptr(),0,C.byref(h)))
    nside=int(round(n**0.5)) if cfg<=3 else None
    p=configs.beam(cfg,theta,n_side=nside)
    if cfg==5: p=configs.beam(5,theta,n_side=(2*kw.get('rings',2)+1)*4.0)
    d=H.shoot_desc(p)
    inp=torch.empty((8,n),dtype=torch.float64,device=dev); o=torch.empty((7,n),dtype=torch.float64,device=dev); io=torch.empty((3,n),dtype=torch.int32,device=dev)
    R.check(R.rbg_shoot(C.byref(d),0,n,*[inp[i].data_ptr() for i in range(8)],0,None))
    r=R.rbg_rays(); r.n=n; r.on_device=1
    for i,k in enumerate(["x","y","z","t","dx","dy","dz","lambda_"]): setattr(r,k,inp[i].data_ptr())
    for i,k in enumerate(["ox","oy","oz","ot","odx","ody","odz"]): setattr(r,k,o[i].data_ptr())
    for i,k in enumerate(["status","last_node","npoints"]): setattr(r,k,io[i].data_ptr())
    res={}
    for spl in (0,-1):
        op=H.opts(disable_fresnel=1 if cfg==2 else 0, steps_per_launch=spl, seed=5)
        for _ in range(2): R.check(R.rbg_trace(h,C.byref(op),C.byref(r),None))
        torch.cuda.synchronize(); l0=R.rbg_launch_count(); t0=time.perf_counter()
        for _ in range(3): R.check(R.rbg_trace(h,C.byref(op),C.byref(r),None))
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
        res[spl]=(n/dt,(R.rbg_launch_count()-l0)//3)
    st=np.bincount(io[0].cpu().numpy(),minlength=6); npts=io[2].float().mean().item()
    print("cfg%d theta=%.1f %s n=%.2g variant=%s wavefront %.3g rays/s (%d launches) single %.3g rays/s  status=%s mean npoints=%.2f"%(cfg,theta,kw,n,R.rbg_scene_kernel_variant(h).decode(),res[0][0],res[0][1],res[-1][0],st.tolist(),npts), flush=True)
    R.rbg_scene_destroy(h)
