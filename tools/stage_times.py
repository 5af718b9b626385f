"""Per-stage CUDA-event times of the solver step on one GPU: python tools/stage_times.py [nx [ny [ndof]]]
(ndof > 3 uses a block-diagonal table of the ndof-3 synthetic stiffness; only the timing matters)."""
import sys, numpy as np, torch
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(ROOT,'user-gfmd_b200')); sys.path.insert(0,ROOT)
import gfmd_b200
from gfmd_b200 import synthetic
nx=int(sys.argv[1]) if len(sys.argv)>1 else 4096
ny=int(sys.argv[2]) if len(sys.argv)>2 else nx
d=int(sys.argv[3]) if len(sys.argv)>3 else 3
s=gfmd_b200.GFMDSolverB200(); s.set_grid_size(nx,ny,d)
print(s.describe())
for k0 in range(0,s.nky,256):
    nk=min(256,s.nky-k0); P3=synthetic.phi_columns(nx,ny,k0,nk)
    if d==3: P=P3
    else:
        P=np.zeros((nx,nk,d,d),dtype=np.complex128)
        for a in range(d//3): P[:,:,3*a:3*a+3,3*a:3*a+3]=P3
    s.set_kernel_columns(P,k0,normalized=False)
s.set_linf(np.zeros(d//3))
u=torch.rand((d,nx*ny),device='cuda',dtype=torch.float64)-0.5; f=torch.empty_like(u); torch.cuda.synchronize()
for i in range(5): s.post_force_device(u,f)
s.synchronize()
import time
t0=time.perf_counter(); N=50
for i in range(N): s.post_force_device(u,f)
s.synchronize(); dt=(time.perf_counter()-t0)/N
s.profile(True)
for i in range(20): s.post_force_device(u,f)
s.profile(False)
st=s.stage_times()
print('solver ms %.4f  steps/s %.1f'%(dt*1e3,1/dt), {k:round(v[0]/max(v[1],1),4) for k,v in st.items() if v[1]})
