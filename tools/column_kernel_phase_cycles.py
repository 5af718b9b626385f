"""Cycle counts of the phases of k_cols_fused_p2_lr (clock64 marks of CTA 0); needs a library built with
GFMD_NVCC_EXTRA=-DGFMD_PHASE_TIMING sh user-gfmd_b200/csrc/build.sh."""
import sys, ctypes, numpy as np, torch
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(ROOT,'user-gfmd_b200')); sys.path.insert(0,ROOT)
import gfmd_b200
from gfmd_b200 import synthetic
nx=ny=4096; d=3
s=gfmd_b200.GFMDSolverB200(); s.set_grid_size(nx,ny,d)
for k0 in range(0,s.nky,256):
    nk=min(256,s.nky-k0); s.set_kernel_columns(synthetic.phi_columns(nx,ny,k0,nk),k0,normalized=False)
s.set_linf(np.zeros(1))
u=torch.rand((d,nx*ny),device='cuda',dtype=torch.float64)-0.5; f=torch.empty_like(u); torch.cuda.synchronize()
lib=s.lib; out=(ctypes.c_longlong*16)()
for i in range(3): s.post_force_device(u,f)
s.synchronize(); lib.gfmd_b200_debug_phase_cycles(out,1)
n=10
for i in range(n): s.post_force_device(u,f)
s.synchronize(); lib.gfmd_b200_debug_phase_cycles(out,1)
names=['pass0_fwd','groupA_rest_fwd','barrier1','groupB_first_fwd','contraction','groupB_first_inv','barrier2','groupA_rest_inv','pass0_inv']
cols_per_cta=(2049+147)//148
tot=sum(out[:9])
for i,nm in enumerate(names): print('%-18s %8.0f cycles/column  %5.1f%%'%(nm,out[i]/n/cols_per_cta,100*out[i]/tot))
print('total per column', tot/n/cols_per_cta)
