"""Per-stage times of the slab-decomposed step: torchrun --nproc-per-node N tools/stage_times_multi_gpu.py
(EXCH=nccl selects the NCCL send/recv exchange, GFMD_B200_CHUNKS=n the pipeline depth)."""
import os, sys, numpy as np, torch, torch.distributed as dist
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(ROOT,'user-gfmd_b200')); sys.path.insert(0,ROOT)
import gfmd_b200
from gfmd_b200 import synthetic
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev=torch.device('cuda',local)
dist.init_process_group('nccl', device_id=dev)
b=torch.zeros(128,dtype=torch.uint8,device=dev)
if rank==0: b.copy_(torch.frombuffer(bytearray(gfmd_b200.get_unique_id()),dtype=torch.uint8))
dist.broadcast(b,0)
grids={1:(4096,4096),2:(4096,8192),4:(8192,8192),8:(16384,8192)}
nx,ny=grids[world]; d=3
s=gfmd_b200.GFMDSolverB200(device=local,rank=rank,nranks=world,unique_id=bytes(b.cpu().numpy().tobytes()))
s.set_grid_size(nx,ny,d)
if os.environ.get('EXCH','ipc')=='ipc': s.enable_peer_copy(gfmd_b200.all_gather_bytes_fn(dev,world))
for k0 in range(s.kylo,s.kylo+s.nky,128):
    nk=min(128,s.kylo+s.nky-k0); s.set_kernel_columns(synthetic.phi_columns(nx,ny,k0,nk),k0,normalized=False)
s.set_linf(np.zeros(1))
nxl=nx//world
u=torch.rand((d,nxl*ny),device=dev,dtype=torch.float64)-0.5; f=torch.empty_like(u); torch.cuda.synchronize()
for i in range(5): s.post_force_device(u,f)
s.synchronize(); dist.barrier()
import time
t0=time.perf_counter(); N=30
for i in range(N): s.post_force_device(u,f)
s.synchronize(); dist.barrier(); dt=(time.perf_counter()-t0)/N
s.profile(True)
for i in range(10): s.post_force_device(u,f)
s.profile(False)
st=s.stage_times()
if rank==0: print('P=%d %dx%d solver ms %.4f steps/s %.1f'%(world,nx,ny,dt*1e3,1/dt), {k:round(v[0]/max(v[1],1),4) for k,v in st.items() if v[1]}, flush=True)
s.close(); dist.destroy_process_group()
