"""Per-stage times of the slab-decomposed step, A/B over the opt-in settings in ONE launch (a
multi-GPU box is charged N x its time, so every setting rides on the same rendezvous):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/stage_times_multi_gpu.py

Settings (AB=<comma separated names> restricts them, default all that apply to the grid):
    default        peer pushes on copy engines + chunked pipeline
    peer_store     GFMD_B200_PEER_STORE=1: the column stage stores its results straight into the peers'
                   return buffers (kernel_cols_lr.cuh, PEER variants)
    direct         GFMD_B200_PEER_DIRECT=1: no transposes, the column stage also LOADS its pieces from the peers
    rows_fused     GFMD_B200_ROWS_VARIANT=ny+6: fused backward row kernel
    rows_r16       GFMD_B200_ROWS_VARIANT=ny+8: radix-16 row kernels
    r16_peer       rows_r16 + peer_store
    r16_direct     rows_r16 + direct
    chunks8        GFMD_B200_CHUNKS=8
    nccl           grouped ncclSend/ncclRecv instead of the peer pushes
Prints, per setting, the solver step (wall clock over 30 steps between barriers), the CUDA-event stage
times of rank 0 and max|f - f_default| over this rank's slab."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'user-gfmd_b200')); sys.path.insert(0, ROOT)
import gfmd_b200
from gfmd_b200 import synthetic

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
grids = {1: (4096, 4096), 2: (4096, 8192), 4: (8192, 8192), 8: (16384, 8192)}
nx, ny = grids[world]; d = 3
if os.environ.get('GRID'): nx, ny = (int(v) for v in os.environ['GRID'].split('x'))      # e.g. GRID=16384x16384
SETTINGS = {
    'default': {},
    'peer_store': {'GFMD_B200_PEER_STORE': '1'},
    'direct': {'GFMD_B200_PEER_DIRECT': '1'},
    'rows_fused': {'GFMD_B200_ROWS_VARIANT': str(ny + 6)},
    'rows_r16': {'GFMD_B200_ROWS_VARIANT': str(ny + 8)},
    'r16_peer': {'GFMD_B200_ROWS_VARIANT': str(ny + 8), 'GFMD_B200_PEER_STORE': '1'},
    'r16_direct': {'GFMD_B200_ROWS_VARIANT': str(ny + 8), 'GFMD_B200_PEER_DIRECT': '1'},
    'chunks8': {'GFMD_B200_CHUNKS': '8'},
    'nccl': {'EXCH': 'nccl'},
    'sync_nccl': {'GFMD_B200_SYNC': 'nccl'},                       # round-1 ordering: all-reduce barriers
    'direct_sync_nccl': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_SYNC': 'nccl'},
    'chunks2': {'GFMD_B200_CHUNKS': '2'},
    'chunks1': {'GFMD_B200_CHUNKS': '1'},
    'rows_r8': {'GFMD_B200_ROWS_VARIANT': str(ny)},                # the round-1 row kernels
    'direct_seq': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_CHUNKS': '1'},      # no transposes, not overlapped
    'direct_x8': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '8'},
    'direct_x12': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '12'},
    'direct_x24': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '24'},
    'direct_32_16': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '32,16'},
    'direct_40_16': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '40,16'},
    'direct_48_16': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '48,16'},
    'direct_40_24': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '40,24'},
    'direct_40_12': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '40,12'},
    'direct_56_16': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '56,16'},
    'direct_48_24': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '48,24'},
    'direct_40_32': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '40,32'},
    'direct_24_40': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '24,40'},
    'direct_32_32': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '32,32'},
    'direct_64_16': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '64,16'},
    'direct_40_16_tl': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '40,16', 'GFMD_B200_TIMELINE': '1'},
    'direct_tl': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_TIMELINE': '1'},
    'direct_x32': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '32'},
    'direct_x24c4': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_XCHG_SMS': '24', 'GFMD_B200_CHUNKS': '4'},
    'direct_c4': {'GFMD_B200_PEER_DIRECT': '1', 'GFMD_B200_CHUNKS': '4'},
    'pushes': {'GFMD_B200_PEER_DIRECT': '0'},                      # copy-engine pushes, chunked pipeline
}
names = os.environ['AB'].split(',') if os.environ.get('AB') else list(SETTINGS)
if world == 1: names = [n for n in names if n in ('default', 'rows_fused', 'rows_r16')]
KEYS = ('GFMD_B200_PEER_STORE', 'GFMD_B200_PEER_DIRECT', 'GFMD_B200_ROWS_VARIANT', 'GFMD_B200_CHUNKS', 'EXCH', 'GFMD_B200_SYNC',
        'GFMD_B200_XCHG_SMS', 'GFMD_B200_TIMELINE')
nxl = nx // world
u = torch.rand((d, nxl * ny), device=dev, dtype=torch.float64, generator=torch.Generator(dev).manual_seed(7 + rank)) - 0.5
f0 = None
for name in names:
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(SETTINGS[name])
    b = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0: b.copy_(torch.frombuffer(bytearray(gfmd_b200.get_unique_id()), dtype=torch.uint8))
    dist.broadcast(b, 0)
    s = gfmd_b200.GFMDSolverB200(device=local, rank=rank, nranks=world, unique_id=bytes(b.cpu().numpy().tobytes()))
    s.set_grid_size(nx, ny, d)
    if world > 1 and os.environ.get('EXCH', 'ipc') == 'ipc': s.enable_peer_copy(gfmd_b200.all_gather_bytes_fn(dev, world))
    # timing only: the sc100 matrices evaluated on the GPU, two layers (one recursion step)
    for k0 in range(s.kylo, s.kylo + s.nky, 64):
        nk = min(64, s.kylo + s.nky - k0)
        s.build_kernel_columns_device(synthetic.sc100_dynamical_matrices_torch(nx, ny, k0, nk, dev), k0, nk, height=2)
    s.set_linf(np.zeros(1))
    f = torch.empty_like(u); torch.cuda.synchronize()
    for i in range(5): s.post_force_device(u, f)
    s.synchronize(); dist.barrier()
    if f0 is None: f0 = f.clone()
    diff = (f - f0).abs().max().item() / f0.abs().max().item()
    t0 = time.perf_counter(); N = 30
    for i in range(N): s.post_force_device(u, f)
    s.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / N
    s.profile(True)
    for i in range(10): s.post_force_device(u, f)
    s.profile(False)
    st = s.stage_times()
    if rank == 0:
        print('%-10s P=%d %dx%d solver ms %.4f steps/s %.1f | %s | rel max|f - f_default| %.2e | %s'
              % (name, world, nx, ny, dt * 1e3, 1 / dt, {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]},
                 diff, s.describe().split('|', 1)[1].strip()), flush=True)
    s.close()
dist.barrier(); dist.destroy_process_group()
