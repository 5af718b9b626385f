"""A/B of the experimental row-kernel variants (GFMD_B200_ROWS_VARIANT, csrc/kernels_fast.cuh) on one GPU:
per-stage CUDA-event times of the solver step and the difference of the forces to the default variant.
  python tools/rows_variants_ab.py [nx [ny]] > gpurun_out/rows_variants.txt"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'user-gfmd_b200')); sys.path.insert(0, ROOT)
import gfmd_b200
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = int(sys.argv[2]) if len(sys.argv) > 2 else nx
d = 3
variants = [0] + [ny + k for k in ((0, 1, 3, 5, 6, 7, 8) if ny == 4096 else (3, 5, 6, 7, 8) if ny == 8192 else (5, 8, 9))]   # 0 = the library default
if os.environ.get('AB_ONLY'): variants = [0] + [int(v) for v in os.environ['AB_ONLY'].split(',')]
rounds = int(os.environ.get('AB_ROUNDS', '2'))
u = torch.rand((d, nx * ny), device='cuda', dtype=torch.float64) - 0.5
f0 = None
for rep in range(rounds):                      # second round: order effects / clocks
    for v in variants:
        if v: os.environ['GFMD_B200_ROWS_VARIANT'] = str(v)
        else: os.environ.pop('GFMD_B200_ROWS_VARIANT', None)
        s = gfmd_b200.GFMDSolverB200(); s.set_grid_size(nx, ny, d)
        for k0 in range(0, s.nky, 256):
            nk = min(256, s.nky - k0)
            P = np.zeros((nx, nk, d, d), dtype=np.complex128); P[..., range(d), range(d)] = 1.0 + 0.001 * k0
            s.set_kernel_columns(P, k0, normalized=False)
        s.set_linf(np.zeros(1))
        f = torch.empty_like(u); torch.cuda.synchronize()
        for i in range(5): s.post_force_device(u, f)
        s.synchronize()
        if f0 is None: f0 = f.clone()
        diff = (f - f0).abs().max().item()
        s.profile(True)
        for i in range(30): s.post_force_device(u, f)
        s.profile(False)
        st = s.stage_times()
        ms = {k: v[0] / max(v[1], 1) for k, v in st.items() if v[1]}
        print('variant %5d round %d | rows_fwd %.4f cols %.4f rows_inv %.4f | solver %.4f ms | max|f - f_default| %.3e | %s'
              % (v, rep, ms['rows_fwd'], ms['cols_fused'], ms['rows_inv'], sum(ms.values()), diff,
                 s.describe().split('|')[1].strip()), flush=True)
        s.close()
