"""Independent CPU value of the potential energy of bench.py's default workload (4096 x 4096, `sc100 height
128`, displacement_field(seed=1, nwaves=8)): numpy rfft2 + the transfer-matrix recursion with np.linalg.solve.
Takes ~16 min on 8 cores; the result, 442.2815166173698, is what bench.py compares its `epot` with
(`epot_rel_err`).  No GPU, no library of this repository involved except the synthetic input generators."""
import sys, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'user-gfmd_b200'))
import numpy as np
from gfmd_b200 import synthetic
nx = ny = 4096; d = 3
t0 = time.time()
u = synthetic.displacement_field(nx, ny, 0, nx, seed=1, nwaves=8)
x, xeq, gid, mask = synthetic.atoms_for_slab(nx, ny, 0, nx, u)
# gather as the fix does: u = x - xeq (minimum image irrelevant at 1e-3 amplitudes)
ug = (x - xeq).T.reshape(3, nx, ny)
uq = np.fft.rfft2(ug, axes=(1, 2))            # [3, nx, ny/2+1]
print("fft done", time.time() - t0, flush=True)
nyh = ny // 2 + 1
esum = 0.0
for k0 in range(0, nyh, 128):
    nk = min(128, nyh - k0)
    M = synthetic.sc100_dynamical_matrices(nx, ny, k0, nk)
    U0, U, V = M[:, :, 0], M[:, :, 1], M[:, :, 2]
    Vd = -np.conj(np.swapaxes(V, -1, -2))
    VT = None
    for it in range(128):
        A = U if VT is None else U + VT
        VT = V @ np.linalg.solve(A, Vd)
    phi = (U0 + VT) / (nx * ny)
    phi = 0.5 * (phi + np.conj(np.swapaxes(phi, -1, -2)))
    q = np.moveaxis(uq[:, :, k0:k0 + nk], 0, -1)      # [nx, nk, 3]
    F = np.einsum('xkij,xkj->xki', phi, q)
    eq = (np.conj(q) * F).sum(axis=-1).real           # [nx, nk]
    w = np.where((np.arange(k0, k0 + nk) == 0) | (2 * np.arange(k0, k0 + nk) == ny), 1.0, 2.0)
    esum += float((eq * w[None, :]).sum())
    print(k0, time.time() - t0, flush=True)
print("epot", 0.5 * esum)
