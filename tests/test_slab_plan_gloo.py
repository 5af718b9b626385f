"""CPU test of the multi-rank plan (world_size 2 and 3, gloo): the slab partition, the
staging-buffer layout, the all-to-all block structure, the ky weights, the gamma-point
owner, the per-rank energies and the u0 all-reduce -- with the oracle's slab stages
(oracle/gfmd_oracle.py rows_forward / columns_contract / rows_inverse) standing in for
the CUDA kernels.  The exchange itself is the same block all-to-all the library issues
through NCCL (csrc/gfmd_b200.cu: exchange)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _all_to_all(blocks, rank, world):
    """blocks[p]: tensor for rank p -> list of tensors received from each rank (gloo)."""
    out = [torch.empty_like(blocks[0]) for _ in range(world)]
    out[rank].copy_(blocks[rank])
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        reqs.append(dist.isend(blocks[p], p))
        reqs.append(dist.irecv(out[p], p))
    for r in reqs:
        r.wait()
    return out


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gfmd_oracle as O
    import gfmd_b200
    g = load_golden(name)
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    plan = gfmd_b200.slab_plan(nx, ny, d, rank, world)
    nxl, x0, kyb, ky0, nky = plan["nx_loc"], plan["x0"], plan["kyb"], plan["ky0"], plan["nky_loc"]
    nyh = plan["nyh"]
    u = g["u_uniform"][:, x0:x0 + nxl, :]
    linf = g["linf"]
    # rows -> send staging [P][d][kyb][nx_loc]
    ut = O.rows_forward(u)                                   # [d, nyh, nx_loc]
    send = np.zeros((world, d, kyb, nxl), dtype=np.complex128)
    for p in range(world):
        k_lo, k_hi = p * kyb, min(nyh, (p + 1) * kyb)
        if k_hi > k_lo:
            send[p, :, :k_hi - k_lo, :] = ut[:, k_lo:k_hi, :]
    recv = _all_to_all([torch.from_numpy(send[p].copy()) for p in range(world)], rank, world)
    # columns of this rank: piece p of every column comes from rank p
    cols = np.concatenate([recv[p].numpy()[:, :nky, :] for p in range(world)], axis=2)   # [d, nky, nx]
    P4 = g["phi"].reshape(nx, ny, d, d)
    phi_cols = np.ascontiguousarray(np.swapaxes(P4[:, ky0:ky0 + nky], 0, 1))              # [nky, nx, d, d]
    ft, e2, u0 = O.columns_contract(cols, phi_cols, linf, ky0, np.array(plan["ky_weights"]))
    epot = 0.5 * e2                                         # this rank's energy (the fix sums them)
    assert (u0 is not None) == (rank == plan["gamma_rank"])
    u0t = torch.from_numpy(u0 if u0 is not None else np.zeros(d))
    dist.all_reduce(u0t)                                    # gfmd_solver_static.cpp:176
    # way back: block p = rows of rank p
    send2 = np.zeros((world, d, kyb, nxl), dtype=np.complex128)
    for p in range(world):
        send2[p, :, :nky, :] = ft[:, :, p * nxl:(p + 1) * nxl]
    recv2 = _all_to_all([torch.from_numpy(send2[p].copy()) for p in range(world)], rank, world)
    fslab_t = np.zeros((d, nyh, nxl), dtype=np.complex128)
    for p in range(world):
        k_lo, k_hi = p * kyb, min(nyh, (p + 1) * kyb)
        if k_hi > k_lo:
            fslab_t[:, k_lo:k_hi, :] = recv2[p].numpy()[:, :k_hi - k_lo, :]
    f = O.rows_inverse(fslab_t, ny)
    et = torch.tensor([epot], dtype=torch.float64)
    dist.all_reduce(et)                                     # fix_gfmd.cpp:935
    fref = g["f_uniform"][:, x0:x0 + nxl, :]
    err = np.abs(f - fref).max() / np.abs(g["f_uniform"]).max()
    eerr = abs(et.item() - float(g["epot_uniform"])) / abs(float(g["epot_uniform"]))
    u0err = np.abs(u0t.numpy() - g["u0_uniform"]).max()
    q.put((rank, err, eerr, u0err))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "C1_sc100_128x128"), (2, "small_fcc111_8x7"),
                                        (2, "C3_fcc100_two_layers_10x10"), (4, "C2_fcc111_64x37")])
def test_slab_plan_with_gloo(world, name, oracle_libs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (hash((world, name)) % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, eerr, u0err in res:
        assert err < 1e-13 and eerr < 1e-13 and u0err < 1e-12, (rank, err, eerr, u0err)


def test_slab_plan_properties():
    sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
    import gfmd_b200
    for nx, ny, P in [(4096, 4096, 8), (16384, 8192, 8), (8, 7, 2), (12, 37, 3), (4, 2, 4)]:
        plans = [gfmd_b200.slab_plan(nx, ny, 3, r, P) for r in range(P)]
        assert sum(p["nx_loc"] for p in plans) == nx
        assert sum(p["nky_loc"] for p in plans) == ny // 2 + 1          # every ky owned exactly once
        assert sum(sum(p["ky_weights"]) for p in plans) == ny            # full-spectrum multiplicity
        assert all(p["block_elems"] == plans[0]["block_elems"] for p in plans)   # uniform all-to-all
    with pytest.raises(gfmd_b200.GFMDError):
        gfmd_b200.slab_plan(10, 8, 3, 0, 4)
