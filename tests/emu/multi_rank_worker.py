"""TEST INFRASTRUCTURE ONLY -- the slab-decomposed (multi-GPU) path on the CUDA-on-CPU emulation
build: the ranks are host threads of this process, the peer pushes are memcpys between their
buffers (in-process "IPC" handles), NCCL is tests/emu/fake_nccl.cpp whose collectives are real
barriers.  Mirrors tests/mgpu_worker.py (which needs real GPUs): golden vectors through the slab
path, and slab == single-rank on grids that select the specialised kernels, the chunked pipeline,
split columns (LP > 0) and long columns.

    python tests/emu/multi_rank_worker.py <nranks> [nccl]      ("nccl": send/recv instead of peer pushes)
"""
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
sys.path.insert(0, HERE)
import build as emu_build  # noqa: E402
import gfmd_b200  # noqa: E402
from gfmd_b200 import synthetic  # noqa: E402

TOL = 1e-11


class Gather:
    """all_gather / broadcast / sum between the rank threads."""

    def __init__(self, n):
        self.n, self.slots, self.bar = n, [None] * n, threading.Barrier(n, timeout=300)

    def all_gather(self, rank, value):
        self.slots[rank] = value
        self.bar.wait()
        out = list(self.slots)
        self.bar.wait()
        return out


def run_rank(rank, world, peer_copy, comm, results):
    ok = True
    log = []

    def new_slab(nx, ny, d):
        uid = comm.all_gather(rank, gfmd_b200.get_unique_id() if rank == 0 else None)[0]
        sl = gfmd_b200.GFMDSolverB200(device=0, rank=rank, nranks=world, unique_id=uid)
        sl.set_grid_size(nx, ny, d)
        if peer_copy:
            sl.enable_peer_copy(lambda b: comm.all_gather(rank, b))
        return sl

    # (a) golden vectors through the slab path (generic kernels)
    for name in ["small_sc100_16x12", "small_fcc111_8x7", "C2_fcc111_64x37"]:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        nx, ny, d = int(z["nx"]), int(z["ny"]), int(z["ndof"])
        if nx % world:
            continue
        s = new_slab(nx, ny, d)
        s.set_kernel(z["phi"], z["linf"])
        nxl = nx // world
        for c in ("uniform", "hertz"):
            u = np.ascontiguousarray(z["u_" + c][:, rank * nxl:(rank + 1) * nxl, :]).reshape(d, nxl * ny)
            f = np.full_like(u, np.nan)
            e = s.post_force(u, f)
            fref = z["f_" + c][:, rank * nxl:(rank + 1) * nxl, :].reshape(d, nxl * ny)
            err = np.abs(f - fref).max() / np.abs(z["f_" + c]).max()
            etot = sum(comm.all_gather(rank, e))              # the fix sums the per-rank energies
            eref = float(z["epot_" + c])
            eerr = abs(etot - eref) / abs(eref)
            u0err = np.abs(s.get_u0() - z["u0_" + c]).max() / max(1.0, np.abs(z["u0_" + c]).max())
            good = bool(err < TOL and eerr < TOL and u0err < TOL)
            ok = ok and good
            log.append("rank %d %s/%s slab force err %.2e epot err %.2e u0 err %.2e %s"
                       % (rank, name, c, err, eerr, u0err, "ok" if good else "FAIL"))
        s.close()

    # (b) specialised kernels: slab result == single-rank result (thin grids keep the emulation cheap)
    grids = [(4096, 4), (2048, 6), (8, 4096), (8192, 4), (16384, 2), (16, 2048)]
    if os.environ.get("GFMD_EMU_GRIDS"):          # e.g. "4096x2048": the chunked pipeline (minutes)
        grids = [tuple(int(v) for v in g.split("x")) for g in os.environ["GFMD_EMU_GRIDS"].split(",")]
    for nx, ny in grids:
        if nx % world:
            continue
        d = 3
        s = new_slab(nx, ny, d)
        for k0 in range(s.kylo, s.kylo + s.nky, 128):
            nk = min(128, s.kylo + s.nky - k0)
            s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        s.set_linf(np.array([0.25]))
        ufull = np.random.default_rng(11).uniform(-0.5, 0.5, size=(d, nx, ny))
        single = comm.all_gather(rank, None)                  # placeholder round keeps ranks in step
        if rank == 0:
            one = gfmd_b200.GFMDSolverB200(device=0)
            one.set_grid_size(nx, ny, d)
            for k0 in range(0, one.nky, 128):
                nk = min(128, one.nky - k0)
                one.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
            one.set_linf(np.array([0.25]))
            ffull = np.full((d, nx * ny), np.nan)
            one.post_force_device(np.ascontiguousarray(ufull.reshape(d, nx * ny)), ffull)
            r1 = one.results()
            one.close()
            single = (ffull.reshape(d, nx, ny), r1["epot"], r1["u0"])
        single = comm.all_gather(rank, single)[0]
        ffull, e1, u01 = single
        nxl = nx // world
        uslab = np.ascontiguousarray(ufull[:, rank * nxl:(rank + 1) * nxl, :]).reshape(d, nxl * ny)
        fslab = np.full_like(uslab, np.nan)
        for rep in range(2):
            s.post_force_device(uslab, fslab)
            rs = s.results()
        err = np.abs(fslab.reshape(d, nxl, ny) - ffull[:, rank * nxl:(rank + 1) * nxl, :]).max() / np.abs(ffull).max()
        etot = sum(comm.all_gather(rank, rs["epot"]))
        eerr = abs(etot - e1) / abs(e1)
        u0err = np.abs(rs["u0"] - u01).max() / np.abs(u01).max()
        good = bool(err < TOL and eerr < TOL and u0err < TOL)
        ok = ok and good
        log.append("rank %d %dx%d [%s] slab vs single force err %.2e epot err %.2e u0 err %.2e %s"
                   % (rank, nx, ny, " |".join(part[:46] for part in s.describe().split("|")[1:]), err, eerr, u0err, "ok" if good else "FAIL"))
        s.close()
    results[rank] = (ok, log)


def main():
    world = int(sys.argv[1])
    peer_copy = not (len(sys.argv) > 2 and sys.argv[2] == "nccl")
    os.environ["GFMD_B200_NCCL_LIB"] = emu_build.FAKE_NCCL
    gfmd_b200._lib = gfmd_b200.load_library(emu_build.build())
    comm = Gather(world)
    results = [None] * world
    threads = [threading.Thread(target=run_rank, args=(r, world, peer_copy, comm, results)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    good = all(r is not None and r[0] for r in results)
    for r in results:
        if r is not None and (not r[0] or r is results[0]):
            print("\n".join(r[1]))
    print("exchange:", "peer pushes (in-process handles) + barrier" if peer_copy else "send/recv")
    print("EMU_MGPU_PARITY_OK" if good else "EMU_MGPU_PARITY_FAIL", flush=True)
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
