"""Multi-GPU parity worker.  Launch: python -m torch.distributed.run --nnodes=1
--nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py

Every rank owns an x-slab.  Checks the slab-decomposed CUDA path (NCCL all-to-all
transposes inside libgfmd_b200) against (a) the committed golden vectors of the
reference solver and (b) the single-GPU path on the same inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
import gfmd_b200  # noqa: E402
from gfmd_b200 import synthetic  # noqa: E402

TOL = 1e-11


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # GFMD_TEST_ONE_DEVICE=1: every rank (process) uses cuda:0 -- the whole slab path (CUDA IPC
    # mappings, peer pushes, flag words, in-kernel peer loads / stores) on a box with ONE GPU.
    # NCCL refuses two ranks on one device, so there is no communicator: the ranks order their
    # transfers through the flag words alone; the host-side plumbing (handle exchange) is gloo.
    one_device = os.environ.get("GFMD_TEST_ONE_DEVICE", "0") == "1"
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_device:
        dist.init_process_group("gloo")
        cdev = torch.device("cpu")
    else:
        dist.init_process_group("nccl", device_id=dev)
        cdev = dev
    s = None
    ok = True

    peer_copy = os.environ.get("GFMD_TEST_EXCHANGE", "ipc") == "ipc" or one_device
    use_nccl = not one_device and os.environ.get("GFMD_TEST_NO_COMM", "0") != "1"
    if rank == 0:
        print("exchange:", ("CUDA IPC peer copies" if peer_copy else "NCCL send/recv") +
              (", NCCL communicator present" if use_nccl else ", no NCCL communicator") +
              (", all ranks on cuda:0" if one_device else ""), flush=True)

    def new_slab(nx, ny, d):
        uid = None
        if use_nccl:
            # a fresh communicator per solver: rank 0 hands out a new id
            b = torch.zeros(gfmd_b200.UNIQUE_ID_BYTES, dtype=torch.uint8, device=cdev)
            if rank == 0:
                b.copy_(torch.frombuffer(bytearray(gfmd_b200.get_unique_id()), dtype=torch.uint8))
            dist.broadcast(b, 0)
            uid = bytes(b.cpu().numpy().tobytes())
        sl = gfmd_b200.GFMDSolverB200(device=local, rank=rank, nranks=world, unique_id=uid)
        sl.set_grid_size(nx, ny, d)
        if peer_copy:
            sl.enable_peer_copy(gfmd_b200.all_gather_bytes_fn(cdev, world))
        return sl

    # (a) golden vectors through the slab path (generic kernels)
    for name in ["C1_sc100_128x128", "small_fcc111_8x7", "C2_fcc111_64x37"]:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        nx, ny, d = int(z["nx"]), int(z["ny"]), int(z["ndof"])
        if nx % world:
            continue
        s = new_slab(nx, ny, d)
        s.set_kernel(z["phi"], z["linf"])            # each rank extracts its q columns
        nxl = nx // world
        for c in ("uniform", "hertz"):
            u = np.ascontiguousarray(z["u_" + c][:, rank * nxl:(rank + 1) * nxl, :]).reshape(d, nxl * ny)
            f = np.zeros_like(u)
            e = s.post_force(u, f)
            fref = z["f_" + c][:, rank * nxl:(rank + 1) * nxl, :].reshape(d, nxl * ny)
            err = np.abs(f - fref).max() / np.abs(z["f_" + c]).max()
            et = torch.tensor([e], device=cdev, dtype=torch.float64)
            dist.all_reduce(et)                       # the fix sums the per-rank energies
            eref = float(z["epot_" + c])
            eerr = abs(et.item() - eref) / abs(eref)
            u0err = np.abs(s.get_u0() - z["u0_" + c]).max() / max(1.0, np.abs(z["u0_" + c]).max())
            good = err < TOL and eerr < TOL and u0err < TOL
            ok = ok and good
            if rank == 0 or not good:
                print("rank %d %s/%s slab force err %.2e epot err %.2e u0 err %.2e %s"
                      % (rank, name, c, err, eerr, u0err, "ok" if good else "FAIL"), flush=True)
        s.close()

    # (b) large grid, specialised kernels: slab result == single-GPU result
    for nx, ny in [(4096, 2048), (2048, 4096), (4096, 96), (8192, 4096), (16384, 256)]:
        d = 3
        s = new_slab(nx, ny, d)
        for k0 in range(s.kylo, s.kylo + s.nky, 128):
            nk = min(128, s.kylo + s.nky - k0)
            s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        s.set_linf(np.array([0.25]))
        one = gfmd_b200.GFMDSolverB200(device=local)
        one.set_grid_size(nx, ny, d)
        for k0 in range(0, one.nky, 128):
            nk = min(128, one.nky - k0)
            one.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        one.set_linf(np.array([0.25]))
        gen = torch.Generator(device=dev).manual_seed(11)          # same field on every rank
        ufull = torch.rand((d, nx, ny), generator=gen, device=dev, dtype=torch.float64) - 0.5
        ffull = torch.empty_like(ufull)
        nxl = nx // world
        uslab = ufull[:, rank * nxl:(rank + 1) * nxl, :].contiguous()
        fslab = torch.empty_like(uslab)
        torch.cuda.synchronize()
        one.post_force_device(ufull, ffull)
        r1 = one.results()
        for rep in range(2):
            s.post_force_device(uslab, fslab)
            rs = s.results()
        err = (fslab - ffull[:, rank * nxl:(rank + 1) * nxl, :]).abs().max().item() / ffull.abs().max().item()
        et = torch.tensor([rs["epot"]], device=cdev, dtype=torch.float64)
        dist.all_reduce(et)
        eerr = abs(et.item() - r1["epot"]) / abs(r1["epot"])
        u0err = np.abs(rs["u0"] - r1["u0"]).max() / np.abs(r1["u0"]).max()
        good = err < TOL and eerr < TOL and u0err < TOL
        ok = ok and good
        if rank == 0 or not good:
            print("rank %d %dx%d [%s] slab vs single force err %.2e epot err %.2e u0 err %.2e %s"
                  % (rank, nx, ny, " |".join(part[:46] for part in s.describe().split("|")[1:]), err, eerr, u0err,
                     "ok" if good else "FAIL"), flush=True)
        s.close()
        one.close()

    flag = torch.tensor([1 if ok else 0], device=cdev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_PARITY_OK" if flag.item() == 1 else "MGPU_PARITY_FAIL", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
