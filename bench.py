#!/usr/bin/env python
"""GFMD force-evaluation throughput benchmark (BASELINE.json metric: GFMD force-eval steps/s vs
surface grid at 1/2/4/8 B200; % of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU solver

A "step" is one elastic-force evaluation of `fix gfmd` on a synthetic sc100 surface (ndof 3,
stiffness kernel `sc100 height 128`): gather (atoms -> grid), forward 2-D FFT, Phi(q).u(q)
contraction with energy and gamma point, inverse FFT, scatter (grid -> atoms).

Workload.  The SAME surface at every N: 16384 x 16384 (the north-star's strong-scaling grid; it
fits one B200), x-slabs over the N GPUs -> "scaling": "strong".  At N = 1 the line also carries
`grid_4096`: the 4096 x 4096 surface the north-star quotes the single-GPU roofline target on, with
its own roofline, e2e, cpu_baseline and the energy check against an independent CPU value.

  value     steps/s with atoms and grids resident in HBM (gfmd_b200_full_step)
  e2e       steps/s through the solver-plugin boundary GFMDSolver::post_force with HOST u_xy / f_xy:
            H2D of u, the GPU step, D2H of f and of (epot, u0) inside the timed region
  roofline  the kernel with the largest share of the step: algorithmic bytes over its
            CUDA-event time; `stages` lists every stage the same way
  parity_max_rel_err  (N > 1) slab-decomposed forces against a single-GPU run of the same surface
            on rank 0 of the same job; (N = 1) the size-independent identity E = -1/2 sum f.u
  cpu_baseline / --impl reference  the reference's own solver sources (oracle/_ref; FFTW replaced
            by oracle/fft_plain.c) on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
sys.path.insert(0, ROOT)

METRIC = "gfmd_force_eval_steps_per_sec"
UNIT = "steps/s"
STRONG_GRID = 16384
EPOT_4096 = 442.2815166173698      # tools/epot_reference_4096.py (numpy rfft2 + np.linalg.solve, ~16 min of CPU)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="nx = ny (default %d)" % STRONG_GRID)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-4096", action="store_true", help="skip the grid_4096 sub-record at N = 1")
    ap.add_argument("--no-parity", action="store_true", help="skip the single-GPU comparison at N > 1")
    return ap.parse_args()


def workload(args):
    n = args.grid if args.grid > 0 else STRONG_GRID
    return n, n, 3


def workload_string(nx, ny, d):
    return ("synthetic sc100 surface %dx%d, stiffness kernel `sc100 height 128`, ndof %d, 1 atom/cell"
            % (nx, ny, d))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ clocks ---

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------- reference ---

REF_SAMPLE_GRID = 4096       # the largest surface the reference arm runs in full (2.4 GB table, ~1.3 s / step)


def host_threads():
    """All host cores, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its children."""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return cores


def reference_run(nx, ny, d, steps, warmup):
    """Times the reference's GFMDSolverStatic::post_force (its own sources, FFT3d shim backed by
    oracle/fft_plain.c with OpenMP) on host arrays, on all host cores.  Surfaces up to 4096^2 run in
    full; a larger one is SAMPLED by the full 4096^2 surface and the rate scaled by the cell ratio
    (flagged `extrapolated`): the reference keeps the full complex d x d table in host memory,
    38.7 GB at 16384^2 and 45 minutes of single-threaded fill_phi_buffer."""
    cores = host_threads()
    from oracle import gfmd_oracle as O
    from gfmd_b200 import synthetic
    kind = "reference" if O.ref_available() else "port"
    cores = O.set_fft_threads(cores) or cores     # explicit: independent of the launcher's environment
    nxs, nys = min(nx, REF_SAMPLE_GRID), min(ny, REF_SAMPLE_GRID)
    phi = synthetic.phi_full(nxs, nys)
    u = synthetic.displacement_field(nxs, nys, seed=1, nwaves=8)
    linf = np.zeros(1)
    if kind == "reference":
        s = O.RefSolver(nxs, nys, d, fft_backend=1)
        s.set_phi(phi, linf)
        fn = lambda: s.post_force(u)
    else:
        fn = lambda: O.c_post_force(u, phi, linf, 1)
    for _ in range(max(warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    scale = (nxs * nys) / float(nx * ny)
    sample = ("GFMDSolverStatic::post_force on host u_xy/f_xy, %dx%d ndof %d in full, %d timed steps, %d warm-up"
              % (nxs, nys, d, steps, warmup))
    if scale != 1.0:
        sample += ("; the %dx%d workload is SAMPLED by this surface and the rate scaled by the cell ratio %.4g "
                   "(extrapolated)" % (nx, ny, scale))
    return {"value": scale / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": sample + "; FFT = oracle/fft_plain.c, OpenMP (no FFTW / MPI / LAMMPS on this box)",
            "ms_per_step": 1e3 * dt / scale, "extrapolated": scale != 1.0,
            "measured": {"grid": "%dx%d" % (nxs, nys), "value": 1.0 / dt, "ms_per_step": 1e3 * dt},
            "fft": "oracle/fft_plain.c (not FFTW)"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, d = workload(args)
    # bounded: the driver's K and W are honoured up to what a few minutes allow (1.3 s / step)
    steps = min(args.steps, 20)
    warmup = min(args.warmup, 3)
    r = reference_run(nx, ny, d, steps, warmup)
    cb = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "fft")}
    out = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
           "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_string(nx, ny, d),
                      "step": "GFMDSolver::post_force(u_xy, f_xy) on host arrays (solver only: no gather / scatter)",
                      "steps_requested": args.steps, "warmup_requested": args.warmup},
           "extrapolated": r["extrapolated"],
           "reference_fft": "oracle/fft_plain.c, not FFTW: a real FFTW build would be faster",
           "cpu_baseline": cb,
           "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "grid_4096": {"workload": workload_string(REF_SAMPLE_GRID, REF_SAMPLE_GRID, d),
                         "value": r["measured"]["value"], "unit": UNIT, "ms_per_step": r["measured"]["ms_per_step"],
                         "extrapolated": False,
                         "e2e": {"value": r["measured"]["value"], "unit": UNIT}},
           "gpu_launches": 0}
    print(json.dumps(out))


# -------------------------------------------------------------------- b200 ---

def ncu_traffic(stage, nx, ny, d, world, kernels):
    """DRAM bytes per launch of a stage's kernel from the committed ncu --set full capture of the SAME kernel on
    the SAME grid (profiles/r2_ncu_full_summary.json, refreshed by tools/ncu_refresh.sh + tools/ncu_summary.py);
    None when there is no such capture -- never a number for another kernel or grid."""
    prefix = {"rows_fwd": "k_rows_fwd_r16<", "rows_inv": "k_rows_inv_r16<", "cols_fused": "k_cols_fused_p2_lr<"}.get(stage)
    if prefix is None or world != 1:
        return None, None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_full_summary.json")))
    except Exception:
        return None, None
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for name, m in summ.items():
        if name.startswith(prefix) and m.get("grid") == "%dx%d ndof %d" % (nx, ny, d):
            if stage.startswith("rows") and "variant 4104" not in kernels:
                continue
            if stage == "cols_fused" and "k_cols_fused_p2_lr" not in kernels:
                continue
            def gb(x):
                v, unit = x.split()[:2]
                return float(v) * mult[unit]
            return gb(m["dram__bytes_read.sum"]) + gb(m["dram__bytes_write.sum"]), \
                "ncu --set full capture of %s on this grid, profiles/r2_ncu_full_summary.json (not re-measured in this run)" % name
    return None, None


def stage_bytes_per_cell(d, fused_gather=False, fused_scatter=False):
    """Algorithmic bytes per cell of every stage; long columns (nx > 4096) add the
    top-digit pass of the column transform, one more read + write of the half spectrum each way.
    Fused atom I/O (cell -> atom map): the forward rows read x, xeq (48 B) and the map (4 B) per atom
    and write the half spectrum (8 B per dof); the inverse rows read it and the map and update f
    (24 B read + 24 B written per atom) -- the u_xy / f_xy grids do not exist."""
    nu = d // 3
    b = {"gather": 64 * nu + 8 * d, "rows_fwd": 16 * d, "cols_top_fwd": 16 * d, "cols_fused": 16 * d + 4 * d * d,
         "cols_top_inv": 16 * d, "rows_inv": 16 * d, "scatter": 8 * d + 64 * nu}
    if fused_gather:
        b["rows_fwd"] = 52 * nu + 8 * d
    if fused_scatter:
        b["rows_inv"] = 8 * d + 52 * nu
    return b


KERNEL_OF_STAGE = {
    "gather": "k_gather", "rows_fwd": "k_rows_fwd_*", "cols_top_fwd": "k_cols_top_pass<-1>",
    "cols_fused": "k_cols_fused_p2_lr (x-FFT + Phi.u + energy + gamma point + x-IFFT)",
    "cols_top_inv": "k_cols_top_pass<+1>", "rows_inv": "k_rows_inv_*", "scatter": "k_scatter + k_sum_partials"}


class Ctx:
    pass


def run_config(c, nx, ny, d, legacy_inputs, steps, warmup, e2e_steps, want_clocks):
    """One workload on the ranks of this job: set-up, timed steps, stage times, e2e.  Returns a dict
    (rank 0) with everything measured, and keeps the solver + fields in c for the parity check."""
    import torch
    import torch.distributed as dist
    import gfmd_b200
    from gfmd_b200 import synthetic
    world, rank, local, dev = c.world, c.rank, c.local, c.dev

    uid = None
    if world > 1:
        buf = torch.zeros(gfmd_b200.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(gfmd_b200.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    s = gfmd_b200.GFMDSolverB200(device=local, rank=rank, nranks=world, unique_id=uid)
    s.set_grid_size(nx, ny, d)
    exchange = "none (single GPU)"
    if world > 1:
        try:
            s.enable_peer_copy(gfmd_b200.all_gather_bytes_fn(dev, world))
            exchange = "CUDA IPC peer mappings over NVLink"
        except gfmd_b200.GFMDError as ex:          # still a GPU path: grouped ncclSend/ncclRecv
            exchange = "NCCL send/recv (peer mappings unavailable: %s)" % ex
    # stiffness table of the reference's `sc100 height 128` kernel for this rank's q columns: the per-q
    # matrices U0, U, V in closed form (== the plugin's get_dynamical_matrices, tests/test_oracle.py),
    # evaluated on the GPU; the transfer-matrix recursion runs on the GPU (gfmd_b200_build_phi_columns*)
    t_setup = time.time()
    ch = max(8, min(64, (1 << 20) // nx))
    for k0 in range(s.kylo, s.kylo + s.nky, ch):
        nk = min(ch, s.kylo + s.nky - k0)
        if legacy_inputs:
            s.build_kernel_columns(synthetic.sc100_dynamical_matrices(nx, ny, k0, nk), k0, height=128)
        else:
            s.build_kernel_columns_device(synthetic.sc100_dynamical_matrices_torch(nx, ny, k0, nk, dev), k0, nk,
                                          height=128)
    s.set_linf(np.zeros(d // 3))

    nx_loc, x0 = nx // world, rank * (nx // world)
    if legacy_inputs:       # the host generators the stored energy of the 4096^2 surface was computed with
        u_h = synthetic.displacement_field(nx, ny, x0, nx_loc, seed=1, nwaves=8)
        x, xeq, gid, mask = synthetic.atoms_for_slab(nx, ny, x0, nx_loc, u_h)
        du = torch.tensor(u_h, device=dev)
        dx, dxeq = torch.tensor(x, device=dev), torch.tensor(xeq, device=dev)
        dgid, dmask = torch.tensor(gid, device=dev), torch.tensor(mask, device=dev)
    else:                   # every value a function of the global cell index, generated on the GPU
        du = synthetic.displacement_field_torch(nx, ny, x0, nx_loc, dev, seed=1, nwaves=8)
        dx, dxeq, dgid, dmask = synthetic.atoms_for_slab_torch(nx, ny, x0, nx_loc, du)
    nat = dx.shape[0]
    dfat = torch.zeros((nat, 3), device=dev, dtype=torch.float64)
    stream = c.stream
    s.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    # fused gather / scatter through a cell -> atom map: OPT-IN (BENCH_FUSED_IO=1).  Measured slower than the
    # separate kernels (profiles/r2_fused_atom_io_ab.md: per-dof CTAs touch 8 of every 24 bytes of x / xeq / f)
    fused_io = os.environ.get("BENCH_FUSED_IO", "0") == "1" and s.build_cell_map(dgid, dmask, 1, nat, nat)
    t_setup = time.time() - t_setup

    def step():
        s.full_step(dx, dxeq, dgid, dmask, 1, nat, nat, float(nx), float(ny), dfat)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    sampler = ClockSampler(local) if (rank == 0 and want_clocks) else None
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        # short sustained load so that the clock samples are taken under load
        t_end = time.time() + 0.6
        while time.time() < t_end:
            for _ in range(3):
                step()
            s.synchronize()
        barrier()
        l0 = s.launch_count()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = s.launch_count() - l0
        res = s.results()
        dfat.zero_()
        step()                      # forces of exactly one step, for the parity check
        s.synchronize()
        f_one = dfat.clone()
        t_tail = time.time() + 0.3
        while time.time() < t_tail:
            for _ in range(3):
                step()
            s.synchronize()
    ms = max_over_ranks(ms)
    clocks = sampler.stop(t_wall0 - 0.7, t_wall1 + 0.4) if sampler else None
    ms_per_step = ms / steps

    # solver only (device-resident grids), and per-stage times with CUDA events
    with torch.cuda.stream(stream):
        barrier()
        e0.record(stream)
        for _ in range(steps):
            s.post_force_device()
        e1.record(stream)
        barrier()
        ms_solver = max_over_ranks(e0.elapsed_time(e1)) / steps
        s.profile(True)
        for _ in range(min(steps, 10)):
            step()
        s.profile(False)
        barrier()
    stages = s.stage_times()
    stage_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in stages.items()}

    # roofline per stage: algorithmic bytes of this rank's cells over the stage's CUDA-event time
    peaks, peak_src = measured_peaks()
    cells_loc = nx * ny / world
    fused_gather = fused_io and stage_ms.get("gather", 0.0) == 0.0
    bpc = stage_bytes_per_cell(d, fused_gather, fused_io)
    kernel_of = dict(KERNEL_OF_STAGE)
    if fused_gather:
        kernel_of["rows_fwd"] = "k_rows_fwd_* with fused gather (atoms -> half spectrum)"
    if fused_io:
        kernel_of["rows_inv"] = "k_rows_inv_* with fused scatter (half spectrum -> atoms)"
    roof_stages = {}
    for k, b in bpc.items():
        t = stage_ms.get(k, 0.0)
        if t > 0:
            a = b * cells_loc / (t * 1e-3) / 1e9
            roof_stages[k] = {"kernel": kernel_of[k], "ms": t, "alg_bytes_per_cell": b, "achieved": a,
                              "frac": a / peaks["hbm_gbs"]}
    solver_keys = ("rows_fwd", "cols_top_fwd", "cols_fused", "cols_top_inv", "rows_inv")
    dom = max((k for k in solver_keys if k in roof_stages), key=lambda k: roof_stages[k]["ms"])
    rs = roof_stages[dom]
    roofline = {"kernel": rs["kernel"], "stage": dom, "bound": "hbm", "achieved": rs["achieved"],
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": rs["frac"],
                "traffic": None, "traffic_source": None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rs["alg_bytes_per_cell"] * cells_loc, "kernel_ms": rs["ms"],
                "share_of_step": rs["ms"] / ms_per_step,
                "dominant_of": "FFT + contraction stages of this workload (gather / scatter listed in `stages`)",
                "stages": roof_stages,
                "solver_bytes_per_step": (80 * d + 4 * d * d) * cells_loc,
                "solver_frac": ((80 * d + 4 * d * d) * cells_loc / (ms_solver * 1e-3) / 1e9) / peaks["hbm_gbs"]}

    roofline["traffic"], roofline["traffic_source"] = ncu_traffic(dom, nx, ny, d, world, s.describe())
    # size-independent property on the timed workload (linf = 0): E = -1/2 sum_r f.u  
    fu = float((f_one * (dx - dxeq)).sum().item())
    esum = res["epot"]
    if world > 1:
        t = torch.tensor([fu, esum], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        fu, esum = float(t[0].item()), float(t[1].item())
    identity_err = abs(esum + 0.5 * fu) / abs(esum) if esum else None

    # end to end through the plugin boundary with pinned host buffers
    n_e2e = e2e_steps if e2e_steps > 0 else max(3, min(steps, 10))
    hu = torch.empty((d, nx_loc * ny), dtype=torch.float64, pin_memory=True)
    hu.copy_(du.reshape(d, nx_loc * ny))
    hf = torch.empty((d, nx_loc * ny), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()

    def time_e2e():
        for _ in range(2):
            s.post_force(hu, hf)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            s.post_force(hu, hf)
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / n_e2e)

    hp_default = s.host_pipeline()
    dt_e2e = time_e2e()
    grid_bytes = d * nx_loc * ny * 8
    e2e = {"value": 1.0 / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": grid_bytes,
           "d2h_bytes_per_step": grid_bytes + 8 * (d + 1), "steps": n_e2e,
           "call": "gfmd_b200_post_force_host (GFMDSolver::post_force boundary), pinned host u_xy/f_xy, per rank",
           "host_pipeline": hp_default,
           "pcie_floor_ms": 2 * grid_bytes / 55e9 * 1e3}
    del hu, hf

    out = {"value": 1e3 / ms_per_step, "ms_per_step": ms_per_step, "gpu_launches": int(launches),
           "clocks": clocks, "e2e": e2e, "roofline": roofline,
           "solver_only": {"value": 1e3 / ms_solver, "unit": UNIT, "ms_per_step": ms_solver},
           "stage_ms": stage_ms, "epot": esum, "energy_identity_rel_err": identity_err,
           "setup_s": t_setup, "exchange": exchange, "kernels": s.describe(), "fused_atom_io": bool(fused_io),
           "l2": "inputs larger than L2 (%.0f MB of atoms + grids per rank and step)"
                 % ((nat * (48 + 16 + 24) + 2 * grid_bytes) / 1e6)}
    c.solver, c.f_one, c.nat = s, f_one, nat
    return out


def slab_parity(c, nx, ny, d):
    """N > 1: rank 0 runs the same surface on ITS GPU alone (single-GPU handle, same inputs) and
    compares every rank's slab forces and the summed energy with it."""
    import torch
    import torch.distributed as dist
    import gfmd_b200
    from gfmd_b200 import synthetic
    world, rank, dev = c.world, c.rank, c.dev
    nx_loc = nx // world
    n_loc = nx_loc * ny
    e_slab = torch.tensor([c.solver.results()["epot"]], device=dev, dtype=torch.float64)
    dist.all_reduce(e_slab)
    out = None
    if rank == 0:
        one = gfmd_b200.GFMDSolverB200(device=c.local)
        one.set_grid_size(nx, ny, d)
        ch = max(8, min(64, (1 << 20) // nx))
        for k0 in range(0, one.nky, ch):
            nk = min(ch, one.nky - k0)
            one.build_kernel_columns_device(synthetic.sc100_dynamical_matrices_torch(nx, ny, k0, nk, dev), k0, nk,
                                            height=128)
        one.set_linf(np.zeros(d // 3))
        du = synthetic.displacement_field_torch(nx, ny, 0, nx, dev, seed=1, nwaves=8)
        dx, dxeq, dgid, dmask = synthetic.atoms_for_slab_torch(nx, ny, 0, nx, du)
        del du
        nat = dx.shape[0]
        fat = torch.zeros((nat, 3), device=dev, dtype=torch.float64)
        torch.cuda.synchronize()

        def step1():
            one.full_step(dx, dxeq, dgid, dmask, 1, nat, nat, float(nx), float(ny), fat)
        for _ in range(2):
            step1()
        one.synchronize()
        t0 = time.perf_counter()
        n1 = 5
        for _ in range(n1):
            step1()
        one.synchronize()
        ms1 = (time.perf_counter() - t0) / n1 * 1e3
        fat.zero_()
        step1()
        r1 = one.results()
        fmax = float(fat.abs().max().item())
        err = float((c.f_one - fat[:n_loc]).abs().max().item()) / fmax
        buf = torch.empty((n_loc, 3), device=dev, dtype=torch.float64)
        for r in range(1, world):
            dist.recv(buf, src=r)
            err = max(err, float((buf - fat[r * n_loc:(r + 1) * n_loc]).abs().max().item()) / fmax)
        out = {"parity_max_rel_err": err,
               "parity_epot_rel_err": abs(float(e_slab.item()) - r1["epot"]) / abs(r1["epot"]),
               "parity_against": "single-GPU run of the same %dx%d surface on rank 0 of this job (gfmd_b200_full_step)"
                                 % (nx, ny),
               "single_gpu_same_box": {"value": 1e3 / ms1, "unit": UNIT, "ms_per_step": ms1, "steps": n1,
                                       "timing": "wall clock around %d synchronised steps" % n1}}
        one.close()
    else:
        dist.send(c.f_one, dst=0)
    dist.barrier()
    return out


def latency_configs(c):
    """BASELINE configs C1-C3 (small grids): device time of one solver step through the captured CUDA
    graph (gfmd_b200_use_graph) next to the reference solver's CPU time on the same table."""
    import torch
    import gfmd_b200
    out = {}
    cases = [("C1_sc100_128x128", "tests/TEST_Hertz_sc100_128x128"),
             ("C2_fcc111_64x37", "tests/TEST_Hertz_fcc111_64x37_2"),
             ("C3_fcc100_two_layers_10x10", "tests/TEST_energy_conservation_two_layers")]
    for name, ref in cases:
        p = os.path.join(ROOT, "tests", "golden", name + ".npz")
        if not os.path.exists(p):
            continue
        z = np.load(p)
        nx, ny, d = int(z["nx"]), int(z["ny"]), int(z["ndof"])
        s = gfmd_b200.GFMDSolverB200(device=c.local)
        s.set_grid_size(nx, ny, d)
        s.set_kernel(z["phi"], z["linf"])
        s.set_stream(c.stream.cuda_stream)
        u = torch.tensor(z["u_uniform"].reshape(d, nx * ny), device=c.dev)
        f = torch.empty_like(u)
        rec = {"grid": "%dx%d" % (nx, ny), "ndof": d, "reference_test": ref}
        for mode in ("launches", "graph"):
            s.use_graph(mode == "graph")
            with torch.cuda.stream(c.stream):
                for _ in range(20):
                    s.post_force_device(u, f)
                s.synchronize()
                n = 500
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(c.stream)
                for _ in range(n):
                    s.post_force_device(u, f)
                e1.record(c.stream)
                s.synchronize()
                us = e0.elapsed_time(e1) / n * 1e3
            rec["us_per_step_" + mode] = us
        rec["steps_per_s"] = 1e6 / min(rec["us_per_step_launches"], rec["us_per_step_graph"])
        try:
            from oracle import gfmd_oracle as O
            if O.ref_available():
                rs = O.RefSolver(nx, ny, d, fft_backend=1)
                rs.set_phi(z["phi"], z["linf"])
                uu = np.ascontiguousarray(z["u_uniform"])
                for _ in range(3):
                    rs.post_force(uu)
                t0 = time.perf_counter()
                nrep = 50
                for _ in range(nrep):
                    rs.post_force(uu)
                rec["reference_cpu_us_per_step"] = (time.perf_counter() - t0) / nrep * 1e6
        except Exception as ex:
            rec["reference_cpu_us_per_step"] = None
            rec["reference_error"] = repr(ex)
        s.close()
        out[name] = rec
    return out


def main_b200(args):
    import torch
    import torch.distributed as dist

    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL prints its version banner with a plain printf to stdout when NCCL_DEBUG is set;
    # keep stdout for the single JSON line: everything else of this process goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    c.stream = torch.cuda.Stream(device=c.dev)
    world, rank = c.world, c.rank
    if world > 1:
        dist.init_process_group("nccl", device_id=c.dev)

    nx, ny, d = workload(args)
    # N = 1: the 4096 x 4096 sub-record first, on a GPU that the long 16384^2 run has not yet driven into
    # its power cap (the order is stated in the line; each record carries its own clock samples)
    sub = None
    if world == 1 and not args.no_4096 and args.grid in (0, STRONG_GRID):
        sub = run_config(c, 4096, 4096, d, True, max(args.steps, 50), max(args.warmup, 5), args.e2e_steps, True)
        c.solver.close()
        del c.f_one
        torch.cuda.empty_cache()
        sub["workload"] = workload_string(4096, 4096, d)
        sub["measured"] = "before the 16384 x 16384 record of the same line"
        sub["epot_reference"] = EPOT_4096
        sub["epot_reference_source"] = ("stored constant; regenerate with tools/epot_reference_4096.py "
                                        "(independent numpy computation, ~16 min of CPU)")
        sub["epot_rel_err"] = abs(sub["epot"] - EPOT_4096) / EPOT_4096

    main = run_config(c, nx, ny, d, False, args.steps, args.warmup, args.e2e_steps, True)

    parity = None
    if world > 1 and not args.no_parity:
        parity = slab_parity(c, nx, ny, d)

    nvlink = None
    if world > 1:
        cells_loc = nx * ny / world
        per_dir = 8.0 * d * cells_loc * (world - 1) / world          # bytes sent per GPU per transpose
        sm = main["stage_ms"]
        nvlink = {"bytes_out_per_gpu_per_step": 2 * per_dir, "bytes_in_per_gpu_per_step": 2 * per_dir,
                  "peak_gbs_per_dir": 770.0, "peak_source": "B200_PROFILING.md measured peer copy",
                  "min_transfer_ms_per_step": 2 * per_dir / 770e9 * 1e3,
                  "achieved_gbs_per_dir_if_fully_exposed": 2 * per_dir / (main["ms_per_step"] * 1e-3) / 1e9,
                  "frac": (2 * per_dir / 770e9 * 1e3) / main["ms_per_step"],
                  "frac_meaning": "share of the step the transfers would take alone at the peak rate; the exchange "
                                  "overlaps the column stage and two thirds of the forward rows (DESIGN.md 5)",
                  "non_overlapped_exchange_ms": sm.get("exchange_fwd", 0.0) + sm.get("exchange_inv", 0.0)}
    c.solver.close()
    del c.f_one

    lat = cpu = None
    if world == 1 and rank == 0:
        try:
            lat = latency_configs(c)
        except Exception as ex:
            lat = {"error": repr(ex)}
        if not args.no_cpu_baseline:
            try:
                r = reference_run(nx, ny, d, 3, 1)
                cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "fft")}
                cpu["measured"] = r["measured"]
                if sub is not None:
                    sub["cpu_baseline"] = dict(r["measured"], unit=UNIT, cores=r["cores"], kind=r["kind"],
                                               sample="GFMDSolverStatic::post_force, 4096x4096 in full, 3 timed steps; "
                                                      "FFT = oracle/fft_plain.c")
            except Exception as ex:   # the oracle is a checker; its absence must not hide the GPU number
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        out = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_string(nx, ny, d),
                          "step": "gather + forward FFT + Phi.u + inverse FFT + scatter, device resident",
                          "decomposition": "x-slabs over %d GPU(s); transposes: %s" % (world, main["exchange"]),
                          "l2": main["l2"], "kernels": main["kernels"],
                          "fused_atom_io": main["fused_atom_io"],
                          "inputs": "displacement field and atoms generated on the GPU from the global cell index "
                                    "(gfmd_b200.synthetic.*_torch), atoms in grid order",
                          "latency": lat},
               "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
               "roofline": main["roofline"], "cpu_baseline": cpu, "nvlink": nvlink,
               "solver_only": main["solver_only"], "stage_ms": main["stage_ms"], "epot": main["epot"],
               "energy_identity_rel_err": main["energy_identity_rel_err"], "setup_s": main["setup_s"]}
        if parity:
            out.update(parity)
        elif world == 1:
            out["parity_max_rel_err"] = main["energy_identity_rel_err"]
            out["parity_against"] = "size-independent identity E = -1/2 sum f.u on the timed surface (linf = 0)"
        if sub is not None:
            out["grid_4096"] = sub
            try:    # the north-star's single-GPU roofline target is quoted on THIS surface: repeat its fractions up front
                st = sub["roofline"]["stages"]
                out["roofline"]["north_star_surface_4096x4096"] = {
                    "frac_rows_fwd": st["rows_fwd"]["frac"], "frac_cols_fused": st["cols_fused"]["frac"],
                    "frac_rows_inv": st["rows_inv"]["frac"], "peak": sub["roofline"]["peak"], "unit": "GB/s",
                    "note": "FFT and fused Phi(q) stages of the 4096x4096 sc100 surface (grid_4096.roofline.stages); the "
                            "top-level roofline object describes the dominant kernel of the 16384x16384 workload of "
                            "`value`, whose rows (one 128 KB row per CTA) are the weakest stage"}
            except Exception:
                pass
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
