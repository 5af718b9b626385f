"""The slab-decomposed (multi-GPU) path without GPUs: tests/emu/multi_rank_worker.py runs the
ranks as host threads of one process on the CUDA-on-CPU emulation build, with in-process "IPC"
handles for the peer pushes and tests/emu/fake_nccl.cpp (collectives = real barriers) for NCCL.
Same checks as tests/mgpu_worker.py on real GPUs: golden vectors through the slab path, and
slab result == single-rank result, bit for bit, on grids that select the specialised kernels,
split columns and long columns.  Test infrastructure only (see tests/emu)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "emu", "multi_rank_worker.py")


@pytest.mark.parametrize("nranks,mode", [(2, "ipc"), (4, "ipc"), (2, "nccl")])
def test_slab_parity_emulated(nranks, mode):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    r = subprocess.run([sys.executable, WORKER, str(nranks)] + (["nccl"] if mode == "nccl" else []),
                       capture_output=True, text=True, timeout=1200)
    assert "EMU_MGPU_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_chunked_pipeline_emulated():
    """4096 x 2048 on two ranks: the grid size at which the chunked pipeline (per-dof row
    kernels, column chunks overlapping their own transposes, pipelined_step) is selected."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_EMU_GRIDS="4096x2048")
    r = subprocess.run([sys.executable, WORKER, "2"], env=env, capture_output=True, text=True, timeout=2400)
    assert "EMU_MGPU_PARITY_OK" in r.stdout and "4096x2048" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_slab_parity_emulated_split_columns():
    """The three-phase column stage (kernel_cols_split.cuh, forced with GFMD_B200_COLS_SPLIT) behind
    the exchange: columns arrive in P pieces and are transformed in place in the receive buffer."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_B200_COLS_SPLIT="2", GFMD_EMU_GRIDS="8x4096,16x2048,48x36")
    r = subprocess.run([sys.executable, WORKER, "2"], env=env, capture_output=True, text=True, timeout=1200)
    assert "EMU_MGPU_PARITY_OK" in r.stdout and "k_cols_split_fft" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4])
def test_in_kernel_peer_stores_emulated(nranks):
    """GFMD_B200_PEER_STORE=1: the last kernel of the column stage (the fused column kernel, or the
    backward top-digit pass of long columns) stores each result piece straight into its owner's
    return buffer; no return pushes.  nx = 4096 / 8192 / 16384 cover both kernels and 1-4 pieces
    per sub-column; 4096 x 32 on two ranks also takes the chunked pipeline."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_B200_PEER_STORE="1", GFMD_EMU_GRIDS="4096x8,4096x32,8192x8,16384x4")
    r = subprocess.run([sys.executable, WORKER, str(nranks)], env=env, capture_output=True, text=True, timeout=1800)
    assert "EMU_MGPU_PARITY_OK" in r.stdout and "in-kernel peer stores" in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4])
def test_no_transposes_emulated(nranks):
    """GFMD_B200_PEER_DIRECT=1 with the row-output buffers mapped (gfmd_b200_ipc_import_stage): after a
    barrier the column stage loads its pieces straight from the ranks that produced them (pass 0 of
    the column kernel, or the forward top-digit pass of long columns) and stores its results straight
    into their return buffers.  Two steps per grid, so a buffer reused too early would show."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_B200_PEER_DIRECT="1", GFMD_EMU_GRIDS="4096x8,4096x32,8192x8,16384x4,16x2048")
    r = subprocess.run([sys.executable, WORKER, str(nranks)], env=env, capture_output=True, text=True, timeout=1800)
    assert "EMU_MGPU_PARITY_OK" in r.stdout and "transposes: none" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
