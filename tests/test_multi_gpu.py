"""Slab-decomposed (multi-GPU, NCCL) parity; needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run_worker(nranks, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))


# ---- the slab path on a box with ONE GPU: the ranks are processes that all use cuda:0 ----------
# Same library code as on N GPUs (CUDA IPC mappings, copy-engine pushes, flag words written and
# awaited by the streams, in-kernel peer loads / stores); no NCCL (it refuses two ranks on a device).

ONE_DEVICE_MODES = {
    "pushes": {"GFMD_B200_PEER_DIRECT": "0"},          # copy-engine pushes also where the default is the step without transposes
    "peer_store": {"GFMD_B200_PEER_STORE": "1"},
    "no_transposes": {"GFMD_B200_PEER_DIRECT": "1"},
    "chunks8": {"GFMD_B200_CHUNKS": "8"},
}


ONE_DEVICE_MODES["default"] = {}


@pytest.mark.parametrize("nranks,mode", [(2, "pushes"), (4, "pushes"), (2, "peer_store"), (4, "no_transposes"),
                                         (2, "chunks8"), (4, "default")])
def test_slab_parity_one_device(nranks, mode):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    offset = sorted(ONE_DEVICE_MODES).index(mode)
    r = run_worker(nranks, 29700 + 10 * offset + nranks, dict(ONE_DEVICE_MODES[mode], GFMD_TEST_ONE_DEVICE="1"))
    assert "MGPU_PARITY_OK" in r.stdout and "stream flags" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    if mode == "peer_store":
        assert "in-kernel peer stores" in r.stdout
    if mode == "no_transposes":
        assert "transposes: none" in r.stdout


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity_nccl_barriers(nranks):
    """GFMD_B200_SYNC=nccl: the round-1 ordering (one-element all-reduces as barriers)."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    r = run_worker(nranks, 29670 + nranks, {"GFMD_B200_SYNC": "nccl"})
    assert "MGPU_PARITY_OK" in r.stdout and "NCCL all-reduce barriers" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity_send_recv(nranks):
    """No peer mappings at all: grouped ncclSend / ncclRecv transposes."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    r = run_worker(nranks, 29690 + nranks, {"GFMD_TEST_EXCHANGE": "nccl"})
    assert "MGPU_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + nranks),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MGPU_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity_in_kernel_peer_stores(nranks):
    """GFMD_B200_PEER_STORE=1 (opt-in): the column stage
    stores its result pieces straight into the peers' return buffers over NVLink."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29630 + nranks),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GFMD_B200_PEER_STORE="1"))
    assert "MGPU_PARITY_OK" in r.stdout and "in-kernel peer stores" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity_no_transposes(nranks):
    """GFMD_B200_PEER_DIRECT=1 (the default for >= 4 ranks with long columns): the column stage loads
    and stores its pieces in the peers' memory itself; no transposes."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + nranks),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GFMD_B200_PEER_DIRECT="1"))
    assert "MGPU_PARITY_OK" in r.stdout and "transposes: none" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
