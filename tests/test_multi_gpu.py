"""Slab-decomposed (multi-GPU, NCCL) parity; needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


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
    """GFMD_B200_PEER_STORE=1 (opt-in, emulator-verified, not yet run on GPUs): the column stage
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
    """GFMD_B200_PEER_DIRECT=1 (opt-in, emulator-verified, not yet run on GPUs): the column stage loads
    and stores its pieces in the peers' memory itself; no transposes."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + nranks),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GFMD_B200_PEER_DIRECT="1"))
    assert "MGPU_PARITY_OK" in r.stdout and "transposes: none" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
