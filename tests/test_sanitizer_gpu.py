"""compute-sanitizer over the small-grid cases of every kernel family (the race / memory
checks the reference leaves to valgrind-style tooling).  memcheck: out-of-bounds and misaligned accesses;
racecheck: shared-memory hazards."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = os.path.join(ROOT, "tests", "sanitizer_worker.py")


def sanitizer():
    for c in ("compute-sanitizer", "/usr/local/cuda/bin/compute-sanitizer"):
        p = shutil.which(c)
        if p:
            return p
    return None


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_kernels_under_compute_sanitizer(tool):
    exe = sanitizer()
    if exe is None:
        pytest.skip("compute-sanitizer not installed")
    r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "9", sys.executable, WORKER],
                       capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert "SANITIZER_WORKER_OK" in out, out[-4000:]
    assert r.returncode == 0, out[-4000:]
    if tool == "memcheck":
        assert "ERROR SUMMARY: 0 errors" in out, out[-4000:]
    else:
        assert "RACECHECK SUMMARY: 0 hazards" in out, out[-4000:]
