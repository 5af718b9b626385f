import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_cases(g):
    return sorted(k[2:] for k in g if k.startswith("u_"))


@pytest.fixture(scope="session")
def oracle_libs():
    """Builds oracle/_ref/libgfmd_oracle.so (always) and libgfmd_ref.so (when the
    reference sources are present) if they are missing."""
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(odir, "_ref", "libgfmd_oracle.so")):
        subprocess.check_call(["make", "-C", odir, "oracle"], stdout=subprocess.DEVNULL)
    if (not os.path.exists(os.path.join(odir, "_ref", "libgfmd_ref.so"))
            and os.path.isdir("/root/reference/src")):
        subprocess.check_call(["make", "-C", odir, "ref"], stdout=subprocess.DEVNULL)
    from oracle import gfmd_oracle
    return gfmd_oracle


def rel_err(a, b):
    """max-norm error relative to max|b|."""
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)
