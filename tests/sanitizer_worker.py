"""Small-grid steps of every kernel family, run under compute-sanitizer by tests/test_sanitizer_gpu.py
(memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards between barriers --
what the CPU emulation's thread-order permutation cannot see).  Checks results too, so that a run
that "passes" the sanitizer has really executed the kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
import gfmd_b200  # noqa: E402
from gfmd_b200 import synthetic  # noqa: E402


def golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    nx, ny, d = int(z["nx"]), int(z["ny"]), int(z["ndof"])
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(z["phi"], z["linf"])
    u = np.ascontiguousarray(z["u_uniform"].reshape(d, nx * ny))
    f = np.zeros_like(u)
    e = s.post_force(u, f)
    err = np.abs(f.reshape(d, nx, ny) - z["f_uniform"]).max() / np.abs(z["f_uniform"]).max()
    assert err < 1e-11 and abs(e - float(z["epot_uniform"])) <= 1e-11 * abs(float(z["epot_uniform"])), (name, err)
    s.close()
    return s


def synthetic_case(nx, ny, expect):
    d = 3
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    assert expect in s.describe(), s.describe()
    for k0 in range(0, s.nky, 512):
        nk = min(512, s.nky - k0)
        s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
    s.set_linf(np.zeros(1))
    rng = np.random.default_rng(nx + ny)
    u = rng.uniform(-1e-3, 1e-3, size=(d, nx * ny))
    f = np.full_like(u, np.nan)
    e = s.post_force(u, f)
    assert np.isfinite(f).all() and abs(e + 0.5 * np.sum(f * u)) <= 1e-11 * abs(e)      # E = -1/2 sum f.u (linf = 0)
    s.close()


def main():
    golden("small_sc100_16x12")            # generic radix kernels
    golden("C2_fcc111_64x37")              # Bluestein rows, ndof 6
    synthetic_case(4, 4096, "k_rows_*_r16")        # radix-16 rows (2 rows / CTA)
    synthetic_case(2, 8192, "k_rows_*_r16")        # half-unit radix-16 rows
    synthetic_case(2, 16384, "k_rows_*_r16")       # one row per CTA
    synthetic_case(4096, 4, "k_cols_fused_p2_lr")  # specialised fused column kernel (software-pipelined)
    synthetic_case(8192, 4, "top radix 2")         # + top-digit passes
    os.environ["GFMD_B200_ROWS_VARIANT"] = "16393"
    synthetic_case(2, 16384, "2-CTA clusters")     # cluster rows, distributed shared memory
    print("SANITIZER_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
