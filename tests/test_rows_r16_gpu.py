"""GPU parity of the radix-16 row kernels (user-gfmd_b200/csrc/kernels_rows_r16.cuh, variant ids ny + 8 of
GFMD_B200_ROWS_VARIANT: the default since round 2) and of the two-CTA-cluster form (ny + 9).  History:
emulator-verified first (tests/test_emulated_kernels.py::test_row_kernel_variants, also under other
thread orders), run and timed on B200s in round 2 (profiles/r2_rows_variants.txt).
Tolerance: <= 1e-11 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-11


@pytest.fixture(scope="module")
def B():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gfmd_b200
    gfmd_b200.load_library()          # raises if the CUDA library is missing: no fallback
    return gfmd_b200


def random_case(nx, ny, d):
    rng = np.random.default_rng(1000 * nx + ny + d)
    Dr = rng.standard_normal((nx, ny, d, d)) * np.exp(-0.3 * rng.random((nx, ny, 1, 1)) * 10)
    Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
    Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
    phi = np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)
    return phi, rng.standard_normal(d // 3), rng.uniform(-0.1, 0.1, size=(d, nx, ny))


VARIANTS = [(64, 4096, 3, 4104), (6, 4096, 6, 4104), (32, 8192, 3, 8200), (2, 8192, 6, 8200), (16, 16384, 3, 16392),
            (3, 16384, 6, 16392), (16, 16384, 3, 16393), (4, 16384, 6, 16393)]


@pytest.mark.parametrize("nx,ny,d,variant", VARIANTS)
def test_r16_rows_against_oracle(B, nx, ny, d, variant, oracle_libs, monkeypatch):
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, u0_ref = O.post_force(u, phi, linf)
    monkeypatch.setenv("GFMD_B200_ROWS_VARIANT", str(variant))
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    assert "variant %d" % variant in s.describe() and "k_rows_*_r16" in s.describe(), s.describe()
    assert ("2-CTA clusters" in s.describe()) == (variant == ny + 9)
    s.set_kernel(phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    f = np.full_like(uu, np.nan)
    e = s.post_force(uu, f)
    assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL
    assert abs(e - e_ref) <= TOL * abs(e_ref)
    assert np.abs(s.get_u0() - u0_ref).max() <= TOL * max(1.0, np.abs(u0_ref).max())
    s.close()


@pytest.mark.parametrize("ny,variant", [(4096, 4104), (8192, 8200), (16384, 16392)])
def test_r16_rows_match_default_rows(B, ny, variant, monkeypatch):
    """Many row tiles (256 rows x 3 dofs, generic columns): radix-16 rows (the default since round 2)
    against the radix-8 rows of the same grid (variant id ny + 0), to rounding."""
    from gfmd_b200 import synthetic
    nx, d = 256, 3
    rng = np.random.default_rng(3)
    u = rng.uniform(-1e-3, 1e-3, size=(d, nx * ny))
    out = []
    for v in (ny, variant):
        monkeypatch.setenv("GFMD_B200_ROWS_VARIANT", str(v))
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        assert ("k_rows_*_r16" in s.describe()) == (v == variant), s.describe()
        for k0 in range(0, s.nky, 512):
            nk = min(512, s.nky - k0)
            s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        s.set_linf(np.array([0.25]))
        f = np.full_like(u, np.nan)
        e = s.post_force(u, f)
        out.append((f, e))
        s.close()
    assert rel_err(out[1][0], out[0][0]) < 1e-12
    assert abs(out[1][1] - out[0][1]) <= 1e-12 * abs(out[0][1])


def test_cluster_rows_bit_identical_to_single_cta_rows(B, monkeypatch):
    """ny = 16384: the two-CTA-cluster row kernels (variant ny + 9: staging accesses split by frequency,
    partner points through distributed shared memory) run the same arithmetic per element as the
    one-CTA-per-row kernels (ny + 8): forces and energy must agree bit for bit."""
    from gfmd_b200 import synthetic
    nx, ny, d = 64, 16384, 3
    rng = np.random.default_rng(9)
    u = rng.uniform(-1e-3, 1e-3, size=(d, nx * ny))
    out = []
    for v in (ny + 8, ny + 9):
        monkeypatch.setenv("GFMD_B200_ROWS_VARIANT", str(v))
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        assert "variant %d" % v in s.describe(), s.describe()
        for k0 in range(0, s.nky, 1024):
            nk = min(1024, s.nky - k0)
            s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        s.set_linf(np.array([0.25]))
        f = np.full_like(u, np.nan)
        e = s.post_force(u, f)
        out.append((f, e))
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
