"""GPU tests of the off-path solver services through the C ABI:
the q-space fields GFMDSolverFFT::dump writes on `dumpq_every` steps and
GFMDSolverStatic::prec_gradient, against outputs of the reference's own sources recorded in
tests/golden/aux (tests/golden/make_golden_aux.py) and against the oracle.
Tolerance 1e-11 relative (2e-10 for fields parsed back from the reference's text dumps)."""
import numpy as np
import pytest

import aux_checks
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gfmd_b200
    gfmd_b200.load_library()          # raises if the CUDA library is missing: no fallback
    return gfmd_b200


@pytest.mark.parametrize("name", aux_checks.aux_names())
def test_spectrum_and_dump_fields(B, name, oracle_libs):
    aux_checks.check_spectrum(B, oracle_libs, name)


@pytest.mark.parametrize("name", aux_checks.aux_names())
def test_prec_gradient(B, name, oracle_libs):
    aux_checks.check_prec_gradient(B, oracle_libs, name)


@pytest.mark.parametrize("nx,ny,d", [(2048, 16, 3), (4096, 8, 3), (243, 50, 3), (11, 13, 15), (64, 2048, 6)])
def test_spectrum_on_other_layouts(B, nx, ny, d, oracle_libs):
    """Specialised (digit-reversed, interleaved) table layout, Bluestein columns, run-time ndof,
    specialised row kernels selected for the per-step path."""
    O = oracle_libs
    rng = np.random.default_rng(nx + ny + d)
    Dr = rng.standard_normal((nx, ny, d, d)) * np.exp(-3 * rng.random((nx, ny, 1, 1)))
    Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
    Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
    phi = np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)
    u = rng.uniform(-0.1, 0.1, size=(d, nx, ny))
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(phi, np.zeros(d // 3))
    uq, fq = s.spectrum(np.ascontiguousarray(u.reshape(d, nx * ny)))
    uq_ref, fq_ref = O.spectrum(u, phi)
    assert rel_err(uq, uq_ref) < aux_checks.TOL and rel_err(fq, fq_ref) < aux_checks.TOL
    s.close()


@pytest.mark.parametrize("nx,ny,d", [(4096, 8, 6), (8192, 4, 3), (4096, 4096, 3), (6000, 6, 6)])
def test_services_on_column_sets_beyond_one_cta(B, nx, ny, d, oracle_libs):
    """Round 2: no 227 KB limit any more -- e.g. a `dumpq_every` step on the 4096 x 4096 bench surface."""
    if nx * ny > 1 << 22:
        # full surface: spectrum only against numpy's FFT of a smooth field (the oracle's table route needs 2.4 GB)
        from gfmd_b200 import synthetic
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        for k0 in range(0, s.nky, 256):
            nk = min(256, s.nky - k0)
            s.set_kernel_columns(synthetic.phi_columns(nx, ny, k0, nk), k0, normalized=False)
        s.set_linf(np.zeros(1))
        u = synthetic.displacement_field(nx, ny, seed=3, nwaves=4)
        uq, _ = s.spectrum(np.ascontiguousarray(u.reshape(d, nx * ny)), with_force=False)
        ref = np.moveaxis(np.fft.fft2(u, axes=(1, 2)).reshape(d, nx * ny), 0, 1)
        assert rel_err(uq, ref) < aux_checks.TOL
        s.close()
        return
    aux_checks.check_large_column_sets(B, oracle_libs, nx, ny, d)


def test_unsupported_sizes_fail_loudly(B):
    s = B.GFMDSolverB200()
    s.set_grid_size(16384, 4, 3)       # one column (256 KB) exceeds a CTA's shared memory
    from gfmd_b200 import synthetic
    s.set_kernel_columns(synthetic.phi_columns(16384, 4, 0, s.nky), 0, normalized=False)
    u = np.zeros((3, 16384 * 4))
    with pytest.raises(B.GFMDError) as ei:
        s.spectrum(u)
    assert ei.value.code == 4
    s.close()
