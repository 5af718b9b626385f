"""Device-side table builder with `height -1` (semi-infinite substrate: iterate the continued
fraction to the reference's 1e-8 convergence, surface_stiffness.cpp:849-851 / iterate_Gnn :493-548)
against the plugin's own table.  Emulator-verified
(tests/test_emulated_kernels.py::test_device_built_table_equals_plugin_table) and GPU-verified.  Tolerance 1e-9 on forces here: both sides stop at a
1e-8 change of VT, not at an exact fixed point, and they stop at the same iteration only up to
rounding (Gauss-Jordan vs elimination with partial pivoting)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,nx,ny", [
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height -1", 16, 12),
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height -1", 8, 7),
])
def test_converged_table_equals_plugin_table(kernel, nx, ny, oracle_libs):
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    import gfmd_b200
    k = O.RefKernel(kernel)
    assert k.height() < 0
    d = k.ndof
    u = np.random.default_rng(9).uniform(-0.1, 0.1, size=(d, nx * ny))
    out = []
    for mode in ("plugin", "device"):
        s = gfmd_b200.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        if mode == "plugin":
            s.set_kernel(k.phi(nx, ny), k.linf())
        else:
            s.build_kernel_columns(k.dynamical_matrices(nx, ny, 0, s.nky), 0, height=k.height())
            s.set_linf(k.linf())
        f = np.zeros_like(u)
        e = s.post_force(u, f)
        out.append((f, e))
        s.close()
    assert rel_err(out[1][0], out[0][0]) < 1e-9
    assert abs(out[1][1] - out[0][1]) <= 1e-9 * abs(out[0][1])
    k.close()
