"""Device-side stiffness-table builder (gfmd_b200_build_phi_columns: the transfer-matrix
recursion of surface_stiffness.cpp:811-873 on the GPU) against the reference plugin's own
tables, end to end: forces and energy of a step with the device-built table must equal the
golden vectors / the host-plugin table to 1e-11."""
import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-11


def test_sc100_table_built_on_device_reproduces_golden_forces():
    import gfmd_b200
    from gfmd_b200 import synthetic
    g = load_golden("C1_sc100_128x128")            # plugin table: ft sc100 ... height 128 == sc100 height 128
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    for k0 in range(0, s.nky, 16):
        nk = min(16, s.nky - k0)
        s.build_kernel_columns(synthetic.sc100_dynamical_matrices(nx, ny, k0, nk), k0, height=128)
    s.set_linf(g["linf"])
    for c in ("uniform", "hertz", "shift"):
        u = np.ascontiguousarray(g["u_" + c].reshape(d, nx * ny))
        f = np.zeros_like(u)
        e = s.post_force(u, f)
        assert rel_err(f.reshape(d, nx, ny), g["f_" + c]) < TOL, c
        assert abs(e - float(g["epot_" + c])) <= TOL * abs(float(g["epot_" + c])), c
    s.close()


@pytest.mark.parametrize("kernel,nx,ny", [
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 128", 64, 37),
    ("ft fcc100 1.0 2 pair-potential 2 1.0 -0.1 height 10", 10, 10),
    ("sc100 height 0", 16, 12),
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 7", 2048, 8),     # specialised column layout
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 5", 4096, 4),        # position-ordered table (k_cols_fft_p2)
])
def test_device_built_table_equals_plugin_table(kernel, nx, ny, oracle_libs):
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    import gfmd_b200
    k = O.RefKernel(kernel)
    d = k.ndof
    phi = k.phi(nx, ny)
    rng = np.random.default_rng(9)
    u = rng.uniform(-0.1, 0.1, size=(d, nx * ny))
    out = []
    for mode in ("plugin", "device"):
        s = gfmd_b200.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        if mode == "plugin":
            s.set_kernel(phi, k.linf())
        else:
            s.build_kernel_columns(k.dynamical_matrices(nx, ny, 0, s.nky), 0, height=k.height())
            s.set_linf(k.linf())
        f = np.zeros_like(u)
        e = s.post_force(u, f)
        out.append((f, e))
        s.close()
    assert rel_err(out[1][0], out[0][0]) < TOL
    assert abs(out[1][1] - out[0][1]) <= TOL * abs(out[0][1])
    k.close()


@pytest.mark.parametrize("nx,ny,height", [(6, 5, 3), (16, 12, 2)])
def test_explicit_spring_network(nx, ny, height):
    """Reference-independent physical anchor of the whole CUDA path (tests/spring_network.py):
    closed-form per-q matrices -> transfer-matrix recursion on the device -> FFT, contraction,
    inverse FFT, against an explicit spring network relaxed by a dense solve.  Pins the FFT
    sign convention, the substrate orientation and the meaning of `height` without any table
    or FFT library."""
    import gfmd_b200
    import spring_network
    from gfmd_b200 import synthetic
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, 3)
    s.build_kernel_columns(synthetic.sc100_dynamical_matrices(nx, ny, 0, s.nky), 0, height=height)
    s.set_linf(np.zeros(1))
    u0 = np.random.default_rng(4).uniform(-0.1, 0.1, size=(3, nx, ny))
    f = np.full((3, nx * ny), np.nan)
    e = s.post_force(np.ascontiguousarray(u0.reshape(3, nx * ny)), f)
    s.close()
    fs = spring_network.surface_force(u0, height)
    assert rel_err(f.reshape(3, nx, ny), fs) < 1e-12
    assert abs(e + 0.5 * float((fs * u0).sum())) <= 1e-12 * abs(e)
