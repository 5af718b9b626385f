"""CPU tests: pin the oracle restatements (numpy, plain C) against the golden
vectors produced by the reference's own sources, and -- where oracle/_ref was
built -- against those sources directly."""
import numpy as np
import pytest

from conftest import golden_cases, golden_names, load_golden, rel_err

TOL = 1e-13   # restatements differ from the reference arithmetic only by FFT rounding


@pytest.mark.parametrize("name", golden_names())
def test_numpy_oracle_matches_golden(name, oracle_libs):
    O = oracle_libs
    g = load_golden(name)
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    for c in golden_cases(g):
        f, e, u0 = O.post_force(g["u_" + c], g["phi"], g["linf"])
        assert rel_err(f, g["f_" + c]) < TOL
        assert abs(e - g["epot_" + c]) <= TOL * max(1.0, abs(g["epot_" + c]))
        assert np.abs(u0 - g["u0_" + c]).max() <= TOL * max(1.0, np.abs(g["u0_" + c]).max())


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("backend", [0, 1])
def test_c_oracle_matches_golden(name, backend, oracle_libs):
    O = oracle_libs
    g = load_golden(name)
    if backend == 0 and int(g["nx"]) * int(g["ny"]) > 4096:
        pytest.skip("long-double DFT backend only exercised on small grids")
    for c in golden_cases(g):
        f, e, u0 = O.c_post_force(g["u_" + c], g["phi"], g["linf"], backend)
        tol = 0.0 if backend == 0 else TOL     # same DFT as the golden run -> bit-exact
        assert rel_err(f, g["f_" + c]) <= tol
        assert abs(e - g["epot_" + c]) <= tol * max(1.0, abs(g["epot_" + c]))


def test_plain_fft_all_radices(oracle_libs):
    """oracle/fft_plain.c (mixed radix + Bluestein) against numpy's pocketfft."""
    import ctypes
    lib = oracle_libs.clib()
    lib.fftp_plan_2d.restype = ctypes.c_void_p
    lib.fftp_plan_2d.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.fftp_exec_2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.fftp_destroy.argtypes = [ctypes.c_void_p]
    rng = np.random.default_rng(5)
    for nx, ny in [(1, 1), (2, 3), (4, 8), (5, 7), (12, 30), (37, 64), (64, 37), (74, 11), (105, 128),
                   (13, 26), (256, 6)]:
        a = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
        p = lib.fftp_plan_2d(nx, ny)
        b = a.copy()
        lib.fftp_exec_2d(p, b.ctypes.data, -1)
        assert rel_err(b, np.fft.fft2(a)) < 1e-13, (nx, ny)
        lib.fftp_exec_2d(p, b.ctypes.data, +1)
        assert rel_err(b / (nx * ny), a) < 1e-13, (nx, ny)
        lib.fftp_destroy(p)


def test_reference_build_matches_golden(oracle_libs):
    """The reference's own solver sources (when built here) reproduce the
    committed golden vectors bit for bit, and the plugin tables too."""
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built (no /root/reference)")
    for name in ["small_sc100_16x12", "C3_fcc100_two_layers_10x10"]:
        g = load_golden(name)
        nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
        k = O.RefKernel(str(g["kernel"]))
        assert k.ndof == d
        phi = k.phi(nx, ny)
        assert np.array_equal(phi, g["phi"])
        s = O.RefSolver(nx, ny, d, fft_backend=0)
        s.set_kernel(k)
        for c in golden_cases(g):
            f, e, u0 = s.post_force(g["u_" + c])
            assert np.array_equal(f, g["f_" + c])
            assert e == g["epot_" + c]
        s.close()
        k.close()


def test_phi_symmetries(oracle_libs):
    """Phi Hermitian and Phi(-q) = conj Phi(q): what the half-spectrum,
    Hermitian-packed device table relies on."""
    for name in golden_names():
        g = load_golden(name)
        nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
        P = g["phi"].reshape(nx, ny, d, d)
        scale = np.abs(P).max()
        assert np.abs(P - np.conj(np.swapaxes(P, 2, 3))).max() < 1e-14 * scale
        Pm = P[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
        assert np.abs(P - np.conj(Pm)).max() < 1e-14 * scale


def test_energy_identity(oracle_libs):
    """E = -1/2 sum_r f.u when linf = 0 (SURVEY 8a restatement)."""
    g = load_golden("small_sc100_16x12")
    f, e, _ = oracle_libs.post_force(g["u_uniform"], g["phi"], g["linf"])
    assert abs(e + 0.5 * np.sum(f * g["u_uniform"])) < 1e-12 * abs(e)


def test_gather_scatter_oracles_agree(oracle_libs):
    O = oracle_libs
    rng = np.random.default_rng(3)
    nx, ny, nu = 6, 5, 2
    d = 3 * nu
    n = nx * ny * nu
    gid = np.array([(ix, iy, iu) for ix in range(nx) for iy in range(ny) for iu in range(nu)],
                   dtype=np.int32)
    perm = rng.permutation(n)
    gid = gid[perm]
    xeq = np.stack([gid[:, 0] + 0.5, gid[:, 1] + 0.5, -gid[:, 2].astype(float)], axis=1)
    x = xeq + rng.uniform(-0.3, 0.3, size=(n, 3))
    x[:, 0] = np.mod(x[:, 0], nx)     # wrapped positions -> minimum image needed
    x[:, 1] = np.mod(x[:, 1], ny)
    mask = np.where(rng.random(n) < 0.9, 3, 1).astype(np.int32)
    for shift in [(0, 0), (2, -1)]:
        u1, n1 = O.gather(x, xeq, gid.copy(), mask, 2, nx, ny, d, float(nx), float(ny), *shift)
        u2, n2, _ = O.c_gather(x, xeq, gid.copy(), mask, 2, nx, ny, d, float(nx), float(ny), *shift)
        assert n1 == n2 and np.array_equal(u1, u2)
        assert np.abs(u1).max() <= 0.3 + 1e-12
    fxy = rng.standard_normal((d, nx * ny))
    f1, s1, k1 = O.scatter(fxy, gid, mask, 2, np.zeros((n, 3)), nx=nx, ny=ny)
    f2, s2, k2 = O.c_scatter(fxy, gid, mask, 2, np.zeros((n, 3)), nx, ny)
    assert k1 == k2 and np.array_equal(f1, f2) and np.allclose(s1, s2, rtol=0, atol=1e-12)


def test_synthetic_sc100_matrices_equal_the_plugin(oracle_libs):
    """gfmd_b200.synthetic.sc100_dynamical_matrices (used to build the bench's sc100 table on
    the device) against the reference plugin's get_dynamical_matrices."""
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    import sys, os
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "user-gfmd_b200"))
    from gfmd_b200 import synthetic
    k = O.RefKernel("sc100 height 128")
    assert k.height() == 128
    ref = k.dynamical_matrices(12, 10, 0, 6)
    mine = synthetic.sc100_dynamical_matrices(12, 10, 0, 6)
    assert np.abs(ref - mine).max() < 1e-15
    k.close()


# ---- off-path services: q-space dump fields and prec_gradient -------

def _aux_names():
    import aux_checks
    return aux_checks.aux_names()


@pytest.mark.parametrize("name", _aux_names())
def test_numpy_oracle_matches_recorded_prec_gradient_and_dumps(name, oracle_libs):
    """oracle.prec_gradient / oracle.spectrum against what the reference's own
    GFMDSolverStatic::prec_gradient and GFMDSolverFFT::dump produced (tests/golden/aux)."""
    import aux_checks
    O = oracle_libs
    g, a = load_golden(name), aux_checks.load_aux(name)
    nx, ny = int(g["nx"]), int(g["ny"])
    gP = O.prec_gradient(g["u_uniform"], g["phi"], a["cavg"], reference_quirk=True)
    assert rel_err(gP, a["gP"]) < 1e-12
    uq, fq = O.spectrum(g["u_uniform"], g["phi"])
    fields = O.dump_fields(uq, fq, nx, ny)
    for k in (k for k in a if k.startswith("dump_")):
        assert rel_err(fields[k[5:]], a[k]) < aux_checks.TOL_TEXT, k


def test_reference_sources_reproduce_recorded_prec_gradient(oracle_libs):
    """The reference's solver sources, when built here, against the recorded fixture and
    against the oracle's statement of the ndof > 3 behaviour (gfmd_misc.h:113-115)."""
    import aux_checks
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    for name in ("small_sc100_16x12", "small_fcc111_8x7"):
        g, a = load_golden(name), aux_checks.load_aux(name)
        s = O.RefSolver(int(g["nx"]), int(g["ny"]), int(g["ndof"]), fft_backend=0)
        s.set_phi(g["phi"], g["linf"])
        gP = s.prec_gradient(a["cavg"], g["u_uniform"])
        assert np.array_equal(gP, a["gP"])
        full = O.prec_gradient(g["u_uniform"], g["phi"], a["cavg"])
        if int(g["ndof"]) == 3:
            assert rel_err(full, gP) < 1e-12
        else:
            assert rel_err(full, gP) > 1e-3
        s.close()


# ---- FFT-free physical anchor -------------------------------------------------------------

@pytest.mark.parametrize("nx,ny,height", [(6, 5, 3), (4, 7, 1), (5, 5, 0)])
def test_explicit_spring_network_pins_sign_convention_and_table(nx, ny, height, oracle_libs):
    """tests/spring_network.py: the explicit harmonic network behind the `sc100` kernel, relaxed
    by a dense solve, against the reference plugin's table pushed through the oracle path.  The
    transposed FFT sign convention must NOT fit (that is what makes this a sign KAT)."""
    import spring_network
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    k = O.RefKernel("sc100 height %d" % height)
    phi = k.phi(nx, ny)
    u0 = np.random.default_rng(4).uniform(-0.1, 0.1, size=(3, nx, ny))
    f, e, _ = O.post_force(u0, phi, np.zeros(1))
    fs = spring_network.surface_force(u0, height)
    assert rel_err(f, fs) < 1e-13
    assert abs(e + 0.5 * float((fs * u0).sum())) <= 1e-13 * abs(e)
    # the same table with e^{+i q r} forward / e^{-i q r} backward transforms
    uq = np.fft.ifft2(u0, axes=(1, 2)) * (nx * ny)
    F = -np.einsum("qij,qj->qi", phi, np.moveaxis(uq.reshape(3, nx * ny), 0, 1))
    f_wrong = np.fft.fft2(np.moveaxis(F, 0, 1).reshape(3, nx, ny), axes=(1, 2)).real
    if height > 0:
        assert rel_err(f_wrong, fs) > 1e-2
    k.close()
