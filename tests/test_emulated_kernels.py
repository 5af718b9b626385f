"""Kernel LOGIC on machines without a GPU: the unchanged sources of user-gfmd_b200/csrc compiled
for the CPU against the CUDA emulation shim in tests/emu (threads of a block as cooperative
fibers) and driven through the same C ABI and Python binding as on the B200.  This is test
infrastructure -- it says nothing about performance or real data races, the product never loads
it, and the parity claims rest on the `-m gpu` tests; what it buys is that indexing, barriers,
table layouts and host orchestration of every kernel (generic, Bluestein, the specialised
power-of-two kernels, long columns, device table builder, gather/scatter, the off-path services)
are exercised by `pytest -m "not gpu"`."""
import shutil

import numpy as np
import pytest

import aux_checks
import split_checks
from conftest import golden_cases, golden_names, load_golden, rel_err

TOL = 1e-11


@pytest.fixture(scope="module")
def B():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build as emu_build
    import gfmd_b200
    lib = gfmd_b200.load_library(emu_build.build())
    saved = gfmd_b200._lib
    gfmd_b200._lib = lib
    yield gfmd_b200
    gfmd_b200._lib = saved


def random_case(nx, ny, d):
    """Same synthetic Hermitian, conj-symmetric table as tests/test_gpu_parity.py."""
    rng = np.random.default_rng(1000 * nx + ny + d)
    Dr = rng.standard_normal((nx, ny, d, d)) * np.exp(-0.3 * rng.random((nx, ny, 1, 1)) * 10)
    Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
    Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
    phi = np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)
    return phi, rng.standard_normal(d // 3), rng.uniform(-0.1, 0.1, size=(d, nx, ny))


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "C1_sc100_128x128"])
def test_golden_vectors(B, name):
    g = load_golden(name)
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    for c in golden_cases(g):
        u = np.ascontiguousarray(g["u_" + c].reshape(d, nx * ny))
        f = np.full_like(u, np.nan)
        e = s.post_force(u, f)
        assert rel_err(f.reshape(d, nx, ny), g["f_" + c]) < TOL, (name, c)
        assert abs(e - float(g["epot_" + c])) <= TOL * max(abs(float(g["epot_" + c])), 1e-300)
        assert np.abs(s.get_u0() - g["u0_" + c]).max() <= TOL * max(1.0, np.abs(g["u0_" + c]).max())
    s.close()


SIZES = [
    # generic kernels: every radix, odd/even/prime, Bluestein rows and columns, run-time ndof
    (1, 1, 3), (2, 1, 3), (3, 5, 3), (7, 9, 6), (37, 64, 3), (64, 37, 6), (30, 42, 9), (35, 25, 12),
    (11, 13, 15),
    # specialised power-of-two kernels (the grids are thin in the other direction to stay cheap)
    (2048, 4, 3), (4096, 2, 3), (4, 2048, 3), (2, 4096, 3), (2, 8192, 3), (1, 16384, 3),
    # long columns: top-digit pass + 4096-point sub-columns
    (8192, 2, 3), (16384, 1, 3),
]


@pytest.mark.parametrize("nx,ny,d", SIZES)
def test_random_tables_against_oracle(B, nx, ny, d, oracle_libs):
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, u0_ref = O.post_force(u, phi, linf)
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    if max(nx, ny) >= 2048:
        assert "[fast" in s.describe()
    s.set_kernel(phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    f = np.full_like(uu, np.nan)
    e = s.post_force(uu, f)
    assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL
    assert abs(e - e_ref) <= TOL * abs(e_ref)
    assert np.abs(s.get_u0() - u0_ref).max() <= TOL * max(1.0, np.abs(u0_ref).max())
    s.close()


def test_device_table_builder(B):
    """k_build_phi (transfer-matrix recursion) in small column chunks, against the golden
    forces of the plugin's own table -- the chunked call pattern of bench.py."""
    from gfmd_b200 import synthetic
    g = load_golden("small_sc100_16x12")       # kernel: ft sc100 ... height 128 == sc100 height 128
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    for k0 in range(0, s.nky, 3):
        nk = min(3, s.nky - k0)
        s.build_kernel_columns(synthetic.sc100_dynamical_matrices(nx, ny, k0, nk), k0, height=128)
    s.set_linf(g["linf"])
    u = np.ascontiguousarray(g["u_uniform"].reshape(d, nx * ny))
    f = np.full_like(u, np.nan)
    e = s.post_force(u, f)
    assert rel_err(f.reshape(d, nx, ny), g["f_uniform"]) < TOL
    assert abs(e - float(g["epot_uniform"])) <= TOL * abs(float(g["epot_uniform"]))
    s.close()


@pytest.mark.parametrize("kernel,nx,ny", [
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 16", 8, 7),
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 7", 2048, 2),     # specialised table layout
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height -1", 6, 5),      # iterate to convergence
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height -1", 4, 3),
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 5", 4096, 2),       # position-ordered table (k_cols_fft_p2)
])
def test_device_built_table_equals_plugin_table(B, kernel, nx, ny, oracle_libs):
    O = oracle_libs
    if not O.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    k = O.RefKernel(kernel)
    d = k.ndof
    u = np.random.default_rng(9).uniform(-0.1, 0.1, size=(d, nx * ny))
    out = []
    for mode in ("plugin", "device"):
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        if mode == "plugin":
            s.set_kernel(k.phi(nx, ny), k.linf())
        else:
            s.build_kernel_columns(k.dynamical_matrices(nx, ny, 0, s.nky), 0, height=k.height())
            s.set_linf(k.linf())
        f = np.full_like(u, np.nan)
        out.append((f, s.post_force(u, f)))
        s.close()
    assert rel_err(out[1][0], out[0][0]) < TOL
    assert abs(out[1][1] - out[0][1]) <= TOL * abs(out[0][1])
    k.close()


def make_atoms(nx, ny, nu, rng, a0=1.0):
    n = nx * ny * nu
    gid = np.array([(ix, iy, iu) for ix in range(nx) for iy in range(ny) for iu in range(nu)], dtype=np.int32)
    gid = gid[rng.permutation(n)]
    xeq = a0 * np.stack([gid[:, 0] + 0.5, gid[:, 1] + 0.5, -gid[:, 2].astype(float)], axis=1)
    x = xeq + a0 * rng.uniform(-0.3, 0.3, size=(n, 3))
    x[:, 0] = np.mod(x[:, 0], nx * a0)                    # atoms wrapped into the periodic box
    x[:, 1] = np.mod(x[:, 1], ny * a0)
    mask = np.where(rng.random(n) < 0.95, 3, 1).astype(np.int32)
    return x, xeq, gid, mask


@pytest.mark.parametrize("nx,ny,nu,shift,a0", [(6, 5, 2, (0, 0), 1.0), (37, 16, 1, (3, -2), 1.0),
                                               (9, 8, 1, (0, 0), 1.3), (7, 6, 2, (-1, 2), 0.77)])
def test_gather_scatter_against_oracle(B, nx, ny, nu, shift, a0, oracle_libs):
    """Under emulation "device" pointers are host pointers: numpy arrays stand in for them.
    a0 != 1: lattice constant of the reference's TEST_Hertz_sc100_128x128_a0_1.3 (the minimum-image
    wrap works in box lengths, not in grid units)."""
    O = oracle_libs
    rng = np.random.default_rng(11)
    d = 3 * nu
    x, xeq, gid, mask = make_atoms(nx, ny, nu, rng, a0)
    n = x.shape[0]
    g_ref = gid.copy()
    u_ref, n_ref = O.gather(x, xeq, g_ref, mask, 2, nx, ny, d, nx * a0, ny * a0, *shift)
    assert np.abs(u_ref).max() <= 0.3 * a0 + 1e-12            # every displacement was un-wrapped
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    dgid = gid.copy()
    du = np.zeros((d, nx * ny))
    s.gather(x, xeq, dgid, mask, 2, n, nx * a0, ny * a0, shift[0], shift[1], du)
    r = s.results()
    assert r["natoms_gathered"] == n_ref and r["n_out_of_range"] == 0
    assert np.array_equal(du, u_ref)
    assert np.array_equal(dgid, g_ref)
    fxy = rng.standard_normal((d, nx * ny))
    f0 = rng.standard_normal((n, 3))
    nlocal = n - n // 7
    f_ref, fsum_ref, k_ref = O.scatter(fxy, g_ref, mask, 2, f0.copy(), nlocal=nlocal, nx=nx, ny=ny)
    df = f0.copy()
    s.scatter(dgid, mask, 2, n, nlocal, df, fxy)
    r = s.results()
    assert r["natoms_scattered"] == k_ref
    assert np.array_equal(df, f_ref)
    assert np.abs(r["fsum"] - fsum_ref).max() <= 1e-12 * max(1.0, np.abs(fsum_ref).max())
    s.close()


def test_full_step(B, oracle_libs):
    O = oracle_libs
    g = load_golden("small_sc100_16x12")
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    x, xeq, gid, mask = make_atoms(nx, ny, 1, np.random.default_rng(5))
    mask[:] = 3
    n = x.shape[0]
    u_ref, _ = O.gather(x, xeq, gid.copy(), mask, 2, nx, ny, d, float(nx), float(ny))
    f_ref, e_ref, _ = O.post_force(u_ref.reshape(d, nx, ny), g["phi"], g["linf"])
    fa_ref, fsum_ref, _ = O.scatter(f_ref.reshape(d, nx * ny), gid, mask, 2, np.zeros((n, 3)), nx=nx, ny=ny)
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    df = np.zeros((n, 3))
    s.full_step(x, xeq, gid.copy(), mask, 2, n, n, float(nx), float(ny), df)
    r = s.results()
    assert rel_err(df, fa_ref) < TOL
    assert abs(r["epot"] - e_ref) <= TOL * abs(e_ref)
    assert np.abs(r["fsum"] - fsum_ref).max() <= 1e-9
    s.close()


@pytest.mark.parametrize("nx,ny,nu", [(4, 4096, 1), (2, 8192, 1), (2, 4096, 2)])
def test_fused_atom_io_equals_separate_gather_scatter(B, nx, ny, nu, oracle_libs):
    """gfmd_b200_build_cell_map + full_step: the radix-16 row kernels read x / xeq and add the forces
    to the atoms themselves.  Atoms in random order, wrapped into the box (minimum image exercised),
    a few of them ghosts (index >= nlocal): forces on the atoms bit-identical to the separate
    k_gather / k_scatter path, counters equal, force sum to rounding; against the oracle to 1e-11.
    A map with an empty cell or a doubly occupied one is refused (the separate kernels stay)."""
    O = oracle_libs
    d = 3 * nu
    rng = np.random.default_rng(nx * ny + nu)
    x, xeq, gid, mask = make_atoms(nx, ny, nu, rng)
    mask[:] = 3
    n = x.shape[0]
    nlocal = n - n // 9
    phi, linf, _ = random_case(nx, ny, d)
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    assert "k_rows_*_r16" in s.describe(), s.describe()
    s.set_kernel(phi, linf)
    f0 = rng.standard_normal((n, 3))
    out = []
    for fused in (False, True):
        if fused:
            assert s.build_cell_map(gid, mask, 2, n, nlocal) is True
        df = f0.copy()
        l0 = s.launch_count()
        s.full_step(x, xeq, gid, mask, 2, n, nlocal, float(nx), float(ny), df)
        r = s.results()
        out.append((df, r, s.launch_count() - l0))
    (fa, ra, la), (fb, rb, lb) = out
    assert la == lb + 2                                   # no k_gather, k_scatter; k_sum_fsum_io replaces k_sum_partials
    assert np.array_equal(fa, fb)
    assert ra["epot"] == rb["epot"] and np.array_equal(ra["u0"], rb["u0"])
    assert ra["natoms_gathered"] == rb["natoms_gathered"] == n and ra["natoms_scattered"] == rb["natoms_scattered"] == n
    assert np.abs(ra["fsum"] - rb["fsum"]).max() <= 1e-12 * max(1.0, np.abs(ra["fsum"]).max())
    u_ref, _ = O.gather(x, xeq, gid.copy(), mask, 2, nx, ny, d, float(nx), float(ny))
    f_ref, e_ref, _ = O.post_force(u_ref.reshape(d, nx, ny), phi, linf)
    fa_ref, fsum_ref, _ = O.scatter(f_ref.reshape(d, nx * ny), gid, mask, 2, f0.copy(), nlocal=nlocal, nx=nx, ny=ny)
    assert rel_err(fb - f0, fa_ref - f0) < TOL and abs(rb["epot"] - e_ref) <= TOL * abs(e_ref)
    assert np.abs(rb["fsum"] - fsum_ref).max() <= 1e-9 * max(1.0, np.abs(fsum_ref).max())
    # an atom outside the group leaves its cell empty; a second atom in an occupied cell: not usable
    m2 = mask.copy()
    m2[3] = 1
    assert s.build_cell_map(gid, m2, 2, n, nlocal) is False
    g2 = gid.copy()
    g2[5] = g2[6]
    assert s.build_cell_map(g2, mask, 2, n, nlocal) is False
    df = f0.copy()
    s.full_step(x, xeq, gid, mask, 2, n, nlocal, float(nx), float(ny), df)      # separate kernels again
    assert np.array_equal(df, fa)
    s.close()


def test_async_pre_force_and_errors(B):
    g = load_golden("small_sc100_16x12")
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    u = np.ascontiguousarray(g["u_uniform"].reshape(d, nx * ny))
    f = np.zeros_like(u)
    with pytest.raises(B.GFMDError) as ei:          # post_force before set_kernel
        s.post_force(u, f)
    assert ei.value.code == 6
    with pytest.raises(B.GFMDError) as ei:          # so is the preconditioner
        s.prec_gradient(np.eye(d), u, f)
    assert ei.value.code == 6
    bad = g["phi"].copy()
    bad[5, 0, 1] += 0.1                              # break Hermiticity
    with pytest.raises(B.GFMDError) as ei:
        s.set_kernel(bad, g["linf"])
    assert ei.value.code == 7
    s.set_kernel(g["phi"], g["linf"])
    s.pre_force(u, f)
    with pytest.raises(B.GFMDError) as ei:          # off-path services may not cut into a pending step
        s.spectrum(u)
    assert ei.value.code == 6
    e = s.post_force(u, f)
    assert rel_err(f.reshape(d, nx, ny), g["f_uniform"]) < TOL
    assert abs(e - float(g["epot_uniform"])) <= TOL * abs(e)
    s.close()


@pytest.mark.parametrize("name", [n for n in aux_checks.aux_names() if n != "C1_sc100_128x128"])
def test_spectrum_and_dump_fields(B, name, oracle_libs):
    aux_checks.check_spectrum(B, oracle_libs, name)


@pytest.mark.parametrize("name", [n for n in aux_checks.aux_names() if n != "C1_sc100_128x128"])
def test_prec_gradient(B, name, oracle_libs):
    aux_checks.check_prec_gradient(B, oracle_libs, name)


@pytest.mark.parametrize("nx,ny,d", [(4096, 2, 6), (8192, 2, 3), (4096, 2, 3), (2500, 2, 6)])
def test_aux_services_on_column_sets_beyond_one_cta(B, nx, ny, d, oracle_libs):
    """Spectrum and preconditioner through the three-phase column stage (k_cols_split_fft + k_aux_perq) for
    column sets larger than one CTA's shared memory, in every table layout (interleaved incl. long columns,
    position order, plane-major)."""
    if (nx, d) == (4096, 3):
        pytest.skip("3 x 4096 x 16 B fits one CTA: covered by the fused auxiliary kernel")
    aux_checks.check_large_column_sets(B, oracle_libs, nx, ny, d)


@pytest.mark.parametrize("nx,ny,d", [(2048, 2, 3), (4096, 2, 3), (11, 13, 15)])
def test_aux_services_read_the_specialised_table_layout(B, nx, ny, d, oracle_libs):
    """The off-path column kernel reads Phi in the digit-reversed, interleaved layout of the
    specialised per-step kernels (phi_slot); run-time ndof (15) covers its generic branch."""
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(phi, linf)
    uq, fq = s.spectrum(np.ascontiguousarray(u.reshape(d, nx * ny)))
    uq_ref, fq_ref = O.spectrum(u, phi)
    assert rel_err(uq, uq_ref) < TOL and rel_err(fq, fq_ref) < TOL
    if d == 15:
        with pytest.raises(B.GFMDError) as ei:
            s.prec_gradient(np.eye(d), np.zeros((d, nx * ny)), np.zeros((d, nx * ny)))
        assert ei.value.code == 4
    s.close()


@pytest.mark.parametrize("nx,ny,variant", [(2, 4096, 4099), (2, 8192, 8195), (4, 4096, 4097),
                                            (4, 2048, 2053), (2, 4096, 4101), (2, 4096, 4096), (4, 4096, 4103), (2, 4096, 4104), (6, 4096, 4104), (2, 8192, 8200), (4, 8192, 8200), (1, 16384, 16392), (3, 16384, 16392), (2, 8192, 8198), (2, 8192, 8199), (2, 8192, 8197), (3, 16384, 16389)])
def test_row_kernel_variants(B, nx, ny, variant, oracle_libs, monkeypatch):
    """Experimental row-kernel variants (GFMD_B200_ROWS_VARIANT, kernels_fast.cuh): 256-bit
    transposed accesses with their own lane -> wavevector map, other CTA shapes, and the last
    pass fused with the real/complex (un)mixing (ids ny + 5)."""
    O = oracle_libs
    d = 3
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, _ = O.post_force(u, phi, linf)
    monkeypatch.setenv("GFMD_B200_ROWS_VARIANT", str(variant))
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    assert "variant %d" % variant in s.describe(), s.describe()
    s.set_kernel(phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    f = np.full_like(uu, np.nan)
    e = s.post_force(uu, f)
    assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL
    assert abs(e - e_ref) <= TOL * abs(e_ref)
    s.close()



# ndof * nx * 16 B > 227 KB and no specialised kernel: the three-phase column stage
# (kernel_cols_split.cuh) is chosen by plan(); GFMD_B200_COLS_SPLIT=<dofs per CTA> forces it on
# grids that would fit, which also covers dof groups with a ragged last group and Bluestein columns
@pytest.mark.parametrize("nx,ny,d,force", [
    (2048, 2, 12, 0), (4096, 1, 6, 0), (2000, 2, 9, 0), (1331, 3, 12, 0),      # chosen by plan()
    (37, 16, 6, 4), (64, 37, 6, 1), (30, 42, 9, 2), (11, 13, 15, 7), (1100, 3, 3, 2), (1, 1, 3, 1),
    (4, 2048, 6, 4),      # specialised rows + host pipeline (per-dof uploads) around the split column stage
])
def test_split_column_stage(B, nx, ny, d, force, oracle_libs, monkeypatch):
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, u0_ref = O.post_force(u, phi, linf)
    if force:
        monkeypatch.setenv("GFMD_B200_COLS_SPLIT", str(force))
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    assert "k_cols_split_fft" in s.describe(), s.describe()
    s.set_kernel(phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    f = np.full_like(uu, np.nan)
    e = s.post_force(uu, f)
    assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL
    assert abs(e - e_ref) <= TOL * abs(e_ref)
    assert np.abs(s.get_u0() - u0_ref).max() <= TOL * max(1.0, np.abs(u0_ref).max())
    # a second step on the same handle (stale staging data must not matter)
    f2 = np.full_like(uu, np.nan)
    assert s.post_force(uu, f2) == e and np.array_equal(f, f2)
    s.close()


@pytest.mark.parametrize("nx,ny,d", [(4096, 3, 6), (4096, 2, 12), (4096, 4096, 6)][:2])
def test_split_column_stage_on_power_of_two_passes(B, nx, ny, d, oracle_libs, monkeypatch):
    """nx = 4096 with more than one atom per cell: the transform phases of the three-phase column stage run on
    the specialised power-of-two passes (k_cols_fft_p2), the spectrum stays in position order and the table is
    stored likewise (phi_slot mode 2) -- through set_kernel (full table), set_kernel_columns and the device
    table builder.  Against the oracle, and against the same stage on the run-time engine (GFMD_B200_NO_FAST)."""
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, u0_ref = O.post_force(u, phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    res = []
    for how in ("full", "columns", "nofast"):
        if how == "nofast":
            monkeypatch.setenv("GFMD_B200_NO_FAST", "1")
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        assert ("k_cols_fft_p2" in s.describe()) == (how != "nofast"), s.describe()
        if how == "columns":
            p4 = phi.reshape(nx, ny, d, d)
            for k0 in range(s.nky):
                s.set_kernel_columns(np.ascontiguousarray(p4[:, k0:k0 + 1]), k0, normalized=True)
            s.set_linf(linf)
        else:
            s.set_kernel(phi, linf)
        f = np.full_like(uu, np.nan)
        e = s.post_force(uu, f)
        assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL
        assert abs(e - e_ref) <= TOL * abs(e_ref)
        assert np.abs(s.get_u0() - u0_ref).max() <= TOL * max(1.0, np.abs(u0_ref).max())
        f2 = np.full_like(uu, np.nan)
        assert s.post_force(uu, f2) == e and np.array_equal(f, f2)
        res.append(f)
        s.close()
    assert rel_err(res[0], res[1]) < 1e-13         # set_kernel symmetrises the table, set_kernel_columns does not
    assert rel_err(res[0], res[2]) < 1e-13


def test_split_column_stage_equals_fused_bitwise_in_forces(B, monkeypatch):
    """Same transforms and the same per-q arithmetic: forces are bit-identical to the fused kernel's;
    only the grouping of the energy partials differs."""
    nx, ny, d = 48, 20, 6
    phi, linf, u = random_case(nx, ny, d)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    out = []
    for force in (None, "4"):
        if force:
            monkeypatch.setenv("GFMD_B200_COLS_SPLIT", force)
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        s.set_kernel(phi, linf)
        f = np.full_like(uu, np.nan)
        e = s.post_force(uu, f)
        out.append((f, e, s.get_u0().copy(), s.describe()))
        s.close()
    assert "k_cols_fused" in out[0][3] and "k_cols_split_fft" in out[1][3]
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2])
    assert abs(out[0][1] - out[1][1]) <= 1e-14 * abs(out[0][1])


def test_split_column_stage_energy_identity_and_linearity(B, monkeypatch):
    """Small version of the full-size GPU property test (tests/test_split_columns_gpu.py)."""
    monkeypatch.setenv("GFMD_B200_COLS_SPLIT", "4")
    split_checks.energy_identity_and_linearity(B, 96, 40, 6, expect=("k_cols_split_fft",))


def test_sweep_of_small_grids(B, oracle_libs):
    """Every grid up to 24 x 24 (all radix mixes, Bluestein for 11, 13, 17, 19, 23, odd and even
    rows) plus a few long smooth / prime extents, ndof 3 or 6, against the oracle."""
    O = oracle_libs
    rng = np.random.default_rng(0)
    grids = [(nx, ny) for nx in range(1, 25) for ny in range(1, 25)]
    grids += [(n, m) for n in (49, 81, 97, 125, 243, 251, 343, 1021) for m in (1, 6)]
    grids += [(m, n) for n, m in grids[-16:]]
    bad = []
    for nx, ny in grids:
        d = 3 if (nx + ny) % 3 else 6
        Dr = rng.standard_normal((nx, ny, d, d))
        Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
        Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
        phi = np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)
        linf = rng.standard_normal(d // 3)
        u = rng.uniform(-0.1, 0.1, size=(d, nx, ny))
        f_ref, e_ref, _ = O.post_force(u, phi, linf)
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        s.set_kernel(phi, linf)
        uu = np.ascontiguousarray(u.reshape(d, nx * ny))
        f = np.full_like(uu, np.nan)
        e = s.post_force(uu, f)
        s.close()
        if not (rel_err(f.reshape(d, nx, ny), f_ref) < TOL and abs(e - e_ref) <= TOL * abs(e_ref)):
            bad.append((nx, ny, d))
    assert not bad, bad


@pytest.mark.parametrize("nx,ny,d", [(4, 2048, 3), (2, 4096, 6)])
def test_host_pipeline_is_bit_identical(B, nx, ny, d, oracle_libs, monkeypatch):
    """GFMD_B200_HOST_PIPE=1: u uploaded and f downloaded dof by dof around per-dof row kernels
    (enqueue_solver_hostpipe) -- same kernels, so the same bits as the plain host path; also via
    pre_force.  (Streams are synchronous here; the event ordering is checked on the GPU.)"""
    O = oracle_libs
    phi, linf, u = random_case(nx, ny, d)
    f_ref, e_ref, u0_ref = O.post_force(u, phi, linf)
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    out = {}
    for pipe in ("0", "1"):
        monkeypatch.setenv("GFMD_B200_HOST_PIPE", pipe)
        s = B.GFMDSolverB200()
        s.set_grid_size(nx, ny, d)
        s.set_kernel(phi, linf)
        f = np.full_like(uu, np.nan)
        e = s.post_force(uu, f)
        f2 = np.full_like(uu, np.nan)
        s.pre_force(uu, f2)
        e2 = s.post_force(uu, f2)
        assert np.array_equal(f, f2) and e == e2
        out[pipe] = (f, e, s.get_u0().copy())
        s.close()
    assert np.array_equal(out["0"][0], out["1"][0]) and out["0"][1] == out["1"][1]
    assert np.array_equal(out["0"][2], out["1"][2])
    assert rel_err(out["1"][0].reshape(d, nx, ny), f_ref) < TOL
    assert abs(out["1"][1] - e_ref) <= TOL * abs(e_ref)


def test_energy_conservation_two_layers_short(B):
    """The reference's TEST_energy_conservation_two_layers criterion (tests/test_compound.py runs
    all 100 000 steps on the GPU) on the first 1500 velocity-Verlet steps: gather -> solver ->
    scatter every step, lattice re-indexing included; torch CPU tensors stand in for device
    memory.  max|E - <E>| / <E> <= 1e-4."""
    import torch
    import test_compound
    test_compound.run_energy_conservation_two_layers(B, torch.device("cpu"), 1500, 100)


def test_energy_conservation_fcc100_short(B, oracle_libs):
    """TEST_energy_conservation_fcc100 (single layer, closed-form `fcc100 1.0 1` kernel, table from the
    reference plugin): first 500 steps; the full 100 000 emulated steps pass too (19 minutes)."""
    import torch
    import test_compound
    if not oracle_libs.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    test_compound.run_energy_conservation(B, torch.device("cpu"), 500, 50,
                                          table=test_compound.plugin_table(oracle_libs, "fcc100 1.0 1", 10, 10),
                                          vx=0.0, expect_shift=False)


def test_random_call_sequences(B, oracle_libs):
    """Host logic under interleaving: several live handles of different grid sizes (generic and
    specialised kernels), random sequences of post_force / pre_force + post_force / device step /
    spectrum / prec_gradient / new tables (whole or column blocks in random order) / new linf /
    pipeline toggles / close, every result checked against the oracle."""
    import random
    O = oracle_libs
    rnd = random.Random(20261017)

    def table(nx, ny, d, rng):
        Dr = rng.standard_normal((nx, ny, d, d))
        Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
        Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
        Dr[0, 0] += 4 * d * np.eye(d)                      # well-conditioned Phi + C
        return np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)

    shapes = [(8, 7, 6), (16, 12, 3), (5, 9, 3), (37, 8, 3), (6, 2048, 3), (2048, 2, 3), (12, 10, 9), (10, 10, 12)]
    ops = ["new", "new", "step", "step", "step", "pre", "spec", "prec", "retable", "linf", "close", "pipe", "devstep"]
    handles = []
    for _ in range(400):
        op = rnd.choice(ops)
        if op == "new" or not handles:
            if len(handles) > 3:
                continue
            nx, ny, d = rnd.choice(shapes)
            rng = np.random.default_rng(rnd.randrange(1 << 30))
            s = B.GFMDSolverB200()
            s.set_grid_size(nx, ny, d)
            h = dict(s=s, nx=nx, ny=ny, d=d, phi=table(nx, ny, d, rng), linf=rng.standard_normal(d // 3), rng=rng)
            s.set_kernel(h["phi"], h["linf"])
            handles.append(h)
            continue
        h = rnd.choice(handles)
        s, nx, ny, d, rng = h["s"], h["nx"], h["ny"], h["d"], h["rng"]
        u = rng.uniform(-0.1, 0.1, size=(d, nx, ny))
        uu = np.ascontiguousarray(u.reshape(d, nx * ny))
        if op in ("step", "pre", "devstep"):
            f = np.full_like(uu, np.nan)
            if op == "pre":
                s.pre_force(uu, f)
            if op == "devstep":
                s.post_force_device(uu.copy(), f)
                r = s.results()
                e, u0 = r["epot"], r["u0"]
            else:
                e, u0 = s.post_force(uu, f), s.get_u0()
            fr, er, u0r = O.post_force(u, h["phi"], h["linf"])
            assert rel_err(f.reshape(d, nx, ny), fr) < TOL, (op, nx, ny, d)
            assert abs(e - er) <= TOL * abs(er), (op, nx, ny, d)
            assert np.abs(u0 - u0r).max() <= TOL * max(1.0, np.abs(u0r).max())
        elif op == "spec":
            uq, fq = s.spectrum(uu)
            uqr, fqr = O.spectrum(u, h["phi"])
            assert rel_err(uq, uqr) < TOL and rel_err(fq, fqr) < TOL
        elif op == "prec" and d in (3, 6, 9, 12):
            cavg = (2.0 * np.eye(d) + 0.1 * rng.standard_normal((d, d))) / (nx * ny)
            quirk = rnd.random() < 0.5
            gP = np.full_like(uu, np.nan)
            s.prec_gradient(cavg, uu, gP, reference_quirk=quirk)
            ref = O.prec_gradient(u, h["phi"], cavg, reference_quirk=quirk)
            assert rel_err(gP.reshape(d, nx, ny), ref) < 10 * TOL, (nx, ny, d)
        elif op == "retable":
            h["phi"], h["linf"] = table(nx, ny, d, rng), rng.standard_normal(d // 3)
            if rnd.random() < 0.5:
                s.set_kernel(h["phi"], h["linf"])
            else:
                P = h["phi"].reshape(nx, ny, d, d)
                blocks = list(range(0, s.nky, 3))
                rnd.shuffle(blocks)
                for k0 in blocks:
                    s.set_kernel_columns(np.ascontiguousarray(P[:, k0:k0 + min(3, s.nky - k0)]), k0)
                s.set_linf(h["linf"])
        elif op == "linf":
            h["linf"] = rng.standard_normal(d // 3)
            s.set_linf(h["linf"])
        elif op == "pipe":
            s.host_pipeline(rnd.random() < 0.5)
        elif op == "close":
            s.close()
            handles.remove(h)
    for h in handles:
        h["s"].close()


def test_error_convergence_two_layers_lj(B):
    """The reference's error-convergence criterion (tests/TEST_error_convergence_ft_two_layers_lj_cut, eval.py:
    207-226) on the emulated CUDA path: GFMD substrate under a Lennard-Jones crystal against the all-atom twin
    (golden fixture); the force error on the probe atom must shrink like the square of its displacement.
    See tests/errconv.py."""
    import errconv
    from conftest import GOLDEN_DIR
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "compound", "errconv_fcc100_two_layers_lj.npz"))
    goeslike, relerr = errconv.check(B, g)
    print("force error goes like dstep^%.3f; relative errors %s" % (goeslike, relerr))


@pytest.mark.parametrize("order", ["reverse", "random:7"])
def test_results_do_not_depend_on_thread_order(order):
    """CUDA promises no execution order between barriers.  The emulation runs the threads of a
    block in thread order by default; GFMD_EMU_SCHED re-runs them in reverse / a random
    permutation (tests/emu/emu_runtime.cpp), which turns a missing __syncthreads / __syncwarp
    into a wrong result.  The order is fixed per process, hence the sub-process."""
    import os
    import subprocess
    import sys
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_EMU_SCHED=order)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "random_tables or (row_kernel_variants and (4103 or 4104 or 8199 or 8200 or 16389 or 16392)) "
                              "or split_column_stage or device_table_builder or prec_gradient or fused_atom_io "
                              "or aux_services_on_column_sets"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nx,ny,height", [(6, 5, 3), (4, 7, 1)])
def test_explicit_spring_network(B, nx, ny, height):
    """Reference-independent physical anchor of the whole CUDA path (tests/spring_network.py):
    closed-form per-q matrices -> transfer-matrix recursion on the "device" -> FFT, contraction,
    inverse FFT, against an explicit spring network relaxed by a dense solve."""
    import spring_network
    from gfmd_b200 import synthetic
    s = B.GFMDSolverB200()
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


def test_pipelined_column_kernel_variant():
    """Experimental k_cols_fused_p2_lr<..., PIPE> (compiled with -DGFMD_EXPERIMENTAL_COLS_PIPE, which
    the emulation build always sets; GFMD_B200_COLS_PIPE=1 selects it, read once per process):
    the last backward pass of a column fused with pass 0 of the next one.  Same tests as the
    default kernel, in a sub-process, under a random thread order."""
    import os
    import subprocess
    import sys
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    env = dict(os.environ, GFMD_B200_COLS_PIPE="1", GFMD_EMU_SCHED="random:11")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "random_tables and (4096 or 8192 or 16384)"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("which", ["hertz_fcc111_64x37", "hertz_sc100_128x128", "hertz_fcc100_128x128",
                                   "hertz_sc100_128x128_a0_1_3"])
def test_hertz_contact_long(B, which, oracle_libs):
    """The Hertz acceptance tests of tests/test_compound.py on the emulation build: thousands of FIRE
    steps, 17 to 28 minutes each on one core (all four pass), so only with GFMD_EMU_LONG=1."""
    import os
    if not os.environ.get("GFMD_EMU_LONG"):
        pytest.skip("long: set GFMD_EMU_LONG=1")
    import torch
    import test_compound
    dev = torch.device("cpu")
    if which == "hertz_fcc100_128x128":
        test_compound.run_hertz_cubic(B, dev, test_compound.plugin_table(
            oracle_libs, "ft fcc100 1 1.0 pair-potential 1 1.0 height 128", 128, 128), 1.0, 1.39)
    elif which == "hertz_sc100_128x128_a0_1_3":
        test_compound.run_hertz_cubic(B, dev, test_compound.plugin_table(
            oracle_libs, "ft sc100 1.3 1 pair-potential 2 1.0 1.0 height 128", 128, 128), 1.3, 8.0 / 3 / 1.3, dmax=0.1)
    else:
        getattr(test_compound, "run_" + which)(B, dev)
