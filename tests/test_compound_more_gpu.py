"""More of the reference's compound acceptance tests on the CUDA path, through the helpers of
tests/test_compound.py, with the stiffness tables taken from the reference plugin at test time
(oracle/_ref).  All three pass on the emulation build too (tests/test_emulated_kernels.py, GFMD_EMU_LONG=1)
and on the GPU (round 2)."""
import pytest

from test_compound import plugin_table, run_energy_conservation, run_hertz_cubic

pytestmark = pytest.mark.gpu


def test_energy_conservation_fcc100(oracle_libs):
    """tests/TEST_energy_conservation_fcc100: single layer, kernel `fcc100 1.0 1` (closed form, no
    transfer matrix), 10 x 10, random displacements 0.1, no initial velocity, same criterion.  The
    table comes from the reference plugin at test time (oracle/_ref)."""
    import torch
    import gfmd_b200
    if not oracle_libs.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    run_energy_conservation(gfmd_b200, torch.device("cuda"), 100000, 1000,
                            table=plugin_table(oracle_libs, "fcc100 1.0 1", 10, 10), vx=0.0, expect_shift=False)


def test_hertz_fcc100_128x128(oracle_libs):
    """tests/TEST_Hertz_fcc100_128x128: kernel `ft fcc100 1 1.0 pair-potential 1 1.0 height 128` on the
    simple-cubic 128 x 128 layer, contact modulus E = 1.39 (eval.py:37), residual < 1e-2."""
    import torch
    import gfmd_b200
    if not oracle_libs.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    run_hertz_cubic(gfmd_b200, torch.device("cuda"),
                    plugin_table(oracle_libs, "ft fcc100 1 1.0 pair-potential 1 1.0 height 128", 128, 128), 1.0, 1.39)


def test_hertz_sc100_128x128_a0_1_3(oracle_libs):
    """tests/TEST_Hertz_sc100_128x128_a0_1.3: lattice constant 1.3 (`ft sc100 1.3 1 ...`; the atoms
    sit at 1.3 i + 0.5), E* = 8/3 / 1.3, residual of the PRESSURE f / a0^2 < 1e-2 (eval.py:38-41, :82-84)."""
    import torch
    import gfmd_b200
    if not oracle_libs.ref_available():
        pytest.skip("oracle/_ref/libgfmd_ref.so not built")
    run_hertz_cubic(gfmd_b200, torch.device("cuda"),
                    plugin_table(oracle_libs, "ft sc100 1.3 1 pair-potential 2 1.0 1.0 height 128", 128, 128),
                    1.3, 8.0 / 3 / 1.3, dmax=0.1)


def test_error_convergence_two_layers_lj():
    """TEST_error_convergence_ft_two_layers_lj_cut restated (tests/errconv.py): a Lennard-Jones fcc(100) crystal on
    a GFMD substrate (two GFMD layers, ndof 6, the reference plugin's `ft fcc100 ... pair-potential 2x ...` table
    with the LJ force constants) against its all-atom twin (golden fixture, plain numpy): the force error on the
    displaced probe atom must go like dstep^2 (exponent >= 1.9) and stay below 10 % at dstep = 0.01."""
    import os
    import numpy as np
    import gfmd_b200
    import errconv
    from conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "compound", "errconv_fcc100_two_layers_lj.npz"))
    goeslike, relerr = errconv.check(gfmd_b200, g)
    print("force error goes like dstep^%.3f; relative errors %s" % (goeslike, relerr))
