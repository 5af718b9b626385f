"""The C++ host glue (user-gfmd_b200/host/gfmd_solver_b200.cpp -- what a LAMMPS build
compiles) against the reference's GFMDSolverStatic, both driven through the
reference's own GFMDSolver plugin interface with the same StiffnessKernel object
(oracle/hostshim_driver.cpp)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "oracle", "_ref", "libgfmd_hostshim_test.so")
LIB_EMU = os.path.join(ROOT, "oracle", "_ref", "libgfmd_hostshim_emu_test.so")

KERNELS = [
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128", 16, 12, 0),
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 128", 64, 37, 1),
    ("ft fcc100 1.0 2 pair-potential 2 1.0 -0.1 height 10", 10, 10, 1),
]


def test_host_shim_mirrors_the_reference_interface():
    """Every virtual of GFMDSolver that GFMDSolverStatic overrides is overridden (CPU check
    on the sources; the reference header is only read when it is present)."""
    hdr = open(os.path.join(ROOT, "user-gfmd_b200", "host", "gfmd_solver_b200.h")).read()
    for member in ["set_grid_size(int, int, int)", "set_kernel(StiffnessKernel *, bool normalize = true)",
                   "pre_force(void *, void *)", "post_force(void *, void *, char *)", "memory_usage()",
                   "init()", "prec_gradient(double *, double **, double **)", "dump_stiffness()",
                   "dump_greens_function()"]:
        assert member in hdr, member
    src = open(os.path.join(ROOT, "user-gfmd_b200", "host", "gfmd_solver_b200.cpp")).read()
    assert 'strcpy(name, "static/b200")' in src          # factory name check, gfmd_solver.cpp:244-253
    assert "error->one(FLERR" in src                       # the reference's only error path
    assert "fill_phi_buffer(" in src and "get_force_at_gamma_point" in src


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,nx,ny,pre", [
    ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128", 16, 12, 0),
    ("ft fcc111 1 1.0 pair-potential 1 1.0 height 128", 64, 37, 1),
    ("ft fcc100 1.0 2 pair-potential 2 1.0 -0.1 height 10", 10, 10, 1),
    ("sc100 height 16", 128, 128, 0),
])
def test_b200_solver_equals_reference_solver_through_plugin_interface(kernel, nx, ny, pre):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libgfmd_hostshim_test.so not built (needs /root/reference at build time)")
    lib = ctypes.CDLL(LIB)
    lib.hostshim_compare.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_double)]
    out = (ctypes.c_double * 4)()
    rc = lib.hostshim_compare(kernel.encode(), nx, ny, 1234, pre, out)
    assert rc == 0
    assert out[0] < 1e-11 and out[1] < 1e-11 and out[2] < 1e-11, list(out)


def _compare(lib_path, kernel, nx, ny, pre):
    lib = ctypes.CDLL(lib_path)
    lib.hostshim_compare.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_double)]
    out = (ctypes.c_double * 4)()
    rc = lib.hostshim_compare(kernel.encode(), nx, ny, 1234, pre, out)
    assert rc == 0
    assert out[0] < 1e-11 and out[1] < 1e-11 and out[2] < 1e-11, list(out)


def _compare_aux(lib_path, kernel, nx, ny, tmp_path):
    """dump_stiffness, dump_greens_function, a `dumpq_every` step and prec_gradient: the same
    files with the same numbers from both solvers (the reference prints 7 / 11 significant
    digits; entries may differ in the last printed digit)."""
    lib = ctypes.CDLL(lib_path)
    lib.hostshim_compare_aux.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                         ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double,
                                         ctypes.POINTER(ctypes.c_double)]
    a, b = tmp_path / "ref", tmp_path / "b200"
    a.mkdir()
    b.mkdir()
    out = (ctypes.c_double * 4)()
    rc = lib.hostshim_compare_aux(kernel.encode(), nx, ny, 77, str(a).encode(), str(b).encode(), 1.0, out)
    assert rc == 0
    assert out[0] < 1e-11 and out[1] < 1e-11, list(out)
    fa, fb = sorted(p.name for p in a.iterdir()), sorted(p.name for p in b.iterdir())
    ndof = 3 if "sc100" in kernel else 6
    assert fa == fb and len(fa) == 4 * ndof + 3 + 2 * (2 * ndof * ndof + 2)
    for n in fa:
        ta, tb = (a / n).read_text(), (b / n).read_text()
        if ta == tb:
            continue
        xa, xb = np.loadtxt(a / n, ndmin=2), np.loadtxt(b / n, ndmin=2)
        assert xa.shape == xb.shape, n
        tol = 2e-10 if n.startswith("dump.q.") else 2e-6          # " %20.10e " / " %e "
        assert np.abs(xa - xb).max() <= tol * max(np.abs(xa).max(), 1e-300), n


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,nx,ny,pre", KERNELS)
def test_dumps_and_prec_gradient_equal_reference_through_plugin_interface(kernel, nx, ny, pre, tmp_path):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libgfmd_hostshim_test.so not built (needs /root/reference at build time)")
    _compare_aux(LIB, kernel, nx, ny, tmp_path)


def _emu_shim():
    """The same host glue objects linked against the CUDA-on-CPU emulation build of the
    library (tests/emu): test infrastructure for machines without a GPU."""
    if not os.path.isdir("/root/reference/src") and not os.path.exists(LIB_EMU):
        pytest.skip("needs the reference sources (build container only)")
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build as emu_build
    emu_build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "hostshim_emu"])
    return LIB_EMU


@pytest.mark.parametrize("kernel,nx,ny,pre", KERNELS)
def test_host_glue_on_the_emulated_library(kernel, nx, ny, pre, tmp_path):
    lib = _emu_shim()
    _compare(lib, kernel, nx, ny, pre)
    _compare_aux(lib, kernel, nx, ny, tmp_path)
