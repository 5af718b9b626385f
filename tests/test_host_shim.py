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


def test_host_shim_mirrors_the_reference_interface():
    """Every virtual of GFMDSolver that GFMDSolverStatic overrides is overridden (CPU check
    on the sources; the reference header is only read when it is present)."""
    hdr = open(os.path.join(ROOT, "user-gfmd_b200", "host", "gfmd_solver_b200.h")).read()
    for member in ["set_grid_size(int, int, int)", "set_kernel(StiffnessKernel *, bool normalize = true)",
                   "pre_force(void *, void *)", "post_force(void *, void *, char *)", "memory_usage()",
                   "init()"]:
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
