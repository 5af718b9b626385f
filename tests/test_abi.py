"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU,
exports every symbol include/gfmd_b200.h declares, and fails loudly (no CPU
fallback) when no device is usable."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "gfmd_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gfmd_b200_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import gfmd_b200
    if not os.path.exists(gfmd_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return gfmd_b200.load_library()


def test_header_declares_the_expected_boundary():
    syms = declared_symbols()
    for s in ["gfmd_b200_create", "gfmd_b200_set_phi", "gfmd_b200_post_force_host",
              "gfmd_b200_pre_force_async_host", "gfmd_b200_gather", "gfmd_b200_scatter",
              "gfmd_b200_last_error", "gfmd_b200_destroy", "gfmd_b200_comm_init"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    import gfmd_b200
    for s in declared_symbols():
        assert hasattr(lib, s), "libgfmd_b200.so does not export " + s
        assert s in gfmd_b200.ABI, "python binding does not bind " + s
    assert b"sm_100a" in lib.gfmd_b200_version()


def test_no_torch_or_cpp_types_in_header():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in txt.lower()
    assert "std::" not in txt and "at::" not in txt and "&" not in txt


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gfmd_b200
    s = gfmd_b200.GFMDSolverB200()
    with pytest.raises(gfmd_b200.GFMDError) as ei:
        s.set_grid_size(16, 16, 3)
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_argument_validation_needs_no_gpu(lib):
    import gfmd_b200
    h = ctypes.c_void_p()
    assert lib.gfmd_b200_create(ctypes.byref(h), 0, 8, 3, 0) == 1
    assert lib.gfmd_b200_create(ctypes.byref(h), 8, 8, 4, 0) == 1       # ndof % 3
    assert lib.gfmd_b200_create(ctypes.byref(h), 8, 8, 27, 0) == 1      # > MAX_NDOF
    assert lib.gfmd_b200_create_slab(ctypes.byref(h), 8, 8, 3, 0, 2, 2) == 1
    assert lib.gfmd_b200_create_slab(ctypes.byref(h), 9, 8, 3, 0, 0, 2) == 4   # nx % nranks
    assert b"divisible" in lib.gfmd_b200_last_error(None)
    with pytest.raises(gfmd_b200.GFMDError):
        gfmd_b200.gfmd_solver_factory("static/cuda")


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (CPU fallback)."""
    for dp, _, files in os.walk(os.path.join(ROOT, "user-gfmd_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\".*oracle", src, flags=re.M), fn
                assert "gfmd_oracle" not in src and "libgfmd_ref" not in src, fn


def test_integration_doc_covers_every_entry_point():
    """INTEGRATION.md maps every declared entry point to the reference interface it replaces."""
    import re
    hdr = open(os.path.join(ROOT, "include", "gfmd_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(gfmd_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(syms) >= 38
    assert [s for s in syms if s not in doc] == []
