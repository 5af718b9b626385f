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


def test_specialised_kernels_keep_their_register_and_stack_budget():
    """The occupancy the specialised kernels are designed for is a property of
    the built machine code: rows and radix-16 rows 128 registers (two CTAs of 256 threads, or one of
    512, per SM), the 16-warp column kernel 128, the 8-warp one 255, and next to no local memory.
    cuobjdump reads it from the library without a GPU; a change that makes one of them spill shows
    here, not only as a slower bench."""
    import re
    import shutil
    import subprocess
    import gfmd_b200
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(gfmd_b200.__file__))), "libgfmd_b200.so")
    if not os.path.exists(lib):
        pytest.skip("libgfmd_b200.so not built")
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {m.group(1): (int(m.group(2)), int(m.group(3)))
             for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)}
    assert usage, "no resource usage found in " + lib
    budgets = [                                    # (mangled-name fragment, max registers, max stack bytes)
        ("k_cols_fused_p2_lrI", 128, 160), ("k_cols_fused_p2I", 255, 64),
        ("k_cols_top_passI", 64, 160), ("8k_gather", 40, 16), ("9k_scatter", 40, 16),
    ]
    seen = set()
    nrows = 0
    for name, (reg, stack) in usage.items():
        for frag, rmax, smax in budgets:
            if frag in name:
                seen.add(frag)
                assert reg <= rmax and stack <= smax, (name, reg, stack)
        m = re.search(r"k_rows_(?:fwd|inv)_(?:p2|r16h?)ILi(\d+)ELi(\d+)ELi(\d+)E", name) or \
            re.search(r"k_rows_(?:fwd|inv)_r16wILi(\d+)E()Li(\d+)E", name)
        if m:      # <NR, RB, T, ...>: two CTAs per SM up to 256 threads, one beyond; 64 K registers per SM
            nrows += 1
            threads = int(m.group(3))
            ctas = 2 if threads <= 256 else 1
            assert reg * threads * ctas <= 65536 and stack <= 64, (name, reg, stack)
    assert seen == {b[0] for b in budgets}, seen
    assert nrows >= 30
