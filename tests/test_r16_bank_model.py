"""The conflict-free claim of the radix-16 row kernels as a test: the bank model
of tools/r16_bank_model.py over every shared-memory access of the three kernels."""
import os
import sys

from conftest import ROOT


def test_radix16_row_kernels_are_bank_conflict_free(capsys):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import r16_bank_model as M
    assert max(M.model_r16(), M.model_r16h(), M.model_r16w()) == 1
