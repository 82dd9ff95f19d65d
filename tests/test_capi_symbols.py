"""
The C-ABI library loads (no GPU needed for that) and exports every symbol that
include/reheatfunq_b200.h declares; without a device, calls fail loudly instead
of falling back to the CPU.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "reheatfunq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rhfq_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    lib = g.build_cuda()
    dll = C.CDLL(lib)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(dll, s), "missing symbol " + s


def test_no_cpu_fallback():
    if has_gpu():
        pytest.skip("GPU present")
    from reheatfunq_b200 import PosteriorBatch
    q = np.array([60.0, 61.0, 70.0])
    c = np.array([1e-9, 3e-9, 2e-9])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        PosteriorBatch([[(q, c)]], [[1.0]], (1.0, 0.0, 0.0, 0.0))


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ or the host emulation."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "reheatfunq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle|liborc|libengine_emu|libhostmath", txt, re.M):
                    bad.append(f)
    assert not bad, bad
