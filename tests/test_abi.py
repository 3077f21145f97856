"""The C-ABI shared library loads and exports every symbol include/qutip_b200.h declares
(no compute call is made, so this runs without a GPU)."""
import ctypes
import os
import re

import pytest

import qutip_b200
from qutip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "qutip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_all_declared_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert set(_lib.SYMBOLS) == set(syms)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device compute calls must fail loudly, never fall back."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    import numpy as np
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(qutip_b200.QbError):
        qutip_b200.DeviceDense.from_numpy(np.ones(4, dtype=complex))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qutip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), f
