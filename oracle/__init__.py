"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the hot path (numpy + a small C
library).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; the product package ``qutip_b200``
never does (tests/test_abi.py::test_product_never_imports_oracle greps for it).

Parity status: PINNED.  The restatement is checked against outputs of the reference
itself (the unmodified qutip 5.4.0.dev build in ``oracle/_ref``, see build_ref.py) via
the committed fixtures under ``tests/golden/`` (generator: tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile oracle/spmv_oracle.c into oracle/liboracle.so (gcc, ~1 s)."""
    src = os.path.join(_HERE, "spmv_oracle.c")
    out = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.orc_wrmn_error.restype = ctypes.c_double
    return _LIB


def ref_path():
    """Directory to put on sys.path to import the unmodified reference, or None."""
    p = os.path.join(_HERE, "_ref")
    return p if os.path.exists(os.path.join(p, ".built")) else None
