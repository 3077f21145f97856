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


def test_options_struct_matches_header_and_defaults():
    """qb_options as declared in the header, the ctypes mirrors (package and test emulator)
    and the defaults of the reference (qutip_integrator.py:51-59, mcsolve.py:460-465)."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    text = open(os.path.join(ROOT, "include", "qutip_b200.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} qb_options;", text).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    from engine_fields import package_fields, emulator_fields
    assert names == package_fields() == emulator_fields()
    from qutip_b200.engine import make_options
    o = make_options()
    assert (o.atol, o.rtol, o.nsteps, o.first_step, o.min_step, o.max_step) == (1e-8, 1e-6, 1000, 0, 0, 0)
    assert (o.interpolate, o.norm_steps, o.norm_t_tol, o.norm_tol, o.norm_min_step, o.mc_corr_eps) == \
        (1, 25, 1e-6, 1e-4, 0.1, 1e-10)
    assert o.max_order == 0 and o.jump_prob_floor == 0.0 and o.no_jump == 0
    with pytest.raises(KeyError):
        make_options(bogus=1)
