"""The scipy-built synthetic systems equal what the reference's constructors build."""
import sys

import numpy as np
import pytest

import oracle
from qutip_b200 import models


@pytest.mark.ref
def test_tfim_matches_reference_constructors():
    ref = oracle.ref_path()
    if ref is None:
        pytest.skip("reference build not present")
    sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore")
    import qutip
    n = 4
    H, c_ops, sz = models.tfim(n)
    sx_, sz_, sm_ = [], [], []
    for i in range(n):
        ops = [qutip.qeye(2)] * n
        ops[i] = qutip.sigmax(); sx_.append(qutip.tensor(ops))
        ops[i] = qutip.sigmaz(); sz_.append(qutip.tensor(ops))
        ops[i] = qutip.sigmam(); sm_.append(qutip.tensor(ops))
    Hq = 0
    for i in range(n - 1):
        Hq = Hq - sz_[i] * sz_[i + 1]
    for i in range(n):
        Hq = Hq - sx_[i]
    cq = [np.sqrt(0.1) * s for s in sm_]
    assert np.abs(H.toarray() - Hq.full()).max() == 0
    Lq = qutip.liouvillian(Hq, cq)
    L = models.liouvillian(H, c_ops)
    assert np.abs(L.toarray() - Lq.full()).max() < 1e-15
    solver = qutip.MCSolver(Hq, cq, options={"progress_bar": False})
    heff_q = sum(e.full() for e in solver.rhs().to_list())
    assert np.abs(models.heff(H, c_ops).toarray() - heff_q).max() < 1e-15


def test_program_combinators_used_by_the_matrix_free_binding():
    """-i f(t), +i conj(f(t)) and |g(t)|^2 as built by solve.lindblad_matrix_free /
    plugin.matrix_free_system, evaluated with the python restatement of the byte-code."""
    import cmath
    from qutip_b200 import coeffs
    f = coeffs.compile_expr("A * exp(1j * w * t) + 0.5", {"A": 0.3, "w": 2.0})
    for t in (0.0, 0.7, 3.1):
        val = 0.3 * cmath.exp(2j * t) + 0.5
        assert coeffs.evaluate(f, t) == pytest.approx(val)
        assert coeffs.evaluate(f.scaled(-1j), t) == pytest.approx(-1j * val)
        assert coeffs.evaluate(f.conj().scaled(1j), t) == pytest.approx(1j * val.conjugate())
        assert coeffs.evaluate(f.norm(), t) == pytest.approx(abs(val) ** 2)
        assert coeffs.evaluate(f * f.conj(), t) == pytest.approx(abs(val) ** 2)
        assert coeffs.evaluate(f + coeffs.constant(2 - 1j), t) == pytest.approx(val + 2 - 1j)
