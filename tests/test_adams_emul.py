"""No-GPU checks of the device-resident Adams integrator (qutip_b200/csrc/qb_adams.h + the
QL_AD_* states of qb_control.h) through the host emulator: the generated Adams-Moulton
coefficients against their textbook values, and whole mesolve / mcsolve runs against the
reference fixtures at the accuracy the reference asks of its own Adams method
(tests/solver/test_integrator.py:71-98, test_mesolve.py:110-123: 5e-5 / 1e-5)."""
import ctypes as C
from fractions import Fraction as F

import numpy as np
import pytest

from _emul import FMT_CSR, FMT_DIAM, FMT_SELL, EmulSystem, default_options, lib
from _golden import coeff_spec, load, op_arrays
from _systems import functional_of, merged_constant_rhs
from qutip_b200 import coeffs

ADAMS = 3


def _sp_arrays(m):
    import scipy.sparse as sp
    m = sp.csr_matrix(m)
    m.sort_indices()
    return ("csr", m.shape, dict(data=m.data, col=m.indices, rowptr=m.indptr))


def _table(nq):
    el = (C.c_double * 13)()
    tq = (C.c_double * 3)()
    lib().emul_adams_table(nq, el, tq)
    return list(el)[:nq + 1], list(tq)


def test_adams_moulton_coefficients():
    # Nordsieck corrector vectors of the Adams-Moulton methods (Gear 1971, table 9.2)
    known = {
        1: [1, 1],
        2: [F(1, 2), 1, F(1, 2)],
        3: [F(5, 12), 1, F(3, 4), F(1, 6)],
        4: [F(3, 8), 1, F(11, 12), F(1, 3), F(1, 24)],
        5: [F(251, 720), 1, F(25, 24), F(35, 72), F(5, 48), F(1, 120)],
        6: [F(95, 288), 1, F(137, 120), F(5, 8), F(17, 96), F(1, 40), F(1, 720)],
    }
    for nq, l in known.items():
        el, tq = _table(nq)
        np.testing.assert_allclose(el, [float(x) for x in l], rtol=1e-14)
    # error constants C_{q+1} of Adams-Moulton order q: 1/2, 1/12, 1/24, 19/720, 3/160
    for nq, c in {1: F(1, 2), 2: F(1, 12), 3: F(1, 24), 4: F(19, 720), 5: F(3, 160)}.items():
        assert _table(nq)[1][1] == pytest.approx(1.0 / float(c), rel=1e-13)
    # the higher orders stay consistent: l_1 = 1, l_q = 1/q!, sum_j (-1)^j l_j = 0 for q >= 2
    for nq in range(2, 13):
        el, tq = _table(nq)
        assert el[1] == 1.0
        fact = 1.0
        for i in range(2, nq + 1):
            fact *= i
        assert el[nq] == pytest.approx(1.0 / fact, rel=1e-12)
        assert abs(sum((-1) ** j * x for j, x in enumerate(el))) < 1e-12
        assert tq[1] > 0 and (nq == 12 or tq[2] > 0) and tq[0] > 0


@pytest.mark.parametrize("name", ["c1_jc", "c2_tfim4", "c4_driven", "c5_kerr_0"])
@pytest.mark.parametrize("fmt", [FMT_DIAM, FMT_SELL])
def test_mesolve_adams_vs_reference_fixtures(name, fmt):
    g = load(name)
    s = EmulSystem(len(g["y0"]), 0, fmt)
    for i in range(int(g["n_elements"])):
        spec = coeff_spec(g["el%d_coeff" % i])
        prog = coeffs.compile_expr(spec[0], spec[1]) if spec is not None else None
        s.add_element(*op_arrays(g, "el%d" % i), prog=prog)
    for i in range(int(g["n_eops"])):
        s.add_eop(*_sp_arrays(functional_of(g["eop%d" % i])))
    s.set_functional(1)
    r = s.run(0, ADAMS, g["y0"], g["tlist"], opt=default_options(store_states=1, nsteps=2500))
    assert r["status"][0] == 1
    scale = max(1.0, np.abs(g["states_vern7"]).max())
    assert np.abs(r["states"][0] - g["states_vern7"]).max() < 5e-5 * scale
    escale = max(1.0, np.abs(g["expect_vern7"]).max())
    assert np.abs(r["expect"][0] - g["expect_vern7"]).max() < 5e-5 * escale
    nrhs, nacc, nrej, npass = r["stats"][0]
    assert nacc > 10 and nrej < nacc


def test_adams_tightening_tolerances_converges():
    g = load("c1_jc")
    errs = []
    for rtol in (1e-4, 1e-6, 1e-8):
        s = EmulSystem(len(g["y0"]), 0, FMT_CSR)
        for i in range(int(g["n_elements"])):
            s.add_element(*op_arrays(g, "el%d" % i))
        r = s.run(0, ADAMS, g["y0"], g["tlist"],
                  opt=default_options(store_states=1, nsteps=100000, rtol=rtol, atol=rtol * 1e-2))
        assert r["status"][0] == 1
        errs.append(np.abs(r["states"][0] - g["states_vern7"]).max())
    assert errs[0] > errs[1] > errs[2]
    assert errs[2] < 1e-6


def test_adams_nsteps_limit_and_first_step():
    g = load("c1_jc")
    s = EmulSystem(len(g["y0"]), 0, FMT_CSR)
    for i in range(int(g["n_elements"])):
        s.add_element(*op_arrays(g, "el%d" % i))
    r = s.run(0, ADAMS, g["y0"], g["tlist"], opt=default_options(nsteps=5))
    assert r["status"][0] == -1                       # too much work
    r = s.run(0, ADAMS, g["y0"], g["tlist"], opt=default_options(nsteps=2500, first_step=1e-3,
                                                               max_step=0.05, store_states=1))
    assert r["status"][0] == 1
    assert np.abs(r["states"][0] - g["states_vern7"]).max() < 2e-5


@pytest.mark.parametrize("name,nslots", [("c3_tfim6_mc", 24), ("c3_tfim4_mc_strong", 5)])
def test_mcsolve_adams_vs_reference_fixture(name, nslots):
    """Same thresholds as the reference run: the jump sequence must be the reference's
    (collapse operators identical, collapse times to the integration accuracy)."""
    g = load(name)
    s = EmulSystem(len(g["psi0"]), 0, FMT_DIAM)
    s.add_element(*_sp_arrays(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
    for i in range(int(g["n_eops"])):
        s.add_eop(*op_arrays(g, "eop%d" % i))
    ntraj = int(g["ntraj"])
    r = s.run(1, ADAMS, g["psi0"], g["tlist"], ntraj=ntraj, nslots=nslots, draws=g["draws"],
              opt=default_options(store_states=1, nsteps=2500))
    assert (r["status"] == 1).all()
    assert np.array_equal(r["ncol"], g["col_count"])
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        n = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r["col_t"][j, :n], g["col_times"][cc[j]:cc[j + 1]],
                                   rtol=0, atol=1e-3)    # root finder stops at norm_tol = 1e-4
    assert np.abs(np.transpose(r["expect"], (1, 0, 2)) - g["runs_expect"]).max() < 5e-4


def _adams_system(g, fmt=FMT_DIAM):
    s = EmulSystem(len(g["y0"]), 0, fmt)
    for i in range(int(g["n_elements"])):
        spec = coeff_spec(g["el%d_coeff" % i])
        prog = coeffs.compile_expr(spec[0], spec[1]) if spec is not None else None
        s.add_element(*op_arrays(g, "el%d" % i), prog=prog)
    for i in range(int(g["n_eops"])):
        s.add_eop(*_sp_arrays(functional_of(g["eop%d" % i])))
    s.set_functional(1)
    return s


@pytest.mark.parametrize("name", ["c1_jc", "c2_tfim4", "c4_driven"])
def test_adams_pinned_against_zvode_golden(name):
    """tests/golden/adams_zvode.npz holds the reference's method='adams' (SciPy zvode) runs
    (make_golden.py adams).  Step-level parity with zvode is out of reach (compiled SciPy,
    not in the reference tree), so the device Adams method is pinned against zvode's
    solutions: (1) at tight tolerances both agree within the north-star 1e-8 / 1e-6;
    (2) at the default tolerances its distance to the converged solution is of the size of
    zvode's own, with a comparable number of RHS evaluations."""
    z = load("adams_zvode")
    g = load(name)
    tight = z[name + "_tight_expect"]
    s = _adams_system(g)
    r = s.run(0, ADAMS, g["y0"], g["tlist"],
              opt=default_options(store_states=1, nsteps=100000, atol=1e-12, rtol=1e-10))
    assert r["status"][0] == 1
    np.testing.assert_allclose(r["expect"][0], tight, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(r["states"][0][-1], z[name + "_tight_final"], rtol=1e-6, atol=1e-8)
    r = s.run(0, ADAMS, g["y0"], g["tlist"], opt=default_options(store_states=1, nsteps=2500))
    assert r["status"][0] == 1
    zerr = max(np.abs(z[name + "_default_expect"] - tight).max(), 1e-7)
    assert np.abs(r["expect"][0] - tight).max() < 4 * zerr
    nfe_z = int(z[name + "_default_nst_nfe"][1])
    assert r["stats"][0][0] < 1.5 * nfe_z + 20
