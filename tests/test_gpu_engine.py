"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
fixtures generated from the reference.  Tolerances are north_star's: expectation values
and states within atol 1e-8 / rtol 1e-6; jump counts and collapse indices bit-exact."""
import numpy as np
import pytest

import qutip_b200 as qb
from qutip_b200 import engine as E
from _golden import load, op_arrays, orc_op, orc_rhs
from _systems import dev_op, mc_system_from_golden, me_system_from_golden
from oracle.rk_oracle import OrcEvo, mcsolve_oracle

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-8, 1e-6


# ------------------------------------------------------------------ data layer
@pytest.mark.parametrize("kind", ["csr", "dia", "dense"])
@pytest.mark.parametrize("fmt", [qb.FMT_AUTO, qb.FMT_CSR, qb.FMT_DIAM, qb.FMT_SELL, qb.FMT_RSELL])
def test_matmul_vs_reference(kind, fmt):
    g = load("matmul")
    if kind == "dense" and fmt != qb.FMT_AUTO:
        pytest.skip("dense has one layout")
    op = dev_op(*op_arrays(g, kind), fmt=fmt)
    x = qb.DeviceDense.from_numpy(g["x"])
    for si, sc in enumerate(g["scales"]):
        # in-place accumulate semantics: out pre-filled with ones (reference
        # tests/core/data/test_mathematics.py:864-911)
        out = qb.DeviceDense.from_numpy(np.ones(len(g["x"]), dtype=complex))
        E.matmul(op, x, sc, out)
        ref = 1.0 + g["%s_mul_s%d" % (kind, si)]
        np.testing.assert_allclose(out.to_numpy().ravel(), ref, rtol=1e-12, atol=1e-12)


def test_matmul_multicolumn_orders():
    g = load("matmul")
    rng = np.random.default_rng(3)
    n = len(g["x"])
    X = rng.random((n, 5)) + 1j * rng.random((n, 5))
    for kind in ("csr", "dia", "dense"):
        A = orc_op(g, kind).to_dense_array()
        op = dev_op(*op_arrays(g, kind))
        for order in ("F", "C"):
            x = qb.DeviceDense.from_numpy(np.array(X, order=order))
            out = E.matmul(op, x, 0.5j)
            np.testing.assert_allclose(out.to_numpy(), 0.5j * A @ X, rtol=1e-12, atol=1e-12)


def test_shape_errors():
    g = load("matmul")
    op = dev_op(*op_arrays(g, "csr"))
    x = qb.DeviceDense.from_numpy(np.ones(7, dtype=complex))
    with pytest.raises(qb.QbError, match="incompatible matrix shapes"):
        E.matmul(op, x)


def test_vector_ops():
    rng = np.random.default_rng(1)
    n = 100003
    a = rng.random(n) + 1j * rng.random(n)
    b = rng.random(n) + 1j * rng.random(n)
    da, db = qb.DeviceDense.from_numpy(a), qb.DeviceDense.from_numpy(b)
    assert abs(E.nrm2(da) - np.linalg.norm(a)) < 1e-9 * np.linalg.norm(a)
    np.testing.assert_allclose(E.inner(da, db), np.vdot(a, b), rtol=1e-12)
    E.axpy(da, 0.3 - 0.2j, db)
    np.testing.assert_allclose(db.to_numpy().ravel(), b + (0.3 - 0.2j) * a, rtol=1e-13)
    E.scal(da, 2j)
    np.testing.assert_allclose(da.to_numpy().ravel(), 2j * a, rtol=1e-13)
    diff, state = rng.random(n) * 1e-7 + 0j, a
    w = E.wrms_error(qb.DeviceDense.from_numpy(diff), qb.DeviceDense.from_numpy(state),
                     1e-8, 1e-6)
    ref = np.sqrt(np.mean((np.abs(diff) / (1e-8 + 1e-6 * np.abs(state))) ** 2))
    assert abs(w - ref) < 1e-10 * ref


def test_expect_ops():
    g = load("matmul")
    n = len(g["x"])
    rng = np.random.default_rng(2)
    for kind in ("csr", "dia", "dense"):
        A = orc_op(g, kind).to_dense_array()
        op = dev_op(*op_arrays(g, kind))
        x = g["x"]
        np.testing.assert_allclose(E.expect_ket(op, qb.DeviceDense.from_numpy(x)),
                                   np.vdot(x, A @ x), rtol=1e-12)
        rho = rng.random((n, n)) + 1j * rng.random((n, n))
        np.testing.assert_allclose(
            E.expect_dm(op, qb.DeviceDense.from_numpy(np.asfortranarray(rho))),
            np.trace(A @ rho), rtol=1e-12)
    m = 9
    rho = rng.random((m, m)) + 1j * rng.random((m, m))
    v = qb.DeviceDense.from_numpy(rho.ravel("F"))
    np.testing.assert_allclose(E.trace_oper_ket(v), np.trace(rho), rtol=1e-13)


# ------------------------------------------------------------------ mesolve
ME_CASES = [("c1_jc", "vern7"), ("c1_jc", "vern9"), ("c1_jc", "tsit5"), ("c2_tfim4", "vern7"),
            ("c2_tfim4", "vern9"), ("c4_driven", "vern7"), ("c5_kerr_0", "vern7")]


@pytest.mark.parametrize("name,method", ME_CASES)
@pytest.mark.parametrize("fmt", [qb.FMT_AUTO, qb.FMT_CSR, qb.FMT_DIAM, qb.FMT_SELL, qb.FMT_RSELL])
def test_mesolve_vs_reference(name, method, fmt):
    g = load(name)
    system = me_system_from_golden(g, fmt)
    eng = qb.Engine(system, method, nslots=1, store_states=1)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert r.status[0] == 1
    np.testing.assert_allclose(r.states[0], g["states_" + method], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r.expect[0], g["expect_" + method], rtol=RTOL, atol=ATOL)
    # much tighter in practice: same step sequence as the reference
    assert np.abs(r.states[0] - g["states_" + method]).max() < 1e-10


def test_mesolve_step_counts_match_oracle():
    from oracle.rk_oracle import mesolve_oracle
    g = load("c1_jc")
    rhs = orc_rhs(g)
    o = mesolve_oracle(rhs, g["y0"], g["tlist"], "vern7")
    eng = qb.Engine(me_system_from_golden(g), "vern7")
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert r.stats[0][0] == rhs.nevals
    assert r.stats[0][1] == o["rk"].n_accept and r.stats[0][2] == o["rk"].n_reject


@pytest.mark.parametrize("k", [1, 2, 3])
def test_mesolve_kerr_stability_limited(k):
    g = load("c5_kerr_%d" % k)
    eng = qb.Engine(me_system_from_golden(g), "vern7")
    r = eng.run_mesolve(g["y0"], g["tlist"])
    np.testing.assert_allclose(r.expect[0], g["expect_vern7"], rtol=RTOL, atol=ATOL)


def test_mesolve_batch_of_identical_systems():
    g = load("c2_tfim4")
    eng = qb.Engine(me_system_from_golden(g), "vern7", nslots=3)
    r = eng.run_mesolve(g["y0"], g["tlist"], ntraj=7)
    assert (r.status == 1).all()
    for j in range(7):
        np.testing.assert_allclose(r.expect[j], g["expect_vern7"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(r.expect[j], r.expect[0])      # deterministic


def test_integrator_protocol():
    """set_state / integrate / mcstep forward and backward (reference
    tests/solver/test_integrator.py:71-98 pattern)."""
    g = load("c1_jc")
    eng = qb.Engine(me_system_from_golden(g), "vern7")
    eng.set_state(0.0, g["y0"])
    ref = g["states_vern7"]
    for i in (10, 20, 30):
        t, st = eng.integrate(g["tlist"][i])
        assert t == g["tlist"][i] and st >= 0
        _, y = eng.get_state()
        np.testing.assert_allclose(y, ref[i], rtol=RTOL, atol=ATOL)
    t_front, st = eng.integrate(100.0, step=True)     # one step at most, may stop short
    assert g["tlist"][30] < t_front < 100.0
    t_front2, st = eng.integrate(100.0, step=True)
    assert t_front2 > t_front
    t_mid = 0.5 * (t_front + t_front2)
    t_back, st = eng.integrate(t_mid, step=True)      # backward: interpolation
    assert t_back == t_mid and st == 1
    t_out, st = eng.integrate(g["tlist"][30] - 5.0)
    assert st == -3                                     # OUTSIDE_RANGE


# ------------------------------------------------------------------ mcsolve
@pytest.mark.parametrize("name,method,nslots", [("c3_tfim6_mc", "vern7", 24),
                                               ("c3_tfim6_mc", "vern7", 5),
                                               ("c3_tfim4_mc_strong", "vern9", 7),
                                               ("c3_tfim4_mc_tsit5", "tsit5", 5)])
@pytest.mark.parametrize("fmt", [qb.FMT_AUTO, qb.FMT_CSR, qb.FMT_DIAM, qb.FMT_SELL, qb.FMT_RSELL])
def test_mcsolve_vs_reference(name, method, nslots, fmt):
    g = load(name)
    eng = qb.Engine(mc_system_from_golden(g, fmt), method, nslots=nslots)
    ntraj = int(g["ntraj"])
    r = eng.run_mcsolve(g["psi0"], g["tlist"], g["draws"], ntraj=ntraj,
                        final_states=nslots >= ntraj)
    assert (r.status == 1).all()
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    assert np.array_equal(r.ncol, g["col_count"])                # jump counts: bit-exact
    for j in range(ntraj):
        n = r.ncol[j]
        assert np.array_equal(r.col_which[j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r.col_t[j, :n], g["col_times"][cc[j]:cc[j + 1]],
                                   rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.transpose(r.expect, (1, 0, 2)), g["runs_expect"],
                               rtol=RTOL, atol=ATOL)
    if r.final_states is not None:
        np.testing.assert_allclose(r.final_states / np.linalg.norm(r.final_states, axis=1,
                                                                   keepdims=True),
                                   g["final_states"], rtol=RTOL, atol=ATOL)


def test_mcsolve_vs_oracle_random_system():
    """Seeded random system at a size the oracle finishes in seconds, unsorted CSR."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    n = 200
    H = sp.random(n, n, 0.03, random_state=7) + 1j * sp.random(n, n, 0.03, random_state=8)
    H = sp.csr_matrix(H + H.conj().T)
    cs = [sp.csr_matrix(sp.random(n, n, 0.01, random_state=20 + k) * 0.6) for k in range(3)]
    ns = [sp.csr_matrix(c.conj().T @ c) for c in cs]
    heff = sp.csr_matrix(-1j * H - 0.5 * sum(ns))
    e = sp.csr_matrix(sp.diags(rng.random(n)))
    psi0 = rng.random(n) + 1j * rng.random(n)
    psi0 /= np.linalg.norm(psi0)
    tlist = np.linspace(0, 1.0, 9)
    draws = rng.random((6, 64))
    s = qb.System(n)
    s.add_element(qb.DeviceOp.from_scipy(heff))
    for c, nn in zip(cs, ns):
        s.add_collapse(qb.DeviceOp.from_scipy(c), qb.DeviceOp.from_scipy(nn))
    s.add_eop(qb.DeviceOp.from_scipy(e))
    r = qb.Engine(s, "vern7", nslots=4).run_mcsolve(psi0, tlist, draws)
    assert (r.status == 1).all()
    from oracle.rk_oracle import OrcOp
    rhs = OrcEvo([(OrcOp.from_scipy(heff), 1.0)])
    ocs = [OrcEvo([(OrcOp.from_scipy(c), 1.0)]) for c in cs]
    ons = [OrcEvo([(OrcOp.from_scipy(nn), 1.0)]) for nn in ns]
    for j in range(6):
        o = mcsolve_oracle(rhs, ocs, ons, psi0, tlist, draws[j], [OrcOp.from_scipy(e)])
        assert r.ncol[j] == len(o["collapses"])
        assert list(r.col_which[j, :r.ncol[j]]) == [w for _, w in o["collapses"]]
        np.testing.assert_allclose(r.expect[j], o["expect"], rtol=RTOL, atol=ATOL)


def test_threshold_table_exhaustion_is_reported():
    g = load("c3_tfim4_mc_strong")
    eng = qb.Engine(mc_system_from_golden(g), "vern9", nslots=4)
    r = eng.run_mcsolve(g["psi0"], g["tlist"], g["draws"][:, :3], ntraj=int(g["ntraj"]))
    assert (r.status == -12).any()           # QB_ST_RNG_EXHAUSTED, never silent
    ok = r.status == 1
    assert np.array_equal(r.ncol[ok], g["col_count"][ok])


def test_mcsolve_super_operator_hamiltonian_vs_reference():
    """mcsolve with a super-operator H (solver/mcsolve.py:481-490) on the device: tr(rho) in
    place of the squared norm, tr(n_k rho) probabilities, renormalisation by the trace."""
    from _emul import FMT_CSR  # noqa: F401
    from _systems import functional_of, merged_constant_rhs
    g = load("c3_tfim3_mc_super")
    n = int(g["super_n"])
    s = qb.System(len(g["psi0"]))
    s.add_element(qb.DeviceOp.from_scipy(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(dev_op(*op_arrays(g, "cop%d" % i)), dev_op(*op_arrays(g, "nop%d" % i)))
    for i in range(int(g["n_eops"])):
        s.add_eop(qb.DeviceOp.from_scipy(functional_of(g["eop%d_full" % i])))
    s.set_functional(True)
    s.set_mc_trace(n)
    ntraj = int(g["ntraj"])
    for method, nslots in (("vern7", 5), ("vern7", 16)):
        eng = qb.Engine(s, method, nslots=nslots, store_states=1)
        r = eng.run_mcsolve(g["psi0"], g["tlist"], g["draws"], ntraj=ntraj)
        assert (r.status == 1).all()
        assert np.array_equal(r.ncol, g["col_count"])
        cc = np.concatenate([[0], np.cumsum(g["col_count"])])
        for j in range(ntraj):
            k = r.ncol[j]
            assert np.array_equal(r.col_which[j, :k], g["col_which"][cc[j]:cc[j + 1]])
            np.testing.assert_allclose(r.col_t[j, :k], g["col_times"][cc[j]:cc[j + 1]], rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.transpose(r.expect, (1, 0, 2)), g["runs_expect"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(r.states[:, -1, :], g["final_states"], rtol=1e-6, atol=1e-8)
