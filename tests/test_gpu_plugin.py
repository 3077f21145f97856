"""GPU tests of the QuTiP plug-in: the registered integrators, data-layer types and the
'b200' map against the reference's own solvers run side by side (unmodified qutip
5.4.0.dev from oracle/_ref) with identical options."""
import pickle
import sys
import warnings

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

_ref = oracle.ref_path()
if _ref is None:
    pytest.skip("reference build (oracle/_ref) not present", allow_module_level=True)
sys.path.insert(0, _ref)
warnings.filterwarnings("ignore")
import qutip  # noqa: E402
from qutip import (QobjEvo, basis, destroy, mcsolve, mesolve, qeye, sigmam, sigmax,  # noqa
                   sigmaz, tensor)
import qutip_b200.plugin as plugin  # noqa: E402

ATOL, RTOL = 1e-8, 1e-6
OPT = {"progress_bar": False}


def tfim(n, gamma=0.1):
    sx, sz, sm = [], [], []
    for i in range(n):
        ops = [qeye(2)] * n
        ops[i] = sigmax(); sx.append(tensor(ops))
        ops[i] = sigmaz(); sz.append(tensor(ops))
        ops[i] = sigmam(); sm.append(tensor(ops))
    H = 0
    for i in range(n - 1):
        H = H - sz[i] * sz[i + 1]
    for i in range(n):
        H = H - sx[i]
    return H, [np.sqrt(gamma) * s for s in sm], sz


def jc():
    N = 10
    a = tensor(destroy(N), qeye(2)); sm = tensor(qeye(N), destroy(2))
    H = 2 * np.pi * a.dag() * a + 2 * np.pi * sm.dag() * sm \
        + 2 * np.pi * 0.05 * (a.dag() * sm + a * sm.dag())
    c_ops = [np.sqrt(0.1) * a, np.sqrt(0.05) * sm]
    psi0 = tensor(basis(N, 3), basis(2, 0))
    return H, c_ops, psi0, [a.dag() * a, tensor(qeye(N), sigmaz())]


def test_registered():
    assert "b200_vern7" in qutip.MESolver.avail_integrators()
    assert "b200_vern9" in qutip.MCSolver.avail_integrators()
    assert plugin.B200Dense in qutip.core.data.to.dtypes
    assert qutip.solver.parallel._maps["b200"] is plugin.b200_map


@pytest.mark.parametrize("method", ["vern7", "vern9", "tsit5"])
def test_mesolve_c1_jc(method):
    H, c_ops, psi0, e_ops = jc()
    tl = np.linspace(0, 10, 101)
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method=method, store_states=True))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops,
                  options=dict(OPT, method="b200_" + method, store_states=True))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)
    for a, b in zip(out.states, ref.states):
        np.testing.assert_allclose(a.full(), b.full(), rtol=RTOL, atol=ATOL)


def test_mesolve_c4_time_dependent_string_and_array():
    Nc = 8
    a = tensor(destroy(Nc), qeye(3)); b = tensor(qeye(Nc), destroy(3))
    H0 = 5 * a.dag() * a + 4.5 * b.dag() * b - 0.15 * b.dag() * b.dag() * b * b \
        + 0.1 * (a.dag() * b + a * b.dag())
    tl = np.linspace(0, 5, 51)
    env = np.exp(-((tl - 2.5) / 1.0) ** 2)
    H = QobjEvo([H0, [a + a.dag(), "A*cos(w*t)"], [b + b.dag(), env]],
                args={"A": 0.2, "w": 5.0}, tlist=tl)
    c_ops = [np.sqrt(0.01) * a, np.sqrt(0.02) * b, np.sqrt(0.03) * b.dag() * b]
    psi0 = tensor(basis(Nc, 0), basis(3, 0))
    e_ops = [a.dag() * a, b.dag() * b]
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="vern7"))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)


def test_sesolve_and_options():
    H, _, psi0, e_ops = jc()
    tl = np.linspace(0, 3, 31)
    o = dict(OPT, atol=1e-10, rtol=1e-8, first_step=1e-3, max_step=0.05)
    ref = qutip.sesolve(H, psi0, tl, e_ops=e_ops, options=dict(o, method="vern7"))
    out = qutip.sesolve(H, psi0, tl, e_ops=e_ops, options=dict(o, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)
    with pytest.raises(KeyError):
        qutip.sesolve(H, psi0, tl, options=dict(OPT, method="b200_vern7", bogus=1))


def test_nsteps_failure_raises_reference_exception():
    H, c_ops, psi0, e_ops = jc()
    with pytest.raises(qutip.solver.integrator.IntegratorException, match="Too much work"):
        mesolve(H, psi0, [0, 50], c_ops, options=dict(OPT, method="b200_vern7", nsteps=3))


def test_refuses_host_only_forms():
    a = destroy(5)
    # operator-valued python functions cannot be split into (operator, scalar coefficient)
    H = QobjEvo(lambda t: a.dag() * a * np.cos(t))
    with pytest.raises(TypeError, match="device"):
        mesolve(H, basis(5, 0), [0, 1], [a], options=dict(OPT, method="b200_vern7"))
    with pytest.raises(TypeError, match="QobjEvo.matmul_data"):
        plugin.B200Vern7(lambda t, y: y, {})
    # python-callable scalar coefficients cannot be used by the whole-batch map
    Hc = QobjEvo([a.dag() * a, [a + a.dag(), lambda t: np.cos(t)]])
    with pytest.raises(TypeError, match="b200"):
        mcsolve(Hc, basis(5, 1), [0, 1], [a], ntraj=2, options=dict(OPT, map="b200"))


@pytest.mark.parametrize("method,dev,rtol,atol", [("vern7", "b200_vern7", RTOL, ATOL),
                                                  ("adams", "b200_zvode", RTOL, ATOL),
                                                  ("adams", "b200_adams", 1e-4, 5e-5)])
def test_python_function_coefficients_evaluated_by_host(method, dev, rtol, atol):
    """FunctionCoefficient: the scalar is evaluated on the host for each stage time while the
    matvecs, stage combinations and step control stay on the device."""
    a = destroy(8)
    H = QobjEvo([a.dag() * a, [a + a.dag(), lambda t, A: A * np.cos(2 * t)]], args={"A": 0.4})
    c_ops = [QobjEvo([a, lambda t, k: np.sqrt(k * np.exp(-t))], args={"k": 0.5})]
    tl = np.linspace(0, 3, 16)
    ref = mesolve(H, basis(8, 3), tl, c_ops, e_ops=[a.dag() * a], options=dict(OPT, method=method))
    out = mesolve(H, basis(8, 3), tl, c_ops, e_ops=[a.dag() * a], options=dict(OPT, method=dev))
    np.testing.assert_allclose(out.expect[0], ref.expect[0], rtol=rtol, atol=atol)
    # args updated between steps (Solver.step(t, args=...), reference test_mesolver_stepping)
    s_ref = qutip.MESolver(H, c_ops, options=dict(OPT, method=method))
    s_out = qutip.MESolver(H, c_ops, options=dict(OPT, method=dev))
    for s in (s_ref, s_out):
        s.start(basis(8, 3), 0)
    for t, args in ((1.0, None), (2.0, {"k": 0.0, "A": 0.1})):
        x, y = s_ref.step(t, args=args), s_out.step(t, args=args)
        np.testing.assert_allclose(y.full(), x.full(), rtol=max(rtol, 1e-5), atol=max(atol, 1e-7))


def test_integrator_pickle_roundtrip():
    H, c_ops, psi0, e_ops = jc()
    s = qutip.MESolver(H, c_ops, options=dict(OPT, method="b200_vern7"))
    s.start(psi0, 0.0)
    s.step(0.5)
    integ = pickle.loads(pickle.dumps(s._integrator))
    t1, y1 = s._integrator.integrate(1.0)
    t2, y2 = integ.integrate(1.0)
    assert t1 == t2 == 1.0
    np.testing.assert_allclose(y1.to_array(), y2.to_array(), rtol=RTOL, atol=ATOL)


def test_mcsolve_with_registered_integrator_python_driver():
    """Stock MCIntegrator (python jump logic) on top of the device integrator."""
    H, c_ops, sz = tfim(5)
    psi0 = basis([2] * 5, [0] * 5)
    tl = np.linspace(0, 2, 11)
    kw = dict(e_ops=[sz[0]], ntraj=6, seeds=np.random.SeedSequence(3))
    ref = mcsolve(H, psi0, tl, c_ops, options=dict(OPT, method="vern7", keep_runs_results=True), **kw)
    kw["seeds"] = np.random.SeedSequence(3)
    out = mcsolve(H, psi0, tl, c_ops, options=dict(OPT, method="b200_vern7", keep_runs_results=True), **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("method", ["vern7", "vern9", "tsit5"])
def test_mcsolve_b200_map_matches_reference(method):
    """Whole batch on the device through options['map']='b200': identical collapse records
    (TestSeeds-style determinism, reference tests/solver/test_mcsolve.py:326-404)."""
    H, c_ops, sz = tfim(6)
    psi0 = basis([2] * 6, [0] * 6)
    tl = np.linspace(0, 2, 21)
    e_ops = [sz[0], sz[2]]
    ref = mcsolve(H, psi0, tl, c_ops, e_ops=e_ops, ntraj=20, seeds=np.random.SeedSequence(7),
                  options=dict(OPT, method=method, keep_runs_results=True))
    out = mcsolve(H, psi0, tl, c_ops, e_ops=e_ops, ntraj=20, seeds=np.random.SeedSequence(7),
                  options=dict(OPT, method=method, map="b200", keep_runs_results=True))
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.array(out.std_expect), np.array(ref.std_expect), rtol=1e-5, atol=1e-7)
    assert out.num_trajectories == 20


def test_mcsolve_b200_map_states_and_time_dependent_cops():
    a = destroy(6)
    H = a.dag() * a
    c_ops = [QobjEvo([a, "sqrt(g)*exp(-k*t)"], args={"g": 0.8, "k": 0.3}), 0.2 * a.dag() * a]
    psi0 = basis(6, 4)
    tl = np.linspace(0, 3, 13)
    o = dict(OPT, method="vern7", keep_runs_results=True, store_final_state=True)
    ref = mcsolve(H, psi0, tl, c_ops, e_ops=[a.dag() * a], ntraj=12, seeds=5, options=o)
    out = mcsolve(H, psi0, tl, c_ops, e_ops=[a.dag() * a], ntraj=12, seeds=5,
                  options=dict(o, map="b200"))
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)
    for x, y in zip(out.runs_final_states, ref.runs_final_states):
        np.testing.assert_allclose(x.full(), y.full(), rtol=RTOL, atol=ATOL)


def test_data_layer_types_in_stock_rk_loop():
    """Stock vern7 running on the registered device data types through the dispatchers
    (QobjEvo.matmul_data -> matmul[B200Operator, B200Dense], add, mul, norms ...)."""
    H, c_ops, psi0, e_ops = jc()
    tl = np.linspace(0, 2, 11)
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="vern7"))
    L = qutip.liouvillian(H, c_ops).to("B200Operator")
    rho0 = qutip.operator_to_vector(qutip.ket2dm(psi0)).to("b200")
    solver = qutip.MESolver(L, options=dict(OPT, method="vern7"))
    integ = solver._integrator
    integ.set_state(0.0, rho0.data)
    assert isinstance(integ.get_state()[1], plugin.B200Dense)
    for i, t in enumerate(tl[1:], 1):
        _, y = integ.integrate(t)
        assert isinstance(y, plugin.B200Dense)
        rho = qutip.vector_to_operator(qutip.Qobj(y.to_array(), dims=rho0.dims))
        for k, e in enumerate(e_ops):
            assert abs(qutip.expect(e, rho) - ref.expect[k][i]) < 1e-7


def test_data_layer_specialisations_direct():
    from qutip.core import data as _data
    rng = np.random.default_rng(0)
    n = 40
    A = qutip.rand_herm(n, density=0.2, seed=1).data
    x = rng.random((n, 1)) + 1j * rng.random((n, 1))
    dA = _data.to(plugin.B200Operator, A)
    dx = _data.to(plugin.B200Dense, _data.Dense(x))
    y = _data.matmul(dA, dx, 0.5j)
    assert isinstance(y, plugin.B200Dense)
    np.testing.assert_allclose(y.to_array(), 0.5j * A.to_array() @ x, atol=1e-13)
    np.testing.assert_allclose(_data.add(dx, y, 2.0).to_array(), x + 2 * y.to_array(), atol=1e-13)
    np.testing.assert_allclose(_data.mul(dx, 3j).to_array(), 3j * x, atol=1e-13)
    assert abs(_data.norm.l2(dx) - np.linalg.norm(x)) < 1e-12
    assert abs(_data.expect(dA, dx) - np.vdot(x, A.to_array() @ x)) < 1e-12
    assert abs(_data.inner(dx, y) - np.vdot(x, y.to_array())) < 1e-12
    z = _data.zeros[plugin.B200Dense](3, 1)
    assert np.all(z.to_array() == 0)
    back = _data.to(_data.Dense, y)
    np.testing.assert_allclose(back.to_array(), y.to_array())
    with pytest.raises(ValueError, match="incompatible matrix shapes"):
        _data.matmul(dA, _data.to(plugin.B200Dense, _data.Dense(np.ones((3, 1), dtype=complex))))


def test_c4_full_size_time_dependent():
    """C4 at full size (cavity 40 x transmon 3, Liouvillian 14400^2, 4 elements with the
    string coefficient and its conjugate evaluated on the device)."""
    Nc = 40
    a = tensor(destroy(Nc), qeye(3)); b = tensor(qeye(Nc), destroy(3))
    H0 = 5 * a.dag() * a + 4.5 * b.dag() * b - 0.15 * b.dag() * b.dag() * b * b \
        + 0.1 * (a.dag() * b + a * b.dag())
    H = QobjEvo([H0, [a + a.dag(), "A*cos(w*t)"]], args={"A": 0.2, "w": 5.0})
    c_ops = [np.sqrt(0.01) * a, np.sqrt(0.02) * b, np.sqrt(0.03) * b.dag() * b]
    psi0 = tensor(basis(Nc, 0), basis(3, 0))
    tl = np.linspace(0, 5, 51)
    e_ops = [a.dag() * a, b.dag() * b]
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="vern7"))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)


def test_c5_sweep_members_match_reference():
    """Batched parameter sweep: per-member constant coefficients (U, Delta, F) read from the
    per-trajectory argument table; a few members checked against the reference mesolve."""
    from qutip_b200 import coeffs, models, solve
    N = 30
    Ls, args, a_sp = models.kerr_sweep(N, grid=16)
    idx = {"U": 0, "D": 1, "F": 2}
    elements = [(Ls[0], coeffs.compile_expr("U", arg_index=idx)),
                (Ls[1], coeffs.compile_expr("D", arg_index=idx)),
                (Ls[2], coeffs.compile_expr("F", arg_index=idx)),
                (Ls[3], None)]
    n_op = (a_sp.conj().T @ a_sp).toarray()
    rho0 = np.zeros(N * N, dtype=complex); rho0[0] = 1.0
    tl = np.linspace(0, 10, 21)
    pick = [0, 17, 255, 1000, 2222, 4095]
    r = solve.mesolve(elements, rho0, tl, e_ops=[n_op], args=args[pick], nargs=3,
                      store_states=False)
    a = destroy(N)
    for k, j in enumerate(pick):
        U, D, F = args[j].real
        Hq = 0.5 * U * a.dag() * a.dag() * a * a - D * a.dag() * a + F * (a + a.dag())
        ref = mesolve(Hq, basis(N, 0), tl, [a], e_ops=[a.dag() * a], options=dict(OPT, method="vern7"))
        np.testing.assert_allclose(r.expect[k, 0].real, ref.expect[0], rtol=RTOL, atol=ATOL)


def test_zvode_with_device_rhs():
    """method='b200_zvode': SciPy zvode (the reference's Adams) with the RHS on the device."""
    H, c_ops, psi0, e_ops = jc()
    tl = np.linspace(0, 5, 26)
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="adams"))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="b200_zvode"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)
    # time-dependent + mcsolve driver on top of it
    a = destroy(6)
    Ht = QobjEvo([a.dag() * a, [a + a.dag(), "0.3*sin(2*t)"]])
    ref = mcsolve(Ht, basis(6, 2), np.linspace(0, 2, 9), [0.5 * a], e_ops=[a.dag() * a], ntraj=5,
                  seeds=3, options=dict(OPT, method="adams", keep_runs_results=True))
    out = mcsolve(Ht, basis(6, 2), np.linspace(0, 2, 9), [0.5 * a], e_ops=[a.dag() * a], ntraj=5,
                  seeds=3, options=dict(OPT, method="b200_zvode", keep_runs_results=True))
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=1e-5, atol=1e-7)


def test_device_resident_adams():
    """method='b200_adams': the Nordsieck Adams-Moulton method on the device against the
    reference's zvode-based 'adams' at the accuracy the reference asks of it
    (tests/solver/test_integrator.py:71-98, test_mesolve.py:110-123), same option keys."""
    stock = qutip.solver.integrator.scipy_integrator.IntegratorScipyAdams.integrator_options
    assert plugin.B200Adams.integrator_options == stock
    H, c_ops, psi0, e_ops = jc()
    tl = np.linspace(0, 5, 26)
    exact = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="vern9", atol=1e-12, rtol=1e-10))
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="adams"))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="b200_adams", store_states=True))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=1e-4, atol=5e-5)
    err_out = np.abs(np.array(out.expect) - np.array(exact.expect)).max()
    err_ref = np.abs(np.array(ref.expect) - np.array(exact.expect)).max()
    assert err_out < 5e-5 and err_out < 50 * max(err_ref, 1e-7)
    assert abs(out.states[-1].tr() - 1) < 1e-6
    # tighter tolerances and a capped order are honoured
    tight = mesolve(H, psi0, tl, c_ops, e_ops=e_ops,
                    options=dict(OPT, method="b200_adams", atol=1e-11, rtol=1e-9, order=5, nsteps=20000))
    assert np.abs(np.array(tight.expect) - np.array(exact.expect)).max() < 2e-7
    with pytest.raises(Exception, match="Too much work"):
        mesolve(H, psi0, tl, c_ops, options=dict(OPT, method="b200_adams", nsteps=3))
    # time-dependent Hamiltonian, mcsolve driver on top of the integrator (mcstep protocol)
    a = destroy(6)
    Ht = QobjEvo([a.dag() * a, [a + a.dag(), "0.3*sin(2*t)"]])
    kw = dict(e_ops=[a.dag() * a], ntraj=5, seeds=3)
    ref = mcsolve(Ht, basis(6, 2), np.linspace(0, 2, 9), [0.5 * a],
                  options=dict(OPT, method="adams", keep_runs_results=True), **kw)
    out = mcsolve(Ht, basis(6, 2), np.linspace(0, 2, 9), [0.5 * a],
                  options=dict(OPT, method="b200_adams", keep_runs_results=True), **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.concatenate(out.col_times), np.concatenate(ref.col_times), atol=1e-3)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=1e-3, atol=1e-3)
    # whole-batch mcsolve on the device with the Adams method
    out = mcsolve(Ht, basis(6, 2), np.linspace(0, 2, 9), [0.5 * a],
                  options=dict(OPT, method="adams", map="b200", keep_runs_results=True), **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=1e-3, atol=1e-3)


def test_matrix_valued_state_unitary_evolution():
    """sesolve of an operator (propagator-style): N x N state through the device integrator."""
    H = qutip.rand_herm(6, seed=2)
    U0 = qutip.qeye(6)
    tl = np.linspace(0, 2, 9)
    ref = qutip.sesolve(H, U0, tl, options=dict(OPT, method="vern7"))
    out = qutip.sesolve(H, U0, tl, options=dict(OPT, method="b200_vern7"))
    for a, b in zip(out.states, ref.states):
        np.testing.assert_allclose(a.full(), b.full(), rtol=RTOL, atol=ATOL)


def test_b200_map_improved_sampling():
    """options['improved_sampling']: thresholds floored at the no-jump probability
    (mcsolve.py:276-279) and trajectories weighted by 1 - p_nojump."""
    H, c_ops, sz = tfim(4, gamma=0.05)
    psi0 = basis([2] * 4, [0] * 4)
    tl = np.linspace(0, 1.5, 7)
    o = dict(OPT, method="vern7", improved_sampling=True, keep_runs_results=True)
    ref = mcsolve(H, psi0, tl, c_ops, e_ops=[sz[0]], ntraj=10, seeds=9, options=o)
    out = mcsolve(H, psi0, tl, c_ops, e_ops=[sz[0]], ntraj=10, seeds=9, options=dict(o, map="b200"))
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect),
                               rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("method", ["b200_vern7", "b200_adams", "b200_zvode"])
def test_matrix_form_option(method):
    """options['matrix_form']=True (LindbladMatrixForm RHS on the un-vectorised rho) through
    the device integrators, against the reference's matrix-form and superoperator results."""
    H, c_ops, psi0, e_ops = jc()
    tl = np.linspace(0, 3, 13)
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method="vern7", matrix_form=True,
                                                               store_states=True))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(OPT, method=method, matrix_form=True,
                                                               store_states=True))
    tol = {"b200_vern7": dict(rtol=RTOL, atol=ATOL), "b200_zvode": dict(rtol=1e-4, atol=1e-6),
           "b200_adams": dict(rtol=1e-4, atol=5e-5)}[method]
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), **tol)
    np.testing.assert_allclose(out.states[-1].full(), ref.states[-1].full(), **tol)


def test_matrix_form_binds_matrix_free(monkeypatch):
    """options['matrix_form']=True with the matrix-free binding forced (Kronecker operators
    I (x) H_nh, conj(H_nh) (x) I and the sparse jump part instead of a fused superoperator),
    against the reference's own matrix-form run; time-dependent H term and collapse rate."""
    from qutip_b200 import engine as E
    monkeypatch.setenv("QUTIP_B200_MATRIX_FREE", "1")
    n = 4
    H, c_ops, sz = tfim(n, 0.09)
    psi0 = basis([2] * n, [0] * n)
    ops = [qeye(2)] * n
    ops[0] = sigmax()
    Ht = [H, [tensor(ops), "0.4 * sin(2 * t)"]]
    c_all = list(c_ops) + [[sz[1], "0.2 * exp(-t)"]]
    e_ops = [sz[0], sz[n - 1]]
    tl = np.linspace(0, 2, 9)
    opt = dict(OPT, atol=1e-8, rtol=1e-6, store_states=True, matrix_form=True)
    ref = mesolve(Ht, psi0, tl, c_all, e_ops=e_ops, options=dict(opt, method="vern7"))
    out = mesolve(Ht, psi0, tl, c_all, e_ops=e_ops, options=dict(opt, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out.states[-1].full(), ref.states[-1].full(), rtol=RTOL, atol=ATOL)
    solver = qutip.MESolver(QobjEvo(Ht), [QobjEvo(c) if isinstance(c, list) else c for c in c_all],
                            options=dict(opt, method="b200_vern7"))
    solver.start(qutip.ket2dm(psi0), 0.0)
    fmts = [el.info()["fmt"] for el in solver._integrator._system.element_ops]
    # (left, right) x (drive term, |g|^2 c^dag c term of the time-dependent collapse, constant H_nh)
    assert fmts.count(E.FMT_KRON) == 6, fmts
    # python-callable coefficients fall back to the fused superoperator binding
    Hf = [H, [tensor(ops), lambda t: 0.4 * np.sin(2 * t)]]
    ref = mesolve(Hf, psi0, tl, c_ops, e_ops=e_ops, options=dict(opt, method="vern7"))
    out = mesolve(Hf, psi0, tl, c_ops, e_ops=e_ops, options=dict(opt, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("improved", [False, True])
def test_b200_map_mixed_initial_states(improved):
    """mcsolve from a statistical mixture (density matrix or [(ket, weight)] list,
    multitraj.py:285-352, mcsolve.py:752-792): one device batch per initial state, weights
    and the per-state no-jump floors of improved sampling as in the reference."""
    a = destroy(6)
    H = a.dag() * a + 0.3 * (a + a.dag())
    c_ops = [0.5 * a, 0.2 * a.dag() * a]
    ics = [(basis(6, 3), 0.5), (basis(6, 1), 0.3), ((basis(6, 0) + basis(6, 2)).unit(), 0.2)]
    tl = np.linspace(0, 2, 9)
    kw = dict(e_ops=[a.dag() * a, a + a.dag()], ntraj=[6, 4, 3], seeds=11)
    o = dict(OPT, keep_runs_results=True, improved_sampling=improved)
    ref = qutip.MCSolver(H, c_ops, options=dict(o, method="vern7")).run(ics, tl, **kw)
    out = qutip.MCSolver(H, c_ops, options=dict(o, method="vern7", map="b200")).run(ics, tl, **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect), rtol=RTOL, atol=ATOL)
    assert list(out.ntraj_per_initial_state) == list(ref.ntraj_per_initial_state)
    # a density matrix as initial state is decomposed by the reference into its eigenstates
    rho0 = 0.6 * qutip.ket2dm(basis(6, 2)) + 0.4 * qutip.ket2dm(basis(6, 4))
    kw2 = dict(e_ops=[a.dag() * a], ntraj=8, seeds=5)
    ref = mcsolve(H, rho0, tl, c_ops, options=dict(o, method="vern7"), **kw2)
    out = mcsolve(H, rho0, tl, c_ops, options=dict(o, method="vern7", map="b200"), **kw2)
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect), rtol=RTOL, atol=ATOL)


def test_propagator_and_nm_mcsolve_reuse_the_device_integrator():
    """Callers that reuse the same integrators: propagator() drives
    MESolver with a matrix-valued state; NonMarkovianMCSolver subclasses MCIntegrator."""
    a = destroy(5)
    H = a.dag() * a + 0.3 * (a + a.dag())
    U_ref = qutip.propagator(H, 1.5, c_ops=[0.2 * a], options=dict(OPT, method="vern7"))
    U_out = qutip.propagator(H, 1.5, c_ops=[0.2 * a], options=dict(OPT, method="b200_vern7"))
    np.testing.assert_allclose(U_out.full(), U_ref.full(), rtol=1e-5, atol=1e-7)
    ops_and_rates = [(a, "0.3 + 0.2 * sin(t)")]
    kw = dict(e_ops=[a.dag() * a], ntraj=6, seeds=4)
    ref = qutip.nm_mcsolve(H, basis(5, 3), np.linspace(0, 2, 9), ops_and_rates,
                           options=dict(OPT, method="vern7", keep_runs_results=True), **kw)
    out = qutip.nm_mcsolve(H, basis(5, 3), np.linspace(0, 2, 9), ops_and_rates,
                           options=dict(OPT, method="b200_vern7", keep_runs_results=True), **kw)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=1e-5, atol=1e-7)
    # the 'b200' map does not take over subclasses of MCSolver: nm_mcsolve keeps its own
    # trajectory function (martingale weights), run in-process with the device integrator
    out = qutip.nm_mcsolve(H, basis(5, 3), np.linspace(0, 2, 9), ops_and_rates,
                           options=dict(OPT, method="b200_vern7", map="b200", keep_runs_results=True), **kw)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=1e-5, atol=1e-7)


def test_b200_map_bulk_feed_matches_per_trajectory_add():
    """keep_runs_results=False takes the vectorised feed (running sums of McResult updated in
    bulk): averages, std, collapse records and counters must equal both the reference's own
    run and the per-trajectory feed (multitrajresult.py:402-434,1116-1124)."""
    H, c_ops, sz = tfim(6)
    psi0 = basis([2] * 6, [0] * 6)
    tl = np.linspace(0, 2, 21)
    e_ops = [sz[0], sz[1] * sz[2], sigmam_like(6)]          # two Hermitian, one non-Hermitian e_op
    kw = dict(e_ops=e_ops, ntraj=40)
    ref = mcsolve(H, psi0, tl, c_ops, seeds=np.random.SeedSequence(11),
                  options=dict(OPT, method="vern7"), **kw)
    out = mcsolve(H, psi0, tl, c_ops, seeds=np.random.SeedSequence(11),
                  options=dict(OPT, method="vern7", map="b200"), **kw)
    assert out.num_trajectories == ref.num_trajectories == 40
    assert len(out.seeds) == 40 and out.stats["end_condition"] == ref.stats["end_condition"]
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    for a, b in zip(out.average_expect, ref.average_expect):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)
    for a, b in zip(out.std_expect, ref.std_expect):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)
    assert out.runs_weights == ref.runs_weights


def sigmam_like(n):
    ops = [qeye(2)] * n
    ops[0] = sigmam()
    return tensor(ops)


def test_b200_map_more_collapses_than_initial_record(monkeypatch):
    """A strongly damped oscillator jumps more often than the initial capacity of the device
    collapse record: those trajectories are re-run with a larger record instead of failing
    (the reference appends to a python list, mcsolve.py:371-406)."""
    monkeypatch.setattr(plugin, "_MAX_COLLAPSES", 4)
    a = destroy(12)
    H = a.dag() * a
    c_ops = [np.sqrt(2.0) * a, np.sqrt(1.5) * a.dag()]
    psi0 = basis(12, 6)
    tl = np.linspace(0, 3, 7)
    o = dict(OPT, method="vern7", keep_runs_results=True)
    ref = mcsolve(H, psi0, tl, c_ops, e_ops=[a.dag() * a], ntraj=8, seeds=2, options=o)
    out = mcsolve(H, psi0, tl, c_ops, e_ops=[a.dag() * a], ntraj=8, seeds=2, options=dict(o, map="b200"))
    assert max(len(w) for w in ref.col_which) > 4
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)


def test_b200_map_sharded_over_devices():
    """plugin.configure(devices=[0, 1]): contiguous blocks of the seed list per device, one
    ncclAllReduce of the expectation sums -- identical records to the single-device run."""
    import qutip_b200 as qb
    if qb.device_count() < 2:
        pytest.skip("needs two GPUs")
    H, c_ops, sz = tfim(6)
    psi0 = basis([2] * 6, [0] * 6)
    tl = np.linspace(0, 2, 21)
    kw = dict(e_ops=[sz[0], sz[3]], ntraj=31)
    for keep in (False, True):
        o = dict(OPT, method="vern7", map="b200", keep_runs_results=keep)
        plugin.configure(None)
        one = mcsolve(H, psi0, tl, c_ops, seeds=np.random.SeedSequence(5), options=o, **kw)
        plugin.configure([0, 1])
        try:
            two = mcsolve(H, psi0, tl, c_ops, seeds=np.random.SeedSequence(5), options=o, **kw)
        finally:
            plugin.configure(None)
        assert [list(w) for w in two.col_which] == [list(w) for w in one.col_which]
        assert [list(t) for t in two.col_times] == [list(t) for t in one.col_times]
        for a, b in zip(two.average_expect, one.average_expect):
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13)
        for a, b in zip(two.std_expect, one.std_expect):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-7)


def test_c2_full_size_mesolve_vs_reference():
    """BASELINE config 2 at full size (TFIM 10 spins, rho-vector 2^20) through qutip.mesolve:
    the reference's own vern7 (matrix_form, so that the 24.6 M-nnz Liouvillian need not be
    built on the host) against method='b200_vern7' on t in [0, 0.1]."""
    H, c_ops, sz = tfim(10)
    psi0 = basis([2] * 10, [0] * 10)
    tl = np.linspace(0, 0.1, 3)
    e_ops = [sz[0], sz[4] * sz[5]]
    o = dict(OPT, matrix_form=True, store_states=False, store_final_state=True)
    ref = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(o, method="vern7"))
    out = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=dict(o, method="b200_vern7"))
    np.testing.assert_allclose(np.array(out.expect), np.array(ref.expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out.final_state.full(), ref.final_state.full(), rtol=RTOL, atol=ATOL)


_C3_REF_SCRIPT = r"""
import os, sys, warnings
import numpy as np
sys.path.insert(0, sys.argv[1])
warnings.filterwarnings("ignore")
from qutip import basis, mcsolve, qeye, sigmam, sigmax, sigmaz, tensor
n, ntraj = 14, int(sys.argv[3])
sx, sz, sm = [], [], []
for i in range(n):
    ops = [qeye(2)] * n
    ops[i] = sigmax(); sx.append(tensor(ops))
    ops[i] = sigmaz(); sz.append(tensor(ops))
    ops[i] = sigmam(); sm.append(tensor(ops))
H = 0
for i in range(n - 1):
    H = H - sz[i] * sz[i + 1]
for i in range(n):
    H = H - sx[i]
c_ops = [np.sqrt(0.1) * s for s in sm]
cores = len(os.sched_getaffinity(0))
ref = mcsolve(H, basis([2] * n, [0] * n), np.linspace(0, 2, 21), c_ops, e_ops=[sz[0]], ntraj=ntraj,
              seeds=np.random.SeedSequence(7),
              options={"progress_bar": False, "method": "vern7", "keep_runs_results": True,
                       "map": "parallel" if cores > 1 else "serial", "num_cpus": cores})
order = np.argsort([s.spawn_key[-1] for s in ref.seeds])
np.savez(sys.argv[2], runs_expect=np.array(ref.runs_expect)[:, order],
         ncol=np.array([len(ref.col_which[i]) for i in order]),
         col_which=np.concatenate([np.asarray(ref.col_which[i], dtype=int) for i in order] + [np.zeros(0, dtype=int)]),
         col_times=np.concatenate([np.asarray(ref.col_times[i], dtype=float) for i in order] + [np.zeros(0)]))
"""


def test_c3_full_size_256_trajectories_vs_reference(tmp_path):
    """BASELINE config 3 at full size (TFIM 14 spins, dim 16384): 256 trajectories of the
    reference's own mcsolve (all host cores; run in a fresh interpreter because its process
    pool forks) against the b200 map -- jump counts and collapse indices bit-exact, times to
    1e-9, expectation values within 1e-8 / 1e-6."""
    import os
    import subprocess
    ntraj = 256
    out_file = str(tmp_path / "c3_ref.npz")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    subprocess.run([sys.executable, "-c", _C3_REF_SCRIPT, _ref, out_file, str(ntraj)], check=True,
                   timeout=500, env=env)
    ref = np.load(out_file)
    H, c_ops, sz = tfim(14)
    psi0 = basis([2] * 14, [0] * 14)
    tl = np.linspace(0, 2, 21)
    out = mcsolve(H, psi0, tl, c_ops, e_ops=[sz[0]], ntraj=ntraj, seeds=np.random.SeedSequence(7),
                  options=dict(OPT, method="vern7", keep_runs_results=True, map="b200"))
    oo = np.argsort([s.spawn_key[-1] for s in out.seeds])
    assert len(oo) == ntraj
    cc = np.concatenate([[0], np.cumsum(ref["ncol"])])
    oe_ = np.array(out.runs_expect)
    for i, j in enumerate(oo):
        assert list(out.col_which[j]) == list(ref["col_which"][cc[i]:cc[i + 1]])
        np.testing.assert_allclose(out.col_times[j], ref["col_times"][cc[i]:cc[i + 1]], rtol=0, atol=1e-9)
        np.testing.assert_allclose(oe_[:, j], ref["runs_expect"][:, i], rtol=RTOL, atol=ATOL)


def test_b200_map_super_operator_hamiltonian():
    """mcsolve(liouvillian(H), ...) through the b200 map (solver/mcsolve.py:481-490,
    311-319): same collapse records and expectation values as the reference's own run."""
    H, c_ops, sz = tfim(3, gamma=0.8)
    psi0 = basis([2] * 3, [0] * 3)
    tl = np.linspace(0, 3, 13)
    L = qutip.liouvillian(H)
    o = dict(OPT, method="vern7", keep_runs_results=True, store_final_state=True)
    kw = dict(e_ops=[sz[0], sz[1]], ntraj=16, seeds=np.random.SeedSequence(21))
    ref = mcsolve(L, psi0, tl, c_ops, options=o, **kw)
    kw["seeds"] = np.random.SeedSequence(21)
    out = mcsolve(L, psi0, tl, c_ops, options=dict(o, map="b200"), **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    assert sum(len(w) for w in ref.col_which) > 10
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=RTOL, atol=ATOL)
    for x, y in zip(out.runs_final_states, ref.runs_final_states):
        np.testing.assert_allclose(x.full(), y.full(), rtol=RTOL, atol=ATOL)


def test_matmul_dag_specialisations():
    """matmul_dag (core/data/matmul.pyx:1084-1116) registered for the device types: the
    square rho @ A^dagger product runs matrix-free on the device, everything against numpy."""
    from qutip.core import data as _data
    rng = np.random.default_rng(4)
    n = 24
    A = (rng.random((n, n)) < 0.2) * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    rho = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    dA = plugin.B200Operator(_data.to(_data.CSR, _data.Dense(A)))
    drho = plugin.B200Dense(np.asfortranarray(rho))
    out = _data.matmul_dag(drho, dA, 0.5 - 2j)
    assert isinstance(out, plugin.B200Dense)
    np.testing.assert_allclose(out.to_array(), (0.5 - 2j) * rho @ A.conj().T, rtol=1e-12, atol=1e-12)
    B = rng.standard_normal((7, n)) + 1j * rng.standard_normal((7, n))
    out = _data.matmul_dag(plugin.B200Dense(np.asfortranarray(B)), dA, 1.5j)      # non-square left
    np.testing.assert_allclose(out.to_array(), 1.5j * B @ A.conj().T, rtol=1e-12, atol=1e-12)
    C = rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))
    out = _data.matmul_dag(plugin.B200Dense(np.asfortranarray(B)), plugin.B200Dense(np.asfortranarray(C)))
    np.testing.assert_allclose(out.to_array(), B @ C.conj().T, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        _data.matmul_dag(plugin.B200Dense(np.zeros((3, 4), dtype=complex)), dA)
    # the dispatcher picks the device specialisation for the registered types
    assert _data.matmul_dag[plugin.B200Dense, plugin.B200Operator, plugin.B200Dense] is not None


@pytest.mark.parametrize("rate,tight", [("-0.08 + 0.05*sin(2*t)", True), ("0.25*sin(2*t) + 0.05", False)])
def test_nm_mcsolve_b200_map_matches_reference(rate, tight):
    """NonMarkovianMCSolver (solver/nm_mcsolve.py) through the b200 map: rate-shifted collapse
    operators compiled to device coefficient programs (RateShiftCoefficient,
    SqrtRealCoefficient), the influence martingale attached from the collapse records
    (nm_mcsolve.py:562-570) -- same records, expectation values and traces as the reference.

    The second rate changes sign: the shifted rates have kinks at t = 1.67 and 3.04, and behind
    a kink the adaptive step sequence is ill-conditioned -- the REFERENCE run against itself with
    the rate perturbed by 2 ulp (`0.25*sin(2*t)*(1+4e-16) + 0.05`) moves the collapse times of
    trajectories 0, 12 and 15 by 3e-9 ... 3.5e-7 and the expectation values by 5e-8 (measured
    here with oracle/_ref), so that case is compared at the reference's own collapse-time
    resolution norm_t_tol = 1e-6; the smooth (always negative) rate is compared tightly."""
    from qutip import nm_mcsolve, sigmap, coefficient
    H = 0.5 * sigmaz() + 0.2 * sigmax()
    ops_and_rates = [(sigmam(), coefficient(rate)), (sigmap(), 0.15)]
    psi0 = (basis(2, 0) + 0.5 * basis(2, 1)).unit()
    tl = np.linspace(0, 4, 17)
    o = dict(OPT, method="vern7", keep_runs_results=True, store_final_state=True)
    kw = dict(e_ops=[sigmaz(), sigmax()], ntraj=24)
    ref = nm_mcsolve(H, psi0, tl, ops_and_rates, seeds=np.random.SeedSequence(3), options=o, **kw)
    out = nm_mcsolve(H, psi0, tl, ops_and_rates, seeds=np.random.SeedSequence(3),
                     options=dict(o, map="b200"), **kw)
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    assert sum(len(w) for w in ref.col_which) > 5
    t_atol = 1e-9 if tight else 2e-6
    e_rtol, e_atol = (RTOL, ATOL) if tight else (1e-5, 1e-6)
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=t_atol)
    np.testing.assert_allclose(np.array(out.runs_trace), np.array(ref.runs_trace), rtol=1e-9 if tight else 1e-6,
                               atol=1e-12)
    np.testing.assert_allclose(np.array(out.runs_expect), np.array(ref.runs_expect), rtol=e_rtol, atol=e_atol)
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect), rtol=e_rtol, atol=e_atol)
    np.testing.assert_allclose(np.array(out.average_trace), np.array(ref.average_trace),
                               rtol=1e-9 if tight else 1e-6, atol=1e-12)


def test_heom_and_bloch_redfield_solvers_reuse_the_device_integrators():
    """Solver subclasses outside mesolve/mcsolve whose right-hand side is a constant QobjEvo --
    HEOMSolver's hierarchy generator (solver/heom/bofin_solvers.py:699-703, 929-940: one large
    sparse matrix acting on the stacked ADOs) and BRSolver's Bloch-Redfield tensor
    (solver/brmesolve.py:323-333) -- have `b200_*` registered (`add_integrator`) and run
    the same QobjEvo.matmul_data hot path on the device."""
    from qutip.solver.heom import DrudeLorentzBath, HEOMSolver
    H = 0.5 * sigmaz() + 0.25 * sigmax()
    bath = DrudeLorentzBath(sigmaz(), lam=0.05, gamma=0.5, T=1.0, Nk=2)
    rho0 = basis(2, 0) * basis(2, 0).dag()
    tl = np.linspace(0, 5, 21)
    res = {}
    for m in ("vern7", "b200_vern7", "b200_adams"):
        o = dict(OPT, method=m, store_states=True)
        if "adams" not in m:
            o.update(atol=1e-10, rtol=1e-8)
        res[m] = HEOMSolver(H, bath, max_depth=4, options=o).run(rho0, tl, e_ops=[sigmaz(), sigmax()])
    np.testing.assert_allclose(np.array(res["b200_vern7"].expect), np.array(res["vern7"].expect),
                               rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(res["b200_vern7"].states[-1].full(), res["vern7"].states[-1].full(),
                               rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(np.array(res["b200_adams"].expect), np.array(res["vern7"].expect),
                               rtol=1e-4, atol=1e-5)
    from qutip import brmesolve
    a = destroy(6)
    Hb = a.dag() * a + 0.1 * (a + a.dag())
    a_ops = [(a + a.dag(), "0.05 * (w > 0)")]
    psi0 = basis(6, 3)
    tlb = np.linspace(0, 4, 17)
    ref = brmesolve(Hb, psi0, tlb, a_ops, e_ops=[a.dag() * a], options=dict(OPT, method="vern7"))
    out = brmesolve(Hb, psi0, tlb, a_ops, e_ops=[a.dag() * a], options=dict(OPT, method="b200_vern7"))
    np.testing.assert_allclose(out.expect[0], ref.expect[0], rtol=RTOL, atol=ATOL)


def test_nm_mcsolve_mixed_initial_states_b200_map():
    """NonMarkovianMCSolver with a mixed initial state (solver/multitraj.py:285-352 picks the
    state per trajectory, nm_mcsolve.py:562-570 adds the martingale): batched on the device per
    initial state, identical to the reference run with the same seeds."""
    from qutip import NonMarkovianMCSolver, coefficient, sigmap
    H = 0.5 * sigmaz() + 0.2 * sigmax()
    ops_and_rates = [(sigmam(), coefficient("-0.08 + 0.05*sin(2*t)")), (sigmap(), 0.15)]
    ics = [(basis(2, 1), 0.25), ((basis(2, 1) + basis(2, 0)).unit(), 0.75)]
    tl = np.linspace(0, 3, 13)
    res = {}
    for mp in ("serial", "b200"):
        o = dict(OPT, method="vern7", map=mp, keep_runs_results=True)
        solver = NonMarkovianMCSolver(H, ops_and_rates, options=o)
        res[mp] = solver.run(ics, tl, [6, 18], e_ops=[sigmaz(), sigmax()], seeds=np.random.SeedSequence(11))
    ref, out = res["serial"], res["b200"]
    assert [list(w) for w in out.col_which] == [list(w) for w in ref.col_which]
    assert sum(len(w) for w in ref.col_which) > 3
    for a, b in zip(out.col_times, ref.col_times):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.array(out.runs_trace), np.array(ref.runs_trace), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.array(out.average_expect), np.array(ref.average_expect), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.array(out.average_trace), np.array(ref.average_trace), rtol=1e-9, atol=1e-12)


def test_floquet_markov_solver_reuses_the_device_integrators():
    """FMESolver (solver/floquet.py:785-): the Floquet-Markov tensor is a constant QobjEvo, so
    `method="b200_vern7"` integrates it on the device -- same result as the stock vern7."""
    from qutip import fmmesolve, num
    delta, eps0, A = 2 * np.pi, 2 * np.pi, 0.5 * 2 * np.pi
    omega = np.sqrt(delta ** 2 + eps0 ** 2)
    T = 2 * np.pi / omega
    tl = np.linspace(0.0, 2 * T, 41)
    psi0 = (basis(2, 0) + 0.3j * basis(2, 1)).unit()
    H = [-eps0 / 2.0 * sigmaz() - delta / 2.0 * sigmax(), [A / 2.0 * sigmax(), "sin(w * t)"]]

    def spectrum(w):
        return (w > 0) * w * 0.5 * 0.25 / (2 * np.pi)

    res = {}
    for m in ("vern7", "b200_vern7"):
        res[m] = fmmesolve(H, psi0, tl, [sigmax()], [spectrum], T, e_ops=[num(2)], args={"w": omega},
                           options=dict(OPT, method=m)).expect[0]
    np.testing.assert_allclose(res["b200_vern7"], res["vern7"], rtol=RTOL, atol=ATOL)
