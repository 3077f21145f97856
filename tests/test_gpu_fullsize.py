"""Parity at BASELINE.json's full sizes: oracle comparisons where the oracle finishes in
seconds, size-independent properties (linearity, trace / hermiticity preservation, format
independence) otherwise."""
import numpy as np
import pytest

import qutip_b200 as qb
from qutip_b200 import models, solve
from qutip_b200 import engine as E
from oracle.rk_oracle import OrcEvo, OrcOp, mcsolve_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    H, c_ops, sz = models.tfim(10)
    L = models.liouvillian(H, c_ops)
    return L, sz


def test_c2_liouvillian_spmv_full_size(c2):
    L, _ = c2
    N = L.shape[0]
    assert N == 2 ** 20 and L.nnz == 24641535
    op = qb.DeviceOp.from_scipy(L)
    info = op.info()
    assert info["format"] == "diam" and info["nnz"] == L.nnz
    system = qb.System(N)
    system.add_element(op)
    eng = qb.Engine(system, "vern7", nslots=1)
    rng = np.random.default_rng(0)
    x = rng.random(N) + 1j * rng.random(N)
    y = rng.random(N) + 1j * rng.random(N)
    dx, dy = qb.DeviceDense.from_numpy(x), qb.DeviceDense.from_numpy(y)
    ox, oy, oz = (qb.DeviceDense.zeros(N, 1) for _ in range(3))
    eng.rhs(0.0, dx, ox)
    # oracle (C restatement of matmul_csr_vector.cpp) on the same input
    ref = OrcOp.from_scipy(L).matvec(x)
    got = ox.to_numpy().ravel()
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
    # linearity
    a, b = 0.3 - 0.7j, -1.1 + 0.2j
    eng.rhs(0.0, dy, oy)
    dz = qb.DeviceDense.from_numpy(a * x + b * y)
    eng.rhs(0.0, dz, oz)
    lin = a * got + b * oy.to_numpy().ravel()
    assert np.abs(oz.to_numpy().ravel() - lin).max() < 1e-11 * np.abs(lin).max()
    # forced CSR kernel agrees with the DIAM kernel
    s2 = qb.System(N)
    s2.add_element(qb.DeviceOp.from_scipy(L, qb.FMT_CSR))
    o2 = qb.DeviceDense.zeros(N, 1)
    qb.Engine(s2, "vern7", nslots=1).rhs(0.0, dx, o2)
    assert np.abs(o2.to_numpy().ravel() - got).max() < 1e-12 * np.abs(ref).max()


def test_c2_mesolve_preserves_trace_and_hermiticity(c2):
    L, sz = c2
    N, n = L.shape[0], 1024
    els = [L]
    rho0 = np.zeros(N, dtype=complex); rho0[0] = 1.0
    r = solve.mesolve(els, rho0, np.linspace(0, 0.3, 4), e_ops=[sz[0], np.eye(n)],
                      store_states=True)
    for k in range(4):
        rho = r.states[0, k].reshape(n, n, order="F")
        assert abs(np.trace(rho) - 1.0) < 1e-9
        assert np.abs(rho - rho.conj().T).max() < 1e-9
        assert np.diag(rho).real.min() > -1e-9
    np.testing.assert_allclose(r.expect[0, 1].real, 1.0, atol=1e-9)     # tr(rho) functional
    assert r.expect[0, 0, 0].real == pytest.approx(1.0)


def test_c3_full_size_trajectories_vs_oracle():
    n = 14
    H, c_ops, sz = models.tfim(n)
    heff = models.heff(H, c_ops)
    psi0 = models.basis_state(n)
    tlist = np.linspace(0, 2, 21)
    ntraj = 48
    draws = solve.make_thresholds(7, ntraj, 64)
    res = solve.mcsolve([heff], c_ops, psi0, tlist, ntraj, e_ops=[sz[0]], draws=draws, nslots=32)
    assert res.ncol.sum() > 20
    rhs = OrcEvo([(OrcOp.from_scipy(heff), 1.0)])
    ocs = [OrcEvo([(OrcOp.from_scipy(c), 1.0)]) for c in c_ops]
    ons = [OrcEvo([(OrcOp.from_scipy((c.conj().T @ c).tocsr()), 1.0)]) for c in c_ops]
    for j in (0, 5, 17, 40):
        o = mcsolve_oracle(rhs, ocs, ons, psi0, tlist, draws[j], [OrcOp.from_scipy(sz[0])])
        assert res.ncol[j] == len(o["collapses"])
        assert list(res.col_which[j]) == [w for _, w in o["collapses"]]
        np.testing.assert_allclose(res.col_times[j], [t for t, _ in o["collapses"]], atol=1e-9)
        np.testing.assert_allclose(res.runs_expect[0, j], o["expect"][0], rtol=1e-6, atol=1e-8)
        assert res.stats[j, 0] == rhs.nevals
        rhs.nevals = 0
    # statistics of the batch: |<sz>| <= 1, average decays from +1
    assert np.abs(res.runs_expect).max() <= 1 + 1e-9
    assert res.average_expect[0, 0].real == pytest.approx(1.0)
