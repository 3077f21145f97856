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
    # rule-compressed slices (xor key): constant diagonals of the Hamiltonian part cost nothing,
    # 60 MB instead of 402 MB (DIAM) / 497 MB (CSR) are read per product
    assert info["format"] == "rsell" and info["nnz"] == L.nnz and info["device_bytes"] < 80e6
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


def test_c2_matrix_free_rhs_equals_liouvillian(c2):
    """C2 at full size: the matrix-free right-hand side (Kronecker operators + jump part, both
    the explicit and the sandwich form) against the 24.6 M-nnz Liouvillian on the same rho."""
    L, _ = c2
    H, c_ops, _ = models.tfim(10)
    N = L.shape[0]
    rng = np.random.default_rng(5)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    ref = L @ x
    dx = qb.DeviceDense.from_numpy(x)
    for jump in ("explicit", "sandwich"):
        system = qb.System(N)
        for op, prog in solve.lindblad_matrix_free([H], c_ops, jump=jump):
            system.add_element(op, prog)
        eng = qb.Engine(system, "vern7", nslots=1)
        out = qb.DeviceDense.zeros(N, 1)
        eng.rhs(0.0, dx, out)
        np.testing.assert_allclose(out.to_numpy().ravel(), ref, rtol=1e-12, atol=1e-11)


def test_matrix_free_mesolve_12_spins():
    """Beyond what a Liouvillian allows on the host: dissipative TFIM with 12 spins (rho is
    4096 x 4096 = 268 MB; L would have 4.5e8 non-zeros).  The right-hand side is checked on
    sampled columns / rows against the n x n operators applied with scipy, and a short
    mesolve must preserve trace and hermiticity and reproduce the product-state short-time
    expansion of <sz_0>."""
    nsp = 12
    H, c_ops, sz = models.tfim(nsp)
    n = H.shape[0]
    N = n * n
    els = solve.lindblad_matrix_free([H], c_ops)
    assert [o.info()["format"] for o, _ in els] == ["kron", "kron", "kron"]   # sandwich jumps (auto)
    assert sum(o.info()["device_bytes"] for o, _ in els) < 8e6
    system = qb.System(N)
    for op, prog in els:
        system.add_element(op, prog)
    system.add_eop(qb.DeviceOp.from_scipy(solve.trace_functional(sz[0])))
    system.set_functional(True)
    eng = qb.Engine(system, "vern7", nslots=1, store_states=1)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    out = qb.DeviceDense.zeros(N, 1)
    eng.rhs(0.0, qb.DeviceDense.from_numpy(x), out)
    z = out.to_numpy().reshape(n, n, order="F")
    rho = x.reshape(n, n, order="F")
    Hn = H.copy().astype(complex)
    for c in c_ops:
        Hn = Hn - 0.5j * (c.conj().T @ c)
    Hn = Hn.tocsr()
    cols = [0, 1, 777, 4095]
    # column j of  -i Hn rho + i rho Hn^dag + sum C rho C^dag
    rhoHd = (Hn.conj() @ rho.T).T[:, cols]                 # (rho Hn^dag)[:, cols]
    jump = np.zeros((n, len(cols)), dtype=complex)
    for c in c_ops:
        cr = c @ rho                                        # C rho
        jump += (c.conj() @ cr.T).T[:, cols]                # (C rho C^dag)[:, cols]
    ref = -1j * (Hn @ rho[:, cols]) + 1j * rhoHd + jump
    np.testing.assert_allclose(z[:, cols], ref, rtol=1e-11, atol=1e-10)
    # short evolution from |0...0><0...0|
    y0 = np.zeros(N, dtype=complex)
    y0[0] = 1.0
    tl = np.linspace(0, 0.05, 3)
    r = eng.run_mesolve(y0, tl)
    assert (r.status == 1).all()
    rho_t = r.states[0, -1].reshape(n, n, order="F")
    assert abs(np.trace(rho_t) - 1.0) < 1e-9
    assert np.max(np.abs(rho_t - rho_t.conj().T)) < 1e-10
    # <sz_0>(t) = 1 - 2 gamma t - 2 t^2 + O(t^3): sigma^- (gamma = 0.1) flips the spin out of
    # basis state 0 (sz = +1), -sx makes it precess at frequency 2
    t = tl[-1]
    assert abs(r.expect[0, 0, -1].real - (1 - 2 * 0.1 * t - 2 * t * t)) < 3e-4
