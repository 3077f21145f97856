"""Matrix-free Lindblad right-hand side (LindbladMatrixForm on the device):
the Kronecker operators I (x) A and conj(A) (x) I against explicit scipy Kronecker products,
and whole mesolve runs against the fused-superoperator path that the golden fixtures pin."""
import numpy as np
import pytest
import scipy.sparse as sp

import qutip_b200 as qb
from qutip_b200 import coeffs, engine as E, models, solve

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-8, 1e-6


def _random_op(n, density, seed):
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, density=density, random_state=rng, format="csr", dtype=float)
    m = m + 1j * sp.random(n, n, density=density, random_state=rng, format="csr", dtype=float)
    m = sp.csr_matrix(m)
    m.sort_indices()
    return m


@pytest.mark.parametrize("n", [5, 30, 32, 96])
@pytest.mark.parametrize("side", [0, 1])
def test_kron_matmul_vs_scipy(n, side):
    A = _random_op(n, 0.25, 10 * n + side)
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    I = sp.identity(n, dtype=complex, format="csr")
    K = sp.kron(I, A) if side == 0 else sp.kron(A.conj(), I)
    op = qb.DeviceOp.kron(A, side)
    assert op.shape == (n * n, n * n)
    out = qb.DeviceDense.from_numpy(np.ones(n * n, dtype=complex))
    E.matmul(op, qb.DeviceDense.from_numpy(x), 0.5 - 2j, out)
    ref = 1.0 + (0.5 - 2j) * (K @ x)
    np.testing.assert_allclose(out.to_numpy().ravel(), ref, rtol=1e-12, atol=1e-12)
    # as rho -> A rho / rho A^dagger on the un-stacked matrix
    rho = x.reshape(n, n, order="F")
    mat = A.toarray() @ rho if side == 0 else rho @ A.toarray().conj().T
    np.testing.assert_allclose((K @ x).reshape(n, n, order="F"), mat, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n", [6, 32, 40])
def test_sandwich_matmul_vs_scipy(n):
    ops = [_random_op(n, 0.15, 7 * n + k) for k in range(3)]
    ops.append(sp.csr_matrix((n, n), dtype=complex))          # an empty operator in the stack
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    op = qb.DeviceOp.sandwich(ops)
    out = qb.DeviceDense.from_numpy(np.zeros(n * n, dtype=complex))
    E.matmul(op, qb.DeviceDense.from_numpy(x), 1.0, out)
    rho = x.reshape(n, n, order="F")
    ref = sum(c.toarray() @ rho @ c.toarray().conj().T for c in ops)
    np.testing.assert_allclose(out.to_numpy().reshape(n, n, order="F"), ref, rtol=1e-12, atol=1e-12)
    assert op.info()["nnz"] == sum(c.nnz ** 2 for c in ops)


def test_kron_rejects_bad_input():
    with pytest.raises(ValueError):
        qb.DeviceOp.kron(sp.csr_matrix((3, 4), dtype=complex), 0)
    A = _random_op(8, 0.3, 1)
    op = qb.DeviceOp.kron(A, 0)
    x = qb.DeviceDense.from_numpy(np.ones(10, dtype=complex))
    out = qb.DeviceDense.from_numpy(np.ones(64, dtype=complex))
    with pytest.raises(Exception, match="incompatible matrix shapes"):
        E.matmul(op, x, 1.0, out)


@pytest.mark.parametrize("jump", ["explicit", "sandwich"])
@pytest.mark.parametrize("nspins", [3, 5, 6])
def test_mesolve_matrix_free_vs_superoperator_tfim(nspins, jump):
    """n = 8 (per-lane CSR products), n = 32 / 64 (SELL left product, warp-uniform right)."""
    H, c_ops, sz = models.tfim(nspins)
    n = H.shape[0]
    rho0 = np.zeros((n, n), dtype=complex)
    rho0[0, 0] = 1.0
    y0 = rho0.reshape(-1, order="F")
    tl = np.linspace(0, 1.5, 7)
    ref = solve.mesolve([models.liouvillian(H, c_ops)], y0, tl, e_ops=[sz[0], sz[-1]])
    out = solve.mesolve(solve.lindblad_matrix_free([H], c_ops, jump=jump), y0, tl,
                        e_ops=[sz[0], sz[-1]])
    np.testing.assert_allclose(out.expect, ref.expect, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out.states, ref.states, rtol=RTOL, atol=ATOL)
    # same controller decisions: the two right-hand sides differ at round-off only
    assert list(out.stats[0]) == list(ref.stats[0])


@pytest.mark.parametrize("jump", ["explicit", "sandwich"])
def test_mesolve_matrix_free_time_dependent(jump):
    """C4-like driven cavity x transmon (n = 30, not a multiple of 32) with a cos drive and a
    time-dependent collapse rate."""
    H0, H1, c_ops, a, b = models.driven_cavity_transmon(10)
    n = H0.shape[0]
    drive = coeffs.compile_expr("0.3 * cos(4.8 * t)")
    rate = coeffs.compile_expr("exp(-0.2 * t)")
    L0 = models.liouvillian(H0, c_ops[1:])
    I = sp.identity(n, dtype=complex, format="csr")
    L1 = -1j * (sp.kron(I, H1) - sp.kron(H1.T, I))
    c = c_ops[0]
    cdc = c.conj().T @ c
    Lc = sp.kron(c.conj(), c) - 0.5 * sp.kron(I, cdc) - 0.5 * sp.kron(cdc.T, I)
    psi = np.zeros(n, dtype=complex)
    psi[3 * 2 + 1] = 1.0
    y0 = np.outer(psi, psi.conj()).reshape(-1, order="F")
    tl = np.linspace(0, 2, 9)
    e_ops = [sp.csr_matrix(a.conj().T @ a), sp.csr_matrix(b.conj().T @ b)]
    ref = solve.mesolve([(sp.csr_matrix(L1), drive), (sp.csr_matrix(Lc), rate.norm()), L0],
                        y0, tl, e_ops=e_ops)
    els = solve.lindblad_matrix_free([(H1, drive), H0], [(c, rate)] + list(c_ops[1:]), jump=jump)
    out = solve.mesolve(els, y0, tl, e_ops=e_ops)
    np.testing.assert_allclose(out.expect, ref.expect, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out.states, ref.states, rtol=RTOL, atol=ATOL)
