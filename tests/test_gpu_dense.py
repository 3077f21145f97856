"""Dense batched path: complex128 ZGEMM on the FP64 tensor cores (DMMA) and the dense-H_eff
mcsolve block that uses it (BASELINE config 5)."""
import numpy as np
import pytest
import scipy.sparse as sp

import qutip_b200 as qb
from qutip_b200 import engine as E
from oracle.rk_oracle import OrcEvo, OrcOp, mcsolve_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,k,n", [(64, 64, 8), (100, 70, 33), (256, 256, 256), (130, 257, 65)])
def test_zgemm_dmma_vs_numpy(m, k, n):
    rng = np.random.default_rng(m + k + n)
    A = rng.random((m, k)) + 1j * rng.random((m, k))
    X = rng.random((k, n)) + 1j * rng.random((k, n))
    C0 = rng.random((m, n)) + 1j * rng.random((m, n))
    dA = qb.DeviceDense.from_numpy(np.asfortranarray(A))
    dX = qb.DeviceDense.from_numpy(np.asfortranarray(X))
    dC = qb.DeviceDense.from_numpy(np.asfortranarray(C0))
    E.matmul(dA, dX, 0.5 - 0.25j, dC)           # routed to qb_zgemm (>= 8 columns)
    ref = C0 + (0.5 - 0.25j) * (A @ X)
    np.testing.assert_allclose(dC.to_numpy(), ref, rtol=1e-12, atol=1e-12)


def _dense_mc_case(dim, ntraj, seed):
    rng = np.random.default_rng(seed)
    H = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    H = (H + H.conj().T) / np.sqrt(dim)
    cs = []
    for k in range(2):
        c = sp.random(dim, dim, 4.0 / dim, random_state=seed + k, dtype=float) * 0.7
        cs.append(sp.csr_matrix(c).astype(complex))
    heff = -1j * H - 0.5 * sum((c.conj().T @ c).toarray() for c in cs)
    psi0 = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    psi0 /= np.linalg.norm(psi0)
    e = sp.diags(rng.random(dim)).tocsr().astype(complex)
    draws = rng.random((ntraj, 64))
    return heff, cs, psi0, e, draws


@pytest.mark.parametrize("dim,ntraj", [(64, 24), (100, 9)])
def test_dense_heff_mcsolve_block_vs_oracle(dim, ntraj):
    heff, cs, psi0, e, draws = _dense_mc_case(dim, ntraj, 3)
    tlist = np.linspace(0, 1.5, 7)
    s = qb.System(dim)
    s.add_element(qb.DeviceDense.from_numpy(np.asfortranarray(heff)))
    for c in cs:
        s.add_collapse(qb.DeviceOp.from_scipy(c), qb.DeviceOp.from_scipy(sp.csr_matrix(c.conj().T @ c)))
    s.add_eop(qb.DeviceOp.from_scipy(e))
    eng = qb.Engine(s, "vern7", nslots=16)          # >= 8 slots -> DMMA ZGEMM pre-pass
    r = eng.run_mcsolve(psi0, tlist, draws)
    assert (r.status == 1).all()
    rhs = OrcEvo([(OrcOp.dense(heff), 1.0)])
    ocs = [OrcEvo([(OrcOp.from_scipy(c), 1.0)]) for c in cs]
    ons = [OrcEvo([(OrcOp.from_scipy(sp.csr_matrix(c.conj().T @ c)), 1.0)]) for c in cs]
    for j in range(ntraj):
        o = mcsolve_oracle(rhs, ocs, ons, psi0, tlist, draws[j], [OrcOp.from_scipy(e)])
        assert r.ncol[j] == len(o["collapses"])
        assert list(r.col_which[j, :r.ncol[j]]) == [w for _, w in o["collapses"]]
        np.testing.assert_allclose(r.expect[j], o["expect"], rtol=1e-6, atol=1e-8)
    # the same system through the per-row zgemv path (fewer than 8 slots) agrees
    eng1 = qb.Engine(s, "vern7", nslots=4)
    r1 = eng1.run_mcsolve(psi0, tlist, draws)
    assert np.array_equal(r1.ncol, r.ncol)
    np.testing.assert_allclose(r1.expect, r.expect, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("dim,ncols", [(1024, 256), (4096, 256)])
def test_zgemm_dmma_baseline_dims_vs_numpy(dim, ncols):
    """BASELINE config 5 sizes (dense H_eff dim 1024 / 4096 times a block of trajectories):
    every output element against numpy's zgemm (matmul_dense, core/data/matmul.pyx:275-347)."""
    rng = np.random.default_rng(dim)
    A = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) / np.sqrt(dim)
    X = rng.standard_normal((dim, ncols)) + 1j * rng.standard_normal((dim, ncols))
    dA = qb.DeviceDense.from_numpy(np.asfortranarray(A))
    dX = qb.DeviceDense.from_numpy(np.asfortranarray(X))
    dC = qb.DeviceDense.zeros(dim, ncols)
    E.matmul(dA, dX, 1.0, dC)
    ref = A @ X
    # accumulation order differs (k-blocked DMMA vs BLAS): 1e-6 / 1e-8 is the north-star
    # tolerance, observed ~1e-13
    np.testing.assert_allclose(dC.to_numpy(), ref, rtol=1e-6, atol=1e-8)
    assert np.abs(dC.to_numpy() - ref).max() < 1e-11
