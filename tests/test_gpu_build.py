"""Liouvillian assembled on the device (qb_liouvillian_build) against the reference's
construction: qutip.liouvillian (core/superoperator.py:116-142) where the reference build
travels with the repo, the scipy restatement in qutip_b200.models otherwise -- structure
(canonical CSR, same pattern) and values."""
import sys
import time

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import qutip_b200 as qb
from qutip_b200 import engine as E, models

pytestmark = pytest.mark.gpu


def _rand_sparse(rng, n, density):
    m = sp.random(n, n, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))),
                  format="csr", dtype=float)
    m = m.astype(complex)
    m.data = m.data + 1j * rng.standard_normal(m.nnz)
    return sp.csr_matrix(m)


def _same(got, want, rtol=1e-13):
    got = sp.csr_matrix(got); want = sp.csr_matrix(want)
    got.sort_indices(); want.sum_duplicates(); want.eliminate_zeros(); want.sort_indices()
    assert got.shape == want.shape
    diff = (got - want)
    scale = max(1e-300, np.abs(want.data).max() if want.nnz else 1.0)
    assert (np.abs(diff.data).max() if diff.nnz else 0.0) <= rtol * scale
    # canonical: sorted, no duplicates
    for r in range(0, got.shape[0], max(1, got.shape[0] // 64)):
        c = got.indices[got.indptr[r]:got.indptr[r + 1]]
        assert (np.diff(c) > 0).all()


@pytest.mark.parametrize("n,nc,density", [(2, 0, 1.0), (3, 1, 1.0), (5, 2, 0.6), (12, 3, 0.3), (33, 4, 0.1),
                                          (64, 1, 0.05)])
def test_liouvillian_matches_restatement(n, nc, density):
    rng = np.random.default_rng(100 + n)
    H = _rand_sparse(rng, n, density)
    H = sp.csr_matrix(H + H.conj().T)
    c_ops = [_rand_sparse(rng, n, density) for _ in range(nc)]
    want = models.liouvillian(H, c_ops)
    op = qb.DeviceOp.liouvillian(H, c_ops, qb.FMT_CSR)
    got = op.to_scipy()
    _same(got, want)
    want.eliminate_zeros()           # scipy's kron (block format) stores explicit zeros
    assert got.nnz == want.nnz and np.array_equal(got.indices, want.indices)


def test_liouvillian_one_by_one_cancels():
    # n = 1: |c|^2 - |c|^2 cancels; scipy's sum order may leave rounding noise, the reference's
    # tidy-up (and ours) drops it
    c = sp.csr_matrix(np.array([[0.3 - 0.7j]]))
    got = qb.DeviceOp.liouvillian(sp.csr_matrix(np.array([[2.0 + 0j]])), [c], qb.FMT_CSR, tol=1e-14).to_scipy()
    assert got.shape == (1, 1) and got.nnz == 0


def test_liouvillian_dissipator_only_and_errors():
    rng = np.random.default_rng(5)
    c_ops = [_rand_sparse(rng, 7, 0.5) for _ in range(2)]
    want = models.liouvillian(sp.csr_matrix((7, 7), dtype=complex), c_ops)
    _same(qb.DeviceOp.liouvillian(None, c_ops, qb.FMT_CSR).to_scipy(), want)
    with pytest.raises(ValueError):
        qb.DeviceOp.liouvillian(None, [])
    with pytest.raises(ValueError):
        qb.DeviceOp.liouvillian(sp.identity(3), [sp.identity(4)])
    with pytest.raises(qb.QbError):          # not a CSR-format operator
        qb.DeviceOp.liouvillian(sp.identity(64, dtype=complex, format="csr"), [], qb.FMT_RSELL).to_scipy()


def test_liouvillian_against_reference_constructor():
    ref = oracle.ref_path()
    if ref is None:
        pytest.skip("reference build (oracle/_ref) not present")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import qutip
    N = 6
    a = qutip.tensor(qutip.destroy(N), qutip.qeye(2))
    sm = qutip.tensor(qutip.qeye(N), qutip.sigmam())
    H = a.dag() * a + 0.5 * sm.dag() * sm + 0.3 * (a.dag() * sm + a * sm.dag())
    c_ops = [np.sqrt(0.2) * a, np.sqrt(0.05) * sm, np.sqrt(0.01) * a.dag()]
    want = qutip.liouvillian(H, c_ops).to("CSR").data.as_scipy()
    got = qb.DeviceOp.liouvillian(H.to("CSR").data.as_scipy(), [c.to("CSR").data.as_scipy() for c in c_ops],
                                  qb.FMT_CSR, tol=qutip.settings.core["auto_tidyup_atol"]).to_scipy()
    _same(got, want)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)


def test_liouvillian_compressed_format_same_product():
    H, c_ops, _ = models.tfim(6)
    want = models.liouvillian(H, c_ops)
    op = qb.DeviceOp.liouvillian(H, c_ops)            # auto format: host slice analysers
    N = want.shape[0]
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y = E.matmul(op, qb.DeviceDense.from_numpy(x)).to_numpy().ravel()
    ref = want @ x
    assert np.abs(y - ref).max() < 1e-13 * np.abs(ref).max()
    assert op.info()["nnz"] == want.nnz


def test_c2_liouvillian_full_size_on_device():
    """C2: 2^20 x 2^20, 24.6 M non-zeros -- assembled on the device, bit-for-bit pattern of the
    host construction, values within rounding of the different summation order."""
    H, c_ops, _ = models.tfim(10)
    t0 = time.perf_counter()
    op = qb.DeviceOp.liouvillian(H, c_ops, qb.FMT_CSR)
    t_dev = time.perf_counter() - t0
    info = op.info()
    t0 = time.perf_counter()
    want = models.liouvillian(H, c_ops)
    t_host = time.perf_counter() - t0
    want.eliminate_zeros()
    assert info["rows"] == 2 ** 20 and info["nnz"] == want.nnz and abs(want.nnz - 24641535) < 8
    got = op.to_scipy()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.abs(got.data - want.data).max() <= 1e-14 * np.abs(want.data).max()
    print("C2 Liouvillian assembly: device %.3f s, host (scipy kron/add) %.3f s" % (t_dev, t_host))
    assert t_dev < t_host


@pytest.mark.parametrize("n,density", [(7, 0.5), (32, 0.2), (100, 0.05), (257, 0.02)])
def test_csr_to_diam_on_device_matches_host_analyser(n, density):
    rng = np.random.default_rng(n)
    m = _rand_sparse(rng, n, density) + sp.diags(rng.standard_normal(n) + 0j, 0) \
        + sp.diags(rng.standard_normal(n - 3) + 0j, 3)
    m = sp.csr_matrix(m)
    m.sum_duplicates(); m.sort_indices()
    host = qb.DeviceOp.from_scipy(m, qb.FMT_DIAM)
    dev = qb.DeviceOp.from_scipy(m, qb.FMT_CSR).convert(qb.FMT_DIAM)
    hi, di = host.info(), dev.info()
    assert di["format"] == "diam" and di["nnz"] == hi["nnz"] == m.nnz
    assert abs(di["device_bytes"] - hi["device_bytes"]) <= 16
    x = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    dx = qb.DeviceDense.from_numpy(np.asfortranarray(x))
    yh = E.matmul(host, dx).to_numpy()
    yd = E.matmul(dev, dx).to_numpy()
    assert np.array_equal(yh, yd)                       # same entries in the same order: bit-equal
    np.testing.assert_allclose(yd, m @ x, rtol=1e-13, atol=1e-13)
    # through the host analysers: RSELL / SELL / auto
    for fmt in (qb.FMT_SELL, qb.FMT_RSELL, qb.FMT_AUTO):
        y = E.matmul(qb.DeviceOp.from_scipy(m, qb.FMT_CSR).convert(fmt), dx).to_numpy()
        np.testing.assert_allclose(y, m @ x, rtol=1e-13, atol=1e-13)
    with pytest.raises(qb.QbError):
        host.convert(qb.FMT_CSR)


def test_c2_liouvillian_device_build_to_diam_never_visits_host():
    H, c_ops, _ = models.tfim(8)
    want = models.liouvillian(H, c_ops)
    want.eliminate_zeros()
    op = qb.DeviceOp.liouvillian(H, c_ops, qb.FMT_DIAM)
    assert op.info()["format"] == "diam" and op.info()["nnz"] == want.nnz
    N = want.shape[0]
    rng = np.random.default_rng(2)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y = E.matmul(op, qb.DeviceDense.from_numpy(x)).to_numpy().ravel()
    ref = want @ x
    assert np.abs(y - ref).max() < 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("sa,sb", [((3, 4), (5, 2)), ((1, 1), (7, 7)), ((6, 6), (1, 3)), ((40, 33), (9, 20))])
def test_kron_product_on_device(sa, sb):
    rng = np.random.default_rng(sa[0] * 100 + sb[0])

    def rnd(shape):
        m = sp.random(shape[0], shape[1], density=0.3, format="csr", dtype=float,
                      random_state=np.random.RandomState(int(rng.integers(1 << 30)))).astype(complex)
        m.data = m.data + 1j * rng.standard_normal(m.nnz)
        return sp.csr_matrix(m)

    a, b = rnd(sa), rnd(sb)
    want = sp.kron(a, b, format="csr")
    want.sum_duplicates(); want.sort_indices()
    got = qb.DeviceOp.kron_product(a, b).to_scipy()
    assert got.shape == want.shape and got.nnz == want.nnz
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=4e-16, atol=0)
    ref = oracle.ref_path()
    if ref is not None:                                  # the reference's own kron_csr: bit-equal
        if ref not in sys.path:
            sys.path.insert(0, ref)
        from qutip.core import data as _data
        q = _data.kron(_data.CSR(a), _data.CSR(b)).as_scipy()
        q.sort_indices()
        assert np.array_equal(q.indices, got.indices) and np.array_equal(q.data, got.data)
    if min(want.shape) >= 1 and want.shape[0] == want.shape[1]:
        x = rng.standard_normal(want.shape[1]) + 0j
        y = E.matmul(qb.DeviceOp.kron_product(a, b, qb.FMT_DIAM), qb.DeviceDense.from_numpy(x)).to_numpy().ravel()
        np.testing.assert_allclose(y, want @ x, rtol=1e-13, atol=1e-13)


def test_kron_product_empty_and_errors():
    z = sp.csr_matrix((4, 4), dtype=complex)
    got = qb.DeviceOp.kron_product(z, sp.identity(3, dtype=complex, format="csr")).to_scipy()
    assert got.shape == (12, 12) and got.nnz == 0
    with pytest.raises(qb.QbError):
        qb.DeviceOp.kron_product(sp.identity(70000, dtype=complex, format="csr"),
                                 sp.identity(70000, dtype=complex, format="csr"))
