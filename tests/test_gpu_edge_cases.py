"""Edge cases of the device path: empty operators, one-element systems, rows that are not a
multiple of the 32-row slices / 256-row tiles, single-point tlists, no collapse operators,
more slots than trajectories, unsorted / duplicate CSR indices
(reference: tests/core/data/test_mathematics.py:159-161, conftest shapes)."""
import numpy as np
import pytest
import scipy.sparse as sp

import qutip_b200 as qb
from qutip_b200 import engine as E, solve

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["vern7", "tsit5", "adams"])
def test_zero_operator_keeps_the_state(method):
    N = 37
    Z = sp.csr_matrix((N, N), dtype=complex)
    y0 = (np.arange(N) + 1j).astype(complex)
    r = solve.mesolve([Z], y0, np.linspace(0, 3, 4), method=method)
    np.testing.assert_allclose(r.states[0], np.tile(y0, (4, 1)), rtol=0, atol=0)


@pytest.mark.parametrize("N", [1, 2, 31, 33, 255, 257])
@pytest.mark.parametrize("fmt", [qb.FMT_AUTO, qb.FMT_CSR, qb.FMT_DIAM, qb.FMT_SELL, qb.FMT_RSELL])
def test_ragged_sizes_decay(N, fmt):
    """dy/dt = -diag(k) y + coupling: sizes around the slice / tile boundaries."""
    rng = np.random.default_rng(N)
    A = sp.diags(-(0.1 + rng.random(N)) - 1j * rng.random(N)).tocsr()
    if N > 1:
        A = A + sp.random(N, N, density=min(1.0, 3.0 / N), random_state=rng, format="csr") * 0.2j
    A = sp.csr_matrix(A)
    y0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    system = solve.build_system([A], fmt=fmt)
    eng = qb.Engine(system, "vern7", nslots=1, store_states=1)
    r = eng.run_mesolve(y0, np.array([0.0, 0.7, 1.5]))
    assert r.status[0] == 1
    import scipy.linalg
    ref = scipy.linalg.expm(A.toarray() * 1.5) @ y0
    np.testing.assert_allclose(r.states[0, -1], ref, rtol=1e-6, atol=1e-8)


def test_single_point_tlist_and_more_slots_than_trajectories():
    N = 20
    A = sp.diags(-np.linspace(0.1, 1, N)).tocsr().astype(complex)
    y0 = np.ones(N, dtype=complex)
    r = solve.mesolve([A], y0, np.array([0.5]), e_ops=[])
    assert r.states.shape == (1, 1, N)
    np.testing.assert_array_equal(r.states[0, 0], y0)
    r = solve.mesolve([A], np.stack([y0, 2 * y0, 3 * y0]), np.array([0.0, 1.0]), nslots=64)
    np.testing.assert_allclose(r.states[:, -1, :], np.outer([1, 2, 3], np.exp(-np.linspace(0.1, 1, N))),
                               rtol=1e-6, atol=1e-8)


def test_mcsolve_without_jumps_is_the_no_jump_evolution():
    """With thresholds of zero no trajectory ever jumps: the result is exp(-i H_eff t) psi,
    normalised (mcsolve.py:286-302)."""
    N = 8
    rng = np.random.default_rng(2)
    H = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    H = H + H.conj().T
    c = np.diag(np.sqrt(np.linspace(0.1, 0.5, N))).astype(complex)
    heff = sp.csr_matrix(-1j * H - 0.5 * c.conj().T @ c)
    psi0 = np.zeros(N, dtype=complex); psi0[0] = 1
    draws = np.zeros((3, 8))
    r = solve.mcsolve([heff], [sp.csr_matrix(c)], psi0, np.array([0.0, 0.4, 0.8]), ntraj=3,
                      draws=draws, e_ops=[sp.csr_matrix(np.diag(np.arange(N, dtype=complex)))],
                      options=dict(store_states=1))
    assert (r.ncol == 0).all()
    import scipy.linalg
    psi = scipy.linalg.expm(heff.toarray() * 0.8) @ psi0
    psi /= np.linalg.norm(psi)
    np.testing.assert_allclose(r.states[0, -1], psi, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(r.runs_expect[0, :, -1].real, (np.arange(N) * np.abs(psi) ** 2).sum(), rtol=1e-6)


def test_unsorted_and_duplicate_csr_indices_all_formats():
    data = np.array([1.0, 2.0, 3.0, 4.0, 5.0, -1.0], dtype=complex)
    col = np.array([2, 0, 2, 1, 1, 0], dtype=np.int32)          # unsorted, duplicate (0,2)
    rowptr = np.array([0, 3, 4, 6], dtype=np.int32)
    dense = np.array([[2, 0, 4], [0, 4, 0], [-1, 5, 0]], dtype=complex)
    x = np.array([1.0, 10.0, 100.0], dtype=complex)
    for fmt in (qb.FMT_AUTO, qb.FMT_CSR, qb.FMT_DIAM, qb.FMT_SELL, qb.FMT_RSELL):
        op = qb.DeviceOp.from_csr(data, col, rowptr, (3, 3), fmt)
        out = E.matmul(op, qb.DeviceDense.from_numpy(x))
        np.testing.assert_allclose(out.to_numpy().ravel(), dense @ x)


def test_bad_arguments_are_reported():
    with pytest.raises(Exception):
        qb.DeviceOp.from_csr(np.ones(2, dtype=complex), np.array([0, 5], dtype=np.int32),
                             np.array([0, 1, 2], dtype=np.int32), (2, 2))      # column out of range
    with pytest.raises(ValueError):
        qb.Engine(qb.System(4), "rk45")
    s = qb.System(4)
    s.add_element(qb.DeviceOp.from_scipy(sp.identity(4, dtype=complex, format="csr")))
    eng = qb.Engine(s, "vern7", nslots=1)
    with pytest.raises(ValueError):
        eng.run_mesolve(np.ones(5, dtype=complex), np.array([0.0, 1.0]))
