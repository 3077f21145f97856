"""Synthetic systems of BASELINE.json's configs, built with scipy.sparse only (no QuTiP
needed on the GPU box).  tests/test_models.py checks them against the reference's own
constructors (liouvillian / tensor / destroy) when the reference build is importable."""
import numpy as np
import scipy.sparse as sp

_sx = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=complex))
_sz = sp.csr_matrix(np.array([[1, 0], [0, -1]], dtype=complex))
_sm = sp.csr_matrix(np.array([[0, 0], [1, 0]], dtype=complex))     # qutip.sigmam()


def _site(op, i, n):
    """op on site i of n spins; site 0 is the left-most tensor factor (qutip.tensor)."""
    return sp.kron(sp.kron(sp.identity(2 ** i, dtype=complex), op),
                   sp.identity(2 ** (n - i - 1), dtype=complex), format="csr")


def tfim(n, gamma=0.1):
    """H = -sum sz_i sz_{i+1} - sum sx_i, c_i = sqrt(gamma) sigma^-_i."""
    H = sp.csr_matrix((2 ** n, 2 ** n), dtype=complex)
    for i in range(n - 1):
        H = H - _site(_sz, i, n) @ _site(_sz, i + 1, n)
    for i in range(n):
        H = H - _site(_sx, i, n)
    c_ops = [np.sqrt(gamma) * _site(_sm, i, n) for i in range(n)]
    sz = [_site(_sz, i, n) for i in range(n)]
    return sp.csr_matrix(H), c_ops, sz


def liouvillian(H, c_ops):
    """Column-stacked Lindblad superoperator, as qutip.liouvillian builds it
    (core/superoperator.py:116-142): vec(A rho B) = (B^T kron A) vec(rho)."""
    n = H.shape[0]
    I = sp.identity(n, dtype=complex, format="csr")
    L = -1j * (sp.kron(I, H) - sp.kron(H.T, I))
    for c in c_ops:
        cdc = c.conj().T @ c
        L = L + sp.kron(c.conj(), c) - 0.5 * sp.kron(I, cdc) - 0.5 * sp.kron(cdc.T, I)
    L = sp.csr_matrix(L)
    L.sum_duplicates()
    L.sort_indices()
    return L


def heff(H, c_ops):
    """mcsolve's -i H - 1/2 sum c^dag c (solver/mcsolve.py:493-496), constants merged."""
    out = -1j * H
    for c in c_ops:
        out = out - 0.5 * (c.conj().T @ c)
    out = sp.csr_matrix(out)
    out.sum_duplicates()
    out.sort_indices()
    return out


def basis_state(n_spins):
    """basis([2]*n, [0]*n)"""
    psi = np.zeros(2 ** n_spins, dtype=complex)
    psi[0] = 1.0
    return psi


def csr_algorithmic_bytes(nnz, rows, cols, ncols_x=1):
    """bytes one CSR RHS evaluation must move (single column)."""
    return nnz * 20 + (rows + 1) * 4 + 16 * cols * ncols_x + 16 * rows * ncols_x
