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


# ---------------------------------------------------------------- bosonic configs (C1, C4, C5)
def destroy(n):
    return sp.diags(np.sqrt(np.arange(1, n, dtype=float)), 1, format="csr", dtype=complex)


def commutator_super(H):
    """-i (I kron H - H^T kron I): the Hamiltonian part of the Liouvillian."""
    n = H.shape[0]
    I = sp.identity(n, dtype=complex, format="csr")
    return sp.csr_matrix(-1j * (sp.kron(I, H) - sp.kron(H.T, I)))


def dissipator_super(c_ops):
    n = c_ops[0].shape[0]
    I = sp.identity(n, dtype=complex, format="csr")
    L = sp.csr_matrix((n * n, n * n), dtype=complex)
    for c in c_ops:
        cdc = c.conj().T @ c
        L = L + sp.kron(c.conj(), c) - 0.5 * sp.kron(I, cdc) - 0.5 * sp.kron(cdc.T, I)
    return sp.csr_matrix(L)


def kerr_sweep(N=30, grid=16):
    """C5: H = U/2 a^dag^2 a^2 - Delta a^dag a + F (a + a^dag), c = [a], (U, Delta, F) on a
    fixed grid^3 lattice.  Returns the four superoperators [L_U, L_Delta, L_F, L_diss], the
    per-system argument table [grid^3][3] and a, so that
    L(U, Delta, F) = U L_U + Delta L_Delta + F L_F + L_diss."""
    a = destroy(N)
    ad = sp.csr_matrix(a.conj().T)
    H_U = 0.5 * ad @ ad @ a @ a
    H_D = -(ad @ a)
    H_F = a + ad
    Us = np.linspace(0.1, 1.0, grid)
    Ds = np.linspace(-1.0, 1.0, grid)
    Fs = np.linspace(0.1, 1.5, grid)
    args = np.array([(u, d, f) for u in Us for d in Ds for f in Fs], dtype=complex)
    return ([commutator_super(H_U), commutator_super(H_D), commutator_super(H_F),
             dissipator_super([a])], args, a)


def driven_cavity_transmon(Nc=40):
    """C4: H0 + A cos(w t) (a + a^dag) with decay/dephasing; returns (H0, H1, c_ops, a, b)."""
    a = sp.kron(destroy(Nc), sp.identity(3, dtype=complex), format="csr")
    b = sp.kron(sp.identity(Nc, dtype=complex), destroy(3), format="csr")
    ad, bd = sp.csr_matrix(a.conj().T), sp.csr_matrix(b.conj().T)
    H0 = 5 * ad @ a + 4.5 * bd @ b - 0.15 * bd @ bd @ b @ b + 0.1 * (ad @ b + a @ bd)
    H1 = a + ad
    c_ops = [np.sqrt(0.01) * a, np.sqrt(0.02) * b, np.sqrt(0.03) * (bd @ b)]
    return sp.csr_matrix(H0), sp.csr_matrix(H1), c_ops, a, b
