"""Pin the oracle (oracle/rk_oracle.py + oracle/spmv_oracle.c) against outputs of the
unmodified reference stored in tests/golden/ (generator: tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest

from _golden import load, orc_op, orc_rhs
from oracle.rk_oracle import OrcEvo, mcsolve_oracle, mesolve_oracle


@pytest.mark.parametrize("kind", ["csr", "dia", "dense"])
def test_matvec_matches_reference(kind):
    g = load("matmul")
    op = orc_op(g, kind)
    for si, sc in enumerate(g["scales"]):
        r = op.matvec(g["x"], sc)
        # reference's own bar for these products: atol 1e-10 / rtol 1e-7
        np.testing.assert_allclose(r, g["%s_mul_s%d" % (kind, si)], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name,method,tol", [
    ("c1_jc", "vern7", 1e-11), ("c1_jc", "vern9", 1e-10), ("c1_jc", "tsit5", 1e-10),
    ("c2_tfim4", "vern7", 1e-12), ("c2_tfim4", "vern9", 1e-10),
    ("c4_driven", "vern7", 1e-11),
    ("c5_kerr_0", "vern7", 1e-11),
])
def test_mesolve_matches_reference(name, method, tol):
    g = load(name)
    rhs = orc_rhs(g)
    n = int(g["n"])
    eops = [g["eop%d" % i] for i in range(int(g["n_eops"]))]
    ef = [(lambda E: (lambda t, y: np.trace(E @ y.reshape(n, n, order="F"))))(E)
          for E in eops]
    r = mesolve_oracle(rhs, g["y0"], g["tlist"], method, ef)
    assert np.abs(r["states"] - g["states_" + method]).max() < tol
    assert np.abs(r["expect"] - g["expect_" + method]).max() < tol


@pytest.mark.parametrize("k", [1, 2, 3])
def test_mesolve_kerr_stability_limited(k):
    """Stability-limited (many rejected steps) Kerr cases: round-off in the stiff
    components is amplified to the integration tolerance, so only north_star's
    tolerance (atol 1e-8 / rtol 1e-6 on expectation values) is meaningful."""
    g = load("c5_kerr_%d" % k)
    rhs = orc_rhs(g)
    n = int(g["n"])
    E = g["eop0"]
    r = mesolve_oracle(rhs, g["y0"], g["tlist"], "vern7",
                       [lambda t, y: np.trace(E @ y.reshape(n, n, order="F"))])
    np.testing.assert_allclose(r["expect"], g["expect_vern7"], rtol=1e-6, atol=1e-8)
    assert np.abs(r["states"] - g["states_vern7"]).max() < 1e-6


@pytest.mark.parametrize("name,method", [("c3_tfim6_mc", "vern7"),
                                         ("c3_tfim4_mc_strong", "vern9"),
                                         ("c3_tfim4_mc_tsit5", "tsit5")])
def test_mcsolve_matches_reference(name, method):
    g = load(name)
    rhs = orc_rhs(g)
    nc = int(g["n_cops"])
    cops = [OrcEvo([(orc_op(g, "cop%d" % i), 1.0)]) for i in range(nc)]
    nops = [OrcEvo([(orc_op(g, "nop%d" % i), 1.0)]) for i in range(nc)]
    eops = [orc_op(g, "eop%d" % i) for i in range(int(g["n_eops"]))]
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    total_jumps = 0
    for j in range(int(g["ntraj"])):
        r = mcsolve_oracle(rhs, cops, nops, g["psi0"], g["tlist"], g["draws"][j], eops,
                           method)
        ct = g["col_times"][cc[j]:cc[j + 1]]
        cw = g["col_which"][cc[j]:cc[j + 1]]
        # jump counts and collapse indices: bit-exact
        assert [w for _, w in r["collapses"]] == list(cw)
        np.testing.assert_allclose([t for t, _ in r["collapses"]], ct, rtol=0, atol=1e-10)
        np.testing.assert_allclose(r["expect"], g["runs_expect"][:, j, :], rtol=1e-6,
                                   atol=1e-9)
        np.testing.assert_allclose(r["states"][-1], g["final_states"][j], atol=1e-9)
        total_jumps += len(cw)
    assert total_jumps > 0
