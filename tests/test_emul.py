"""No-GPU check of the engine's control logic and operator format.

tests/emul/emul.cpp compiles the SAME headers the CUDA kernels use (qb_control.h step
controller, qb_coeff.h byte-code, qb_diam.h format conversion) with g++ and executes each
vector pass with a serial loop.  Compared here with the golden fixtures of the reference:
if these pass, the flattened state machine reproduces the reference's step sequence,
RHS-evaluation counts and collapse records."""
import numpy as np
import pytest

from _emul import FMT_CSR, FMT_DIAM, FMT_RSELL, FMT_SELL, EmulSystem, default_options
from _golden import coeff_spec, load, op_arrays, orc_op, orc_rhs
from _systems import functional_of, merged_constant_rhs
from qutip_b200 import coeffs


def _sp_arrays(m):
    import scipy.sparse as sp
    m = sp.csr_matrix(m)
    m.sort_indices()
    return ("csr", m.shape, dict(data=m.data, col=m.indices, rowptr=m.indptr))


@pytest.mark.parametrize("kind", ["csr", "dia"])
@pytest.mark.parametrize("fmt", [FMT_DIAM, FMT_SELL, FMT_RSELL])
def test_diam_format_matvec(kind, fmt):
    g = load("matmul")
    n = len(g["x"])
    s = EmulSystem(n, 0, fmt)
    s.add_element(*op_arrays(g, kind))
    np.testing.assert_allclose(s.matvec(0, g["x"]), g["%s_mul_s0" % kind], rtol=1e-13,
                               atol=1e-13)


def test_diam_duplicate_entries():
    # duplicate (row, col) pairs are legal in CSR; they must be summed
    data = np.array([1.0, 2.0, 3.0, 4.0], dtype=complex)
    col = np.array([1, 1, 0, 2], dtype=np.int32)
    rowptr = np.array([0, 2, 3, 4], dtype=np.int32)
    s = EmulSystem(3, 0, FMT_DIAM)
    s.add_element("csr", (3, 3), dict(data=data, col=col, rowptr=rowptr))
    x = np.array([1.0, 10.0, 100.0], dtype=complex)
    np.testing.assert_allclose(s.matvec(0, x), [30.0, 3.0, 400.0])
    for fmt in (FMT_SELL, FMT_RSELL):
        s = EmulSystem(3, 0, fmt)
        s.add_element("csr", (3, 3), dict(data=data, col=col, rowptr=rowptr))
        np.testing.assert_allclose(s.matvec(0, x), [30.0, 3.0, 400.0])


def test_rsell_spin_chain_is_rule_compressed():
    # H_eff of the dissipative TFIM (C3 model, 8 spins): every slot of every slice follows the
    # xor rule (sigma_x flips one bit of the row index) and only the diagonal needs a value
    # block -- no column index is stored at all (qb_types.h, RSELL)
    from qutip_b200 import models
    H, c_ops, _ = models.tfim(8)
    A = models.heff(H, c_ops)
    n = A.shape[0]
    s = EmulSystem(n, 0, FMT_RSELL)
    s.add_element(*_sp_arrays(A))
    st = s.rsell_stats(0)
    assert st["slots"] == (n // 32) * 9 and st["xor_slots"] == st["slots"]
    assert st["col_blocks"] == 0 and st["val_blocks"] == n // 32
    assert st["unique_descriptors"] == 9        # every slice shares ONE descriptor list
    assert st["bytes"] < 0.25 * (A.nnz * 20)
    x = np.random.default_rng(1).random(n) + 1j * np.random.default_rng(2).random(n)
    np.testing.assert_allclose(s.matvec(0, x), A @ x, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("n", [32, 33, 63, 100, 257, 400])
def test_rsell_banded_and_random(n):
    import scipy.sparse as sp
    rng = np.random.default_rng(n)
    a = sp.diags([np.sqrt(np.arange(1, n))], [1]).tocsr()            # destroy(n): row + 1 rule
    band = (a + a.T + sp.diags([np.arange(n) * (1 - 0.5j)], [0])).tocsr()
    rnd = sp.random(n, n, density=0.05, random_state=rng, dtype=float).tocsr() * (1 + 2j)
    for m in (band, rnd, sp.csr_matrix((n, n), dtype=complex), (band + rnd).tocsr()):
        m = sp.csr_matrix(m, dtype=complex)
        s = EmulSystem(n, 0, FMT_RSELL)
        s.add_element(*_sp_arrays(m))
        x = rng.random(n) + 1j * rng.random(n)
        np.testing.assert_allclose(s.matvec(0, x), m @ x, rtol=1e-13, atol=1e-13)
    s = EmulSystem(n, 0, FMT_RSELL)
    s.add_element(*_sp_arrays(band))
    st = s.rsell_stats(0)
    assert st["col_blocks"] <= 2 * 3          # only the first / last slice can leave the band rule


@pytest.mark.parametrize("name,method", [("c1_jc", "vern7"), ("c1_jc", "vern9"), ("c1_jc", "tsit5"),
                                         ("c2_tfim4", "vern7"), ("c2_tfim4", "vern9"),
                                         ("c4_driven", "vern7"), ("c5_kerr_0", "vern7")])
@pytest.mark.parametrize("fmt", [FMT_DIAM, FMT_CSR, FMT_SELL, FMT_RSELL])
def test_mesolve_state_machine(name, method, fmt):
    g = load(name)
    s = EmulSystem(len(g["y0"]), 0, fmt)
    for i in range(int(g["n_elements"])):
        spec = coeff_spec(g["el%d_coeff" % i])
        prog = coeffs.compile_expr(spec[0], spec[1]) if spec is not None else None
        s.add_element(*op_arrays(g, "el%d" % i), prog=prog)
    for i in range(int(g["n_eops"])):
        s.add_eop(*_sp_arrays(functional_of(g["eop%d" % i])))
    s.set_functional(1)
    r = s.run(0, {"vern7": 0, "vern9": 1, "tsit5": 2}[method], g["y0"], g["tlist"],
              opt=default_options(store_states=1))
    assert r["status"][0] == 1
    assert np.abs(r["states"][0] - g["states_" + method]).max() < 1e-10
    assert np.abs(r["expect"][0] - g["expect_" + method]).max() < 1e-10


def test_mesolve_counts_match_oracle():
    from oracle.rk_oracle import mesolve_oracle
    g = load("c1_jc")
    rhs = orc_rhs(g)
    o = mesolve_oracle(rhs, g["y0"], g["tlist"], "vern7")
    s = EmulSystem(len(g["y0"]), 0, FMT_DIAM)
    for i in range(int(g["n_elements"])):
        s.add_element(*op_arrays(g, "el%d" % i))
    r = s.run(0, 0, g["y0"], g["tlist"])
    assert r["stats"][0][0] == rhs.nevals
    assert (r["stats"][0][1], r["stats"][0][2]) == (o["rk"].n_accept, o["rk"].n_reject)


@pytest.mark.parametrize("name,method,nslots", [("c3_tfim6_mc", "vern7", 24),
                                               ("c3_tfim6_mc", "vern7", 5),
                                               ("c3_tfim4_mc_strong", "vern9", 7),
                                               ("c3_tfim4_mc_tsit5", "tsit5", 5)])
def test_mcsolve_state_machine(name, method, nslots):
    g = load(name)
    s = EmulSystem(len(g["psi0"]), 0, FMT_DIAM)
    s.add_element(*_sp_arrays(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
    for i in range(int(g["n_eops"])):
        s.add_eop(*op_arrays(g, "eop%d" % i))
    ntraj = int(g["ntraj"])
    r = s.run(1, {"vern7": 0, "vern9": 1, "tsit5": 2}[method], g["psi0"], g["tlist"], ntraj=ntraj,
              nslots=nslots, draws=g["draws"], opt=default_options(store_states=1))
    assert (r["status"] == 1).all()
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    assert np.array_equal(r["ncol"], g["col_count"])
    for j in range(ntraj):
        n = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r["col_t"][j, :n], g["col_times"][cc[j]:cc[j + 1]],
                                   rtol=0, atol=1e-10)
    assert np.abs(np.transpose(r["expect"], (1, 0, 2)) - g["runs_expect"]).max() < 1e-9
    assert np.abs(r["states"][:, -1, :] - g["final_states"]).max() < 1e-9


def test_rng_exhaustion_status():
    g = load("c3_tfim4_mc_strong")
    s = EmulSystem(len(g["psi0"]), 0, FMT_DIAM)
    s.add_element(*_sp_arrays(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
    r = s.run(1, 1, g["psi0"], g["tlist"], ntraj=int(g["ntraj"]), draws=g["draws"][:, :3])
    assert (r["status"] == -12).any()


def test_coefficient_bytecode():
    import ctypes as C
    from _emul import lib
    L = lib()
    cases = [("A*cos(w*t)", {"A": 0.2, "w": 5.0}), ("exp(-t/tau)*sin(2*pi*f*t)**2", {"tau": 3.0, "f": 0.7}),
             ("conj(exp(1j*w*t))", {"w": 2.0}), ("sqrt(abs(t-1.5))+real((1+2j)*t)", {}),
             ("-(t**2)/(1+t)", {}), ("tanh(t)+cosh(0.1*t)-sinh(0.2*t)", {})]
    for expr, args in cases:
        p = coeffs.compile_expr(expr, args)
        for t in (0.0, 0.3, 1.7, 4.2):
            out = (C.c_double * 2)()
            rc = L.emul_eval_prog(p.as_ctypes(), len(p), C.c_double(t), None, out)
            assert rc == 0
            ref = coeffs.evaluate(p, t)
            assert abs(complex(out[0], out[1]) - ref) < 1e-13 * max(1, abs(ref))
    # nm_mcsolve's rate shift 2*|min(0, r1, r2)| and sqrt(real(r + shift)) (solver/cy/nm_mcsolve.pyx)
    r1, r2 = coeffs.compile_expr("0.25*sin(2*t)+0.05"), coeffs.compile_expr("0.15-0.1*t")
    shift = coeffs.rate_shift([r1, r2])
    for p, fn in ((shift, lambda t: 2 * abs(min(0.0, 0.25 * np.sin(2 * t) + 0.05, 0.15 - 0.1 * t))),
                  ((r1 + shift).sqrt_real(),
                   lambda t: np.sqrt(0.25 * np.sin(2 * t) + 0.05
                                     + 2 * abs(min(0.0, 0.25 * np.sin(2 * t) + 0.05, 0.15 - 0.1 * t))))):
        for t in (0.0, 0.9, 2.0, 2.6, 4.2):
            out = (C.c_double * 2)()
            assert L.emul_eval_prog(p.as_ctypes(), len(p), C.c_double(t), None, out) == 0
            assert abs(complex(out[0], out[1]) - fn(t)) < 1e-13
            assert abs(coeffs.evaluate(p, t) - fn(t)) < 1e-13
    with pytest.raises(TypeError):
        coeffs.compile_expr("__import__('os').system('x')")
    with pytest.raises(TypeError):
        coeffs.compile_expr("foo(t)")


@pytest.mark.parametrize("name,method,nslots", [("c3_tfim6_mc", "vern7", 24),
                                               ("c3_tfim4_mc_strong", "vern9", 7),
                                               ("c3_tfim4_mc_tsit5", "tsit5", 5)])
def test_mcsolve_state_machine_tile_mode(name, method, nslots):
    """Fixed slot labels (explicit y_prev <- y_front copies, chunked expectation passes) as
    used by the trajectory-interleaved tile engine: same results as the relabelling mode."""
    from _emul import lib
    g = load(name)
    lib().emul_set_tile_mode(1, 3)
    try:
        s = EmulSystem(len(g["psi0"]), 0, FMT_CSR)
        s.add_element(*_sp_arrays(merged_constant_rhs(g)))
        for i in range(int(g["n_cops"])):
            s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
        for i in range(int(g["n_eops"])):
            s.add_eop(*op_arrays(g, "eop%d" % i))
        ntraj = int(g["ntraj"])
        r = s.run(1, {"vern7": 0, "vern9": 1, "tsit5": 2}[method], g["psi0"], g["tlist"], ntraj=ntraj,
                  nslots=nslots, draws=g["draws"], opt=default_options(store_states=1))
    finally:
        lib().emul_set_tile_mode(0, 0)
    assert (r["status"] == 1).all()
    assert np.array_equal(r["ncol"], g["col_count"])
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        n = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :n], g["col_which"][cc[j]:cc[j + 1]])
    assert np.abs(np.transpose(r["expect"], (1, 0, 2)) - g["runs_expect"]).max() < 1e-9
    assert np.abs(r["states"][:, -1, :] - g["final_states"]).max() < 1e-9


def test_mcsolve_super_operator_hamiltonian_state_machine():
    """mcsolve with a super-operator H (solver/mcsolve.py:481-490): trajectories of the
    column-stacked rho, thresholds compared with tr(rho) (mcsolve.py:311-319), probabilities
    tr(n_k rho), renormalisation by the trace -- against the reference fixture."""
    g = load("c3_tfim3_mc_super")
    n = int(g["super_n"])
    s = EmulSystem(len(g["psi0"]), 0, FMT_CSR)
    s.add_element(*_sp_arrays(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
    for i in range(int(g["n_eops"])):
        s.add_eop(*_sp_arrays(functional_of(g["eop%d_full" % i])))
    s.set_functional(1)
    s.set_mc_trace(n)
    ntraj = int(g["ntraj"])
    r = s.run(1, 0, g["psi0"], g["tlist"], ntraj=ntraj, nslots=5, draws=g["draws"],
              opt=default_options(store_states=1))
    assert (r["status"] == 1).all()
    assert np.array_equal(r["ncol"], g["col_count"]) and g["col_count"].sum() > 10
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        k = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :k], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r["col_t"][j, :k], g["col_times"][cc[j]:cc[j + 1]], rtol=0, atol=1e-9)
    assert np.abs(np.transpose(r["expect"], (1, 0, 2)) - g["runs_expect"]).max() < 1e-9
    assert np.abs(r["states"][:, -1, :] - g["final_states"]).max() < 1e-9


@pytest.mark.parametrize("fmt", [FMT_DIAM, FMT_CSR])
def test_nm_mcsolve_state_machine(fmt):
    """nm_mcsolve (solver/nm_mcsolve.py): rate-shifted collapse operators with time-dependent
    coefficients c_k(t) = sqrt(rate_k + shift), n_k(t) = |c_k|^2 and the matching H_eff terms,
    as compiled device programs (MIN_RE / SQRT_RE byte-code) -- the reference's trajectories
    (fixture nm_two_level, generated by make_golden.py nm): collapse records, un-weighted
    expectation values and final states.  The influence martingale is host post-processing of
    the collapse records (plugin._b200_batch) and is checked on the GPU against the reference."""
    g = load("nm_two_level")
    N = len(g["psi0"])
    s = EmulSystem(N, 0, fmt)

    def prog(key):
        a = g[key]
        return coeffs.Program([(int(op), int(ia), float(re), float(im)) for op, ia, re, im in a]) if len(a) else None

    for i in range(int(g["n_elements"])):
        s.add_element(*op_arrays(g, "el%d" % i), prog=prog("el%d_prog" % i))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d_0" % i), op_arrays(g, "nop%d_0" % i),
                       cprog=prog("cop%d_0_prog" % i), nprog=prog("nop%d_0_prog" % i))
    for i in range(int(g["n_eops"])):
        s.add_eop(*op_arrays(g, "eop%d" % i))
    ntraj = int(g["ntraj"])
    # the collapse search runs with NonMarkovianMCSolver's own defaults (norm_steps 5, norm_min_step 0)
    opt = default_options(store_states=1, norm_steps=int(g["norm_steps"]), norm_min_step=float(g["norm_min_step"]),
                          norm_tol=float(g["norm_tol"]), norm_t_tol=float(g["norm_t_tol"]))
    r = s.run(1, 0, g["psi0"], g["tlist"], ntraj=ntraj, nslots=5, draws=g["draws"], opt=opt)
    assert (r["status"] == 1).all()
    assert np.array_equal(r["ncol"], g["col_count"]) and g["col_count"].sum() >= 8
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        n = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r["col_t"][j, :n], g["col_times"][cc[j]:cc[j + 1]], rtol=0, atol=1e-9)
    assert np.abs(r["expect"] - g["raw_expect"]).max() < 1e-8
    assert np.abs(r["states"][:, -1, :] - g["final_states"]).max() < 1e-8


def test_mcsolve_improved_sampling_state_machine():
    """options["improved_sampling"] (solver/mcsolve.py:716-747): the first threshold of every
    trajectory is floored at the no-jump probability (mcsolve.py:276-279) -- the reference's
    trajectories (fixture c3_tfim4_mc_improved) against `jump_prob_floor`; the no-jump trajectory
    itself is the `no_jump` option (threshold 0)."""
    g = load("c3_tfim4_mc_improved")
    s = EmulSystem(len(g["psi0"]), 0, FMT_DIAM)
    s.add_element(*_sp_arrays(merged_constant_rhs(g)))
    for i in range(int(g["n_cops"])):
        s.add_collapse(op_arrays(g, "cop%d" % i), op_arrays(g, "nop%d" % i))
    for i in range(int(g["n_eops"])):
        s.add_eop(*op_arrays(g, "eop%d" % i))
    ntraj = int(g["ntraj"])
    p = float(g["no_jump_prob"])
    assert 0.05 < p < 0.95
    r = s.run(1, 0, g["psi0"], g["tlist"], ntraj=ntraj, nslots=6, draws=g["draws"],
              opt=default_options(store_states=1, jump_prob_floor=p))
    assert (r["status"] == 1).all()
    assert np.array_equal(r["ncol"], g["col_count"]) and (g["col_count"] >= 1).all()
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        n = r["ncol"][j]
        assert np.array_equal(r["col_which"][j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r["col_t"][j, :n], g["col_times"][cc[j]:cc[j + 1]], rtol=0, atol=1e-9)
    assert np.abs(r["expect"] - g["raw_expect"]).max() < 1e-9
    assert np.abs(r["states"][:, -1, :] - g["final_states"]).max() < 1e-9
    # the deterministic no-jump trajectory and its survival probability
    r0 = s.run(1, 0, g["psi0"], g["tlist"], ntraj=1, draws=g["draws"][:1], opt=default_options(store_states=1, no_jump=1))
    assert r0["ncol"][0] == 0 and np.abs(r0["expect"][0] - g["no_jump_expect"]).max() < 1e-9
