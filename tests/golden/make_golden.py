#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(qutip 5.4.0.dev built into oracle/_ref by oracle/build_ref.py) in this container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box's tests as a hard requirement, so its outputs
on reduced-size versions of BASELINE.json's five configs are committed as small .npz
files; tests/test_oracle.py pins oracle/ against them and the GPU tests compare the CUDA
path with them too.  Operators are stored exactly as the reference's constructors built
them (CSR / Dia arrays), so layouts match SURVEY.md section 8d.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
warnings.filterwarnings("ignore")

import qutip  # noqa: E402
from qutip import (basis, destroy, liouvillian, mcsolve, mesolve, qeye, sigmam,  # noqa
                   sigmax, sigmaz, tensor, QobjEvo, operator_to_vector, num)
from qutip.core import data as _data  # noqa: E402


def pack_op(prefix, d, out):
    """Store a data-layer object under keys prefix_*."""
    if isinstance(d, _data.CSR):
        s = d.as_scipy()
        out[prefix + "_kind"] = "csr"
        out[prefix + "_data"] = s.data.copy()
        out[prefix + "_col"] = s.indices.astype(np.int32)
        out[prefix + "_rowptr"] = s.indptr.astype(np.int32)
    elif isinstance(d, _data.Dia):
        s = d.as_scipy()
        out[prefix + "_kind"] = "dia"
        out[prefix + "_data"] = s.data.copy()
        out[prefix + "_offsets"] = s.offsets.astype(np.int32)
    else:
        out[prefix + "_kind"] = "dense"
        out[prefix + "_arr"] = d.to_array()
    out[prefix + "_shape"] = np.array(d.shape)


def tfim(n, gamma=0.1):
    sx, sz, sm = [], [], []
    for i in range(n):
        ops = [qeye(2)] * n
        ops[i] = sigmax(); sx.append(tensor(ops))
        ops[i] = sigmaz(); sz.append(tensor(ops))
        ops[i] = sigmam(); sm.append(tensor(ops))
    H = 0
    for i in range(n - 1):
        H = H - sz[i] * sz[i + 1]
    for i in range(n):
        H = H - sx[i]
    c_ops = [np.sqrt(gamma) * s for s in sm]
    return H, c_ops, sz


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


OPT = {"progress_bar": False, "store_states": True}


def golden_mesolve(name, H, psi0, tlist, c_ops, e_ops, methods=("vern7", "vern9"), args=None,
                   force=None):
    out = {}
    solver = qutip.MESolver(H if isinstance(H, QobjEvo) else QobjEvo(H, args=args) if
                            isinstance(H, list) else H, c_ops,
                            options=dict(OPT, method=methods[0]))
    rhs = solver.rhs
    if force is not None:
        rhs = rhs.to(force)
    lst = rhs.to_list()
    nel = 0
    for el in lst:
        if isinstance(el, qutip.Qobj):
            pack_op("el%d" % nel, el.data, out)
            out["el%d_coeff" % nel] = ""
        else:
            pack_op("el%d" % nel, el[0].data, out)
            # StrCoefficient state: (values..., args, code, names...); conj wraps it
            out["el%d_coeff" % nel] = describe_coeff(el[1])
        nel += 1
    out["n_elements"] = nel
    out["tlist"] = np.asarray(tlist, dtype=float)
    rho0 = qutip.ket2dm(psi0) if psi0.isket else psi0
    out["y0"] = rho0.full().ravel("F")
    out["n"] = rho0.shape[0]
    for i, e in enumerate(e_ops):
        out["eop%d" % i] = e.full()
    out["n_eops"] = len(e_ops)
    for m in methods:
        r = mesolve(H, psi0, tlist, c_ops, e_ops=e_ops, args=args,
                    options=dict(OPT, method=m))
        out["states_" + m] = np.array([s.full().ravel("F") for s in r.states])
        out["expect_" + m] = np.array(r.expect)
    save(name, **out)


def describe_coeff(c):
    """Text form of a Coefficient for the fixtures: 'str:<expr>|k=v,...' or
    'conj(<...>)'."""
    cls = type(c).__name__
    if cls == "ConjCoefficient":
        return "conj(" + describe_coeff(c.__reduce__()[2][1]) + ")"
    if cls.startswith("StrCoefficient"):
        st = c.__reduce__()
        # generated class, auto_pickle -> (rebuild, (cls, checksum, state))
        state = st[2] if len(st) > 2 and st[2] is not None else st[1][2]
        # (named values..., extracted numeric constants..., args, code, names...): the code
        # string is the first str; the names follow it and own the leading values
        i_code = next(i for i, x in enumerate(state) if isinstance(x, str))
        code, names = state[i_code], state[i_code + 1:]
        vals = state[:len(names)]
        return "str:" + code + "|" + ",".join("%s=%r" % (k, complex(v))
                                              for k, v in zip(names, vals))
    raise NotImplementedError(cls)


def golden_mcsolve(name, H, c_ops, psi0, tlist, e_ops, ntraj, seed, method="vern7", super_h=False):
    """super_h: the Hamiltonian is handed over as a super-operator (liouvillian(H)); the
    trajectories then evolve the column-stacked density matrix, the jump logic runs on
    tr(rho) and the collapse operators are spre(c) spost(c^dag) (solver/mcsolve.py:481-490)."""
    out = {}
    if super_h:
        H = liouvillian(H)
    solver = qutip.MCSolver(H, c_ops, options={"progress_bar": False, "method": method,
                                               "keep_runs_results": True,
                                               "store_final_state": True})
    # NB: the reference keeps -iH and every -0.5*n_op as separate constant elements
    els = solver.rhs().to_list()
    for i, el in enumerate(els):
        pack_op("el%d" % i, el.data, out)
        out["el%d_coeff" % i] = ""
    out["n_elements"] = len(els)
    for i, (c, n) in enumerate(zip(solver._c_ops, solver._n_ops)):
        pack_op("cop%d" % i, c.to_list()[0].data, out)
        pack_op("nop%d" % i, n.to_list()[0].data, out)
    out["n_cops"] = len(c_ops)
    for i, e in enumerate(e_ops):
        pack_op("eop%d" % i, e.data, out)
        out["eop%d_full" % i] = e.full()
    out["n_eops"] = len(e_ops)
    out["tlist"] = np.asarray(tlist, dtype=float)
    out["psi0"] = qutip.ket2dm(psi0).full().ravel("F") if super_h else psi0.full().ravel()
    out["super_n"] = psi0.shape[0] if super_h else 0
    ss = np.random.SeedSequence(seed)
    r = solver.run(psi0, tlist, ntraj=ntraj, e_ops=e_ops, seeds=ss)
    out["seed"] = seed
    out["ntraj"] = ntraj
    out["runs_expect"] = np.array(r.runs_expect)        # [n_e][ntraj][nt]
    out["avg_expect"] = np.array(r.average_expect)
    ncol = np.array([len(c) for c in r.col_times])
    out["col_count"] = ncol
    out["col_times"] = np.concatenate([np.asarray(c, dtype=float) for c in r.col_times]
                                      + [np.zeros(0)])
    out["col_which"] = np.concatenate([np.asarray(c, dtype=np.int64) for c in r.col_which]
                                      + [np.zeros(0, dtype=np.int64)])
    out["final_states"] = np.array([s.full().ravel("F") if super_h else s.full().ravel()
                                    for s in r.runs_final_states])
    # thresholds exactly as the reference draws them
    kids = np.random.SeedSequence(seed).spawn(ntraj)
    out["draws"] = np.stack([np.random.default_rng(k).random(64) for k in kids])
    save(name, **out)


def golden_matmul():
    rng = np.random.default_rng(0)
    out = {}
    n = 96
    dense = rng.random((n, n)) + 1j * rng.random((n, n))
    mask = rng.random((n, n)) < 0.1
    A = dense * mask
    csr = _data.to(_data.CSR, _data.Dense(A))
    # unsorted column indices, as the reference's tests cover
    s = csr.as_scipy()
    for r in range(n):
        lo, hi = s.indptr[r], s.indptr[r + 1]
        perm = rng.permutation(hi - lo)
        s.indices[lo:hi] = s.indices[lo:hi][perm]
        s.data[lo:hi] = s.data[lo:hi][perm]
    pack_op("csr", csr, out)
    B = np.zeros((n, n), dtype=complex)
    for off in (-40, -3, -1, 0, 2, 7, 50):
        B += np.diag(rng.random(n - abs(off)) + 1j * rng.random(n - abs(off)), off)
    dia = _data.to(_data.Dia, _data.Dense(B))
    pack_op("dia", dia, out)
    pack_op("dense", _data.Dense(dense), out)
    x = rng.random(n) + 1j * rng.random(n)
    out["x"] = x
    xd = _data.Dense(x.reshape(-1, 1))
    for nm, op in (("csr", csr), ("dia", dia), ("dense", _data.Dense(dense))):
        for si, sc in enumerate((1.0, 0.5, 0.5j, -0.3 + 0.7j)):
            pre = np.ones((n, 1), dtype=complex)
            o = _data.Dense(pre.copy())
            r = _data.matmul(op, xd, sc)
            out["%s_mul_s%d" % (nm, si)] = r.to_array().ravel()
    out["scales"] = np.array([1.0, 0.5, 0.5j, -0.3 + 0.7j])
    save("matmul", **out)


def golden_adams():
    """zvode (the reference's method='adams', solver/integrator/scipy_integrator.py:20-196;
    SciPy's compiled zvode, tied to the installed SciPy) on C1 and reduced C2 / C4: expectation
    values and final states at the default tolerances and at tight ones, with zvode's own step
    and RHS-evaluation counts.  The device-resident Adams method is a different member of the
    ODEPACK family (no step-level parity): it is pinned against these solutions."""
    import scipy
    out = {"scipy_version": scipy.__version__}
    N = 10
    a = tensor(destroy(N), qeye(2)); sm = tensor(qeye(N), destroy(2))
    H1 = 2 * np.pi * a.dag() * a + 2 * np.pi * sm.dag() * sm \
        + 2 * np.pi * 0.05 * (a.dag() * sm + a * sm.dag())
    cases = {"c1_jc": (H1, tensor(basis(N, 3), basis(2, 0)), np.linspace(0, 10, 101),
                       [np.sqrt(0.1) * a, np.sqrt(0.05) * sm], [a.dag() * a, tensor(qeye(N), sigmaz())])}
    H2, c2, sz = tfim(4)
    cases["c2_tfim4"] = (H2, basis([2] * 4, [0] * 4), np.linspace(0, 1, 11), c2, [sz[0]])
    Nc = 8
    a4 = tensor(destroy(Nc), qeye(3)); b4 = tensor(qeye(Nc), destroy(3))
    H0 = 5 * a4.dag() * a4 + 4.5 * b4.dag() * b4 - 0.15 * b4.dag() * b4.dag() * b4 * b4 \
        + 0.1 * (a4.dag() * b4 + a4 * b4.dag())
    Ht = QobjEvo([H0, [a4 + a4.dag(), "A*cos(w*t)"]], args={"A": 0.2, "w": 5.0})
    cases["c4_driven"] = (Ht, tensor(basis(Nc, 0), basis(3, 0)), np.linspace(0, 5, 51),
                          [np.sqrt(0.01) * a4, np.sqrt(0.02) * b4, np.sqrt(0.03) * b4.dag() * b4],
                          [a4.dag() * a4, b4.dag() * b4])
    for name, (H, psi0, tl, c_ops, e_ops) in cases.items():
        for tag, tol in (("default", {}), ("tight", {"atol": 1e-12, "rtol": 1e-10, "nsteps": 100000})):
            solver = qutip.MESolver(H, c_ops, options=dict(OPT, method="adams", **tol))
            r = solver.run(psi0, tl, e_ops=e_ops)
            iw = solver._integrator._ode_solver._integrator.iwork
            out["%s_%s_expect" % (name, tag)] = np.array(r.expect)
            out["%s_%s_final" % (name, tag)] = r.states[-1].full().ravel("F")
            out["%s_%s_nst_nfe" % (name, tag)] = np.array([int(iw[10]), int(iw[11])])
    save("adams_zvode", **out)



def golden_nm_mcsolve(name="nm_two_level", ntraj=32, seed=3):
    """nm_mcsolve (solver/nm_mcsolve.py) on a two-level system with one always-negative rate
    (smooth shifted rates: well conditioned, see tests/test_gpu_plugin.py) and one constant rate.
    Stored: the effective system the reference integrates (elements of H_eff, rate-shifted
    collapse operators c_k(t) C_k and n_k(t) C_k^dagger C_k) with every coefficient as the
    byte-code `qutip_b200.plugin.coefficient_to_program` compiles it to (tests/test_plugin_cpu.py
    checks those programs against the reference coefficients), and the reference's per-trajectory
    collapse records, un-weighted expectation values and influence martingales."""
    sys.path.insert(0, ROOT)
    import qutip_b200.plugin as plugin
    from qutip import NonMarkovianMCSolver, coefficient, sigmap
    H = 0.5 * sigmaz() + 0.2 * sigmax()
    ops_and_rates = [(sigmam(), coefficient("-0.08 + 0.05*sin(2*t)")), (sigmap(), 0.15)]
    psi0 = (basis(2, 0) + 0.5 * basis(2, 1)).unit()
    tlist = np.linspace(0, 4, 17)
    e_ops = [sigmaz(), sigmax()]
    solver = NonMarkovianMCSolver(H, ops_and_rates, options={"progress_bar": False, "method": "vern7",
                                                              "keep_runs_results": True,
                                                              "store_final_state": True})
    out = {}

    def pack_terms(prefix, qevo):
        """constant elements merged (as the device binding does), each td element with its program"""
        n = 0
        const = None
        for el in qevo.to_list():
            if isinstance(el, qutip.Qobj):
                const = el if const is None else const + el
        for el in qevo.to_list():
            if isinstance(el, (list, tuple)):
                pack_op("%s%d" % (prefix, n), el[0].to("CSR").data, out)
                out["%s%d_prog" % (prefix, n)] = np.array(plugin.coefficient_to_program(el[1]).instrs, dtype=float)
                n += 1
        if const is not None:
            pack_op("%s%d" % (prefix, n), const.to("CSR").data, out)
            out["%s%d_prog" % (prefix, n)] = np.zeros((0, 4))
            n += 1
        return n

    out["n_elements"] = pack_terms("el", solver.rhs.rhs)
    for i, (c, n) in enumerate(zip(solver.rhs.c_ops, solver.rhs.n_ops)):
        assert pack_terms("cop%d_" % i, c) == 1 and pack_terms("nop%d_" % i, n) == 1
    out["n_cops"] = len(solver.rhs.c_ops)
    for i, e in enumerate(e_ops):
        pack_op("eop%d" % i, e.to("CSR").data, out)
    out["n_eops"] = len(e_ops)
    out["tlist"] = tlist
    out["psi0"] = psi0.full().ravel()
    r = solver.run(psi0, tlist, ntraj=ntraj, e_ops=e_ops, seeds=np.random.SeedSequence(seed))
    out["seed"], out["ntraj"] = seed, ntraj
    # NonMarkovianMCSolver's own defaults differ from MCSolver's (nm_mcsolve.py:352-372)
    out["norm_steps"], out["norm_min_step"] = solver.options["norm_steps"], solver.options["norm_min_step"]
    out["norm_tol"], out["norm_t_tol"] = solver.options["norm_tol"], solver.options["norm_t_tol"]
    out["raw_expect"] = np.array([[np.asarray(tr.e_data[k]) for k in range(len(e_ops))]
                                  for tr in r.trajectories])          # [ntraj][n_e][nt], un-weighted
    out["runs_trace"] = np.array(r.runs_trace)
    out["avg_expect"] = np.array(r.average_expect)
    out["col_count"] = np.array([len(c) for c in r.col_times])
    out["col_times"] = np.concatenate([np.asarray(c, dtype=float) for c in r.col_times] + [np.zeros(0)])
    out["col_which"] = np.concatenate([np.asarray(c, dtype=np.int64) for c in r.col_which]
                                      + [np.zeros(0, dtype=np.int64)])
    out["final_states"] = np.array([s.full().ravel() if s.isket else s.full().ravel("F")
                                    for s in r.runs_final_states])
    kids = np.random.SeedSequence(seed).spawn(ntraj)
    out["draws"] = np.stack([np.random.default_rng(k).random(64) for k in kids])
    save(name, **out)


def golden_mcsolve_improved(name="c3_tfim4_mc_improved", ntraj=16, seed=9):
    """mcsolve with options["improved_sampling"] (solver/mcsolve.py:716-747): the no-jump trajectory
    is integrated first, its survival probability p becomes the floor of every trajectory's FIRST
    threshold (mcsolve.py:276-279) and the trajectories carry the weight 1 - p (:565)."""
    H, c_ops, sz = tfim(4, gamma=0.25)
    psi0 = basis([2] * 4, [0] * 4)
    tlist = np.linspace(0, 2, 11)
    e_ops = [sz[0], sz[2]]
    solver = qutip.MCSolver(H, c_ops, options={"progress_bar": False, "method": "vern7",
                                               "keep_runs_results": True, "store_final_state": True,
                                               "improved_sampling": True})
    out = {}
    els = solver.rhs().to_list()
    for i, el in enumerate(els):
        pack_op("el%d" % i, el.data, out)
        out["el%d_coeff" % i] = ""
    out["n_elements"] = len(els)
    for i, (c, n) in enumerate(zip(solver._c_ops, solver._n_ops)):
        pack_op("cop%d" % i, c.to_list()[0].data, out)
        pack_op("nop%d" % i, n.to_list()[0].data, out)
    out["n_cops"] = len(c_ops)
    for i, e in enumerate(e_ops):
        pack_op("eop%d" % i, e.data, out)
    out["n_eops"] = len(e_ops)
    out["tlist"] = tlist
    out["psi0"] = psi0.full().ravel()
    r = solver.run(psi0, tlist, ntraj=ntraj, e_ops=e_ops, seeds=np.random.SeedSequence(seed))
    det = r.deterministic_trajectories[0]
    out["no_jump_expect"] = np.array([np.asarray(det.e_data[k]) for k in range(len(e_ops))])
    out["no_jump_prob"] = float(r.deterministic_weights[0]) if hasattr(r, "deterministic_weights") \
        else float(solver._no_jump_simulation(solver._prepare_state(psi0), tlist, e_ops)[1])
    out["seed"], out["ntraj"] = seed, ntraj
    out["raw_expect"] = np.array([[np.asarray(tr.e_data[k]) for k in range(len(e_ops))] for tr in r.trajectories])
    out["avg_expect"] = np.array(r.average_expect)
    out["col_count"] = np.array([len(c) for c in r.col_times])
    out["col_times"] = np.concatenate([np.asarray(c, dtype=float) for c in r.col_times] + [np.zeros(0)])
    out["col_which"] = np.concatenate([np.asarray(c, dtype=np.int64) for c in r.col_which]
                                      + [np.zeros(0, dtype=np.int64)])
    out["final_states"] = np.array([s.full().ravel() for s in r.runs_final_states])
    kids = np.random.SeedSequence(seed).spawn(ntraj)
    out["draws"] = np.stack([np.random.default_rng(k).random(64) for k in kids])
    save(name, **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "improved":
        return golden_mcsolve_improved()
    if len(sys.argv) > 1 and sys.argv[1] == "nm":
        return golden_nm_mcsolve()
    if len(sys.argv) > 1 and sys.argv[1] == "adams":
        return golden_adams()
    if len(sys.argv) > 1 and sys.argv[1] == "super":
        H, c_ops, sz = tfim(3, gamma=0.8)
        return golden_mcsolve("c3_tfim3_mc_super", H, c_ops, basis([2] * 3, [0] * 3), np.linspace(0, 3, 13),
                              [sz[0], sz[1]], 16, 21, super_h=True)
    golden_matmul()

    # C1: damped Jaynes-Cummings, cavity N=10 (x) qubit  
    N = 10
    a = tensor(destroy(N), qeye(2)); sm = tensor(qeye(N), destroy(2))
    H = 2 * np.pi * a.dag() * a + 2 * np.pi * sm.dag() * sm \
        + 2 * np.pi * 0.05 * (a.dag() * sm + a * sm.dag())
    c_ops = [np.sqrt(0.1) * a, np.sqrt(0.05) * sm]
    psi0 = tensor(basis(N, 3), basis(2, 0))
    golden_mesolve("c1_jc", H, psi0, np.linspace(0, 10, 101), c_ops,
                   [a.dag() * a, tensor(qeye(N), sigmaz())], methods=("vern7", "vern9", "tsit5"))

    # C2 reduced: dissipative TFIM, 4 spins (L 256^2), CSR from the reference ctor
    H, c_ops, sz = tfim(4)
    psi0 = basis([2] * 4, [0] * 4)
    golden_mesolve("c2_tfim4", H, psi0, np.linspace(0, 1, 11), c_ops, [sz[0]])

    # C3 reduced: mcsolve TFIM 6 spins (dim 64), 24 trajectories, SeedSequence(7)
    H, c_ops, sz = tfim(6)
    psi0 = basis([2] * 6, [0] * 6)
    golden_mcsolve("c3_tfim6_mc", H, c_ops, psi0, np.linspace(0, 2, 21), [sz[0]], 24, 7)
    # stronger damping -> many jumps per trajectory
    H, c_ops, sz = tfim(4, gamma=1.5)
    psi0 = basis([2] * 4, [0] * 4)
    golden_mcsolve("c3_tfim4_mc_strong", H, c_ops, psi0, np.linspace(0, 3, 16),
                   [sz[0], sz[1]], 16, 11, method="vern9")
    golden_mcsolve("c3_tfim4_mc_tsit5", H, c_ops, psi0, np.linspace(0, 3, 16),
                   [sz[0], sz[1]], 12, 13, method="tsit5")

    # C4 reduced: cos-driven cavity 8 (x) transmon 3, string coefficient
    Nc = 8
    a = tensor(destroy(Nc), qeye(3)); b = tensor(qeye(Nc), destroy(3))
    H0 = 5 * a.dag() * a + 4.5 * b.dag() * b - 0.15 * b.dag() * b.dag() * b * b \
        + 0.1 * (a.dag() * b + a * b.dag())
    H1 = a + a.dag()
    args = {"A": 0.2, "w": 5.0}
    Ht = QobjEvo([H0, [H1, "A*cos(w*t)"]], args=args)
    c_ops = [np.sqrt(0.01) * a, np.sqrt(0.02) * b, np.sqrt(0.03) * b.dag() * b]
    psi0 = tensor(basis(Nc, 0), basis(3, 0))
    golden_mesolve("c4_driven", Ht, psi0, np.linspace(0, 5, 51), c_ops,
                   [a.dag() * a, b.dag() * b], methods=("vern7",))

    # C5 reduced: driven Kerr N=12, 4 parameter points
    Nk = 12
    a = destroy(Nk)
    for i, (U, D, F) in enumerate([(0.5, 0.2, 0.4), (1.0, -0.5, 0.8), (0.1, 1.0, 0.3),
                                   (2.0, 0.0, 1.0)]):
        H = 0.5 * U * a.dag() * a.dag() * a * a - D * a.dag() * a + F * (a + a.dag())
        golden_mesolve("c5_kerr_%d" % i, H, basis(Nk, 0), np.linspace(0, 10, 21), [a],
                       [a.dag() * a], methods=("vern7",))


if __name__ == "__main__":
    main()
