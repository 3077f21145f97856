"""mesolve / mcsolve on plain arrays (scipy.sparse operators, numpy states).

These are the calls a user makes when the whole batch should run on the device: all
trajectories (or sweep members) advance concurrently in the fused engine, per-trajectory
records come back as arrays shaped like the reference's ``McResult`` fields
(``runs_expect``, ``col_times``, ``col_which``; solver/multitrajresult.py).  The QuTiP-facing
wrappers in plugin.py translate Qobj/QobjEvo into these calls.

Multi-GPU: trajectories are independent given their seed-derived thresholds
(multitraj.py:250-256), so each device takes a contiguous block of the spawned seed list and
only the expectation sums are reduced -- ONE ncclAllReduce issued through the C ABI
(engine.Comm).  ``mcsolve(..., devices=[0, 1, ...])`` drives all devices from this process
(one host thread per device); ``mcsolve_sharded`` is the one-process-per-device form.
"""
import threading

import numpy as np
import scipy.sparse as sp

from . import engine as E
from ._lib import QbError  # noqa: F401


def make_thresholds(seeds, ntraj, ndraws=64, bitgenerator=None, first=0):
    """The successive ``generator.random()`` values of trajectories [first, first+ntraj),
    drawn exactly as the reference does (multitraj.py:354-393): SeedSequence(seed).spawn
    then default_rng(child) (PCG64) or the named bit generator.  ``Generator.random(K)``
    returns the same K doubles as K scalar calls."""
    if isinstance(seeds, (list, tuple)):
        kids = [s if isinstance(s, np.random.SeedSequence) else np.random.SeedSequence(s)
                for s in seeds[first:first + ntraj]]
    else:
        ss = seeds if isinstance(seeds, np.random.SeedSequence) else np.random.SeedSequence(seeds)
        if ss.n_children_spawned == 0:
            # child i of spawn() is SeedSequence(entropy, spawn_key + (i,)): build the block
            # [first, first + ntraj) directly instead of spawning `first` children to drop them
            kids = [np.random.SeedSequence(ss.entropy, spawn_key=tuple(ss.spawn_key) + (i,),
                                           pool_size=ss.pool_size) for i in range(first, first + ntraj)]
        else:
            kids = ss.spawn(first + ntraj)[first:]
    out = np.empty((ntraj, ndraws), dtype=np.float64)
    for j, k in enumerate(kids):
        if bitgenerator:
            gen = np.random.Generator(getattr(np.random, bitgenerator)(k))
        else:
            gen = np.random.default_rng(k)
        out[j] = gen.random(ndraws)
    return out


def _as_op(m, fmt=E.FMT_AUTO):
    if isinstance(m, (E.DeviceOp, E.DeviceDense)):
        return m
    if sp.issparse(m):
        return E.DeviceOp.from_scipy(m, fmt)
    return E.DeviceDense.from_numpy(np.asfortranarray(np.asarray(m, dtype=np.complex128)))


def trace_functional(E_op):
    """diag(vec(E^T)): sum_r w[r] * vec(rho)[r] = tr(E rho) for a column-stacked rho
    (the mesolve branch of expect, core/data/expect.pyx:146-158)."""
    Ed = E_op.toarray() if sp.issparse(E_op) else np.asarray(E_op)
    w = Ed.T.reshape(-1, order="F")
    return sp.dia_matrix((w.reshape(1, -1), [0]), shape=(w.size, w.size))


def build_system(elements, c_ops=(), n_ops=None, e_ops=(), functional=False, nargs=0,
                 fmt=E.FMT_AUTO):
    """elements: list of operator or (operator, Program) pairs (time-dependent first,
    constant last, like QobjEvo.compress).  c_ops: operators (or (op, Program))."""
    first = elements[0][0] if isinstance(elements[0], (tuple, list)) else elements[0]
    N = first.shape[0]
    s = E.System(N, nargs)
    for el in elements:
        op, prog = el if isinstance(el, (tuple, list)) else (el, None)
        s.add_element(_as_op(op, fmt), prog)
    for i, c in enumerate(c_ops):
        cop, cprog = c if isinstance(c, (tuple, list)) else (c, None)
        if n_ops is not None:
            n = n_ops[i]
            nop, nprog = n if isinstance(n, (tuple, list)) else (n, None)
        else:
            cm = sp.csr_matrix(cop)
            nop, nprog = sp.csr_matrix(cm.conj().T @ cm), (cprog.norm() if cprog else None)
        s.add_collapse(_as_op(cop, fmt), _as_op(nop, fmt), cprog, nprog)
    for e in e_ops:
        op, prog = e if isinstance(e, (tuple, list)) else (e, None)
        s.add_eop(_as_op(op, fmt), prog)
    if functional:
        s.set_functional(True)
    return s


# jump part sum conj(C) (x) C: built as an explicit sparse superoperator up to this many
# non-zeros (sum nnz(C)^2), matrix-free ("sandwich") beyond.  The explicit form is ~20 % faster
# per RHS evaluation on C2 but costs a host-side Kronecker product and format conversion
# (0.27 s for 2.6 M non-zeros), which only long integrations amortise.
JUMP_EXPLICIT_MAX_NNZ = 1 << 16


def _jump_operator(ops, jump):
    nnz2 = sum(int(sp.csr_matrix(c).nnz) ** 2 for c in ops)
    if jump == "explicit" or (jump == "auto" and nnz2 <= JUMP_EXPLICIT_MAX_NNZ):
        S = None
        for c in ops:
            c = sp.csr_matrix(c, dtype=complex)
            K = sp.kron(c.conj(), c, format="csr")
            S = K if S is None else S + K
        S = sp.csr_matrix(S)
        S.sum_duplicates()
        S.sort_indices()
        return E.DeviceOp.from_scipy(S)
    return E.DeviceOp.sandwich(ops)


def lindblad_matrix_free(H_terms, c_ops, jump="auto"):
    """Elements of the Lindblad right-hand side in matrix form (LindbladMatrixForm,
    core/cy/lindblad_matrix_form.pyx:105-203) for ``build_system`` / ``mesolve``: nothing of
    the size of the Liouvillian's Hamiltonian part is built.  ``H_terms``: operators or
    (operator, Program) pairs summing to H; ``c_ops``: operators or (operator, Program).
    Returns [(DeviceOp, Program|None)]: per term of ``H_nh = H - i/2 sum c^dag c`` the
    matrix-free products ``-i f A rho`` and ``+i conj(f) rho A^dagger``, and the jump part
    ``sum |g|^2 C rho C^dagger`` -- ``jump="explicit"``: the sparse superoperator
    ``sum conj(C) (x) C`` (nnz = sum nnz(C)^2), ``"sandwich"``: matrix-free, ``"auto"``:
    explicit while it stays small (JUMP_EXPLICIT_MAX_NNZ)."""
    from . import coeffs
    const = None
    td = []
    for t in H_terms:
        op, prog = t if isinstance(t, (tuple, list)) else (t, None)
        op = sp.csr_matrix(op, dtype=complex)
        if prog is None:
            const = op if const is None else const + op
        else:
            td.append((op, prog))
    jump_const, jumps = [], []
    for c in c_ops:
        op, prog = c if isinstance(c, (tuple, list)) else (c, None)
        op = sp.csr_matrix(op, dtype=complex)
        cdc = sp.csr_matrix(op.conj().T @ op)
        if prog is None:
            const = -0.5j * cdc if const is None else const - 0.5j * cdc
            jump_const.append(op)
        else:
            td.append((-0.5j * cdc, prog.norm()))
            jumps.append((op, prog.norm()))
    out = []
    for op, prog in td:
        out.append((E.DeviceOp.kron(op, 0), prog.scaled(-1j)))
        out.append((E.DeviceOp.kron(op, 1), prog.conj().scaled(1j)))
    out += [(_jump_operator([op], jump), pg) for op, pg in jumps]
    if const is not None:
        const = sp.csr_matrix(const)
        const.sum_duplicates()
        out.append((E.DeviceOp.kron(const, 0), coeffs.constant(-1j)))
        out.append((E.DeviceOp.kron(const, 1), coeffs.constant(1j)))
    if jump_const:
        out.append((_jump_operator(jump_const, jump), None))
    return out


class McResult(dict):
    __getattr__ = dict.__getitem__


def default_nslots(N, method="vern7", fraction=0.6):
    """Resident trajectory slots when the caller does not say: as many as fit in ``fraction``
    of the free device memory (every slot owns S + 5 state-sized vectors), at most 16384."""
    V = {"vern7": 21, "vern9": 31, "tsit5": 12, "adams": 33}.get(method, 31)
    free, _ = E.mem_info()
    return int(max(1, min(16384, fraction * free // (V * N * 16 + 4096))))


def mcsolve(heff_elements, c_ops, psi0, tlist, ntraj, seeds=None, e_ops=(), method="vern7",
            nslots=None, ndraws=64, options=None, n_ops=None, draws=None, engine=None,
            devices=None, first=0):
    """Monte-Carlo trajectories on the device.  ``heff_elements`` is mcsolve's rhs
    (-iH - 1/2 sum c^dag c, solver/mcsolve.py:493-496).  Returns per-trajectory
    expectation values [n_e][ntraj][nt], their average / std as the reference computes
    them (multitrajresult.py:261-279,1116-1124) and the collapse records."""
    if devices is not None and len(devices) > 1:
        return mcsolve_multi(heff_elements, c_ops, psi0, tlist, ntraj, seeds, devices, e_ops=e_ops,
                             method=method, nslots=nslots, ndraws=ndraws, options=options,
                             n_ops=n_ops, draws=draws)
    if devices:
        E.set_device(devices[0])
    opts = dict(options or {})
    if engine is None:
        system = build_system(heff_elements, c_ops, n_ops, e_ops)
        nslots = min(ntraj, nslots or default_nslots(system.N, method))
        engine = E.Engine(system, method, nslots=nslots, **opts)
    if draws is None:
        draws = make_thresholds(seeds, ntraj, ndraws, first=first)
    r = engine.run_mcsolve(psi0, tlist, draws, ntraj=ntraj)
    # trajectories whose threshold table ran out are re-run with a longer one
    todo = np.nonzero(r.status == -12)[0]
    while todo.size:
        ndraws *= 4
        full = make_thresholds(seeds, ntraj, ndraws, first=first) if seeds is not None else None
        if full is None:
            raise QbError(-12, E.STATUS_MESSAGES[-12])
        r2 = engine.run_mcsolve(psi0, tlist, full[todo], ntraj=len(todo))
        for k in ("expect", "status", "ncol", "stats"):
            r[k][todo] = r2[k]
        if r.states is not None:
            r.states[todo] = r2.states
        w = r2.col_t.shape[1]
        r.col_t[todo, :w] = r2.col_t
        r.col_which[todo, :w] = r2.col_which
        todo = todo[r2.status == -12]
    bad = np.nonzero(r.status != 1)[0]
    if bad.size:
        st = int(r.status[bad[0]])
        raise QbError(st, "trajectory %d: %s" % (bad[0], E.STATUS_MESSAGES.get(st, "failed")))
    runs = np.transpose(r.expect, (1, 0, 2))
    avg = runs.mean(axis=1)
    avg2 = (np.abs(runs) ** 2).mean(axis=1) if np.iscomplexobj(runs) else (runs ** 2).mean(axis=1)
    std = np.sqrt(np.abs(avg2 - np.abs(avg) ** 2))
    col_times = [r.col_t[j, :r.ncol[j]].copy() for j in range(ntraj)]
    col_which = [r.col_which[j, :r.ncol[j]].copy() for j in range(ntraj)]
    return McResult(runs_expect=runs, average_expect=avg, std_expect=std, col_times=col_times,
                    col_which=col_which, ncol=r.ncol, stats=r.stats, rounds=r.rounds,
                    states=r.states, gpu_ms=r.gpu_ms, engine=engine)


def mesolve(elements, y0, tlist, e_ops=(), method="vern7", args=None, nargs=0,
            store_states=True, options=None, engine=None, ntraj=None, nslots=None):
    """Integrate d y/dt = (sum_k c_k(t) A_k) y through tlist for one system or -- with
    per-member ``args`` [nsys][nargs] -- a batched parameter sweep.  ``e_ops`` are n x n
    operators; tr(E rho) is evaluated on the device as a linear functional of the
    column-stacked rho."""
    opts = dict(options or {})
    y0 = np.atleast_2d(np.asarray(y0, dtype=np.complex128))
    nsys = ntraj or (len(args) if args is not None else y0.shape[0])
    if engine is None:
        fun = [trace_functional(e) for e in e_ops]
        system = build_system(elements, e_ops=fun, functional=True, nargs=nargs)
        engine = E.Engine(system, method, nslots=min(nsys, nslots or 4096),
                          store_states=int(bool(store_states)), **opts)
    init_map = np.arange(nsys, dtype=np.int32) % y0.shape[0] if y0.shape[0] > 1 else None
    r = engine.run_mesolve(y0, tlist, ntraj=nsys, args=args, init_map=init_map)
    bad = np.nonzero(r.status != 1)[0]
    if bad.size:
        st = int(r.status[bad[0]])
        raise QbError(st, "system %d: %s" % (bad[0], E.STATUS_MESSAGES.get(st, "failed")))
    return McResult(expect=r.expect, states=r.states, stats=r.stats, rounds=r.rounds,
                    gpu_ms=r.gpu_ms, engine=engine)


# ------------------------------------------------------------------ multi-GPU sharding
def shard_range(ntraj, rank, world):
    """Contiguous block of the spawned seed list for this rank."""
    per = (ntraj + world - 1) // world
    lo = min(ntraj, rank * per)
    return lo, min(ntraj, lo + per)


def local_expect_sums(runs_expect):
    """[2][n_e][nt]: (sum_j e_j, sum_j (re_j^2, im_j^2)) of runs_expect[n_e][ntraj][nt] -- the
    quantities _TrajectorySum.reduce_expect accumulates (multitrajresult.py:1116-1124), laid
    out like the device reduction (qb_reduce_expect)."""
    runs = np.asarray(runs_expect, dtype=np.complex128)
    return np.stack([runs.sum(axis=1), (runs.real ** 2).sum(axis=1) + 1j * (runs.imag ** 2).sum(axis=1)])


def finish_expect_sums(s1, s2, ntraj):
    """average and standard deviation as multitrajresult.py:261-279 forms them"""
    avg = s1 / ntraj
    avg2 = (s2.real + s2.imag) / ntraj
    return avg, np.sqrt(np.abs(avg2 - np.abs(avg) ** 2))


def reduce_expect_sums(runs_expect, comm=None, allreduce=None):
    """Expectation sums over all ranks with ONE collective: ``comm`` (engine.Comm, NCCL) on
    GPUs; ``allreduce`` (callable summing a float64 array in place over the ranks) lets the
    CPU tests exercise the same host logic over gloo."""
    local = np.ascontiguousarray(local_expect_sums(runs_expect))
    if comm is not None:
        comm.allreduce_sum([local])
    elif allreduce is not None:
        allreduce(local.view(np.float64).reshape(-1))
    return local[0], local[1]


def mcsolve_sharded(heff_elements, c_ops, psi0, tlist, ntraj, seeds, e_ops=(), rank=0, world=1,
                    comm=None, allreduce=None, **kw):
    """One process per device: this rank runs trajectories [lo, hi) of the global seed list;
    the averages come from one all-reduce of the expectation sums."""
    lo, hi = shard_range(ntraj, rank, world)
    draws = make_thresholds(seeds, hi - lo, kw.pop("ndraws", 64), first=lo)
    res = mcsolve(heff_elements, c_ops, psi0, tlist, hi - lo, e_ops=e_ops, draws=draws, **kw)
    if comm is not None and len(e_ops):
        s1, s2 = comm.reduce_expect([res.engine], len(e_ops), len(tlist))
    else:
        s1, s2 = reduce_expect_sums(res.runs_expect, allreduce=allreduce)
    res["global_average_expect"], res["global_std_expect"] = finish_expect_sums(s1, s2, ntraj)
    res["shard"] = (lo, hi)
    return res


def mcsolve_multi(heff_elements, c_ops, psi0, tlist, ntraj, seeds, devices, e_ops=(),
                  method="vern7", nslots=None, ndraws=64, options=None, n_ops=None, draws=None):
    """All ``devices`` of the box from ONE process: device d gets the contiguous block
    ``shard_range(ntraj, d, len(devices))`` of the seed list, its own copy of the operators
    and one host thread (the C ABI calls release the GIL); per-trajectory records are
    concatenated in seed order; the averages come from ONE ncclAllReduce of the device-side
    expectation sums (engine.Comm.reduce_expect)."""
    devices = list(devices)
    world = len(devices)
    comm = E.Comm.all(devices)
    parts = [None] * world
    errors = []

    def work(i):
        try:
            lo, hi = shard_range(ntraj, i, world)
            if hi <= lo:
                return
            E.set_device(devices[i])
            d = draws[lo:hi] if draws is not None else make_thresholds(seeds, hi - lo, ndraws, first=lo)
            parts[i] = mcsolve(heff_elements, c_ops, psi0, tlist, hi - lo, seeds=seeds, e_ops=e_ops,
                               method=method, nslots=nslots, ndraws=ndraws, options=options,
                               n_ops=n_ops, draws=d, first=lo)
        except BaseException as exc:      # re-raised in the calling thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    live = [p for p in parts if p is not None]
    out = McResult(
        runs_expect=np.concatenate([p.runs_expect for p in live], axis=1),
        col_times=[c for p in live for c in p.col_times],
        col_which=[c for p in live for c in p.col_which],
        ncol=np.concatenate([p.ncol for p in live]), stats=np.concatenate([p.stats for p in live]),
        rounds=max(p.rounds for p in live), gpu_ms=max(p.gpu_ms for p in live),
        states=None if live[0].states is None else np.concatenate([p.states for p in live]),
        engine=[None if p is None else p.engine for p in parts], devices=devices)
    if len(e_ops):
        s1, s2 = comm.reduce_expect(out.engine, len(e_ops), len(tlist))
        out["average_expect"], out["std_expect"] = finish_expect_sums(s1, s2, ntraj)
    else:
        out["average_expect"] = out["std_expect"] = np.zeros((0, len(tlist)))
    comm.free()
    return out
