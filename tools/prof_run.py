#!/usr/bin/env python
"""Small driver for ncu captures (never a source of bench numbers).
    python tools/prof_run.py c3 [ntraj]   one mcsolve batch of config C3
    python tools/prof_run.py c2|c2mf      C2 SpMV + one short mesolve (c2mf: matrix-free RHS)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qutip_b200 as qb  # noqa: E402
from qutip_b200 import models, solve  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "c3"
if what == "c3":
    ntraj = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    n = 14
    H, c_ops, sz = models.tfim(n)
    system = solve.build_system([models.heff(H, c_ops)], c_ops, e_ops=[sz[0]])
    eng = qb.Engine(system, os.environ.get("QB_METHOD", "vern7"), nslots=ntraj, nsteps=100000)
    draws = solve.make_thresholds(7, ntraj, 64)
    reps = int(os.environ.get("QB_REPS", "1"))
    ms = []
    for _ in range(reps):
        r = eng.run_mcsolve(models.basis_state(n), np.linspace(0, 2, 21), draws, ntraj=ntraj)
        ms.append(r.gpu_ms)
    print("c3", ntraj, "traj", r.rounds, "rounds", min(ms), "ms", (r.status == 1).all(),
          " ".join("%.0f" % m for m in ms))
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    H, c_ops, sz = models.tfim(n)
    L = models.liouvillian(H, c_ops)
    N = L.shape[0]
    system = qb.System(N)
    if what.startswith("c2mf"):             # matrix-free Lindblad right-hand side (c2mfs: sandwich jumps)
        for op_k, prog_k in solve.lindblad_matrix_free([H], c_ops, jump="sandwich" if what == "c2mfs" else "explicit"):
            system.add_element(op_k, prog_k)
    else:
        system.add_element(qb.DeviceOp.from_scipy(L))
    eng = qb.Engine(system, os.environ.get("QB_METHOD", "vern7"), nslots=1, nsteps=100000)
    x = qb.DeviceDense.from_numpy(np.random.default_rng(0).random(N) + 0j)
    out = qb.DeviceDense.zeros(N, 1)
    eng.rhs_bench(0.0, x, out, iters=3)
    print("spmv ms", eng.rhs_bench(0.0, x, out, iters=20) / 20)
    rho0 = np.zeros(N, dtype=complex); rho0[0] = 1
    reps = int(os.environ.get("QB_REPS", "1"))
    ms = []
    for _ in range(reps):
        r = eng.run_mesolve(rho0, np.linspace(0, float(os.environ.get("QB_TEND", "0.2")), 3))
        ms.append(r.gpu_ms)
    line = "c2 mesolve %s min %.2f ms (%s), %d rounds" % (r.stats[0], min(ms), " ".join("%.1f" % m for m in ms), r.rounds)
    if reps > 1:
        eng.set_profiling(True)
        r = eng.run_mesolve(rho0, np.linspace(0, 0.2, 3))
        pr = eng.profile()
        line += "; profiled: pass kernel %.1f us avg over %d launches, total %.2f ms" % (
            1e3 * pr["pass_ms"] / max(1, pr["pass_launches"]), pr["pass_launches"], r.gpu_ms)
    print(line)
