#!/usr/bin/env python
"""C1 (damped Jaynes-Cummings, Liouvillian 400^2, 101 output times) through QuTiP's own mesolve:
the reference's vern7 on the host cores against method="b200_vern7" (whole tlist in one device
run, Integrator.run) and the one-call-per-output-time protocol.  Prints one JSON line."""
import json
import os
import sys
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

if oracle.ref_path() is None:
    print(json.dumps({"unavailable": "oracle/_ref not built"}))
    sys.exit(0)
sys.path.insert(0, oracle.ref_path())
import numpy as np  # noqa: E402
from qutip import basis, destroy, mesolve, qeye, sigmaz, tensor  # noqa: E402
import qutip_b200.plugin  # noqa: E402,F401

N = 10
a = tensor(destroy(N), qeye(2)); sm = tensor(qeye(N), destroy(2))
H = 2 * np.pi * a.dag() * a + 2 * np.pi * sm.dag() * sm + 2 * np.pi * 0.05 * (a.dag() * sm + a * sm.dag())
c_ops = [np.sqrt(0.1) * a, np.sqrt(0.05) * sm]
psi0 = tensor(basis(N, 3), basis(2, 0))
tl = np.linspace(0, 10, 101)
e_ops = [a.dag() * a, tensor(qeye(N), sigmaz())]


def best(method, reps=5, env=None):
    if env:
        os.environ.update(env)
    ts, r = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options={"method": method, "progress_bar": False})
        ts.append(time.perf_counter() - t0)
    if env:
        for k in env:
            os.environ.pop(k, None)
    return min(ts), np.array(r.expect)


t_ref, e_ref = best("vern7")
t_dev, e_dev = best("b200_vern7")
t_one, e_one = best("b200_vern7", env={"QUTIP_B200_NO_BATCHED_RUN": "1"})
print(json.dumps({"workload": "C1 mesolve damped JC N=10 (Liouvillian 400^2), 101 output times, vern7",
                  "reference_vern7_wall_ms": 1e3 * t_ref, "b200_vern7_wall_ms": 1e3 * t_dev,
                  "b200_vern7_one_call_per_time_wall_ms": 1e3 * t_one,
                  "max_abs_diff_expect": float(np.abs(e_dev - e_ref).max()),
                  "max_abs_diff_expect_one_call": float(np.abs(e_one - e_ref).max())}))
