#!/usr/bin/env python
"""C2 through QuTiP's own API with the plug-in: mesolve(..., options={"method": "b200_*",
"matrix_form": True}) on the dissipative TFIM chain, wall-clock including QuTiP's host work
and the binding (needs the reference build in oracle/_ref).  Prints one JSON line.
    python tools/plugin_c2_matrix_form.py [n_spins] [ref]     ref: also time the reference's
                                                               vern7 on t in [0, 0.1]"""
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
import qutip  # noqa: E402
from qutip import basis, qeye, sigmam, sigmax, sigmaz, tensor  # noqa: E402
import qutip_b200.plugin  # noqa: E402,F401

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
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
c_ops = [np.sqrt(0.1) * s for s in sm]
psi0 = basis([2] * n, [0] * n)
tl = np.linspace(0, 1, 11)
opt = dict(progress_bar="", store_states=False, store_final_state=False, matrix_form=True)
out = {"workload": "mesolve dissipative TFIM %d spins via qutip.mesolve + plug-in, matrix_form, "
                   "tlist linspace(0,1,11)" % n}
qutip.mesolve(H, psi0, tl[:2], c_ops, e_ops=[sz[0]], options=dict(opt, method="b200_vern7"))   # warm-up
for method in ("b200_vern7", "b200_adams"):
    t0 = time.perf_counter()
    r = qutip.mesolve(H, psi0, tl, c_ops, e_ops=[sz[0]], options=dict(opt, method=method))
    out[method] = {"wall_s": time.perf_counter() - t0, "expect_sz0_final": float(r.expect[0][-1])}
if len(sys.argv) > 2:
    tl2 = np.linspace(0, 0.1, 2)
    t0 = time.perf_counter()
    ref = qutip.mesolve(H, psi0, tl2, c_ops, e_ops=[sz[0]], options=dict(opt, method="vern7"))
    wall = time.perf_counter() - t0
    dev = qutip.mesolve(H, psi0, tl2, c_ops, e_ops=[sz[0]], options=dict(opt, method="b200_vern7"))
    out["reference_vern7_t0.1"] = {"wall_s": wall, "expect_sz0": float(ref.expect[0][-1]),
                                   "plugin_expect_sz0": float(dev.expect[0][-1])}
print(json.dumps(out))
