#!/usr/bin/env python
"""Per-launch roofline of the pass kernel on C3: CUDA-event time and algorithmic bytes of every
round of one mcsolve batch (engine profiling mode), summarised by how full the launch was.
    python tools/round_profile.py [ntraj] [out.csv]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qutip_b200 as qb  # noqa: E402
from qutip_b200 import models, solve  # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = 14
H, c_ops, sz = models.tfim(n)
heff = models.heff(H, c_ops)
N = heff.shape[0]
system = solve.build_system([heff], c_ops, e_ops=[sz[0]])
eng = qb.Engine(system, "vern7", nslots=ntraj)
draws = solve.make_thresholds(7, ntraj, 64)
tl = np.linspace(0, 2, 21)
eng.run_mcsolve(models.basis_state(n), tl, draws, ntraj=ntraj)          # warm-up
eng.set_profiling(True)
r = eng.run_mcsolve(models.basis_state(n), tl, draws, ntraj=ntraj)
ms, acc = eng.profile_rounds()
op_alg = models.csr_algorithmic_bytes(heff.nnz, N, N) - 32 * N
gb = (acc * 16.0 * N + op_alg) / 1e9
peak = 6543.4
full = gb.max()
out = {"rounds": int(len(ms)), "total_ms": float(ms.sum()), "total_GB": float(gb.sum()),
       "overall_TBps": float(gb.sum() / ms.sum()), "overall_frac": float(gb.sum() / ms.sum() * 1e3 / peak)}
bins = [(0.9, 1.01), (0.6, 0.9), (0.3, 0.6), (0.1, 0.3), (0.0, 0.1)]
for lo, hi in bins:
    sel = (gb >= lo * full) & (gb < hi * full)
    if sel.any():
        out["launches with %.0f-%.0f%% of the heaviest launch's bytes" % (100 * lo, 100 * hi)] = {
            "n": int(sel.sum()), "ms": float(ms[sel].sum()), "GB": float(gb[sel].sum()),
            "frac_of_peak": float(gb[sel].sum() / ms[sel].sum() * 1e3 / peak)}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    np.savetxt(sys.argv[2], np.stack([ms, gb], axis=1), delimiter=",", header="ms,GB", comments="")
