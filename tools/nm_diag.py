"""diagnostic: nm_mcsolve through the b200 map against the reference, per-trajectory differences"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref"))
import numpy as np
import qutip
from qutip import nm_mcsolve, sigmap, sigmam, sigmaz, sigmax, basis, coefficient
import qutip_b200.plugin  # noqa

H = 0.5 * sigmaz() + 0.2 * sigmax()
ops_and_rates = [(sigmam(), coefficient("0.25*sin(2*t) + 0.05")), (sigmap(), 0.15)]
psi0 = (basis(2, 0) + 0.5 * basis(2, 1)).unit()
tl = np.linspace(0, 4, 17)
for extra in [dict()]:
    o = dict(progress_bar=False, method="vern7", keep_runs_results=True, store_final_state=True, **extra)
    kw = dict(e_ops=[sigmaz(), sigmax()], ntraj=24)
    ref = nm_mcsolve(H, psi0, tl, ops_and_rates, seeds=np.random.SeedSequence(3), options=o, **kw)
    out = nm_mcsolve(H, psi0, tl, ops_and_rates, seeds=np.random.SeedSequence(3), options=dict(o, map="b200"), **kw)
    print("options", extra)
    print(" which equal", [list(w) for w in out.col_which] == [list(w) for w in ref.col_which])
    for i, (a, b) in enumerate(zip(out.col_times, ref.col_times)):
        a = np.asarray(a); b = np.asarray(b)
        if len(a) == len(b) and len(a):
            d = np.abs(a - b)
            if d.max() > 1e-10:
                print("  traj", i, "times", b, "diff", a - b, "which", list(ref.col_which[i]))
    print(" trace diff", np.abs(np.array(out.runs_trace) - np.array(ref.runs_trace)).max())
    print(" expect diff", np.abs(np.array(out.runs_expect) - np.array(ref.runs_expect)).max())
    print(" avg expect diff", np.abs(np.array(out.average_expect) - np.array(ref.average_expect)).max())
