"""HEOMSolver (solver/heom/bofin_solvers.py) with the stock vern7 and with b200_vern7: the hierarchy
generator is one large constant sparse QobjEvo, i.e. the same matmul_data hot path.  Prints wall
times and the largest difference of the expectation values.  usage: heom_demo.py [max_depth] [Nk]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np
import qutip
from qutip import basis, sigmax, sigmaz, tensor, qeye
from qutip.solver.heom import DrudeLorentzBath, HEOMSolver
import qutip_b200.plugin  # noqa

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
Nk = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sz1, sz2 = tensor(sigmaz(), qeye(2)), tensor(qeye(2), sigmaz())
sx1, sx2 = tensor(sigmax(), qeye(2)), tensor(qeye(2), sigmax())
H = 0.5 * sz1 + 0.4 * sz2 + 0.25 * (sx1 * sx2)
baths = [DrudeLorentzBath(sz1, lam=0.05, gamma=0.5, T=1.0, Nk=Nk),
         DrudeLorentzBath(sz2, lam=0.04, gamma=0.7, T=1.0, Nk=Nk)]
rho0 = qutip.ket2dm(tensor(basis(2, 0), basis(2, 1)))
tl = np.linspace(0, 10, 41)
out = {}
for m in ("b200_vern7", "vern7"):
    t0 = time.perf_counter()
    s = HEOMSolver(H, baths, max_depth=depth, options=dict(progress_bar=False, method=m, atol=1e-8, rtol=1e-6))
    t1 = time.perf_counter()
    r = s.run(rho0, tl, e_ops=[sz1, sz2])
    t2 = time.perf_counter()
    out[m] = r
    print("%-11s hierarchy dim %d, nnz %d: build %.2f s, run %.3f s" % (
        m, s.rhs.shape[0], s.rhs(0).to("CSR").data.as_scipy().nnz, t1 - t0, t2 - t1), flush=True)
print("max |diff| of expectation values:",
      float(np.abs(np.array(out["vern7"].expect) - np.array(out["b200_vern7"].expect)).max()))
