"""cProfile of the C1 mesolve through the plug-in (host-side breakdown of the 14 ms)"""
import cProfile, io, os, pstats, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
sys.path.insert(0, oracle.ref_path())
import numpy as np
from qutip import basis, destroy, mesolve, qeye, sigmaz, tensor
import qutip_b200.plugin  # noqa
N = 10
a = tensor(destroy(N), qeye(2)); sm = tensor(qeye(N), destroy(2))
H = 2 * np.pi * a.dag() * a + 2 * np.pi * sm.dag() * sm + 2 * np.pi * 0.05 * (a.dag() * sm + a * sm.dag())
c_ops = [np.sqrt(0.1) * a, np.sqrt(0.05) * sm]
psi0 = tensor(basis(N, 3), basis(2, 0))
tl = np.linspace(0, 10, 101)
e_ops = [a.dag() * a, tensor(qeye(N), sigmaz())]
opt = {"method": "b200_vern7", "progress_bar": False}
for _ in range(3):
    mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=opt)
t0 = time.perf_counter(); r = mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=opt); print("wall ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    mesolve(H, psi0, tl, c_ops, e_ops=e_ops, options=opt)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(32); print(s.getvalue()[:6000])
