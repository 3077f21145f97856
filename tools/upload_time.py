"""time DeviceOp.from_scipy (host format analysis + upload) for the C2 Liouvillian"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qutip_b200 as qb
from qutip_b200 import models
H, c_ops, _ = models.tfim(10)
L = models.liouvillian(H, c_ops)
qb.DeviceOp.from_scipy(models.liouvillian(*models.tfim(4)[:2]))      # CUDA init
for fmt, name in ((qb.FMT_AUTO, "auto"), (qb.FMT_RSELL, "rsell"), (qb.FMT_DIAM, "diam"), (qb.FMT_CSR, "csr")):
    t0 = time.perf_counter()
    op = qb.DeviceOp.from_scipy(L, fmt)
    t = time.perf_counter() - t0
    print(name, "%.3f s" % t, op.info()["format"], op.info()["device_bytes"])
    op.free()
t0 = time.perf_counter(); op = qb.DeviceOp.liouvillian(H, c_ops, qb.FMT_AUTO); print("device build + auto format %.3f s" % (time.perf_counter() - t0), op.info()["format"])
t0 = time.perf_counter(); op = qb.DeviceOp.liouvillian(H, c_ops, qb.FMT_DIAM); print("device build + DIAM %.3f s" % (time.perf_counter() - t0), op.info()["format"])
