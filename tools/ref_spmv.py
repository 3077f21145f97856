#!/usr/bin/env python
"""The reference's own CPU kernel for metric (2): `qutip.core.data.matmul(L, rho)` with the C2
Liouvillian as CSR (matmul_csr_dense_dense -> _matmul_csr_vector,
core/data/matmul.pyx:226-272, src/matmul_csr_vector.cpp), timed on the box's host.
Needs the reference build in oracle/_ref.  Prints one JSON line.
    python tools/ref_spmv.py [n_spins] [repetitions]"""
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
from qutip.core import data as _data  # noqa: E402
from qutip_b200 import models  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, c_ops, _ = models.tfim(n)
L = models.liouvillian(H, c_ops)
N = L.shape[0]
Lq = _data.CSR((L.data, L.indices.astype(np.int32), L.indptr.astype(np.int32)), shape=L.shape, copy=False)
rng = np.random.default_rng(0)
x = _data.Dense((rng.random(N) + 1j * rng.random(N)).reshape(-1, 1))
y = _data.matmul(Lq, x)                       # warm-up
t0 = time.perf_counter()
for _ in range(reps):
    y = _data.matmul(Lq, x)
per = (time.perf_counter() - t0) / reps
ref = L @ x.to_array().ravel()
alg = models.csr_algorithmic_bytes(L.nnz, N, N)
print(json.dumps({"kernel": "qutip.core.data.matmul (CSR x Dense, reference CPU, 1 thread)",
                  "n_spins": n, "nnz": int(L.nnz), "ms": per * 1e3, "gbs": alg / per / 1e9,
                  "repetitions": reps,
                  "max_abs_diff_vs_scipy": float(np.max(np.abs(y.to_array().ravel() - ref)))}))
