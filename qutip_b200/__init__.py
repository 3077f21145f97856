"""qutip_b200 -- B200-native (sm_100a) accelerator for QuTiP's time-evolution hot path.

Scope: the complex128 Liouvillian / effective-Hamiltonian matvec (QobjEvo.matmul_data over
CSR, Dia and Dense operators, or matrix-free Kronecker operators for the Lindblad matrix
form) inside every explicit Runge-Kutta step (vern7 / vern9 / tsit5) and a device-resident
Adams method of mesolve and mcsolve, including mcsolve's norm-threshold jump detection and
collapse-operator selection.  All compute runs in hand-written CUDA kernels behind the C
ABI of ``libqutip_b200.so`` (include/qutip_b200.h); there is no CPU fallback.

Layers
  qutip_b200.engine   operators / dense states / systems / engines (numpy in, numpy out)
  qutip_b200.coeffs   coefficient compiler (strings, splines -> device byte-code)
  qutip_b200.solve    mesolve / mcsolve on plain arrays (batched, multi-GPU sharding)
  qutip_b200.plugin   registration with QuTiP's own extension surfaces (data-layer type,
                      Dispatcher specialisations, MESolver/MCSolver.add_integrator)
"""
from ._lib import QbError, LIB_PATH, device_count, launch_count  # noqa: F401
from .engine import (DeviceDense, DeviceOp, Engine, System, FMT_AUTO, FMT_CSR,  # noqa: F401
                     FMT_DIAM, FMT_SELL, FMT_RSELL, STATUS_MESSAGES, make_options)
from . import coeffs  # noqa: F401

__version__ = "0.1.0"
