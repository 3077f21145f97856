"""Build product-side systems (qutip_b200.System) from the golden fixtures."""
import numpy as np
import scipy.sparse as sp

import qutip_b200 as qb
from qutip_b200 import coeffs
from _golden import coeff_spec, op_arrays


def dev_op(kind, shape, a, fmt=qb.FMT_AUTO):
    if kind == "csr":
        return qb.DeviceOp.from_csr(a["data"], a["col"], a["rowptr"], shape, fmt)
    if kind == "dia":
        return qb.DeviceOp.from_dia(a["data"], a["offsets"], shape, fmt)
    return qb.DeviceDense.from_numpy(np.asfortranarray(a["arr"]))


def merged_constant_rhs(g):
    """The reference keeps -iH and each -0.5*n_op as separate constant elements
    (mcsolve.py:494-496); the plug-in sums constant elements at bind time."""
    tot = None
    for i in range(int(g["n_elements"])):
        assert str(g["el%d_coeff" % i]) == ""
        k, shape, a = op_arrays(g, "el%d" % i)
        m = sp.csr_matrix((a["data"], a["col"], a["rowptr"]), shape=shape)
        tot = m if tot is None else tot + m
    tot = sp.csr_matrix(tot)
    tot.sort_indices()
    return tot


def functional_of(E):
    """diag(w) with w = vec(E^T): sum_r w[r]*vec(rho)[r] = tr(E rho) for column-stacked rho."""
    w = np.asarray(E).T.reshape(-1, order="F")
    return sp.dia_matrix((w.reshape(1, -1), [0]), shape=(w.size, w.size))


def me_system_from_golden(g, fmt=qb.FMT_AUTO):
    N = len(g["y0"])
    s = qb.System(N)
    for i in range(int(g["n_elements"])):
        spec = coeff_spec(g["el%d_coeff" % i])
        prog = coeffs.compile_expr(spec[0], spec[1]) if spec is not None else None
        s.add_element(dev_op(*op_arrays(g, "el%d" % i), fmt=fmt), prog)
    for i in range(int(g["n_eops"])):
        s.add_eop(qb.DeviceOp.from_scipy(functional_of(g["eop%d" % i]), fmt))
    s.set_functional(True)
    return s


def mc_system_from_golden(g, fmt=qb.FMT_AUTO):
    N = len(g["psi0"])
    s = qb.System(N)
    s.add_element(qb.DeviceOp.from_scipy(merged_constant_rhs(g), fmt))
    for i in range(int(g["n_cops"])):
        s.add_collapse(dev_op(*op_arrays(g, "cop%d" % i), fmt=fmt),
                       dev_op(*op_arrays(g, "nop%d" % i), fmt=fmt))
    for i in range(int(g["n_eops"])):
        s.add_eop(dev_op(*op_arrays(g, "eop%d" % i), fmt=fmt))
    return s
