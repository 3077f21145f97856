"""ctypes wrapper around tests/emul/emul.cpp (host-side harness of the step controller;
test infrastructure only, see the header of emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from qutip_b200.coeffs import Program, QbInstr

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FMT_CSR, FMT_DIAM, FMT_SELL, FMT_RSELL = 0, 1, 3, 5


class QbOptions(C.Structure):
    _fields_ = [("atol", C.c_double), ("rtol", C.c_double), ("nsteps", C.c_int),
                ("first_step", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
                ("interpolate", C.c_int), ("norm_steps", C.c_int), ("norm_t_tol", C.c_double),
                ("norm_tol", C.c_double), ("norm_min_step", C.c_double),
                ("mc_corr_eps", C.c_double), ("store_states", C.c_int),
                ("max_collapses", C.c_int), ("no_jump", C.c_int),
                ("jump_prob_floor", C.c_double), ("max_order", C.c_int),
                ("pad_", C.c_int)]


def default_options(**kw):
    o = QbOptions(1e-8, 1e-6, 1000, 0.0, 0.0, 0.0, 1, 25, 1e-6, 1e-4, 0.1, 1e-10, 0, 64, 0, 0.0, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def lib():
    global _LIB
    if _LIB is None:
        src = os.path.join(HERE, "emul", "emul.cpp")
        out = os.path.join(HERE, "emul", "libemul.so")
        deps = [src] + [os.path.join(HERE, "..", "qutip_b200", "csrc", f)
                        for f in ("qb_control.h", "qb_coeff.h", "qb_types.h", "qb_diam.h",
                                  "qb_tableaux.h", "qb_adams.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out)
                                          for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread",
                                   "-o", out, src])
        _LIB = C.CDLL(out)
        _LIB.emul_diam_avg_lanes.restype = C.c_double
    return _LIB


def _csr(kind, shape, arrs):
    if kind == "csr":
        return (np.ascontiguousarray(arrs["data"], dtype=np.complex128),
                np.ascontiguousarray(arrs["col"], dtype=np.int32),
                np.ascontiguousarray(arrs["rowptr"], dtype=np.int32))
    if kind == "dia":
        m = sp.dia_matrix((arrs["data"], arrs["offsets"]), shape=shape).tocsr()
    else:
        m = sp.csr_matrix(arrs["arr"])
    m.sort_indices()
    return (m.data.astype(np.complex128), m.indices.astype(np.int32),
            m.indptr.astype(np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmulSystem:
    def __init__(self, N, nargs=0, fmt=FMT_DIAM):
        self.N, self.nargs, self.fmt = N, nargs, fmt
        self.L = lib()
        self.L.emul_reset(C.c_int64(N), nargs)
        self.neops = 0
        self.ncops = 0
        self._keep = []

    def add_element(self, kind, shape, arrs, prog=None):
        d, c, r = _csr(kind, shape, arrs)
        self._keep += [d, c, r]
        prog = prog or Program()
        self.L.emul_add_element(_p(d), _p(c), _p(r), self.fmt, prog.as_ctypes(), len(prog))

    def add_collapse(self, c_op, n_op, cprog=None, nprog=None):
        cd, cc, cr = _csr(*c_op)
        nd, nc, nr = _csr(*n_op)
        self._keep += [cd, cc, cr, nd, nc, nr]
        if cprog is None and nprog is None:
            self.L.emul_add_collapse(_p(cd), _p(cc), _p(cr), _p(nd), _p(nc), _p(nr), self.fmt)
        else:
            cprog, nprog = cprog or Program(), nprog or Program()
            self.L.emul_add_collapse_td(_p(cd), _p(cc), _p(cr), _p(nd), _p(nc), _p(nr), self.fmt,
                                        cprog.as_ctypes(), len(cprog), nprog.as_ctypes(), len(nprog))
        self.ncops += 1

    def add_eop(self, kind, shape, arrs):
        d, c, r = _csr(kind, shape, arrs)
        self._keep += [d, c, r]
        self.L.emul_add_eop(_p(d), _p(c), _p(r), self.fmt)
        self.neops += 1

    def set_functional(self, f):
        self.L.emul_set_functional(int(f))

    def set_mc_trace(self, n):
        self.L.emul_set_mc_trace(int(n))

    def rsell_stats(self, which):
        out = (C.c_longlong * 6)()
        self.L.emul_rsell_stats(which, out)
        return dict(slots=out[0], col_blocks=out[1], val_blocks=out[2], bytes=out[3], xor_slots=out[4],
                    unique_descriptors=out[5])

    def matvec(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.zeros(self.N, dtype=np.complex128)
        self.L.emul_matvec(which, _p(x), _p(y))
        return y

    def run(self, mode, tableau, init_states, tlist, ntraj=1, nslots=None, args=None,
            draws=None, opt=None, init_map=None, max_rounds=10 ** 7):
        opt = opt or default_options()
        nslots = nslots or ntraj
        init_states = np.ascontiguousarray(np.atleast_2d(init_states), dtype=np.complex128)
        tlist = np.ascontiguousarray(tlist, dtype=np.float64)
        nt = len(tlist)
        N = self.N
        expect = np.zeros((ntraj, max(1, self.neops), nt), dtype=np.complex128)
        status = np.zeros(ntraj, dtype=np.int32)
        ncol = np.zeros(ntraj, dtype=np.int32)
        col_t = np.zeros((ntraj, opt.max_collapses), dtype=np.float64)
        col_w = np.zeros((ntraj, opt.max_collapses), dtype=np.int32)
        stats = np.zeros((ntraj, 4), dtype=np.int32)
        states = np.zeros((ntraj, nt, N) if opt.store_states else (1,), dtype=np.complex128)
        ndraws = 0
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            ndraws = draws.shape[1]
        if args is not None:
            args = np.ascontiguousarray(args, dtype=np.complex128)
        if init_map is not None:
            init_map = np.ascontiguousarray(init_map, dtype=np.int32)
        rc = self.L.emul_run(
            mode, tableau, C.byref(opt), C.c_int64(ntraj), nslots, _p(init_states),
            _p(init_map) if init_map is not None else None, _p(tlist), nt,
            _p(args) if args is not None else None,
            _p(draws) if draws is not None else None, ndraws, _p(expect), _p(status),
            _p(ncol), _p(col_t), _p(col_w), _p(stats), _p(states), C.c_int64(max_rounds))
        if rc < 0:
            raise RuntimeError("emulator failed rc=%d" % rc)
        return dict(expect=expect[:, :self.neops], status=status, ncol=ncol, col_t=col_t,
                    col_which=col_w, stats=stats, states=states, rounds=rc)
