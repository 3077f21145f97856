"""Host-side mirror of the device objects: operators, dense states, systems, engines.

Everything here is a thin owner of an opaque handle of libqutip_b200.so; numpy arrays go
in and come out, all arithmetic happens in the CUDA kernels.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import QbOptions, as_c128, check, ptr
from .coeffs import Program

FMT_AUTO, FMT_CSR, FMT_DIAM, FMT_SELL = 0, 1, 2, 3
FMT_RSELL = 5     # rule-compressed sliced ELLPACK (qb_types.h); same code as upload format and as reported format
FMT_KRON = 4      # reported by DeviceOp.info() for matrix-free Kronecker operators
FMT_NAMES = {0: "csr", 1: "diam", 2: "dense", 3: "sell", 4: "kron", 5: "rsell"}
TABLEAUX = {"vern7": 0, "vern9": 1, "tsit5": 2, "adams": 3}

STATUS_MESSAGES = {
    # texts of explicit_rk.pyx:476-492 so callers can raise the reference's messages
    2: "Internal state at the desired time.",
    1: "Internal state past the desired time and interpolation to step done.",
    0: "No work done.",
    -1: ("Too much work done in one call. Try to increase the nsteps parameter or "
         "increasing the tolerance."),
    -2: "Step size becomes too small. Try increasing tolerance.",
    -3: "Step outside available range.",
    -4: "Not initialized.",
    -10: ("Could not find the collapse time within desired tolerance. Increase accuracy of "
          "the ODE solver or lower the tolerance with the options 'norm_steps', 'norm_tol', "
          "'norm_t_tol'."),
    -11: "collapse operator selection ran past the last operator (IndexError in the reference)",
    -12: "threshold table exhausted: more random draws are needed for this trajectory",
    -13: "more collapses than max_collapses in one trajectory",
    -14: "coefficient program failed on the device",
    -15: ("Repeated convergence failures of the Adams corrector (perhaps wrong step size or "
          "tolerances too tight)."),
}


class _Handle:
    def __init__(self, h):
        self._h = h

    @property
    def handle(self):
        if self._h is None:
            raise ValueError("handle already freed")
        return self._h

    def free(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().qb_free(self._h)
            except Exception:
                pass
            self._h = None

    def __del__(self):
        self.free()


class DeviceDense(_Handle):
    """complex128 matrix on the device (mirror of core/data/dense.pxd:9-22)."""

    def __init__(self, h, shape, fortran):
        super().__init__(h)
        self.shape = tuple(shape)
        self.fortran = bool(fortran)

    @classmethod
    def from_numpy(cls, arr):
        arr = np.asarray(arr, dtype=np.complex128)
        if arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        fortran = arr.flags.f_contiguous and not (arr.flags.c_contiguous and arr.shape[1] != 1)
        if not (arr.flags.f_contiguous or arr.flags.c_contiguous):
            arr = np.asfortranarray(arr)
            fortran = True
        h = C.c_void_p()
        check(_lib.load().qb_dense_upload(ptr(arr), arr.shape[0], arr.shape[1], int(fortran),
                                          C.byref(h)))
        return cls(h, arr.shape, fortran)

    @classmethod
    def zeros(cls, rows, cols, fortran=True):
        h = C.c_void_p()
        check(_lib.load().qb_dense_zeros(rows, cols, int(fortran), C.byref(h)))
        return cls(h, (rows, cols), fortran)

    def to_numpy(self):
        out = np.empty(self.shape, dtype=np.complex128, order="F" if self.fortran else "C")
        check(_lib.load().qb_dense_download(self.handle, ptr(out)))
        return out

    def copy(self):
        h = C.c_void_p()
        check(_lib.load().qb_dense_copy(self.handle, C.byref(h)))
        return DeviceDense(h, self.shape, self.fortran)

    def reshape(self, rows, cols):
        """in-place change of shape of a column-major buffer (stack / unstack columns)"""
        check(_lib.load().qb_dense_reshape(self.handle, int(rows), int(cols)))
        self.shape, self.fortran = (int(rows), int(cols)), True
        return self

    def write(self, arr):
        """overwrite the device buffer from a host array of the same size"""
        arr = as_c128(arr)
        if arr.size != self.shape[0] * self.shape[1]:
            raise ValueError("size mismatch")
        check(_lib.load().qb_dense_write(self.handle, ptr(arr)))

    def read_into(self, out):
        check(_lib.load().qb_dense_download(self.handle, ptr(out)))
        return out


class DeviceOp(_Handle):
    """Sparse operator on the device: diagonal-masked slices or CSR."""

    def __init__(self, h, shape):
        super().__init__(h)
        self.shape = tuple(shape)

    @classmethod
    def from_csr(cls, data, col, rowptr, shape, fmt=FMT_AUTO):
        data = as_c128(data)
        col = np.ascontiguousarray(col, dtype=np.int32)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        h = C.c_void_p()
        check(_lib.load().qb_csr_upload(ptr(data), ptr(col), ptr(rowptr), shape[0], shape[1],
                                        int(rowptr[-1]), fmt, C.byref(h)))
        return cls(h, shape)

    @classmethod
    def from_dia(cls, data, offsets, shape, fmt=FMT_AUTO):
        data = as_c128(data)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if data.ndim != 2 or data.shape != (len(offsets), shape[1]):
            raise ValueError("Dia data must be [num_diag][ncols]")
        h = C.c_void_p()
        check(_lib.load().qb_dia_upload(ptr(data), ptr(offsets), len(offsets), shape[0],
                                        shape[1], fmt, C.byref(h)))
        return cls(h, shape)

    @classmethod
    def kron(cls, m, side):
        """Matrix-free superoperator of the n x n operator ``m`` on a column-stacked n x n
        state: ``side=0`` is ``I (x) m`` (rho -> m rho), ``side=1`` is ``conj(m) (x) I``
        (rho -> rho m^dagger).  Nothing of size n^2 x n^2 is built."""
        import scipy.sparse as sp
        m = sp.csr_matrix(m)
        m.sort_indices()
        n = m.shape[0]
        if m.shape[0] != m.shape[1]:
            raise ValueError("Kronecker operator needs a square matrix")
        data = as_c128(m.data)
        col = np.ascontiguousarray(m.indices, dtype=np.int32)
        rowptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
        h = C.c_void_p()
        check(_lib.load().qb_kron_upload(ptr(data), ptr(col), ptr(rowptr), n, int(rowptr[-1]),
                                         int(side), C.byref(h)))
        return cls(h, (n * n, n * n))

    @classmethod
    def sandwich(cls, ops):
        """Matrix-free ``rho -> sum_c C_c rho C_c^dagger`` on a column-stacked n x n state for
        the list of n x n operators ``ops`` (the jump part of the Lindblad equation)."""
        import scipy.sparse as sp
        ops = [sp.csr_matrix(c) for c in ops]
        n = ops[0].shape[0]
        if any(c.shape != (n, n) for c in ops):
            raise ValueError("sandwich operators must all be n x n")
        m = sp.vstack(ops, format="csr")
        m.sort_indices()
        data = as_c128(m.data)
        col = np.ascontiguousarray(m.indices, dtype=np.int32)
        rowptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
        h = C.c_void_p()
        check(_lib.load().qb_sandwich_upload(ptr(data), ptr(col), ptr(rowptr), n, len(ops),
                                             int(rowptr[-1]), C.byref(h)))
        return cls(h, (n * n, n * n))

    @classmethod
    def liouvillian(cls, H=None, c_ops=(), fmt=FMT_AUTO, tol=0.0):
        """The Liouvillian ``qutip.liouvillian(H, c_ops)`` (core/superoperator.py:116-142) of
        n x n matrices, assembled ON THE DEVICE: only the n x n effective generator
        ``A = -iH - 1/2 sum C^dagger C`` is formed on the host; the Kronecker rows are counted,
        filled, sorted and merged by one thread per row.  ``fmt=FMT_CSR`` keeps the device CSR
        as it is; the compressed formats pass through the host slice analysers.  ``tol`` is the
        reference's ``auto_tidyup_atol`` (0: drop exact zeros only, as scipy's sparse add does)."""
        import scipy.sparse as sp
        c_ops = [sp.csr_matrix(c, dtype=np.complex128) for c in c_ops]
        if H is None and not c_ops:
            raise ValueError("The liouvillian need an Hamiltonian and/or c_ops")
        n = (sp.csr_matrix(H) if H is not None else c_ops[0]).shape[0]
        a = sp.csr_matrix((n, n), dtype=np.complex128)
        if H is not None:
            H = sp.csr_matrix(H, dtype=np.complex128)
            if H.shape != (n, n):
                raise ValueError("H must be square")
            a = a + (-1j) * H
        for c in c_ops:
            if c.shape != (n, n):
                raise ValueError("collapse operators must all be n x n")
            a = a - 0.5 * (c.conj().T @ c)
        a = sp.csr_matrix(a)
        a.sum_duplicates()
        a.sort_indices()
        if c_ops:
            cs = sp.vstack(c_ops, format="csr")
            cs.sum_duplicates()
            cs.sort_indices()
        else:
            cs = sp.csr_matrix((0, n), dtype=np.complex128)
        ad, ac, ap = as_c128(a.data), a.indices.astype(np.int32), a.indptr.astype(np.int32)
        cd, cc, cp = as_c128(cs.data), cs.indices.astype(np.int32), cs.indptr.astype(np.int32)
        h = C.c_void_p()
        check(_lib.load().qb_liouvillian_build(ptr(ad), ptr(ac), ptr(ap), int(ap[-1]),
                                               ptr(cd), ptr(cc), ptr(cp), int(cp[-1]) if len(cp) else 0,
                                               n, len(c_ops), float(tol), fmt, C.byref(h)))
        return cls(h, (n * n, n * n))

    @classmethod
    def kron_product(cls, a, b, fmt=FMT_CSR):
        """``kron(a, b)`` (core/data/kron.pyx) of two sparse matrices assembled on the device;
        ``fmt=FMT_CSR`` / ``FMT_DIAM`` never visit the host."""
        import scipy.sparse as sp
        mats = []
        for m in (a, b):
            m = sp.csr_matrix(m, dtype=np.complex128)
            m.sum_duplicates()
            m.sort_indices()
            mats.append((as_c128(m.data), m.indices.astype(np.int32), m.indptr.astype(np.int32), m.shape))
        (ad, ac, ap, ash), (bd, bc, bp, bsh) = mats
        h = C.c_void_p()
        check(_lib.load().qb_kron_build(ptr(ad), ptr(ac), ptr(ap), ash[0], ash[1], ptr(bd), ptr(bc), ptr(bp),
                                        bsh[0], bsh[1], int(fmt), C.byref(h)))
        return cls(h, (ash[0] * bsh[0], ash[1] * bsh[1]))

    def convert(self, fmt):
        """A CSR-format operator in another device format (``FMT_DIAM``: converted on the
        device; the others pass through the host analysers)."""
        h = C.c_void_p()
        check(_lib.load().qb_op_convert(self.handle, int(fmt), C.byref(h)))
        return type(self)(h, self.shape)

    def to_scipy(self):
        """Copy a CSR-format operator back to the host (scipy.sparse.csr_matrix)."""
        import scipy.sparse as sp
        inf = self.info()
        data = np.empty(inf["nnz"], dtype=np.complex128)
        col = np.empty(inf["nnz"], dtype=np.int32)
        rowptr = np.empty(inf["rows"] + 1, dtype=np.int32)
        check(_lib.load().qb_op_csr_download(self.handle, ptr(data), ptr(col), ptr(rowptr)))
        return sp.csr_matrix((data, col, rowptr), shape=(inf["rows"], inf["cols"]))

    @classmethod
    def from_scipy(cls, m, fmt=FMT_AUTO):
        import scipy.sparse as sp
        if isinstance(m, (sp.dia_matrix, sp.dia_array)):
            return cls.from_dia(m.data, m.offsets, m.shape, fmt)
        m = sp.csr_matrix(m)
        return cls.from_csr(m.data, m.indices, m.indptr, m.shape, fmt)

    def info(self):
        fmt, rows, cols = C.c_int(), C.c_int64(), C.c_int64()
        nnz, nbytes = C.c_int64(), C.c_int64()
        check(_lib.load().qb_op_info(self.handle, C.byref(fmt), C.byref(rows), C.byref(cols),
                                     C.byref(nnz), C.byref(nbytes)))
        return dict(format=FMT_NAMES[fmt.value], fmt=fmt.value, rows=rows.value, cols=cols.value,
                    nnz=nnz.value, device_bytes=nbytes.value)


# --------------------------------------------------------------------- data-layer ops
def matmul(op, x, scale=1.0, out=None):
    """out += scale * op @ x  (matmul.pyx:226-347,429-507); allocates zeros if out is None."""
    if out is None:
        out = DeviceDense.zeros(op.shape[0], x.shape[1], x.fortran or x.shape[1] == 1)
    s = complex(scale)
    check(_lib.load().qb_matmul(op.handle, x.handle, s.real, s.imag, out.handle))
    return out


def zgemm_bench(a, x, out, iters=10):
    """milliseconds for ``iters`` back-to-back dense ZGEMMs out += a @ x (CUDA events)."""
    ms = C.c_double()
    check(_lib.load().qb_zgemm_bench(a.handle, x.handle, out.handle, int(iters), C.byref(ms)))
    return ms.value


def dmma_peak_tflops(iters=4096):
    """measured FP64 tensor-core (DMMA m8n8k4) peak of the current device, TFLOP/s"""
    t = C.c_double()
    check(_lib.load().qb_dmma_peak_bench(int(iters), C.byref(t)))
    return t.value


def axpy(x, a, y):
    a = complex(a)
    check(_lib.load().qb_axpy(x.handle, a.real, a.imag, y.handle))
    return y


def scal(x, a):
    a = complex(a)
    check(_lib.load().qb_scal(x.handle, a.real, a.imag))
    return x


def nrm2(x):
    out = C.c_double()
    check(_lib.load().qb_nrm2(x.handle, C.byref(out)))
    return out.value


def wrms_error(diff, state, atol, rtol):
    out = C.c_double()
    check(_lib.load().qb_wrms_error(diff.handle, state.handle, atol, rtol, C.byref(out)))
    return out.value


def _c2(fn, *args):
    out = (C.c_double * 2)()
    check(fn(*args, out))
    return complex(out[0], out[1])


def inner(a, b, conj=True):
    return _c2(_lib.load().qb_inner, a.handle, b.handle, int(conj))


def expect_ket(op, x):
    return _c2(_lib.load().qb_expect_ket, op.handle, x.handle)


def expect_dm(op, rho):
    return _c2(_lib.load().qb_expect_dm, op.handle, rho.handle)


def expect_super(op, vec):
    return _c2(_lib.load().qb_expect_super, op.handle, vec.handle)


def trace_oper_ket(vec):
    return _c2(_lib.load().qb_trace_oper_ket, vec.handle)


# --------------------------------------------------------------------- system / engine
def _prog_args(prog):
    if prog is None or len(prog) == 0:
        return None, 0
    return prog.as_ctypes(), len(prog)


class System(_Handle):
    """What QobjEvo.matmul_data sums, plus mcsolve's collapse operators and the e_ops."""

    def __init__(self, N, nargs=0):
        h = C.c_void_p()
        check(_lib.load().qb_system_create(N, nargs, C.byref(h)))
        super().__init__(h)
        self.N, self.nargs = int(N), int(nargs)
        self.nelem = self.ncops = self.neops = 0
        self.functional = False
        self._keep = []     # operators must outlive the system
        self.element_ops = []

    def add_element(self, op, prog=None):
        p, n = _prog_args(prog)
        check(_lib.load().qb_system_add_element(self.handle, op.handle, p, n))
        self._keep.append(op)
        self.element_ops.append(op)
        self.nelem += 1

    def add_collapse(self, c_op, n_op, cprog=None, nprog=None):
        cp, cn = _prog_args(cprog)
        np_, nn = _prog_args(nprog)
        check(_lib.load().qb_system_add_collapse(self.handle, c_op.handle, cp, cn,
                                                 n_op.handle, np_, nn))
        self._keep += [c_op, n_op]
        self.ncops += 1

    def add_eop(self, op, prog=None):
        p, n = _prog_args(prog)
        check(_lib.load().qb_system_add_eop(self.handle, op.handle, p, n))
        self._keep.append(op)
        self.neops += 1

    def set_functional(self, flag=True):
        check(_lib.load().qb_system_set_eop_functional(self.handle, int(flag)))
        self.functional = bool(flag)

    def set_mc_trace(self, n):
        """mcsolve of a super-operator H: the state is the column-stacked n x n rho and tr(rho)
        takes the place of the squared norm (solver/mcsolve.py:311-319,481-490)"""
        check(_lib.load().qb_system_set_mc_trace(self.handle, int(n)))
        self.mc_trace = int(n)

    def add_spline(self, tlist, poly, dt=0.0):
        tlist = np.ascontiguousarray(tlist, dtype=np.float64)
        poly = as_c128(poly)
        if poly.ndim == 1:
            poly = poly.reshape(1, -1)
        sid = C.c_int()
        check(_lib.load().qb_system_add_spline(self.handle, ptr(tlist), ptr(poly), len(tlist),
                                               poly.shape[0] - 1, float(dt), C.byref(sid)))
        return sid.value


def make_options(**kw):
    o = QbOptions()
    check(_lib.load().qb_options_default(C.byref(o)))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError("unknown engine option %r" % k)
        setattr(o, k, v)
    return o


class RunResult(dict):
    __getattr__ = dict.__getitem__


class Engine(_Handle):
    """Fused RK / Monte-Carlo engine bound to one system."""

    def __init__(self, system, method="vern7", nslots=1, **options):
        if method not in TABLEAUX:
            raise ValueError("unknown method %r (device engine has %s)" % (method, list(TABLEAUX)))
        self.system = system
        self.method = method
        self.nslots = int(nslots)
        self.opt = make_options(**options)
        h = C.c_void_p()
        check(_lib.load().qb_engine_create(system.handle, TABLEAUX[method], self.nslots,
                                           C.byref(self.opt), C.byref(h)))
        super().__init__(h)

    # ---- batched runs ------------------------------------------------------------
    def run(self, mode, init_states, tlist, ntraj=1, init_map=None, args=None, draws=None,
            final_states=False):
        s = self.system
        init_states = as_c128(np.atleast_2d(init_states))
        if init_states.shape[1] != s.N:
            raise ValueError("incompatible state size %d for a system of size %d"
                             % (init_states.shape[1], s.N))
        tlist = np.ascontiguousarray(tlist, dtype=np.float64)
        nt = len(tlist)
        maxcol = self.opt.max_collapses
        expect = np.zeros((ntraj, max(1, s.neops), nt), dtype=np.complex128)
        status = np.zeros(ntraj, dtype=np.int32)
        ncol = np.zeros(ntraj, dtype=np.int32)
        col_t = np.zeros((ntraj, maxcol), dtype=np.float64)
        col_w = np.zeros((ntraj, maxcol), dtype=np.int32)
        stats = np.zeros((ntraj, 4), dtype=np.int32)
        states = (np.zeros((ntraj, nt, s.N), dtype=np.complex128)
                  if self.opt.store_states else None)
        fstates = np.zeros((ntraj, s.N), dtype=np.complex128) if final_states else None
        ndraws = 0
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            if draws.shape[0] != ntraj:
                raise ValueError("draws must be [ntraj][ndraws]")
            ndraws = draws.shape[1]
        if args is not None:
            args = as_c128(args).reshape(ntraj, s.nargs)
        if init_map is not None:
            init_map = np.ascontiguousarray(init_map, dtype=np.int32)
        check(_lib.load().qb_engine_run(
            self.handle, mode, ntraj, ptr(init_states), init_states.shape[0], ptr(init_map),
            ptr(tlist), nt, ptr(args), ptr(draws), ndraws, ptr(expect), ptr(status), ptr(ncol),
            ptr(col_t), ptr(col_w), ptr(stats), ptr(states), ptr(fstates)))
        rounds, ms = C.c_int64(), C.c_double()
        check(_lib.load().qb_engine_last_run_info(self.handle, C.byref(rounds), C.byref(ms)))
        return RunResult(expect=expect[:, :s.neops], status=status, ncol=ncol, col_t=col_t,
                         col_which=col_w, stats=stats, states=states, final_states=fstates,
                         rounds=rounds.value, gpu_ms=ms.value)

    def run_mesolve(self, y0, tlist, **kw):
        return self.run(0, y0, tlist, **kw)

    def run_mcsolve(self, psi0, tlist, draws, ntraj=None, **kw):
        ntraj = len(draws) if ntraj is None else ntraj
        return self.run(1, psi0, tlist, ntraj=ntraj, draws=draws, **kw)

    # ---- Integrator protocol (slot 0) ---------------------------------------------
    def set_args(self, args):
        args = as_c128(args)
        check(_lib.load().qb_integ_set_args(self.handle, ptr(args)))

    # host_coeffs: callable t -> sequence of nelem complex values, used when some element's
    # coefficient is a python callable (program coeffs.host()); see qb_integ_resume
    host_coeffs = None

    def _serve_coefficients(self, status, t_out=0.0):
        lib = _lib.load()
        while status == 3:
            if self.host_coeffs is None:
                raise _lib.QbError(-6, "engine needs host-evaluated coefficients but none are bound")
            t = C.c_double()
            check(lib.qb_integ_pending_coef(self.handle, C.byref(t)))
            vals = as_c128(np.asarray(self.host_coeffs(t.value), dtype=complex))
            tt, st = C.c_double(), C.c_int()
            check(lib.qb_integ_resume(self.handle, ptr(vals), C.byref(tt), C.byref(st)))
            t_out, status = tt.value, st.value
        return t_out, status

    def set_state(self, t, y):
        y = as_c128(y).reshape(-1)
        if y.size != self.system.N:
            raise ValueError("incompatible state size")
        rc = _lib.load().qb_integ_set_state(self.handle, float(t), ptr(y))
        if rc == 3:
            _, st = self._serve_coefficients(3)
            if st < 0:
                raise _lib.QbError(st, STATUS_MESSAGES.get(st, "set_state failed"))
            return
        check(rc)

    def integrate(self, t, step=False):
        """returns (t_reached, status)."""
        t_out, st = C.c_double(), C.c_int()
        check(_lib.load().qb_integ_integrate(self.handle, float(t), int(step), C.byref(t_out),
                                             C.byref(st)))
        return self._serve_coefficients(st.value, t_out.value)

    def get_state(self):
        t = C.c_double()
        y = np.empty(self.system.N, dtype=np.complex128)
        check(_lib.load().qb_integ_get_state(self.handle, C.byref(t), ptr(y)))
        return t.value, y

    def stats(self):
        st = (C.c_int64 * 4)()
        check(_lib.load().qb_integ_stats(self.handle, st))
        return dict(rhs_evals=st[0], accepted=st[1], rejected=st[2], passes=st[3])

    def rhs(self, t, x, out):
        check(_lib.load().qb_engine_rhs(self.handle, float(t), x.handle, out.handle))
        return out

    def rhs_coef(self, vals, x, out):
        vals = as_c128(vals)
        check(_lib.load().qb_engine_rhs_coef(self.handle, ptr(vals), x.handle, out.handle))
        return out

    def rhs_bench(self, t, x, out, iters=20):
        """milliseconds for ``iters`` back-to-back RHS evaluations (CUDA events)."""
        ms = C.c_double()
        check(_lib.load().qb_engine_rhs_bench(self.handle, float(t), x.handle, out.handle,
                                              int(iters), C.byref(ms)))
        return ms.value

    def set_profiling(self, on=True):
        check(_lib.load().qb_engine_set_profiling(self.handle, int(on)))

    def profile_rounds(self, max_rounds=8192):
        """(ms, vector accesses) per pass launch of the last profiled run"""
        ms = np.zeros(max_rounds); cum = np.zeros(max_rounds)
        n = C.c_int64()
        check(_lib.load().qb_engine_profile_rounds(self.handle, ptr(ms), ptr(cum), max_rounds, C.byref(n)))
        k = min(max_rounds, n.value)
        acc = np.diff(np.concatenate([[0.0], cum[:k]]))
        return ms[:k], acc

    def profile(self):
        ms, n, v = C.c_double(), C.c_int64(), C.c_double()
        check(_lib.load().qb_engine_profile(self.handle, C.byref(ms), C.byref(n), C.byref(v)))
        return dict(pass_ms=ms.value, pass_launches=n.value, state_vector_accesses=v.value)


# --------------------------------------------------------------------- multi-GPU
def set_device(dev):
    """make ``dev`` the current CUDA device of the calling thread (handles created afterwards
    live there; engine calls switch to their own device by themselves)."""
    check(_lib.load().qb_set_device(int(dev)))


def mem_info():
    """(free, total) bytes of the current device"""
    free, total = C.c_int64(), C.c_int64()
    check(_lib.load().qb_device_mem_info(C.byref(free), C.byref(total)))
    return free.value, total.value


class Comm(_Handle):
    """NCCL group for the one collective of the sharded workloads: the sum of the expectation
    sums over the devices (solver/multitrajresult.py:1116-1124 across shards; C ABI
    ``qb_comm_*``).  ``Comm.all(devices)``: this process drives all ``devices``
    (ncclCommInitAll).  ``Comm.rank(nranks, rank, uid)``: one process per device; ``uid`` is
    ``Comm.unique_id()`` of rank 0, distributed by the launcher."""

    ID_BYTES = 128

    def __init__(self, h, nranks, nlocal):
        super().__init__(h)
        self.nranks, self.nlocal = nranks, nlocal

    @classmethod
    def all(cls, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(_lib.load().qb_comm_init_all(len(devices), devs, C.byref(h)))
        return cls(h, len(devices), len(devices))

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * Comm.ID_BYTES)()
        check(_lib.load().qb_comm_unique_id(buf, Comm.ID_BYTES))
        return bytes(buf)

    @classmethod
    def rank(cls, nranks, rank, uid):
        buf = (C.c_ubyte * Comm.ID_BYTES).from_buffer_copy(bytes(uid))
        h = C.c_void_p()
        check(_lib.load().qb_comm_init_rank(int(nranks), int(rank), buf, C.byref(h)))
        return cls(h, int(nranks), 1)

    def allreduce_sum(self, arrays):
        """in-place sum over the group of one float64 / complex128 array per local member"""
        arrays = list(arrays)
        if len(arrays) != self.nlocal:
            raise ValueError("one array per local member")
        views = [a.view(np.float64).reshape(-1) for a in arrays]
        if any(not a.flags.c_contiguous for a in arrays) or len({v.size for v in views}) != 1:
            raise ValueError("arrays must be contiguous and of equal size")
        ptrs = (C.c_void_p * self.nlocal)(*[v.ctypes.data for v in views])
        check(_lib.load().qb_comm_allreduce_sum(self.handle, ptrs, views[0].size))
        return arrays

    def allreduce_sum_device(self, device_pointers, count):
        """in-place sum over the group of ``count`` doubles at one device pointer per member"""
        ptrs = (C.c_void_p * self.nlocal)(*[int(p) for p in device_pointers])
        check(_lib.load().qb_comm_allreduce_sum_device(self.handle, ptrs, int(count)))

    def reduce_expect(self, engines, neops, nt):
        """(sum_j e_j, sum_j (re^2, im^2)) over ALL trajectories of the group, [n_e][n_t] each:
        every engine (``None`` for a member that ran nothing) reduces the expectation values
        its last run left on its device, then ONE ncclAllReduce."""
        if len(engines) != self.nlocal:
            raise ValueError("one engine per local member")
        hs = (C.c_void_p * self.nlocal)(*[None if e is None else e.handle for e in engines])
        out = np.zeros((2, neops, nt), dtype=np.complex128)
        check(_lib.load().qb_comm_reduce_expect(self.handle, hs, int(neops), int(nt), ptr(out)))
        return out[0], out[1]
