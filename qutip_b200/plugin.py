"""Registration with QuTiP's own extension surfaces (drop-in boundary).

``import qutip_b200.plugin`` (with QuTiP importable) registers

* data-layer types ``B200Dense`` (device state) and ``B200Operator`` (device sparse
  operator) with ``qutip.core.data.to.add_conversions`` (core/data/convert.pyx:208-329) and
  specialisations with ``Dispatcher.add_specialisations`` (core/data/dispatch.pyx:267-337)
  for the operations the stock RK loop and solvers call on a state: matmul, add, mul,
  imul, neg, zeros-like copies, norm.l2 / norm.frobenius, ode.wrmn_error, expect,
  expect_super, trace_oper_ket, inner;
* integrators ``"b200_vern7"`` / ``"b200_vern9"`` / ``"b200_tsit5"`` with
  ``MESolver.add_integrator``, ``SESolver.add_integrator`` and ``MCSolver.add_integrator``
  (solver/solver_base.py:475-492): the whole adaptive RK stepping of
  ``Explicit_RungeKutta`` runs fused on the device, same ``integrator_options`` keys and
  defaults as ``IntegratorVern7`` (solver/integrator/qutip_integrator.py:51-59);
  ``"b200_adams"`` (Nordsieck Adams-Moulton on the device, option keys of the stock
  ``adams``) and ``"b200_zvode"`` (SciPy zvode with the right-hand side on the device);
  ``options["matrix_form"]`` is bound matrix-free (Kronecker / sandwich operators);
* the map ``options["map"] = "b200"`` in ``qutip.solver.parallel._maps``
  (solver/parallel.py:541-559): ``MCSolver.run`` hands all seeds to the device engine,
  which runs every trajectory (RK steps, jump detection, collapse selection) without
  host round trips and feeds per-trajectory ``Result`` objects to ``McResult.add``
  (pure and mixed initial states, improved sampling).

Anything that cannot run on the device (python-function coefficients or operators,
feedback arguments) raises ``TypeError`` naming the stock method to use instead -- there is no
CPU fallback in this package.  Python-callable ``e_ops`` in the batched map are evaluated by
the host on the trajectories' downloaded states (they are host code by definition).
"""
import os
import threading
import warnings

import numpy as np
import scipy.sparse as sp

import qutip
from qutip.core import data as _data
from qutip.core.cy.qobjevo import QobjEvo
from qutip.core.cy.coefficient import Coefficient as _Coefficient
from qutip.solver.integrator.integrator import Integrator, IntegratorException
from qutip.solver.mcsolve import MCSolver
from qutip.solver.mesolve import MESolver
from qutip.solver.sesolve import SESolver
from qutip.solver import parallel as _qparallel
import qutip.solver.integrator.scipy_integrator  # noqa: F401

from . import _lib, coeffs, engine as E, solve
from .solve import make_thresholds  # noqa: F401

__all__ = ["B200Dense", "B200Operator", "B200Vern7", "B200Vern9", "B200Tsit5", "B200Adams", "bind_qobjevo", "b200_map",
           "register", "configure"]


# ------------------------------------------------------------------ QobjEvo -> device system
def _data_to_host(d):
    """scipy / numpy view of a reference data-layer object."""
    if isinstance(d, _data.CSR):
        return d.as_scipy()
    if isinstance(d, _data.Dia):
        return d.as_scipy()
    if isinstance(d, _data.Dense):
        return np.asfortranarray(d.to_array())
    if isinstance(d, B200Operator):
        return d
    return _data.to(_data.CSR, d).as_scipy()


def _coeff_state(c):
    st = c.__reduce__()
    if len(st) > 2 and st[2] is not None:
        return st[2]
    return st[1][2]


def coefficient_to_program(c, system=None):
    """Compile a reference Coefficient (core/cy/coefficient.pyx) into device byte-code."""
    name = type(c).__name__
    if name == "ConstantCoefficient":
        return coeffs.constant(_coeff_state(c)[1])
    if name.startswith("StrCoefficient"):
        state = _coeff_state(c)
        # (named values..., extracted numeric constants..., args, code, names...): the code
        # string is the first str; the names follow it and own the leading values
        i_code = next(i for i, x in enumerate(state) if isinstance(x, str))
        code, names = state[i_code], state[i_code + 1:]
        vals = state[:len(names)]
        prog = coeffs.compile_expr(code, dict(zip(names, vals)))
        # the pickled members are ordered by attribute name (typed args: _arg_cpl*, _arg_dbl*,
        # _arg_int*), which need not be the order of the names: check the compiled program
        # against the coefficient itself before trusting the binding
        for t in (0.0, 0.37, 1.9):
            try:
                want, got = complex(c(t)), complex(coeffs.evaluate(prog, t))
            except (NotImplementedError, ArithmeticError, ValueError):
                continue                      # instruction without a python mirror / singular point
            if not (np.isfinite(want.real) and np.isfinite(want.imag)):
                continue
            if not abs(want - got) <= 1e-12 * max(1.0, abs(want)):
                raise TypeError("string coefficient %r could not be bound to its argument "
                                "values (typed arguments); evaluated by the host instead" % code)
        return prog
    if name == "ConjCoefficient":
        return coefficient_to_program(_coeff_state(c)[1], system).conj()
    if name == "NormCoefficient":
        return coefficient_to_program(_coeff_state(c)[1], system).norm()
    if name == "SumCoefficient":
        s = _coeff_state(c)
        return coefficient_to_program(s[1], system) + coefficient_to_program(s[2], system)
    if name == "MulCoefficient":
        s = _coeff_state(c)
        return coefficient_to_program(s[1], system) * coefficient_to_program(s[2], system)
    if name == "RateShiftCoefficient":          # nm_mcsolve (solver/cy/nm_mcsolve.pyx:15-78)
        return coeffs.rate_shift([coefficient_to_program(x, system) for x in c.__reduce__()[1][0]])
    if name == "SqrtRealCoefficient":           # nm_mcsolve (solver/cy/nm_mcsolve.pyx:80-119)
        return coefficient_to_program(_coeff_state(c)[1], system).sqrt_real()
    if name == "InterCoefficient":
        if system is None:
            raise TypeError("array coefficients need a system to hold the spline table")
        tl, poly, dt = c.__reduce__()[1]
        return coeffs.spline(system.add_spline(tl, poly, dt or 0.0))
    raise TypeError(
        "coefficient of type %s cannot be evaluated on the device (python callables have no "
        "device form); use a string or array coefficient, or a stock method such as 'vern7'"
        % name)


class _HostCoefficient(coeffs.Program):
    """Program placeholder for a coefficient only the host can evaluate (python callable);
    ``coeff(t)`` is the reference Coefficient itself."""

    def __init__(self, coeff):
        super().__init__(coeffs.host().instrs)
        self.coeff = coeff


def bind_qobjevo(qevo, system=None, allow_host=False):
    """[(host operator, Program|None)] for ``sum_k coeff_k(t) A_k`` with all constant
    elements merged into one operator placed last (QobjEvo.compress order,
    core/cy/qobjevo.pyx:816-867).  Raises TypeError for forms that only exist on the host."""
    if isinstance(qevo, qutip.Qobj):
        qevo = QobjEvo(qevo)
    if getattr(qevo, "_feedback_functions", None):
        raise TypeError("feedback arguments rebuild the operator from the state on the host at "
                        "every RHS evaluation and cannot run on the device; use method='vern7'")
    const = None
    out = []
    for el in qevo.to_list():
        if isinstance(el, qutip.Qobj):
            h = _data_to_host(el.data)
            h = sp.csr_matrix(h)
            const = h if const is None else const + h
        elif isinstance(el, (list, tuple)) and isinstance(el[0], qutip.Qobj) \
                and isinstance(el[1], _Coefficient):
            try:
                prog = coefficient_to_program(el[1], system)
            except (TypeError, ValueError) as exc:
                if not allow_host:
                    if isinstance(exc, ValueError):
                        raise TypeError(str(exc)) from exc
                    raise
                prog = _HostCoefficient(el[1])        # python callable: evaluated by the host
            out.append((_data_to_host(el[0].data), prog))
        else:
            raise TypeError("QobjEvo contains a function / map / product element, which is a "
                            "python callable returning a Qobj and cannot run on the device; "
                            "use method='vern7'")
    if const is not None:
        const = sp.csr_matrix(const)
        const.sum_duplicates()
        const.sort_indices()
        out.append((const, None))
    if not out:
        n = qevo.shape[0]
        out.append((sp.csr_matrix((n, n), dtype=complex), None))
    return out


def _single_element(qevo, what):
    els = bind_qobjevo(qevo)
    if len(els) != 1:
        raise TypeError("%s with several time-dependent terms is not supported on the device"
                        % what)
    return els[0]


def _upload(h):
    if isinstance(h, (E.DeviceOp, E.DeviceDense)):
        return h
    if isinstance(h, B200Operator):
        return h.dev
    if sp.issparse(h):
        return E.DeviceOp.from_scipy(h)
    return E.DeviceDense.from_numpy(np.asfortranarray(h))


def system_from_qobjevo(qevo, c_ops=(), n_ops=(), e_ops=(), functional=False, allow_host=False,
                        ncols=1):
    """``ncols`` > 1: the state is an N x ncols matrix (propagator-style evolution,
    solver/propagator.py:401-417); its columns are stacked and every operator becomes
    block-diagonal kron(I_ncols, A), so one error norm covers the whole matrix exactly as
    in the reference's RK loop."""
    N = qevo.shape[0] * ncols
    system = E.System(N)
    system.coeff_objects = []       # per element: host-evaluated reference Coefficient or None
    system.programs = []            # per element: compiled Program or None (constant 1)
    for h, prog in bind_qobjevo(qevo, system, allow_host):
        if ncols > 1:
            h = sp.kron(sp.identity(ncols, dtype=complex, format="csr"), sp.csr_matrix(h),
                        format="csr")
        system.add_element(_upload(h), prog)
        system.coeff_objects.append(getattr(prog, "coeff", None))
        system.programs.append(prog)
    system.has_host = any(c is not None for c in system.coeff_objects)
    for c, n in zip(c_ops, n_ops):
        ch, cp = _single_element(c, "a collapse operator")
        nh, npg = _single_element(n, "a collapse operator")
        system.add_collapse(_upload(ch), _upload(nh), cp, npg)
    for e in e_ops:
        eh, ep = _single_element(e, "an e_op")
        system.add_eop(_upload(eh), ep)
    if functional:
        system.set_functional(True)
    return system


def _device_qevo(qevo):
    """The QobjEvo the device integrates.  ``options["matrix_form"]`` hands the integrator a
    LindbladMatrixForm (core/cy/lindblad_matrix_form.pyx:27-203), whose RHS
    ``-i(H_nh rho - rho H_nh^dag) + sum c rho c^dag`` acts on the un-vectorised n x n rho.
    On the device the same linear map is applied to the column-stacked rho as the
    superoperator ``-i(spre(H_nh) - spost(H_nh^dag)) + sum sprepost(c, c^dag)``: identical
    result, one fused SpMV per stage instead of 2 + 2*len(c_ops) sparse-dense products."""
    if type(qevo).__name__ != "LindbladMatrixForm":
        return qevo
    H_nh = qevo.H_nh
    sup = -1j * (qutip.spre(H_nh) - qutip.spost(H_nh.dag()))
    for c in qevo.c_ops:
        sup = sup + qutip.sprepost(c, c.dag())
    return QobjEvo(sup)


# n (operator dimension) from which LindbladMatrixForm is bound matrix-free instead of as
# the fused n^2 x n^2 superoperator; QUTIP_B200_MATRIX_FREE=0/1 forces either
MATRIX_FREE_MIN_DIM = 256


def matrix_free_system(lmf, ncols=1):
    """Device system for a LindbladMatrixForm without materialising any n^2 x n^2 Hamiltonian
    part (core/cy/lindblad_matrix_form.pyx:105-203): for every term ``f(t) A`` of ``H_nh`` the
    matrix-free Kronecker operators ``-i f (I (x) A)`` and ``+i conj(f) (conj(A) (x) I)``
    (rho -> A rho, rho -> rho A^dagger), plus the jump part ``sum |g|^2 C rho C^dagger`` -- an
    explicit sparse superoperator ``sum conj(C) (x) C`` while nnz = sum nnz(C)^2 is small, the
    matrix-free sandwich operator beyond (solve._jump_operator).
    Raises TypeError when a collapse operator has several terms or a python coefficient."""
    if ncols != 1:
        raise TypeError("matrix-valued states are not combined with matrix_form")
    n = lmf.shape[0]
    system = E.System(n * n)
    system.coeff_objects, system.programs = [], []

    def add(op, prog):
        system.add_element(op, prog)
        system.coeff_objects.append(None)
        system.programs.append(prog)

    for h, prog in bind_qobjevo(lmf.H_nh, system, allow_host=False):
        h = sp.csr_matrix(h)
        if prog is None:
            pl, pr = coeffs.constant(-1j), coeffs.constant(1j)
        else:
            pl, pr = prog.scaled(-1j), prog.conj().scaled(1j)
        add(E.DeviceOp.kron(h, 0), pl)
        add(E.DeviceOp.kron(h, 1), pr)
    from .solve import _jump_operator
    jump_const = []
    for c in lmf.c_ops:
        ch, cp = _single_element(c, "a collapse operator (matrix_form)")
        ch = sp.csr_matrix(ch)
        if cp is None:
            jump_const.append(ch)
        else:
            add(_jump_operator([ch], "auto"), cp.norm())
    if jump_const:
        add(_jump_operator(jump_const, "auto"), None)
    system.has_host = False
    return system


def _bind_system(qevo, ncols):
    """(device system, base dimension) for the integrators."""
    if type(qevo).__name__ == "LindbladMatrixForm" and ncols == 1:
        import os
        force = os.environ.get("QUTIP_B200_MATRIX_FREE")
        want = (qevo.shape[0] >= MATRIX_FREE_MIN_DIM) if force is None else force == "1"
        if want:
            try:
                return matrix_free_system(qevo), qevo.shape[0] ** 2
            except TypeError:
                pass            # python coefficients / multi-term c_ops: fused superoperator
    dev_qevo = _device_qevo(qevo)
    return system_from_qobjevo(dev_qevo, allow_host=True, ncols=ncols), dev_qevo.shape[0]


def _state_columns(arr_shape, base_n):
    """1 when the state is the (possibly n x n, to be stacked) vector the system acts on;
    k for an N x k matrix-valued state evolved column by column."""
    if arr_shape[0] * arr_shape[1] == base_n:
        return 1
    if arr_shape[0] == base_n:
        return arr_shape[1]
    raise TypeError("incompatible dimensions %s for a system of size %d" % (arr_shape, base_n))


# ------------------------------------------------------------------ integrators
class _B200Integrator(Integrator):
    """Device-resident adaptive Runge-Kutta (restates Explicit_RungeKutta,
    solver/integrator/explicit_rk.pyx, on the GPU).  ``rhs_format`` is "callable" like the
    stock Verner integrators: the derivative must be the bound ``QobjEvo.matmul_data`` of
    the system (what Solver._get_integrator passes, solver_base.py:292-295)."""
    integrator_options = {
        'atol': 1e-8,
        'rtol': 1e-6,
        'nsteps': 1000,
        'first_step': 0,
        'max_step': 0,
        'min_step': 0,
        'interpolate': True,
    }
    rhs_format = "callable"
    _tableau = "vern7"
    method = "b200_vern7"

    def _prepare(self):
        qevo = getattr(self.derivative, "__self__", None)
        if not isinstance(qevo, QobjEvo) or getattr(self.derivative, "__name__", "") != "matmul_data":
            raise TypeError(
                "%s integrates QobjEvo systems on the device; the derivative must be "
                "`QobjEvo.matmul_data` (python functions cannot run on the GPU). Use "
                "method='%s' for arbitrary callables." % (self.method, self._tableau))
        self._qevo = qevo
        self._build()
        self.name = self.method

    _ncols = 1

    def _build(self):
        o = self._options
        self._system, self._base_n = _bind_system(self._qevo, self._ncols)
        self._engine = E.Engine(
            self._system, self._tableau, nslots=1, atol=o['atol'], rtol=o['rtol'],
            nsteps=int(o['nsteps']), first_step=float(o['first_step'] or 0),
            min_step=float(o['min_step'] or 0), max_step=float(o['max_step'] or 0),
            interpolate=int(bool(o.get('interpolate', True))), max_order=int(o.get('order', 0) or 0),
            store_states=1)          # only used by run(tlist); the per-call protocol ignores it
        if self._system.has_host:
            objs = self._system.coeff_objects
            self._engine.host_coeffs = lambda t: [1.0 if c is None else complex(c(t)) for c in objs]
        self._shape = None

    def arguments(self, args):
        # QobjEvo.arguments() was already applied by the solver; re-bind the coefficients
        if self._is_set:
            state = self.get_state()
        self._build()
        if self._is_set:
            self.set_state(*state)

    def reset(self, hard=False):
        # Solver._argument() updates the QobjEvo's args and then calls reset()
        # (solver_base.py:456-460): the coefficient programs hold the bound args, so re-bind
        self.arguments(None)

    def set_state(self, t, state):
        if getattr(self._qevo, "_feedback_functions", None):
            # feedback registered on the QobjEvo after this integrator was built (solvers do it
            # at the start of a run, solver_base.py / mcsolve.py _register_feedback)
            raise TypeError("feedback arguments rebuild the operator from the state on the host at "
                            "every RHS evaluation and cannot run on the device; use method='vern7'")
        arr = _data.to(_data.Dense, state).to_array()
        ncols = _state_columns(arr.shape, self._base_n)
        if ncols != self._ncols:                 # matrix-valued state: re-bind block-diagonal
            self._ncols = ncols
            self._build()
        self._shape = arr.shape
        self._y_set = np.ascontiguousarray(arr.reshape(-1, order="F"))
        self._engine.set_state(t, self._y_set)
        self._is_set = True
        self._fresh_at = t           # no step taken since: run(tlist) may take the batched path

    def _wrap(self, y):
        return _data.Dense(y.reshape(self._shape, order="F"), copy=False)

    def get_state(self, copy=True):
        t, y = self._engine.get_state()
        return t, self._wrap(y)

    def run(self, tlist):
        """``Integrator.run`` (solver/integrator/integrator.py:197-212; consumed by
        ``Solver.run``, solver_base.py:218-220).  Right after ``set_state(tlist[0], ...)`` the
        whole list is integrated by ONE device run -- the same sequence of ``integrate(t)``
        calls the base class makes, without a host round trip per output time; otherwise
        (python-evaluated coefficients, an integrator that has already stepped) the calls are
        made one by one."""
        tlist = np.asarray(tlist, dtype=float)
        if (getattr(self, "_fresh_at", None) is None or len(tlist) < 3 or tlist[0] != self._fresh_at
                or self._system.has_host or os.environ.get("QUTIP_B200_NO_BATCHED_RUN")):
            for t in tlist[1:]:
                yield self.integrate(t, False)
            return
        self._fresh_at = None
        r = self._engine.run_mesolve(self._y_set, tlist)
        st = int(r.status[0])
        # states of the output times reached before a failure are handed over first, as the
        # one-by-one loop would have done
        good = len(tlist) if st == 1 else 0
        for i in range(1, good):
            yield float(tlist[i]), self._wrap(r.states[0, i].copy())
        if st != 1:
            # reproduce the failure through the protocol so that the exception (and the states
            # yielded before it) are those of the reference's loop
            self._engine.set_state(tlist[0], self._y_set)
            for t in tlist[1:]:
                yield self.integrate(t, False)

    def _run(self, t, step):
        self._fresh_at = None
        t_out, status = self._engine.integrate(t, step)
        if status < 0:
            raise IntegratorException(E.STATUS_MESSAGES.get(status, "integration failed"))
        return self.get_state(False)

    def integrate(self, t, copy=True):
        return self._run(t, False)

    def mcstep(self, t, copy=True):
        return self._run(t, True)

    def stats(self):
        return self._engine.stats()

    def __getstate__(self):
        d = dict(self.__dict__)
        saved = self.get_state() if self._is_set else None
        for k in ("_system", "_engine"):
            d.pop(k, None)
        d["_saved_state"] = saved
        return d

    def __setstate__(self, d):
        saved = d.pop("_saved_state", None)
        self.__dict__.update(d)
        self._build()
        if saved is not None:
            self.set_state(*saved)

    @property
    def options(self):
        """
        Supported options by the device Verner methods (same keys and defaults as the
        stock ``vern7`` / ``vern9``):

        atol : float, default: 1e-8
            Absolute tolerance.

        rtol : float, default: 1e-6
            Relative tolerance.

        nsteps : int, default: 1000
            Max. number of internal steps/call.

        first_step : float, default: 0
            Size of initial step (0 = automatic).

        min_step : float, default: 0
            Minimum step size (0 = automatic).

        max_step : float, default: 0
            Maximum step size (0 = automatic)

        interpolate : bool, default: True
            Whether to use interpolation step, faster most of the time.
        """
        return self._options

    @options.setter
    def options(self, new_options):
        Integrator.options.fset(self, new_options)


class B200Vern7(_B200Integrator):
    """Verner 7(6) "most efficient" pair, fused on the B200.  ``method="b200_vern7"``."""
    _tableau = "vern7"
    method = "b200_vern7"


class B200Vern9(_B200Integrator):
    """Verner 9(8) "most efficient" pair, fused on the B200.  ``method="b200_vern9"``."""
    _tableau = "vern9"
    method = "b200_vern9"


class B200Tsit5(_B200Integrator):
    """Tsitouras 5(4) pair (first-same-as-last), fused on the B200.  ``method="b200_tsit5"``."""
    _tableau = "tsit5"
    method = "b200_tsit5"


class B200Adams(_B200Integrator):
    """``method="b200_adams"``: variable-order (1..12), variable-step Adams-Moulton method in
    Nordsieck form running entirely on the device (csrc/qb_adams.h) -- the device-resident
    counterpart of the reference's ``method="adams"`` (IntegratorScipyAdams,
    solver/integrator/scipy_integrator.py:20-196: SciPy zvode, Adams, order 12).  Same option
    keys and defaults as the reference integrator; results agree with it within the
    integration tolerance (the step sequence is not zvode's -- use ``b200_zvode`` for that)."""
    integrator_options = {
        'atol': 1e-8,
        'rtol': 1e-6,
        'nsteps': 2500,
        'order': 12,
        'first_step': 0,
        'max_step': 0,
        'min_step': 0,
    }
    _tableau = "adams"
    method = "b200_adams"

    @property
    def options(self):
        """
        Supported options by the device Adams method (keys and defaults of the stock
        ``adams``, scipy_integrator.py:22-30):

        atol : float, default: 1e-8
            Absolute tolerance.

        rtol : float, default: 1e-6
            Relative tolerance.

        order : int, default: 12
            Highest order used (<= 12).

        nsteps : int, default: 2500
            Max. number of internal steps/call.

        first_step : float, default: 0
            Size of initial step (0 = automatic).

        min_step : float, default: 0
            Minimum step size (0 = automatic).

        max_step : float, default: 0
            Maximum step size (0 = automatic)
        """
        return self._options

    @options.setter
    def options(self, new_options):
        Integrator.options.fset(self, new_options)


class B200Zvode(qutip.solver.integrator.scipy_integrator.IntegratorScipyAdams):
    """``method="b200_zvode"``: the reference's Adams integrator (SciPy zvode, variable-order
    Adams-Moulton; solver/integrator/scipy_integrator.py:20-196) with the RHS callback
    ``_mul_np_vec`` (:62-71) evaluated on the device: every ``QobjEvo.matmul_data`` call
    becomes one fused ``qb_engine_rhs`` launch.  Step control and order selection stay in
    SciPy's compiled zvode, exactly as in the reference, so the step sequence is the
    reference's; the state crosses PCIe once per RHS evaluation (``b200_adams`` is the
    device-resident Adams method)."""
    method = "adams"

    _ncols = 1

    def _bind(self):
        self._system, self._base_n = _bind_system(self._qevo, self._ncols)
        self._engine = E.Engine(self._system, "vern7", nslots=1)
        n = self._system.N
        self._dx = E.DeviceDense.zeros(n, 1)
        self._dout = E.DeviceDense.zeros(n, 1)
        self._hout = np.empty(n, dtype=np.complex128)
        self._progs = self._system.programs

    def _prepare(self):
        qevo = getattr(self.derivative, "__self__", None)
        if not isinstance(qevo, QobjEvo) or getattr(self.derivative, "__name__", "") != "matmul_data":
            raise TypeError("b200_zvode integrates QobjEvo systems on the device; use "
                            "method='adams' for arbitrary callables")
        self._qevo = qevo
        self._bind()
        super()._prepare()
        self.name = "b200 device RHS + scipy zvode adams"

    @staticmethod
    def _mul_np_vec(t, vec, self):
        self._dx.write(vec)
        if self._system.has_host:
            vals = [1.0 if c is None else complex(c(t)) for c in self._system.coeff_objects]
            # device-compiled elements are evaluated by the same byte-code on the host side
            for k, c in enumerate(self._system.coeff_objects):
                if c is None and self._progs[k] is not None:
                    vals[k] = coeffs.evaluate(self._progs[k], t)
            self._engine.rhs_coef(vals, self._dx, self._dout)
        else:
            self._engine.rhs(t, self._dx, self._dout)
        return self._dout.read_into(self._hout)

    def set_state(self, t, state0):
        ncols = _state_columns(state0.shape, self._base_n)
        if ncols != self._ncols:                 # matrix-valued state: block-diagonal binding
            self._ncols = ncols
            self._bind()
        super().set_state(t, _data.to(_data.Dense, state0))

    def arguments(self, args):
        self._bind()

    def reset(self, hard=False):
        self.arguments(None)
        super().reset(hard)

    def __getstate__(self):
        raise TypeError("b200_zvode integrators hold device handles and SciPy zvode state; "
                        "re-create them instead of pickling")


# ------------------------------------------------------------------ data-layer types
class B200Dense(_data.Data):
    """Dense complex128 matrix resident in HBM (mirror of core/data/dense.pxd:9-22)."""

    def __init__(self, dev, shape=None):
        if not isinstance(dev, E.DeviceDense):
            dev = E.DeviceDense.from_numpy(np.asarray(dev, dtype=complex))
        super().__init__(tuple(dev.shape))
        self.dev = dev

    @classmethod
    def sparcity(cls):
        return "dense"

    def to_array(self):
        return self.dev.to_numpy()

    def copy(self):
        return B200Dense(self.dev.copy())

    def conj(self):
        return B200Dense(np.conj(self.to_array()))

    def transpose(self):
        return B200Dense(self.to_array().T)

    def adjoint(self):
        return B200Dense(np.conj(self.to_array().T))

    def trace(self):
        return complex(np.trace(self.to_array()))


class B200Operator(_data.Data):
    """Sparse operator resident in HBM (diagonal-masked slices or CSR).  Keeps the host
    object it was converted from so that conversions back are exact."""

    def __init__(self, host):
        if not isinstance(host, (_data.CSR, _data.Dia)):
            host = _data.to(_data.CSR, host)
        super().__init__(tuple(host.shape))
        self.host = host
        self.dev = E.DeviceOp.from_scipy(host.as_scipy())

    @classmethod
    def sparcity(cls):
        return "sparse"

    def to_array(self):
        return self.host.to_array()

    def copy(self):
        return B200Operator(self.host.copy())

    def conj(self):
        return B200Operator(self.host.conj())

    def transpose(self):
        return B200Operator(self.host.transpose())

    def adjoint(self):
        return B200Operator(self.host.adjoint())

    def trace(self):
        return self.host.trace()


def _dense_from(d):
    arr = d.to_array()
    return B200Dense(E.DeviceDense.from_numpy(arr))


def _to_dense(d):
    return _data.Dense(d.to_array(), copy=False)


def _check_mm(left, right):
    if left.shape[1] != right.shape[0]:
        raise ValueError("incompatible matrix shapes " + str(left.shape) + " and "
                         + str(right.shape))


def _matmul_op_dense(left, right, scale=1):
    _check_mm(left, right)
    return B200Dense(E.matmul(left.dev, right.dev, scale))


def _matmul_dense_dense(left, right, scale=1):
    _check_mm(left, right)
    if not left.dev.fortran and min(left.shape) > 1:
        left = B200Dense(np.asfortranarray(left.to_array()))
    return B200Dense(E.matmul(left.dev, right.dev, scale))


def _matmul_dag_dense_op(left, right, scale=1):
    """scale * left @ right^dagger (matmul_dag, core/data/matmul.pyx:1084-1116; the product
    LindbladMatrixForm and _BaseElement.adjoint_rmatmul_data_t are made of, _element.pyx:225-237).
    For the square column-major rho and an n x n operator this is the matrix-free right product
    (conj(A) (x) I) vec(rho) on the device (KRON side 1); other shapes go through
    (A left^dagger)^dagger."""
    if left.shape[1] != right.shape[1]:
        raise ValueError("incompatible matrix shapes " + str(left.shape) + " and "
                         + str(right.shape[::-1]))
    n = right.shape[0]
    if left.shape == (n, n) and right.shape == (n, n) and (left.dev.fortran or n == 1):
        k = getattr(right, "_kron_right", None)
        if k is None:
            k = right._kron_right = E.DeviceOp.kron(sp.csr_matrix(right.host.as_scipy()), 1)
        vec = left.dev.copy().reshape(n * n, 1)
        return B200Dense(E.matmul(k, vec, scale).reshape(n, n))
    out = E.matmul(right.dev, left.adjoint().dev, np.conj(complex(scale)))     # A left^dag conj(s)
    return B200Dense(out).adjoint()


def _matmul_dag_dense_dense(left, right, scale=1):
    if left.shape[1] != right.shape[1]:
        raise ValueError("incompatible matrix shapes " + str(left.shape) + " and "
                         + str(right.shape[::-1]))
    return _matmul_dense_dense(left, right.adjoint(), scale)


def _add_dense(left, right, scale=1):
    if left.shape != right.shape:
        raise ValueError("incompatible matrix shapes " + str(left.shape) + " and "
                         + str(right.shape))
    out = left.dev.copy()
    E.axpy(right.dev, scale, out)
    return B200Dense(out)


def _iadd_dense(left, right, scale=1):
    E.axpy(right.dev, scale, left.dev)
    return left


def _sub_dense(left, right):
    return _add_dense(left, right, -1)


def _mul_dense(matrix, value):
    out = matrix.dev.copy()
    E.scal(out, value)
    return B200Dense(out)


def _imul_dense(matrix, value):
    E.scal(matrix.dev, value)
    return matrix


def _neg_dense(matrix):
    return _mul_dense(matrix, -1)


def _zeros_dense(rows, cols):
    return B200Dense(E.DeviceDense.zeros(rows, cols, True))


def _l2_dense(vector):
    return E.nrm2(vector.dev)


def _wrmn_dense(diff, state, atol, rtol):
    return E.wrms_error(diff.dev, state.dev, atol, rtol)


def _expect_op_dense(op, state):
    if state.shape[1] == 1:
        return E.expect_ket(op.dev, state.dev)
    return E.expect_dm(op.dev, state.dev)


def _expect_super_op_dense(op, state):
    return E.expect_super(op.dev, state.dev)


def _trace_oper_ket_dense(matrix):
    return E.trace_oper_ket(matrix.dev)


def _inner_dense(left, right, scalar_is_ket=False):
    if left.shape[1] == 1 and left.shape[0] != 1 or (left.shape == (1, 1) and scalar_is_ket):
        return E.inner(left.dev, right.dev, True)       # <left|right>, left a ket
    return E.inner(left.dev, right.dev, False)          # left is a bra


_registered = False


def register():
    """Idempotent registration of the types, specialisations, integrators and the map."""
    global _registered
    if _registered:
        return
    _data.to.add_conversions([
        (B200Dense, _data.Dense, _dense_from, 1),
        (_data.Dense, B200Dense, _to_dense, 1),
        (B200Operator, _data.CSR, B200Operator, 1),
        (B200Operator, _data.Dia, B200Operator, 1),
        (_data.CSR, B200Operator, lambda d: _data.to(_data.CSR, d.host), 1),
    ])
    _data.to.register_aliases(["b200", "B200Dense", "b200_dense"], B200Dense)
    _data.to.register_aliases(["B200Operator", "b200_operator", "b200_sparse"], B200Operator)
    _data.matmul.add_specialisations([
        (B200Operator, B200Dense, B200Dense, _matmul_op_dense),
        (B200Dense, B200Dense, B200Dense, _matmul_dense_dense),
    ])
    _data.matmul_dag.add_specialisations([
        (B200Dense, B200Operator, B200Dense, _matmul_dag_dense_op),
        (B200Dense, B200Dense, B200Dense, _matmul_dag_dense_dense),
    ])
    _data.add.add_specialisations([(B200Dense, B200Dense, B200Dense, _add_dense)])
    _data.sub.add_specialisations([(B200Dense, B200Dense, B200Dense, _sub_dense)])
    _data.mul.add_specialisations([(B200Dense, B200Dense, _mul_dense)])
    _data.imul.add_specialisations([(B200Dense, B200Dense, _imul_dense)])
    _data.neg.add_specialisations([(B200Dense, B200Dense, _neg_dense)])
    if hasattr(_data, "iadd"):
        _data.iadd.add_specialisations([(B200Dense, B200Dense, B200Dense, _iadd_dense)])
    _data.zeros.add_specialisations([(B200Dense, _zeros_dense)])
    _data.norm.l2.add_specialisations([(B200Dense, _l2_dense)])
    _data.norm.frobenius.add_specialisations([(B200Dense, _l2_dense)])
    _data.ode.wrmn_error.add_specialisations([(B200Dense, B200Dense, _wrmn_dense)])
    _data.expect.add_specialisations([(B200Operator, B200Dense, _expect_op_dense)])
    _data.expect_super.add_specialisations([(B200Operator, B200Dense, _expect_super_op_dense)])
    _data.trace_oper_ket.add_specialisations([(B200Dense, _trace_oper_ket_dense)])
    _data.inner.add_specialisations([(B200Dense, B200Dense, _inner_dense)])

    # also the other Solver subclasses whose right-hand side is a QobjEvo: HEOMSolver's hierarchy
    # generator, BRSolver's constant Bloch-Redfield tensor, FMESolver.  (Not the Solver base class:
    # its integrators must accept arbitrary python callables, tests/solver/test_integrator.py.)
    solvers = [MESolver, SESolver, MCSolver]
    try:
        from qutip.solver.brmesolve import BRSolver
        from qutip.solver.floquet import FMESolver
        from qutip.solver.heom.bofin_solvers import HEOMSolver
        solvers += [BRSolver, FMESolver, HEOMSolver]
    except ImportError:                                   # pragma: no cover
        pass
    for solver in solvers:
        solver.add_integrator(B200Vern7, "b200_vern7")
        solver.add_integrator(B200Vern9, "b200_vern9")
        solver.add_integrator(B200Tsit5, "b200_tsit5")
        solver.add_integrator(B200Adams, "b200_adams")
        solver.add_integrator(B200Zvode, "b200_zvode")
    _qparallel._maps["b200"] = b200_map
    import os
    if os.environ.get("QUTIP_B200_DEFAULT_MAP") == "1":
        # acceptance runs of the reference's own mcsolve tests: every MCSolver / nm_mcsolve
        # call that does not name a map takes the device map
        from qutip.solver.nm_mcsolve import NonMarkovianMCSolver
        for cls in (MCSolver, NonMarkovianMCSolver):
            cls.solver_options = dict(cls.solver_options, map="b200")
    default_method = os.environ.get("QUTIP_B200_DEFAULT_METHOD")
    if default_method:
        # acceptance runs of the reference's solver tests with a device integrator as THE default
        for cls in (MESolver, SESolver, MCSolver):
            cls.solver_options = dict(cls.solver_options, method=default_method)
    _registered = True


# ------------------------------------------------------------------ whole-batch mcsolve
def _device_method(method):
    if method in ("b200_vern7", "vern7"):
        return "vern7"
    if method in ("b200_vern9", "vern9"):
        return "vern9"
    if method in ("b200_tsit5", "tsit5"):
        return "tsit5"
    if method in ("b200_adams", "adams"):
        return "adams"
    raise TypeError("the b200 map runs vern7 / vern9 / tsit5 / adams on the device, not method=%r"
                    % (method,))


def b200_map(task, values, task_args=None, task_kwargs=None, reduce_func=None, map_kw=None,
             progress_bar=None, progress_bar_kwargs={}):
    """Map function for ``MCSolver.run(..., options={"map": "b200"})`` (protocol of
    solver/parallel.py:49-132).  Trajectory tasks -- ``MCSolver._run_one_traj`` (one pure
    initial state, ``values`` = seeds, ``task_args = (state0, tlist, e_ops)``) and
    ``_run_one_traj_mixed`` (mixed initial states, ``values`` = trajectory ids or
    ``(id, jump_prob_floor)`` pairs, multitraj.py:285-352, mcsolve.py:752-792) -- run as device
    batches, one per initial state; each trajectory is handed to ``reduce_func`` as the
    ``(seed, Result, weight)`` triple ``McResult.add`` expects
    (solver/multitrajresult.py:402-434).  Any other task (the no-jump simulations of
    improved sampling) is executed in the calling process like ``serial_map``."""
    task_args = tuple(task_args or ())
    task_kwargs = dict(task_kwargs or {})
    inner = getattr(task, "func", task)                   # mcsolve._unpack_arguments wrapper
    solver = getattr(inner, "__self__", None)
    name = getattr(inner, "__name__", "")
    # NonMarkovianMCSolver (solver/nm_mcsolve.py): the same trajectories with rate-shifted collapse
    # operators; the influence martingale of a trajectory depends only on its collapse record
    # and is attached to the results afterwards (_b200_batch), for pure and mixed initial states
    # (_run_one_traj_mixed only picks the state and multiplies the weight, mcsolve.py:752-792).
    # Other subclasses override the trajectory function and run through their own code, one
    # trajectory at a time, with a RuntimeWarning.
    from qutip.solver.nm_mcsolve import NonMarkovianMCSolver
    batched = type(solver) in (MCSolver, NonMarkovianMCSolver)
    if not batched or name not in ("_run_one_traj", "_run_one_traj_mixed"):
        if name.startswith("_run_one_traj"):
            warnings.warn("the 'b200' map batches MCSolver / NonMarkovianMCSolver trajectories on the device; "
                          "%s.%s is run one trajectory at a time through the solver's own code"
                          % (type(solver).__name__, name), RuntimeWarning, stacklevel=2)
        results = []
        for v in values:
            out = task(v, *task_args, **task_kwargs)
            if reduce_func is not None:
                remaining = reduce_func(out)
                if remaining is not None and remaining <= 0:
                    break
            else:
                results.append(out)
        return None if reduce_func is not None else results
    if bool(task_kwargs.pop("no_jump", False)):
        raise TypeError("the 'b200' map does not run forced no-jump trajectories")
    if name == "_run_one_traj":
        floor = float(task_kwargs.pop("jump_prob_floor", 0.0))      # improved sampling
        if task_kwargs:
            raise TypeError("unsupported trajectory arguments %s for the 'b200' map" % list(task_kwargs))
        state0, tlist, e_ops = task_args
        _b200_batch(solver, state0, tlist, e_ops, list(values), floor, 1.0, reduce_func, task)
        return None
    # mixed initial states: group the trajectory ids by initial state, one batch per group
    if inner is task:                 # values = ids, task_args = (seeds, ics, tlist, e_ops)
        seeds, ics, tlist, e_ops = task_args
        items = [(int(v), float(task_kwargs.get("jump_prob_floor", 0.0))) for v in values]
    else:                             # values = (id, jump_prob_floor), the rest by keyword
        seeds, ics = task_kwargs["seeds"], task_kwargs["ics"]
        tlist, e_ops = task_kwargs["tlist"], task_kwargs["e_ops"]
        items = [(int(v[0]), float(v[1])) for v in values]
    groups = {}
    for tid, floor in items:
        groups.setdefault((ics.get_state_index(tid), floor), []).append(tid)
    for (_, floor), ids in sorted(groups.items(), key=lambda kv: kv[1][0]):
        state, weight = ics.get_state_and_weight(ids[0])
        stop = _b200_batch(solver, state, tlist, e_ops, [seeds[t] for t in ids], floor, weight,
                           reduce_func, solver._run_one_traj)
        if stop:
            break
    return None


# devices the "b200" map shards over (None: the current device only).  Set with
# ``configure(devices=[0, 1, ...])`` / ``configure(devices="all")`` or QUTIP_B200_DEVICES=0,1,..
_DEVICES = None


def configure(devices=None):
    """Choose the GPUs ``options={"map": "b200"}`` uses: a list of device indices, ``"all"``
    (every device of the box) or ``None`` (the current device).  Trajectories are sharded in
    contiguous blocks of the seed list, every device holds its own copy of the
    operators, and the expectation sums are combined by ONE ncclAllReduce."""
    global _DEVICES
    if devices == "all":
        devices = list(range(max(1, _lib.device_count())))
    _DEVICES = None if devices is None else [int(d) for d in devices]
    return _DEVICES


def _map_devices():
    if _DEVICES is not None:
        return list(_DEVICES)
    env = os.environ.get("QUTIP_B200_DEVICES", "").strip()
    if env == "all":
        return list(range(max(1, _lib.device_count())))
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return None


def _run_engine_batch(make_engine, psi0, tlist, draws, gens, max_collapses):
    """One engine, trajectories = rows of ``draws``: run, then re-run the trajectories whose
    threshold table (status -12) or collapse record (status -13) was too short with longer
    ones -- the reference has neither limit (mcsolve.py:371-406 appends to python lists)."""
    ntraj, ndraws = draws.shape
    eng = make_engine(max_collapses)
    r = eng.run_mcsolve(psi0, tlist, draws, ntraj=ntraj)
    while ((r.status == -12) | (r.status == -13)).any():
        todo = np.nonzero((r.status == -12) | (r.status == -13))[0]
        if (r.status[todo] == -13).any():
            max_collapses *= 4
            eng_retry = make_engine(max_collapses)
        else:
            eng_retry = eng
        if (r.status[todo] == -12).any():
            if gens is None:
                raise IntegratorException(E.STATUS_MESSAGES[-12])
            more = np.stack([gens[j].random(3 * ndraws) for j in range(ntraj)])
            draws = np.concatenate([draws, more], axis=1)
            ndraws = draws.shape[1]
        r2 = eng_retry.run_mcsolve(psi0, tlist, np.ascontiguousarray(draws[todo]), ntraj=len(todo))
        if r2.col_t.shape[1] > r.col_t.shape[1]:
            for key in ("col_t", "col_which"):
                wide = np.zeros((ntraj, r2[key].shape[1]), dtype=r[key].dtype)
                wide[:, :r[key].shape[1]] = r[key]
                r[key] = wide
        w = r2.col_t.shape[1]
        for key in ("expect", "status", "ncol", "stats"):
            r[key][todo] = r2[key]
        r.col_t[todo, :w] = r2.col_t
        r.col_which[todo, :w] = r2.col_which
        if r.states is not None:
            r.states[todo] = r2.states
    r["engine"] = eng
    return r


def _b200_batch(solver, state0, tlist, e_ops, seeds, floor, weight, reduce_func, task):
    """One device batch: all trajectories of one initial state.  Returns True when
    ``reduce_func`` asked to stop."""
    if floor >= 1 - solver.options["norm_tol"]:
        # dark initial state under improved sampling: the reference returns all-zero
        # trajectories without integrating (mcsolve.py:543-556); nothing to run on the device
        for seed in seeds:
            sd, res, w = solver._run_one_traj(seed, state0, tlist, e_ops, jump_prob_floor=floor)
            if reduce_func is not None:
                remaining = reduce_func((sd, res, w * weight))
                if remaining is not None and remaining <= 0:
                    return True
        return False
    ntraj = len(seeds)
    opts = solver.options
    method = _device_method(opts["method"])
    rhs = solver.rhs
    e_dict = e_ops if isinstance(e_ops, dict) else dict(enumerate(e_ops or []))
    # python-callable e_ops f(t, state) (solver/result.py:29-77) cannot run on the device: the
    # trajectories' states are brought back and the callables evaluated on them, the operator
    # e_ops stay device-side expectation passes
    host_keys = [k for k, e in e_dict.items() if not isinstance(e, (qutip.Qobj, QobjEvo))]
    for k in host_keys:
        if not callable(e_dict[k]):
            raise TypeError("e_ops must be Qobj, QobjEvo or callables f(t, state)")
    dev_dict = {k: e for k, e in e_dict.items() if k not in host_keys}
    dev_index = {k: m for m, k in enumerate(dev_dict)}
    issuper = bool(rhs.rhs.issuper)
    if issuper and any(not isinstance(e, qutip.Qobj) for e in dev_dict.values()):
        raise TypeError("time-dependent e_ops are not combined with a superoperator Hamiltonian in "
                        "the 'b200' map")
    want_states = bool(opts["store_states"]) or (opts["store_states"] is None and not e_dict)
    want_final = bool(opts["store_final_state"])
    need_states = want_states or bool(host_keys)
    if issuper:
        # the trajectories evolve the column-stacked rho (mcsolve.py:481-490): tr(E rho) is the
        # linear functional sum_r vec(E^T)[r] rho[r] (core/data/expect.pyx:146-158)
        if any(not e.isoper for e in dev_dict.values()):
            raise TypeError("e_ops must be operators for a superoperator Hamiltonian in the 'b200' map")
        e_evos = [QobjEvo(qutip.Qobj(sp.csr_matrix(solve.trace_functional(e.full()))))
                  for e in dev_dict.values()]
    else:
        e_evos = [QobjEvo(e) if isinstance(e, qutip.Qobj) else e for e in dev_dict.values()]
    iopt = solver._integrator._integrator.options
    psi0 = _data.to(_data.Dense, state0).to_array().reshape(-1, order="F")
    tlist = np.asarray(tlist, dtype=float)
    N = psi0.size
    ndraws = 64
    gens = [s if hasattr(s, "random") else solver._get_generator(s) for s in seeds]
    draws = np.stack([g.random(ndraws) for g in gens])
    devices = _map_devices()
    multi = devices is not None and len(devices) > 1 and ntraj >= 2 * len(devices)
    store = int(need_states or want_final)

    def fingerprint():
        """identity of everything the device system is built from; the objects are kept alive
        by the cache entry so that ids cannot be recycled.  ``solver.run(args=...)`` replaces
        the coefficient objects of the rhs, which changes the key."""
        objs = []
        for q in [rhs.rhs] + list(rhs.c_ops) + list(rhs.n_ops):
            for el in q.to_list():
                objs.extend(el if isinstance(el, (list, tuple)) else [el])
        objs.extend(dev_dict.values())
        return tuple(id(o) for o in objs), objs

    opt_key = (method, floor, store, tuple(sorted((k, repr(v)) for k, v in iopt.items())),
               tuple(repr(opts[k]) for k in ("norm_steps", "norm_t_tol", "norm_tol", "norm_min_step",
                                             "mc_corr_eps")))

    def shard(lo, hi, device):
        if device is not None:
            E.set_device(device)
        # device system and engines are kept on the solver between runs (operators uploaded and
        # converted once, state pool allocated once)
        ids, objs = fingerprint()
        cache = solver.__dict__.setdefault("_b200_cache", _DeviceCache())
        key = (device, ids)
        entry = cache.get(key)
        if entry is None:
            system = system_from_qobjevo(rhs.rhs, rhs.c_ops, rhs.n_ops, e_evos, allow_host=True,
                                         functional=issuper)
            if issuper:
                system.set_mc_trace(int(round(np.sqrt(N))))
            if system.has_host:
                raise TypeError("python-callable coefficients need a host evaluation per RHS call; the "
                                "'b200' map runs whole batches on the device and cannot use them. Use "
                                "string/array coefficients, or method='b200_vern7' with a stock map.")
            for k in [k for k in cache if k[0] == device]:
                del cache[k]                       # one system per device and solver
            entry = cache[key] = dict(system=system, engines={}, refs=objs)
        system = entry["system"]
        nslots = min(hi - lo, solve.default_nslots(N, method))

        def make_engine(max_collapses):
            ek = (opt_key, max_collapses)
            eng = entry["engines"].get(ek)
            if eng is not None and eng.nslots >= min(hi - lo, nslots):
                return eng
            entry["engines"].clear()               # frees the previous state pool first
            eng = entry["engines"][ek] = E.Engine(
                system, method, nslots=nslots, atol=iopt['atol'], rtol=iopt['rtol'],
                nsteps=int(iopt['nsteps']), first_step=float(iopt['first_step'] or 0),
                min_step=float(iopt['min_step'] or 0), max_step=float(iopt['max_step'] or 0),
                interpolate=int(bool(iopt.get('interpolate', True))),
                max_order=int(iopt.get('order', 0) or 0), norm_steps=int(opts['norm_steps']),
                norm_t_tol=opts['norm_t_tol'], norm_tol=opts['norm_tol'],
                norm_min_step=opts['norm_min_step'], mc_corr_eps=opts['mc_corr_eps'],
                store_states=store, jump_prob_floor=floor, max_collapses=max_collapses)
            return eng

        return _run_engine_batch(make_engine, psi0, tlist, np.ascontiguousarray(draws[lo:hi]),
                                 gens[lo:hi], _MAX_COLLAPSES)

    sums = None
    if not multi:
        r = shard(0, ntraj, devices[0] if devices else None)
    else:
        # contiguous blocks of the seed list, one host thread per device (the C ABI calls
        # release the GIL); ONE ncclAllReduce of the device-side expectation sums
        world = len(devices)
        comm = E.Comm.all(devices)
        parts, errors = [None] * world, []

        def work(i):
            try:
                lo, hi = solve.shard_range(ntraj, i, world)
                if hi > lo:
                    parts[i] = shard(lo, hi, devices[i])
            except BaseException as exc:
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        live = [q for q in parts if q is not None]
        if dev_dict:
            sums = comm.reduce_expect([None if q is None else q.engine for q in parts],
                                      len(dev_dict), len(tlist))
        comm.free()
        width = max(q.col_t.shape[1] for q in live)

        def cat(key, pad=False):
            arrs = [q[key] for q in live]
            if pad:
                arrs = [np.pad(x, ((0, 0), (0, width - x.shape[1]))) for x in arrs]
            return np.concatenate(arrs)

        r = E.RunResult(expect=cat("expect"), status=cat("status"), ncol=cat("ncol"),
                        col_t=cat("col_t", True), col_which=cat("col_which", True), stats=cat("stats"),
                        states=None if live[0].states is None else cat("states"))
    bad = np.nonzero(r.status != 1)[0]
    if bad.size:
        st = int(r.status[bad[0]])
        if st == -10:
            raise RuntimeError(E.STATUS_MESSAGES[-10])
        raise IntegratorException(E.STATUS_MESSAGES.get(st, "integration failed"))
    herm = [bool(e.isherm) if isinstance(e, qutip.Qobj) else False for e in dev_dict.values()]
    w_traj = (1 - floor) * weight                                        # mcsolve.py:565

    def make_result(j):
        res = solver._trajectory_resultclass(e_dict, solver.options)
        res.times = list(tlist)
        qs = None
        if need_states or want_final:
            qs = [solver._restore_state(_data.Dense(r.states[j, i].reshape(-1, 1)), copy=False)
                  for i in (range(len(tlist)) if need_states else [len(tlist) - 1])]
        for k in res.e_data:
            if k in dev_index:
                m = dev_index[k]
                vals = r.expect[j, m]
                res.e_data[k].extend((vals.real if herm[m] else vals).tolist())
            else:                                   # callable e_op on the downloaded states
                res.e_data[k].extend([e_dict[k](t, q) for t, q in zip(tlist, qs)])
        if want_states or want_final:
            if want_states:
                res.states = qs
            res._final_state = qs[-1]
        res.collapse = [(float(r.col_t[j, i]), int(r.col_which[j, i])) for i in range(r.ncol[j])]
        if martingale is not None:
            # NonMarkovianMCSolver._run_one_traj (nm_mcsolve.py:562-570): the influence martingale
            # at the output times, from the solver's own InfluenceMartingale object (continuous
            # part pre-computed by NonMarkovianMCSolver.run, discrete part from the collapses)
            martingale.initialize(tlist[0], cache='keep')
            for tc, ch in res.collapse:
                martingale.add_collapse(tc, ch)
            res.trace = [martingale.value(t) for t in tlist]
        return res

    first = 0
    martingale = getattr(solver, "_martingale", None)
    target = getattr(reduce_func, "__self__", None)
    if _bulk_feed_ok(target, need_states, want_final) and ntraj > 1:
        # averages only: trajectory 0 goes through McResult.add (it sizes the accumulators),
        # the others are added to the running sums in bulk -- what _TrajectorySum.reduce_expect
        # (multitrajresult.py:1116-1124) and _McBaseResult._add_collapse do one by one
        remaining = reduce_func((seeds[0], make_result(0), w_traj))
        if remaining is not None and remaining <= 0:
            return True
        room = ntraj - 1
        if target._target_ntraj is not None:
            room = min(room, int(target._target_ntraj - target.num_trajectories))
        if room > 0:
            sel = slice(1, 1 + room)
            target.seeds.extend(seeds[sel])
            target._trajectories_weight_info.extend([w_traj] * room)
            target.num_trajectories += room
            for m in range(len(dev_dict)):
                vals = r.expect[sel, m]
                vals = vals.real if herm[m] else vals
                if sums is not None and room == ntraj - 1 and not np.iscomplexobj(vals):
                    # multi-device: the NCCL-reduced sums of ALL trajectories minus trajectory 0
                    v0 = r.expect[0, m].real
                    s1 = sums[0][m].real - v0
                    s2 = sums[1][m].real - v0 * v0
                else:
                    s1, s2 = vals.sum(axis=0), (vals ** 2).sum(axis=0)
                target._sum_rel.sum_expect[m] += w_traj * s1
                target._sum_rel.sum2_expect[m] += w_traj * s2
            ct, cw, nc = r.col_t, r.col_which, r.ncol
            target.collapse.extend(
                [list(zip(ct[j, :nc[j]].tolist(), cw[j, :nc[j]].tolist())) for j in range(1, 1 + room)])
        remaining = target._early_finish_check()
        return remaining is not None and remaining <= 0
    for j in range(first, ntraj):
        if reduce_func is not None:
            remaining = reduce_func((seeds[j], make_result(j), w_traj))
            if remaining is not None and remaining <= 0:
                return True
    return False


class _DeviceCache(dict):
    """device systems / engines kept on a solver between runs; device handles do not travel,
    so a pickled solver (process maps) starts with an empty cache"""

    def __reduce__(self):
        return (_DeviceCache, ())


_MAX_COLLAPSES = 64       # initial capacity of the per-trajectory collapse record (grown on demand)


def _bulk_feed_ok(target, want_states, want_final):
    """The vectorised feed replaces exactly the processors a plain averaging McResult runs per
    trajectory (_increment_traj, _reduce_expect, _add_collapse); anything else -- kept
    trajectories, averaged states, target tolerances, subclasses -- takes McResult.add."""
    from qutip.solver.multitrajresult import McResult, MultiTrajResult, _McBaseResult
    if type(target) is not McResult or want_states or want_final:
        return False
    if target.options["keep_runs_results"] or target.runs_e_data:
        return False
    try:
        procs = {getattr(pr, "__func__", None) for pr in target._state_processors}
        allowed = {MultiTrajResult._increment_traj, MultiTrajResult._reduce_expect,
                   _McBaseResult._add_collapse}
        check = getattr(target._early_finish_check, "__func__", None)
        return procs <= allowed and check in (MultiTrajResult._fixed_end, MultiTrajResult._no_end)
    except AttributeError:
        return False


register()
