"""oracle/rk_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's time-evolution hot path:

* operators  : CSR / Dia / Dense matvec through oracle/spmv_oracle.c
* OrcEvo     : QobjEvo.matmul_data        (qutip/core/cy/qobjevo.pyx:1103-1116,
                                           _element.pyx:117-182)
* ExplicitRK : Explicit_RungeKutta        (qutip/solver/integrator/explicit_rk.pyx:47-530)
* MCIntegratorOracle : MCIntegrator       (qutip/solver/mcsolve.py:228-414)
* mesolve_oracle / mcsolve_oracle : the Solver.run / MultiTrajSolver loops
                                          (solver_base.py:159-226, multitraj.py:260-283)

Control flow is kept line-for-line equivalent to the reference so that the step
sequence, RHS-evaluation counts and jump records can be compared exactly; see
tests/test_oracle.py for the pinning against the reference build.
"""
import ctypes
import json
import math
import os

import numpy as np

from . import lib

_c = ctypes
_P = _c.c_void_p


def _ptr(a):
    return _c.c_void_p(a.ctypes.data)


# --------------------------------------------------------------------------- operators
class OrcOp:
    """Operator in one of the reference's three layouts (core/data/{csr,dia,dense}.pxd)."""

    def __init__(self, kind, shape, **arrs):
        self.kind, self.shape = kind, tuple(shape)
        self.__dict__.update(arrs)

    @classmethod
    def csr(cls, data, col, rowptr, shape):
        return cls("csr", shape,
                   data=np.ascontiguousarray(data, dtype=np.complex128),
                   col=np.ascontiguousarray(col, dtype=np.int32),
                   rowptr=np.ascontiguousarray(rowptr, dtype=np.int32))

    @classmethod
    def dia(cls, data, offsets, shape):
        return cls("dia", shape,
                   data=np.ascontiguousarray(data, dtype=np.complex128),
                   offsets=np.ascontiguousarray(offsets, dtype=np.int32))

    @classmethod
    def dense(cls, arr):
        arr = np.asfortranarray(arr, dtype=np.complex128)
        return cls("dense", arr.shape, arr=arr)

    @classmethod
    def from_scipy(cls, m):
        import scipy.sparse as sp
        if sp.isspmatrix_dia(m) or isinstance(m, sp.dia_array):
            return cls.dia(m.data, m.offsets, m.shape)
        m = sp.csr_matrix(m)
        return cls.csr(m.data, m.indices, m.indptr, m.shape)

    def to_dense_array(self):
        out = np.zeros(self.shape, dtype=np.complex128)
        eye = np.eye(self.shape[1], dtype=np.complex128)
        for c in range(self.shape[1]):
            col = np.zeros(self.shape[0], dtype=np.complex128)
            self.matvec_acc(eye[:, c].copy(), 1.0, col)
            out[:, c] = col
        return out

    def matvec_acc(self, x, scale, out):
        """out += scale * A @ x   (x, out: 1-D complex128 contiguous)."""
        L = lib()
        scale = complex(scale)
        nrows, ncols = self.shape
        assert x.shape == (ncols,) and out.shape == (nrows,)
        if self.kind == "csr":
            L.orc_csr_matvec(_ptr(self.data), _ptr(self.col), _ptr(self.rowptr),
                             _c.c_int64(nrows), _ptr(x), _c.c_double(scale.real),
                             _c.c_double(scale.imag), _ptr(out))
        elif self.kind == "dia":
            # reference: un-scaled product into a zeroed temp, then out += scale*tmp
            # (matmul.pyx:431-440,499-505)
            tmp = np.zeros(nrows, dtype=np.complex128)
            L.orc_dia_matvec_noscale(_ptr(self.data), _ptr(self.offsets),
                                     _c.c_int64(len(self.offsets)), _c.c_int64(nrows),
                                     _c.c_int64(ncols), _ptr(x), _ptr(tmp))
            L.orc_axpy(_c.c_int64(nrows), _c.c_double(scale.real),
                       _c.c_double(scale.imag), _ptr(tmp), _ptr(out))
        else:
            L.orc_dense_matvec_f(_ptr(self.arr), _c.c_int64(nrows), _c.c_int64(ncols),
                                 _ptr(x), _c.c_double(scale.real),
                                 _c.c_double(scale.imag), _ptr(out))
        return out

    def matvec(self, x, scale=1.0):
        return self.matvec_acc(np.ascontiguousarray(x, dtype=np.complex128), scale,
                               np.zeros(self.shape[0], dtype=np.complex128))

    def expect_ket(self, y):
        """<y|A|y>  (core/data/expect.pyx:133-144)."""
        return np.vdot(y, self.matvec(y))


class OrcEvo:
    """sum_k coeff_k(t) * A_k.  ``elements`` = [(OrcOp, coeff)], coeff a complex
    constant or a callable t -> complex.  Order follows QobjEvo.compress():
    time-dependent elements first, the constant part last (qobjevo.pyx:816-867)."""

    def __init__(self, elements):
        self.elements = list(elements)
        self.shape = self.elements[0][0].shape
        self.nevals = 0

    def matmul(self, t, x, out=None, scale=1.0):
        # qobjevo.pyx:1103-1116 : out += scale * coeff(t) * A x for each element
        if out is None:
            out = np.zeros(self.shape[0], dtype=np.complex128)
        for op, cf in self.elements:
            c = cf(t) if callable(cf) else cf
            op.matvec_acc(x, complex(c) * scale, out)
        self.nevals += 1
        return out

    def expect(self, t, y):
        # QobjEvo.expect_data (qobjevo.pyx:1018-1072) on a ket
        tot = 0j
        for op, cf in self.elements:
            c = cf(t) if callable(cf) else cf
            tot += complex(c) * op.expect_ket(y)
        return tot


# --------------------------------------------------------------------------- tableaux
def load_tableau(name):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "qutip_b200", "tableaux.json")
    js = json.load(open(path))[name]
    fh = float.fromhex
    return {
        "order": js["order"],
        "a": np.array([[fh(x) for x in r] for r in js["a"]]),
        "b": np.array([fh(x) for x in js["b"]]),
        "c": np.array([fh(x) for x in js["c"]]),
        "e": np.array([fh(x) for x in js["e"]]),
        "bi": np.array([[fh(x) for x in r] for r in js["bi"]]),
    }


AT_FRONT, INTERPOLATED, NORMAL = 2, 1, 0
TOO_MUCH_WORK, DT_UNDERFLOW, OUTSIDE_RANGE, NOT_INITIATED = -1, -2, -3, -4


def wrmn_error(diff, state, atol, rtol):
    # core/data/ode.pyx:39-64
    return float(lib().orc_wrmn_error(_c.c_int64(diff.size), _ptr(diff), _ptr(state),
                                      _c.c_double(atol), _c.c_double(rtol)))


def frobenius(x):
    # core/data/norm.pyx:127-130 (dznrm2)
    return float(np.linalg.norm(x))


class ExplicitRK:
    """Restatement of Explicit_RungeKutta (explicit_rk.pyx:47-530)."""

    def __init__(self, deriv, tableau, rtol=1e-6, atol=1e-8, nsteps=1000, first_step=0,
                 min_step=0, max_step=0, interpolate=True):
        self.deriv = deriv                      # deriv(t, y) -> new array
        self.atol, self.rtol = atol, rtol
        self._dt_safe = atol
        self.max_numsteps = nsteps
        self.first_step = first_step
        self.min_step = min_step or 1e-15      # :115
        self.max_step = max_step
        self.order = tableau["order"]
        self.a, self.b, self.c = tableau["a"], tableau["b"], tableau["c"]
        self.e, self.bi = tableau.get("e"), tableau.get("bi")
        self.rk_step = self.b.shape[0]
        self.rk_extra_step = self.c.shape[0]
        self.adaptative_step = self.e is not None
        self.interpolate = bool(interpolate) and self.bi is not None   # :176
        if self.interpolate:
            self.denseout_order = self.bi.shape[1]
        self.first_same_as_last = bool(np.allclose(
            self.a[self.rk_step - 1, :self.rk_step], self.b, atol=1e-14))   # :190-192
        self._y_prev = None
        self._status = NOT_INITIATED
        # statistics the device engine must reproduce
        self.n_accept = 0
        self.n_reject = 0
        self.step_log = []        # (t_prev, dt, err) of every attempted step

    # :204-230
    def set_initial_value(self, y0, t):
        self._t = self._t_prev = self._t_front = t
        self._dt_int = 0.0
        self._y = y0
        n_k = self.rk_extra_step if self.interpolate else self.rk_step
        self.k = [y0.copy() for _ in range(n_k)]
        self._y_temp = y0.copy()
        self._y_front = y0.copy()
        self._y_prev = y0.copy()
        if self.first_same_as_last:
            self._k_fsal = self.deriv(t, y0)
        if not self.first_step:
            self._dt_safe = self._estimate_first_step(t, self._y)
        else:
            self._dt_safe = self.first_step
        self._status = NORMAL

    # :232-276
    def _estimate_first_step(self, t, y0):
        if not self.adaptative_step:
            return 0.0
        norm = frobenius(y0)
        tol = self.atol + norm * self.rtol
        if self.first_same_as_last:
            self.k[0] = self._k_fsal.copy()
        else:
            self.k[0] = self.deriv(t, y0)
        if norm <= self.atol:
            norm = 1
        tmp_norm = frobenius(self.k[0])
        factorial = 1.0
        for i in range(1, self.order + 1):
            factorial *= i
        if tmp_norm >= self.atol * 1e-6:
            dt1 = (tol * factorial * norm ** self.order) ** (1 / (self.order + 1)) / tmp_norm
        else:
            dt1 = (tol * factorial) ** (1 / (self.order + 1)) * norm * 0.5
        t1 = t + dt1 / 100
        self._y_temp = y0 + (dt1 / 100) * self.k[0]
        self.k[1] = self.deriv(t1, self._y_temp)
        tmp_norm = frobenius(self.k[1])
        if tmp_norm >= self.atol * 1e-6:
            dt2 = (tol * factorial * norm ** self.order) ** (1 / (self.order + 1)) / tmp_norm
        else:
            dt2 = dt1
        dt = min(dt1, dt2)
        if self.max_step:
            dt = min(self.max_step, dt)
        if self.min_step:
            dt = max(self.min_step, dt)
        return dt

    # :278-331
    def integrate(self, t, step=False):
        nsteps_left = self.max_numsteps
        if self._y_prev is None:
            self._status = NOT_INITIATED
            return
        if t == self._t:
            return
        if t < self._t_prev:
            self._status = OUTSIDE_RANGE
            return
        if self.interpolate and t < self._t_front:
            if self._status != INTERPOLATED:
                self._prep_dense_out()
                self._status = INTERPOLATED
            self._y = self._interpolate_step(t)
            self._t = t
            return
        self._status = NORMAL
        if step and self._t < self._t_front and t > self._t_front:
            t = self._t_front
        while self._t_front < t and self._status >= 0:
            self._y_prev = self._y_front.copy()
            self._t_prev = self._t_front
            nsteps_left -= self._step_in_err(t, nsteps_left)
            if step:
                break
        if self._status < 0:
            return
        if self._t_front > t:
            self._prep_dense_out()
            self._status = INTERPOLATED
            self._t = t
            self._y = self._interpolate_step(t)
        else:
            self._status = AT_FRONT
            self._t = self._t_front
            self._y = self._y_front.copy()

    # :333-356
    def _step_in_err(self, t, max_step):
        error = 1.0
        nsteps = 0
        while error >= 1:
            dt = self._get_timestep(t)
            error = self._compute_step(dt)
            self._dt_int = dt
            self._recompute_safe_step(error, dt)
            self.step_log.append((self._t_prev, dt, error))
            if error >= 1:
                self.n_reject += 1
            else:
                self.n_accept += 1
            if dt == self.min_step and error > 1:
                self._status = DT_UNDERFLOW
                break
            nsteps += 1
            if nsteps > max_step:
                self._status = TOO_MUCH_WORK
                break
        if self.first_same_as_last:
            self._k_fsal = self.k[self.rk_step - 1].copy()
        return nsteps

    def _accumulate(self, target, factors, dt, size):
        # :432-437 with iadd_data's skip of exact zeros (:34-39)
        for i in range(size):
            f = dt * factors[i]
            if f == 0:
                continue
            target += f * self.k[i]
        return target

    # :358-389
    def _compute_step(self, dt):
        if self.first_same_as_last:
            self.k[0] = self._k_fsal.copy()
        else:
            self.k[0] = self.deriv(self._t_prev, self._y_prev)
        for i in range(1, self.rk_step):
            self._y_temp = self._accumulate(self._y_prev.copy(), self.a[i, :], dt, i)
            self.k[i] = self.deriv(self._t_prev + self.c[i] * dt, self._y_temp)
        self._y_front = self._accumulate(self._y_prev.copy(), self.b, dt, self.rk_step)
        self._t_front = self._t_prev + dt
        return self._error(self._y_front, dt)

    # :391-397
    def _error(self, y_new, dt):
        if not self.adaptative_step:
            return 0.0
        self._y_temp = self._accumulate(np.zeros_like(y_new), self.e, dt, self.rk_step)
        return wrmn_error(self._y_temp, y_new, self.atol, self.rtol)

    # :399-410
    def _prep_dense_out(self):
        dt = self._dt_int
        for i in range(self.rk_step, self.rk_extra_step):
            self._y_temp = self._accumulate(self._y_prev.copy(), self.a[i, :], dt, i)
            self.k[i] = self.deriv(self._t_prev + self.c[i] * dt, self._y_temp)

    # :412-430
    def _interpolate_step(self, t):
        t0, dt = self._t_prev, self._dt_int
        tau = (t - t0) / dt
        bf = np.zeros(self.rk_extra_step)
        for i in range(self.rk_extra_step):
            v = 0.0
            for j in range(self.denseout_order - 1, -1, -1):
                v += self.bi[i, j]
                v *= tau
            bf[i] = v
        return self._accumulate(self._y_prev.copy(), bf, dt, self.rk_extra_step)

    # :440-449
    def _get_timestep(self, t):
        dt_needed = t - self._t_prev
        if not self.adaptative_step:
            return dt_needed
        if self.interpolate:
            return self._dt_safe
        if dt_needed <= self._dt_safe:
            return dt_needed
        return dt_needed / (int(dt_needed / self._dt_safe) + 1)

    # :451-467
    def _recompute_safe_step(self, err, dt):
        if not self.adaptative_step:
            return
        if err == 0:
            factor = 10.0
        else:
            factor = 0.9 * err ** (-1 / (self.order + 1))
            factor = min(10, factor)
            factor = max(0.2, factor)
        self._dt_safe = dt * factor
        if self.max_step:
            self._dt_safe = min(self.max_step, self._dt_safe)
        if self.min_step:
            self._dt_safe = max(self.min_step, self._dt_safe)

    # wrapper semantics of IntegratorVern7 (qutip_integrator.py:69-92)
    def set_state(self, t, y):
        self.set_initial_value(np.array(y, dtype=np.complex128, copy=True), t)

    def get_state(self):
        return self._t, self._y

    def mcstep(self, t):
        self.integrate(t, step=True)
        if self._status < 0:
            raise RuntimeError("integration failed, status %d" % self._status)
        return self.get_state()

    def integrate_to(self, t):
        self.integrate(t, step=False)
        if self._status < 0:
            raise RuntimeError("integration failed, status %d" % self._status)
        return self.get_state()


# --------------------------------------------------------------------------- mcsolve
MC_DEFAULTS = dict(norm_steps=25, norm_t_tol=1e-6, norm_tol=1e-4, norm_min_step=0.1,
                   mc_corr_eps=1e-10)


class MCIntegratorOracle:
    """Restatement of MCIntegrator (solver/mcsolve.py:228-414) for ket trajectories.
    ``draws`` supplies the successive generator.random() values."""

    def __init__(self, rk, c_ops, n_ops, options=None):
        self.rk, self.c_ops, self.n_ops = rk, c_ops, n_ops
        self.options = dict(MC_DEFAULTS)
        self.options.update(options or {})

    def set_state(self, t, y0, draws, no_jump=False, jump_prob_floor=0.0):
        self.collapses = []
        self.draws = iter(draws)
        self.n_draws = 0
        if no_jump:
            self.target_norm = 0.0
        else:
            self.target_norm = self._random() * (1 - jump_prob_floor) + jump_prob_floor
        self.rk.set_state(t, y0)

    def _random(self):
        self.n_draws += 1
        return float(next(self.draws))

    @staticmethod
    def _prob(y):
        return frobenius(y) ** 2          # :311-314

    def integrate(self, t):
        # :286-302
        t_old, y_old = self.rk.get_state()
        norm_old = self._prob(y_old)
        while t_old < t:
            t_step, state = self.rk.mcstep(t)
            norm = self._prob(state)
            if norm <= self.target_norm:
                t_col, state = self._find_collapse_time(norm_old, norm, t_old, t_step)
                self._do_collapse(t_col, state)
                t_old, y_old = self.rk.get_state()
                norm_old = 1.0
            else:
                t_old, y_old = t_step, state
                norm_old = norm
        return t_old, y_old * (1 / frobenius(y_old))

    def _find_collapse_time(self, norm_old, norm, t_prev, t_final):
        # :321-369
        o = self.options
        tries = 0
        while tries < o["norm_steps"]:
            tries += 1
            if (t_final - t_prev) < o["norm_t_tol"]:
                t_guess = t_final
                _, state = self.rk.get_state()
                break
            dt = t_final - t_prev
            ratio = np.log(norm_old / self.target_norm) / np.log(norm_old / norm)
            if ratio < o["norm_min_step"]:
                ratio = o["norm_min_step"]
            if ratio > (1 - o["norm_min_step"]):
                ratio = 1 - o["norm_min_step"]
            t_guess = t_prev + dt * ratio
            if (t_guess - t_prev) < o["norm_t_tol"]:
                t_guess = t_prev + o["norm_t_tol"]
            _, state = self.rk.mcstep(t_guess)
            norm2_guess = self._prob(state)
            if abs(self.target_norm - norm2_guess) < o["norm_tol"] * self.target_norm:
                break
            elif norm2_guess < self.target_norm:
                t_final = t_guess
                norm = norm2_guess
            else:
                t_prev = t_guess
                norm_old = norm2_guess
        if tries >= o["norm_steps"]:
            raise RuntimeError("Could not find the collapse time within desired tolerance.")
        return t_guess, state

    def _do_collapse(self, t, state):
        # :371-406
        num_ops = len(self.n_ops)
        if num_ops == 1:
            which = 0
        else:
            probs = [n_op.expect(t, state).real for n_op in self.n_ops]
            target = sum(probs) * self._random() - probs[0]
            which = 0
            while target > 0 and which <= num_ops:
                which += 1
                target -= probs[which]
        state_new = self.c_ops[which].matmul(t, state)
        new_norm = frobenius(state_new)
        if new_norm < self.options["mc_corr_eps"]:
            state_new = state * (1 / frobenius(state))
        else:
            state_new = state_new * (1 / new_norm)
            self.collapses.append((t, which))
            self.target_norm = self._random()
        self.rk.set_state(t, state_new)


def make_thresholds(seed_seq_or_int, ntraj, ndraws):
    """Per-trajectory generator.random() draws exactly as the reference makes them:
    SeedSequence(seed).spawn(ntraj) -> default_rng(child) (multitraj.py:354-393).
    Generator.random(K) yields the same K doubles as K scalar random() calls."""
    ss = seed_seq_or_int
    if not isinstance(ss, np.random.SeedSequence):
        ss = np.random.SeedSequence(ss)
    kids = ss.spawn(ntraj)
    return np.stack([np.random.default_rng(k).random(ndraws) for k in kids])


def mesolve_oracle(rhs, y0, tlist, method="vern7", e_funcs=(), options=None):
    """Solver.run loop (solver_base.py:159-226) on an already vectorised state.
    e_funcs: callables (t, y) -> complex evaluated at every tlist point."""
    rk = ExplicitRK(rhs.matmul, load_tableau(method), **(options or {}))
    y0 = np.ascontiguousarray(y0, dtype=np.complex128)
    rk.set_state(tlist[0], y0)
    states = [y0.copy()]
    for t in tlist[1:]:
        _, y = rk.integrate_to(t)
        states.append(y.copy())
    expect = np.array([[f(t, y) for t, y in zip(tlist, states)] for f in e_funcs])
    return dict(states=np.array(states), expect=expect, rk=rk)


def mcsolve_oracle(rhs, c_ops, n_ops, psi0, tlist, draws, e_ops=(), method="vern7",
                   options=None, mc_options=None):
    """One trajectory of MultiTrajSolver._run_one_traj (multitraj.py:260-283)."""
    rk = ExplicitRK(rhs.matmul, load_tableau(method), **(options or {}))
    mc = MCIntegratorOracle(rk, c_ops, n_ops, mc_options)
    psi0 = np.ascontiguousarray(psi0, dtype=np.complex128)
    mc.set_state(tlist[0], psi0, draws)
    states = [psi0.copy()]
    for t in tlist[1:]:
        _, y = mc.integrate(t)
        states.append(y.copy())
    expect = np.array([[e.expect_ket(y) for y in states] for e in e_ops])
    return dict(states=np.array(states), expect=expect, collapses=list(mc.collapses),
                n_draws=mc.n_draws, rk=rk)
