"""Compile time-dependent coefficients into the engine's stack byte-code.

Host-side counterpart of the reference's Coefficient family
(qutip/core/cy/coefficient.pyx:88-988, qutip/core/coefficient.py:399-589).  The reference
JIT-compiles string coefficients into Cython modules and calls them from the RK loop on
the host; here each coefficient becomes a short program (``QbInstr`` in
csrc/qb_types.h) that the device-side step controller evaluates for every stage time,
so a time-dependent system needs no host round trip per RHS evaluation.

Supported sources
  * python / numpy numbers                         -> constant
  * expression strings in ``t`` and named args     -> parsed with ``ast`` (the grammar the
    reference documents for string coefficients: + - * / **, sin cos tan exp log sqrt
    sinh cosh tanh abs real imag conj pi, complex literals)
  * ``("spline", id)`` array coefficients           -> InterCoefficient tables
  * combinators ``conj``, ``sum``, ``mul``, ``norm`` (ConjCoefficient, SumCoefficient,
    MulCoefficient, NormCoefficient)
Anything else (python callables, feedback) cannot run on the device and raises
``TypeError`` -- there is no CPU fallback.
"""
import ast
import ctypes

import numpy as np

# op-codes: keep in sync with csrc/qb_types.h
(I_CONST, I_T, I_ARG, I_ADD, I_SUB, I_MUL, I_DIV, I_NEG, I_CONJ, I_SIN, I_COS, I_TAN, I_EXP,
 I_LOG, I_SQRT, I_ABS, I_REAL, I_IMAG, I_POW, I_SINH, I_COSH, I_TANH, I_SPLINE, I_ASIN,
 I_ACOS, I_ATAN, I_NORM2, I_HEAVISIDE_GE, I_HOST, I_MIN_RE, I_SQRT_RE) = range(31)

_FUNCS = {
    "sin": I_SIN, "cos": I_COS, "tan": I_TAN, "exp": I_EXP, "log": I_LOG, "sqrt": I_SQRT,
    "abs": I_ABS, "real": I_REAL, "imag": I_IMAG, "conj": I_CONJ, "sinh": I_SINH,
    "cosh": I_COSH, "tanh": I_TANH, "asin": I_ASIN, "acos": I_ACOS, "atan": I_ATAN,
    "arcsin": I_ASIN, "arccos": I_ACOS, "arctan": I_ATAN, "norm": I_NORM2,
}
_CONSTS = {"pi": np.pi, "e": np.e}


class QbInstr(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int), ("iarg", ctypes.c_int), ("re", ctypes.c_double),
                ("im", ctypes.c_double)]


class Program:
    """A compiled coefficient: list of (op, iarg, re, im)."""

    def __init__(self, instrs=()):
        self.instrs = list(instrs)

    def __len__(self):
        return len(self.instrs)

    def as_ctypes(self):
        arr = (QbInstr * max(1, len(self.instrs)))()
        for i, (op, ia, re, im) in enumerate(self.instrs):
            arr[i].op, arr[i].iarg, arr[i].re, arr[i].im = op, ia, re, im
        return arr

    # combinators -------------------------------------------------------------------
    def conj(self):
        return Program(self.instrs + [(I_CONJ, 0, 0.0, 0.0)])

    def norm(self):
        return Program(self.instrs + [(I_NORM2, 0, 0.0, 0.0)])

    def __add__(self, other):
        return Program(self.instrs + other.instrs + [(I_ADD, 0, 0.0, 0.0)])

    def __mul__(self, other):
        return Program(self.instrs + other.instrs + [(I_MUL, 0, 0.0, 0.0)])

    def sqrt_real(self):
        """sqrt(real(.)) -- SqrtRealCoefficient (solver/cy/nm_mcsolve.pyx:80-119)"""
        return Program(self.instrs + [(I_SQRT_RE, 0, 0.0, 0.0)])

    def scaled(self, z):
        z = complex(z)
        return Program(self.instrs + [(I_CONST, 0, z.real, z.imag), (I_MUL, 0, 0.0, 0.0)])


def rate_shift(programs):
    """2 * abs(min(0, real(c_1(t)), real(c_2(t)), ...)) -- RateShiftCoefficient
    (solver/cy/nm_mcsolve.pyx:15-78)"""
    instrs = [(I_CONST, 0, 0.0, 0.0)]
    for p in programs:
        instrs += list(p.instrs) + [(I_MIN_RE, 0, 0.0, 0.0)]
    instrs += [(I_ABS, 0, 0.0, 0.0), (I_CONST, 0, 2.0, 0.0), (I_MUL, 0, 0.0, 0.0)]
    return Program(instrs)


def constant(z):
    z = complex(z)
    return Program([(I_CONST, 0, z.real, z.imag)])


def host():
    """Placeholder program: the value is supplied by the host for every evaluation time
    (python-callable coefficients)."""
    return Program([(I_HOST, 0, 0.0, 0.0)])


def spline(spline_id):
    """coefficient = InterCoefficient table ``spline_id`` evaluated at t."""
    return Program([(I_T, 0, 0.0, 0.0), (I_SPLINE, int(spline_id), 0.0, 0.0)])


class _Compiler(ast.NodeVisitor):
    def __init__(self, arg_index, arg_values):
        self.arg_index = arg_index          # name -> slot in the per-trajectory args
        self.arg_values = arg_values        # name -> bound constant
        self.out = []

    def generic_visit(self, node):
        raise TypeError("unsupported syntax in coefficient string: %s"
                        % type(node).__name__)

    def visit_Expression(self, node):
        self.visit(node.body)

    def visit_Constant(self, node):
        if not isinstance(node.value, (int, float, complex)) or isinstance(node.value, bool):
            raise TypeError("unsupported constant %r in coefficient" % (node.value,))
        z = complex(node.value)
        self.out.append((I_CONST, 0, z.real, z.imag))

    def visit_Name(self, node):
        n = node.id
        if n == "t":
            self.out.append((I_T, 0, 0.0, 0.0))
        elif n in self.arg_index:
            self.out.append((I_ARG, self.arg_index[n], 0.0, 0.0))
        elif n in self.arg_values:
            z = complex(self.arg_values[n])
            self.out.append((I_CONST, 0, z.real, z.imag))
        elif n in _CONSTS:
            self.out.append((I_CONST, 0, float(_CONSTS[n]), 0.0))
        else:
            raise TypeError("unknown name %r in coefficient string" % n)

    def visit_Attribute(self, node):
        # np.pi, numpy.pi, np.e
        if isinstance(node.value, ast.Name) and node.value.id in ("np", "numpy", "math") \
                and node.attr in _CONSTS:
            self.out.append((I_CONST, 0, float(_CONSTS[node.attr]), 0.0))
        else:
            raise TypeError("unsupported attribute in coefficient string")

    def visit_UnaryOp(self, node):
        self.visit(node.operand)
        if isinstance(node.op, ast.USub):
            self.out.append((I_NEG, 0, 0.0, 0.0))
        elif not isinstance(node.op, ast.UAdd):
            raise TypeError("unsupported unary operator")

    def visit_BinOp(self, node):
        ops = {ast.Add: I_ADD, ast.Sub: I_SUB, ast.Mult: I_MUL, ast.Div: I_DIV, ast.Pow: I_POW}
        if type(node.op) not in ops:
            raise TypeError("unsupported operator %s" % type(node.op).__name__)
        self.visit(node.left)
        self.visit(node.right)
        self.out.append((ops[type(node.op)], 0, 0.0, 0.0))

    def visit_Call(self, node):
        f = node.func
        name = f.id if isinstance(f, ast.Name) else (
            f.attr if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name)
            and f.value.id in ("np", "numpy", "math", "cmath") else None)
        if name not in _FUNCS or len(node.args) != 1 or node.keywords:
            raise TypeError("unsupported function call in coefficient string")
        self.visit(node.args[0])
        self.out.append((_FUNCS[name], 0, 0.0, 0.0))


def compile_expr(expr, args=None, arg_index=None):
    """Compile ``expr`` (python syntax in ``t``).  ``args`` binds names to constants;
    names listed in ``arg_index`` are read from the per-trajectory argument vector
    instead (parameter sweeps)."""
    tree = ast.parse(expr.strip(), mode="eval")
    c = _Compiler(dict(arg_index or {}), dict(args or {}))
    c.visit(tree)
    if len(c.out) > 256:
        raise TypeError("coefficient expression too long for the device")
    return Program(c.out)


def evaluate(prog, t, args=()):
    """Reference evaluation in python (used by tests to check the device interpreter)."""
    import cmath
    st = []
    un = {I_NEG: lambda a: -a, I_CONJ: lambda a: a.conjugate(), I_SIN: cmath.sin,
          I_COS: cmath.cos, I_TAN: cmath.tan, I_EXP: cmath.exp, I_LOG: cmath.log,
          I_SQRT: cmath.sqrt, I_ABS: lambda a: complex(abs(a)),
          I_REAL: lambda a: complex(a.real), I_IMAG: lambda a: complex(a.imag),
          I_SINH: cmath.sinh, I_COSH: cmath.cosh, I_TANH: cmath.tanh,
          I_NORM2: lambda a: complex(abs(a) ** 2)}
    for op, ia, re, im in prog.instrs:
        if op == I_CONST:
            st.append(complex(re, im))
        elif op == I_T:
            st.append(complex(t))
        elif op == I_ARG:
            st.append(complex(args[ia]))
        elif op in (I_ADD, I_SUB, I_MUL, I_DIV, I_POW):
            b = st.pop(); a = st.pop()
            st.append({I_ADD: a + b, I_SUB: a - b, I_MUL: a * b,
                       I_DIV: a / b if op == I_DIV else 0, I_POW: a ** b if op == I_POW else 0}[op])
        elif op == I_MIN_RE:
            b = st.pop(); a = st.pop()
            st.append(complex(min(a.real, b.real)))
        elif op == I_SQRT_RE:
            st.append(complex(np.sqrt(st.pop().real)))
        elif op in un:
            st.append(complex(un[op](st.pop())))
        else:
            raise NotImplementedError(op)
    assert len(st) == 1
    return st[0]
