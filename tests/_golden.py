"""Helpers shared by the tests: load golden fixtures into oracle objects."""
import cmath
import math
import os

import numpy as np

from oracle.rk_oracle import OrcEvo, OrcOp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def op_arrays(g, prefix):
    kind = str(g[prefix + "_kind"])
    shape = tuple(int(x) for x in g[prefix + "_shape"])
    if kind == "csr":
        return kind, shape, dict(data=g[prefix + "_data"], col=g[prefix + "_col"],
                                 rowptr=g[prefix + "_rowptr"])
    if kind == "dia":
        return kind, shape, dict(data=g[prefix + "_data"], offsets=g[prefix + "_offsets"])
    return kind, shape, dict(arr=g[prefix + "_arr"])


def orc_op(g, prefix):
    kind, shape, a = op_arrays(g, prefix)
    if kind == "csr":
        return OrcOp.csr(a["data"], a["col"], a["rowptr"], shape)
    if kind == "dia":
        return OrcOp.dia(a["data"], a["offsets"], shape)
    return OrcOp.dense(a["arr"])


def parse_coeff(desc):
    """'' -> None (constant 1); 'str:expr|k=v,..' ; 'conj(...)' -> python callable."""
    desc = str(desc)
    if desc == "":
        return None
    if desc.startswith("conj(") and desc.endswith(")"):
        inner = parse_coeff(desc[5:-1])
        return lambda t: complex(inner(t)).conjugate()
    assert desc.startswith("str:")
    expr, _, kv = desc[4:].partition("|")
    env = {k: getattr(cmath, k) for k in ("sin", "cos", "exp", "sqrt", "tan", "log", "pi")}
    env["abs"] = abs
    for item in kv.split(","):
        if item:
            k, v = item.split("=")
            env[k] = complex(v.strip("()"))
    code = compile(expr, "<coeff>", "eval")
    return lambda t: complex(eval(code, {"__builtins__": {}}, dict(env, t=t)))


def coeff_spec(desc):
    """(expr, args-dict) for the product's coefficient compiler, or None."""
    desc = str(desc)
    if desc == "":
        return None
    conj = False
    if desc.startswith("conj("):
        conj = True
        desc = desc[5:-1]
    expr, _, kv = desc[4:].partition("|")
    args = {}
    for item in kv.split(","):
        if item:
            k, v = item.split("=")
            args[k] = complex(v.strip("()"))
    return ("conj(%s)" % expr if conj else expr), args


def orc_rhs(g):
    els = []
    for i in range(int(g["n_elements"])):
        cf = parse_coeff(g["el%d_coeff" % i])
        els.append((orc_op(g, "el%d" % i), 1.0 if cf is None else cf))
    return OrcEvo(els)
