"""ctypes binding of libqutip_b200.so (C ABI declared in include/qutip_b200.h).

There is deliberately no fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or
``qutip_b200/csrc/build.sh``) importing any compute path raises ``ImportError``, and
calling it without a CUDA device raises ``QbError`` from the first CUDA call.
"""
import ctypes as C
import os

import numpy as np

from .coeffs import QbInstr

_HERE = os.path.dirname(os.path.abspath(__file__))
# QUTIP_B200_LIB selects another build of the same library (A/B kernel tuning only)
LIB_PATH = os.environ.get("QUTIP_B200_LIB") or os.path.join(_HERE, "libqutip_b200.so")


class QbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qutip_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class QbOptions(C.Structure):
    _fields_ = [("atol", C.c_double), ("rtol", C.c_double), ("nsteps", C.c_int),
                ("first_step", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
                ("interpolate", C.c_int), ("norm_steps", C.c_int), ("norm_t_tol", C.c_double),
                ("norm_tol", C.c_double), ("norm_min_step", C.c_double),
                ("mc_corr_eps", C.c_double), ("store_states", C.c_int),
                ("max_collapses", C.c_int), ("no_jump", C.c_int),
                ("jump_prob_floor", C.c_double), ("max_order", C.c_int),
                ("pad_", C.c_int)]


# every symbol include/qutip_b200.h declares (tests check the library exports them all)
SYMBOLS = [
    "qb_version", "qb_last_error", "qb_device_count", "qb_set_device", "qb_synchronize",
    "qb_launch_count", "qb_device_mem_info",
    "qb_dense_upload", "qb_dense_zeros", "qb_dense_download", "qb_dense_write", "qb_dense_copy", "qb_dense_reshape", "qb_dense_info",
    "qb_csr_upload", "qb_dia_upload", "qb_kron_upload", "qb_sandwich_upload", "qb_liouvillian_build", "qb_op_csr_download", "qb_op_convert", "qb_kron_build",
    "qb_op_info", "qb_free",
    "qb_matmul", "qb_axpy", "qb_scal", "qb_copy", "qb_zero", "qb_nrm2", "qb_wrms_error",
    "qb_inner", "qb_expect_ket", "qb_expect_dm", "qb_expect_super", "qb_trace_oper_ket",
    "qb_system_create", "qb_system_add_element", "qb_system_add_collapse", "qb_system_add_eop",
    "qb_system_set_eop_functional", "qb_system_set_mc_trace", "qb_system_add_spline", "qb_options_default",
    "qb_engine_create", "qb_engine_run", "qb_engine_run_device", "qb_reduce_expect",
    "qb_engine_last_run_info", "qb_integ_set_state", "qb_integ_integrate",
    "qb_integ_get_state", "qb_integ_set_args", "qb_integ_stats", "qb_engine_rhs",
    "qb_engine_rhs_bench", "qb_engine_set_profiling", "qb_engine_profile", "qb_engine_profile_rounds",
    "qb_zgemm", "qb_zgemm_bench", "qb_dmma_peak_bench", "qb_integ_pending_coef", "qb_integ_resume",
    "qb_engine_rhs_coef",
    "qb_comm_nccl_version", "qb_comm_init_all", "qb_comm_unique_id", "qb_comm_init_rank",
    "qb_comm_info", "qb_comm_allreduce_sum", "qb_comm_allreduce_sum_device", "qb_comm_reduce_expect",
]

_lib = None
_load_pid = None


def load():
    """Load the shared library (no CUDA call is made by loading)."""
    global _lib, _load_pid
    if _lib is not None:
        return _lib
    _load_pid = os.getpid()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "qutip_b200: %s is missing. Build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` (needs nvcc). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.qb_last_error.restype = C.c_char_p
    lib.qb_launch_count.restype = C.c_int64
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    pp = C.POINTER(C.c_void_p)
    sig = {
        "qb_dense_upload": [vp, i64, i64, i32, pp],
        "qb_dense_zeros": [i64, i64, i32, pp],
        "qb_dense_download": [vp, vp],
        "qb_dense_write": [vp, vp],
        "qb_dense_copy": [vp, pp],
        "qb_dense_reshape": [vp, i64, i64],
        "qb_dense_info": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), pp],
        "qb_csr_upload": [vp, vp, vp, i64, i64, i64, i32, pp],
        "qb_dia_upload": [vp, vp, i64, i64, i64, i32, pp],
        "qb_kron_upload": [vp, vp, vp, i64, i64, i32, pp],
        "qb_sandwich_upload": [vp, vp, vp, i64, i64, i64, pp],
        "qb_liouvillian_build": [vp, vp, vp, i64, vp, vp, vp, i64, i64, i64, dbl, i32, pp],
        "qb_op_csr_download": [vp, vp, vp, vp],
        "qb_op_convert": [vp, i32, pp],
        "qb_kron_build": [vp, vp, vp, i64, i64, vp, vp, vp, i64, i64, i32, pp],
        "qb_op_info": [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                       C.POINTER(i64)],
        "qb_free": [vp],
        "qb_matmul": [vp, vp, dbl, dbl, vp],
        "qb_axpy": [vp, dbl, dbl, vp],
        "qb_zgemm": [vp, vp, dbl, dbl, vp],
        "qb_zgemm_bench": [vp, vp, vp, i32, C.POINTER(dbl)],
        "qb_dmma_peak_bench": [i32, C.POINTER(dbl)],
        "qb_scal": [vp, dbl, dbl],
        "qb_copy": [vp, vp],
        "qb_zero": [vp],
        "qb_nrm2": [vp, C.POINTER(dbl)],
        "qb_wrms_error": [vp, vp, dbl, dbl, C.POINTER(dbl)],
        "qb_inner": [vp, vp, i32, C.POINTER(dbl)],
        "qb_expect_ket": [vp, vp, C.POINTER(dbl)],
        "qb_expect_dm": [vp, vp, C.POINTER(dbl)],
        "qb_expect_super": [vp, vp, C.POINTER(dbl)],
        "qb_trace_oper_ket": [vp, C.POINTER(dbl)],
        "qb_system_create": [i64, i32, pp],
        "qb_system_add_element": [vp, vp, C.POINTER(QbInstr), i32],
        "qb_system_add_collapse": [vp, vp, C.POINTER(QbInstr), i32, vp, C.POINTER(QbInstr), i32],
        "qb_system_add_eop": [vp, vp, C.POINTER(QbInstr), i32],
        "qb_system_set_eop_functional": [vp, i32],
        "qb_system_set_mc_trace": [vp, i32],
        "qb_system_add_spline": [vp, vp, vp, i32, i32, dbl, C.POINTER(i32)],
        "qb_options_default": [C.POINTER(QbOptions)],
        "qb_engine_create": [vp, i32, i32, C.POINTER(QbOptions), pp],
        "qb_engine_run": [vp, i32, i64, vp, i64, vp, vp, i32, vp, vp, i32,
                          vp, vp, vp, vp, vp, vp, vp, vp],
        "qb_engine_run_device": [vp, i32, i64, vp, i64, vp, vp, i32, vp, vp, i32,
                                 vp, vp, vp, vp, vp, vp, vp],
        "qb_reduce_expect": [vp, i64, i32, i32, vp],
        "qb_engine_last_run_info": [vp, C.POINTER(i64), C.POINTER(dbl)],
        "qb_integ_set_state": [vp, dbl, vp],
        "qb_integ_integrate": [vp, dbl, i32, C.POINTER(dbl), C.POINTER(i32)],
        "qb_integ_get_state": [vp, C.POINTER(dbl), vp],
        "qb_integ_pending_coef": [vp, C.POINTER(dbl)],
        "qb_integ_resume": [vp, vp, C.POINTER(dbl), C.POINTER(i32)],
        "qb_integ_set_args": [vp, vp],
        "qb_integ_stats": [vp, C.POINTER(i64)],
        "qb_engine_rhs": [vp, dbl, vp, vp],
        "qb_engine_rhs_coef": [vp, vp, vp, vp],
        "qb_engine_rhs_bench": [vp, dbl, vp, vp, i32, C.POINTER(dbl)],
        "qb_engine_set_profiling": [vp, i32],
        "qb_engine_profile": [vp, C.POINTER(dbl), C.POINTER(i64), C.POINTER(dbl)],
        "qb_engine_profile_rounds": [vp, vp, vp, i64, C.POINTER(i64)],
        "qb_device_count": [C.POINTER(i32)],
        "qb_set_device": [i32],
        "qb_device_mem_info": [C.POINTER(i64), C.POINTER(i64)],
        "qb_comm_nccl_version": [C.POINTER(i32)],
        "qb_comm_init_all": [i32, C.POINTER(i32), pp],
        "qb_comm_unique_id": [vp, i32],
        "qb_comm_init_rank": [i32, i32, vp, pp],
        "qb_comm_info": [vp, C.POINTER(i32), C.POINTER(i32)],
        "qb_comm_allreduce_sum": [vp, pp, i64],
        "qb_comm_allreduce_sum_device": [vp, pp, i64],
        "qb_comm_reduce_expect": [vp, pp, i32, i32, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().qb_last_error().decode("utf-8", "replace")
        if _load_pid is not None and os.getpid() != _load_pid and "initialization error" in msg:
            msg += (" -- this process was forked after CUDA had been initialised in its parent and CUDA "
                    "contexts do not survive fork(): use the 'spawn' start method for process maps, or "
                    "options={'map': 'serial'} / {'map': 'b200'}")
        raise QbError(rc, msg)


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int(0)
    rc = load().qb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def launch_count():
    return int(load().qb_launch_count())


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)
