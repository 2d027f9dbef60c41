"""ctypes binding of liblsx.so (the C-ABI declared in include/lsx.h).

There is no CPU fallback: if the shared library has not been built, importing this module
raises, and every operation needs a CUDA device (``lsx_create`` fails without one).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LSX_LIB_PATH") or os.path.join(_HERE, "liblsx.so")   # override: kernel experiments

if not os.path.exists(SO_PATH):
    raise ImportError(
        "linalg_solver_b200: %s is missing. Build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
        "(or `make -C linalg_solver_b200/csrc`). There is no CPU fallback." % SO_PATH
    )

lib = ctypes.CDLL(SO_PATH)

OK = 0
ERR_BAD_SHAPE, ERR_CUDA, ERR_BOUND, ERR_NULL, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6
ERR_NAMES = {-1: "BAD_SHAPE", -2: "CUDA", -3: "BOUND", -4: "NULL", -5: "UNSUPPORTED", -6: "NO_DEVICE"}
MEM_HOST, MEM_DEVICE = 0, 1
ST_SINGULAR, ST_INCONSISTENT, ST_BOUND, ST_NO_GOOD_PRIME, ST_GEN_TRUNC, ST_RETRIED = 1, 2, 4, 8, 16, 32
OP_RREF, OP_INVERSE, OP_DET, OP_RANK, OP_SOLVE = 1, 2, 3, 4, 5
ABI_VERSION = 1
TABLE_PRIMES = 2048


class Plan(ctypes.Structure):
    """Mirror of ``struct lsx_plan`` (include/lsx.h)."""
    _fields_ = [
        ("op", ctypes.c_int32), ("m", ctypes.c_int32), ("n", ctypes.c_int32), ("bar_col", ctypes.c_int32),
        ("max_rank", ctypes.c_int32), ("n_primes", ctypes.c_int32), ("limbs", ctypes.c_int32),
        ("pivot_slots", ctypes.c_int32), ("gen_cap", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("a_abs_max", ctypes.c_int64), ("b_abs_max", ctypes.c_int64), ("log2_bound", ctypes.c_double),
    ]

    def __repr__(self):
        return "Plan(op=%d, %dx%d bar=%d, primes=%d, limbs=%d, bound=2^%.1f)" % (
            self.op, self.m, self.n, self.bar_col, self.n_primes, self.limbs, self.log2_bound)


_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_pp = ctypes.POINTER(Plan)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against lsx.h
SIGNATURES = {
    "lsx_abi_version": (_i, []),
    "lsx_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "lsx_create_multi": (_i, [ctypes.POINTER(_i), _i, ctypes.POINTER(_vp)]),
    "lsx_device_count": (_i, [_vp]),
    "lsx_multi_uses_nccl": (_i, [_vp]),
    "lsx_destroy": (None, [_vp]),
    "lsx_last_error": (ctypes.c_char_p, [_vp]),
    "lsx_set_stream": (_i, [_vp, _vp]),
    "lsx_synchronize": (_i, [_vp]),
    "lsx_launch_count": (_i64, [_vp]),
    "lsx_last_prime_count": (_i, [_vp, ctypes.POINTER(_i)]),
    "lsx_timing_enable": (_i, [_vp, _i]),
    "lsx_timing_read": (_i, [_vp, _vp, _i, ctypes.POINTER(_i)]),
    "lsx_debug_set_primes": (_i, [_vp, _vp, _i]),
    "lsx_get_primes": (_i, [_vp, _vp, _i]),
    "lsx_plan_rref": (_i, [_i, _i, _i, _i64, _i64, _i, _pp]),
    "lsx_plan_inverse": (_i, [_i, _i64, _pp]),
    "lsx_plan_det": (_i, [_i, _i64, _pp]),
    "lsx_plan_rank": (_i, [_i, _i, _i64, _pp]),
    "lsx_plan_solve": (_i, [_i, _i, _i64, _i64, _i, _i, _pp]),
    "lsx_rref_batch": (_i, [_vp, _pp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "lsx_inverse_batch": (_i, [_vp, _pp, _vp, _i64, _i, _vp, _vp, _vp]),
    "lsx_inverse_batch_i8": (_i, [_vp, _pp, _vp, _i64, _i, _vp, _vp, _vp]),
    "lsx_det_batch": (_i, [_vp, _pp, _vp, _i64, _i, _vp, _vp, _vp]),
    "lsx_rank_batch": (_i, [_vp, _pp, _vp, _i64, _i, _vp, _vp]),
    "lsx_solve_batch": (_i, [_vp, _pp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lsx_lowest_terms": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "lsx_rref_trace_max_ops": (_i, [_i, _i, _i]),
    "lsx_rref_trace": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lsx_rref_trace_q": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lsx_det_large_prime_count": (_i, [_i, _i64, ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_double)]),
    "lsx_det_large_prime_count_for": (_i, [_vp, _vp, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_double)]),
    "lsx_det_large_residues": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "lsx_rank_large": (_i, [_vp, _vp, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "lsx_det_large": (_i, [_vp, _vp, _i, _i, _vp, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "lsx_crt_signed": (_i, [_vp, _vp, _i, _i, _i, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here means the .so is stale: rebuild
    _fn.restype = _res
    _fn.argtypes = _args

if lib.lsx_abi_version() != ABI_VERSION:
    raise ImportError("liblsx.so has ABI version %d, the Python layer expects %d: rebuild"
                      % (lib.lsx_abi_version(), ABI_VERSION))
