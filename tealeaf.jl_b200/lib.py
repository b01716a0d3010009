"""ctypes binding of libtealeaf_b200.so (include/tealeaf_b200.h).

This is the Python equivalent of the `ccall` stubs in julia/TeaLeafB200.jl.  It fails
loudly: a missing library raises ImportError, a missing GPU surfaces as TL_ERR_NO_DEVICE
from `tl_create`.  There is no CPU fallback anywhere on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TEALEAF_B200_LIB overrides the in-tree build (A/B runs of two builds, packaged installs)
LIB_PATH = os.environ.get("TEALEAF_B200_LIB") or os.path.join(HERE, "csrc", "libtealeaf_b200.so")

TL_OK = 0
TL_ERR_ARG, TL_ERR_CUDA, TL_ERR_NO_DEVICE, TL_ERR_EIGEN, TL_ERR_COMM, TL_ERR_STATE = -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "TL_OK", -1: "TL_ERR_ARG", -2: "TL_ERR_CUDA", -3: "TL_ERR_NO_DEVICE", -4: "TL_ERR_EIGEN",
                -5: "TL_ERR_COMM", -6: "TL_ERR_STATE"}


class TeaLeafError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class SolveInfo(C.Structure):
    """tl_solve_info"""
    _fields_ = [("iters", C.c_int), ("cg_iters", C.c_int), ("cheby_iters", C.c_int), ("est_iters", C.c_int),
                ("inner_total", C.c_int), ("halo_depth_k", C.c_int), ("error", C.c_double), ("eigmin", C.c_double),
                ("eigmax", C.c_double), ("solve_ms", C.c_double), ("kernel_launches", C.c_longlong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_DP = C.POINTER(C.c_double)

class PaintState(C.Structure):
    """tl_state in include/tealeaf_b200.h"""
    _fields_ = [("density", _D), ("energy", _D), ("xmin", _D), ("ymin", _D), ("xmax", _D), ("ymax", _D),
                ("radius", _D), ("geometry", _I), ("reserved", _I)]


# name -> (restype, argtypes); every symbol include/tealeaf_b200.h declares
SIGNATURES = {
    "tl_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I]),
    "tl_create_tile": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _I, _I, _I]),
    "tl_create_multi": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, C.POINTER(_I), _I, _I]),
    "tl_destroy": (None, [_P]),
    "tl_last_error": (C.c_char_p, [_P]),
    "tl_abi_version": (_I, []),
    "tl_set_option": (_I, [_P, C.c_char_p, _D]),
    "tl_get_option": (_I, [_P, C.c_char_p, _DP]),
    "tl_comm_blob_size": (_I, []),
    "tl_comm_export": (_I, [_P, _P]),
    "tl_comm_unique_id": (_I, [_P]),
    "tl_comm_connect": (_I, [_P, _P, _P]),
    "tl_set_field": (_I, [_P, _I, _P, C.c_long]),
    "tl_get_field": (_I, [_P, _I, _P, C.c_long]),
    "tl_copy_field": (_I, [_P, _I, _I]),
    "tl_paint_states": (_I, [_P, _I, C.POINTER(PaintState), _D, _D, _D, _D, _I, _I]),
    "tl_halo_update": (_I, [_P, C.c_uint, _I]),
    "tl_cg_init": (_I, [_P, _I, _D, _D, _DP]),
    "tl_cg_calc_w": (_I, [_P, _DP]),
    "tl_cg_calc_ur": (_I, [_P, _D, _DP]),
    "tl_cg_calc_p": (_I, [_P, _D]),
    "tl_copy_u": (_I, [_P]),
    "tl_calc_residual": (_I, [_P]),
    "tl_finalise": (_I, [_P]),
    "tl_solve_finished": (_I, [_P, _I]),
    "tl_norm2": (_I, [_P, _I, _DP]),
    "tl_cheby_init": (_I, [_P, _D, _DP]),
    "tl_cheby_iterate": (_I, [_P, _D, _D, _I, _DP]),
    "tl_ppcg_init_sd": (_I, [_P, _D]),
    "tl_ppcg_inner": (_I, [_P, _DP, _DP, _I]),
    "tl_jacobi_init": (_I, [_P, _I, _D, _D]),
    "tl_jacobi_iterate": (_I, [_P, _DP]),
    "tl_field_summary": (_I, [_P, _D, _DP, _DP, _DP, _DP]),
    "tl_cg_solve": (_I, [_P, _I, _D, _D, _D, _I, C.POINTER(SolveInfo), _DP, _DP]),
    "tl_cheby_solve": (_I, [_P, _I, _D, _D, _D, _I, _I, _D, _I, C.POINTER(SolveInfo)]),
    "tl_ppcg_solve": (_I, [_P, _I, _D, _D, _D, _I, _I, _D, _I, _I, _I, C.POINTER(SolveInfo)]),
    "tl_jacobi_solve": (_I, [_P, _I, _D, _D, _D, _I, C.POINTER(SolveInfo)]),
    "tl_time_kernel": (_I, [_P, C.c_char_p, _I, _DP]),
    "tl_timer_start": (_I, [_P]),
    "tl_timer_stop": (_I, [_P, _DP]),
    "tl_launch_count": (_I, [_P, C.POINTER(C.c_longlong)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc: int):
    if rc != TL_OK:
        msg = ""
        if ctx:
            raw = load().tl_last_error(ctx)
            msg = raw.decode() if raw else ""
        raise TeaLeafError(rc, msg)
