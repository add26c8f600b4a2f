"""ctypes binding of libfdm_b200.so (include/fdm_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible,
every solve raises.  ``FDM_B200_LIB`` may point at an alternative build of the library.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libfdm_b200.so")

ABI_VERSION = 3

FLAG_CALL = 0x01
FLAG_AMERICAN = 0x02
FLAG_BLACK76 = 0x04
FLAG_SNAP = 0x08
FLAG_KO_DOWN = 0x10
FLAG_KO_UP = 0x20
FLAG_BDRY_TOP_NEXT = 0x40
FLAG_PUT_BDRY_SMIN = 0x80
FLAG_INVALID = 0x100

STATUS_OK, STATUS_NONFINITE, STATUS_STIFF, STATUS_SINGULAR, STATUS_NOT_SOLVED = 0, 1, 2, 3, -1
ROWVAR_MAX_NODES = 1024

MAX_N_SPACE = 8193
WARP_MAX_N_SPACE = 513

ERR_NAMES = {0: "FDM_OK", -1: "FDM_ERR_INVALID", -2: "FDM_ERR_CUDA", -3: "FDM_ERR_UNSUPPORTED",
             -4: "FDM_ERR_NO_DEVICE"}

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class FdmBatch(C.Structure):
    _fields_ = [
        ("n_solves", C.c_int64), ("n_space_max", C.c_int32), ("n_seg_total", C.c_int32),
        ("n_mon_total", C.c_int32), ("reserved", C.c_int32),
        ("x_min", C.c_void_p), ("dx", C.c_void_p), ("sigma", C.c_void_p), ("r", C.c_void_p),
        ("b", C.c_void_p), ("spot", C.c_void_p), ("strike", C.c_void_p), ("barrier", C.c_void_p),
        ("rebate", C.c_void_p), ("n_space", C.c_void_p), ("flags", C.c_void_p),
        ("rannacher", C.c_void_p), ("seg_begin", C.c_void_p), ("mon_begin", C.c_void_p),
        ("seg_steps", C.c_void_p), ("seg_dt", C.c_void_p), ("seg_tau0", C.c_void_p),
        ("seg_restart", C.c_void_p), ("seg_div", C.c_void_p), ("mon_step", C.c_void_p),
        ("b_bdry", C.c_void_p), ("barrier_up", C.c_void_p), ("mon_rebate", C.c_void_p),
        ("mon_end", C.c_void_p),
    ]


class FdmOut(C.Structure):
    _fields_ = [
        ("price", C.c_void_p), ("delta", C.c_void_p), ("gamma", C.c_void_p),
        ("stencil_v", C.c_void_p), ("stencil_s", C.c_void_p), ("idx", C.c_void_p),
        ("status", C.c_void_p), ("v_dump", C.c_void_p), ("v_dump_stride", C.c_int64),
        ("ex_trace", C.c_void_p), ("ex_trace_stride", C.c_int64),
    ]


class FdmAmericanBook(C.Structure):
    _fields_ = [
        ("n_options", C.c_int64),
        ("spot", C.c_void_p), ("strike", C.c_void_p), ("T", C.c_void_p), ("sigma", C.c_void_p),
        ("r", C.c_void_p), ("b", C.c_void_p), ("flags", C.c_void_p), ("n_space", C.c_void_p),
        ("n_time", C.c_void_p), ("rannacher", C.c_void_p),
        ("s_max_mult", C.c_double), ("dv_sigma", C.c_double), ("what", C.c_int32), ("use_richardson", C.c_int32),
        ("div_begin", C.c_void_p), ("div_tau", C.c_void_p), ("div_cash", C.c_void_p), ("n_div_total", C.c_int64),
        ("n_space_max", C.c_int32), ("hints", C.c_int32),
    ]


class FdmBookOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("price_log2", "price", "delta", "gamma", "vega", "theta",
                                          "ex_lo", "ex_hi", "status", "out_index")]


class FdmBarrierBook(C.Structure):
    _fields_ = [
        ("n_options", C.c_int64),
        ("S0", C.c_void_p), ("K", C.c_void_p), ("T", C.c_void_p), ("sigma", C.c_void_p), ("r_disc", C.c_void_p),
        ("b_carry", C.c_void_p), ("x_min", C.c_void_p), ("dx", C.c_void_p), ("barrier", C.c_void_p),
        ("barrier_up", C.c_void_p), ("rebate", C.c_void_p), ("flags", C.c_void_p), ("n_space", C.c_void_p),
        ("n_time", C.c_void_p), ("mon_begin", C.c_void_p), ("mon_step", C.c_void_p), ("n_mon_total", C.c_int64),
        ("dv_sigma", C.c_double), ("what", C.c_int32), ("n_space_max", C.c_int32), ("hints", C.c_int32),
        ("reserved", C.c_int32),
    ]


class FdmRowVarBatch(C.Structure):
    _fields_ = [("n_solves", C.c_int64), ("n_nodes_max", C.c_int32), ("n_mon_total", C.c_int32)] + [
        (n, C.c_void_p) for n in ("node_begin", "s_nodes", "v_terminal", "sigma", "r", "strike", "tenor", "dt", "lo_bar",
                                  "up_bar", "h_minus", "h_plus", "s_eff", "n_steps", "rannacher", "flags", "bar_j",
                                  "mon_begin", "mon_step")]


class FdmRowVarOut(C.Structure):
    _fields_ = [("price", C.c_void_p), ("status", C.c_void_p), ("v_dump", C.c_void_p), ("v_dump_stride", C.c_int64)]


BOOK_PRICE_LOG, BOOK_PRICE_LOG2, BOOK_GREEKS_LOG2, BOOK_SCENARIO = 0, 1, 2, 3
BARRIER_BOOK_SOLVE, BARRIER_BOOK_GREEKS = 0, 1

EXPORTS = ("fdm_version", "fdm_sizeof", "fdm_last_error", "fdm_device_count", "fdm_solve_batch",
           "fdm_solve_batch_host", "fdm_launch_hints", "fdm_book_solves_per_option", "fdm_american_book_host",
           "fdm_american_book_workspace", "fdm_american_book", "fdm_barrier_book_solves_per_option",
           "fdm_barrier_book_workspace", "fdm_barrier_book", "fdm_barrier_book_host", "fdm_rowvar_solve",
           "fdm_rowvar_solve_host", "fdm_release_workspace",
           "fdm_fp64_peak")


class FdmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None
_lib_path = None


def lib_path() -> str:
    return os.environ.get("FDM_B200_LIB", DEFAULT_LIB)


def load(path: str | None = None):
    """Load (once) and return the CDLL.  Raises if the CUDA extension has not been built."""
    global _lib, _lib_path
    path = path or lib_path()
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise FdmError(-4, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  finite_difference_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.fdm_version.restype = C.c_int
    L.fdm_sizeof.restype = C.c_int
    L.fdm_sizeof.argtypes = [C.c_int]
    L.fdm_last_error.restype = C.c_char_p
    L.fdm_device_count.restype = C.c_int
    L.fdm_solve_batch.restype = C.c_int
    L.fdm_solve_batch.argtypes = [C.POINTER(FdmBatch), C.POINTER(FdmOut), C.c_void_p, _pi]
    L.fdm_solve_batch_host.restype = C.c_int
    L.fdm_solve_batch_host.argtypes = [C.POINTER(FdmBatch), C.POINTER(FdmOut), _pi, C.POINTER(C.c_float)]
    L.fdm_launch_hints.restype = C.c_int
    L.fdm_launch_hints.argtypes = [_pi, _pi, C.c_int64, _pd, C.c_int64]
    L.fdm_book_solves_per_option.restype = C.c_int
    L.fdm_book_solves_per_option.argtypes = [C.c_int32, C.c_int32]
    L.fdm_american_book_host.restype = C.c_int
    L.fdm_american_book_host.argtypes = [C.POINTER(FdmAmericanBook), C.POINTER(FdmBookOut), C.POINTER(C.c_int64),
                                         _pi, C.POINTER(C.c_float)]
    L.fdm_american_book_workspace.restype = C.c_size_t
    L.fdm_american_book_workspace.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    L.fdm_american_book.restype = C.c_int
    L.fdm_american_book.argtypes = [C.POINTER(FdmAmericanBook), C.POINTER(FdmBookOut), C.c_void_p, C.c_size_t,
                                    C.c_void_p, _pi]
    L.fdm_barrier_book_solves_per_option.restype = C.c_int
    L.fdm_barrier_book_solves_per_option.argtypes = [C.c_int32]
    L.fdm_barrier_book_workspace.restype = C.c_size_t
    L.fdm_barrier_book_workspace.argtypes = [C.c_int64, C.c_int32]
    L.fdm_barrier_book.restype = C.c_int
    L.fdm_barrier_book.argtypes = [C.POINTER(FdmBarrierBook), C.POINTER(FdmBookOut), C.c_void_p, C.c_size_t,
                                   C.c_void_p, _pi]
    L.fdm_barrier_book_host.restype = C.c_int
    L.fdm_barrier_book_host.argtypes = [C.POINTER(FdmBarrierBook), C.POINTER(FdmBookOut), C.POINTER(C.c_int64),
                                        _pi, C.POINTER(C.c_float)]
    L.fdm_rowvar_solve.restype = C.c_int
    L.fdm_rowvar_solve.argtypes = [C.POINTER(FdmRowVarBatch), C.POINTER(FdmRowVarOut), C.c_void_p, _pi]
    L.fdm_rowvar_solve_host.restype = C.c_int
    L.fdm_rowvar_solve_host.argtypes = [C.POINTER(FdmRowVarBatch), C.POINTER(FdmRowVarOut), _pi, C.POINTER(C.c_float)]
    L.fdm_fp64_peak.restype = C.c_int
    L.fdm_fp64_peak.argtypes = [_pd, C.POINTER(C.c_float)]
    if L.fdm_version() != ABI_VERSION:
        raise FdmError(-1, f"ABI mismatch: library {L.fdm_version()} vs binding {ABI_VERSION}")
    _lib, _lib_path = L, path
    return L


def check(rc: int):
    if rc != 0:
        raise FdmError(rc, load().fdm_last_error().decode("utf-8", "replace"))


def fp64_peak_tflops() -> tuple[float, float]:
    """(TFLOP/s, ms) of the register-resident DFMA micro-benchmark on the current device."""
    t = C.c_double()
    ms = C.c_float()
    check(load().fdm_fp64_peak(C.byref(t), C.byref(ms)))
    return t.value, ms.value
