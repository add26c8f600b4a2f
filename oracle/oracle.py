"""ctypes binding of the CPU oracle (oracle/fdm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

The date -> tau mapping of dividends (``fd_american_option_pricer.py:304-315``) and the
pricer-level compositions stay in Python here; the PDE solve itself
(``_solve_grid``, ``fd_american_option_pricer.py:554-602``;
``discrete_barrier_fdm_pricer_cn.py:219-302``) is the C restatement.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfdm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "fdm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


class _FdoAmerican(C.Structure):
    _fields_ = [
        ("spot", C.c_double), ("strike", C.c_double), ("T", C.c_double), ("sigma", C.c_double),
        ("r", C.c_double), ("b", C.c_double), ("s_max_mult", C.c_double),
        ("is_call", C.c_int32), ("model", C.c_int32), ("n_space", C.c_int32),
        ("n_time", C.c_int32), ("rannacher", C.c_int32), ("n_div", C.c_int32),
        ("div_tau", C.POINTER(C.c_double)), ("div_cash", C.POINTER(C.c_double)),
    ]


class _FdoAmericanInfo(C.Structure):
    _fields_ = [
        ("spot_idx", C.c_int32), ("strike_idx", C.c_int32), ("n_steps_total", C.c_int32),
        ("ex_lo", C.c_int32), ("ex_hi", C.c_int32),
        ("dx", C.c_double), ("S_min", C.c_double), ("S_max", C.c_double), ("price", C.c_double),
        ("seg_steps", C.c_int32 * 16), ("n_seg", C.c_int32),
    ]


class _FdoBarrier(C.Structure):
    _fields_ = [
        ("S0", C.c_double), ("K", C.c_double), ("T", C.c_double), ("sigma", C.c_double),
        ("r_disc", C.c_double), ("b_carry", C.c_double),
        ("lower_barrier", C.c_double), ("upper_barrier", C.c_double), ("rebate", C.c_double),
        ("is_call", C.c_int32), ("ko_kind", C.c_int32), ("apply_KO", C.c_int32),
        ("N_space", C.c_int32), ("N_time", C.c_int32), ("n_monitor", C.c_int32),
        ("monitor_times", C.POINTER(C.c_double)),
    ]


class _FdoBarrierInfo(C.Structure):
    _fields_ = [
        ("S_min", C.c_double), ("S_max", C.c_double), ("dx", C.c_double),
        ("price", C.c_double), ("delta", C.c_double), ("gamma", C.c_double),
        ("n_monitor_steps", C.c_int32), ("greek_idx", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.fdo_american_solve.restype = C.c_int
        L.fdo_american_solve.argtypes = [C.POINTER(_FdoAmerican), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                         C.POINTER(_FdoAmericanInfo)]
        L.fdo_barrier_configure.restype = C.c_int
        L.fdo_barrier_configure.argtypes = [C.POINTER(_FdoBarrier), C.c_int32, C.c_int32,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.fdo_barrier_solve.restype = C.c_int
        L.fdo_barrier_solve.argtypes = [C.POINTER(_FdoBarrier), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                        C.POINTER(_FdoBarrierInfo)]
        L.fdo_barrier_solve_ext.restype = C.c_int
        L.fdo_barrier_solve_ext.argtypes = [C.POINTER(_FdoBarrier), C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                            C.POINTER(_FdoBarrierInfo)]
        L.fdo_american_batch.restype = C.c_int
        L.fdo_american_batch.argtypes = [C.c_int64, C.POINTER(_FdoAmerican), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
        L.fdo_barrier_batch.restype = C.c_int
        L.fdo_barrier_batch.argtypes = [C.c_int64, C.POINTER(_FdoBarrier), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32]
        L.fdo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# --------------------------------------------------------------------------------------
# American
# --------------------------------------------------------------------------------------

@dataclass
class AmericanResult:
    V: np.ndarray
    s_nodes: np.ndarray
    price: float
    spot_idx: int
    strike_idx: int
    ex_lo: int
    ex_hi: int
    dx: float
    S_min: float
    S_max: float
    seg_steps: List[int]
    ex_trace: Optional[np.ndarray]


def _am_struct(spot, strike, T, sigma, r, b, is_call, model, n_space, n_time, rannacher,
               s_max_mult, div_tau, div_cash):
    p = _FdoAmerican()
    p.spot, p.strike, p.T, p.sigma, p.r, p.b = spot, strike, T, sigma, r, b
    p.s_max_mult = s_max_mult
    p.is_call, p.model = int(bool(is_call)), int(model)
    p.n_space, p.n_time, p.rannacher = int(n_space), int(n_time), int(rannacher)
    p.n_div = len(div_tau)
    p.div_tau = _dptr(div_tau) if len(div_tau) else None
    p.div_cash = _dptr(div_cash) if len(div_cash) else None
    return p


def american_solve(spot: float, strike: float, T: float, sigma: float, r: float, b: float,
                   is_call: bool, model: int = 0, n_space: int = 400, n_time: int = 400,
                   rannacher: int = 2, s_max_mult: float = 4.5,
                   divs_tau: Sequence[Tuple[float, float]] = (),
                   trace: bool = False) -> AmericanResult:
    """One ``_solve_grid`` (fd_american_option_pricer.py:554-602) + ``_interp_price``.

    ``divs_tau`` = [(tau_div, cash)] ascending in tau, as ``_div_times_tau`` returns.
    ``model``: 0 equity, 1 Black-76 (fd_american_black76.py).
    """
    div_tau = np.ascontiguousarray([d[0] for d in divs_tau], dtype=np.float64)
    div_cash = np.ascontiguousarray([d[1] for d in divs_tau], dtype=np.float64)
    p = _am_struct(spot, strike, T, sigma, r, b, is_call, model, n_space, n_time, rannacher,
                   s_max_mult, div_tau, div_cash)
    V = np.empty(n_space + 1, dtype=np.float64)
    s = np.empty(n_space + 1, dtype=np.float64)
    max_steps = n_time + 4 * (len(divs_tau) + 1) + 8
    ex = np.full(2 * max_steps, -2, dtype=np.int32) if trace else None
    info = _FdoAmericanInfo()
    rc = lib().fdo_american_solve(C.byref(p), _dptr(V), _dptr(s),
                                  ex.ctypes.data_as(C.POINTER(C.c_int32)) if trace else None,
                                  C.byref(info))
    if rc != 0:
        raise ValueError(f"fdo_american_solve failed rc={rc}")
    return AmericanResult(
        V=V, s_nodes=s, price=info.price, spot_idx=info.spot_idx, strike_idx=info.strike_idx,
        ex_lo=info.ex_lo, ex_hi=info.ex_hi, dx=info.dx, S_min=info.S_min, S_max=info.S_max,
        seg_steps=[info.seg_steps[i] for i in range(min(info.n_seg, 16))],
        ex_trace=ex[: 2 * info.n_steps_total].reshape(-1, 2) if trace else None,
    )


def local_cubic_delta_gamma(s: np.ndarray, V: np.ndarray, Sx: float) -> Tuple[float, float]:
    """``_local_cubic_delta_gamma`` (fd_american_option_pricer.py:632-653)."""
    n = len(s) - 1
    i_near = int(np.argmin(np.abs(s - Sx)))          # first minimum, like min(range, key=)
    i_near = min(max(i_near, 1), n - 2)
    idx = [i_near - 1, i_near, i_near + 1, i_near + 2]
    x = np.array([s[j] for j in idx], dtype=float)
    y = np.array([V[j] for j in idx], dtype=float)
    z = x - Sx
    M = np.vstack([z ** 3, z ** 2, z, np.ones_like(z)]).T
    a, b, c, d = np.linalg.solve(M, y)
    return float(c), float(2.0 * b)


def american_batch_prices(params: Sequence[dict], n_threads: int = 0):
    """OpenMP batch of American solves (CPU baseline).  Each dict = american_solve kwargs."""
    n = len(params)
    arr = (_FdoAmerican * n)()
    keep = []
    for i, kw in enumerate(params):
        divs = kw.get("divs_tau", ())
        div_tau = np.ascontiguousarray([d[0] for d in divs], dtype=np.float64)
        div_cash = np.ascontiguousarray([d[1] for d in divs], dtype=np.float64)
        keep.append((div_tau, div_cash))
        arr[i] = _am_struct(kw["spot"], kw["strike"], kw["T"], kw["sigma"], kw["r"], kw["b"],
                            kw["is_call"], kw.get("model", 0), kw.get("n_space", 400),
                            kw.get("n_time", 400), kw.get("rannacher", 2),
                            kw.get("s_max_mult", 4.5), div_tau, div_cash)
    price = np.empty(n, dtype=np.float64)
    lo = np.empty(n, dtype=np.int32)
    hi = np.empty(n, dtype=np.int32)
    rc = lib().fdo_american_batch(n, arr, _dptr(price), lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                  hi.ctypes.data_as(C.POINTER(C.c_int32)), int(n_threads))
    if rc != 0:
        raise ValueError("fdo_american_batch: at least one solve failed")
    return price, lo, hi


def _struct_dtype(struct) -> np.dtype:
    """numpy structured dtype with the exact layout of a ctypes Structure (pointers as u8)."""
    names, formats, offsets = [], [], []
    for name, typ in struct._fields_:
        f = getattr(struct, name)
        names.append(name)
        offsets.append(f.offset)
        formats.append({C.c_double: "f8", C.c_int32: "i4"}.get(typ, "u8"))
    return np.dtype(dict(names=names, formats=formats, offsets=offsets, itemsize=C.sizeof(struct)))


def american_batch_prices_arrays(spot, strike, T, sigma, r, b, is_call, model, n_space, n_time,
                                 rannacher=2, s_max_mult=4.5, n_threads: int = 0):
    """OpenMP batch of dividend-free American solves from broadcastable arrays (CPU baseline)."""
    arrs = np.broadcast_arrays(*[np.asarray(v) for v in (spot, strike, T, sigma, r, b, is_call, model,
                                                         n_space, n_time, rannacher)])
    n = len(arrs[0])
    rec = np.zeros(n, dtype=_struct_dtype(_FdoAmerican))
    for name, a in zip(("spot", "strike", "T", "sigma", "r", "b", "is_call", "model", "n_space", "n_time",
                        "rannacher"), arrs):
        rec[name] = a
    rec["s_max_mult"] = s_max_mult
    price = np.empty(n, dtype=np.float64)
    lo = np.empty(n, dtype=np.int32)
    hi = np.empty(n, dtype=np.int32)
    rc = lib().fdo_american_batch(n, C.cast(rec.ctypes.data, C.POINTER(_FdoAmerican)), _dptr(price),
                                  lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                  hi.ctypes.data_as(C.POINTER(C.c_int32)), int(n_threads))
    if rc != 0:
        raise ValueError("fdo_american_batch: at least one solve failed")
    return price, lo, hi


# --------------------------------------------------------------------------------------
# Barrier
# --------------------------------------------------------------------------------------

@dataclass
class BarrierResult:
    V: np.ndarray
    s_nodes: np.ndarray
    price: float
    delta: float
    gamma: float
    S_min: float
    S_max: float
    dx: float
    monitor_steps: List[int]
    greek_idx: int
    N_space: int
    N_time: int


_KO = {"none": 0, "down-and-out": 1, "up-and-out": 2, "double-out": 3}


def _bar_struct(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type, lower_barrier,
                upper_barrier, rebate, mon, N_space, N_time, apply_KO):
    p = _FdoBarrier()
    p.S0, p.K, p.T, p.sigma, p.r_disc, p.b_carry = S0, K, T, sigma, r_disc, b_carry
    p.lower_barrier = float("nan") if lower_barrier is None else float(lower_barrier)
    p.upper_barrier = float("nan") if upper_barrier is None else float(upper_barrier)
    p.rebate = rebate
    p.is_call = int(bool(is_call))
    p.ko_kind = _KO.get(barrier_type, 0)
    p.apply_KO = int(bool(apply_KO))
    p.N_space = int(N_space or 0)
    p.N_time = int(N_time or 0)
    p.n_monitor = len(mon)
    p.monitor_times = _dptr(mon) if len(mon) else None
    return p


def barrier_configure(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type,
                      lower_barrier=None, upper_barrier=None, rebate=0.0, monitor_times=None,
                      N_space=None, N_time=None):
    """``configure_grid`` (discrete_barrier_fdm_pricer_cn.py:59-118) -> (S_min,S_max,N_space,N_time)."""
    mon = np.ascontiguousarray(monitor_times or [], dtype=np.float64)
    p = _bar_struct(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type, lower_barrier,
                    upper_barrier, rebate, mon, N_space, N_time, True)
    smin, smax = C.c_double(), C.c_double()
    ns, nt = C.c_int32(), C.c_int32()
    rc = lib().fdo_barrier_configure(C.byref(p), int(N_space or 0), int(N_time or 0),
                                     C.byref(smin), C.byref(smax), C.byref(ns), C.byref(nt))
    if rc != 0:
        raise ValueError("T, sigma and S0 must be positive")
    return smin.value, smax.value, ns.value, nt.value


def barrier_solve(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type,
                  lower_barrier=None, upper_barrier=None, rebate=0.0, monitor_times=None,
                  N_space=None, N_time=None, apply_KO=True, rannacher=0, rebate_discounted=False) -> BarrierResult:
    """One barrier ``_solve_grid(apply_KO)`` + price/delta/gamma from the grid.

    ``rannacher`` / ``rebate_discounted`` / ``barrier_type="double-out"`` are the engine options of SURVEY.md 8f-2
    (``fdo_barrier_solve_ext``; no runnable reference: parity unpinned)."""
    _, _, ns, nt = barrier_configure(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type,
                                     lower_barrier, upper_barrier, rebate, monitor_times,
                                     N_space, N_time)
    mon = np.ascontiguousarray(monitor_times or [], dtype=np.float64)
    p = _bar_struct(S0, K, T, sigma, r_disc, b_carry, is_call, barrier_type, lower_barrier,
                    upper_barrier, rebate, mon, ns, nt, apply_KO)
    V = np.empty(ns + 1, dtype=np.float64)
    s = np.empty(ns + 1, dtype=np.float64)
    flags = np.zeros(nt + 1, dtype=np.int32)
    info = _FdoBarrierInfo()
    if rannacher or rebate_discounted or barrier_type == "double-out":
        rc = lib().fdo_barrier_solve_ext(C.byref(p), C.c_int32(int(rannacher)), C.c_int32(int(bool(rebate_discounted))),
                                         _dptr(V), _dptr(s), flags.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(info))
    else:
        rc = lib().fdo_barrier_solve(C.byref(p), _dptr(V), _dptr(s),
                                     flags.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(info))
    if rc != 0:
        raise ValueError(f"fdo_barrier_solve failed rc={rc}")
    return BarrierResult(V=V, s_nodes=s, price=info.price, delta=info.delta, gamma=info.gamma,
                         S_min=info.S_min, S_max=info.S_max, dx=info.dx,
                         monitor_steps=[int(k) for k in np.nonzero(flags)[0]],
                         greek_idx=info.greek_idx, N_space=ns, N_time=nt)


def barrier_batch_prices(params: Sequence[dict], n_threads: int = 0):
    """OpenMP batch of barrier solves (CPU baseline).  Each dict = barrier_solve kwargs with
    N_space / N_time already fixed."""
    n = len(params)
    arr = (_FdoBarrier * n)()
    keep = []
    for i, kw in enumerate(params):
        mon = np.ascontiguousarray(kw.get("monitor_times") or [], dtype=np.float64)
        keep.append(mon)
        arr[i] = _bar_struct(kw["S0"], kw["K"], kw["T"], kw["sigma"], kw["r_disc"], kw["b_carry"],
                             kw["is_call"], kw["barrier_type"], kw.get("lower_barrier"),
                             kw.get("upper_barrier"), kw.get("rebate", 0.0), mon,
                             kw["N_space"], kw["N_time"], kw.get("apply_KO", True))
    price = np.empty(n, dtype=np.float64)
    delta = np.empty(n, dtype=np.float64)
    gamma = np.empty(n, dtype=np.float64)
    rc = lib().fdo_barrier_batch(n, arr, _dptr(price), _dptr(delta), _dptr(gamma), int(n_threads))
    if rc != 0:
        raise ValueError("fdo_barrier_batch: at least one solve failed")
    return price, delta, gamma


class _FdoDbfp(C.Structure):
    _fields_ = [("S_min", C.c_double), ("S_max", C.c_double), ("S0_eff", C.c_double), ("S_greek", C.c_double),
                ("K", C.c_double), ("T", C.c_double), ("sigma", C.c_double),
                ("r", C.c_double), ("b", C.c_double), ("q", C.c_double),
                ("lower_barrier", C.c_double), ("upper_barrier", C.c_double), ("rebate_amount", C.c_double),
                ("rebate_at_hit", C.c_int32), ("is_call", C.c_int32), ("ko_kind", C.c_int32),
                ("apply_KO", C.c_int32), ("num_space_nodes", C.c_int32), ("num_time_steps", C.c_int32),
                ("N_time", C.c_int32), ("rannacher_steps", C.c_int32), ("n_monitor", C.c_int32),
                ("monitor_times", C.POINTER(C.c_double))]


class _FdoDbfpInfo(C.Structure):
    _fields_ = [("dx", C.c_double), ("price", C.c_double), ("delta", C.c_double), ("gamma", C.c_double),
                ("v_len", C.c_int32), ("greek_idx", C.c_int32), ("n_monitor_steps", C.c_int32)]


@dataclass
class DbfpResult:
    V: np.ndarray           # value vector after the march (n entries, see fdo_dbfp_solve)
    s: np.ndarray           # n+1 node prices
    price: float
    delta: float
    gamma: float
    dx: float
    greek_idx: int
    monitor_steps: np.ndarray


DBFP_KO = {"none": 0, "down-and-in": 0, "up-and-in": 0, "double-in": 0,
           "down-and-out": 1, "up-and-out": 2, "double-out": 3}


def dbfp_solve(S_min, S_max, S0_eff, S_greek, K, T, sigma, r, b, q, is_call, barrier_type, lower_barrier,
               upper_barrier, rebate_amount, rebate_at_hit, monitor_times, num_space_nodes, num_time_steps,
               rannacher_steps, apply_KO=True, N_time=None) -> DbfpResult:
    """DiscreteBarrierFDMPricer._solve_grid + _interp_price + _delta_gamma_from_grid
    (discrete_barrier_fdm_pricer.py:442-547, :629-646, :949-977)."""
    L = lib()
    mt = np.ascontiguousarray(monitor_times, dtype=np.float64)
    nan = float("nan")
    p = _FdoDbfp(S_min, S_max, S0_eff, S_greek, K, T, sigma, r, b, q,
                 nan if lower_barrier is None else float(lower_barrier),
                 nan if upper_barrier is None else float(upper_barrier), float(rebate_amount),
                 int(bool(rebate_at_hit)), int(bool(is_call)), DBFP_KO[barrier_type], int(bool(apply_KO)),
                 int(num_space_nodes), int(num_time_steps), int(N_time if N_time is not None else num_time_steps),
                 int(rannacher_steps), len(mt), mt.ctypes.data_as(C.POINTER(C.c_double)))
    n = int(num_space_nodes)
    V = np.zeros(n + 1)
    s = np.zeros(n + 1)
    mon = np.zeros(int(num_time_steps) + 1, dtype=np.int32)
    info = _FdoDbfpInfo()
    L.fdo_dbfp_solve.restype = C.c_int
    rc = L.fdo_dbfp_solve(C.byref(p), _dptr(V), _dptr(s), mon.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(info))
    if rc != 0:
        raise RuntimeError(f"fdo_dbfp_solve failed: {rc}")
    return DbfpResult(V=V[:info.v_len].copy(), s=s, price=info.price, delta=info.delta, gamma=info.gamma,
                      dx=info.dx, greek_idx=info.greek_idx, monitor_steps=np.nonzero(mon)[0])


def max_threads() -> int:
    return int(lib().fdo_max_threads())


def norm_cdf(x: float) -> float:
    """discrete_barrier_fdm_pricer_cn.py:16-18"""
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


# --------------------------------------------------------------------------------------
# DiscreteBarrierFDMPricer2 (P2 = discrete_barrier_fdm_pricer_2.py): row-varying S-space solver
# --------------------------------------------------------------------------------------

class _FdoPricer2(C.Structure):
    _fields_ = [("sigma", C.c_double), ("r", C.c_double), ("strike", C.c_double), ("tenor", C.c_double), ("dt", C.c_double),
                ("lo_bar", C.c_double), ("up_bar", C.c_double), ("h_minus", C.c_double), ("h_plus", C.c_double),
                ("n_nodes", C.c_int32), ("n_steps", C.c_int32), ("rannacher", C.c_int32), ("is_call", C.c_int32),
                ("ko_kind", C.c_int32), ("bar_j", C.c_int32),
                ("s_nodes", C.POINTER(C.c_double)), ("v_terminal", C.POINTER(C.c_double)),
                ("project_after", C.POINTER(C.c_uint8))]


P2_KO = {"none": 0, "down-and-out": 1, "up-and-out": 2, "double-out": 3, "down-and-in": 0, "up-and-in": 0, "double-in": 0}


def p2_year_fraction(day_count: str, d0, d1) -> float:
    """P2:120-130."""
    dc = day_count.upper()
    if dc in ("ACT/365", "ACT/365F", "ACT/365 FIXED"):
        return max(0, (d1 - d0).days) / 365.0
    if dc in ("ACT/360",):
        return max(0, (d1 - d0).days) / 360.0
    if dc in ("30/360", "30E/360"):
        y0, m0, dd0 = d0.year, d0.month, min(d0.day, 30)
        y1, m1, dd1 = d1.year, d1.month, min(d1.day, 30)
        return ((y1 - y0) * 360 + (m1 - m0) * 30 + (dd1 - dd0)) / 360.0
    return max(0, (d1 - d0).days) / 365.0


def pricer2_prepare(spot, strike, valuation_date, maturity_date, volatility, option_type, barrier_type="none",
                    lower_barrier=None, upper_barrier=None, monitoring_dates=None, flat_rate_nacc=0.0, dividends=None,
                    num_space_nodes=600, num_time_steps=600, rannacher_steps=2, day_count="ACT/365",
                    smooth_payoff_around_strike=True, payoff_smoothing_half_width_nodes=2) -> dict:
    """Everything ``DiscreteBarrierFDMPricer2`` does outside ``_solve_pde_backward``, in the reference's own operation
    order (constructor P2:51-116, grid P2:146-168, n_lim / BGK decision P2:173-226, payoff P2:231-249, monitoring map
    P2:455-466, escrow shift P2:468-476, barrier interval P2:297-321).  Returns the inputs of ``fdo_pricer2_solve``."""
    yf = lambda a, b: p2_year_fraction(day_count, a, b)
    mons = sorted(monitoring_dates or [])
    tenor = yf(valuation_date, maturity_date)
    M = int(num_time_steps)
    dt = tenor / max(1, M)
    # grid
    anchors = [float(spot), float(strike)]
    if lower_barrier:
        anchors.append(lower_barrier)
    if upper_barrier:
        anchors.append(upper_barrier)
    s_max = 4.0 * max(anchors) * math.exp(float(volatility) * math.sqrt(max(tenor, 1e-12)))
    N = max(200, int(num_space_nodes))
    dS = (s_max - 0.0) / N
    nodes = [0.0 + i * dS for i in range(N + 1)]
    for x in (float(strike), lower_barrier, upper_barrier):
        if x is None:
            continue
        j = min(range(len(nodes)), key=lambda i: abs(nodes[i] - x))
        nodes[j] = float(x)
    # decision + BGK
    use_bgk, lo_adj, up_adj, k0, k1 = False, lower_barrier, upper_barrier, None, None
    if not (barrier_type == "none" or len(mons) == 0):
        first_mon, last_mon = min(mons), max(mons)
        inside = [d for d in mons if valuation_date < d <= maturity_date]
        if last_mon > first_mon and len(inside) > 0:
            dt_uniform = tenor / max(1, M)
            intervals = [yf(inside[i - 1], inside[i]) for i in range(1, len(inside))]
            N_hat = sum(max(1, int(round(ti / dt_uniform))) for ti in intervals)
            use_bgk = N_hat > 5 * M
            a_b = yf(first_mon, last_mon) / max(1, len(inside))
            adj = math.exp(0.5826 * float(volatility) * a_b)
            if lower_barrier is not None:
                lo_adj = lower_barrier / adj
            if upper_barrier is not None:
                up_adj = upper_barrier * adj
            k0 = max(0, min(M, int(round(yf(valuation_date, first_mon) / dt))))
            k1 = max(0, min(M, int(round(yf(valuation_date, last_mon) / dt))))
            k0, k1 = min(k0, k1), max(k0, k1)
    lo_eff, up_eff = (lo_adj, up_adj) if use_bgk else (lower_barrier, upper_barrier)
    # monitoring map
    project = np.zeros(M + 1, dtype=np.uint8)
    if use_bgk:
        project[k0:k1 + 1] = 1
    else:
        for d in mons:
            if valuation_date < d <= maturity_date:
                k = int(round(yf(valuation_date, d) / dt))
                if 0 <= k <= M:
                    project[k] = 1
    # escrow shift
    pv = 0.0
    for pay_date, amount in (dividends or []):
        tau = yf(valuation_date, pay_date)
        if tau > 0:
            pv += float(amount) * math.exp(-float(flat_rate_nacc) * tau)
    s = [max(x - pv, 0.0) for x in nodes]
    # terminal payoff
    K = float(strike)
    V = [max(x - K, 0.0) if option_type == "call" else max(K - x, 0.0) for x in s]
    if smooth_payoff_around_strike and payoff_smoothing_half_width_nodes > 0:
        m = int(payoff_smoothing_half_width_nodes)
        k_star = min(range(len(s)), key=lambda i: abs(s[i] - K))
        i0, i1 = max(0, k_star - m), min(len(s) - 1, k_star + m)
        S0, V0, S1, V1 = s[i0], V[i0], s[i1], V[i1]
        a = (V1 - V0) / ((S1 - S0) ** 2) if S1 != S0 else 0.0
        for i in range(i0, i1 + 1):
            V[i] = a * (s[i] - S0) ** 2 + V0
    # barrier interval (only knock-OUT types have one, P2:304, :312)
    bar_j, hm, hp = -1, 0.0, 0.0
    H = None
    if barrier_type in ("down-and-out", "double-out") and lo_eff is not None:
        H = lo_eff
    elif barrier_type in ("up-and-out", "double-out") and up_eff is not None:
        H = up_eff
    if H is not None:
        if H <= s[0]:
            bar_j, hm, hp = 0, 1e-12, s[1] - s[0]
        elif H >= s[-1]:
            bar_j, hm, hp = N - 1, s[N - 1] - s[N - 2], 1e-12
        else:
            j = max(0, min(N - 1, next(k for k in range(N) if s[k] <= H <= s[k + 1])))
            bar_j, hm, hp = j, max(1e-12, H - s[j]), max(1e-12, s[j + 1] - H)
    return dict(sigma=float(volatility), r=float(flat_rate_nacc), strike=K, tenor=tenor, dt=dt, lo_bar=lo_eff, up_bar=up_eff,
                h_minus=hm, h_plus=hp, n_steps=M, rannacher=int(rannacher_steps), is_call=option_type == "call",
                barrier_type=barrier_type, bar_j=bar_j, s_nodes=np.array(s), v_terminal=np.array(V), project_after=project,
                S_eff=float(spot) - pv, use_bgk=use_bgk, k_first=k0, k_last=k1, pv_divs=pv)


def pricer2_march(prep: dict, knock_out: bool = True) -> np.ndarray:
    """``_solve_pde_backward`` (P2:336-428) on prepared inputs.  ``knock_out=False``: the vanilla leg of a knock-in
    (``_solve_pde_backward(None, None, {}, S_shifted)``, P2:473-475)."""
    L = lib()
    s = np.ascontiguousarray(prep["s_nodes"], dtype=np.float64)
    v = np.ascontiguousarray(prep["v_terminal"], dtype=np.float64)
    proj = np.ascontiguousarray(prep["project_after"] if knock_out else np.zeros_like(prep["project_after"]), dtype=np.uint8)
    nan = float("nan")
    ko = P2_KO[prep["barrier_type"]] if knock_out else 0
    p = _FdoPricer2(prep["sigma"], prep["r"], prep["strike"], prep["tenor"], prep["dt"],
                    nan if (prep["lo_bar"] is None or not knock_out) else float(prep["lo_bar"]),
                    nan if (prep["up_bar"] is None or not knock_out) else float(prep["up_bar"]),
                    prep["h_minus"], prep["h_plus"], len(s), prep["n_steps"], prep["rannacher"], int(prep["is_call"]), ko,
                    prep["bar_j"] if knock_out else -1, _dptr(s), _dptr(v), proj.ctypes.data_as(C.POINTER(C.c_uint8)))
    V = np.empty(len(s))
    L.fdo_pricer2_solve.restype = C.c_int
    rc = L.fdo_pricer2_solve(C.byref(p), _dptr(V))
    if rc != 0:
        raise ZeroDivisionError("Tridiagonal pivot ~ 0") if rc == -1 else RuntimeError(f"fdo_pricer2_solve failed: {rc}")
    return V


def p2_interp_linear(x: float, xs, ys) -> float:
    """P2:433-449."""
    if x <= xs[0]:
        return float(ys[0])
    if x >= xs[-1]:
        return float(ys[-1])
    lo, hi = 0, len(xs) - 1
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if x < xs[mid]:
            hi = mid
        else:
            lo = mid
    x0, x1, y0, y1 = xs[lo], xs[hi], ys[lo], ys[hi]
    w = (x - x0) / (x1 - x0)
    return float((1 - w) * y0 + w * y1)


def pricer2_price(**kw) -> Tuple[float, np.ndarray, dict]:
    """``DiscreteBarrierFDMPricer2.price()`` (P2:468-482): (price, value vector, prepared inputs)."""
    prep = pricer2_prepare(**kw)
    V = pricer2_march(prep, True)
    if prep["barrier_type"] in ("down-and-in", "up-and-in", "double-in"):
        Vv = pricer2_march(prep, False)
        V = np.array([Vv[i] - V[i] for i in range(len(V))])
    s = prep["s_nodes"]
    return p2_interp_linear(prep["S_eff"], [float(x) for x in s], [float(x) for x in V]), V, prep
