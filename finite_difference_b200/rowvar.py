"""Row-varying S-space solves: the march of ``DiscreteBarrierFDMPricer2`` (discrete_barrier_fdm_pricer_2.py:336-428)
as batches through ``fdm_rowvar_solve_host`` / ``fdm_rowvar_solve``.

``march_inputs`` derives, for ONE contract, what the reference class computes outside ``_solve_pde_backward``: the
uniform S grid with the strike and barriers snapped onto nodes (:146-168), the n_lim decision with the BGK shift of a
frequently monitored barrier (:173-226), the escrow shift of discrete dividends (:468-470), the smoothed terminal
payoff (:231-249), the monitoring map (:455-466) and the interval that holds the active barrier with its one-sided
spacings (:297-321).  Plain ``math`` arithmetic in the reference's operation order: the grid, the payoff and every
scalar are bit-identical to the reference's.  ``RowVarBatch`` packs any number of such marches for the C ABI.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FLAG_CALL, FLAG_KO_DOWN, FLAG_KO_UP

BGK_BETA = 0.5826          # discrete_barrier_fdm_pricer_2.py:46
N_LIM = 5                  # :47
MIN_INTERVAL_STEPS = 1     # :48
_KO_FLAGS = {"down-and-out": FLAG_KO_DOWN, "up-and-out": FLAG_KO_UP, "double-out": FLAG_KO_DOWN | FLAG_KO_UP}
KNOCK_IN = ("down-and-in", "up-and-in", "double-in")


def year_fraction(day_count: str, d0, d1) -> float:
    """:120-130 (unknown conventions fall back to ACT/365)."""
    days = max(0, (d1 - d0).days)
    if day_count == "ACT/360":
        return days / 360.0
    if day_count in ("30/360", "30E/360"):
        return ((d1.year - d0.year) * 360 + (d1.month - d0.month) * 30 + (min(d1.day, 30) - min(d0.day, 30))) / 360.0
    return days / 365.0


@dataclass
class MarchInputs:
    """One march of ``_solve_pde_backward``; field names follow FdmRowVarBatch."""
    s_nodes: np.ndarray
    v_terminal: np.ndarray
    sigma: float
    r: float
    strike: float
    tenor: float
    dt: float
    n_steps: int
    rannacher: int
    is_call: bool
    barrier_type: str
    lo_bar: Optional[float]
    up_bar: Optional[float]
    bar_j: int
    h_minus: float
    h_plus: float
    mon_keys: List[int]              # keys of the reference's monitoring map
    s_eff: float
    extras: Dict[str, object] = field(default_factory=dict)

    def vanilla_leg(self) -> "MarchInputs":
        """``_solve_pde_backward(None, None, {}, S_shifted)`` -- the second march of a knock-in (:473-475)."""
        return MarchInputs(self.s_nodes, self.v_terminal, self.sigma, self.r, self.strike, self.tenor, self.dt, self.n_steps,
                           self.rannacher, self.is_call, "none", None, None, -1, 0.0, 0.0, [], self.s_eff, self.extras)


def space_grid(spot, strike, lower, upper, sigma, tenor, num_space_nodes) -> List[float]:
    """:146-168.  Truthiness of the barriers as in the reference (a barrier of 0.0 is ignored as an anchor)."""
    s_ref = max([spot, strike] + [b for b in (lower, upper) if b])
    s_max = 4.0 * s_ref * math.exp(sigma * math.sqrt(max(tenor, 1e-12)))
    N = max(200, int(num_space_nodes))
    dS = (s_max - 0.0) / N
    nodes = [0.0 + i * dS for i in range(N + 1)]
    for level in (strike, lower, upper):
        if level is not None:
            nearest = min(range(N + 1), key=lambda i: abs(nodes[i] - level))
            nodes[nearest] = float(level)
    return nodes


def continuous_window(barrier_type, lower, upper, mons, valuation, maturity, sigma, tenor, dt, M, day_count):
    """:173-226 -> (use_bgk, lower_adj, upper_adj, k_first, k_last)."""
    off = (False, lower, upper, None, None)
    if barrier_type == "none" or not mons or max(mons) <= min(mons):
        return off
    inside = [d for d in mons if valuation < d <= maturity]
    if not inside:
        return off
    yf = lambda a, b: year_fraction(day_count, a, b)
    dt_uniform = tenor / max(1, M)
    n_hat = sum(max(MIN_INTERVAL_STEPS, int(round(yf(inside[i - 1], inside[i]) / dt_uniform))) for i in range(1, len(inside)))
    a_b = yf(min(mons), max(mons)) / max(1, len(inside))
    adj = math.exp(BGK_BETA * sigma * a_b)
    lo = lower / adj if lower is not None else lower
    up = upper * adj if upper is not None else upper
    k0 = max(0, min(M, int(round(yf(valuation, min(mons)) / dt))))
    k1 = max(0, min(M, int(round(yf(valuation, max(mons)) / dt))))
    return (n_hat > N_LIM * M, lo, up, min(k0, k1), max(k0, k1))


def terminal_payoff(s: Sequence[float], strike: float, is_call: bool, smooth: bool, half_width: int) -> List[float]:
    """:231-249: payoff, replaced by a parabola through the ends of a window around the node nearest the strike."""
    V = [max(x - strike, 0.0) if is_call else max(strike - x, 0.0) for x in s]
    if smooth and half_width > 0:
        centre = min(range(len(s)), key=lambda i: abs(s[i] - strike))
        i0, i1 = max(0, centre - half_width), min(len(s) - 1, centre + half_width)
        curv = (V[i1] - V[i0]) / ((s[i1] - s[i0]) ** 2) if s[i1] != s[i0] else 0.0
        base, left = V[i0], s[i0]
        for i in range(i0, i1 + 1):
            V[i] = curv * (s[i] - left) ** 2 + base
    return V


def barrier_interval(s: Sequence[float], barrier_type: str, lo, up):
    """:297-321 -> (bar_j, h_minus, h_plus); bar_j = -1 when no knock-out barrier is active."""
    N = len(s) - 1
    H = None
    if barrier_type in ("down-and-out", "double-out") and lo is not None:
        H = lo
    elif barrier_type in ("up-and-out", "double-out") and up is not None:
        H = up
    if H is None:
        return -1, 0.0, 0.0
    if H <= s[0]:
        return 0, 1e-12, s[1] - s[0]
    if H >= s[-1]:
        return N - 1, s[N - 1] - s[N - 2], 1e-12
    j = next(k for k in range(N) if s[k] <= H <= s[k + 1])
    return j, max(1e-12, H - s[j]), max(1e-12, s[j + 1] - H)


def march_inputs(spot, strike, valuation_date, maturity_date, volatility, option_type, barrier_type="none",
                 lower_barrier=None, upper_barrier=None, monitoring_dates=None, flat_rate_nacc=0.0, dividends=None,
                 num_space_nodes=600, num_time_steps=600, rannacher_steps=2, day_count="ACT/365",
                 smooth_payoff_around_strike=True, payoff_smoothing_half_width_nodes=2, s_grid=None) -> MarchInputs:
    """Inputs of the knock-out march of one contract (``_solve_grid_once``, :468-476).  ``s_grid``: the un-shifted grid if
    the caller already holds it (the class builds it once, at the construction volatility, :108)."""
    dc = day_count.upper()
    spot, strike, sigma, r = float(spot), float(strike), float(volatility), float(flat_rate_nacc)
    mons = sorted(monitoring_dates or [])
    M = int(num_time_steps)
    tenor = year_fraction(dc, valuation_date, maturity_date)
    dt = tenor / max(1, M)
    nodes = list(s_grid) if s_grid is not None else space_grid(spot, strike, lower_barrier, upper_barrier, sigma, tenor, num_space_nodes)
    use_bgk, lo_adj, up_adj, k0, k1 = continuous_window(barrier_type, lower_barrier, upper_barrier, mons, valuation_date,
                                                        maturity_date, sigma, tenor, dt, M, dc)
    lo_eff, up_eff = (lo_adj, up_adj) if use_bgk else (lower_barrier, upper_barrier)
    if use_bgk:
        keys = list(range(k0, k1 + 1))
    else:
        keys = sorted({int(round(year_fraction(dc, valuation_date, d) / dt)) for d in mons if valuation_date < d <= maturity_date})
    pv = 0.0
    for pay_date, amount in (dividends or []):
        tau = year_fraction(dc, valuation_date, pay_date)
        if tau > 0:
            pv += float(amount) * math.exp(-r * tau)
    s = [max(x - pv, 0.0) for x in nodes]
    is_call = option_type == "call"
    V = terminal_payoff(s, strike, is_call, bool(smooth_payoff_around_strike), int(payoff_smoothing_half_width_nodes))
    bar_j, hm, hp = barrier_interval(s, barrier_type, lo_eff, up_eff)
    return MarchInputs(np.array(s), np.array(V), sigma, r, strike, tenor, dt, M, int(rannacher_steps), is_call, barrier_type,
                       lo_eff, up_eff, bar_j, hm, hp, keys, spot - pv,
                       dict(use_bgk=use_bgk, bgk_lower=lo_adj, bgk_upper=up_adj, k_first=k0, k_last=k1, pv_divs=pv, grid=nodes))


class RowVarBatch:
    """Marches packed for ``FdmRowVarBatch`` (include/fdm_b200.h)."""

    def __init__(self, marches: Sequence[MarchInputs]):
        self.marches = list(marches)
        n = len(self.marches)
        counts = np.array([len(m.s_nodes) for m in self.marches], dtype=np.int64)
        self.node_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.s_nodes = np.ascontiguousarray(np.concatenate([m.s_nodes for m in self.marches]) if n else np.zeros(0))
        self.v_terminal = np.ascontiguousarray(np.concatenate([m.v_terminal for m in self.marches]) if n else np.zeros(0))
        f = lambda name: np.ascontiguousarray([getattr(m, name) for m in self.marches], dtype=np.float64)
        self.sigma, self.r, self.strike, self.tenor, self.dt = f("sigma"), f("r"), f("strike"), f("tenor"), f("dt")
        self.h_minus, self.h_plus, self.s_eff = f("h_minus"), f("h_plus"), f("s_eff")
        nanv = lambda v: float("nan") if v is None else float(v)
        self.lo_bar = np.ascontiguousarray([nanv(m.lo_bar) for m in self.marches])
        self.up_bar = np.ascontiguousarray([nanv(m.up_bar) for m in self.marches])
        i = lambda vals: np.ascontiguousarray(vals, dtype=np.int32)
        self.n_steps = i([m.n_steps for m in self.marches])
        self.rannacher = i([m.rannacher for m in self.marches])
        self.flags = i([(FLAG_CALL if m.is_call else 0) | _KO_FLAGS.get(m.barrier_type, 0) for m in self.marches])
        self.bar_j = i([m.bar_j for m in self.marches])
        # projection after (M - k) completed steps for every key k < M of the map (:423-425); ascending
        mon = [sorted({m.n_steps - k for k in m.mon_keys if 0 <= k < m.n_steps}) if m.barrier_type in _KO_FLAGS else []
               for m in self.marches]
        self.mon_begin = np.concatenate([[0], np.cumsum([len(x) for x in mon])]).astype(np.int32)
        self.mon_step = i([k for x in mon for k in x])
        self.n_nodes_max = int(counts.max()) if n else 0

    @property
    def n_solves(self) -> int:
        return len(self.marches)

    def node_steps(self) -> int:
        return int(sum(len(m.s_nodes) * m.n_steps for m in self.marches))

    def struct(self, ptr) -> "_lib.FdmRowVarBatch":
        b = _lib.FdmRowVarBatch()
        b.n_solves, b.n_nodes_max, b.n_mon_total = self.n_solves, self.n_nodes_max, len(self.mon_step)
        for name in ("node_begin", "s_nodes", "v_terminal", "sigma", "r", "strike", "tenor", "dt", "lo_bar", "up_bar", "h_minus",
                     "h_plus", "s_eff", "n_steps", "rannacher", "flags", "bar_j"):
            setattr(b, name, ptr(name, getattr(self, name)))
        if len(self.mon_step):
            b.mon_begin, b.mon_step = ptr("mon_begin", self.mon_begin), ptr("mon_step", self.mon_step)
        return b


@dataclass
class RowVarResults:
    price: np.ndarray
    status: np.ndarray
    V: Optional[np.ndarray] = None       # [n, n_nodes_max] when dump=True
    kernel_ms: float = 0.0
    n_launches: int = 0


def solve_rowvar_host(batch: RowVarBatch, dump: bool = False) -> RowVarResults:
    """Run every march of ``batch`` on the current CUDA device from host buffers (no CPU fallback)."""
    n = batch.n_solves
    res = RowVarResults(price=np.empty(n), status=np.empty(n, dtype=np.int32))
    if n == 0:
        return res
    if batch.n_nodes_max > _lib.ROWVAR_MAX_NODES:
        raise _lib.FdmError(-3, f"{batch.n_nodes_max} grid nodes exceed the largest compiled kernel variant ({_lib.ROWVAR_MAX_NODES})")
    L = _lib.load()
    keep = []

    def host_ptr(_name, arr):
        keep.append(arr)
        return arr.ctypes.data

    b = batch.struct(host_ptr)
    out = _lib.FdmRowVarOut()
    out.price, out.status = res.price.ctypes.data, res.status.ctypes.data
    if dump:
        res.V = np.zeros((n, batch.n_nodes_max))
        out.v_dump, out.v_dump_stride = res.V.ctypes.data, batch.n_nodes_max
    nl, ms = C.c_int32(0), C.c_float(0.0)
    _lib.check(L.fdm_rowvar_solve_host(C.byref(b), C.byref(out), C.byref(nl), C.byref(ms)))
    res.kernel_ms, res.n_launches = float(ms.value), int(nl.value)
    return res


__all__ = ["MarchInputs", "RowVarBatch", "RowVarResults", "march_inputs", "solve_rowvar_host", "year_fraction", "space_grid",
           "continuous_window", "terminal_payoff", "barrier_interval"]
