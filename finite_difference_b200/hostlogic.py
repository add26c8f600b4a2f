"""Host-side preparation of PDE solves: everything the reference does between the pricer
constructor and the time loop, restated as (a) scalar functions that follow the reference's
``math``-module arithmetic operation for operation (used by the drop-in classes) and (b)
numpy-vectorised builders for large batches.

Reference line numbers: AM = fd_american_option_pricer.py, B76 = fd_american_black76.py,
BAR = discrete_barrier_fdm_pricer_cn.py (lines 1-537), DBF = discrete_barrier_fdm_pricer.py (lines 1-745, 883-1027).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import (FLAG_AMERICAN, FLAG_BLACK76, FLAG_BDRY_TOP_NEXT, FLAG_CALL, FLAG_KO_DOWN, FLAG_KO_UP,
                   FLAG_PUT_BDRY_SMIN, FLAG_SNAP)
from .batch import SolveBatch

# --------------------------------------------------------------------------------------
# American pricers: grid
# --------------------------------------------------------------------------------------


def american_domain(spot: float, strike: float, T: float, sigma: float, s_max_mult: float) -> Tuple[float, float]:
    """``_configure_grid`` (AM:209-229, B76:209-234) -> (S_min, S_max)."""
    s_low = min(spot, strike)
    s_high = max(spot, strike)
    s_c = math.sqrt(max(s_low * s_high, 1e-12))
    band = s_max_mult * sigma * math.sqrt(max(T, 1e-12))
    x_c = math.log(s_c)
    x_min = x_c - 0.5 * band
    x_max = x_c + 0.5 * band
    S_min = math.exp(x_min)
    S_max = math.exp(x_max)
    S_min = min(S_min, 0.5 * s_low)
    S_max = max(S_max, 2.0 * s_high)
    return max(S_min, 1e-8), S_max


def log_grid(S_min: float, S_max: float, n: int) -> Tuple[float, float]:
    """``_build_log_grid`` (AM:231-241, BAR:129-131) -> (x_min, dx)."""
    x_min = math.log(S_min)
    x_max = math.log(S_max)
    return x_min, (x_max - x_min) / float(n)


def grid_nodes(x_min: float, dx: float, n: int) -> Tuple[List[float], List[float]]:
    """x_i = x_min + i*dx, s_i = exp(x_i) (AM:239-240)."""
    x_nodes = [x_min + i * dx for i in range(n + 1)]
    return x_nodes, [math.exp(x) for x in x_nodes]


def nearest_node(s_nodes: Sequence[float], level: float) -> int:
    """First-minimum argmin of |s_i - level| (AM:253, AM:261)."""
    return min(range(len(s_nodes)), key=lambda i: abs(s_nodes[i] - level))


def american_domain_vec(spot, strike, T, sigma, s_max_mult):
    """Vectorised :func:`american_domain` (numpy transcendental functions, <= 1 ulp from libm)."""
    spot, strike, T, sigma = (np.asarray(v, dtype=np.float64) for v in (spot, strike, T, sigma))
    s_low = np.minimum(spot, strike)
    s_high = np.maximum(spot, strike)
    s_c = np.sqrt(np.maximum(s_low * s_high, 1e-12))
    band = s_max_mult * sigma * np.sqrt(np.maximum(T, 1e-12))
    x_c = np.log(s_c)
    S_min = np.minimum(np.exp(x_c - 0.5 * band), 0.5 * s_low)
    S_max = np.maximum(np.exp(x_c + 0.5 * band), 2.0 * s_high)
    return np.maximum(S_min, 1e-8), S_max


# --------------------------------------------------------------------------------------
# American pricers: time segments
# --------------------------------------------------------------------------------------


def american_segments(T: float, base_N: int, div_taus: Sequence[Tuple[float, float]], is_call: bool,
                      black76: bool):
    """Segment table of ``_solve_grid`` (AM:559-600; B76:450-472 is the no-dividend case).

    Returns lists (n_steps, dt, tau_start, restart_rannacher, cash_div_after).
    """
    if black76:
        div_taus = []                                        # B76:89
        base_N = max(int(base_N), 1)                         # B76:457-458
    base_N = int(base_N)
    base_dt = T / float(base_N)                              # AM:563
    tau_points = [0.0] + [tau for tau, _ in div_taus] + [T]  # AM:565
    n_segments = len(tau_points) - 1
    seg_lengths = [tau_points[i + 1] - tau_points[i] for i in range(n_segments)]
    seg_steps: List[int] = []
    remaining = base_N
    for L in seg_lengths[:-1]:                               # AM:571-574
        n = max(1, int(round(L / base_dt)))
        seg_steps.append(n)
        remaining -= n
    seg_steps.append(max(1, remaining))                      # AM:575
    steps, dts, tau0s, restarts, cash = [], [], [], [], []
    for i in range(n_segments):
        n_steps = seg_steps[i]
        dt = (tau_points[i + 1] - tau_points[i]) / float(n_steps)   # AM:412
        if not dt > 0.0:
            raise ValueError("zero-length time segment (two dividends map to the same tau); "
                             "the reference divides by zero here (fd_american_option_pricer.py:498)")
        steps.append(n_steps)
        dts.append(dt)
        tau0s.append(tau_points[i])
        restarts.append(1 if (black76 or i == 0 or is_call) else 0)  # AM:583-586, B76:470
        cash.append(float(div_taus[i][1]) if i < len(div_taus) else 0.0)
    return steps, dts, tau0s, restarts, cash


def build_american_batch(spot, strike, T, sigma, r, b, is_call, n_space, n_time, rannacher=2,
                         black76=False, s_max_mult=4.5, div_taus: Optional[Sequence] = None,
                         x_min=None, dx=None) -> SolveBatch:
    """Vectorised descriptor builder for American solves.  All array arguments broadcast.

    ``div_taus``: optional per-solve list of [(tau_div, cash)] (ascending tau) as
    ``_div_times_tau`` returns (AM:304-315); ``None`` / empty = single segment.
    ``x_min`` / ``dx``: override the grid (the scalar drop-in classes pass their libm values).
    """
    arrs = np.broadcast_arrays(*[np.asarray(v) for v in
                                 (spot, strike, T, sigma, r, b, is_call, n_space, n_time, rannacher, black76)])
    spot, strike, T, sigma, r, b = (np.ascontiguousarray(a, dtype=np.float64) for a in arrs[:6])
    is_call = arrs[6].astype(bool)
    n_space = np.maximum(arrs[7].astype(np.int64), 3).astype(np.int32)      # AM:142
    n_time = arrs[8].astype(np.int64)
    rannacher = np.maximum(arrs[9].astype(np.int64), 0).astype(np.int32)    # AM:144
    black76 = arrs[10].astype(bool)
    n = len(spot)
    if np.any(~(spot > 0)) or np.any(~(strike > 0)) or np.any(~(sigma > 0)):
        raise ValueError("spot, strike and sigma must be positive.")       # AM:68-69
    if np.any(~(T > 0)):
        raise ValueError("time_to_expiry must be positive.")               # AM:123-124
    if x_min is None:
        S_min, S_max = american_domain_vec(spot, strike, T, sigma, s_max_mult)
        x_min = np.log(S_min)
        dx = (np.log(S_max) - x_min) / n_space.astype(np.float64)
    else:
        x_min = np.broadcast_to(np.asarray(x_min, dtype=np.float64), (n,)).copy()
        dx = np.broadcast_to(np.asarray(dx, dtype=np.float64), (n,)).copy()

    flags = (FLAG_AMERICAN | FLAG_SNAP) + np.where(is_call, FLAG_CALL, 0) + np.where(black76, FLAG_BLACK76, 0)

    has_div = np.zeros(n, dtype=bool)
    if div_taus is not None:
        for i, d in enumerate(div_taus):
            if d and not black76[i]:
                has_div[i] = True
    base_N = np.where(black76, np.maximum(n_time, 1), n_time)
    if np.any(base_N < 1):
        raise ValueError("n_time must be >= 1")
    if not has_div.any():
        seg_begin = np.arange(n + 1, dtype=np.int32)
        seg_steps = base_N.astype(np.int32)
        seg_dt = (T - 0.0) / base_N.astype(np.float64)
        seg_tau0 = np.zeros(n)
        seg_restart = np.ones(n, dtype=np.int32)
        seg_div = np.zeros(n)
    else:
        sb, ss, sd, st, sr, sv = [0], [], [], [], [], []
        for i in range(n):
            dts = list(div_taus[i]) if has_div[i] else []
            steps, dtl, tau0s, restarts, cash = american_segments(float(T[i]), int(base_N[i]), dts,
                                                                  bool(is_call[i]), bool(black76[i]))
            ss += steps; sd += dtl; st += tau0s; sr += restarts; sv += cash
            sb.append(len(ss))
        seg_begin, seg_steps, seg_dt, seg_tau0, seg_restart, seg_div = sb, ss, sd, st, sr, sv
    return SolveBatch(x_min=x_min, dx=dx, sigma=sigma, r=r, b=b, spot=spot, strike=strike,
                      n_space=n_space, flags=flags.astype(np.int32), rannacher=rannacher,
                      seg_begin=seg_begin, seg_steps=seg_steps, seg_dt=seg_dt, seg_tau0=seg_tau0,
                      seg_restart=seg_restart, seg_div=seg_div)


# --------------------------------------------------------------------------------------
# Discrete barrier dataclass
# --------------------------------------------------------------------------------------


def barrier_configure(S0, K, T, sigma, lower_barrier, upper_barrier, monitor_times, N_space, N_time):
    """``configure_grid`` (BAR:59-118) -> (S_min, S_max, N_space, N_time)."""
    if T <= 0.0:
        raise ValueError("T must be positive")
    if sigma <= 0.0:
        raise ValueError("sigma must be positive")
    if S0 <= 0.0:
        raise ValueError("S0 must be positive")
    candidates = [S0, K]
    if lower_barrier is not None and lower_barrier > 0:
        candidates.append(lower_barrier)
    if upper_barrier is not None and upper_barrier > 0:
        candidates.append(upper_barrier)
    s_low = min(candidates)
    s_high = max(candidates)
    S_min = max(1e-8, s_low / 4.0)
    S_max = s_high * 4.0
    if S_min >= S_max:
        S_min = S0 / 5.0
        S_max = S0 * 5.0
    x_min, x_max = math.log(S_min), math.log(S_max)
    x_range = x_max - x_min
    dx_target = sigma * math.sqrt(T) / 12
    if dx_target <= 0.0:
        dx_target = x_range / 300.0
    if N_space is None:
        N_space = max(int(math.ceil(x_range / dx_target)), 300)
    if N_time is None:
        dx = x_range / N_space
        Ntime_opt = int(math.ceil(0.5 * sigma * sigma * T / (0.4 * dx * dx)))
        valid_mon = [t for t in (monitor_times or []) if 0.0 < t < T]
        N_time = max(Ntime_opt, N_space, 10 * (len(valid_mon) + 1))
    return S_min, S_max, N_space, N_time


def barrier_monitor_steps(T: float, N_time: int, monitor_times) -> List[int]:
    """``_monitor_indices_tau`` (BAR:177-193), sorted.  Python ``round`` is half-to-even."""
    dt = T / N_time
    idx = set()
    for t_mon in (monitor_times or []):
        if t_mon <= 0.0 or t_mon >= T:
            continue
        k = int(round((T - t_mon) / dt))
        if 0 < k < N_time:
            idx.add(k)
    return sorted(idx)


def barrier_monitor_steps_vec(T, N_time, monitor_times_2d, valid_mask):
    """Vectorised monitor mapping.  ``monitor_times_2d``: [n, max_mon] padded, ``valid_mask`` same shape.
    Returns CSR (mon_begin, mon_step).  ``np.rint`` rounds half to even like Python ``round``."""
    T = np.asarray(T, dtype=np.float64)[:, None]
    N_time = np.asarray(N_time, dtype=np.int64)[:, None]
    dt = T / N_time
    t = np.asarray(monitor_times_2d, dtype=np.float64)
    ok = valid_mask & (t > 0.0) & (t < T)
    with np.errstate(invalid="ignore"):
        k = np.rint((T - t) / dt)
    ok &= (k > 0) & (k < N_time)
    k = np.where(ok, k, np.iinfo(np.int32).max).astype(np.int64)
    k.sort(axis=1)
    # drop duplicates within a row (the reference keeps a set)
    dup = np.zeros_like(ok)
    dup[:, 1:] = k[:, 1:] == k[:, :-1]
    keep = (k != np.iinfo(np.int32).max) & ~dup
    counts = keep.sum(axis=1)
    mon_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return mon_begin, k[keep].astype(np.int32)


_KO_FLAG = {"none": 0, "down-and-out": FLAG_KO_DOWN, "up-and-out": FLAG_KO_UP, "double-out": FLAG_KO_DOWN | FLAG_KO_UP}


def build_barrier_batch(S0, K, T, sigma, r_disc, b_carry, is_call, ko_type: Sequence[str], lower_barrier,
                        upper_barrier, rebate, monitor_times: Sequence, N_space, N_time,
                        S_min=None, S_max=None, rannacher=0, rebate_discounted=False) -> SolveBatch:
    """Descriptors for barrier-dataclass solves ``_solve_grid(apply_KO)`` (BAR:219-302).

    ``ko_type[i]`` in {"none", "down-and-out", "up-and-out"} is the projection actually applied
    (``apply_KO`` and ``barrier_type`` combined, BAR:199-213, :299).  ``monitor_times[i]`` is a list.
    ``N_space`` / ``N_time`` must be resolved (see :func:`barrier_configure`).
    ``lower_barrier`` / ``upper_barrier`` entries may be ``None`` / NaN.

    Engine options beyond the importable dataclass (after the fragments of
    ``fd_discrete_barrier_option.py:851-1007``; no runnable reference, checked against the extended oracle):
    ``ko_type[i] == "double-out"`` (knocked out at or below the lower and at or above the upper level),
    ``rannacher`` fully implicit start-up steps, ``rebate_discounted`` (a knocked-out node is worth
    ``rebate * exp(-r_disc * tau)`` at time-to-expiry ``tau``: the rebate is paid at expiry).
    """
    n = len(S0)
    S0, K, T, sigma, r_disc, b_carry, rebate = (np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
                                                for v in (S0, K, T, sigma, r_disc, b_carry, rebate))
    low = np.array([np.nan if v is None else float(v) for v in lower_barrier], dtype=np.float64)
    up = np.array([np.nan if v is None else float(v) for v in upper_barrier], dtype=np.float64)
    if np.any(~(T > 0)):
        raise ValueError("T must be positive")
    if np.any(~(sigma > 0)):
        raise ValueError("sigma must be positive")
    if np.any(~(S0 > 0)):
        raise ValueError("S0 must be positive")
    N_space = np.ascontiguousarray(np.broadcast_to(np.asarray(N_space, dtype=np.int64), (n,)))
    N_time = np.ascontiguousarray(np.broadcast_to(np.asarray(N_time, dtype=np.int64), (n,)))
    if S_min is None:
        # BAR:69-87, vectorised
        lo_ok = np.where(low > 0, low, np.inf)
        up_ok = np.where(up > 0, up, np.inf)
        s_low = np.minimum(np.minimum(S0, K), np.minimum(lo_ok, up_ok))
        lo_ok2 = np.where(low > 0, low, -np.inf)
        up_ok2 = np.where(up > 0, up, -np.inf)
        s_high = np.maximum(np.maximum(S0, K), np.maximum(lo_ok2, up_ok2))
        S_min = np.maximum(1e-8, s_low / 4.0)
        S_max = s_high * 4.0
        bad = S_min >= S_max
        S_min = np.where(bad, S0 / 5.0, S_min)
        S_max = np.where(bad, S0 * 5.0, S_max)
    S_min = np.broadcast_to(np.asarray(S_min, dtype=np.float64), (n,))
    S_max = np.broadcast_to(np.asarray(S_max, dtype=np.float64), (n,))
    x_min = np.log(S_min)
    dx = (np.log(S_max) - x_min) / N_space                                   # BAR:129-131

    ko = np.array([_KO_FLAG[k] for k in ko_type], dtype=np.int64)
    flags = ko + np.where(np.asarray(is_call, dtype=bool), FLAG_CALL, 0)
    both = ko == (FLAG_KO_DOWN | FLAG_KO_UP)
    level = np.where(ko == FLAG_KO_UP, up, np.where(ko == 0, 0.0, low))
    # BAR:203 / :209: the projection is skipped when the matching barrier is None
    missing = np.isnan(level) | (both & np.isnan(up))
    flags = np.where(missing, flags & ~(FLAG_KO_DOWN | FLAG_KO_UP), flags)
    level = np.where(missing, 0.0, level)
    level_up = np.where(both & ~missing, up, 0.0) if np.any(both) else None
    rannacher = np.ascontiguousarray(np.broadcast_to(np.asarray(rannacher, dtype=np.int32), (n,)))

    max_mon = max((len(m) if m is not None else 0) for m in monitor_times) if n else 0
    if max_mon:
        mt = np.zeros((n, max_mon))
        valid = np.zeros((n, max_mon), dtype=bool)
        for i, m in enumerate(monitor_times):
            if m:
                mt[i, :len(m)] = m
                valid[i, :len(m)] = True
        mon_begin, mon_step = barrier_monitor_steps_vec(T, N_time, mt, valid)
    else:
        mon_begin, mon_step = np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    mon_rebate = None
    if rebate_discounted and len(mon_step):
        owner = np.repeat(np.arange(n), np.diff(mon_begin))
        tau_k = mon_step * (T / N_time)[owner]                    # tau_next = (m + 1) * dt at the projection
        mon_rebate = rebate[owner] * np.exp(-r_disc[owner] * tau_k)

    return SolveBatch(
        x_min=x_min, dx=dx, sigma=sigma, r=r_disc, b=b_carry, spot=S0, strike=K,
        n_space=N_space.astype(np.int32), flags=flags.astype(np.int32), rannacher=rannacher,
        seg_begin=np.arange(n + 1, dtype=np.int32), seg_steps=N_time.astype(np.int32),
        seg_dt=T / N_time, seg_tau0=np.zeros(n), seg_restart=(rannacher > 0).astype(np.int32),
        seg_div=np.zeros(n), barrier=level, rebate=rebate, mon_begin=mon_begin, mon_step=mon_step,
        barrier_up=level_up, mon_rebate=mon_rebate)


# --------------------------------------------------------------------------------------
# Date/curve-aware barrier pricer DiscreteBarrierFDMPricer (DBF)
# --------------------------------------------------------------------------------------


def dbfp_grid_parameters(S0: float, K: float, lower_barrier, upper_barrier, T: float, sigma: float, N_time: int,
                         k_ppf: float) -> Tuple[int, int, float, float]:
    """``choose_grid_parameters`` (DBF:270-320): ``(N_space, N_time, S_min, S_max)``.  ``k_ppf`` is
    ``scipy.stats.norm.ppf(0.99999)``; ``T`` is ``time_to_expiry``."""
    candidates = [S0, K]
    if lower_barrier is not None and lower_barrier > 0.0:
        candidates.append(lower_barrier)
    if upper_barrier is not None and upper_barrier > 0.0:
        candidates.append(upper_barrier)
    s_low = min(candidates)
    s_high = max(candidates)
    domain_width = 2.0 * k_ppf * sigma * math.sqrt(T)
    s_c = math.sqrt(s_low * s_high)
    x_c = math.log(s_c)
    x_min = x_c - 0.5 * domain_width
    x_max = x_c + 0.5 * domain_width
    S_min = math.exp(x_min)
    S_max = math.exp(x_max)
    S_min = min(S_min, 0.5 * s_low)
    S_max = max(S_max, 2 * s_high)
    N_space = math.ceil((domain_width * N_time) / (2 * sigma * math.sqrt(T)))
    return N_space, N_time, S_min, S_max


def dbfp_monitor_steps(T: float, dt: float, num_time_steps: int, monitor_times) -> List[int]:
    """``_monitor_indices_tau`` (DBF:395-411), sorted: floor mapping clamped to [1, num_time_steps]."""
    idx = set()
    for t_mon in (monitor_times or []):
        if t_mon <= 0.0 or t_mon > T:
            continue
        tau_mon = T - t_mon
        k = int(math.floor(tau_mon / dt + 1e-9))
        idx.add(max(1, min(num_time_steps, k)))
    return sorted(idx)


_DBFP_KO = {"none": 0, "down-and-out": FLAG_KO_DOWN, "up-and-out": FLAG_KO_UP, "double-out": FLAG_KO_DOWN | FLAG_KO_UP}


def build_dbfp_batch(S0_eff, K, T, sigma, r, b, q, is_call, ko_type: Sequence[str], lower_barrier, upper_barrier,
                     rebate_amount, rebate_at_hit, monitor_times: Sequence, num_space_nodes, num_time_steps,
                     rannacher_steps, S_min, S_max, n_march=None, shared_monitors: bool = False) -> SolveBatch:
    """Descriptors for ``DiscreteBarrierFDMPricer._solve_grid`` (DBF:442-547), one entry per solve.

    ``ko_type[i]``: the projection actually applied ("none" for knock-in types or ``apply_KO=False``, DBF:415-416).
    The kernel grid has ``num_space_nodes - 1`` intervals (DBF:448 with :540), the top Dirichlet value is taken one
    node above it, the drift carries ``b - q`` and the boundaries ``b``; rebates not paid at hit are discounted per
    monitor step (DBF:421-424).  ``num_space_nodes`` / ``S_min`` / ``S_max`` come from :func:`dbfp_grid_parameters`.
    """
    n = len(S0_eff)
    f = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
    S0_eff, K, T, sigma, r, b, q, rebate_amount, S_min, S_max = (f(v) for v in (S0_eff, K, T, sigma, r, b, q,
                                                                                 rebate_amount, S_min, S_max))
    ns = np.ascontiguousarray(np.broadcast_to(np.asarray(num_space_nodes, dtype=np.int64), (n,)))
    nt = np.ascontiguousarray(np.broadcast_to(np.asarray(num_time_steps, dtype=np.int64), (n,)))
    march = nt if n_march is None else np.ascontiguousarray(np.broadcast_to(np.asarray(n_march, dtype=np.int64), (n,)))
    ran = np.ascontiguousarray(np.broadcast_to(np.asarray(rannacher_steps, dtype=np.int64), (n,)))
    at_hit = np.broadcast_to(np.asarray(rebate_at_hit, dtype=bool), (n,))
    call = np.broadcast_to(np.asarray(is_call, dtype=bool), (n,))
    if np.any(ns < 4):
        raise ValueError("num_space_nodes must be >= 4")
    x_min = np.log(S_min)
    dx = (np.log(S_max) - x_min) / ns                                           # DBF:349-354
    dt = T / nt                                                                 # DBF:451
    low = np.array([np.nan if v is None else float(v) for v in lower_barrier], dtype=np.float64)
    up = np.array([np.nan if v is None else float(v) for v in upper_barrier], dtype=np.float64)
    ko = np.array([_DBFP_KO[k] for k in ko_type], dtype=np.int64)
    # DBF:428-435: a side whose level is None never knocks out
    ko = np.where(np.isnan(low), ko & ~FLAG_KO_DOWN, ko)
    ko = np.where(np.isnan(up), ko & ~FLAG_KO_UP, ko)
    both = (ko & (FLAG_KO_DOWN | FLAG_KO_UP)) == (FLAG_KO_DOWN | FLAG_KO_UP)
    level = np.where((ko & FLAG_KO_DOWN) != 0, low, np.where((ko & FLAG_KO_UP) != 0, up, 0.0))
    level_up = np.where(both, up, 0.0)
    flags = ko | np.where(call, FLAG_CALL, 0) | FLAG_BDRY_TOP_NEXT | FLAG_PUT_BDRY_SMIN

    if shared_monitors:
        steps = np.array([k for k in dbfp_monitor_steps(float(T[0]), float(dt[0]), int(nt[0]), monitor_times)
                          if k <= int(march[0])], dtype=np.int32)
        has = ko != 0
        counts = np.where(has, len(steps), 0)
        mon_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        mon_step = np.tile(steps, int(has.sum()))
        if len(steps):
            tau = steps.astype(np.float64) * float(dt[0])
            reb = np.where(at_hit[has, None], rebate_amount[has, None],
                           rebate_amount[has, None] * np.exp(-b[has, None] * tau[None, :]))          # DBF:421-424
            mon_rebate = reb.reshape(-1)
        else:
            mon_rebate = np.zeros(0)
        return SolveBatch(
            x_min=x_min, dx=dx, sigma=sigma, r=r, b=b - q, spot=S0_eff, strike=K,
            n_space=(ns - 1).astype(np.int32), flags=flags.astype(np.int32), rannacher=np.maximum(ran, 0).astype(np.int32),
            seg_begin=np.arange(n + 1, dtype=np.int32), seg_steps=march.astype(np.int32), seg_dt=dt,
            seg_tau0=np.zeros(n), seg_restart=np.ones(n, dtype=np.int32), seg_div=np.zeros(n),
            barrier=level, rebate=rebate_amount, mon_begin=mon_begin, mon_step=mon_step, b_bdry=b, barrier_up=level_up,
            mon_rebate=mon_rebate)

    mon_begin = [0]
    mon_step: List[int] = []
    mon_rebate: List[float] = []
    for i in range(n):
        if ko[i]:
            steps = [k for k in dbfp_monitor_steps(float(T[i]), float(dt[i]), int(nt[i]), monitor_times[i])
                     if k <= int(march[i])]
            mon_step.extend(steps)
            if at_hit[i]:
                mon_rebate.extend([float(rebate_amount[i])] * len(steps))
            else:                                                               # DBF:424, tau_left = (m+1) dt
                mon_rebate.extend(float(rebate_amount[i]) * math.exp(-float(b[i]) * (k * float(dt[i]))) for k in steps)
        mon_begin.append(len(mon_step))

    return SolveBatch(
        x_min=x_min, dx=dx, sigma=sigma, r=r, b=b - q, spot=S0_eff, strike=K,
        n_space=(ns - 1).astype(np.int32), flags=flags.astype(np.int32), rannacher=np.maximum(ran, 0).astype(np.int32),
        seg_begin=np.arange(n + 1, dtype=np.int32), seg_steps=march.astype(np.int32), seg_dt=dt,
        seg_tau0=np.zeros(n), seg_restart=np.ones(n, dtype=np.int32), seg_div=np.zeros(n),
        barrier=level, rebate=rebate_amount, mon_begin=np.asarray(mon_begin, dtype=np.int32),
        mon_step=np.asarray(mon_step, dtype=np.int32), b_bdry=b, barrier_up=level_up,
        mon_rebate=np.asarray(mon_rebate, dtype=np.float64))
