"""Drop-in for the reference module ``discrete_barrier_fdm_pricer_2`` (class ``DiscreteBarrierFDMPricer2``): the
S-space Black-Scholes theta scheme with a row-varying tridiagonal matrix, a non-symmetric stencil next to the barrier,
BGK-shifted continuous window and escrowed dividends (discrete_barrier_fdm_pricer_2.py:16-591) -- with the time march
on the GPU (``fdm_rowvar_kernel`` through ``fdm_rowvar_solve_host``; no CPU fallback).

Same constructor, attributes, ``price()``, ``greeks(vega_bump)``, ``_solve_grid_once()`` and ``print_details()``.
``greeks()`` prices its three volatilities in ONE batch.  What the class derives outside the march lives in
``rowvar.py``; the barrier-aware delta / gamma stencils (:487-553) run on the host from the value vector.

Parity: the reference's explicit half-step has a sign error on the standard rows (:373-375), so with theta = 1/2
steps its value vector grows without bound (``nan`` on the default 600 x 600 grid).  The goldens
(tests/golden/pricer2.json, made by running the unmodified class) pin the regime in which it prices -- every step
implicit, ``rannacher_steps >= num_time_steps``; this class matches them to 1e-10.  With theta = 1/2 steps the kernel
applies the same coefficients as the reference, sign included, and diverges with it: same arithmetic, unpinned.
"""
from __future__ import annotations

import dataclasses
from datetime import date
from typing import Dict, List, Optional, Tuple

from . import rowvar

__all__ = ["DiscreteBarrierFDMPricer2"]


class DiscreteBarrierFDMPricer2:
    BGK_BETA = rowvar.BGK_BETA
    N_LIM = rowvar.N_LIM
    MIN_INTERVAL_STEPS = rowvar.MIN_INTERVAL_STEPS
    DEFAULT_DAYCOUNT = "ACT/365"

    def __init__(self, spot: float, strike: float, valuation_date: date, maturity_date: date, volatility: float,
                 option_type: str, barrier_type: str = "none", lower_barrier: Optional[float] = None,
                 upper_barrier: Optional[float] = None, monitoring_dates: Optional[List[date]] = None,
                 flat_rate_nacc: float = 0.0, dividends: Optional[List[Tuple[date, float]]] = None,
                 num_space_nodes: int = 600, num_time_steps: int = 600, rannacher_steps: int = 2,
                 day_count: str = DEFAULT_DAYCOUNT, smooth_payoff_around_strike: bool = True,
                 payoff_smoothing_half_width_nodes: int = 2):
        self.spot_price, self.strike_price = float(spot), float(strike)
        self.valuation_date, self.maturity_date = valuation_date, maturity_date
        self.option_type, self.barrier_type = option_type, barrier_type
        self.barrier_lower, self.barrier_upper = lower_barrier, upper_barrier
        self.monitoring_dates = sorted(monitoring_dates or [])
        self.volatility, self.r_flat = float(volatility), float(flat_rate_nacc)
        self.day_count = day_count.upper()
        self.dividends = [(d, float(a)) for d, a in (dividends or [])]
        self.num_space_nodes, self.num_time_steps = int(num_space_nodes), int(num_time_steps)
        self.rannacher_steps = int(rannacher_steps)
        self.smooth_payoff_around_strike = bool(smooth_payoff_around_strike)
        self.payoff_smoothing_half_width_nodes = int(payoff_smoothing_half_width_nodes)
        self.year_fraction = self._year_fraction
        self.tenor_years = self.year_fraction(self.valuation_date, self.maturity_date)
        self.dt = self.tenor_years / max(1, self.num_time_steps)
        # grid and continuous-window decision are frozen at the construction volatility (:108-116)
        self.S_nodes = rowvar.space_grid(self.spot_price, self.strike_price, self.barrier_lower, self.barrier_upper,
                                         self.volatility, self.tenor_years, self.num_space_nodes)
        self.dS = self.S_nodes[1] - self.S_nodes[0]
        (self.use_bgk_correction, self.bgk_lower, self.bgk_upper, self.k_first_cont, self.k_last_cont) = rowvar.continuous_window(
            self.barrier_type, self.barrier_lower, self.barrier_upper, self.monitoring_dates, self.valuation_date,
            self.maturity_date, self.volatility, self.tenor_years, self.dt, self.num_time_steps, self.day_count)
        self._frozen = None

    # ---- helpers with the reference's names ---------------------------------------------------------------------
    def _year_fraction(self, d0: date, d1: date) -> float:
        return rowvar.year_fraction(self.day_count, d0, d1)

    def _pv_dividends_escrow(self) -> float:
        return self._march().extras["pv_divs"]

    def _effective_barriers_for_pricing(self):
        if self.use_bgk_correction:
            return self.bgk_lower, self.bgk_upper
        return self.barrier_lower, self.barrier_upper

    def _monitoring_step_map(self) -> Dict[int, bool]:
        return {k: True for k in self._march().mon_keys}

    def _locate_barrier_interval(self, s_nodes, lo_bar, up_bar):
        j, hm, hp = rowvar.barrier_interval(s_nodes, self.barrier_type, lo_bar, up_bar)
        if j < 0:
            return (None, None, None, None)
        side = "down" if (self.barrier_type in ("down-and-out", "double-out") and lo_bar is not None) else "up"
        return (side, j, hm, hp)

    def _march(self) -> "rowvar.MarchInputs":
        """Inputs of the knock-out march at the CURRENT volatility on the frozen grid / window."""
        if self._frozen is None:
            m = rowvar.march_inputs(
                self.spot_price, self.strike_price, self.valuation_date, self.maturity_date, self.volatility, self.option_type,
                self.barrier_type, self.barrier_lower, self.barrier_upper, self.monitoring_dates, self.r_flat, self.dividends,
                self.num_space_nodes, self.num_time_steps, self.rannacher_steps, self.day_count,
                self.smooth_payoff_around_strike, self.payoff_smoothing_half_width_nodes, s_grid=self.S_nodes)
            # window and effective levels as decided in the constructor
            lo, up = self._effective_barriers_for_pricing()
            keys = (list(range(self.k_first_cont, self.k_last_cont + 1)) if self.use_bgk_correction else m.mon_keys)
            j, hm, hp = rowvar.barrier_interval(m.s_nodes, self.barrier_type, lo, up)
            self._frozen = dataclasses.replace(m, lo_bar=lo, up_bar=up, mon_keys=keys, bar_j=j, h_minus=hm, h_plus=hp)
        return dataclasses.replace(self._frozen, sigma=self.volatility)

    def _grids_for(self, sigmas) -> List[Tuple[List[float], List[float], float]]:
        """``_solve_grid_once`` (:468-476) for several volatilities in one batch."""
        base = self._march()
        marches = []
        for sg in sigmas:
            m = dataclasses.replace(base, sigma=float(sg))
            marches.append(m)
            if self.barrier_type in rowvar.KNOCK_IN:
                marches.append(m.vanilla_leg())
        res = rowvar.solve_rowvar_host(rowvar.RowVarBatch(marches), dump=True)
        if (res.status == 3).any():
            raise ZeroDivisionError("Tridiagonal pivot ~ 0")                       # :271-279
        n = len(base.s_nodes)
        out, k = [], 0
        s = [float(x) for x in base.s_nodes]
        for _ in sigmas:
            V = [float(v) for v in res.V[k, :n]]
            k += 1
            if self.barrier_type in rowvar.KNOCK_IN:
                Vv = res.V[k, :n]
                k += 1
                V = [float(Vv[i]) - V[i] for i in range(n)]
            out.append((s, V, base.s_eff))
        return out

    def _solve_grid_once(self):
        return self._grids_for([self.volatility])[0]

    @staticmethod
    def _interp_linear(x: float, xs: List[float], ys: List[float]) -> float:
        """:433-449."""
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
        w = (x - xs[lo]) / (xs[hi] - xs[lo])
        return float((1 - w) * ys[lo] + w * ys[hi])

    # ---- public API ----------------------------------------------------------------------------------------------
    def price(self) -> float:
        Sg, Vg, S_eff = self._solve_grid_once()
        return self._interp_linear(S_eff, Sg, Vg)

    def _delta_gamma_from_grid(self, s_nodes, V, S_eff, lo_bar, up_bar) -> Tuple[float, float]:
        """Barrier-aware stencils (:487-553): one-sided delta and a gamma blended with the PDE limit in the interval
        next to the barrier, a half-and-half blend in the second interval, central differences elsewhere."""
        N = len(s_nodes) - 1
        dS = s_nodes[1] - s_nodes[0]
        iS = max(1, min(N - 1, min(range(N), key=lambda k: abs(S_eff - s_nodes[k]))))
        central = ((V[iS + 1] - V[iS - 1]) / (2.0 * dS), (V[iS + 1] - 2.0 * V[iS] + V[iS - 1]) / (dS * dS))
        side, j, _hm, _hp = self._locate_barrier_interval(s_nodes, lo_bar, up_bar)
        if side is None:
            return float(central[0]), float(central[1])

        def one_sided(i):
            if side == "down":
                return (1.5 * V[i] - 2.0 * V[i - 1] + 0.5 * V[min(N, i + 1)]) / dS
            return (2.0 * V[i + 1] - 1.5 * V[i] - 0.5 * V[max(0, i - 1)]) / dS

        if iS in (j, j + 1):
            i = j + 1 if side == "down" else j
            d_os = one_sided(i)
            g_ns = (V[i + 1] - 2.0 * V[i] + V[i - 1]) / (dS * dS)
            denom = max(1e-14, (self.volatility * self.volatility) * s_nodes[i] * s_nodes[i])
            g_lim = 2.0 * (self.r_flat * V[i] - 0.0 * s_nodes[i] * d_os) / denom
            return float(d_os), float(0.5 * g_ns + (1.0 - 0.5) * g_lim)
        if iS in (j - 1, j + 2):
            d_os = one_sided(iS)
            g_os = (V[iS + 1] - 2.0 * V[iS] + V[iS - 1]) / (dS * dS)
            return float(0.5 * d_os + (1 - 0.5) * central[0]), float(0.5 * g_os + (1 - 0.5) * central[1])
        return float(central[0]), float(central[1])

    def greeks(self, vega_bump: float = 0.01) -> Dict[str, float]:
        lo, up = self._effective_barriers_for_pricing()
        sig0 = self.volatility
        (Sg, Vg, S_eff), (_, Vu, _), (_, Vd, _) = self._grids_for([sig0, sig0 + vega_bump, sig0 - vega_bump])
        delta, gamma = self._delta_gamma_from_grid(Sg, Vg, S_eff, lo, up)
        vega = (self._interp_linear(S_eff, Sg, Vu) - self._interp_linear(S_eff, Sg, Vd)) / (2.0 * vega_bump)   # :565-570
        return {"delta": float(delta), "gamma": float(gamma), "vega": float(vega)}

    def print_details(self) -> None:
        lo, up = self._effective_barriers_for_pricing()
        price, g = self.price(), self.greeks()
        print("==== Discrete Barrier Option (FD + CN, Rannacher) ====")
        print(f"Maturity Date           : {self.maturity_date.isoformat()}")
        print(f"T (years)               : {self.tenor_years:.9f}   [{self.day_count}]")
        print(f"Volatility (sigma)      : {self.volatility:.9f}")
        print(f"Flat r (NACC)           : {self.r_flat:.9f}")
        print(f"PV(dividends, escrow)   : {self._pv_dividends_escrow():.9f}")
        print("")
        print(f"Barrier type            : {self.barrier_type}")
        print(f"KO lower / upper        : {self.barrier_lower} / {self.barrier_upper}")
        print(f"BGK lower / upper       : {self.bgk_lower} / {self.bgk_upper}")
        print(f"BGK window steps        : {self.k_first_cont} .. {self.k_last_cont} (use_bgk={self.use_bgk_correction})")
        print("")
        print(f"Grid (space,time)       : {self.num_space_nodes}, {self.num_time_steps} (Rannacher {self.rannacher_steps})")
        print(f"Spot / Strike           : {self.spot_price:.6f} / {self.strike_price:.6f}")
        print(f"Effective barriers      : {lo} / {up}")
        print("")
        print(f"Price                   : {price:.9f}")
        print(f"Greeks                  : delta={g['delta']:.9f}, gamma={g['gamma']:.9f}, vega={g['vega']:.9f}")
