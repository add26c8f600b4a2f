"""A whole barrier scenario sweep as arrays: the host side of ``run_config_scenarios.run_all_scenarios`` without one
Python pricer object per row.

``DiscreteBarrierFDMPricer`` (discrete_barrier_fdm_pricer.py) derives, per scenario, a handful of scalars from dates and
a daily curve (:141-165), a ``norm.ppf``-sized grid (:270-320) and a monitor table (:395-411), then runs two PDE solves
(sigma, sigma + 1e-4; :883-903) and, for knock-in types, a Black-76 vanilla leg (:648-693, :694-745).  On a B200 the two
solves of a 2133 x 500 grid take ~8 us; the per-object Python path costs ~0.2 ms per row.  This module restates the
same derivations with numpy over the rows (every scenario of a sweep shares dates, monitors, option type and
day count, as in run_config_scenarios.py:197-245; spot, strike, sigma, the flat NACA rate and the barrier vary).

numpy's ``exp`` / ``log`` / ``power`` may differ from ``math`` in the last bit, so the descriptors are equal to the
per-object path's to ~1e-16 relative, not bit for bit; prices agree to ~1e-13 (tests/test_gpu_dbfp.py).
Cash dividends (escrowed spot, a per-row yield) are left to the per-object path.
"""
from __future__ import annotations

import datetime as _dt
import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import hostlogic
from .batch import solve_host
from .discrete_barrier_fdm_pricer import DiscreteBarrierFDMPricer, _norm_ppf_99999

_KO_OF = {"down-and-out": "down-and-out", "up-and-out": "up-and-out", "down-and-in": "down-and-out",
          "up-and-in": "up-and-out"}


def _ndtr(x):
    from scipy.special import ndtr
    return ndtr(x)


class BarrierScenarioBook:
    """Rows: ``S0, K, sigma, rate (flat NACA), barrier_type, lower_barrier, upper_barrier`` (arrays / lists of equal
    length; barriers may be None / NaN).  Shared: the keyword arguments of ``run_config_scenarios.run_scenario``."""

    def __init__(self, S0, K, sigma, rate, barrier_type: Sequence[str], lower_barrier, upper_barrier, valuation: _dt.date,
                 maturity: _dt.date, monitor_dates: list, opt_type: str = "call", rebate_amount: float = 0.0,
                 rebate_at_hit: bool = True, already_hit: bool = False, already_in: bool = False,
                 underlying_spot_days: int = 0, option_days: int = 0, option_settlement_days: int = 0,
                 day_count: str = "ACT/365", num_time_steps: int = 500, rannacher_steps: int = 2, **_ignored):
        n = len(S0)
        f = lambda v: np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n))
        self.S0, self.K, self.sigma, self.rate = f(S0), f(K), f(sigma), f(rate)
        self.low = np.array([np.nan if (v is None or (isinstance(v, float) and math.isnan(v))) else float(v)
                             for v in lower_barrier], dtype=np.float64)
        self.up = np.array([np.nan if (v is None or (isinstance(v, float) and math.isnan(v))) else float(v)
                            for v in upper_barrier], dtype=np.float64)
        self.barrier_type = [str(b).lower() for b in barrier_type]
        if np.any(self.S0 <= 0) or np.any(self.K <= 0) or np.any(self.sigma <= 0):
            raise ValueError("spot, strike, sigma must be positive.")
        bad = set(self.barrier_type) - set(_KO_OF) - {"none"}
        if bad:
            raise ValueError(f"Unsupported barrier_type: {sorted(bad)[0]}")
        self.is_call = opt_type == "call"
        self.opt_type = opt_type
        self.rebate_amount, self.rebate_at_hit = float(rebate_amount), bool(rebate_at_hit)
        self.already_hit, self.already_in = bool(already_hit), bool(already_in)
        self.num_time_steps, self.rannacher_steps = int(num_time_steps), int(rannacher_steps)
        # dates, year fractions and monitor times do not depend on the row: take them from one reference-shaped object
        from .utils import create_rate_df_iso
        from .american import CurveLookup
        probe = DiscreteBarrierFDMPricer(
            spot=1.0, strike=1.0, valuation_date=valuation, maturity_date=maturity, sigma=0.2, option_type=opt_type,
            monitor_dates=monitor_dates, discount_curve=CurveLookup(create_rate_df_iso(0.05)), forward_curve=None,
            underlying_spot_days=underlying_spot_days, option_days=option_days,
            option_settlement_days=option_settlement_days, num_time_steps=num_time_steps, day_count=day_count)
        self.T, self.T_carry, self.T_disc = probe.time_to_expiry, probe.time_to_carry, probe.time_to_discount
        self.monitor_times = probe.monitor_times
        yf = probe._year_fraction
        v = probe.valuation_date
        self._tau = {k: yf(v, getattr(probe, k)) for k in ("discount_start_date", "discount_end_date",
                                                           "carry_start_date", "carry_end_date")}

    # ---- constructor-derived scalars (:148-165), vectorised for a flat curve ------------------------------------
    def _forward_nacc(self, tau_near: float, tau_far: float, span: float) -> np.ndarray:
        base = 1.0 + self.rate
        df_far, df_near = np.power(base, -tau_far), np.power(base, -tau_near)      # get_discount_factor, :205-214
        return -np.log(df_far / df_near) / max(1e-12, span)                         # get_forward_nacc_rate, :226-230

    def _derive(self):
        self.r = self._forward_nacc(self._tau["discount_start_date"], self._tau["discount_end_date"], self.T_disc)
        self.b = self._forward_nacc(self._tau["carry_start_date"], self._tau["carry_end_date"], self.T_carry)

    # ---- choose_grid_parameters (:270-320), vectorised ---------------------------------------------------------------
    def _grid(self, sigma: np.ndarray):
        k = _norm_ppf_99999()
        lo_c = np.where(self.low > 0, self.low, np.inf)
        up_c = np.where(self.up > 0, self.up, np.inf)
        s_low = np.minimum(np.minimum(self.S0, self.K), np.minimum(lo_c, up_c))
        lo_c = np.where(self.low > 0, self.low, -np.inf)
        up_c = np.where(self.up > 0, self.up, -np.inf)
        s_high = np.maximum(np.maximum(self.S0, self.K), np.maximum(lo_c, up_c))
        width = 2.0 * k * sigma * math.sqrt(self.T)
        x_c = np.log(np.sqrt(s_low * s_high))
        S_min = np.minimum(np.exp(x_c - 0.5 * width), 0.5 * s_low)
        S_max = np.maximum(np.exp(x_c + 0.5 * width), 2 * s_high)
        n_space = np.ceil((width * self.num_time_steps) / (2 * sigma * math.sqrt(self.T))).astype(np.int64)
        return n_space, S_min, S_max

    # ---- Black-76 vanilla leg (:648-745), vectorised --------------------------------------------------------------------
    def _vanilla(self, S, sigma, T):
        F = S * np.exp(self.b * self.T_carry)
        sq = np.sqrt(T)
        d1 = (np.log(F / self.K) + (0.5 * sigma * sigma) * T) / (sigma * sq)
        d2 = d1 - sigma * sq
        disc = np.exp(-self.r * self.T_disc)
        if self.is_call:
            return disc * (F * _ndtr(d1) - self.K * _ndtr(d2))
        return disc * (self.K * (1.0 - _ndtr(d2)) - F * (1.0 - _ndtr(d1)))

    def _vanilla_greeks(self, dS=0.0001, dSigma=0.0001, dT=0.0001) -> Dict[str, np.ndarray]:
        S0, sg, T0 = self.S0, self.sigma, self.T
        h = S0 * dS
        p0 = self._vanilla(S0, sg, T0)
        pu, pd = self._vanilla(S0 + h, sg, T0), self._vanilla(S0 - h, sg, T0)
        vega = (self._vanilla(S0, sg + dSigma, T0) - p0) / (100 * dSigma)
        if T0 > 2.0 * dT:
            theta = -(self._vanilla(S0, sg, T0 + dT) - self._vanilla(S0, sg, T0 - dT)) / (2.0 * dT)
        else:
            theta = -(p0 - self._vanilla(S0, sg, max(T0 - dT, 1e-8))) / dT
        return {"price": p0, "delta": (pu - pd) / (2.0 * h), "gamma": (pu - 2.0 * p0 + pd) / (h ** 2), "theta": theta,
                "vega": vega}

    # ---- the sweep -------------------------------------------------------------------------------------------------------------
    def run(self, dv_sigma: float = 0.0001) -> Dict[str, np.ndarray]:
        """``price_log2()`` and ``greeks_log2()`` of every row (:907-1026): arrays ``price, delta, gamma, vega, theta``
        plus ``n_solves`` / ``node_steps`` of the PDE work that was launched."""
        n = len(self.S0)
        self._derive()
        bt = np.array(self.barrier_type)
        is_ko = (bt == "down-and-out") | (bt == "up-and-out")
        is_ki = (bt == "down-and-in") | (bt == "up-and-in")
        pde = (is_ko & (not self.already_hit)) | (is_ki & (not self.already_in))
        out = {k: np.zeros(n) for k in ("price", "delta", "gamma", "vega", "theta")}
        van = self._vanilla_greeks() if (np.any(is_ki) or np.any(bt == "none")) else None
        n_solves = node_steps = 0
        ko = {k: np.zeros(n) for k in out}
        if np.any(pde):
            idx = np.nonzero(pde)[0]
            sub = _Rows(self, idx)
            parts = []
            for sg in (sub.sigma, sub.sigma + dv_sigma):
                ns, S_min, S_max = sub._grid(sg)
                parts.append(dict(sigma=sg, num_space_nodes=ns, S_min=S_min, S_max=S_max))
            m = len(idx)
            cat = lambda a, b_: np.concatenate([a, b_])
            twice = lambda a: np.concatenate([a, a])
            ko_type = [_KO_OF[self.barrier_type[i]] for i in idx] * 2
            batch = hostlogic.build_dbfp_batch(
                S0_eff=twice(sub.S0), K=twice(sub.K), T=self.T, sigma=cat(parts[0]["sigma"], parts[1]["sigma"]),
                r=twice(sub.r), b=twice(sub.b), q=0.0, is_call=self.is_call, ko_type=ko_type,
                lower_barrier=[None if np.isnan(v) else v for v in twice(sub.low)],
                upper_barrier=[None if np.isnan(v) else v for v in twice(sub.up)], rebate_amount=self.rebate_amount,
                rebate_at_hit=self.rebate_at_hit, monitor_times=self.monitor_times,
                num_space_nodes=cat(parts[0]["num_space_nodes"], parts[1]["num_space_nodes"]),
                num_time_steps=self.num_time_steps, rannacher_steps=self.rannacher_steps,
                S_min=cat(parts[0]["S_min"], parts[1]["S_min"]), S_max=cat(parts[0]["S_max"], parts[1]["S_max"]),
                shared_monitors=True)
            res = solve_host(batch)
            if int(res.status.max()) != 0:
                raise RuntimeError("non-finite PDE solution")
            n_solves, node_steps = batch.n_solves, batch.node_steps()
            p0, p1 = res.price[:m], res.price[m:]
            d, g = res.delta[:m], res.gamma[:m]
            ko["price"][idx], ko["delta"][idx], ko["gamma"][idx] = p0, d, g
            ko["vega"][idx] = (p1 - p0) / (dv_sigma * 100)                                              # :896
            ko["theta"][idx] = -(0.5 * sub.sigma * sub.sigma * sub.S0 * sub.S0 * g + sub.b * sub.S0 * d - sub.r * p0)   # :898-902
        for k in out:
            v = out[k]
            v[is_ko] = ko[k][is_ko]
            if van is not None:
                v[is_ki] = van[k][is_ki] - ko[k][is_ki]                                                   # :1016
                v[bt == "none"] = van[k][bt == "none"]
        if self.already_hit and np.any(is_ko):                                                            # :922-925, :994-1001
            df_end = np.power(1.0 + self.rate, -self._tau["discount_end_date"])
            for k in out:
                out[k][is_ko] = 0.0
        if self.already_in and np.any(is_ki):                                                             # :930-931, :1006-1007
            for k in out:
                out[k][is_ki] = van[k][is_ki]
        out["price_log2"] = out["price"].copy()               # price_log2() == greeks_log2()["price"] except when already hit
        if self.already_hit and np.any(is_ko):
            out["price_log2"][is_ko] = (self.rebate_amount * df_end)[is_ko]
        out["n_solves"], out["node_steps"] = n_solves, node_steps
        return out


class _Rows:
    """View of a subset of a book's rows (the ones that need PDE solves)."""

    def __init__(self, book: BarrierScenarioBook, idx: np.ndarray):
        self.S0, self.K, self.sigma = book.S0[idx], book.K[idx], book.sigma[idx]
        self.low, self.up, self.r, self.b = book.low[idx], book.up[idx], book.r[idx], book.b[idx]
        self.T, self.num_time_steps = book.T, book.num_time_steps

    _grid = BarrierScenarioBook._grid


__all__ = ["BarrierScenarioBook"]
