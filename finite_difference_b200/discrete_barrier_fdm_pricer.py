"""Drop-in for the reference's date/curve-aware barrier pricer ``DiscreteBarrierFDMPricer``
(discrete_barrier_fdm_pricer.py:33-745 and :883-1027 -- the file's two parsable line ranges; the class is what
``run_config_scenarios.py`` drives over ``config_scenarios*.csv``).

Same constructor, attributes, method names and exceptions.  Everything outside the time loop (dates, curve
look-ups, escrowed dividends, the ``norm.ppf``-sized grid, monitor mapping) is done on the host with the
reference's own ``math`` / ``scipy`` calls; ``_solve_grid`` (:442-547) runs on the GPU through the C ABI.
There is no CPU fallback.

Reference behaviours kept on purpose (they are the reference's results on identical inputs):

* ``_solve_grid`` marches a vector ONE NODE SHORTER than ``s_nodes`` (``N = num_space_nodes - 1`` at :448 and the
  slice assignment ``V[1:-1] = V_int`` at :540): the kernel grid is ``num_space_nodes - 1`` intervals and the top
  Dirichlet value is evaluated at ``s_nodes[-1]``, one node above it (``FDM_FLAG_BDRY_TOP_NEXT``).
* the put boundary ``K e^{-r tau} * S_min * e^{(b-r) tau}`` (:391, ``FDM_FLAG_PUT_BDRY_SMIN``).
* the drift uses ``b - q`` (:461) while the boundary uses ``b`` (:382): ``b`` / ``b_bdry`` of the ABI.
* monitor mapping by ``floor(tau/dt + 1e-9)`` clamped to ``[1, N]`` (:404-410): the expiry monitor lands on step 1.
* a rebate that is not paid at hit is ``rebate * exp(-carry_rate * tau)`` at the monitor (:421-424): per-monitor
  values through ``mon_rebate``.
* ``configure_grid`` overwrites ``num_space_nodes`` with ``ceil(norm.ppf(0.99999) * num_time_steps)`` (:302-306).
* vega is one-sided and per vol point, ``(p(sigma + h) - p(sigma)) / (100 h)`` (:889-896).

The constructor of the reference calls ``self._build_stock_price_grid()`` (:167), which the file never defines; here
``stock_grid`` / ``grid_spacing`` are simply not provided.
"""
from __future__ import annotations

import datetime as _dt
import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import hostlogic
from .american import CurveLookup
from .batch import SolveBatch, solve_host
from .calendar_sa import SouthAfrica

_KO_TYPES = ("down-and-out", "up-and-out", "double-out")
_NORM_PPF_99999 = None


def _norm_ppf_99999() -> float:
    """``scipy.stats.norm.ppf(0.99999)`` (discrete_barrier_fdm_pricer.py:302), evaluated once."""
    global _NORM_PPF_99999
    if _NORM_PPF_99999 is None:
        from scipy.stats import norm
        _NORM_PPF_99999 = float(norm.ppf(0.99999))
    return _NORM_PPF_99999


def _norm_cdf(x: float) -> float:
    from scipy.stats import norm
    return float(norm.cdf(x))


class DiscreteBarrierFDMPricer:
    """CN FDM pricer for discretely monitored European barrier options with daily curves
    (discrete_barrier_fdm_pricer.py:33-41)."""

    def __init__(self, spot, strike, valuation_date, maturity_date, sigma, option_type, barrier_type="none",
                 lower_barrier=None, upper_barrier=None, monitor_dates=None, rebate_amount=0.0, rebate_at_hit=False,
                 already_hit=False, already_in=False, underlying_spot_days=3, option_days=0,
                 option_settlement_days=0, discount_curve=None, forward_curve=None, dividend_schedule=None,
                 trade_id=None, direction="long", quantity=1, contract_multiplier=1.0,
                 min_substeps_between_monitors=1, grid_type="uniform", sinh_alpha=1.5, lambda_diff_target=0.5,
                 num_space_nodes=400, num_time_steps=400, rannacher_steps=2, s_max_mult=4.5,
                 restart_on_monitoring=False, use_one_sided_greeks_near_barrier=True, mollify_final=True,
                 mollify_band_nodes=2, price_extrapolation=False, day_count="ACT/365",
                 calculate_greeks_in_pde=True) -> None:
        if any(x <= 0 for x in (spot, strike, sigma)):                                       # :84-87
            raise ValueError("spot, strike, sigma must be positive.")
        if maturity_date <= valuation_date:
            raise ValueError("maturity_date must be after valuation_date.")
        self.spot = spot
        self.strike = strike
        self.valuation_date = valuation_date
        self.maturity_date = maturity_date
        self.sigma = sigma
        self.option_type = option_type
        self.barrier_type = barrier_type
        self.lower_barrier = lower_barrier
        self.upper_barrier = upper_barrier
        self.monitor_dates = sorted(monitor_dates or [])
        self.rebate_amount = rebate_amount
        self.rebate_at_hit = rebate_at_hit
        self.already_hit = already_hit
        self.already_in = already_in
        self.underlying_spot_days = underlying_spot_days
        self.option_days = option_days
        self.option_settlement_days = option_settlement_days
        self.calendar = SouthAfrica()
        if isinstance(discount_curve, CurveLookup):
            # internal fast path of the batch runner: one shared date -> NACA table instead of a DataFrame copy and
            # a fresh table per pricer (the reference copies the frame, :118-119, and scans it per look-up)
            self._discount_lookup = discount_curve
            self.discount_curve_df = discount_curve.df
            self.forward_curve_df = forward_curve.df if isinstance(forward_curve, CurveLookup) else forward_curve
        else:
            self.discount_curve_df = discount_curve.copy() if discount_curve is not None else None
            self.forward_curve_df = forward_curve.copy() if forward_curve is not None else None
            self._discount_lookup = CurveLookup(self.discount_curve_df) if self.discount_curve_df is not None else None
        self.dividend_schedule = sorted(dividend_schedule or [], key=lambda x: x[0])
        self.trade_id = trade_id
        self.direction = direction
        self.quantity = int(quantity)
        self.contract_multiplier = float(contract_multiplier)
        self.num_space_nodes = int(num_space_nodes)
        self.num_time_steps = int(num_time_steps)
        self.rannacher_steps = int(rannacher_steps)
        self.min_substeps = max(1, int(min_substeps_between_monitors))
        self.lambda_diff_target = float(lambda_diff_target)
        self.s_max_mult = s_max_mult
        self.restart_on_monitoring = restart_on_monitoring
        self.mollify_final = mollify_final
        self.mollify_band_nodes = int(mollify_band_nodes)
        self.price_extrapolation = price_extrapolation
        self.use_one_sided_greeks_near_barrier = use_one_sided_greeks_near_barrier
        self.calculate_greeks_in_pde = calculate_greeks_in_pde
        self.day_count = day_count.upper().replace("F", "")
        self._year_denominator = self._infer_denominator(self.day_count)

        add = self.calendar.add_working_days                                                  # :141-145
        self.carry_start_date = add(self.valuation_date, self.underlying_spot_days)
        self.carry_end_date = add(self.maturity_date, self.underlying_spot_days)
        self.discount_start_date = add(self.valuation_date, self.option_days)
        self.discount_end_date = add(self.maturity_date, self.option_settlement_days)
        self.time_to_expiry = self._year_fraction(self.valuation_date, self.maturity_date)
        self.time_to_carry = self._year_fraction(self.carry_start_date, self.carry_end_date)
        self.time_to_discount = self._year_fraction(self.discount_start_date, self.discount_end_date)

        self.discount_rate_nacc = self.get_forward_nacc_rate(self.discount_start_date, self.discount_end_date)
        self.carry_rate_nacc = self.get_forward_nacc_rate(self.carry_start_date, self.carry_end_date)
        self.div_yield_nacc = self.dividend_yield_nacc()
        self.pv_divs = self.pv_dividends()
        self.forward_price = self.spot * math.exp((self.carry_rate_nacc - self.div_yield_nacc) * self.time_to_carry)
        self.b = math.log(self.forward_price / self.spot) / self.time_to_carry

        self.grid_type = grid_type
        self.sinh_alpha = sinh_alpha
        self.time_spacing = self.time_to_expiry / self.num_time_steps
        self.time_grid = [i * self.time_to_expiry / self.num_time_steps for i in range(self.num_time_steps + 1)]
        self.monitor_times = self._build_monitor_times_exact()
        self.s_nodes: List[float] = []
        self._S_min = 0.0
        self._S_max = 0.0

    # ---- day count / curves (:174-255) -------------------------------------------------------------
    def _infer_denominator(self, day_count: str) -> int:
        if day_count in ("ACT/365", "ACT/365F"):
            return 365
        if day_count in ("ACT/360", "ACT/364"):
            return 360 if day_count == "ACT/360" else 364
        if day_count in ("30/360", "BOND", "US30/360"):
            return 360
        return 365

    def _year_fraction(self, start_date: _dt.date, end_date: _dt.date) -> float:
        if end_date <= start_date:
            return 0.0
        if self.day_count in ("ACT/365", "ACT/365F", "ACT/360", "ACT/364"):
            return (end_date - start_date).days / float(self._year_denominator)
        if self.day_count in ("30/360", "BOND", "US30/360"):
            y1, m1, d1 = start_date.year, start_date.month, start_date.day
            y2, m2, d2 = end_date.year, end_date.month, end_date.day
            d1 = min(d1, 30)
            if d1 == 30:
                d2 = min(d2, 30)
            days = (y2 - y1) * 360 + (m2 - m1) * 30 + (d2 - d1)
            return days / 360.0
        return (end_date - start_date).days / 365.0

    def get_discount_factor(self, lookup_date: _dt.date) -> float:
        if self._discount_lookup is None:
            raise ValueError("No discount curve attached.")
        naca = self._discount_lookup.naca(lookup_date.isoformat())
        tau = self._year_fraction(self.valuation_date, lookup_date)
        return (1.0 + naca) ** (-tau)

    def get_nacc_rate(self, lookup_date: _dt.date) -> float:
        if self._discount_lookup is None:
            return 0.0
        try:
            return math.log(1.0 + self._discount_lookup.naca(lookup_date.isoformat()))
        except ValueError:
            return 0.0

    def get_forward_nacc_rate(self, start_date: _dt.date, end_date: _dt.date) -> float:
        df_far = self.get_discount_factor(end_date)
        df_near = self.get_discount_factor(start_date)
        tau = self._year_fraction(start_date, end_date)
        return -math.log(df_far / df_near) / max(1e-12, tau)

    def pv_dividends(self) -> float:
        if self.dividend_schedule is None:
            return 0.0
        pv = 0.0
        for pay_date, amount in self.dividend_schedule:
            if self.valuation_date < pay_date <= self.maturity_date:
                df = self.get_discount_factor(pay_date) / self.get_discount_factor(self.carry_start_date)
                pv += amount * df
        return pv

    def dividend_yield_nacc(self) -> float:
        pv_divs = self.pv_dividends()
        S = self.spot
        tau = max(1e-12, self.time_to_carry)
        if pv_divs <= 0.0:
            return 0.0
        if pv_divs >= S:
            raise ValueError("PV(dividend_schedule) >= spot.")
        return -math.log((S - pv_divs) / S) / tau

    def _build_monitor_times_exact(self) -> List[float]:                                      # :257-268
        times = []
        for d in self.monitor_dates:
            if self.valuation_date <= d <= self.maturity_date:
                t = self._year_fraction(self.valuation_date, d)
                if 0.0 <= t <= self.time_to_expiry:
                    times.append(t)
        if times[-1] < self.time_to_expiry - 1e-14:       # IndexError on an empty list, as in the reference
            times.append(self.time_to_expiry)
        return sorted(set(times))

    # ---- grid (:270-364) -------------------------------------------------------------------------------
    def choose_grid_parameters(self, S0, K, lower_barrier, upper_barrier, T, sigma) -> Tuple[int, int, float, float]:
        if T <= 0.0:
            raise ValueError("Maturity T must be positive.")
        if sigma <= 0.0:
            raise ValueError("Volatility sigma must be positive.")
        if S0 <= 0.0:
            raise ValueError("Spot S0 must be positive.")
        return hostlogic.dbfp_grid_parameters(S0, K, lower_barrier, upper_barrier, self.time_to_expiry, sigma,
                                              self.num_time_steps, _norm_ppf_99999())

    def configure_grid(self) -> None:
        N_space, N_time, S_min, S_max = self.choose_grid_parameters(
            S0=self.spot - self.pv_divs, K=self.strike, lower_barrier=self.lower_barrier,
            upper_barrier=self.upper_barrier, T=self.time_to_expiry, sigma=self.sigma)
        self.num_space_nodes = N_space
        self.num_time_steps = N_time
        self._S_min = S_min
        self._S_max = S_max

    def _build_log_grid(self) -> float:
        self.configure_grid()
        x_min = math.log(self._S_min)
        x_max = math.log(self._S_max)
        n = self.num_space_nodes
        dx = (x_max - x_min) / n
        self.s_nodes = [math.exp(x_min + i * dx) for i in range(n + 1)]
        self.lower_barrier_log = math.log(self.lower_barrier) if self.lower_barrier else None
        self.upper_barrier_log = math.log(self.upper_barrier) if self.upper_barrier else None
        return dx

    def _terminal_payoff(self) -> List[float]:
        if self.option_type == "call":
            return [max(s - self.strike, 0.0) for s in self.s_nodes]
        return [max(self.strike - s, 0.0) for s in self.s_nodes]

    def _boundary_values(self, tau: float) -> Tuple[float, float]:                            # :372-393
        S_min, S_max = self.s_nodes[0], self.s_nodes[-1]
        r, b, k = self.discount_rate_nacc, self.carry_rate_nacc, self.strike
        if self.option_type.lower() == "call":
            return 0.0, S_max * math.exp((b - r) * tau) - k * math.exp(-r * tau)
        return k * math.exp(-r * tau) * S_min * math.exp((b - r) * tau), 0.0

    def _monitor_indices_tau(self, dt: float) -> set:                                         # :395-411
        return set(hostlogic.dbfp_monitor_steps(self.time_to_expiry, dt, self.num_time_steps, self.monitor_times))

    # ---- the solve ---------------------------------------------------------------------------------------
    def _solve_args(self, apply_KO: bool, N_time: Optional[int] = None, sigma: Optional[float] = None) -> Dict[str, Any]:
        """Arguments of :func:`hostlogic.build_dbfp_batch` for one ``_solve_grid`` call at volatility ``sigma``
        (default ``self.sigma``).  Like the reference, a bumped sigma re-runs ``configure_grid`` (the domain follows it)."""
        sig = self.sigma if sigma is None else sigma
        N_space, N_steps, S_min, S_max = self.choose_grid_parameters(
            S0=self.spot - self.pv_divs, K=self.strike, lower_barrier=self.lower_barrier,
            upper_barrier=self.upper_barrier, T=self.time_to_expiry, sigma=sig)
        n_march = int(N_time) if N_time is not None else int(N_steps)
        assert n_march >= 1
        ko = self.barrier_type if (apply_KO and self.barrier_type in _KO_TYPES) else "none"
        return dict(
            S0_eff=self.spot - self.pv_divs, K=self.strike, T=self.time_to_expiry, sigma=sig,
            r=self.discount_rate_nacc, b=self.carry_rate_nacc, q=self.div_yield_nacc,
            is_call=self.option_type == "call", ko_type=ko, lower_barrier=self.lower_barrier,
            upper_barrier=self.upper_barrier, rebate_amount=self.rebate_amount, rebate_at_hit=self.rebate_at_hit,
            monitor_times=self.monitor_times, num_space_nodes=N_space, num_time_steps=N_steps,
            rannacher_steps=self.rannacher_steps, S_min=S_min, S_max=S_max, n_march=n_march)

    def _solve_descriptor(self, apply_KO: bool, N_time: Optional[int] = None, sigma: Optional[float] = None) -> SolveBatch:
        """Descriptor of one ``_solve_grid`` call."""
        return hostlogic.build_dbfp_batch(**{k: [v] for k, v in self._solve_args(apply_KO, N_time, sigma).items()})

    def _solve_grid(self, apply_KO: bool, N_time: int = None) -> List[float]:
        """One PDE solve on the GPU; returns the value vector (``num_space_nodes`` entries, see module docstring)."""
        self._build_log_grid()
        res = solve_host(self._solve_descriptor(apply_KO, N_time), dump=True)
        if int(res.status[0]) != 0:
            raise RuntimeError("non-finite PDE solution")
        return [float(v) for v in res.V[0, :self.num_space_nodes]]

    # ---- read-out (:629-646, :949-977) ------------------------------------------------------------
    def _interp_price(self, V: List[float]) -> float:
        s = self.s_nodes
        S0 = self.spot - self.pv_divs
        if S0 <= s[0]:
            return float(V[0])
        if S0 >= s[-1]:
            return float(V[-1])
        lo, hi = 0, len(s) - 1
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if S0 < s[mid]:
                hi = mid
            else:
                lo = mid
        w = (S0 - s[lo]) / (s[hi] - s[lo])
        return float((1.0 - w) * V[lo] + w * V[hi])

    def _delta_gamma_from_grid(self, V: List[float]) -> Tuple[float, float]:
        s = self.s_nodes
        S0 = self.spot
        idx = min(range(1, len(s) - 1), key=lambda i: abs(s[i] - S0))
        h1 = s[idx] - s[idx - 1]
        h2 = s[idx + 1] - s[idx]
        V_im1, V_i, V_ip1 = V[idx - 1], V[idx], V[idx + 1]
        delta = (-h2 / (h1 * (h1 + h2)) * V_im1 + (h2 - h1) / (h1 * h2) * V_i + h1 / (h2 * (h1 + h2)) * V_ip1)
        gamma = 2.0 * (V_im1 / (h1 * (h1 + h2)) - V_i / (h1 * h2) + V_ip1 / (h2 * (h1 + h2)))
        return delta, gamma

    def _map_KI_to_KO(self) -> Optional[str]:
        return {"down-and-in": "down-and-out", "up-and-in": "up-and-out", "double-in": "double-out"}.get(
            self.barrier_type)

    # ---- vanilla leg (:648-745) --------------------------------------------------------------------
    def _vanilla_black76_price(self, S: Optional[float] = None, sigma: Optional[float] = None,
                               T: Optional[float] = None) -> float:
        S = self.spot - self.pv_divs if S is None else S - self.pv_divs
        K = self.strike
        time_to_expiry = self.time_to_expiry if T is None else T
        sigma = self.sigma if sigma is None else sigma
        if self.time_to_discount <= 0 or sigma <= 0:
            return max(S - K, 0.0) if self.option_type == "call" else max(K - S, 0.0)
        sqrtT = math.sqrt(time_to_expiry)
        F = S * math.exp(self.carry_rate_nacc * self.time_to_carry)
        d1 = (math.log(F / K) + (0.5 * sigma * sigma) * time_to_expiry) / (sigma * sqrtT)
        d2 = d1 - sigma * sqrtT
        Nd1, Nd2 = _norm_cdf(d1), _norm_cdf(d2)
        disc = math.exp(-self.discount_rate_nacc * self.time_to_discount)
        if self.option_type == "call":
            return disc * (F * Nd1 - K * Nd2)
        return disc * (K * (1.0 - Nd2) - F * (1.0 - Nd1))

    def _vanilla_black76_greeks_fd(self, dS: float = 0.0001, dSigma: float = 0.0001, dT: float = 0.0001) -> Dict[str, float]:
        S0, sigma0, T0 = self.spot, self.sigma, self.time_to_expiry
        dS_pert = S0 * dS
        price = self._vanilla_black76_price
        p0 = price(S=S0, sigma=sigma0, T=T0)
        p_up_S = price(S=S0 + dS_pert, sigma=sigma0, T=T0)
        p_dn_S = price(S=S0 - dS_pert, sigma=sigma0, T=T0)
        delta = (p_up_S - p_dn_S) / (2.0 * dS_pert)
        gamma = (p_up_S - 2.0 * p0 + p_dn_S) / (dS_pert ** 2)
        vega = (price(S=S0, sigma=sigma0 + dSigma, T=T0) - p0) / (100 * dSigma)
        if T0 > 2.0 * dT:
            theta = -(price(S=S0, sigma=sigma0, T=T0 + dT) - price(S=S0, sigma=sigma0, T=T0 - dT)) / (2.0 * dT)
        else:
            theta = -(p0 - price(S=S0, sigma=sigma0, T=max(T0 - dT, 1e-8))) / dT
        return {"price": p0, "delta": delta, "gamma": gamma, "theta": theta, "vega": vega}

    # ---- price / greeks (:883-1026) ---------------------------------------------------------------
    def _pde_price_and_greeks3(self, apply_KO: bool, dv_sigma: float = 0.0001, use_richardson: bool = False) -> Dict[str, float]:
        """Base solve and the sigma + dv_sigma solve as ONE batch of two (:883-903)."""
        self._build_log_grid()
        batch = SolveBatch.concat([self._solve_descriptor(apply_KO),
                                   self._solve_descriptor(apply_KO, sigma=self.sigma + dv_sigma)])
        n0, n1 = int(batch.n_space[0]) + 1, int(batch.n_space[1]) + 1
        res = solve_host(batch, dump=True)
        V_grid = [float(v) for v in res.V[0, :n0]]
        price_base = self._interp_price(V_grid)
        delta, gamma = self._delta_gamma_from_grid(V_grid)
        # the bumped solve has its own grid (configure_grid follows sigma); _interp_price reads self.s_nodes
        orig_sigma = self.sigma
        self.sigma = self.sigma + dv_sigma
        try:
            self._build_log_grid()
            price_up = self._interp_price([float(v) for v in res.V[1, :n1]])
        finally:
            self.sigma = orig_sigma
        vega = (price_up - price_base) / (dv_sigma * 100)
        theta = -(0.5 * self.sigma * self.sigma * self.spot * self.spot * gamma
                  + (self.carry_rate_nacc - self.div_yield_nacc) * self.spot * delta
                  - self.discount_rate_nacc * price_base)
        return {"price": price_base, "delta": delta, "gamma": gamma, "vega": vega, "theta": theta}

    def price_log2(self, apply_KO: bool = True, use_richardson: bool = False) -> float:
        bt = self.barrier_type.lower()
        if bt == "none":
            return self._vanilla_black76_price()
        if bt in ("down-and-out", "up-and-out"):
            if self.already_hit:
                return self.rebate_amount * self.get_discount_factor(self.discount_end_date)
            return self._pde_price_and_greeks3(apply_KO=True, dv_sigma=0.0001, use_richardson=use_richardson)["price"]
        if bt in ("down-and-in", "up-and-in"):
            if self.already_in:
                return self._vanilla_black76_price()
            original_bt = bt
            P_van = self._vanilla_black76_price()
            self.barrier_type = "down-and-out" if bt == "down-and-in" else "up-and-out"
            try:
                g_ko = self._pde_price_and_greeks3(apply_KO=True, dv_sigma=0.0001, use_richardson=use_richardson)
            finally:
                self.barrier_type = original_bt
            return P_van - g_ko["price"]
        raise ValueError(f"Unsupported barrier_type: {self.barrier_type}")

    def greeks_log2(self, dv_sigma: float = 0.0001, use_richardson: bool = False) -> Dict[str, float]:
        bt = self.barrier_type.lower()
        if bt == "none":
            return self._vanilla_black76_greeks_fd()
        if bt in ("down-and-out", "up-and-out"):
            if self.already_hit:
                return {"price": 0.00, "delta": 0.00, "gamma": 0.00, "vega": 0.00, "theta": 0.00}
            return self._pde_price_and_greeks3(apply_KO=True, dv_sigma=dv_sigma, use_richardson=use_richardson)
        if bt in ("down-and-in", "up-and-in"):
            if self.already_in:
                return self._vanilla_black76_greeks_fd()
            original_bt = bt
            self.barrier_type = "none"
            try:
                g_van = self._vanilla_black76_greeks_fd()
                self.barrier_type = "down-and-out" if bt == "down-and-in" else "up-and-out"
                g_ko = self._pde_price_and_greeks3(apply_KO=True, dv_sigma=dv_sigma, use_richardson=use_richardson)
            finally:
                self.barrier_type = original_bt
            return {k: g_van[k] - g_ko[k] for k in g_van.keys()}
        raise ValueError(f"Unsupported barrier_type: {self.barrier_type}")


def price_and_greeks_batch(pricers: Sequence[DiscreteBarrierFDMPricer], dv_sigma: float = 0.0001):
    """``(price_log2(), greeks_log2(dv_sigma))`` of every pricer with ALL PDE solves in one C-ABI call.

    The reference executes four solves per knock-out / knock-in scenario (``price_log2`` and ``greeks_log2`` each run
    ``_pde_price_and_greeks3``, :883-903); they are two distinct solves (sigma and sigma + dv_sigma), so two per
    scenario are launched and shared.  Results are those of the per-object methods: the read-out is the same
    arithmetic in the kernel epilogue (linear interpolation at ``spot - pv_divs``, 3-point stencil), except with
    escrowed dividends, where the stencil is centred on ``spot`` (:952) and the value vector is read back instead.
    Returns ``(prices, greeks, n_solves, node_steps)``.
    """
    n = len(pricers)
    prices: List[Optional[float]] = [None] * n
    greeks: List[Optional[Dict[str, float]]] = [None] * n
    args, owners = [], []
    for i, pr in enumerate(pricers):
        bt = pr.barrier_type.lower()
        if bt == "none" or (bt in ("down-and-out", "up-and-out") and pr.already_hit) \
                or (bt in ("down-and-in", "up-and-in") and pr.already_in):
            prices[i] = pr.price_log2()
            greeks[i] = pr.greeks_log2(dv_sigma)
            continue
        if bt not in ("down-and-out", "up-and-out", "down-and-in", "up-and-in"):
            raise ValueError(f"Unsupported barrier_type: {pr.barrier_type}")
        original_bt = pr.barrier_type
        pr.barrier_type = {"down-and-in": "down-and-out", "up-and-in": "up-and-out"}.get(bt, bt)
        try:
            args.append(pr._solve_args(True))
            args.append(pr._solve_args(True, sigma=pr.sigma + dv_sigma))
        finally:
            pr.barrier_type = original_bt
        owners.append(i)
    n_solves = node_steps = 0
    if args:
        batch = hostlogic.build_dbfp_batch(**{k: [a[k] for a in args] for k in args[0]})
        need_dump = any(pricers[i].pv_divs != 0.0 for i in owners)
        res = solve_host(batch, dump=need_dump)
        n_solves, node_steps = batch.n_solves, batch.node_steps()
        for j, i in enumerate(owners):
            pr = pricers[i]
            bt = pr.barrier_type.lower()
            price_base, price_up = float(res.price[2 * j]), float(res.price[2 * j + 1])
            delta, gamma = float(res.delta[2 * j]), float(res.gamma[2 * j])
            if pr.pv_divs != 0.0:
                pr._build_log_grid()
                delta, gamma = pr._delta_gamma_from_grid([float(v) for v in res.V[2 * j, :pr.num_space_nodes]])
            vega = (price_up - price_base) / (dv_sigma * 100)
            theta = -(0.5 * pr.sigma * pr.sigma * pr.spot * pr.spot * gamma
                      + (pr.carry_rate_nacc - pr.div_yield_nacc) * pr.spot * delta - pr.discount_rate_nacc * price_base)
            g_ko = {"price": price_base, "delta": delta, "gamma": gamma, "vega": vega, "theta": theta}
            if bt in ("down-and-out", "up-and-out"):
                prices[i], greeks[i] = price_base, g_ko
            else:
                prices[i] = pr._vanilla_black76_price() - price_base
                g_van = pr._vanilla_black76_greeks_fd()
                greeks[i] = {k: g_van[k] - g_ko[k] for k in g_van.keys()}
    return prices, greeks, n_solves, node_steps


__all__ = ["DiscreteBarrierFDMPricer", "price_and_greeks_batch"]
