"""Drop-in American pricer classes on top of the CUDA solve.

Shared implementation of the three reference classes

* ``fd_american_option_pricer.AmericanFDMPricer``  (fd_american_option_pricer.py:12-751)
* ``fd_american_equity.AmericanFDMPricer``         (fd_american_equity.py:21-1068, same arithmetic,
  keyword spellings ``n_time`` / ``apply_ko``)
* ``fd_american_black76.AmericanFwdFDMPricer``     (fd_american_black76.py:13-625)

Constructor arguments, public attributes, method names, return types and exceptions are the
reference's.  The host side reproduces, with the same ``math`` calls, everything the reference
does outside the time loop (dates, curve look-ups, grid bounds, snapping, segment tables); the
time loop itself (``_solve_segment``, fd_american_option_pricer.py:392-511) and the dividend jump
(:517-548) run on the GPU through the C ABI.  There is no CPU fallback.

Methods that fan out into several solves (``price_log2``: 2, ``greeks_log2``: 6) submit ONE batch.
"""
from __future__ import annotations

import datetime as _dt
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import hostlogic
from .batch import SolveBatch, solve_host
from .calendar_sa import SouthAfrica


class CurveLookup:
    """Date -> NACA look-up with the reference's semantics: ISO-string equality on column ``Date``,
    first match wins, ``ValueError`` when missing (fd_american_option_pricer.py:190-197).
    Built once per curve instead of one DataFrame scan per look-up."""

    def __init__(self, curve):
        self.df = curve
        dates = curve["Date"]
        rates = curve["NACA"]
        dates = dates.tolist() if hasattr(dates, "tolist") else list(dates)
        rates = rates.tolist() if hasattr(rates, "tolist") else list(rates)
        self._map: Dict[str, float] = {}
        for d, v in zip(dates, rates):
            self._map.setdefault(d, float(v))

    def naca(self, iso: str) -> float:
        try:
            return self._map[iso]
        except KeyError:
            raise ValueError(f"Discount factor not found for date: {iso}") from None


class AmericanPricerBase:
    """See the module docstring.  Subclasses set ``_BLACK76`` and the Richardson step rule."""

    _BLACK76 = False

    def __init__(self, spot, strike, valuation_date, maturity_date, sigma, option_type, discount_curve,
                 forward_curve=None, dividend_schedule=None, trade_id=None, direction="long", quantity=1,
                 contract_multiplier=1.0, underlying_spot_days=0, option_days=0, option_settlement_days=0,
                 day_count="ACT/365", grid_type="uniform", num_space_nodes=400, num_time_steps=400,
                 rannacher_steps=2, s_max_mult=4.5) -> None:
        if spot <= 0.0 or strike <= 0.0 or sigma <= 0.0:
            raise ValueError("spot (forward), strike and sigma must be positive." if self._BLACK76
                             else "spot, strike and sigma must be positive.")
        if maturity_date <= valuation_date:
            raise ValueError("maturity_date must be after valuation_date.")
        if self._BLACK76:
            self.forward0 = float(spot)
        self.spot = float(spot)
        self.strike = float(strike)
        self.valuation_date = valuation_date
        self.maturity_date = maturity_date
        self.sigma = float(sigma)
        self.option_type = option_type.lower()
        if self.option_type not in ("call", "put"):
            raise ValueError("option_type must be 'call' or 'put'.")

        self.discount_curve_df = discount_curve.copy() if hasattr(discount_curve, "copy") else discount_curve
        self.forward_curve_df = (forward_curve.copy() if hasattr(forward_curve, "copy") else forward_curve) \
            if forward_curve is not None else None
        self._discount_lookup = CurveLookup(self.discount_curve_df)
        self.dividend_schedule = [] if self._BLACK76 else sorted(dividend_schedule or [], key=lambda x: x[0])

        self.trade_id = trade_id
        self.direction = direction
        self.quantity = int(quantity)
        self.contract_multiplier = float(contract_multiplier)

        self.calendar = SouthAfrica()
        self.underlying_spot_days = int(underlying_spot_days)
        self.option_days = int(option_days)
        self.option_settlement_days = int(option_settlement_days)
        self.day_count = day_count.upper().replace("F", "")
        self._year_denominator = self._infer_denominator(self.day_count)
        self.grid_type = grid_type.lower()

        add = self.calendar.add_working_days
        self.carry_start_date = add(self.valuation_date, self.underlying_spot_days)
        self.carry_end_date = add(self.maturity_date, self.underlying_spot_days)
        self.discount_start_date = add(self.valuation_date, self.option_days)
        self.discount_end_date = add(self.maturity_date, self.option_settlement_days)

        self.time_to_expiry = self._year_fraction(self.valuation_date, self.maturity_date)
        self.time_to_carry = self._year_fraction(self.carry_start_date, self.carry_end_date)
        self.time_to_discount = self._year_fraction(self.discount_start_date, self.discount_end_date)
        if self.time_to_expiry <= 0.0:
            raise ValueError("time_to_expiry must be positive.")

        self.discount_rate_nacc = self.get_forward_nacc_rate(self.discount_start_date, self.discount_end_date)
        if self.forward_curve_df is not None and not self._BLACK76:
            # the reference reads the DISCOUNT curve between the carry dates (fd_american_option_pricer.py:131-134)
            self.carry_rate_nacc = self.get_forward_nacc_rate(self.carry_start_date, self.carry_end_date)
        else:
            self.carry_rate_nacc = self.discount_rate_nacc
        self.div_yield_nacc = 0.0

        self.num_space_nodes = max(int(num_space_nodes), 3)
        self.num_time_steps = max(int(num_time_steps), 4)
        self.rannacher_steps = max(int(rannacher_steps), 0)
        self.s_max_mult = float(s_max_mult)

        self.snap_spot_to_grid = True
        self.snap_strike_to_grid = True
        self.spot_grid_index: Optional[int] = None
        self.spot_snapped: Optional[float] = None
        self.strike_grid_index: Optional[int] = None
        self.strike_snapped: Optional[float] = None
        self.s_nodes: List[float] = []
        self.x_nodes: List[float] = []
        self._S_min = 0.0
        self._S_max = 0.0
        self._dx = 0.0
        self.last_exercise_interval: Optional[Tuple[int, int]] = None

    # ---- day count / curves (fd_american_option_pricer.py:166-203) --------------------------
    def _infer_denominator(self, day_count: str) -> int:
        if day_count in ("ACT/365", "ACT/365F"):
            return 365
        if day_count in ("ACT/360", "ACT/364"):
            return 360 if day_count == "ACT/360" else 364
        if day_count in ("30/360", "BOND", "US30/360"):
            return 360
        return 365

    def _year_fraction(self, start_date, end_date) -> float:
        if end_date <= start_date:
            return 0.0
        if self.day_count in ("ACT/365", "ACT/365F", "ACT/360", "ACT/364"):
            return (end_date - start_date).days / float(self._year_denominator)
        if self.day_count in ("30/360", "BOND", "US30/360"):
            d1 = min(start_date.day, 30)
            d2 = min(end_date.day, 30) if d1 == 30 else end_date.day
            days = (end_date.year - start_date.year) * 360 + (end_date.month - start_date.month) * 30 + (d2 - d1)
            return days / 360.0
        return (end_date - start_date).days / 365.0

    def get_discount_factor(self, lookup_date) -> float:
        naca = self._discount_lookup.naca(lookup_date.isoformat())
        tau = self._year_fraction(self.valuation_date, lookup_date)
        return (1.0 + naca) ** (-tau)

    def get_forward_nacc_rate(self, start_date, end_date) -> float:
        df_far = self.get_discount_factor(end_date)
        df_near = self.get_discount_factor(start_date)
        tau = self._year_fraction(start_date, end_date)
        return -math.log(df_far / df_near) / max(1e-12, tau)

    # ---- grid (fd_american_option_pricer.py:209-275) -----------------------------------------
    def _configure_grid(self) -> None:
        self._S_min, self._S_max = hostlogic.american_domain(self.spot, self.strike, self.time_to_expiry,
                                                            self.sigma, self.s_max_mult)

    def _build_log_grid(self) -> float:
        self._configure_grid()
        x_min, dx = hostlogic.log_grid(self._S_min, self._S_max, self.num_space_nodes)
        self.x_nodes, self.s_nodes = hostlogic.grid_nodes(x_min, dx, self.num_space_nodes)
        self._dx = dx
        self._snap_critical_levels_to_grid()
        return dx

    def _snap_critical_levels_to_grid(self) -> None:
        s = self.s_nodes
        if not s:
            return
        self.spot_grid_index = hostlogic.nearest_node(s, self.spot)
        self.spot_snapped = s[self.spot_grid_index]
        self.strike_grid_index = hostlogic.nearest_node(s, self.strike)
        self.strike_snapped = s[self.strike_grid_index]

    def _strike_for_pde(self) -> float:
        if self.snap_strike_to_grid and self.strike_snapped is not None:
            return self.strike_snapped
        return self.strike

    def _spot_for_interp(self) -> float:
        if self.snap_spot_to_grid and self.spot_snapped is not None:
            return self.spot_snapped
        return self.spot

    def _intrinsic_payoff(self, S: float) -> float:
        k = self._strike_for_pde()
        return max(S - k, 0.0) if self.option_type == "call" else max(k - S, 0.0)

    def _terminal_payoff(self) -> List[float]:
        return [self._intrinsic_payoff(s) for s in self.s_nodes]

    def _div_times_tau(self) -> List[Tuple[float, float]]:
        """fd_american_option_pricer.py:304-315."""
        res: List[Tuple[float, float]] = []
        for pay_date, amount in self.dividend_schedule:
            if self.valuation_date < pay_date < self.maturity_date:
                t_rel = self._year_fraction(self.valuation_date, pay_date)
                if 0.0 < t_rel < self.time_to_expiry:
                    res.append((self.time_to_expiry - t_rel, float(amount)))
        res.sort(key=lambda x: x[0])
        return res

    # ---- the solves -----------------------------------------------------------------------------
    def _descriptor(self, n_time: Optional[int], sigma: float) -> Tuple[SolveBatch, float, float]:
        base_n = self.num_time_steps if n_time is None else int(n_time)
        S_min, S_max = hostlogic.american_domain(self.spot, self.strike, self.time_to_expiry, sigma, self.s_max_mult)
        x_min, dx = hostlogic.log_grid(S_min, S_max, self.num_space_nodes)
        batch = hostlogic.build_american_batch(
            [self.spot], [self.strike], [self.time_to_expiry], [sigma], [self.discount_rate_nacc],
            [self.carry_rate_nacc], [self.option_type == "call"], [self.num_space_nodes], [base_n],
            [self.rannacher_steps], [self._BLACK76], self.s_max_mult,
            div_taus=[self._div_times_tau()], x_min=[x_min], dx=[dx])
        return batch, S_min, S_max

    def _solve_many(self, specs: Sequence[Tuple[Optional[int], float]]):
        """One C-ABI call for several ``_solve_grid`` variants: ``specs`` = [(n_time, sigma)].
        Returns per spec (V list, SolveResults row index) plus the raw results."""
        parts = [self._descriptor(nt, sig)[0] for nt, sig in specs]
        res = solve_host(SolveBatch.concat(parts), dump=True)
        if int(res.status.max()) != 0:
            raise FloatingPointError("non-finite value in the PDE solution")
        n = self.num_space_nodes
        return [res.V[i, : n + 1].tolist() for i in range(len(specs))], res

    def _solve_grid_impl(self, n_time: Optional[int]) -> List[float]:
        """``_solve_grid`` (fd_american_option_pricer.py:554-602): refreshes the grid attributes from
        the current ``self.sigma`` and returns the value vector at valuation as a list."""
        self._build_log_grid()
        (V,), res = self._solve_many([(n_time, self.sigma)])
        if (int(res.idx[0, 0]), int(res.idx[0, 1])) != (self.spot_grid_index, self.strike_grid_index):
            raise RuntimeError("device and host disagree on the snapped spot/strike node "
                               f"({res.idx[0, :2].tolist()} vs {[self.spot_grid_index, self.strike_grid_index]})")
        self.last_exercise_interval = (int(res.idx[0, 2]), int(res.idx[0, 3]))
        return V

    # ---- read-out (fd_american_option_pricer.py:613-653) ------------------------------------------
    def _interp_price(self, V: Sequence[float]) -> float:
        s = self.s_nodes
        S0 = self._spot_for_interp()
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

    def _local_cubic_delta_gamma(self, V: Sequence[float]) -> Tuple[float, float]:
        s = self.s_nodes
        Sx = self._spot_for_interp()
        n = len(s) - 1
        i_near = hostlogic.nearest_node(s, Sx)
        i_near = min(max(i_near, 1), n - 2)
        idx = [i_near - 1, i_near, i_near + 1, i_near + 2]
        z = np.array([s[j] for j in idx], dtype=float) - Sx
        y = np.array([V[j] for j in idx], dtype=float)
        M = np.vstack([z ** 3, z ** 2, z, np.ones_like(z)]).T
        a, b, c, d = np.linalg.solve(M, y)          # the reference's LAPACK solve, kept on the host
        return float(c), float(2.0 * b)

    # ---- public API ---------------------------------------------------------------------------------
    def _fine_steps_price_log2(self) -> int:
        # fd_american_option_pricer.py:671 / fd_american_equity.py:950 use 2*num_space_nodes
        return 2 * self.num_space_nodes

    def _price_log(self, n_time: Optional[int] = None) -> float:
        return self._interp_price(self._solve_grid_impl(n_time))

    def _price_log2(self, use_richardson: bool = True) -> float:
        if not use_richardson:
            return self._price_log(self.num_time_steps)
        self._build_log_grid()
        (V_N, V_2N), _ = self._solve_many([(self.num_time_steps, self.sigma),
                                           (self._fine_steps_price_log2(), self.sigma)])
        pN, p2N = self._interp_price(V_N), self._interp_price(V_2N)
        return (4.0 * p2N - pN) / 3.0

    def _price_for_sigma(self, sigma: float, N_time: Optional[int] = None) -> float:
        orig_sigma = self.sigma
        try:
            self.sigma = sigma
            return self._price_log(N_time)
        finally:
            self.sigma = orig_sigma

    def _theta(self, sigma0: float, price: float, delta: float, gamma: float) -> float:
        r, b, S0 = self.discount_rate_nacc, self.carry_rate_nacc, self.spot
        return -(0.5 * sigma0 * sigma0 * S0 * S0 * gamma + (b - 0.0) * S0 * delta - r * price)

    def greeks_log2(self, dv_sigma: float = 0.01, use_richardson: bool = True) -> Dict[str, float]:
        """fd_american_option_pricer.py:682-751; all 6 (3) solves go to the GPU as one batch."""
        sigma0, h, N = self.sigma, dv_sigma, self.num_time_steps
        specs = [(N, sigma0)]
        if use_richardson:
            specs += [(2 * N, sigma0), (N, sigma0 + h), (N, sigma0 - h), (N, sigma0 + 2.0 * h), (N, sigma0 - 2.0 * h)]
        else:
            specs += [(N, sigma0 + h), (N, sigma0 - h)]
        Vs, res = self._solve_many(specs)
        self._build_log_grid()                       # base-sigma grid for the read-out
        price_N = self._interp_price(Vs[0])
        delta_N, gamma_N = self._local_cubic_delta_gamma(Vs[0])
        if use_richardson:
            price_2N = self._interp_price(Vs[1])
            delta_2N, gamma_2N = self._local_cubic_delta_gamma(Vs[1])
            price = (4.0 * price_2N - price_N) / 3.0
            delta = (4.0 * delta_2N - delta_N) / 3.0
            gamma = (4.0 * gamma_2N - gamma_N) / 3.0
            bumped = res.price[2:6]                   # V at each bumped grid's own snapped-spot node
            D_h = (bumped[0] - bumped[1]) / (2.0 * h)
            D_2h = (bumped[2] - bumped[3]) / (4.0 * h)
            dV_dsigma = (4.0 * D_h - D_2h) / 3.0
        else:
            price, delta, gamma = price_N, delta_N, gamma_N
            dV_dsigma = (res.price[1] - res.price[2]) / (2.0 * h)
        vega = dV_dsigma / 100.0
        theta = self._theta(sigma0, price, delta, gamma)
        return {"price": float(price), "delta": float(delta), "gamma": float(gamma), "vega": float(vega),
                "theta": float(theta)}
