"""Drop-in ``DiscreteBarrierCrankNicolsonLog`` on top of the CUDA solve.

Shared implementation of the two importable reference dataclasses
(discrete_barrier_fdm_pricer_cn.py:25-537 and fd_discrete_barrier_option.py:6-423): pure
Crank-Nicolson in log-spot, knock-out projection at the monitoring steps, price by linear
interpolation at ``S0``, three-point delta / gamma, vega by a central volatility bump.  The two
modules differ only in how ``price()`` / ``greeks()`` compose solves; see
``discrete_barrier_fdm_pricer_cn.py`` and ``fd_discrete_barrier_option.py`` in this package.

Grid configuration, monitor-step mapping and the read-out formulas run on the host with the
reference's ``math`` calls; the time loop (discrete_barrier_fdm_pricer_cn.py:278-300) runs on the
GPU through the C ABI.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import hostlogic
from .batch import SolveBatch, solve_host

SQRT_2PI = math.sqrt(2.0 * math.pi)


def norm_pdf(x: float) -> float:
    return math.exp(-0.5 * x * x) / SQRT_2PI


def norm_cdf(x: float) -> float:
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


_KO_OF = {"down-and-in": "down-and-out", "up-and-in": "up-and-out", "double-in": "double-out"}
KO_TYPES = ("down-and-out", "up-and-out", "double-out")


@dataclass
class DiscreteBarrierBase:
    S0: float
    K: float
    T: float
    sigma: float
    r_disc: float
    b_carry: float
    option_type: str
    barrier_type: str
    lower_barrier: Optional[float] = None
    upper_barrier: Optional[float] = None
    rebate: float = 0.0
    monitor_times: Optional[List[float]] = None
    N_space: Optional[int] = None
    N_time: Optional[int] = None

    _S_min: float = field(init=False, default=0.0)
    _S_max: float = field(init=False, default=0.0)
    s_nodes: List[float] = field(init=False, default_factory=list)

    # ---- grid (discrete_barrier_fdm_pricer_cn.py:59-135) --------------------------------------
    def configure_grid(self) -> None:
        self._S_min, self._S_max, self.N_space, self.N_time = hostlogic.barrier_configure(
            self.S0, self.K, self.T, self.sigma, self.lower_barrier, self.upper_barrier, self.monitor_times,
            self.N_space, self.N_time)

    def _build_log_grid(self) -> float:
        if self._S_min <= 0.0 or self._S_max <= 0.0 or self.N_space is None:
            self.configure_grid()
        x_min, dx = hostlogic.log_grid(self._S_min, self._S_max, self.N_space)
        _, self.s_nodes = hostlogic.grid_nodes(x_min, dx, self.N_space)
        return dx

    def _monitor_indices_tau(self, dt: float) -> set:
        """discrete_barrier_fdm_pricer_cn.py:177-193 (``dt`` is implied by ``T / N_time``)."""
        return set(hostlogic.barrier_monitor_steps(self.T, self.N_time, self.monitor_times))

    # ---- solves -----------------------------------------------------------------------------------
    def _descriptor(self, apply_KO: bool, sigma: float, barrier_type: Optional[str] = None,
                    N_time: Optional[int] = None) -> SolveBatch:
        bt = (self.barrier_type if barrier_type is None else barrier_type).lower()
        ko = bt if (apply_KO and bt in KO_TYPES) else "none"
        return hostlogic.build_barrier_batch(
            [self.S0], [self.K], [self.T], [sigma], [self.r_disc], [self.b_carry],
            [self.option_type.lower() == "call"], [ko], [self.lower_barrier], [self.upper_barrier], [self.rebate],
            [self.monitor_times or []], [self.N_space], [self.N_time if N_time is None else N_time],
            S_min=[self._S_min], S_max=[self._S_max], rannacher=int(getattr(self, "rannacher_steps", 0)),
            rebate_discounted=bool(getattr(self, "rebate_discounted", False)))

    def _solve_many(self, specs: Sequence[Tuple[bool, float, Optional[str]]]):
        """One C-ABI call for several ``_solve_grid`` variants: specs = [(apply_KO, sigma, barrier_type)]."""
        self.configure_grid()
        self._build_log_grid()
        res = solve_host(SolveBatch.concat([self._descriptor(*s) for s in specs]), dump=True)
        if int(res.status.max()) != 0:
            raise FloatingPointError("non-finite value in the PDE solution")
        return [res.V[i, : self.N_space + 1].tolist() for i in range(len(specs))], res

    def _solve_grid(self, apply_KO: bool) -> List[float]:
        """discrete_barrier_fdm_pricer_cn.py:219-302: value vector at valuation, as a list."""
        (V,), _ = self._solve_many([(apply_KO, self.sigma, None)])
        return V

    # ---- read-out (discrete_barrier_fdm_pricer_cn.py:308-353) ----------------------------------------
    def _interp_price_from_grid(self, V: Sequence[float]) -> float:
        S0, s = self.S0, self.s_nodes
        if S0 <= s[0]:
            return V[0]
        if S0 >= s[-1]:
            return V[-1]
        lo, hi = 0, len(s) - 1
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if S0 < s[mid]:
                hi = mid
            else:
                lo = mid
        w = (S0 - s[lo]) / (s[hi] - s[lo])
        return (1.0 - w) * V[lo] + w * V[hi]

    def _delta_gamma_from_grid(self, V: Sequence[float]) -> Tuple[float, float]:
        s, S0 = self.s_nodes, self.S0
        idx = min(range(1, len(s) - 1), key=lambda i: abs(s[i] - S0))
        h1 = s[idx] - s[idx - 1]
        h2 = s[idx + 1] - s[idx]
        V_im1, V_i, V_ip1 = V[idx - 1], V[idx], V[idx + 1]
        delta = (-h2 / (h1 * (h1 + h2)) * V_im1 + (h2 - h1) / (h1 * h2) * V_i + h1 / (h2 * (h1 + h2)) * V_ip1)
        gamma = 2.0 * (V_im1 / (h1 * (h1 + h2)) - V_i / (h1 * h2) + V_ip1 / (h2 * (h1 + h2)))
        return delta, gamma

    def _pde_theta(self, price: float, delta: float, gamma: float) -> float:
        S0, sigma = self.S0, self.sigma
        return -(0.5 * sigma * sigma * S0 * S0 * gamma + self.b_carry * S0 * delta - self.r_disc * price)

    def _pde_price_and_greeks(self, apply_KO: bool, dv_sigma: float, barrier_type: Optional[str] = None) -> Dict[str, float]:
        """Base, sigma+dv and sigma-dv solves as one batch (discrete_barrier_fdm_pricer_cn.py:429-466)."""
        sig = self.sigma
        (V, V_up, V_dn), _ = self._solve_many([(apply_KO, sig, barrier_type), (apply_KO, sig + dv_sigma, barrier_type),
                                               (apply_KO, sig - dv_sigma, barrier_type)])
        price = self._interp_price_from_grid(V)
        delta, gamma = self._delta_gamma_from_grid(V)
        theta = self._pde_theta(price, delta, gamma)
        vega = (self._interp_price_from_grid(V_up) - self._interp_price_from_grid(V_dn)) / (2.0 * dv_sigma)
        return {"price": price, "delta": delta, "gamma": gamma, "theta": theta, "vega": vega}

    _TYPES = ("none", "down-and-out", "up-and-out", "down-and-in", "up-and-in")    # what the reference accepts (:502)

    def _check_type(self) -> str:
        bt = self.barrier_type.lower()
        if bt not in self._TYPES:
            raise ValueError(f"Unsupported barrier_type: {self.barrier_type}")
        return bt

    # ---- engine options (not in the importable reference dataclass) --------------------------
    def _pde_price_richardson(self, apply_KO: bool, barrier_type: Optional[str] = None,
                              sigma: Optional[float] = None) -> float:
        """Richardson extrapolation in time over N and N/2 steps, ``(4 p_N - p_{N/2}) / 3`` with
        ``N/2 = max(1, N // 2)`` (fd_discrete_barrier_option.py:1979-2005); both solves in one C-ABI call."""
        self.configure_grid()
        self._build_log_grid()
        sig = self.sigma if sigma is None else sigma
        n_full = int(self.N_time)
        if n_full < 2:
            res = solve_host(self._descriptor(apply_KO, sig, barrier_type), dump=True)
            return self._interp_price_from_grid(res.V[0, : self.N_space + 1].tolist())
        res = solve_host(SolveBatch.concat([self._descriptor(apply_KO, sig, barrier_type),
                                            self._descriptor(apply_KO, sig, barrier_type, N_time=max(1, n_full // 2))]),
                         dump=True)
        if int(res.status.max()) != 0:
            raise FloatingPointError("non-finite value in the PDE solution")
        p_n = self._interp_price_from_grid(res.V[0, : self.N_space + 1].tolist())
        p_half = self._interp_price_from_grid(res.V[1, : self.N_space + 1].tolist())
        return (4.0 * p_n - p_half) / 3.0

    def price_richardson(self) -> float:
        """PDE price with Richardson extrapolation in time; knock-in by parity of two PDE legs."""
        bt = self._check_type()
        if bt == "none":
            return self._pde_price_richardson(apply_KO=False)
        if bt in KO_TYPES:
            return self._pde_price_richardson(apply_KO=True)
        return (self._pde_price_richardson(apply_KO=False, barrier_type="none")
                - self._pde_price_richardson(apply_KO=True, barrier_type=_KO_OF[bt]))

    def greeks_one_sided(self, dv_sigma: float = 1e-3) -> Dict[str, float]:
        """Knock-out (or vanilla PDE) greeks with the one-sided volatility bump ``(p(sigma + dv) - p(sigma)) / dv``
        of fd_discrete_barrier_option.py:2076-2088: two solves instead of three."""
        bt = self._check_type()
        if bt not in KO_TYPES and bt != "none":
            raise ValueError("greeks_one_sided: knock-out or vanilla barrier_type only")
        sig = self.sigma
        (V, V_up), _ = self._solve_many([(bt != "none", sig, None), (bt != "none", sig + dv_sigma, None)])
        price = self._interp_price_from_grid(V)
        delta, gamma = self._delta_gamma_from_grid(V)
        vega = (self._interp_price_from_grid(V_up) - price) / dv_sigma
        return {"price": price, "delta": delta, "gamma": gamma, "theta": self._pde_theta(price, delta, gamma),
                "vega": vega}
