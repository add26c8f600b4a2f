"""Drop-in for the importable prefix of the reference module ``discrete_barrier_fdm_pricer_cn``
(lines 1-537): vanilla = closed-form Black-Scholes (:359-423), knock-out = CN PDE (3 solves, even
for ``price()``, :486-487), knock-in = Black-Scholes minus knock-out (:489-500, :522-535)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

from .barrier import DiscreteBarrierBase, KO_TYPES, _KO_OF, norm_cdf, norm_pdf  # noqa: F401


@dataclass
class DiscreteBarrierCrankNicolsonLog(DiscreteBarrierBase):
    def _vanilla_bs_price_and_greeks(self) -> Dict[str, float]:
        S0, K, T, sigma, r, b = self.S0, self.K, self.T, self.sigma, self.r_disc, self.b_carry
        if T <= 0.0 or sigma <= 0.0:
            price = max(S0 - K, 0.0) if self.option_type.lower() == "call" else max(K - S0, 0.0)
            return {"price": price, "delta": 0.0, "gamma": 0.0, "theta": 0.0, "vega": 0.0}
        q = r - b
        sqrtT = math.sqrt(T)
        d1 = (math.log(S0 / K) + (b + 0.5 * sigma * sigma) * T) / (sigma * sqrtT)
        d2 = d1 - sigma * sqrtT
        Nd1, Nd2, nd1 = norm_cdf(d1), norm_cdf(d2), norm_pdf(d1)
        disc_q = math.exp(-q * T)
        disc_r = math.exp(-r * T)
        if self.option_type.lower() == "call":
            price = S0 * disc_q * Nd1 - K * disc_r * Nd2
            delta = disc_q * Nd1
        else:
            price = K * disc_r * norm_cdf(-d2) - S0 * disc_q * norm_cdf(-d1)
            delta = disc_q * (Nd1 - 1.0)
        gamma = disc_q * nd1 / (S0 * sigma * sqrtT)
        vega = S0 * disc_q * nd1 * sqrtT
        theta = -(0.5 * sigma * sigma * S0 * S0 * gamma + b * S0 * delta - r * price)
        return {"price": price, "delta": delta, "gamma": gamma, "theta": theta, "vega": vega}

    def price(self) -> float:
        bt = self._check_type()
        if bt == "none":
            return self._vanilla_bs_price_and_greeks()["price"]
        if bt in KO_TYPES:
            return self._pde_price_and_greeks(apply_KO=True, dv_sigma=1e-3)["price"]
        g_van = self._vanilla_bs_price_and_greeks()
        g_ko = self._pde_price_and_greeks(apply_KO=True, dv_sigma=1e-3, barrier_type=_KO_OF[bt])
        return g_van["price"] - g_ko["price"]

    def greeks(self, dv_sigma: float = 1e-3) -> Dict[str, float]:
        bt = self._check_type()
        if bt == "none":
            return self._vanilla_bs_price_and_greeks()
        if bt in KO_TYPES:
            return self._pde_price_and_greeks(apply_KO=True, dv_sigma=dv_sigma)
        g_van = self._vanilla_bs_price_and_greeks()
        g_ko = self._pde_price_and_greeks(apply_KO=True, dv_sigma=dv_sigma, barrier_type=_KO_OF[bt])
        return {k: g_van[k] - g_ko[k] for k in g_van}
