"""Drop-in for the importable prefix of the reference module ``fd_discrete_barrier_option``
(lines 1-423): same solver as ``discrete_barrier_fdm_pricer_cn``, but vanilla and the knock-in
parity leg come from the PDE (:324-352), ``price()`` is a single solve for knock-outs (:329-331),
and ``greeks()`` bumps through ``price()`` and recurses for knock-ins (:383-423)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

from .barrier import DiscreteBarrierBase, KO_TYPES, _KO_OF


@dataclass
class DiscreteBarrierCrankNicolsonLog(DiscreteBarrierBase):
    def price(self) -> float:
        bt = self._check_type()
        if bt == "none":
            return self._interp_price_from_grid(self._solve_grid(apply_KO=False))
        if bt in KO_TYPES:
            return self._interp_price_from_grid(self._solve_grid(apply_KO=True))
        (V_van, V_ko), _ = self._solve_many([(False, self.sigma, "none"), (True, self.sigma, _KO_OF[bt])])
        return self._interp_price_from_grid(V_van) - self._interp_price_from_grid(V_ko)

    def _direct_greeks(self, bt: str, dv_sigma: float) -> Dict[str, float]:
        # price() of a vanilla / knock-out type is one solve with the bumped sigma on the frozen grid,
        # which is exactly the three-solve batch of the shared base
        return self._pde_price_and_greeks(apply_KO=(bt != "none"), dv_sigma=dv_sigma, barrier_type=bt)

    def greeks(self, dv_sigma: float = 1e-3) -> Dict[str, float]:
        bt = self._check_type()
        if bt == "none" or bt in KO_TYPES:
            return self._direct_greeks(bt, dv_sigma)
        g_van = self._direct_greeks("none", dv_sigma)
        g_ko = self._direct_greeks(_KO_OF[bt], dv_sigma)
        return {key: g_van[key] - g_ko[key] for key in g_van.keys()}
