"""Drop-in for the reference module ``fd_american_black76``: American options on a forward
(``spot`` is F0; drift -sigma^2/2; Black-76 boundaries; dividends ignored; one segment;
fd_american_black76.py:36-160, :291-314, :351, :450-472).  ``price_log2`` refines with
``2*num_time_steps`` (:545) and theta has no carry term (:614-617)."""
from typing import List, Optional

from .american import AmericanPricerBase


class AmericanFwdFDMPricer(AmericanPricerBase):
    _BLACK76 = True

    def _fine_steps_price_log2(self) -> int:
        return 2 * self.num_time_steps

    def _solve_grid(self, N_time: Optional[int] = None) -> List[float]:
        return self._solve_grid_impl(N_time)

    def price_log(self, N_time: Optional[int] = None) -> float:
        return self._price_log(N_time)

    def price_log2(self, apply_KO: bool = True, use_richardson: bool = True) -> float:
        return self._price_log2(use_richardson)

    def _theta(self, sigma0: float, price: float, delta: float, gamma: float) -> float:
        F0 = self.forward0
        return -(0.5 * sigma0 * sigma0 * F0 * F0 * gamma - self.discount_rate_nacc * price)
