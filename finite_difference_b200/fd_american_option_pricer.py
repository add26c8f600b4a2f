"""Drop-in for the reference module ``fd_american_option_pricer`` (same class name, constructor,
methods and keyword spellings: ``N_time``, ``apply_KO``; fd_american_option_pricer.py:659-672)."""
from typing import List, Optional

from .american import AmericanPricerBase


class AmericanFDMPricer(AmericanPricerBase):
    def _solve_grid(self, N_time: Optional[int] = None) -> List[float]:
        return self._solve_grid_impl(N_time)

    def price_log(self, N_time: Optional[int] = None) -> float:
        return self._price_log(N_time)

    def price_log2(self, apply_KO: bool = True, use_richardson: bool = True) -> float:
        return self._price_log2(use_richardson)
