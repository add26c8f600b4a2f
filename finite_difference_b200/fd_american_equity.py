"""Drop-in for the reference module ``fd_american_equity`` (the scenario runner's class,
run_american_scenarios.py:20).  Same arithmetic as ``fd_american_option_pricer``; keyword
spellings ``n_time`` / ``apply_ko`` (fd_american_equity.py:913-951)."""
from typing import List, Optional

from .american import AmericanPricerBase


class AmericanFDMPricer(AmericanPricerBase):
    def _solve_grid(self, n_time: Optional[int] = None) -> List[float]:
        return self._solve_grid_impl(n_time)

    def price_log(self, n_time: Optional[int] = None) -> float:
        return self._price_log(n_time)

    def price_log2(self, apply_ko: bool = True, use_richardson: bool = True) -> float:
        return self._price_log2(use_richardson)

    def _price_for_sigma(self, sigma: float, n_time: Optional[int] = None) -> float:
        return super()._price_for_sigma(sigma, n_time)
