"""Engine options for the barrier dataclass beyond the importable reference prefixes.

The reference's ``fd_discrete_barrier_option.py`` continues after its importable dataclass (lines 1-423) with about
fifteen pasted variants of the same class that do not parse (:425-2421).  They add, between them: fully implicit
Rannacher start-up steps (:912-1007, :1133-1236), a rebate discounted from expiry (:851-867), ``double-out``
(:887-893), Richardson extrapolation over N and N/2 time steps (:1387-1427, :1979-2005) and a one-sided volatility bump
(:1701-1723, :2076-2088).  ``DiscreteBarrierCrankNicolsonLogEx`` offers those on the same CUDA path; with its defaults
it IS the importable dataclass of ``fd_discrete_barrier_option`` (same fields, same results).  The strict drop-in
classes keep the reference's behaviour, including ``ValueError`` for ``double-out``.

No runnable reference exists for these options, so they are checked against the extended CPU oracle
(``fdo_barrier_solve_ext``), which keeps the pinned path's operation order: parity unpinned, stated as such.
"""
from __future__ import annotations

from dataclasses import dataclass

from .fd_discrete_barrier_option import DiscreteBarrierCrankNicolsonLog


@dataclass
class DiscreteBarrierCrankNicolsonLogEx(DiscreteBarrierCrankNicolsonLog):
    rannacher_steps: int = 0           # theta = 1 for the first steps of the march, then 1/2
    rebate_discounted: bool = False    # a knocked-out node is worth rebate * exp(-r_disc * tau): paid at expiry

    _TYPES = DiscreteBarrierCrankNicolsonLog._TYPES + ("double-out", "double-in")
