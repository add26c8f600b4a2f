"""Evidence for SURVEY.md section 8f-3: the reference's S-space pricer diverges on its own defaults.
TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).

``discrete_barrier_fdm_pricer_2.py`` (class ``DiscreteBarrierFDMPricer2``) parses and runs, but the explicit half of
its Crank-Nicolson step has the wrong sign on both off-diagonals of every standard row
(``:373-375``: ``a_expl = -0.5*dt*(1-theta)*(...)``, ``c_expl = -0.5*dt*(1-theta)*(...)`` where ``+(1-theta)*dt*L``
is meant, as the two barrier rows at ``:399-401`` do).  With theta = 1/2 the step matrix ``A^-1 B`` then has
eigenvalues outside the unit circle, and the value vector grows geometrically after the Rannacher steps: the class
returns ``nan`` on its default 600 x 600 grid and 1e13..1e88 on coarser ones.  Only a fully implicit march
(``rannacher_steps >= num_time_steps``, where ``B = I``) gives prices.

There is therefore no meaningful parity target for a row-varying S-space kernel: a drop-in would have to reproduce
an overflow.  The numbers below are what this script printed in the build container (CPython 3.12.3):

    600x600 default (rannacher 2)        down-and-out call  nan          vanilla call  nan
    600x600 fully implicit (rannacher M) down-and-out call  8.636511688471709   vanilla call 10.551156887437646
    200x200 rannacher 2                  down-and-out call  5.037119757715134e+88
    200x50  rannacher 2                  down-and-out call  2.524368729025166e+34
    200x20  rannacher 2                  down-and-out call  50629537317308.65

Run:  python oracle/probe_pricer2.py
"""
from __future__ import annotations

import sys
from datetime import date, timedelta

from ref_loader import REFERENCE_DIR


def main() -> None:
    sys.path.insert(0, REFERENCE_DIR)
    import discrete_barrier_fdm_pricer_2 as ref   # noqa: E402  (parses as shipped; needs only pandas)

    val, mat = date(2025, 7, 28), date(2026, 7, 28)
    mons = [val + timedelta(days=30 * k) for k in range(1, 13)]
    cases = [
        ("600x600 default (rannacher 2)", dict()),
        ("600x600 fully implicit", dict(rannacher_steps=600)),
        ("600x600 vanilla, rannacher 2", dict(barrier_type="none")),
        ("600x600 vanilla, fully implicit", dict(barrier_type="none", rannacher_steps=600)),
        ("200x200 rannacher 2", dict(num_space_nodes=200, num_time_steps=200)),
        ("200x50 rannacher 2", dict(num_space_nodes=200, num_time_steps=50)),
        ("200x20 rannacher 2", dict(num_space_nodes=200, num_time_steps=20)),
    ]
    for name, kw in cases:
        args = dict(barrier_type="down-and-out", lower_barrier=90.0, monitoring_dates=mons, flat_rate_nacc=0.05)
        args.update(kw)
        p = ref.DiscreteBarrierFDMPricer2(100.0, 100.0, val, mat, 0.2, "call", **args)
        print(f"{name:36s} price = {p.price()!r}")


if __name__ == "__main__":
    main()
