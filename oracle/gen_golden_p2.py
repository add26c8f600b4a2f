"""Generates tests/golden/pricer2.json by running the UNMODIFIED ``DiscreteBarrierFDMPricer2`` of the reference
(discrete_barrier_fdm_pricer_2.py).  TEST INFRASTRUCTURE ONLY; build container (needs /root/reference).

    python oracle/gen_golden_p2.py

The class parses and imports as shipped (it needs pandas only).  Its Crank-Nicolson half carries a sign error on
the standard rows (P2:373-375), so a march with theta = 1/2 steps grows without bound (oracle/probe_pricer2.py);
the goldens therefore pin the regime in which the class prices -- every step implicit, ``rannacher_steps >=
num_time_steps`` -- over vanilla / knock-out / double-out / knock-in contracts, the non-symmetric barrier rows, the
BGK continuous window, escrowed dividends, payoff smoothing on and off, three day counts.  Two short marches WITH
theta = 1/2 steps (values already ~1e9) are kept for the CPU oracle only: it reproduces even those bit for bit.
"""
from __future__ import annotations

import json
import os
import sys
from datetime import date, timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pricer2.json")
VAL = date(2025, 7, 28)


def cases():
    mat = date(2026, 7, 28)
    monthly = [VAL + timedelta(days=30 * k) for k in range(1, 13)]
    daily = [VAL + timedelta(days=k) for k in range(1, 366)]
    base = dict(spot=100.0, strike=100.0, valuation_date=VAL, maturity_date=mat, volatility=0.2, option_type="call",
                barrier_type="down-and-out", lower_barrier=90.0, upper_barrier=None, monitoring_dates=monthly,
                flat_rate_nacc=0.05, dividends=[], num_space_nodes=300, num_time_steps=60, rannacher_steps=60,
                day_count="ACT/365", smooth_payoff_around_strike=True, payoff_smoothing_half_width_nodes=2)
    out = []

    def add(cid, greeks=False, gpu=True, **kw):
        out.append(dict(id=cid, greeks=greeks, gpu=gpu, args=dict(base, **kw)))

    add("doc_300x60", greeks=True)
    add("vanilla_call", barrier_type="none", greeks=True)
    add("vanilla_put", barrier_type="none", option_type="put")
    add("uop_120", option_type="put", barrier_type="up-and-out", lower_barrier=None, upper_barrier=120.0, greeks=True)
    add("uoc_130", barrier_type="up-and-out", lower_barrier=None, upper_barrier=130.0)
    add("dop_85", option_type="put", lower_barrier=85.0)
    add("double_out", barrier_type="double-out", upper_barrier=130.0, greeks=True)
    add("down_and_in", barrier_type="down-and-in")
    add("dividend_escrow", dividends=[(date(2025, 12, 1), 0.5), (date(2026, 3, 2), 0.4)])
    add("bgk_daily_40", monitoring_dates=daily, num_time_steps=40, rannacher_steps=40, greeks=True)
    add("bgk_daily_up", monitoring_dates=daily, num_time_steps=40, rannacher_steps=64, option_type="put",
        barrier_type="up-and-out", lower_barrier=None, upper_barrier=115.0)
    add("no_smoothing", smooth_payoff_around_strike=False)
    add("smoothing_5", payoff_smoothing_half_width_nodes=5, strike=103.7)
    add("act360", day_count="ACT/360", maturity_date=date(2026, 1, 28), monitoring_dates=monthly[:6])
    add("thirty360", day_count="30/360", volatility=0.35, flat_rate_nacc=0.08)
    add("grid_600x600", num_space_nodes=600, num_time_steps=600, rannacher_steps=600)
    add("grid_min_200", num_space_nodes=50, num_time_steps=25, rannacher_steps=25)          # N is clamped to 200
    add("grid_1000x100", num_space_nodes=1000, num_time_steps=100, rannacher_steps=100, volatility=0.4)
    add("barrier_off_grid_low", lower_barrier=1e-4)
    add("itm_spot", spot=123.4, strike=95.0, lower_barrier=88.8)
    # theta = 1/2 steps: unstable in the reference; oracle-only
    add("cn_8_steps_unstable", gpu=False, num_time_steps=8, rannacher_steps=2)
    add("cn_12_steps_unstable", gpu=False, num_time_steps=12, rannacher_steps=4, barrier_type="none", option_type="put")
    return out


def main():
    if not ref_loader.reference_available():
        raise SystemExit("needs /root/reference")
    if ref_loader.REFERENCE_DIR not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_DIR)
    import discrete_barrier_fdm_pricer_2 as ref
    recs = []
    for c in cases():
        a = c["args"]
        pr = ref.DiscreteBarrierFDMPricer2(**a)
        Sg, Vg, S_eff = pr._solve_grid_once()
        o = dict(price=pr.price(), S_eff=S_eff, V=[float(v) for v in Vg], s_first=Sg[0], s_second=Sg[1], s_last=Sg[-1],
                 n_nodes=len(Sg), use_bgk=bool(pr.use_bgk_correction), bgk_lower=pr.bgk_lower, bgk_upper=pr.bgk_upper,
                 k_first=pr.k_first_cont, k_last=pr.k_last_cont, tenor=pr.tenor_years, dt=pr.dt,
                 monitor_steps=sorted(pr._monitoring_step_map().keys()))
        if c["greeks"]:
            o["greeks"] = pr.greeks()
        ser = dict(a)
        for k in ("valuation_date", "maturity_date"):
            ser[k] = ser[k].isoformat()
        ser["monitoring_dates"] = [d.isoformat() for d in ser["monitoring_dates"]]
        ser["dividends"] = [[d.isoformat(), x] for d, x in ser["dividends"]]
        recs.append(dict(id=c["id"], gpu=c["gpu"], args=ser, out=o))
        print(c["id"], o["price"], o.get("greeks"))
    with open(OUT, "w") as fh:
        json.dump({"generated_by": "oracle/gen_golden_p2.py", "reference": "discrete_barrier_fdm_pricer_2.py (unmodified)",
                   "cases": recs}, fh)


if __name__ == "__main__":
    main()
