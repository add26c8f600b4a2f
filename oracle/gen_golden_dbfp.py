"""Generates tests/golden/dbfp.json by running the reference's DiscreteBarrierFDMPricer.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_dbfp.py            (build container only: needs /root/reference)

discrete_barrier_fdm_pricer.py does not parse as shipped; oracle/ref_loader.load_dbfp() execs the file's own
lines 1-745 + 883-1027 unmodified and adds the one attribute its constructor calls but no line defines
(_build_stock_price_grid, result unused).  Every number below comes out of that class: the constructor-derived
scalars, the value vector of _solve_grid(apply_KO=True), price_log2() and greeks_log2() -- the two calls
run_config_scenarios.run_scenario makes (run_config_scenarios.py:97-98).

Cases: the rows of the shipped config_scenarios*.csv at reduced grids (the runner's 500 time steps give
2133 x 500 grids, ~25 s per scenario in the reference; two rows are kept at full size) plus synthetic rows for
the features the CSVs do not reach (rebates paid at expiry, escrowed dividends, double-out, calls).
"""
from __future__ import annotations

import datetime as dt
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.gen_golden import OUT_DIR, flat_curve  # noqa: E402

VAL, MAT = "2025-07-28", "2025-08-28"


def business_days(start: dt.date, end: dt.date):
    """The monitoring calendar of run_config_scenarios.py:199-222 (every weekday from valuation to maturity)."""
    out, d = [], start
    while d <= end:
        if d.weekday() < 5:
            out.append(d)
        d += dt.timedelta(days=1)
    return out


def cases():
    import pandas as pd
    out = []
    base = dict(valuation=VAL, maturity=MAT, option_type="put", dividends=[], rebate_amount=0.0, rebate_at_hit=True,
                already_hit=False, already_in=False, rannacher_steps=2, day_count="ACT/365", solve_only=False,
                monitor="weekdays")
    # shipped scenario files (run_config_scenarios.py:248-252 reads config_scenarios_1.csv with opt_type "put")
    for fname, nt_list in (("config_scenarios.csv", (75,)), ("config_scenarios_1.csv", (60, 95))):
        cfg = pd.read_csv(os.path.join(ref_loader.REFERENCE_DIR, fname))
        for i, row in cfg.iterrows():
            nt = nt_list[i % len(nt_list)]
            out.append(dict(base, id=f"{fname[:-4]}_{row.scenario_name}_nt{nt}", spot=float(row.S0), strike=float(row.K),
                            sigma=float(row.sigma), naca=float(row.rate), barrier_type=str(row.barrier_type),
                            lower_barrier=None if pd.isna(row.lower_barrier) else float(row.lower_barrier),
                            upper_barrier=None if pd.isna(row.upper_barrier) else float(row.upper_barrier),
                            num_time_steps=nt))
    syn = dict(base, spot=229.74, strike=220.0, sigma=0.25, naca=0.073085649282, lower_barrier=200.0, upper_barrier=None,
               barrier_type="down-and-out", num_time_steps=75)
    out += [
        dict(syn, id="syn_doc_put_full500", num_time_steps=500),                      # the runner's own grid: 2133 x 500
        dict(syn, id="syn_uoc_call_full500", option_type="call", barrier_type="up-and-out", lower_barrier=None,
             upper_barrier=260.0, num_time_steps=500),
        dict(syn, id="syn_doc_put_nt130", num_time_steps=130),
        dict(syn, id="syn_rebate_at_expiry", rebate_amount=3.0, rebate_at_hit=False),
        dict(syn, id="syn_rebate_at_hit", rebate_amount=3.0, rebate_at_hit=True),
        dict(syn, id="syn_uoc_call_rebate", option_type="call", barrier_type="up-and-out", lower_barrier=None,
             upper_barrier=260.0, rebate_amount=2.5, rebate_at_hit=False, num_time_steps=95),
        dict(syn, id="syn_dividend", dividends=[("2025-08-11", 4.0)], num_time_steps=95),
        dict(syn, id="syn_dividend_call_ui", option_type="call", barrier_type="up-and-in", lower_barrier=None,
             upper_barrier=255.0, dividends=[("2025-08-11", 4.0), ("2025-08-20", 1.25)]),
        dict(syn, id="syn_double_out_solve", barrier_type="double-out", upper_barrier=260.0, rebate_amount=1.5,
             rebate_at_hit=False, solve_only=True),
        dict(syn, id="syn_none_vanilla", barrier_type="none", lower_barrier=None),
        dict(syn, id="syn_already_hit", already_hit=True, rebate_amount=2.0),
        dict(syn, id="syn_already_in", barrier_type="down-and-in", already_in=True),
        dict(syn, id="syn_rannacher0_monthly", rannacher_steps=0, monitor="weekly", maturity="2025-10-28", num_time_steps=95),
        dict(syn, id="syn_act360", day_count="ACT/360", num_time_steps=60),
    ]
    return out


def monitor_dates(case):
    v, m = dt.date.fromisoformat(case["valuation"]), dt.date.fromisoformat(case["maturity"])
    days = business_days(v, m)
    return days if case["monitor"] == "weekdays" else days[::5]


def build(mod, case):
    curve = flat_curve(case["naca"])
    return mod.DiscreteBarrierFDMPricer(
        spot=case["spot"], strike=case["strike"], valuation_date=dt.date.fromisoformat(case["valuation"]),
        maturity_date=dt.date.fromisoformat(case["maturity"]), sigma=case["sigma"], option_type=case["option_type"],
        barrier_type=case["barrier_type"], lower_barrier=case["lower_barrier"], upper_barrier=case["upper_barrier"],
        already_in=case["already_in"], already_hit=case["already_hit"], monitor_dates=monitor_dates(case),
        discount_curve=curve, forward_curve=curve.copy(),
        dividend_schedule=[(dt.date.fromisoformat(d), a) for d, a in case["dividends"]],
        trade_id=201871103, direction="long", quantity=1000, underlying_spot_days=0, option_days=0,
        option_settlement_days=0, rebate_amount=case["rebate_amount"], rebate_at_hit=case["rebate_at_hit"],
        contract_multiplier=1, use_one_sided_greeks_near_barrier=False, num_space_nodes=500,
        num_time_steps=case["num_time_steps"], grid_type="uniform", rannacher_steps=case["rannacher_steps"],
        restart_on_monitoring=False, mollify_final=False, mollify_band_nodes=2, day_count=case["day_count"])


def run(mod, case):
    pr = build(mod, case)
    o = dict(T=pr.time_to_expiry, r=pr.discount_rate_nacc, b=pr.carry_rate_nacc, q=pr.div_yield_nacc,
             pv_divs=pr.pv_divs, forward_price=pr.forward_price, monitor_times=list(pr.monitor_times))
    V = pr._solve_grid(apply_KO=True)
    o.update(S_min=pr._S_min, S_max=pr._S_max, num_space_nodes=pr.num_space_nodes, v_len=len(V),
             s_first=pr.s_nodes[0], s_last=pr.s_nodes[-1], V=[float(v) for v in V],
             solve_price=pr._interp_price(V))
    d, g = pr._delta_gamma_from_grid(V)
    o.update(solve_delta=d, solve_gamma=g)
    if not case["solve_only"]:
        pr = build(mod, case)
        o["price_log2"] = float(pr.price_log2())
        o["greeks_log2"] = {k: float(v) for k, v in pr.greeks_log2().items()}
    return o


def main():
    mod = ref_loader.load_dbfp()
    recs = []
    for case in cases():
        rec = dict(case)
        rec["out"] = run(mod, case)
        recs.append(rec)
        print(case["id"], rec["out"]["num_space_nodes"], rec["out"]["solve_price"], rec["out"].get("price_log2"))
    with open(os.path.join(OUT_DIR, "dbfp.json"), "w") as fh:
        json.dump(dict(generator="oracle/gen_golden_dbfp.py", reference="rwx-gigaba-sonwabo/Finite_Difference",
                       source="discrete_barrier_fdm_pricer.py lines 1-745 + 883-1027 (+ no-op _build_stock_price_grid)",
                       cases=recs), fh)


if __name__ == "__main__":
    main()
