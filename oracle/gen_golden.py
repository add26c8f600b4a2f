"""Generates tests/golden/*.json by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

The reference ships no golden vectors or known-answer tests for this path,
so the parity anchor is "outputs of the reference itself run here": every record below is
produced by importing the reference classes (oracle/ref_loader.py) and calling their public
methods.  The JSON files hold the inputs needed to rebuild each case without the reference.

Exercise intervals: the reference never exposes the exercise set.  ``ex_lo/ex_hi`` stored here
are derived from the reference's final value vector as the first/last interior index k
(0-based in the interior vector, node k+1) with ``V[k+1] == payoff[k+1]`` -- the nodes where
``fd_american_option_pricer.py:496`` selected ``phi_k`` at the last step (an exact tie
``v_candidate == phi_k`` is the only way this differs from ``phi_k > v_candidate``).
"""
from __future__ import annotations

import datetime as dt
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def flat_curve(naca: float):
    """Flat daily NACA curve with ISO dates, as run_american_scenarios.py:135-142 intends."""
    import pandas as pd
    utils = ref_loader.load_utils()
    df = utils.create_rate_df(naca)
    df["Date"] = pd.to_datetime(df["Date"], format="%Y/%m/%d").dt.strftime("%Y-%m-%d")
    return df


def _d(s: str) -> dt.date:
    return dt.date.fromisoformat(s)


# --------------------------------------------------------------------------------------
# American cases
# --------------------------------------------------------------------------------------

def american_cases():
    cases = []
    base = dict(valuation="2025-07-28", maturity="2026-07-28", naca=float(np.exp(0.05) - 1.0),
                spot=100.0, strike=100.0, sigma=0.2, dividends=[], num_space_nodes=400,
                num_time_steps=400, rannacher_steps=2, s_max_mult=4.5, day_count="ACT/365",
                forward_curve=True)
    mods = ["fd_american_option_pricer", "fd_american_equity", "fd_american_black76"]
    # cfg1 (BASELINE config 1) on all three classes, both option types, with the full API
    for mod in mods:
        for ot in ("put", "call"):
            cases.append(dict(id=f"cfg1_{mod}_{ot}", module=mod, option_type=ot, full_api=True,
                              store_s=(mod == "fd_american_option_pricer"), **base))
    # dividends, equity classes only
    divb = dict(base, spot=229.74, strike=220.0, sigma=0.2613190156888, naca=0.073085649282,
                dividends=[["2025-09-12", 7.63], ["2026-04-10", 8.0115]],
                num_space_nodes=500, num_time_steps=500)
    for mod in mods[:2]:
        for ot in ("call", "put"):
            cases.append(dict(id=f"div2_{mod}_{ot}", module=mod, option_type=ot,
                              full_api=(mod == "fd_american_equity"), store_s=False, **divb))
    # one dividend, coarse grid, both types; three dividends
    cases.append(dict(id="div1_coarse_call", module="fd_american_option_pricer", option_type="call",
                      full_api=False, store_s=True,
                      **dict(base, spot=50.0, strike=45.0, sigma=0.35, dividends=[["2026-01-15", 1.25]],
                             num_space_nodes=120, num_time_steps=90)))
    cases.append(dict(id="div3_put", module="fd_american_equity", option_type="put",
                      full_api=False, store_s=False,
                      **dict(base, spot=80.0, strike=95.0, sigma=0.28, naca=0.045,
                             dividends=[["2025-10-01", 0.9], ["2026-01-02", 1.1], ["2026-05-20", 1.3]],
                             num_space_nodes=257, num_time_steps=300)))
    # runner shape (run_american_scenarios.py:283-307): T=31/365, 500x500, put and call
    run = dict(base, valuation="2025-07-28", maturity="2025-08-28", spot=229.74,
               num_space_nodes=500, num_time_steps=500)
    rng = np.random.default_rng(20251018 + 3)
    for i in range(6):
        K = float(rng.uniform(180, 280))
        sig = float(rng.uniform(0.15, 0.45))
        naca = float(rng.uniform(0.03, 0.10))
        ot = "call" if rng.uniform() < 0.5 else "put"
        mod = "fd_american_black76" if i % 2 == 0 else "fd_american_equity"
        cases.append(dict(id=f"runner_{i}_{mod}_{ot}", module=mod, option_type=ot,
                          full_api=(i < 2), store_s=False,
                          **dict(run, strike=K, sigma=sig, naca=naca)))
    # edge grids / parameters
    edge = [
        dict(id="edge_min_grid", num_space_nodes=3, num_time_steps=4, option_type="put"),
        dict(id="edge_n5_m7_call", num_space_nodes=5, num_time_steps=7, option_type="call"),
        dict(id="edge_n33", num_space_nodes=33, num_time_steps=40, option_type="put"),
        dict(id="edge_n416", num_space_nodes=417, num_time_steps=50, option_type="put"),
        dict(id="edge_n513", num_space_nodes=513, num_time_steps=30, option_type="call"),
        dict(id="edge_rannacher0", rannacher_steps=0, num_space_nodes=100, num_time_steps=60, option_type="put"),
        dict(id="edge_rannacher_all", rannacher_steps=80, num_space_nodes=100, num_time_steps=60, option_type="put"),
        dict(id="edge_deep_itm_put", spot=60.0, strike=100.0, num_space_nodes=200, num_time_steps=100, option_type="put"),
        dict(id="edge_deep_otm_put", spot=160.0, strike=100.0, num_space_nodes=200, num_time_steps=100, option_type="put"),
        dict(id="edge_deep_itm_call", spot=160.0, strike=100.0, num_space_nodes=200, num_time_steps=100, option_type="call"),
        dict(id="edge_high_vol", sigma=0.9, num_space_nodes=300, num_time_steps=200, option_type="put"),
        dict(id="edge_low_vol", sigma=0.03, num_space_nodes=300, num_time_steps=200, option_type="put"),
        dict(id="edge_short_T", maturity="2025-08-04", num_space_nodes=200, num_time_steps=50, option_type="call"),
        dict(id="edge_long_T", maturity="2028-07-28", num_space_nodes=300, num_time_steps=300, option_type="put"),
        dict(id="edge_zero_rate", naca=0.0, num_space_nodes=200, num_time_steps=100, option_type="put"),
        dict(id="edge_act360", day_count="ACT/360", num_space_nodes=150, num_time_steps=100, option_type="put"),
        dict(id="edge_no_fwd_curve", forward_curve=False, num_space_nodes=150, num_time_steps=100, option_type="call"),
        dict(id="edge_smax_mult", s_max_mult=9.0, num_space_nodes=250, num_time_steps=100, option_type="put"),
    ]
    for k, e in enumerate(edge):
        mod = mods[k % 3]
        e = dict(e)
        cid = e.pop("id")
        ot = e.pop("option_type")
        cases.append(dict(id=cid + "_" + mod, module=mod, option_type=ot, full_api=False,
                          store_s=(k % 5 == 0), **dict(base, **e)))
    # seeded random sweep on modest grids, all three modules
    rng = np.random.default_rng(20251018)
    for i in range(18):
        mod = mods[i % 3]
        n = int(rng.choice([64, 100, 128, 200, 256, 300]))
        m = int(rng.integers(20, 260))
        S = float(rng.uniform(20, 300))
        K = float(S * rng.uniform(0.7, 1.3))
        days = int(rng.integers(20, 900))
        cases.append(dict(
            id=f"rand_{i}_{mod}", module=mod, option_type="call" if rng.uniform() < 0.5 else "put",
            full_api=False, store_s=False,
            **dict(base, spot=S, strike=K, sigma=float(rng.uniform(0.08, 0.7)),
                   naca=float(rng.uniform(0.0, 0.12)),
                   maturity=(dt.date(2025, 7, 28) + dt.timedelta(days=days)).isoformat(),
                   num_space_nodes=n, num_time_steps=m, rannacher_steps=int(rng.integers(0, 5)))))
    return cases


def run_american(case):
    mod = ref_loader.load_american(case["module"])
    cls = getattr(mod, "AmericanFwdFDMPricer" if case["module"] == "fd_american_black76" else "AmericanFDMPricer")
    curve = flat_curve(case["naca"])
    divs = [(_d(d), float(a)) for d, a in case["dividends"]]
    pr = cls(spot=case["spot"], strike=case["strike"], valuation_date=_d(case["valuation"]),
             maturity_date=_d(case["maturity"]), sigma=case["sigma"], option_type=case["option_type"],
             discount_curve=curve, forward_curve=curve if case["forward_curve"] else None,
             dividend_schedule=divs, day_count=case["day_count"],
             num_space_nodes=case["num_space_nodes"], num_time_steps=case["num_time_steps"],
             rannacher_steps=case["rannacher_steps"], s_max_mult=case["s_max_mult"])
    V = pr._solve_grid()
    price = pr._interp_price(V)
    delta, gamma = pr._local_cubic_delta_gamma(V)
    s = list(pr.s_nodes)
    payoff = [pr._intrinsic_payoff(x) for x in s]
    n = len(s) - 1
    ex = [k for k in range(n - 1) if V[k + 1] == payoff[k + 1]]
    out = dict(
        T=pr.time_to_expiry, r=pr.discount_rate_nacc, b=pr.carry_rate_nacc,
        S_min=pr._S_min, S_max=pr._S_max, dx=pr._dx,
        spot_idx=pr.spot_grid_index, strike_idx=pr.strike_grid_index,
        spot_snapped=pr.spot_snapped, strike_snapped=pr.strike_snapped,
        s_first=s[0], s_last=s[-1],
        price_log=price, delta=delta, gamma=gamma, V=[float(v) for v in V],
        ex_lo=(ex[0] if ex else -1), ex_hi=(ex[-1] if ex else -1),
        ex_contiguous=bool(ex == list(range(ex[0], ex[-1] + 1))) if ex else True,
        div_taus=[[float(t), float(a)] for t, a in pr._div_times_tau()] if hasattr(pr, "_div_times_tau") else [],
    )
    if case["store_s"]:
        out["s_nodes"] = [float(x) for x in s]
    if case["full_api"]:
        out["price_log2"] = pr.price_log2()
        out["price_log2_norich"] = pr.price_log2(use_richardson=False)
        out["greeks_log2"] = pr.greeks_log2()
        out["greeks_log2_norich"] = pr.greeks_log2(use_richardson=False)
        out["greeks_log2_h005"] = pr.greeks_log2(dv_sigma=0.005)
    return out


# --------------------------------------------------------------------------------------
# Barrier cases
# --------------------------------------------------------------------------------------

def barrier_cases():
    cases = []
    monthly = [k / 12.0 for k in range(1, 13)]
    base = dict(S0=100.0, K=100.0, T=1.0, sigma=0.2, r_disc=0.05, b_carry=0.05, option_type="call",
                barrier_type="down-and-out", lower_barrier=90.0, upper_barrier=None, rebate=0.0,
                monitor_times=monthly, N_space=400, N_time=400)
    mods = ["discrete_barrier_fdm_pricer_cn", "fd_discrete_barrier_option"]
    for mod in mods:
        cases.append(dict(id=f"doc_400_{mod}", module=mod, full_api=True, **base))
        cases.append(dict(id=f"doc_auto_{mod}", module=mod, full_api=True,
                          **dict(base, N_space=None, N_time=None)))
        cases.append(dict(id=f"dic_400_{mod}", module=mod, full_api=True,
                          **dict(base, barrier_type="down-and-in")))
        cases.append(dict(id=f"uop_rebate_{mod}", module=mod, full_api=True,
                          **dict(base, option_type="put", barrier_type="up-and-out", lower_barrier=None,
                                 upper_barrier=115.0, rebate=1.5, N_space=300, N_time=350)))
        cases.append(dict(id=f"uip_{mod}", module=mod, full_api=True,
                          **dict(base, option_type="put", barrier_type="up-and-in", lower_barrier=None,
                                 upper_barrier=120.0, N_space=200, N_time=240)))
        cases.append(dict(id=f"none_{mod}", module=mod, full_api=True,
                          **dict(base, barrier_type="none", lower_barrier=None, N_space=200, N_time=200)))
        cases.append(dict(id=f"uoc_{mod}", module=mod, full_api=False,
                          **dict(base, barrier_type="up-and-out", lower_barrier=None, upper_barrier=130.0,
                                 rebate=0.0, N_space=250, N_time=250)))
        cases.append(dict(id=f"dop_nomon_{mod}", module=mod, full_api=False,
                          **dict(base, option_type="put", monitor_times=None, N_space=100, N_time=100)))
        cases.append(dict(id=f"doc_weekly_{mod}", module=mod, full_api=False,
                          **dict(base, monitor_times=[k / 52.0 for k in range(0, 60)], T=1.0,
                                 N_space=200, N_time=None)))
        cases.append(dict(id=f"doc_tiny_{mod}", module=mod, full_api=False,
                          **dict(base, N_space=3, N_time=4)))
    # BASELINE config 2 distribution, reduced to a seeded sample
    rng = np.random.default_rng(20251018 + 2)
    for i in range(16):
        K = float(rng.uniform(80, 120))
        sig = float(rng.uniform(0.10, 0.50))
        r = float(rng.uniform(0.01, 0.08))
        q = float(rng.uniform(0.0, 0.03))
        T = float(rng.choice([0.5, 1.0, 2.0]))
        B = float(100.0 * rng.uniform(0.70, 0.95))
        n_mon = int(round(12 * T))
        big = i < 4
        cases.append(dict(id=f"cfg2_{i}", module=mods[i % 2], full_api=(i < 2),
                          **dict(base, K=K, sigma=sig, r_disc=r, b_carry=r - q, T=T, lower_barrier=B,
                                 monitor_times=[k / 12.0 for k in range(1, n_mon + 1)],
                                 N_space=400 if big else 160, N_time=400 if big else 200)))
    return cases


def run_barrier(case):
    mod = ref_loader.load_barrier(case["module"])
    kw = {k: case[k] for k in ("S0", "K", "T", "sigma", "r_disc", "b_carry", "option_type",
                               "barrier_type", "lower_barrier", "upper_barrier", "rebate",
                               "monitor_times", "N_space", "N_time")}
    pr = mod.DiscreteBarrierCrankNicolsonLog(**kw)
    bt = case["barrier_type"]
    out = {}
    ko_type = {"down-and-in": "down-and-out", "up-and-in": "up-and-out"}.get(bt, bt)
    # the raw KO (or vanilla) PDE solve behind every public method
    pr.barrier_type = ko_type
    V = pr._solve_grid(apply_KO=(ko_type != "none"))
    pr.barrier_type = bt
    dtm = pr.T / pr.N_time
    out.update(
        N_space=pr.N_space, N_time=pr.N_time, S_min=pr._S_min, S_max=pr._S_max,
        s_first=pr.s_nodes[0], s_last=pr.s_nodes[-1],
        monitor_steps=sorted(int(k) for k in pr._monitor_indices_tau(dtm)),
        V=[float(v) for v in V],
        grid_price=pr._interp_price_from_grid(V),
        grid_delta=pr._delta_gamma_from_grid(V)[0],
        grid_gamma=pr._delta_gamma_from_grid(V)[1],
    )
    out["price"] = pr.price()
    if case["full_api"]:
        out["greeks"] = pr.greeks()
        out["greeks_dv5e3"] = pr.greeks(dv_sigma=5e-3)
    return out


def main():
    if not ref_loader.reference_available():
        raise SystemExit("reference checkout not found; goldens can only be generated in the build container")
    os.makedirs(OUT_DIR, exist_ok=True)
    am = []
    for case in american_cases():
        rec = dict(case)
        rec["out"] = run_american(case)
        am.append(rec)
        print("american", case["id"], rec["out"]["price_log"], rec["out"]["ex_lo"], rec["out"]["ex_hi"])
    with open(os.path.join(OUT_DIR, "american.json"), "w") as fh:
        json.dump(dict(generator="oracle/gen_golden.py", reference="rwx-gigaba-sonwabo/Finite_Difference",
                       cases=am), fh)
    ba = []
    for case in barrier_cases():
        rec = dict(case)
        rec["out"] = run_barrier(case)
        ba.append(rec)
        print("barrier", case["id"], rec["out"]["price"])
    with open(os.path.join(OUT_DIR, "barrier.json"), "w") as fh:
        json.dump(dict(generator="oracle/gen_golden.py", reference="rwx-gigaba-sonwabo/Finite_Difference",
                       cases=ba), fh)


if __name__ == "__main__":
    main()
