"""GPU parity of the drop-in pricer classes and runners against goldens produced by the unmodified
reference classes (tests/golden/*.json): same constructor arguments, same method calls.
Tolerances: prices 1e-10 relative (north_star); delta 1e-9, gamma/vega/theta 1e-7 relative -- they
amplify price noise through an ill-conditioned 4x4 solve / a 1e-3..1e-2 volatility bump."""
import datetime as dt
import os

import numpy as np
import pytest

import _golden
from finite_difference_b200 import (discrete_barrier_fdm_pricer_cn, fd_american_black76, fd_american_equity,
                                    fd_american_option_pricer, fd_discrete_barrier_option, run_american_scenarios,
                                    utils)

pytestmark = pytest.mark.gpu

AM = _golden.american_cases()
BAR = _golden.barrier_cases()
MODS = {"fd_american_option_pricer": fd_american_option_pricer.AmericanFDMPricer,
        "fd_american_equity": fd_american_equity.AmericanFDMPricer,
        "fd_american_black76": fd_american_black76.AmericanFwdFDMPricer}
BMODS = {"discrete_barrier_fdm_pricer_cn": discrete_barrier_fdm_pricer_cn.DiscreteBarrierCrankNicolsonLog,
         "fd_discrete_barrier_option": fd_discrete_barrier_option.DiscreteBarrierCrankNicolsonLog}
GREEK_TOL = {"price": 1e-10, "delta": 1e-9, "gamma": 1e-7, "vega": 1e-7, "theta": 1e-7}


def _close(a, b, rel, floor=1e-8):
    assert abs(a - b) <= rel * max(abs(b), floor), (a, b, abs(a - b) / max(abs(b), floor))


def _pricer(case):
    curve = utils.create_rate_df_iso(case["naca"])
    divs = [(_golden.d(d), float(a)) for d, a in case["dividends"]]
    return MODS[case["module"]](
        spot=case["spot"], strike=case["strike"], valuation_date=_golden.d(case["valuation"]),
        maturity_date=_golden.d(case["maturity"]), sigma=case["sigma"], option_type=case["option_type"],
        discount_curve=curve, forward_curve=curve if case["forward_curve"] else None, dividend_schedule=divs,
        day_count=case["day_count"], num_space_nodes=case["num_space_nodes"],
        num_time_steps=case["num_time_steps"], rannacher_steps=case["rannacher_steps"], s_max_mult=case["s_max_mult"])


@pytest.mark.parametrize("case", AM, ids=[c["id"] for c in AM])
def test_american_class_price_log_and_attributes(case):
    o = case["out"]
    pr = _pricer(case)
    assert pr.time_to_expiry == o["T"] and pr.discount_rate_nacc == o["r"] and pr.carry_rate_nacc == o["b"]
    V = pr._solve_grid()
    assert isinstance(V, list) and len(V) == max(case["num_space_nodes"], 3) + 1
    assert (pr._S_min, pr._S_max, pr._dx) == (o["S_min"], o["S_max"], o["dx"])
    assert (pr.spot_grid_index, pr.strike_grid_index) == (o["spot_idx"], o["strike_idx"])
    assert (pr.spot_snapped, pr.strike_snapped) == (o["spot_snapped"], o["strike_snapped"])
    assert pr.s_nodes[0] == o["s_first"] and pr.s_nodes[-1] == o["s_last"]
    _close(pr._interp_price(V), o["price_log"], 1e-10, 1e-300)
    d, g = pr._local_cubic_delta_gamma(V)
    _close(d, o["delta"], 1e-8)
    _close(g, o["gamma"], 1e-7)
    assert pr.last_exercise_interval == (o["ex_lo"], o["ex_hi"])
    _close(pr.price_log(), o["price_log"], 1e-10, 1e-300)
    if hasattr(pr, "_div_times_tau"):
        assert [list(x) for x in pr._div_times_tau()] == o["div_taus"]


FULL = [c for c in AM if c["full_api"]]


@pytest.mark.parametrize("case", FULL, ids=[c["id"] for c in FULL])
def test_american_class_full_api(case):
    o = case["out"]
    pr = _pricer(case)
    _close(pr.price_log2(), o["price_log2"], 1e-10)
    _close(pr.price_log2(use_richardson=False), o["price_log2_norich"], 1e-10)
    for kwargs, key in (({}, "greeks_log2"), ({"use_richardson": False}, "greeks_log2_norich"),
                        ({"dv_sigma": 0.005}, "greeks_log2_h005")):
        g = pr.greeks_log2(**kwargs)
        assert list(g) == ["price", "delta", "gamma", "vega", "theta"]
        for k, ref in o[key].items():
            _close(g[k], ref, GREEK_TOL[k])
    assert pr.sigma == case["sigma"]            # bumps restore the volatility


def test_keyword_spellings_of_the_three_modules():
    c = utils.create_rate_df_iso(0.05)
    args = (100.0, 100.0, dt.date(2025, 7, 28), dt.date(2026, 1, 28), 0.25, "put", c, c)
    kw = dict(num_space_nodes=64, num_time_steps=40)
    a = fd_american_option_pricer.AmericanFDMPricer(*args, **kw)
    e = fd_american_equity.AmericanFDMPricer(*args, **kw)
    b = fd_american_black76.AmericanFwdFDMPricer(*args, **kw)
    assert a.price_log(N_time=50) == e.price_log(n_time=50)
    assert a.price_log2(apply_KO=True) == e.price_log2(apply_ko=True)
    assert b.price_log(N_time=50) > 0 and b.forward0 == 100.0 and b.dividend_schedule == []
    with pytest.raises(TypeError):
        e.price_log(N_time=50)


def test_reference_error_conventions():
    c = utils.create_rate_df_iso(0.05)
    v, m = dt.date(2025, 7, 28), dt.date(2026, 1, 28)
    for cls in MODS.values():
        with pytest.raises(ValueError):
            cls(-1.0, 100.0, v, m, 0.2, "put", c)
        with pytest.raises(ValueError):
            cls(100.0, 100.0, m, v, 0.2, "put", c)
        with pytest.raises(ValueError):
            cls(100.0, 100.0, v, m, 0.2, "straddle", c)
        with pytest.raises(ValueError, match="Discount factor not found"):
            cls(100.0, 100.0, v, dt.date(2031, 1, 1), 0.2, "put", c)


def _barrier(case):
    kw = {k: case[k] for k in ("S0", "K", "T", "sigma", "r_disc", "b_carry", "option_type", "barrier_type",
                               "lower_barrier", "upper_barrier", "rebate", "monitor_times", "N_space", "N_time")}
    return BMODS[case["module"]](**kw)


@pytest.mark.parametrize("case", BAR, ids=[c["id"] for c in BAR])
def test_barrier_class_price_and_greeks(case):
    o = case["out"]
    pr = _barrier(case)
    _close(pr.price(), o["price"], 1e-10)
    pr.configure_grid()        # the closed-form vanilla path of the _cn module never builds the grid
    assert (pr.N_space, pr.N_time) == (o["N_space"], o["N_time"])
    assert (pr._S_min, pr._S_max) == (o["S_min"], o["S_max"])
    assert pr.barrier_type == case["barrier_type"] and pr.sigma == case["sigma"]
    if case["full_api"]:
        for kwargs, key in (({}, "greeks"), ({"dv_sigma": 5e-3}, "greeks_dv5e3")):
            g = pr.greeks(**kwargs)
            for k, ref in o[key].items():
                _close(g[k], ref, 2e-7 if k == "vega" else 1e-9, 1e-6)


def test_barrier_class_rejects_unknown_type_and_bad_inputs():
    cls = BMODS["discrete_barrier_fdm_pricer_cn"]
    with pytest.raises(ValueError, match="Unsupported barrier_type"):
        cls(100.0, 100.0, 1.0, 0.2, 0.05, 0.05, "call", "double-out").price()
    with pytest.raises(ValueError):
        cls(100.0, 100.0, -1.0, 0.2, 0.05, 0.05, "call", "down-and-out", lower_barrier=90.0).price()


def test_scenario_runner_batch_equals_single_row_path(tmp_path):
    """run_all_american_scenarios (one GPU batch) == run_american_scenario per row (pricer class)."""
    import pandas as pd
    rng = np.random.default_rng(5)
    n = 6
    cfg = pd.DataFrame({"scenario_name": [f"scenario_{i}" for i in range(n)], "S0": 229.74,
                        "K": rng.uniform(200, 260, n), "sigma": rng.uniform(0.15, 0.45, n),
                        "rate": rng.uniform(0.03, 0.10, n), "FA_price": rng.uniform(1, 20, n),
                        "FA_delta": [np.nan] + list(rng.uniform(-1, 0, n - 1)), "FA_gamma": 0.01, "FA_vega": 0.2})
    path = os.path.join(tmp_path, "american_config_scenarios.csv")
    cfg.to_csv(path, index=False)
    base = dict(valuation=dt.date(2025, 7, 28), maturity=dt.date(2025, 8, 28), opt_type="put", trade_number=1,
                quantity=1000, contract_size=1, position="long", divs=[(dt.date(2025, 8, 10), 1.5)],
                underlying_spot_days=0, option_days=0, option_settlement_days=0, day_count="ACT/365",
                grid_type="uniform", num_space_nodes=200, num_time_steps=120, rannacher_steps=2)
    out = os.path.join(tmp_path, "results.csv")
    df = run_american_scenarios.run_all_american_scenarios(path, out, base)
    assert os.path.exists(out) and len(df) == n
    assert list(df.columns)[:9] == ["scenario_name", "S0", "K", "sigma", "rate", "model_price", "FA_price",
                                    "price_diff", "price_pct_diff"]
    assert np.isnan(df["delta_diff"][0]) and np.isnan(df["delta_pct_diff"][0])
    for i in range(n):
        row = cfg.iloc[i]
        one = run_american_scenarios.run_american_scenario(
            row["scenario_name"], row["S0"], row["K"], row["sigma"], row["rate"], row["FA_price"], None, 0.01, 0.2,
            **base)
        _close(df["model_price"][i], one["model_price"], 1e-11)
        _close(df["model_delta"][i], one["model_delta"], 1e-8)
        _close(df["model_gamma"][i], one["model_gamma"], 1e-6)
        _close(df["model_vega"][i], one["model_vega"], 1e-8)
