"""Helpers shared by the tests: load golden records and turn them into oracle / product inputs."""
import datetime as dt
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as fh:
        return json.load(fh)["cases"]


def american_cases():
    return load("american")


def barrier_cases():
    return load("barrier")


def dbfp_cases():
    return load("dbfp")


def dbfp_monitor_dates(case):
    """Monitoring calendar of the goldens (oracle/gen_golden_dbfp.py): every weekday (or every fifth) from
    valuation to maturity, as run_config_scenarios.py:199-222 lists them."""
    v, m = d(case["valuation"]), d(case["maturity"])
    days, cur = [], v
    while cur <= m:
        if cur.weekday() < 5:
            days.append(cur)
        cur += dt.timedelta(days=1)
    return days if case["monitor"] == "weekdays" else days[::5]


def dbfp_oracle_kwargs(case):
    o = case["out"]
    return dict(
        S_min=o["S_min"], S_max=o["S_max"], S0_eff=case["spot"] - o["pv_divs"], S_greek=case["spot"], K=case["strike"],
        T=o["T"], sigma=case["sigma"], r=o["r"], b=o["b"], q=o["q"], is_call=case["option_type"] == "call",
        barrier_type=case["barrier_type"], lower_barrier=case["lower_barrier"], upper_barrier=case["upper_barrier"],
        rebate_amount=case["rebate_amount"], rebate_at_hit=case["rebate_at_hit"], monitor_times=o["monitor_times"],
        num_space_nodes=o["num_space_nodes"], num_time_steps=case["num_time_steps"],
        rannacher_steps=case["rannacher_steps"])


def am_model(case):
    return 1 if case["module"] == "fd_american_black76" else 0


def am_oracle_kwargs(case, n_time=None, sigma=None):
    o = case["out"]
    model = am_model(case)
    return dict(
        spot=case["spot"], strike=case["strike"], T=o["T"], sigma=case["sigma"] if sigma is None else sigma,
        r=o["r"], b=o["b"], is_call=case["option_type"] == "call", model=model,
        n_space=max(int(case["num_space_nodes"]), 3),
        n_time=max(int(case["num_time_steps"]), 4) if n_time is None else n_time,
        rannacher=max(int(case["rannacher_steps"]), 0), s_max_mult=case["s_max_mult"],
        divs_tau=[] if model == 1 else [tuple(x) for x in o["div_taus"]],
    )


def bar_oracle_kwargs(case, sigma=None, barrier_type=None, N_space=None, N_time=None):
    bt = case["barrier_type"] if barrier_type is None else barrier_type
    return dict(
        S0=case["S0"], K=case["K"], T=case["T"], sigma=case["sigma"] if sigma is None else sigma,
        r_disc=case["r_disc"], b_carry=case["b_carry"], is_call=case["option_type"] == "call",
        barrier_type=bt, lower_barrier=case["lower_barrier"], upper_barrier=case["upper_barrier"],
        rebate=case["rebate"], monitor_times=case["monitor_times"],
        N_space=case["N_space"] if N_space is None else N_space,
        N_time=case["N_time"] if N_time is None else N_time,
        apply_KO=(bt != "none"),
    )


def d(s):
    return dt.date.fromisoformat(s)
