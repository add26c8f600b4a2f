"""GPU parity of the DiscreteBarrierFDMPricer drop-in (finite_difference_b200/discrete_barrier_fdm_pricer.py) against
goldens produced by the reference class (tests/golden/dbfp.json): same constructor arguments, same method calls.
Grids from 256 to 2133 space nodes: warp kernel, lean kernel (320 / 406 nodes) and both CTA variants.
Tolerances: every node 1e-10 relative (+1e-12 of the vector's scale); prices 1e-10 relative (north_star);
delta 1e-8, gamma 1e-6; vega is a one-sided 1e-4 bump, (p_up - p)/0.01, which amplifies price noise by 1e2/price:
1e-6 relative with a 1e-9 floor; theta follows gamma."""
import numpy as np
import pytest

import _golden
from oracle import oracle
from test_dbfp_host import pricer

pytestmark = pytest.mark.gpu
DBFP = _golden.dbfp_cases()


def _close(a, b, rel, floor):
    assert abs(a - b) <= rel * max(abs(b), floor), (a, b, abs(a - b) / max(abs(b), floor))


@pytest.mark.parametrize("case", DBFP, ids=[c["id"] for c in DBFP])
def test_solve_grid_value_vector(case):
    o = case["out"]
    pr = pricer(case)
    V = np.array(pr._solve_grid(apply_KO=True))
    ref = np.array(o["V"])
    assert V.shape == ref.shape
    tol = 1e-10 * np.abs(ref) + 1e-12 * np.max(np.abs(ref))
    assert np.all(np.abs(V - ref) <= tol), float(np.max(np.abs(V - ref) / (np.abs(ref) + 1e-300)))
    _close(pr._interp_price(list(V)), o["solve_price"], 1e-10, 1e-8)
    d, g = pr._delta_gamma_from_grid(list(V))
    _close(d, o["solve_delta"], 1e-8, 1e-8)
    _close(g, o["solve_gamma"], 1e-6, 1e-8)


@pytest.mark.parametrize("case", [c for c in DBFP if not c["solve_only"]],
                         ids=[c["id"] for c in DBFP if not c["solve_only"]])
def test_price_log2_and_greeks_log2(case):
    o = case["out"]
    pr = pricer(case)
    # knock-in values are vanilla - knock-out (discrete_barrier_fdm_pricer.py:930-943): the tolerance is relative
    # to the legs, not to a difference that may cancel to ~0
    gv = pr._vanilla_black76_greeks_fd()
    leg = {k: max(abs(gv[k]), abs(o["greeks_log2"][k])) for k in gv}
    _close(pr.price_log2(), o["price_log2"], 1e-10, max(leg["price"], 1e-6))
    g = pr.greeks_log2()
    ref = o["greeks_log2"]
    assert set(g) == set(ref)
    _close(g["price"], ref["price"], 1e-10, max(leg["price"], 1e-6))
    _close(g["delta"], ref["delta"], 1e-8, max(leg["delta"], 1e-6))
    _close(g["gamma"], ref["gamma"], 1e-6, max(leg["gamma"], 1e-6))
    _close(g["theta"], ref["theta"], 1e-6, max(leg["theta"], 1e-4))
    _close(g["vega"], ref["vega"], 1e-6, max(leg["vega"], 1e-5))


def test_gpu_matches_oracle_on_unseen_inputs():
    """Seeded inputs outside the golden set, against the C restatement (which is bit-exact to the reference)."""
    rng = np.random.default_rng(20251018)
    base = next(c for c in DBFP if c["id"] == "syn_rebate_at_expiry")
    for i in range(12):
        case = dict(base, strike=float(rng.uniform(190, 270)), sigma=float(rng.uniform(0.12, 0.5)),
                    naca=float(rng.uniform(0.03, 0.1)), num_time_steps=int(rng.choice([60, 75, 88, 95, 110])),
                    option_type=str(rng.choice(["put", "call"])), rebate_amount=float(rng.uniform(0, 4)),
                    rebate_at_hit=bool(rng.integers(2)))
        if rng.integers(2):
            case.update(barrier_type="up-and-out", lower_barrier=None, upper_barrier=float(rng.uniform(235, 300)))
        else:
            case.update(barrier_type="down-and-out", lower_barrier=float(rng.uniform(170, 225)), upper_barrier=None)
        pr = pricer(case)
        V = np.array(pr._solve_grid(apply_KO=True))
        ref = oracle.dbfp_solve(
            S_min=pr._S_min, S_max=pr._S_max, S0_eff=pr.spot - pr.pv_divs, S_greek=pr.spot, K=pr.strike,
            T=pr.time_to_expiry, sigma=pr.sigma, r=pr.discount_rate_nacc, b=pr.carry_rate_nacc, q=pr.div_yield_nacc,
            is_call=pr.option_type == "call", barrier_type=pr.barrier_type, lower_barrier=pr.lower_barrier,
            upper_barrier=pr.upper_barrier, rebate_amount=pr.rebate_amount, rebate_at_hit=pr.rebate_at_hit,
            monitor_times=pr.monitor_times, num_space_nodes=pr.num_space_nodes, num_time_steps=pr.num_time_steps,
            rannacher_steps=pr.rannacher_steps)
        tol = 1e-10 * np.abs(ref.V) + 1e-12 * np.max(np.abs(ref.V))
        assert np.all(np.abs(V - ref.V) <= tol), (i, float(np.max(np.abs(V - ref.V))))
        _close(pr._interp_price(list(V)), ref.price, 1e-10, 1e-8)


def test_run_config_scenarios_batch_equals_per_row(tmp_path):
    """run_all_scenarios (one GPU batch for the CSV) against run_scenario (per-row drop-in class) and the goldens."""
    import pandas as pd
    from finite_difference_b200 import run_config_scenarios as rcs
    cases = [c for c in DBFP if c["id"].startswith("config_scenarios_1_") and c["num_time_steps"] == 60]
    frame = pd.DataFrame([dict(scenario_name=c["id"], S0=c["spot"], K=c["strike"], sigma=c["sigma"], rate=c["naca"],
                               barrier_type=c["barrier_type"], upper_barrier=c["upper_barrier"],
                               lower_barrier=c["lower_barrier"], FA_price=1.0, FA_delta=None, FA_gamma=0.0, FA_vega=2.0)
                          for c in cases])
    cfg, out = tmp_path / "cfg.csv", tmp_path / "res.csv"
    frame.to_csv(cfg, index=False)
    base = dict(valuation=_golden.d(cases[0]["valuation"]), maturity=_golden.d(cases[0]["maturity"]),
                monitor_dates=_golden.dbfp_monitor_dates(cases[0]), opt_type="put", num_time_steps=60)
    df = rcs.run_all_scenarios(str(cfg), str(out), base)                 # vectorised host path (no dividends)
    assert list(df.columns) == list(pd.read_csv(out).columns) and len(df) == len(cases)
    df_obj = rcs.run_scenarios_frame(pd.read_csv(cfg), base, vectorised=False)      # one pricer object per row
    assert list(df_obj.columns) == list(df.columns)
    for k in ("model_price", "model_delta", "model_gamma", "model_vega"):
        scale = np.maximum(np.abs(df_obj[k].to_numpy()), 30.0 if k == "model_price" else 1e-2)
        assert np.all(np.abs(df[k].to_numpy() - df_obj[k].to_numpy()) <= 1e-7 * scale), k
    assert np.all(np.abs(df["model_price"].to_numpy() - df_obj["model_price"].to_numpy()) <= 1e-12 * 30.0)
    for i, c in enumerate(cases):
        ref = c["out"]["greeks_log2"]
        gv = max(abs(ref["price"]), abs(c["out"]["solve_price"]), 1e-6)
        _close(df.loc[i, "model_price"], c["out"]["price_log2"], 1e-10, gv)
        row = rcs.run_scenario(c["id"], c["spot"], c["strike"], c["sigma"], c["naca"], c["barrier_type"],
                               c["upper_barrier"], c["lower_barrier"], 1.0, None, 0.0, 2.0, **base)
        assert list(row.keys()) == list(df.columns)
        # the batch reads prices in the kernel epilogue (CUDA exp for the node prices), the class on the host
        # (glibc exp): one-ulp differences of the ~29.0 legs, amplified 1e2 by the one-sided vega bump
        _close(df.loc[i, "model_price"], row["model_price"], 1e-13, gv)
        for k, rel in (("model_delta", 1e-9), ("model_gamma", 1e-7), ("model_vega", 1e-7)):
            _close(df.loc[i, k], row[k], rel, max(abs(row[k]), 1e-2))
        assert np.isnan(df.loc[i, "FA_delta"]) and np.isnan(df.loc[i, "gamma_pct_diff"])
