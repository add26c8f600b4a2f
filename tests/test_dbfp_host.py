"""Host side of the DiscreteBarrierFDMPricer drop-in (no GPU): constructor-derived scalars, the norm.ppf-sized grid,
monitor mapping and the descriptor the C ABI receives, against goldens produced by the reference class
(tests/golden/dbfp.json, oracle/gen_golden_dbfp.py)."""
import math

import numpy as np
import pytest

import _golden
from finite_difference_b200 import _lib, discrete_barrier_fdm_pricer as dbfp, hostlogic, utils

DBFP = _golden.dbfp_cases()


def pricer(case):
    curve = utils.create_rate_df_iso(case["naca"])
    return dbfp.DiscreteBarrierFDMPricer(
        spot=case["spot"], strike=case["strike"], valuation_date=_golden.d(case["valuation"]),
        maturity_date=_golden.d(case["maturity"]), sigma=case["sigma"], option_type=case["option_type"],
        barrier_type=case["barrier_type"], lower_barrier=case["lower_barrier"], upper_barrier=case["upper_barrier"],
        already_in=case["already_in"], already_hit=case["already_hit"], monitor_dates=_golden.dbfp_monitor_dates(case),
        discount_curve=curve, forward_curve=curve.copy(),
        dividend_schedule=[(_golden.d(d), a) for d, a in case["dividends"]], trade_id=201871103, direction="long",
        quantity=1000, underlying_spot_days=0, option_days=0, option_settlement_days=0,
        rebate_amount=case["rebate_amount"], rebate_at_hit=case["rebate_at_hit"], contract_multiplier=1,
        use_one_sided_greeks_near_barrier=False, num_space_nodes=500, num_time_steps=case["num_time_steps"],
        grid_type="uniform", rannacher_steps=case["rannacher_steps"], restart_on_monitoring=False, mollify_final=False,
        mollify_band_nodes=2, day_count=case["day_count"])


@pytest.mark.parametrize("case", DBFP, ids=[c["id"] for c in DBFP])
def test_constructor_scalars_and_grid_identical(case):
    o = case["out"]
    pr = pricer(case)
    assert (pr.time_to_expiry, pr.discount_rate_nacc, pr.carry_rate_nacc) == (o["T"], o["r"], o["b"])
    assert (pr.div_yield_nacc, pr.pv_divs, pr.forward_price) == (o["q"], o["pv_divs"], o["forward_price"])
    assert pr.monitor_times == o["monitor_times"]
    pr._build_log_grid()
    assert (pr._S_min, pr._S_max, pr.num_space_nodes) == (o["S_min"], o["S_max"], o["num_space_nodes"])
    assert pr.s_nodes[0] == o["s_first"] and pr.s_nodes[-1] == o["s_last"] and len(pr.s_nodes) == o["num_space_nodes"] + 1


def test_descriptor_carries_the_reference_quirks():
    case = next(c for c in DBFP if c["id"] == "syn_rebate_at_expiry")
    pr = pricer(case)
    pr._build_log_grid()
    b = pr._solve_descriptor(apply_KO=True)
    n = pr.num_space_nodes
    assert int(b.n_space[0]) == n - 1                                   # discrete_barrier_fdm_pricer.py:448 with :540
    fl = int(b.flags[0])
    assert fl & _lib.FLAG_BDRY_TOP_NEXT and fl & _lib.FLAG_PUT_BDRY_SMIN and fl & _lib.FLAG_KO_DOWN
    assert not fl & (_lib.FLAG_CALL | _lib.FLAG_AMERICAN | _lib.FLAG_SNAP | _lib.FLAG_KO_UP)
    assert b.b[0] == pr.carry_rate_nacc - pr.div_yield_nacc and b.b_bdry[0] == pr.carry_rate_nacc
    assert int(b.rannacher[0]) == 2 and int(b.seg_restart[0]) == 1
    steps = hostlogic.dbfp_monitor_steps(pr.time_to_expiry, pr.time_to_expiry / pr.num_time_steps, pr.num_time_steps,
                                         pr.monitor_times)
    assert list(b.mon_step) == steps and steps[0] == 1                  # the expiry monitor is clamped to step 1, :409
    dt = pr.time_to_expiry / pr.num_time_steps
    np.testing.assert_array_equal(b.mon_rebate, [3.0 * math.exp(-pr.carry_rate_nacc * (k * dt)) for k in steps])
    # knock-in types and apply_KO=False march without projection (:415-416)
    assert int(pr._solve_descriptor(apply_KO=False).flags[0]) & (_lib.FLAG_KO_DOWN | _lib.FLAG_KO_UP) == 0
    assert len(pr._solve_descriptor(apply_KO=False).mon_step) == 0


def test_double_out_uses_both_levels():
    case = next(c for c in DBFP if c["id"] == "syn_double_out_solve")
    pr = pricer(case)
    pr._build_log_grid()
    b = pr._solve_descriptor(apply_KO=True)
    assert int(b.flags[0]) & _lib.FLAG_KO_DOWN and int(b.flags[0]) & _lib.FLAG_KO_UP
    assert (b.barrier[0], b.barrier_up[0]) == (200.0, 260.0)


def test_reference_exceptions():
    case = dict(DBFP[0])
    with pytest.raises(ValueError, match="must be positive"):
        pricer(dict(case, sigma=0.0))
    with pytest.raises(ValueError, match="after valuation_date"):
        pricer(dict(case, maturity=case["valuation"]))
    pr = pricer(dict(case, barrier_type="double-out"))
    with pytest.raises(ValueError, match="Unsupported barrier_type"):      # discrete_barrier_fdm_pricer.py:945
        pr.price_log2()


def test_launch_hints_families_and_size_classes():
    """fdm_launch_hints (pure host): kernel families, dividend flag, largest warp-sized grid, lean size classes."""
    from finite_difference_b200 import hostlogic
    am = hostlogic.build_american_batch([100.0] * 3, [100.0] * 3, [1.0] * 3, [0.2] * 3, [0.05] * 3, [0.05] * 3,
                                        [False, True, False], [400, 500, 128], [50, 50, 50])
    m = am.family_mask()
    assert m & 1 and not m & 2 and not m & 4 and m & 8 and not m & 16          # American only, no dividends, warp-sized only
    assert (m >> 8) & 0xfff == 500
    assert m & (1 << 29)                                                         # size-class hints valid
    assert m & (1 << 28)                                                         # 128 nodes: outside the lean classes
    assert bin((m >> 20) & 0xff).count("1") == 2                                 # 400 and 500 nodes: two lean classes
    case = next(c for c in DBFP if c["id"] == "syn_doc_put_full500")
    pr = pricer(case)
    pr._build_log_grid()
    b = pr._solve_descriptor(True)
    m = b.family_mask()
    assert m & 2 and not m & 1 and m & 16 and not m & 8                          # linear, CTA-sized grid (2132 intervals)


def test_abi2_arrays_survive_take_and_concat():
    from finite_difference_b200.batch import SolveBatch
    a = pricer(next(c for c in DBFP if c["id"] == "syn_rebate_at_expiry"))
    a._build_log_grid()
    b = pricer(next(c for c in DBFP if c["id"] == "syn_double_out_solve"))
    b._build_log_grid()
    from finite_difference_b200 import hostlogic
    plain = hostlogic.build_barrier_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.03], [True], ["down-and-out"], [90.0],
                                          [None], [1.25], [[0.25, 0.5, 0.75]], [200], [200])
    cat = SolveBatch.concat([a._solve_descriptor(True), plain, b._solve_descriptor(True)])
    assert cat.n_solves == 3 and len(cat.mon_rebate) == len(cat.mon_step)
    lo, hi = int(cat.mon_begin[1]), int(cat.mon_begin[2])
    np.testing.assert_array_equal(cat.mon_rebate[lo:hi], 1.25)                   # the plain solve repeats its rebate
    assert cat.b_bdry[1] == cat.b[1] == 0.03 and cat.barrier_up[2] == 260.0
    sub = cat.take([2, 0])
    assert sub.barrier_up[0] == 260.0 and sub.b_bdry[1] == a.carry_rate_nacc
    assert len(sub.mon_rebate) == len(sub.mon_step) == (int(cat.mon_begin[3]) - hi) + lo
    np.testing.assert_array_equal(sub.mon_rebate[:int(sub.mon_begin[1])], cat.mon_rebate[hi:])


def test_batch_runner_composition_with_oracle_solves(monkeypatch):
    """price_and_greeks_batch / run_scenarios_frame with the GPU call replaced by the C oracle: the host composition
    (two solves per row, one-sided vega per vol point, theta, knock-in parity) reproduces the reference's
    price_log2 / greeks_log2 goldens."""
    import pandas as pd
    from oracle import oracle
    from finite_difference_b200 import batch as batch_mod, discrete_barrier_fdm_pricer as mod, run_config_scenarios as rcs
    cases = [c for c in DBFP if c["id"].startswith("config_scenarios_1_") and c["num_time_steps"] == 60][:6]
    captured = {}

    def fake_solve_host(b, dump=False, trace=False):
        n = b.n_solves
        res = batch_mod.SolveResults(price=np.empty(n), delta=np.empty(n), gamma=np.empty(n), stencil_v=np.zeros((n, 4)),
                                     stencil_s=np.zeros((n, 4)), idx=np.zeros((n, 4), np.int32), status=np.zeros(n, np.int32))
        for i in range(n):
            a = captured["args"][i]
            o = oracle.dbfp_solve(a["S_min"], a["S_max"], a["S0_eff"], a["S0_eff"], a["K"], a["T"], a["sigma"], a["r"], a["b"],
                                  a["q"], a["is_call"], a["ko_type"], a["lower_barrier"], a["upper_barrier"],
                                  a["rebate_amount"], a["rebate_at_hit"], a["monitor_times"], a["num_space_nodes"],
                                  a["num_time_steps"], a["rannacher_steps"])
            res.price[i], res.delta[i], res.gamma[i] = o.price, o.delta, o.gamma
        return res

    real_build = hostlogic.build_dbfp_batch

    def spy_build(**kw):
        n = len(kw["K"])
        shared = kw.get("shared_monitors", False)

        def at(k, v, i):
            if k == "monitor_times":
                return v if shared else v[i]
            if isinstance(v, (list, tuple, np.ndarray)):
                return v[i]
            return v
        captured["args"] = [{k: at(k, v, i) for k, v in kw.items() if k != "shared_monitors"} for i in range(n)]
        return real_build(**kw)

    monkeypatch.setattr(mod, "solve_host", fake_solve_host)
    monkeypatch.setattr(mod.hostlogic, "build_dbfp_batch", spy_build)
    frame = pd.DataFrame([dict(scenario_name=c["id"], S0=c["spot"], K=c["strike"], sigma=c["sigma"], rate=c["naca"],
                               barrier_type=c["barrier_type"], upper_barrier=c["upper_barrier"],
                               lower_barrier=c["lower_barrier"], FA_price=None, FA_delta=None, FA_gamma=None, FA_vega=None)
                          for c in cases])
    base = dict(valuation=_golden.d(cases[0]["valuation"]), maturity=_golden.d(cases[0]["maturity"]),
                monitor_dates=_golden.dbfp_monitor_dates(cases[0]), opt_type="put", num_time_steps=60)
    from finite_difference_b200 import barrier_book
    monkeypatch.setattr(barrier_book, "solve_host", fake_solve_host)
    for vectorised in (False, True):            # one pricer object per row / numpy over the rows
        df = rcs.run_scenarios_frame(frame, base, vectorised=vectorised)
        for i, c in enumerate(cases):
            ref = c["out"]["greeks_log2"]
            leg = max(abs(c["out"]["solve_price"]), 1e-6)
            assert abs(df.loc[i, "model_price"] - c["out"]["price_log2"]) <= 1e-12 * leg, (vectorised, i)
            assert abs(df.loc[i, "model_delta"] - ref["delta"]) <= 1e-10 * max(abs(ref["delta"]), 1.0)
            assert abs(df.loc[i, "model_vega"] - ref["vega"]) <= 1e-8 * max(abs(ref["vega"]), 1.0)
