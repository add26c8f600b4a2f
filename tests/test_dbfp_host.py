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
