"""the row-varying S-space solver (``DiscreteBarrierFDMPricer2``, discrete_barrier_fdm_pricer_2.py).

CPU tier: the C oracle (``fdo_pricer2_solve``) against goldens made by running the unmodified reference class --
bit for bit, every node, including two marches in the reference's unstable theta = 1/2 regime -- and the product's
host preparation (``rowvar.march_inputs``) against the same records.
GPU tier: the drop-in class and the batch API against the goldens of the regime in which the reference prices (every
step implicit): price <= 1e-10 relative, value vector node-wise, delta / gamma / vega."""
import json
import os
from datetime import date

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pricer2.json")) as fh:
    CASES = json.load(fh)["cases"]
GPU_CASES = [c for c in CASES if c["gpu"]]


def _args(c):
    a = dict(c["args"])
    a["valuation_date"] = date.fromisoformat(a["valuation_date"])
    a["maturity_date"] = date.fromisoformat(a["maturity_date"])
    a["monitoring_dates"] = [date.fromisoformat(d) for d in a["monitoring_dates"]]
    a["dividends"] = [(date.fromisoformat(d), x) for d, x in a["dividends"]]
    return a


def test_goldens_cover_the_features():
    ids = {c["id"] for c in CASES}
    assert {"doc_300x60", "vanilla_put", "double_out", "down_and_in", "dividend_escrow", "bgk_daily_40", "grid_600x600",
            "cn_8_steps_unstable"} <= ids
    assert any(c["out"]["use_bgk"] for c in CASES) and len(GPU_CASES) >= 18


@pytest.mark.parametrize("c", CASES, ids=[c["id"] for c in CASES])
def test_oracle_is_bit_exact_to_the_reference(c):
    from oracle import oracle
    price, V, prep = oracle.pricer2_price(**_args(c))
    o = c["out"]
    assert price == o["price"]
    assert np.array_equal(V, np.array(o["V"]))
    assert prep["S_eff"] == o["S_eff"] and prep["use_bgk"] == o["use_bgk"]
    assert (prep["s_nodes"][0], prep["s_nodes"][1], prep["s_nodes"][-1]) == (o["s_first"], o["s_second"], o["s_last"])


@pytest.mark.parametrize("c", CASES, ids=[c["id"] for c in CASES])
def test_host_preparation_matches_the_reference(c):
    from finite_difference_b200 import rowvar
    from oracle import oracle
    a = _args(c)
    m = rowvar.march_inputs(**a)
    o = c["out"]
    assert len(m.s_nodes) == o["n_nodes"] and m.s_eff == o["S_eff"] and m.tenor == o["tenor"] and m.dt == o["dt"]
    assert (m.s_nodes[0], m.s_nodes[1], m.s_nodes[-1]) == (o["s_first"], o["s_second"], o["s_last"])
    assert m.extras["use_bgk"] == o["use_bgk"] and m.extras["bgk_lower"] == o["bgk_lower"] and m.extras["bgk_upper"] == o["bgk_upper"]
    assert sorted(m.mon_keys) == o["monitor_steps"]
    prep = oracle.pricer2_prepare(**a)                       # the oracle's own restatement agrees node for node
    assert np.array_equal(m.s_nodes, prep["s_nodes"]) and np.array_equal(m.v_terminal, prep["v_terminal"])
    assert (m.bar_j, m.h_minus, m.h_plus) == (prep["bar_j"], prep["h_minus"], prep["h_plus"])


def test_escrow_beyond_the_first_interval_divides_by_zero_like_the_reference():
    """PV(dividends) > dS collapses the first two shifted nodes onto 0: dS = 0 and the reference raises
    ZeroDivisionError at :379.  The oracle reports the vanishing pivot the same way."""
    from oracle import oracle
    a = _args(CASES[0])
    a["dividends"] = [(date(2025, 12, 1), 2.5)]
    with pytest.raises(ZeroDivisionError):
        oracle.pricer2_price(**a)


# ---- GPU ---------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("c", GPU_CASES, ids=[c["id"] for c in GPU_CASES])
def test_class_vs_reference_golden(c):
    from finite_difference_b200.discrete_barrier_fdm_pricer_2 import DiscreteBarrierFDMPricer2
    pr = DiscreteBarrierFDMPricer2(**_args(c))
    o = c["out"]
    Sg, Vg, S_eff = pr._solve_grid_once()
    V_ref = np.array(o["V"])
    err = np.abs(np.array(Vg) - V_ref)
    assert np.all(err <= 1e-10 * np.abs(V_ref) + 1e-12 * np.max(np.abs(V_ref))), float(err.max())
    p = pr.price()
    assert abs(p - o["price"]) <= 1e-10 * max(abs(o["price"]), 1e-2)
    assert (pr.use_bgk_correction, pr.bgk_lower, pr.bgk_upper, pr.k_first_cont, pr.k_last_cont) == (
        o["use_bgk"], o["bgk_lower"], o["bgk_upper"], o["k_first"], o["k_last"])
    if "greeks" in o:
        g = pr.greeks()
        for k, tol in (("delta", 1e-9), ("gamma", 1e-7), ("vega", 1e-7)):
            assert abs(g[k] - o["greeks"][k]) <= tol * max(abs(o["greeks"][k]), 1e-3), (k, g[k], o["greeks"][k])


@pytest.mark.gpu
def test_batch_api_and_size_classes():
    """One batch with every size class (201 .. 1001 nodes) and both option types; against the oracle."""
    from finite_difference_b200 import rowvar
    from oracle import oracle
    base = _args(CASES[0])
    marches, refs = [], []
    for n_space, steps in ((50, 10), (224, 12), (300, 9), (416, 15), (600, 20), (800, 8), (1000, 11)):
        for ot, bt, lo, up in (("call", "down-and-out", 90.0, None), ("put", "up-and-out", None, 118.0), ("call", "none", None, None)):
            a = dict(base, num_space_nodes=n_space, num_time_steps=steps, rannacher_steps=steps, option_type=ot, barrier_type=bt,
                     lower_barrier=lo, upper_barrier=up)
            marches.append(rowvar.march_inputs(**a))
            refs.append(oracle.pricer2_price(**a))
    batch = rowvar.RowVarBatch(marches)
    res = rowvar.solve_rowvar_host(batch, dump=True)
    assert np.all(res.status == 0) and res.n_launches >= 5
    for i, (p, V, prep) in enumerate(refs):
        assert abs(res.price[i] - p) <= 1e-10 * max(abs(p), 1e-2), i
        n = len(V)
        assert np.all(np.abs(res.V[i, :n] - V) <= 1e-10 * np.abs(V) + 1e-12 * np.max(np.abs(V))), i


@pytest.mark.gpu
def test_singular_grid_raises_like_the_reference():
    from finite_difference_b200.discrete_barrier_fdm_pricer_2 import DiscreteBarrierFDMPricer2
    a = _args(CASES[0])
    a["dividends"] = [(date(2025, 12, 1), 2.5)]
    with pytest.raises(ZeroDivisionError):
        DiscreteBarrierFDMPricer2(**a).price()


@pytest.mark.gpu
def test_theta_half_steps_follow_the_reference_while_they_stay_small():
    """With theta = 1/2 steps the reference's march is unstable; over a handful of steps the kernel, which applies the
    reference's coefficients sign for sign, still tracks it (error amplified with the solution, hence the loose bar)."""
    from finite_difference_b200.discrete_barrier_fdm_pricer_2 import DiscreteBarrierFDMPricer2
    c = next(x for x in CASES if x["id"] == "cn_8_steps_unstable")
    p = DiscreteBarrierFDMPricer2(**_args(c)).price()
    assert abs(p - c["out"]["price"]) <= 1e-6 * abs(c["out"]["price"])
