"""Host logic of the batch API (no GPU): fan-out layout and the Richardson / vega / theta epilogues,
fed with oracle solves, must reproduce the reference's price_log2 / greeks_log2 goldens."""
import numpy as np
import pytest

import _golden
from finite_difference_b200.book import AmericanBook, BarrierBook
from oracle import oracle

AM_FULL = [c for c in _golden.american_cases() if c["full_api"]]


def _book(case):
    kw = _golden.am_oracle_kwargs(case)
    return AmericanBook([kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]], [kw["is_call"]],
                        case["num_space_nodes"], case["num_time_steps"], case["rannacher_steps"],
                        black76=kw["model"] == 1, s_max_mult=kw["s_max_mult"], div_taus=[kw["divs_tau"]]), kw


def _solve_with_oracle(case, kw, batch):
    price, delta, gamma = (np.empty(batch.n_solves) for _ in range(3))
    steps = batch.steps_per_solve()
    for i in range(batch.n_solves):
        # segment step counts sum to the requested n_time except in degenerate dividend layouts
        o = oracle.american_solve(**dict(kw, sigma=float(batch.sigma[i]), n_time=int(steps[i])))
        price[i] = o.price
        delta[i], gamma[i] = oracle.local_cubic_delta_gamma(o.s_nodes, o.V, o.s_nodes[o.spot_idx])
    return price, delta, gamma


def test_fanout_layout_counts():
    book, _ = _book(AM_FULL[0])
    assert book.fanout("price_log")[0].n_solves == 1
    assert book.fanout("price_log2")[0].n_solves == 2
    assert book.fanout("greeks_log2")[0].n_solves == 6
    assert book.fanout("greeks_log2", use_richardson=False)[0].n_solves == 3
    b, lay = book.fanout("scenario")
    assert b.n_solves == 8 and list(lay.roles) == ["p_N", "p_2N", "g_N", "g_2N", "up_h", "dn_h", "up_2h", "dn_2h"]
    assert book.fanout("scenario", dedupe=True)[0].n_solves in (6, 7)


@pytest.mark.parametrize("case", AM_FULL, ids=[c["id"] for c in AM_FULL])
def test_epilogues_reproduce_reference_goldens(case):
    book, kw = _book(case)
    if kw["divs_tau"]:
        pytest.skip("dividend step tables: n_time is not recoverable from the segment sum")
    o = case["out"]
    for what, rich, h, key in (("greeks_log2", True, 0.01, "greeks_log2"), ("greeks_log2", False, 0.01, "greeks_log2_norich"),
                               ("greeks_log2", True, 0.005, "greeks_log2_h005")):
        batch, lay = book.fanout(what, h, rich)
        p, d, g = _solve_with_oracle(case, kw, batch)
        out = book._combine(what, lay, p, d, g, h, rich)
        for k, ref in o[key].items():
            assert out[k][0] == pytest.approx(ref, rel=1e-13, abs=1e-15), (key, k)
    batch, lay = book.fanout("price_log2")
    p, d, g = _solve_with_oracle(case, kw, batch)
    assert book._combine("price_log2", lay, p, d, g, 0.01, True)["price_log2"][0] == o["price_log2"]
    # scenario fan-out with and without de-duplication gives the same numbers
    for dd in (False, True):
        batch, lay = book.fanout("scenario", dedupe=dd)
        p, d, g = _solve_with_oracle(case, kw, batch)
        out = book._combine("scenario", lay, p, d, g, 0.01, True)
        assert out["price_log2"][0] == o["price_log2"]
        assert out["vega"][0] == pytest.approx(o["greeks_log2"]["vega"], rel=1e-13)


def test_barrier_book_grid_configuration_matches_oracle():
    for c in _golden.barrier_cases():
        book = BarrierBook([c["S0"]], [c["K"]], [c["T"]], [c["sigma"]], [c["r_disc"]], [c["b_carry"]],
                           c["option_type"], c["barrier_type"], [c["lower_barrier"]], [c["upper_barrier"]],
                           [c["rebate"]], [c["monitor_times"]], c["N_space"], c["N_time"])
        assert (int(book.N_space[0]), int(book.N_time[0])) == (c["out"]["N_space"], c["out"]["N_time"])
        b, lay = book.fanout("greeks")
        assert b.n_solves == 3


def test_barrier_book_rejects_unknown_type():
    with pytest.raises(ValueError):
        BarrierBook([100.0], [100.0], [1.0], [0.2], [0.05], [0.05], "call", "double-out")
