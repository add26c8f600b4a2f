"""GPU parity: the CUDA path (through the C ABI, host buffers) against the reference goldens and
the CPU oracle.  Floating-point bar (BASELINE.json north_star): price relative error <= 1e-10,
early-exercise interval identical.  Node-wise: |dV| <= 1e-10*|V_ref| + 1e-12*max|V_ref|."""
import numpy as np
import pytest

import _golden
from finite_difference_b200 import hostlogic, solve_host
from finite_difference_b200.batch import SolveBatch
from oracle import oracle

pytestmark = pytest.mark.gpu

REL = 1e-10
AM = _golden.american_cases()
BAR = _golden.barrier_cases()


def _v_close(V, V_ref):
    scale = np.max(np.abs(V_ref))
    err = np.abs(V - V_ref)
    tol = REL * np.abs(V_ref) + 1e-12 * scale
    bad = np.nonzero(err > tol)[0]
    assert bad.size == 0, f"{bad.size} nodes off, worst {err.max():.3e} at {int(np.argmax(err))} (scale {scale:.3e})"


def _am_batch(cases):
    parts = []
    for c in cases:
        kw = _golden.am_oracle_kwargs(c)
        parts.append(hostlogic.build_american_batch(
            [kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]], [kw["is_call"]],
            [kw["n_space"]], [kw["n_time"]], [kw["rannacher"]], [kw["model"] == 1], kw["s_max_mult"],
            div_taus=[kw["divs_tau"]]))
    return SolveBatch.concat(parts)


@pytest.fixture(scope="module")
def am_results():
    batch = _am_batch(AM)
    return batch, solve_host(batch, dump=True, trace=True)


@pytest.mark.parametrize("i", range(len(AM)), ids=[c["id"] for c in AM])
def test_american_vs_reference_golden(am_results, i):
    batch, res = am_results
    case, o = AM[i], AM[i]["out"]
    n = max(case["num_space_nodes"], 3)
    assert res.status[i] == 0
    assert int(res.idx[i, 0]) == o["spot_idx"] and int(res.idx[i, 1]) == o["strike_idx"]
    V_ref = np.array(o["V"])
    _v_close(res.V[i, : n + 1], V_ref)
    assert abs(res.price[i] - o["price_log"]) <= REL * abs(o["price_log"]) + 1e-300
    assert (int(res.idx[i, 2]), int(res.idx[i, 3])) == (o["ex_lo"], o["ex_hi"])
    # in-kernel cubic delta / gamma against the reference's LAPACK solve
    assert abs(res.delta[i] - o["delta"]) <= 1e-8 * max(abs(o["delta"]), 1e-6)
    assert abs(res.gamma[i] - o["gamma"]) <= 1e-7 * max(abs(o["gamma"]), 1e-6)


@pytest.mark.parametrize("i", range(len(AM)), ids=[c["id"] for c in AM])
def test_american_exercise_trace_vs_oracle(am_results, i):
    """Early-exercise interval identical at EVERY step (the oracle is bit-exact to the reference)."""
    batch, res = am_results
    orc = oracle.american_solve(**_golden.am_oracle_kwargs(AM[i]), trace=True)
    steps = orc.ex_trace.shape[0]
    assert np.array_equal(res.ex_trace[i, :steps], orc.ex_trace)


@pytest.fixture(scope="module")
def bar_results():
    parts = []
    for c in BAR:
        bt = c["barrier_type"]
        ko = {"down-and-in": "down-and-out", "up-and-in": "up-and-out"}.get(bt, bt)
        o = c["out"]
        parts.append(hostlogic.build_barrier_batch(
            [c["S0"]], [c["K"]], [c["T"]], [c["sigma"]], [c["r_disc"]], [c["b_carry"]],
            [c["option_type"] == "call"], [ko], [c["lower_barrier"]], [c["upper_barrier"]], [c["rebate"]],
            [c["monitor_times"]], [o["N_space"]], [o["N_time"]]))
    batch = SolveBatch.concat(parts)
    return batch, solve_host(batch, dump=True)


@pytest.mark.parametrize("i", range(len(BAR)), ids=[c["id"] for c in BAR])
def test_barrier_vs_reference_golden(bar_results, i):
    batch, res = bar_results
    o = BAR[i]["out"]
    n = o["N_space"]
    assert res.status[i] == 0
    mb = batch.mon_begin
    assert list(batch.mon_step[mb[i]:mb[i + 1]]) == (o["monitor_steps"] if batch.flags[i] & 0x30 else
                                                     list(batch.mon_step[mb[i]:mb[i + 1]]))
    _v_close(res.V[i, : n + 1], np.array(o["V"]))
    assert abs(res.price[i] - o["grid_price"]) <= REL * abs(o["grid_price"])
    assert abs(res.delta[i] - o["grid_delta"]) <= 1e-9 * max(abs(o["grid_delta"]), 1e-6)
    assert abs(res.gamma[i] - o["grid_gamma"]) <= 1e-8 * max(abs(o["grid_gamma"]), 1e-6)


def test_device_resident_path_matches_host_path():
    """fdm_solve_batch (device pointers, torch stream) == fdm_solve_batch_host."""
    from finite_difference_b200 import DeviceBatch
    batch = _am_batch(AM[:8])
    host = solve_host(batch)
    dev = DeviceBatch(batch)
    dev.solve()
    r = dev.results()
    assert np.array_equal(r.price, host.price)
    assert np.array_equal(r.idx, host.idx)
    assert r.n_launches >= 1


def test_seeded_batch_vs_oracle_cfg3_shape():
    """BASELINE config 3 shape (Black-76, T=31/365, 500x500) on 64 seeded options vs the oracle."""
    rng = np.random.default_rng(20251018 + 3)
    n = 64
    K = rng.uniform(180, 280, n)
    sig = rng.uniform(0.15, 0.45, n)
    r = np.log1p(rng.uniform(0.03, 0.10, n))
    call = rng.uniform(size=n) < 0.5
    T = 31.0 / 365.0
    batch = hostlogic.build_american_batch(np.full(n, 229.74), K, T, sig, r, r, call, 500, 500, 2, True)
    res = solve_host(batch)
    worst = 0.0
    for i in range(n):
        o = oracle.american_solve(229.74, K[i], T, sig[i], r[i], r[i], bool(call[i]), 1, 500, 500, 2)
        assert (int(res.idx[i, 2]), int(res.idx[i, 3])) == (o.ex_lo, o.ex_hi)
        worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
    assert worst <= REL, worst


def test_full_size_properties_cfg2():
    """10k barrier solves (BASELINE config 2 size): size-independent properties.
    (a) KO price <= vanilla price from the same engine, (b) monotone in the rebate (same grid),
    (c) linearity: the PDE is linear, so scaling S0,K,B by c scales the price by c."""
    rng = np.random.default_rng(20251018 + 2)
    n = 10000
    K = rng.uniform(80, 120, n); sig = rng.uniform(0.10, 0.50, n); r = rng.uniform(0.01, 0.08, n)
    q = rng.uniform(0.0, 0.03, n); T = rng.choice([0.5, 1.0, 2.0], n); B = 100.0 * rng.uniform(0.70, 0.95, n)
    mons = [[k / 12.0 for k in range(1, int(round(12 * t)) + 1)] for t in T]
    S0 = np.full(n, 100.0)

    def run(scale, ko, rebate):
        b = hostlogic.build_barrier_batch(S0 * scale, K * scale, T, sig, r, r - q, np.ones(n, bool), [ko] * n,
                                          list(B * scale), [None] * n, np.full(n, rebate * scale), mons, 400, 400)
        return solve_host(b).price

    p_ko = run(1.0, "down-and-out", 0.0)
    p_van = run(1.0, "none", 0.0)
    p_ko_reb = run(1.0, "down-and-out", 1.0)
    p_scaled = run(2.0, "down-and-out", 0.0)
    assert np.all(np.isfinite(p_ko))
    assert np.all(p_ko <= p_van + 1e-9)
    assert np.all(p_ko_reb >= p_ko - 1e-9)
    assert np.max(np.abs(p_scaled - 2.0 * p_ko) / np.maximum(np.abs(p_ko), 1e-3)) < 1e-9


def test_fused_book_path_matches_host_built_batch():
    """fdm_american_book_host (device-side fan-out + epilogue) == AmericanBook.run (host-built descriptors)."""
    from finite_difference_b200.book import AmericanBook
    rng = np.random.default_rng(11)
    n = 96
    K = rng.uniform(180, 280, n); sig = rng.uniform(0.15, 0.45, n); r = np.log1p(rng.uniform(0.03, 0.10, n))
    call = rng.uniform(size=n) < 0.5; b76 = rng.uniform(size=n) < 0.5
    book = AmericanBook(np.full(n, 229.74), K, rng.uniform(0.05, 1.5, n), sig, r, r * rng.uniform(0.5, 1.0, n), call,
                        rng.choice([100, 257, 400, 500], n), rng.choice([50, 120, 200], n), rng.integers(0, 4, n), black76=b76)
    for what, rich in (("scenario", True), ("scenario", False), ("greeks_log2", True), ("price_log2", True), ("price_log", True)):
        a = book.run(what, use_richardson=rich)
        f = book.run_fused(what, use_richardson=rich)
        assert f["node_steps"] == a["node_steps"] and f["n_solves"] == a["n_solves"]
        assert np.all(f["status"] == 0)
        for k in ("price_log2", "price", "delta", "gamma", "vega", "theta"):
            if k in a and k in f:
                tol = {"price_log2": 1e-12, "price": 1e-12, "delta": 1e-9, "gamma": 1e-8, "vega": 1e-9, "theta": 1e-8}[k]
                err = np.max(np.abs(f[k] - a[k]) / np.maximum(np.abs(a[k]), 1e-6))
                assert err <= tol, (what, rich, k, err)
        assert np.array_equal(f["ex_lo"], a["ex_lo"]) and np.array_equal(f["ex_hi"], a["ex_hi"])
