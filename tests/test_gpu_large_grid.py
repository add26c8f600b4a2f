"""GPU parity of the CTA-per-solve kernel (514..8193 space nodes) against the CPU oracle, which is
pinned bit-exact to the reference (tests/test_oracle_golden.py).  Same bars as the small grids:
node values 1e-10 relative, price 1e-10 relative, exercise interval identical at every step."""
import numpy as np
import pytest

from finite_difference_b200 import hostlogic, solve_host
from finite_difference_b200.batch import SolveBatch
from oracle import oracle

pytestmark = pytest.mark.gpu
REL = 1e-10


def _v_close(V, V_ref):
    scale = np.max(np.abs(V_ref))
    err = np.abs(V - V_ref)
    tol = REL * np.abs(V_ref) + 1e-12 * scale
    bad = np.nonzero(err > tol)[0]
    assert bad.size == 0, f"{bad.size} nodes off, worst {err.max():.3e} at {int(np.argmax(err))} (scale {scale:.3e})"


AM_CASES = [
    # id, spot, strike, T, sigma, r, b, call, model, n, m, rannacher, divs
    ("put_1000", 100.0, 100.0, 1.0, 0.2, 0.05, 0.05, False, 0, 1000, 150, 2, []),
    ("call_b76_1500", 229.74, 240.0, 0.5, 0.3, 0.06, 0.06, True, 1, 1500, 120, 2, []),
    ("put_2049_edge", 50.0, 55.0, 2.0, 0.35, 0.03, 0.01, False, 0, 2049, 60, 3, []),
    ("call_2050_edge", 50.0, 45.0, 0.75, 0.25, 0.04, 0.04, True, 0, 2050, 60, 0, []),
    ("put_div_4096", 229.74, 220.0, 1.0, 0.26, 0.07, 0.07, False, 0, 4096, 90, 2, [(0.3, 8.0), (0.88, 7.6)]),
    ("call_div_8192", 229.74, 250.0, 1.0, 0.15, 0.05, 0.05, True, 0, 8192, 100, 2, [(0.3, 8.0), (0.88, 7.6)]),
    ("put_div_8192_hivol", 229.74, 200.0, 1.0, 0.45, 0.09, 0.09, False, 0, 8192, 100, 2, [(0.3, 3.0), (0.88, 9.0)]),
    ("put_8193_max", 100.0, 120.0, 1.0, 0.3, 0.05, 0.05, False, 0, 8193, 40, 2, []),
    ("put_514_min", 100.0, 90.0, 1.0, 0.3, 0.05, 0.05, False, 0, 514, 80, 2, []),
]


def _kw(c):
    _, spot, strike, T, sigma, r, b, call, model, n, m, ran, divs = c
    return dict(spot=spot, strike=strike, T=T, sigma=sigma, r=r, b=b, is_call=call, model=model, n_space=n,
                n_time=m, rannacher=ran, s_max_mult=4.5, divs_tau=divs)


def _batch(c):
    kw = _kw(c)
    return hostlogic.build_american_batch([kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]],
                                          [kw["is_call"]], [kw["n_space"]], [kw["n_time"]], [kw["rannacher"]],
                                          [kw["model"] == 1], 4.5, div_taus=[kw["divs_tau"]])


@pytest.mark.parametrize("c", AM_CASES, ids=[c[0] for c in AM_CASES])
def test_american_large_grid_vs_oracle(c):
    kw = _kw(c)
    res = solve_host(_batch(c), dump=True, trace=True)
    o = oracle.american_solve(**kw, trace=True)
    n = kw["n_space"]
    assert res.status[0] == 0
    assert (int(res.idx[0, 0]), int(res.idx[0, 1])) == (o.spot_idx, o.strike_idx)
    _v_close(res.V[0, : n + 1], o.V)
    assert abs(res.price[0] - o.price) <= REL * abs(o.price)
    assert (int(res.idx[0, 2]), int(res.idx[0, 3])) == (o.ex_lo, o.ex_hi)
    steps = o.ex_trace.shape[0]
    assert np.array_equal(res.ex_trace[0, :steps], o.ex_trace)
    d, g = oracle.local_cubic_delta_gamma(o.s_nodes, o.V, o.s_nodes[o.spot_idx])
    assert abs(res.delta[0] - d) <= 1e-6 * max(abs(d), 1e-6)


BAR_CASES = [
    ("doc_1200", 100.0, 100.0, 1.0, 0.2, 0.05, 0.03, True, "down-and-out", 90.0, None, 0.0, 1200, 300),
    ("uop_3000_rebate", 100.0, 105.0, 0.5, 0.3, 0.04, 0.04, False, "up-and-out", None, 120.0, 2.0, 3000, 200),
    ("vanilla_5000", 100.0, 95.0, 2.0, 0.25, 0.02, 0.02, True, "none", None, None, 0.0, 5000, 100),
]


@pytest.mark.parametrize("c", BAR_CASES, ids=[c[0] for c in BAR_CASES])
def test_barrier_large_grid_vs_oracle(c):
    _, S0, K, T, sig, r, b, call, bt, lo, up, reb, n, m = c
    mon = [k / 12.0 for k in range(1, int(round(12 * T)) + 1)]
    batch = hostlogic.build_barrier_batch([S0], [K], [T], [sig], [r], [b], [call], [bt], [lo], [up], [reb], [mon], [n], [m])
    res = solve_host(batch, dump=True)
    o = oracle.barrier_solve(S0, K, T, sig, r, b, call, bt, lo, up, reb, mon, n, m, bt != "none")
    assert res.status[0] == 0
    _v_close(res.V[0, : n + 1], o.V)
    assert abs(res.price[0] - o.price) <= REL * abs(o.price)
    assert abs(res.delta[0] - o.delta) <= 1e-8 * max(abs(o.delta), 1e-6)
    assert abs(res.gamma[0] - o.gamma) <= 1e-7 * max(abs(o.gamma), 1e-6)


def test_mixed_batch_small_and_large_grids():
    """One batch holding warp-kernel and CTA-kernel solves, American and linear."""
    parts = [_batch(c) for c in (AM_CASES[0], AM_CASES[4])]
    parts.append(hostlogic.build_american_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.05], [False], [400], [400]))
    mon = [k / 12.0 for k in range(1, 13)]
    parts.append(hostlogic.build_barrier_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.05], [True], ["down-and-out"],
                                               [90.0], [None], [0.0], [mon], [400], [400]))
    parts.append(hostlogic.build_barrier_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.05], [True], ["down-and-out"],
                                               [90.0], [None], [0.0], [mon], [1200], [300]))
    res = solve_host(SolveBatch.concat(parts))
    ref = [oracle.american_solve(**_kw(AM_CASES[0])).price, oracle.american_solve(**_kw(AM_CASES[4])).price,
           oracle.american_solve(100.0, 100.0, 1.0, 0.2, 0.05, 0.05, False, 0, 400, 400, 2).price,
           oracle.barrier_solve(100.0, 100.0, 1.0, 0.2, 0.05, 0.05, True, "down-and-out", 90.0, None, 0.0, mon, 400, 400, True).price,
           oracle.barrier_solve(100.0, 100.0, 1.0, 0.2, 0.05, 0.05, True, "down-and-out", 90.0, None, 0.0, mon, 1200, 300, True).price]
    assert np.all(res.status == 0)
    assert np.max(np.abs(res.price - np.array(ref)) / np.abs(ref)) <= REL


def test_cfg4_full_size_one_solve():
    """BASELINE config 4 at full size (8192 nodes x 10 000 steps, two dividends) against the oracle."""
    kw = dict(spot=229.74, strike=235.0, T=1.0, sigma=0.3, r=0.06, b=0.06, is_call=True, model=0, n_space=8192,
              n_time=10000, rannacher=2, s_max_mult=4.5, divs_tau=[(109.0 / 365.0, 5.5), (319.0 / 365.0, 4.25)])
    batch = hostlogic.build_american_batch([kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]],
                                           [True], [8192], [10000], [2], [False], 4.5, div_taus=[kw["divs_tau"]])
    res = solve_host(batch, dump=True)
    o = oracle.american_solve(**kw)
    assert res.status[0] == 0
    err = abs(res.price[0] - o.price) / abs(o.price)
    assert err <= REL, err
    _v_close(res.V[0, :8193], o.V)
    assert (int(res.idx[0, 2]), int(res.idx[0, 3])) == (o.ex_lo, o.ex_hi)


# Row-0 correction mode of the CTA kernel (fdm_cta.cuh, setup_phase): small time steps so that the correction's NM^i
# tail dies inside the grid, every size class (4x12, 4x16, 5x14, 5x16, 8x12, 8x16, 16x12, 16x16 warps x slots: up to
# 1536 / 2048 / 2240 / 2560 / 3072 / 4096 / 6144 / 8192 interior rows), a full class (no padding) and the first size of
# the next one, row 0 at different slots of its owner thread (pad % C).
ROW0_SIZES = [600, 777, 1234, 1537, 1538, 2047, 2133, 2241, 2242, 2500, 2900, 3073, 3333, 4097, 6000, 6145, 6146, 8000, 8193]


@pytest.mark.parametrize("n", ROW0_SIZES)
def test_american_row0_mode_vs_oracle(n):
    kw = dict(spot=100.0, strike=104.0, T=0.05, sigma=0.32, r=0.05, b=0.03, is_call=(n % 2 == 0), model=0, n_space=n,
              n_time=160, rannacher=2, s_max_mult=4.5, divs_tau=[])
    batch = hostlogic.build_american_batch([kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]],
                                           [kw["is_call"]], [n], [160], [2], [False], 4.5, div_taus=[[]])
    res = solve_host(batch, dump=True, trace=True)
    o = oracle.american_solve(**kw, trace=True)
    assert res.status[0] == 0
    _v_close(res.V[0, : n + 1], o.V)
    assert abs(res.price[0] - o.price) <= REL * abs(o.price)
    assert np.array_equal(res.ex_trace[0, : o.ex_trace.shape[0]], o.ex_trace)


@pytest.mark.parametrize("n", ROW0_SIZES)
def test_barrier_row0_mode_vs_oracle(n):
    S0, K, T, sig, r, b = 100.0, 97.0, 0.25, 0.28, 0.04, 0.02
    mon = [0.05, 0.1, 0.15, 0.2]
    bt, lo, up = ("down-and-out", 88.0, None) if n % 2 else ("up-and-out", None, 125.0)
    batch = hostlogic.build_barrier_batch([S0], [K], [T], [sig], [r], [b], [True], [bt], [lo], [up], [1.5], [mon], [n], [400])
    res = solve_host(batch, dump=True)
    o = oracle.barrier_solve(S0, K, T, sig, r, b, True, bt, lo, up, 1.5, mon, n, 400, True)
    assert res.status[0] == 0
    _v_close(res.V[0, : n + 1], o.V)
    assert abs(res.price[0] - o.price) <= REL * abs(o.price)


# The CTA kernel compiles its time loop per warp role and per form of the warp-to-warp sums (fdm_cta.cuh): which loops
# a solve runs depends on the stiffness theta*dt*alpha of its time step.  Volatility bands at the time steps of the
# benchmark shapes (8192 nodes, dt = 1e-4; 2048 nodes, dt = 2.5e-3) walk through one-term sums without and with the
# third barrier, and multi-term sums; calls and puts put the moving boundary in the last / the first warp.
STIFF_CASES = [(n, m, T, sig, call) for (n, m, T) in ((8192, 200, 0.02), (2048, 100, 0.25), (4096, 120, 0.03), (2200, 100, 0.25),
                                                      (5000, 120, 0.03))
               for sig in (0.16, 0.28, 0.35, 0.45) for call in (False, True)]


@pytest.mark.parametrize("n,m,T,sig,call", STIFF_CASES)
def test_american_stiffness_bands_vs_oracle(n, m, T, sig, call):
    kw = dict(spot=229.74, strike=235.0 if call else 225.0, T=T, sigma=sig, r=0.06, b=0.06, is_call=call, model=0,
              n_space=n, n_time=m, rannacher=2, s_max_mult=4.5, divs_tau=[])
    batch = hostlogic.build_american_batch([kw["spot"]], [kw["strike"]], [T], [sig], [kw["r"]], [kw["b"]], [call], [n], [m],
                                           [2], [False], 4.5, div_taus=[[]])
    res = solve_host(batch, dump=True, trace=True)
    o = oracle.american_solve(**kw, trace=True)
    assert res.status[0] == 0
    _v_close(res.V[0, : n + 1], o.V)
    assert abs(res.price[0] - o.price) <= REL * abs(o.price)
    assert np.array_equal(res.ex_trace[0, : o.ex_trace.shape[0]], o.ex_trace)


@pytest.mark.parametrize("sig", (0.12, 0.3, 0.5))
def test_barrier_stiffness_bands_vs_oracle(sig):
    # 16-warp linear class (table warps + scalar warps) and the 5-warp class of the barrier scenario runner
    for n, m, T in ((8192, 200, 0.02), (2133, 125, 0.25), (2400, 125, 0.25), (1400, 125, 0.25)):
        S0, K, r, b = 100.0, 97.0, 0.04, 0.02
        mon = [T * 0.25, T * 0.5, T * 0.75]
        batch = hostlogic.build_barrier_batch([S0], [K], [T], [sig], [r], [b], [True], ["down-and-out"], [88.0], [None],
                                              [0.0], [mon], [n], [m])
        res = solve_host(batch, dump=True)
        o = oracle.barrier_solve(S0, K, T, sig, r, b, True, "down-and-out", 88.0, None, 0.0, mon, n, m, True)
        assert res.status[0] == 0
        _v_close(res.V[0, : n + 1], o.V)
        assert abs(res.price[0] - o.price) <= REL * abs(o.price)
