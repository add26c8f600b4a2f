"""GPU parity where the hoisted factorisations are least comfortable (ADVICE, round 1): very large alpha dt (few
time steps on fine grids), so that the reference's pivots converge late or not at all inside the grid, the
row-0 correction reaches far, and the growth-factor matching switches between its regimes (imitation of the
reference's steady state / exact row sums / none).  Every kernel family: warp (outside the lean classes), lean, CTA.
Bar: 1e-10 on the price and node-wise on the value vector, status 0 everywhere: what fdm_cta_kernel flags as
FDM_STATUS_STIFF (pivots not converged inside its table) is solved again by fdm_exact_kernel, the reference's Thomas
sweep run literally."""
import numpy as np
import pytest

from finite_difference_b200 import _lib, hostlogic, solve_host
from finite_difference_b200.batch import SolveBatch
from oracle import oracle

pytestmark = pytest.mark.gpu

GRID = [(40, 4), (64, 5), (130, 4), (200, 6), (300, 5), (400, 4), (500, 8), (513, 5), (514, 6), (1000, 5), (2133, 8),
        (4096, 10), (8193, 12), (8193, 60)]


@pytest.mark.parametrize("n_space,n_time", GRID, ids=[f"{a}x{b}" for a, b in GRID])
def test_american_large_alpha_dt(n_space, n_time):
    cases = [dict(K=100.0, sigma=0.6, r=0.03, call=False, T=2.0), dict(K=110.0, sigma=0.45, r=0.08, call=True, T=1.5),
             dict(K=95.0, sigma=0.9, r=0.01, call=False, T=1.0)]
    parts = [hostlogic.build_american_batch([100.0], [c["K"]], [c["T"]], [c["sigma"]], [c["r"]], [c["r"]], [c["call"]], [n_space],
                                            [n_time], [2], [False]) for c in cases]
    res = solve_host(SolveBatch.concat(parts), dump=True)
    for i, c in enumerate(cases):
        o = oracle.american_solve(100.0, c["K"], c["T"], c["sigma"], c["r"], c["r"], c["call"], 0, n_space, n_time, 2)
        V = res.V[i, : n_space + 1]
        scale = np.max(np.abs(o.V))
        tol = 1e-10
        assert res.status[i] == 0, (i, res.status[i])
        assert np.all(np.abs(V - o.V) <= tol * np.abs(o.V) + 1e-2 * tol * scale), (i, float(np.max(np.abs(V - o.V))), res.status[i])
        assert abs(res.price[i] - o.price) <= tol * max(abs(o.price), 1e-2)
        if res.status[i] == 0:
            assert (int(res.idx[i, 2]), int(res.idx[i, 3])) == (o.ex_lo, o.ex_hi)


@pytest.mark.parametrize("n_space,n_time", GRID, ids=[f"{a}x{b}" for a, b in GRID])
def test_barrier_large_alpha_dt(n_space, n_time):
    mon = [0.25, 0.5, 0.75, 1.0, 1.25]
    cases = [dict(K=100.0, sigma=0.6, r=0.03, b=0.01, kind="down-and-out", lo=80.0, up=None, T=1.5),
             dict(K=105.0, sigma=0.4, r=0.06, b=0.06, kind="up-and-out", lo=None, up=140.0, T=1.5)]
    parts = [hostlogic.build_barrier_batch([100.0], [c["K"]], [c["T"]], [c["sigma"]], [c["r"]], [c["b"]], [True], [c["kind"]], [c["lo"]],
                                           [c["up"]], [1.0], [mon], [n_space], [n_time]) for c in cases]
    res = solve_host(SolveBatch.concat(parts), dump=True)
    for i, c in enumerate(cases):
        o = oracle.barrier_solve(100.0, c["K"], c["T"], c["sigma"], c["r"], c["b"], True, c["kind"], c["lo"], c["up"], 1.0, mon,
                                 n_space, n_time, True)
        V = res.V[i, : n_space + 1]
        scale = np.max(np.abs(o.V))
        tol = 1e-10
        assert res.status[i] == 0, (i, res.status[i])
        assert np.all(np.abs(V - o.V) <= tol * np.abs(o.V) + 1e-2 * tol * scale), (i, float(np.max(np.abs(V - o.V))), res.status[i])
        assert abs(res.price[i] - o.price) <= tol * max(abs(o.price), 1e-2)


@pytest.mark.parametrize("n_space,n_time", [(1000, 6), (2133, 9), (8193, 14)], ids=["1000x6", "2133x9", "8193x14"])
def test_flagged_solves_with_dividends_are_rescued(n_space, n_time):
    """The exact kernel also carries the ex-dividend spline jump (serial natural cubic spline, the reference's order)."""
    cases = [dict(K=100.0, sigma=0.6, r=0.03, call=True, divs=[(0.6, 2.0), (1.3, 1.5)]),
             dict(K=105.0, sigma=0.8, r=0.05, call=False, divs=[(0.9, 3.0)])]
    parts = [hostlogic.build_american_batch([100.0], [c["K"]], [2.0], [c["sigma"]], [c["r"]], [c["r"]], [c["call"]], [n_space],
                                            [n_time], [2], [False], div_taus=[c["divs"]]) for c in cases]
    res = solve_host(SolveBatch.concat(parts), dump=True, trace=True)
    for i, c in enumerate(cases):
        o = oracle.american_solve(100.0, c["K"], 2.0, c["sigma"], c["r"], c["r"], c["call"], 0, n_space, n_time, 2, 4.5, c["divs"],
                                  trace=True)
        V = res.V[i, : n_space + 1]
        assert res.status[i] == 0
        assert np.all(np.abs(V - o.V) <= 1e-10 * np.abs(o.V) + 1e-12 * np.max(np.abs(o.V))), (i, float(np.max(np.abs(V - o.V))))
        assert abs(res.price[i] - o.price) <= 1e-10 * max(abs(o.price), 1e-2)
        steps = o.ex_trace.shape[0]
        assert np.array_equal(res.ex_trace[i, :steps], o.ex_trace), i


def test_linear_solve_with_cash_dividends_on_a_large_grid():
    """A European march through cash dividends on a CTA-kernel grid: no pricer of the reference builds one (its
    European / barrier pricers escrow dividends), so there is no oracle for it; the linear CTA kernel keeps no spline
    arrays and hands such a solve to fdm_exact_kernel.  It must come back solved (status 0), not skipped, and agree
    with the same contract on a warp-kernel grid to discretisation accuracy."""
    divs = [[(0.3, 2.0), (0.8, 1.5)]]
    res = {}
    for n in (500, 1000, 2133):
        b = hostlogic.build_american_batch([100.0], [100.0], [1.0], [0.3], [0.05], [0.05], [False], [n], [400], [2], [False],
                                           div_taus=divs)
        b.flags = (b.flags & ~_lib.FLAG_AMERICAN).astype(np.int32)
        r = solve_host(b)
        assert r.status[0] == 0, (n, r.status[0])
        assert np.isfinite(r.price[0]) and r.price[0] > 0.0
        res[n] = float(r.price[0])
    assert abs(res[1000] - res[500]) <= 2e-3 * res[500], res
    assert abs(res[2133] - res[1000]) <= 1e-3 * res[1000], res


def test_extreme_alpha_dt_goes_to_the_literal_sweep():
    """alpha dt = 6e4 (7346 nodes, 21 steps, sigma 0.5): found by the random sweep at 1.2e-10 in the hoisted CTA march;
    beyond theta alpha dt = 1.5e4 the solve is handed to fdm_exact_kernel and meets the bar with room."""
    kw = dict(K=104.0, T=1.0, sig=0.515, r=0.03, q=0.01)
    for call in (False, True):
        b = hostlogic.build_american_batch([100.0], [kw["K"]], [kw["T"]], [kw["sig"]], [kw["r"]], [kw["r"] - kw["q"]], [call], [7346], [21], [2], [False])
        res = solve_host(b, dump=True)
        o = oracle.american_solve(100.0, kw["K"], kw["T"], kw["sig"], kw["r"], kw["r"] - kw["q"], call, 0, 7346, 21, 2)
        assert res.status[0] == 0
        scale = np.max(np.abs(o.V))
        assert np.all(np.abs(res.V[0, :7347] - o.V) <= 2e-11 * np.abs(o.V) + 1e-12 * scale)
        assert abs(res.price[0] - o.price) <= 2e-11 * abs(o.price)
