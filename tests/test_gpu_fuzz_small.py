"""Seeded random sweep of the one-warp kernels (fdm_lean_kernel, fdm_warp_kernel: 3 .. 513 space nodes) against the
oracle: every grid size class and the sizes between them, 0-3 cash dividends (spline jumps inside the lean kernel,
segment restarts), Black-76 forwards, Rannacher 0-4, few and many time steps; barrier grids with and without rebate.
Bar: every node 1e-10 relative (+1e-12 of the vector's scale), price 1e-10, exercise interval identical at every step
-- without dividends.  After a spline jump the far out-of-the-money tail of a put holds values that are cancellation
noise around zero (|V| < 1e-16 of the scale, either sign), and "exercise iff payoff > continuation" (0 > V there) then
follows the last bits of the spline: the kernels eliminate the spline's system in parallel, equal to the reference's
serial sweep to a few ulps but not bit for bit, so the recorded interval's ends can differ on such nodes for a step or
two after a jump while every value agrees to 1e-14 (fdm_exact_kernel is the literal sweep).  With dividends the sweep
therefore compares values at every node and the interval over the nodes that carry a value."""
import numpy as np
import pytest

from finite_difference_b200 import hostlogic, solve_host
from finite_difference_b200.batch import SolveBatch
from oracle import oracle

pytestmark = pytest.mark.gpu
REL = 1e-10


def _v_close(V, ref, what):
    tol = REL * np.abs(ref) + 1e-12 * np.max(np.abs(ref))
    bad = np.abs(V - ref) > tol
    assert not bad.any(), (what, int(bad.sum()), float(np.max(np.abs(V - ref) / (np.abs(ref) + 1e-12 * np.max(np.abs(ref))))))


def american_cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(rng.integers(3, 514))
        if rng.uniform() < 0.4:
            n = int(np.clip(rng.choice([129, 257, 289, 321, 385, 417, 481, 513]) + rng.integers(-3, 3), 3, 513))
        T = float(rng.choice([0.05, 0.25, 1.0, 2.0]))
        b76 = bool(rng.uniform() < 0.3)
        nd = 0 if b76 else int(rng.choice([0, 0, 1, 2, 3]))
        taus = sorted(float(x) for x in rng.uniform(0.08, 0.92, nd) * T)
        divs = [(t, float(rng.uniform(0.5, 6.0))) for t in taus]
        out.append(dict(n=n, m=int(rng.integers(max(8, 4 * (nd + 1)), 420)), T=T, sig=float(rng.uniform(0.08, 0.7)),
                        r=float(rng.uniform(0.0, 0.1)), q=float(rng.uniform(0.0, 0.05)), K=float(rng.uniform(80, 125)),
                        call=bool(rng.uniform() < 0.5), rann=int(rng.choice([0, 2, 4])), b76=b76, divs=divs))
    return out


@pytest.mark.parametrize("seed", [31, 32, 33, 34])
def test_american_random_small_grids(seed):
    cases = american_cases(seed, 40)
    parts = [hostlogic.build_american_batch([100.0], [c["K"]], [c["T"]], [c["sig"]], [c["r"]], [c["r"] - c["q"]], [c["call"]],
                                            [c["n"]], [c["m"]], [c["rann"]], [c["b76"]], div_taus=[c["divs"]]) for c in cases]
    res = solve_host(SolveBatch.concat(parts), dump=True, trace=True)
    for i, c in enumerate(cases):
        o = oracle.american_solve(100.0, c["K"], c["T"], c["sig"], c["r"], c["r"] - c["q"], c["call"], 1 if c["b76"] else 0, c["n"],
                                  c["m"], c["rann"], divs_tau=c["divs"], trace=True)
        assert res.status[i] == 0, (c, res.status[i])
        _v_close(res.V[i, : c["n"] + 1], o.V, c)
        assert abs(res.price[i] - o.price) <= REL * max(abs(o.price), 1e-2), c
        if not c["divs"]:
            assert np.array_equal(res.ex_trace[i, : o.ex_trace.shape[0]], o.ex_trace), c
        else:
            # the returned interval, clipped to the nodes whose value is above the noise floor
            live = np.nonzero(np.abs(o.V[1:-1]) > 1e-13 * np.max(np.abs(o.V)))[0]
            def clip(lo, hi):
                if hi < 0 or len(live) == 0:
                    return (-1, -1)
                lo2, hi2 = max(lo, int(live[0])), min(hi, int(live[-1]))
                return (lo2, hi2) if lo2 <= hi2 else (-1, -1)
            assert clip(int(res.idx[i, 2]), int(res.idx[i, 3])) == clip(o.ex_lo, o.ex_hi), c


@pytest.mark.parametrize("seed", [41, 42, 43])
def test_barrier_random_small_grids(seed):
    rng = np.random.default_rng(seed)
    parts, specs = [], []
    for _ in range(40):
        n = int(rng.integers(8, 514))
        m = int(rng.integers(6, 420))
        T = float(rng.choice([0.05, 0.25, 1.0, 2.0]))
        c = dict(n=n, m=m, T=T, sig=float(rng.uniform(0.08, 0.7)), r=float(rng.uniform(0.0, 0.1)), q=float(rng.uniform(0.0, 0.05)),
                 K=float(rng.uniform(80, 125)), call=bool(rng.uniform() < 0.5))
        kind = str(rng.choice(["down-and-out", "up-and-out"]))
        lo, up = (float(rng.uniform(60, 95)), None) if kind == "down-and-out" else (None, float(rng.uniform(110, 160)))
        mon = sorted(float(x) for x in rng.uniform(0.05, 0.95, int(rng.integers(0, 9))) * T)
        reb = float(rng.choice([0.0, 1.5, 4.0]))
        specs.append((c, kind, lo, up, mon, reb))
        parts.append(hostlogic.build_barrier_batch([100.0], [c["K"]], [T], [c["sig"]], [c["r"]], [c["r"] - c["q"]], [c["call"]],
                                                   [kind], [lo], [up], [reb], [mon], [n], [m]))
    res = solve_host(SolveBatch.concat(parts), dump=True)
    for i, (c, kind, lo, up, mon, reb) in enumerate(specs):
        o = oracle.barrier_solve(100.0, c["K"], c["T"], c["sig"], c["r"], c["r"] - c["q"], c["call"], kind, lo, up, reb, mon,
                                 c["n"], c["m"], True)
        assert res.status[i] == 0, (c, res.status[i])
        _v_close(res.V[i, : c["n"] + 1], o.V, (c, kind, lo, up, reb, len(mon)))
        assert abs(res.price[i] - o.price) <= REL * max(abs(o.price), 1e-2), c
