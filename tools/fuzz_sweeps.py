"""Dev tool: the long random sweeps behind tests/test_gpu_fuzz_*.py (run on a GPU box; ~30 s for the default ranges).
   python tools/fuzz_sweeps.py [first_seed] [n_seeds]
Each sweep draws seeded random cases, runs them through the C ABI and compares every node of the value vector with the
oracle (1e-10 relative + 1e-12 of the vector's scale), the price (1e-10) and, where it is well defined, the exercise
interval.  The tests keep a handful of seeds; this runs as many as asked for and prints what fails."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def sweep_tests(first, count):
    import test_gpu_fuzz_cta as big, test_gpu_fuzz_small as small
    bad = 0
    for seed in range(first, first + count):
        for fn in (big.test_american_random_geometry, big.test_barrier_random_geometry,
                   small.test_american_random_small_grids, small.test_barrier_random_small_grids):
            try:
                fn(seed)
            except AssertionError as e:
                bad += 1
                print("FAIL", fn.__name__, seed, str(e)[:400])
    return bad


def sweep_cta_dividends(first, count):
    from finite_difference_b200 import hostlogic, solve_host
    from finite_difference_b200.batch import SolveBatch
    from oracle import oracle
    bad = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        cases = []
        for _ in range(10):
            T = float(rng.choice([0.1, 0.5, 1.0, 2.0]))
            divs = [(t, float(rng.uniform(0.5, 6.0))) for t in sorted(float(x) for x in rng.uniform(0.08, 0.92, int(rng.choice([1, 2, 3]))) * T)]
            cases.append(dict(n=int(rng.integers(514, 8194)), m=int(rng.integers(20, 200)), T=T, sig=float(rng.uniform(0.1, 0.6)),
                              r=float(rng.uniform(0.0, 0.1)), q=float(rng.uniform(0.0, 0.05)), K=float(rng.uniform(80, 125)),
                              call=bool(rng.uniform() < 0.5), rann=int(rng.choice([0, 2, 4])), divs=divs))
        parts = [hostlogic.build_american_batch([100.0], [c["K"]], [c["T"]], [c["sig"]], [c["r"]], [c["r"] - c["q"]], [c["call"]], [c["n"]],
                                                [c["m"]], [c["rann"]], [False], div_taus=[c["divs"]]) for c in cases]
        res = solve_host(SolveBatch.concat(parts), dump=True)
        for i, c in enumerate(cases):
            o = oracle.american_solve(100.0, c["K"], c["T"], c["sig"], c["r"], c["r"] - c["q"], c["call"], 0, c["n"], c["m"], c["rann"],
                                      divs_tau=c["divs"])
            V = res.V[i, : c["n"] + 1]
            rel = np.abs(V - o.V) / (np.abs(o.V) + 1e-12 * np.max(np.abs(o.V)))
            if res.status[i] != 0 or rel.max() > 1e-10 or abs(res.price[i] - o.price) > 1e-10 * max(abs(o.price), 1e-2):
                bad += 1
                print("FAIL cta+dividends", seed, i, c, "status", res.status[i], "max rel %.3g" % rel.max())
    return bad


def sweep_rowvar(first, count):
    from datetime import date, timedelta
    from finite_difference_b200 import rowvar
    from oracle import oracle
    bad = 0
    val = date(2025, 7, 28)
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        marches, refs, cs = [], [], []
        for _ in range(24):
            days = int(rng.choice([30, 90, 182, 365, 730]))
            bt = str(rng.choice(["none", "down-and-out", "up-and-out", "double-out"]))      # "-in": the reference class returns 0
            lo = float(rng.uniform(70, 95)) if bt in ("down-and-out", "double-out") else None
            up = float(rng.uniform(108, 140)) if bt in ("up-and-out", "double-out") else None
            mons = sorted({val + timedelta(days=int(d)) for d in rng.integers(1, days, int(rng.integers(1, 13)))})
            steps = int(rng.integers(4, 300))
            a = dict(spot=100.0, strike=float(rng.uniform(85, 115)), valuation_date=val, maturity_date=val + timedelta(days=days),
                     volatility=float(rng.uniform(0.1, 0.6)), option_type=str(rng.choice(["call", "put"])), barrier_type=bt,
                     lower_barrier=lo, upper_barrier=up, monitoring_dates=list(mons), flat_rate_nacc=float(rng.uniform(0.0, 0.1)),
                     num_space_nodes=int(rng.integers(40, 1000)), num_time_steps=steps, rannacher_steps=steps)
            try:
                refs.append(oracle.pricer2_price(**a)); marches.append(rowvar.march_inputs(**a)); cs.append(a)
            except Exception:
                continue
        res = rowvar.solve_rowvar_host(rowvar.RowVarBatch(marches), dump=True)
        for i, (p, V, _) in enumerate(refs):
            n = len(V)
            ok = res.status[i] == 0 and abs(res.price[i] - p) <= 1e-10 * max(abs(p), 1e-2) and \
                np.all(np.abs(res.V[i, :n] - V) <= 1e-10 * np.abs(V) + 1e-12 * np.max(np.abs(V)))
            if not ok:
                bad += 1
                print("FAIL rowvar", seed, i, {k: v for k, v in cs[i].items() if k != "monitoring_dates"}, res.price[i], p)
    return bad


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    total = sweep_tests(first, count) + sweep_cta_dividends(first + 1000, max(count // 2, 1)) + sweep_rowvar(first + 2000, max(count // 2, 1))
    print("failures", total)
    sys.exit(1 if total else 0)
