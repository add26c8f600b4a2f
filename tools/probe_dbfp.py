"""Dev tool: time the barrier scenario runner's own solve batch (bench.py dbfp: 2133 x 500, 4096 solves) and the cfg4
shape, device-resident, best of 5.  FDM_B200_LIB selects the library (A/B of kernel variants in one gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datetime as dt
import numpy as np, pandas as pd
from finite_difference_b200 import hostlogic, _lib, run_config_scenarios as rcs
from tools.time_variants import timeit, cfg4, american_T, barrier_T


def dbfp_batch(n=2048, seed=20251018 + 6):
    rng = np.random.default_rng(seed)
    val, mat = dt.date(2025, 7, 28), dt.date(2025, 8, 28)
    mons = [d for d in (val + dt.timedelta(days=i) for i in range(32)) if d.weekday() < 5]
    down = rng.uniform(size=n) < 0.5
    frame = pd.DataFrame(dict(
        scenario_name=[f"s{i}" for i in range(n)], S0=229.74, K=rng.uniform(180, 280, n), sigma=rng.uniform(0.15, 0.45, n),
        rate=rng.choice(np.round(np.linspace(0.03, 0.10, 15), 4), n), barrier_type=np.where(down, "down-and-out", "up-and-out"),
        upper_barrier=np.where(down, np.nan, rng.uniform(240, 330, n)), lower_barrier=np.where(down, rng.uniform(150, 220, n), np.nan)))
    base = dict(valuation=val, maturity=mat, monitor_dates=mons, opt_type="put", num_time_steps=500)
    cache = {}
    a_ = []
    for _, r in frame.iterrows():
        p = rcs._build_pricer(r["S0"], r["K"], r["sigma"], r["rate"], r["barrier_type"], rcs._opt(r["upper_barrier"]),
                              rcs._opt(r["lower_barrier"]), _curve_cache=cache, **base)
        a_.append(p._solve_args(True))
        a_.append(p._solve_args(True, sigma=p.sigma + 1e-4))
    return hostlogic.build_dbfp_batch(**{k: [a[k] for a in a_] for k in a_[0]})


if __name__ == "__main__":
    print("lib", _lib.lib_path())
    what = sys.argv[1:] or ["dbfp", "cfg4"]
    if "dbfp" in what: timeit("dbfp 2133x500 x4096", dbfp_batch(), reps=5)
    if "cfg4" in what: timeit("cfg4 8192x2500 2 divs x592", cfg4(592, 8192, 2500, divs=True), reps=3)
    if "cfg4full" in what: timeit("cfg4 8192x10000 2 divs x296", cfg4(296, 8192, 10000, divs=True), reps=3)
    if "am1000" in what: timeit("am 1000x1000 x4000", american_T(4000, 1000, 1000, 1.0), reps=3)
    if "am3000" in what: timeit("am 3000x1000 x1184", american_T(1184, 3000, 1000, 1.0), reps=3)
    if "am6000" in what: timeit("am 6000x1000 x592", american_T(592, 6000, 1000, 1.0), reps=3)
    if "bar8192" in what: timeit("barrier 8192x1000 x592", barrier_T(592, 8192, 1000, 1.0), reps=3)
