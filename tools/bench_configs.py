"""The other BASELINE.json configurations on one GPU (bench.py measures configs[2] = cfg3).

    python tools/bench_configs.py [cfg1 cfg2 am400 cfg4 cfg5 ...]  > gpurun_out/configs.jsonl

One JSON line per configuration: device-resident throughput (CUDA events, best of 3 after a warm-up),
FP64 roofline fraction with the configuration's algorithmic flops per node-step (22 American, 13
European/barrier), and parity of a seeded subsample against the CPU oracle.
Synthetic inputs follow SURVEY.md 8d (seed 20251018 + config index).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from finite_difference_b200 import DeviceBatch, _lib, hostlogic, solve_host  # noqa: E402
from finite_difference_b200.batch import SolveBatch  # noqa: E402
from oracle import oracle  # noqa: E402

SEED = 20251018


def timed(batch: SolveBatch, reps: int = 3):
    dev = DeviceBatch(batch)
    dev.solve()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_launch = dev.solve()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, n_launch, dev.results()


def line(name, desc, batch, ms, flops_per_ns, peak, parity, extra=None):
    ns = batch.node_steps()
    out = {"config": name, "workload": desc, "n_solves": batch.n_solves, "node_steps": ns, "ms": ms,
           "value": ns / ms * 1e3, "unit": "node-steps/s", "n_gpus": 1,
           "roofline": {"bound": "fp64", "flops_per_node_step": flops_per_ns, "achieved": ns * flops_per_ns / ms / 1e9,
                        "peak": peak, "unit": "TFLOP/s", "frac": ns * flops_per_ns / ms / 1e9 / peak},
           "parity": parity}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def parity_american(batch, res, idx, model, T=1.0):
    """Dividend-free American solves of `batch` against the oracle."""
    worst, same = 0.0, True
    steps = batch.steps_per_solve()
    for i in idx:
        o = oracle.american_solve(float(batch.spot[i]), float(batch.strike[i]), T, float(batch.sigma[i]), float(batch.r[i]),
                                  float(batch.b[i]), bool(batch.flags[i] & 1), model, int(batch.n_space[i]), int(steps[i]),
                                  int(batch.rannacher[i]), 4.5, [])
        worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
        same &= (int(res.idx[i, 2]), int(res.idx[i, 3])) == (o.ex_lo, o.ex_hi)
    return {"n": len(idx), "max_rel_err_price": worst, "exercise_intervals_identical": bool(same), "tolerance": 1e-10}


def cfg1(peak):
    """Single American put, 400 x 400, through the drop-in class (latency of one price_log call)."""
    import datetime as dt
    from finite_difference_b200 import fd_american_option_pricer as m, utils
    c = utils.create_rate_df_iso(float(np.exp(0.05) - 1.0))
    pr = m.AmericanFDMPricer(100.0, 100.0, dt.date(2025, 7, 28), dt.date(2026, 7, 28), 0.2, "put", c, c)
    pr.price_log()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        p = pr.price_log()
    wall = (time.perf_counter() - t0) / reps
    o = oracle.american_solve(100.0, 100.0, 1.0, 0.2, pr.discount_rate_nacc, pr.carry_rate_nacc, False, 0, 400, 400, 2)
    print(json.dumps({"config": "cfg1", "workload": "single American put via fd_american_option_pricer.AmericanFDMPricer.price_log, "
                      "400 nodes x 400 steps (one solve per call, host class + C ABI + D2H of the value vector)",
                      "price": p, "reference_price": 6.089992301161, "rel_err_vs_oracle": abs(p - o.price) / o.price,
                      "exercise_interval": list(pr.last_exercise_interval), "oracle_interval": [o.ex_lo, o.ex_hi],
                      "ms_per_call": wall * 1e3, "value": 160000 / wall, "unit": "node-steps/s", "n_gpus": 1}), flush=True)


def cfg2_inputs(n, seed=SEED + 2):
    rng = np.random.default_rng(seed)
    K = rng.uniform(80, 120, n); sig = rng.uniform(0.10, 0.50, n); r = rng.uniform(0.01, 0.08, n)
    q = rng.uniform(0.0, 0.03, n); T = rng.choice([0.5, 1.0, 2.0], n); B = 100.0 * rng.uniform(0.70, 0.95, n)
    mons = [[k / 12.0 for k in range(1, int(round(12 * t)) + 1)] for t in T]
    return K, sig, r, q, T, B, mons


def barrier_batch(n, sig_shift=0.0, seed=SEED + 2):
    K, sig, r, q, T, B, mons = cfg2_inputs(n, seed)
    return hostlogic.build_barrier_batch(np.full(n, 100.0), K, T, sig + sig_shift, r, r - q, np.ones(n, bool),
                                         ["down-and-out"] * n, list(B), [None] * n, np.zeros(n), mons, 400, 400), (K, sig, r, q, T, B, mons)


def cfg2(peak):
    n = 10000
    for label, shifts in (("price: 1 KO solve per option", (0.0,)), ("greeks: 3 solves per option (sigma, sigma +- 1e-3)", (0.0, 1e-3, -1e-3))):
        parts = [barrier_batch(n, s)[0] for s in shifts]
        batch = SolveBatch.concat(parts)
        ms, _, res = timed(batch)
        K, sig, r, q, T, B, mons = cfg2_inputs(n)
        worst = 0.0
        for i in range(64):
            o = oracle.barrier_solve(100.0, K[i], T[i], sig[i], r[i], r[i] - q[i], True, "down-and-out", B[i], None, 0.0,
                                     mons[i], 400, 400, True)
            worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
        line("cfg2", f"10k discrete down-and-out barrier calls, CN with monthly monitoring, 400 x 400; {label}", batch, ms, 13.0,
             peak, {"n": 64, "max_rel_err_price": worst, "tolerance": 1e-10})


def american_batch(n, n_space, n_time, b76, seed, T=1.0, divs=False, sig_shift=0.0):
    rng = np.random.default_rng(seed)
    K = rng.uniform(180, 280, n); sig = rng.uniform(0.15, 0.45, n) + sig_shift
    r = np.log1p(rng.uniform(0.03, 0.10, n)); call = rng.uniform(size=n) < 0.5
    dts = None
    if divs:
        cash = rng.uniform(2, 9, (n, 2))
        dts = [[(T - 256.0 / 365.0, float(cash[i, 1])), (T - 46.0 / 365.0, float(cash[i, 0]))] for i in range(n)]
    return hostlogic.build_american_batch(np.full(n, 229.74), K, T, sig, r, r, call, n_space, n_time, 2, b76, div_taus=dts), dts


def am400(peak):
    batch, _ = american_batch(100000, 400, 400, False, SEED + 1)
    ms, _, res = timed(batch)
    line("am400", "headline: 100k-option American CN batch, one solve each, 400 nodes x 400 steps (equity, no dividends)", batch, ms,
         22.0, peak, parity_american(batch, res, range(64), 0))


def cfg4(peak):
    n = 2048                      # one GPU's share of the 16 384-option book
    batch, dts = american_batch(n, 8192, 10000, False, SEED + 4, T=1.0, divs=True)
    ms, _, res = timed(batch, reps=1)
    worst, same = 0.0, True
    for i in range(4):
        o = oracle.american_solve(229.74, float(batch.strike[i]), 1.0, float(batch.sigma[i]), float(batch.r[i]), float(batch.b[i]),
                                  bool(batch.flags[i] & 1), 0, 8192, 10000, 2, 4.5, dts[i])
        worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
        same &= (int(res.idx[i, 2]), int(res.idx[i, 3])) == (o.ex_lo, o.ex_hi)
    line("cfg4", "high-resolution grids: 8192 nodes x 10 000 steps, American equity with two cash dividends; 2048 options = "
         "one GPU's share of the 16 384-option book", batch, ms, 22.0, peak,
         {"n": 4, "max_rel_err_price": worst, "exercise_intervals_identical": bool(same), "tolerance": 1e-10},
         {"status_max": int(res.status.max())})


def cfg5(peak):
    n = 62500                     # one GPU's share of the 1M-option mixed book: 62.5k American + 62.5k barrier
    h = 0.01
    am = [american_batch(n, 400, 400, False, SEED + 5, sig_shift=s)[0] for s in (0.0, h, -h, 2 * h, -2 * h)]
    ba = [barrier_batch(n, s, SEED + 5)[0] for s in (0.0, 1e-3, -1e-3)]
    batch = SolveBatch.concat(am + ba)
    ms, n_launch, res = timed(batch)
    flops = (5 * 22.0 + 3 * 13.0) / 8.0
    par = parity_american(batch, res, range(32), 0)
    line("cfg5", "Greeks by bump-and-reprice on a mixed book: 62.5k American (5 grids: sigma, +-h, +-2h) + 62.5k barrier KO "
         "(3 grids) = one GPU's share of the 1M-option book, 400 x 400", batch, ms, flops, peak, par, {"launches": n_launch})


def dbfp(peak):
    """Barrier scenario sweep of run_config_scenarios.py on its own grid: DiscreteBarrierFDMPricer, put, valuation
    2025-07-28, maturity 2025-08-28, weekday monitoring, num_time_steps = 500 -> 2133 space nodes; 2 PDE solves per
    row (sigma, sigma + 1e-4).  Device-resident kernel time and the end-to-end time of the runner (host pricer
    construction + descriptors + H2D + solves + D2H + result rows)."""
    import datetime as dt
    import time
    import pandas as pd
    from finite_difference_b200 import run_config_scenarios as rcs
    from finite_difference_b200.discrete_barrier_fdm_pricer import price_and_greeks_batch
    n = 4096
    rng = np.random.default_rng(SEED + 6)
    val, mat = dt.date(2025, 7, 28), dt.date(2025, 8, 28)
    mons = [d for d in (val + dt.timedelta(days=i) for i in range(32)) if d.weekday() < 5]
    down = rng.uniform(size=n) < 0.5
    frame = pd.DataFrame(dict(
        scenario_name=[f"s{i}" for i in range(n)], S0=229.74, K=rng.uniform(180, 280, n), sigma=rng.uniform(0.15, 0.45, n),
        rate=rng.choice(np.round(np.linspace(0.03, 0.10, 15), 4), n), barrier_type=np.where(down, "down-and-out", "up-and-out"),
        upper_barrier=np.where(down, np.nan, rng.uniform(240, 330, n)), lower_barrier=np.where(down, rng.uniform(150, 220, n), np.nan),
        FA_price=np.nan, FA_delta=np.nan, FA_gamma=np.nan, FA_vega=np.nan))
    base = dict(valuation=val, maturity=mat, monitor_dates=mons, opt_type="put", num_time_steps=500)
    cache = {}
    pricers = [rcs._build_pricer(r["S0"], r["K"], r["sigma"], r["rate"], r["barrier_type"], rcs._opt(r["upper_barrier"]),
                                 rcs._opt(r["lower_barrier"]), _curve_cache=cache, **base) for _, r in frame.iterrows()]
    args = []
    for p in pricers:
        args.append(p._solve_args(True)); args.append(p._solve_args(True, sigma=p.sigma + 1e-4))
    batch = hostlogic.build_dbfp_batch(**{k: [a[k] for a in args] for k in args[0]})
    ms, _, res = timed(batch)
    worst = 0.0
    for i in range(0, 16, 2):
        a = args[i]
        o = oracle.dbfp_solve(a["S_min"], a["S_max"], a["S0_eff"], a["S0_eff"], a["K"], a["T"], a["sigma"], a["r"], a["b"], a["q"],
                              a["is_call"], a["ko_type"], a["lower_barrier"], a["upper_barrier"], a["rebate_amount"],
                              a["rebate_at_hit"], a["monitor_times"], a["num_space_nodes"], a["num_time_steps"],
                              a["rannacher_steps"])
        worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
    rcs.run_scenarios_frame(frame.iloc[:64], base)            # warm-up (scipy import, workspace)
    t0 = time.perf_counter()
    df = rcs.run_scenarios_frame(frame, base)                 # vectorised host path
    e2e_s = time.perf_counter() - t0
    rcs.run_scenarios_frame(frame.iloc[:64], base, vectorised=False)
    t0 = time.perf_counter()
    df_obj = rcs.run_scenarios_frame(frame, base, vectorised=False)
    e2e_obj_s = time.perf_counter() - t0
    worst_vec = float(np.max(np.abs(df["model_price"].to_numpy() - df_obj["model_price"].to_numpy())
                             / np.maximum(np.abs(df_obj["model_price"].to_numpy()), 1e-2)))
    line("dbfp", f"barrier scenario sweep (run_config_scenarios shape, DiscreteBarrierFDMPricer): {n} rows x 2 solves, "
         "2133 space nodes x 500 steps, daily monitoring, Rannacher 2", batch, ms, 13.0, peak,
         {"n": 8, "max_rel_err_price": worst, "tolerance": 1e-10},
         {"e2e_runner": {"value": batch.node_steps() / e2e_s, "unit": "node-steps/s", "seconds": e2e_s,
                         "path": "run_config_scenarios.run_scenarios_frame -> BarrierScenarioBook (numpy over rows) -> "
                                 "fdm_solve_batch_host -> result frame",
                         "per_object_path": {"value": batch.node_steps() / e2e_obj_s, "seconds": e2e_obj_s,
                                             "note": "one Python pricer object per row"},
                         "max_rel_price_diff_between_paths": worst_vec}})


ALL = {"cfg1": cfg1, "cfg2": cfg2, "am400": am400, "cfg4": cfg4, "cfg5": cfg5, "dbfp": dbfp}

if __name__ == "__main__":
    oracle.build()
    peak, _ = _lib.fp64_peak_tflops()
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name](peak)
