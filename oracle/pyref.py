"""The UNMODIFIED Python reference as a timed CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box.  ``stage()`` (called by ``__graft_entry__.build()`` in the build
container) copies the reference's own hot-path files, byte for byte, into the git-ignored ``baseline/_ref/`` --
which travels to the box with the snapshot -- so that ``bench.py``'s ``cpu_baseline`` leg can execute the reference's
Python there ("executing the unmodified reference classes from baseline/_ref/").  Nothing under
``finite_difference_b200/`` imports this module and no reference source enters the git history.

``time_cfg3`` prices a sample of BASELINE config 3 (American Black-76 runner shape, ``price_log2()`` +
``greeks_log2()`` per option, fd_american_black76.py:533-625) with one process per host core and returns both the
throughput and the reference's outputs, which ``bench.py`` also compares with the GPU results: parity against the
reference itself, run on the box.
"""
from __future__ import annotations

import datetime as dt
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_DIR = os.path.join(ROOT, "baseline", "_ref")
SOURCE_DIR = "/root/reference"
HOT_PATH_FILES = ("fd_american_option_pricer.py", "fd_american_equity.py", "fd_american_black76.py",
                  "discrete_barrier_fdm_pricer_cn.py", "fd_discrete_barrier_option.py", "discrete_barrier_fdm_pricer.py",
                  "discrete_barrier_fdm_pricer_2.py", "utils.py", "run_american_scenarios.py", "run_config_scenarios.py")


def stage(source: str = SOURCE_DIR) -> bool:
    """Copy the hot-path reference files into baseline/_ref/ (no-op when the reference checkout is absent)."""
    if not os.path.isfile(os.path.join(source, HOT_PATH_FILES[0])):
        return False
    os.makedirs(STAGE_DIR, exist_ok=True)
    for f in HOT_PATH_FILES:
        src, dst = os.path.join(source, f), os.path.join(STAGE_DIR, f)
        if os.path.isfile(src) and (not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src)):
            shutil.copyfile(src, dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(STAGE_DIR, HOT_PATH_FILES[0])) or os.path.isfile(os.path.join(SOURCE_DIR, HOT_PATH_FILES[0]))


def reference_dir() -> str:
    return STAGE_DIR if os.path.isfile(os.path.join(STAGE_DIR, HOT_PATH_FILES[0])) else SOURCE_DIR


def _loader():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_loader
    ref_loader.REFERENCE_DIR = reference_dir()
    return ref_loader


_CURVES = {}


def _flat_curve(naca: float):
    import pandas as pd
    if naca not in _CURVES:
        utils = _loader().load_utils()
        df = utils.create_rate_df(naca)
        df["Date"] = pd.to_datetime(df["Date"], format="%Y/%m/%d").dt.strftime("%Y-%m-%d")
        _CURVES[naca] = df
    return _CURVES[naca]


def _warm(_):
    _loader().load_american("fd_american_black76")
    _loader().load_utils()
    return 0


def _cfg3_one(args):
    """One option of cfg3 through the reference: AmericanFwdFDMPricer, price_log2() then greeks_log2()."""
    F0, K, sigma, naca, is_call, n_space, n_time = args
    mod = _loader().load_american("fd_american_black76")
    curve = _flat_curve(naca)
    t0 = time.perf_counter()
    pr = mod.AmericanFwdFDMPricer(
        spot=F0, strike=K, valuation_date=dt.date(2025, 7, 28), maturity_date=dt.date(2025, 8, 28), sigma=sigma,
        option_type="call" if is_call else "put", discount_curve=curve, forward_curve=curve,
        num_space_nodes=n_space, num_time_steps=n_time, rannacher_steps=2)
    p2 = pr.price_log2()
    g = pr.greeks_log2()
    return (time.perf_counter() - t0, p2, g["price"], g["delta"], g["gamma"], g["vega"], g["theta"],
            pr.discount_rate_nacc, pr.time_to_expiry)


def time_cfg3(F0, K, sigma, naca, is_call, n_space=500, n_time=500, n_procs: int = 0):
    """Returns ``dict(seconds, node_steps, procs, price_log2, price, delta, gamma, vega, theta, r, T)`` for the sample."""
    import multiprocessing as mp
    import numpy as np
    n = len(K)
    procs = n_procs or max(1, len(os.sched_getaffinity(0)))
    procs = min(procs, n)
    jobs = [(float(F0), float(K[i]), float(sigma[i]), float(naca[i]), bool(is_call[i]), int(n_space), int(n_time))
            for i in range(n)]
    # spawn, not fork: the caller may hold a CUDA context.  Workers import the reference before the clock starts.
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_warm, range(procs), chunksize=1)
        t0 = time.perf_counter()
        rows = pool.map(_cfg3_one, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    arr = np.array(rows, dtype=float)
    return dict(seconds=wall, cpu_seconds=float(arr[:, 0].sum()), node_steps=n * n_space * 10 * n_time,   # 6 solves of n_time steps + 2 of 2 n_time
                procs=procs, price_log2=arr[:, 1], price=arr[:, 2], delta=arr[:, 3], gamma=arr[:, 4], vega=arr[:, 5],
                theta=arr[:, 6], r=arr[:, 7], T=arr[:, 8])
