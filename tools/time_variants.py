"""Dev tool: time the solve kernels on synthetic batches (CUDA events, device-resident descriptors)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import hostlogic, DeviceBatch, _lib

def american(n_opt, n, m, b76, seed=1):
    rng = np.random.default_rng(seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, 31.0 / 365.0 if b76 else 1.0, sig, r, r, call, n, m, 2, b76)

def barrier(n_opt, n, m, seed=2):
    rng = np.random.default_rng(seed)
    K = rng.uniform(80, 120, n_opt); sig = rng.uniform(0.10, 0.50, n_opt); r = rng.uniform(0.01, 0.08, n_opt)
    q = rng.uniform(0.0, 0.03, n_opt); T = rng.choice([0.5, 1.0, 2.0], n_opt); B = 100.0 * rng.uniform(0.70, 0.95, n_opt)
    mons = [[k / 12.0 for k in range(1, int(round(12 * t)) + 1)] for t in T]
    return hostlogic.build_barrier_batch(np.full(n_opt, 100.0), K, T, sig, r, r - q, np.ones(n_opt, bool), ["down-and-out"] * n_opt,
                                         list(B), [None] * n_opt, np.zeros(n_opt), mons, n, m)

def timeit(name, batch, reps=3):
    dev = DeviceBatch(batch)
    dev.solve(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dev.solve(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ns = batch.node_steps()
    print(f"{name}: {batch.n_solves} solves, {ns:.3e} node-steps, {best:.2f} ms -> {ns / best * 1e3:.3e} node-steps/s", flush=True)
    return ns / best * 1e3

def american_div(n_opt, n, m, seed=5):
    rng = np.random.default_rng(seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    cash = rng.uniform(2, 9, (n_opt, 2))
    dts = [[(109.0 / 365.0, float(cash[i, 1])), (319.0 / 365.0, float(cash[i, 0]))] for i in range(n_opt)]
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, 1.0, sig, r, r, call, n, m, 2, False, div_taus=dts)


def cfg4(n_opt, n, m, divs=True, seed=4):
    """cfg4-shaped batch: American equity, T = 1, two cash dividends."""
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    cash = rng.uniform(2, 9, (n_opt, 2))
    dts = [[(109.0 / 365.0, float(cash[i, 1])), (319.0 / 365.0, float(cash[i, 0]))] for i in range(n_opt)] if divs else None
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, 1.0, sig, r, r, call, n, m, 2, False, div_taus=dts)


def american_T(n_opt, n, m, T, seed=4):
    """American equity batch with a chosen maturity (so that dt matches a real configuration)."""
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, T, sig, r, r, call, n, m, 2, False)


def barrier_T(n_opt, n, m, T, seed=2):
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(80, 120, n_opt); sig = rng.uniform(0.10, 0.50, n_opt); r = rng.uniform(0.01, 0.08, n_opt)
    b = r - rng.uniform(0, 0.03, n_opt); B = 100.0 * rng.uniform(0.70, 0.95, n_opt)
    mons = [[k / 12.0 for k in range(1, int(12 * T) + 1)]] * n_opt
    return hostlogic.build_barrier_batch(np.full(n_opt, 100.0), K, np.full(n_opt, T), sig, r, b, np.ones(n_opt, bool),
                                         ["down-and-out"] * n_opt, list(B), [None] * n_opt, np.zeros(n_opt), mons,
                                         np.full(n_opt, n), np.full(n_opt, m))


if __name__ == "__main__":
    print("lib", _lib.lib_path())
    tf, ms = _lib.fp64_peak_tflops()
    print(f"fp64 peak microbench: {tf:.2f} TFLOP/s ({ms:.3f} ms)")
    nopt = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    timeit("american b76 500x500", american(nopt, 500, 500, True))
    timeit("american eq 400x400", american(nopt, 400, 400, False))
    timeit("american eq 256x256", american(nopt, 256, 256, False))
    timeit("american eq 128x128", american(nopt, 128, 128, False))
    timeit("barrier 400x400", barrier(nopt, 400, 400))
    timeit("american eq 500x500, 2 dividends", american_div(nopt // 4, 500, 500))
    tf, ms = _lib.fp64_peak_tflops()
    print(f"fp64 peak microbench (after): {tf:.2f} TFLOP/s ({ms:.3f} ms)")
