"""Dev tool: throughput of the CTA-per-solve kernel on cfg4-shaped batches (8192 nodes, dividends)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import hostlogic, DeviceBatch, _lib

def cfg4(n_opt, n, m, divs=True, seed=4):
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    cash = rng.uniform(2, 9, (n_opt, 2))
    dts = [[(109.0 / 365.0, float(cash[i, 1])), (319.0 / 365.0, float(cash[i, 0]))] for i in range(n_opt)] if divs else None
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, 1.0, sig, r, r, call, n, m, 2, False, div_taus=dts)

def timeit(name, batch, reps=2):
    dev = DeviceBatch(batch)
    dev.solve(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dev.solve(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ns = batch.node_steps()
    print(f"{name}: {batch.n_solves} solves, {ns:.3e} node-steps, {best:.2f} ms -> {ns / best * 1e3:.3e} node-steps/s", flush=True)

if __name__ == "__main__":
    print("lib", _lib.lib_path())
    timeit("cfg4 8192x400 divs, 148 solves", cfg4(148, 8192, 400))
    timeit("cfg4 8192x400 nodiv, 148 solves", cfg4(148, 8192, 400, divs=False))
    timeit("cfg4 8192x400 nodiv, 296 solves", cfg4(296, 8192, 400, divs=False))
    timeit("2048x400 nodiv, 592 solves", cfg4(592, 2048, 400, divs=False))
    timeit("1000x400 nodiv, 592 solves", cfg4(592, 1000, 400, divs=False))
