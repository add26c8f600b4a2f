"""Dev tool: CTA-per-solve kernel on shapes with the time step of the real configurations (cfg4: dt = 1e-4)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import hostlogic, DeviceBatch, _lib

def american(n_opt, n, m, T, seed=4):
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(0.15, 0.45, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, T, sig, r, r, call, n, m, 2, False)

def barrier(n_opt, n, m, T, seed=2):
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(80, 120, n_opt); sig = rng.uniform(0.10, 0.50, n_opt); r = rng.uniform(0.01, 0.08, n_opt)
    b = r - rng.uniform(0, 0.03, n_opt); B = 100.0 * rng.uniform(0.70, 0.95, n_opt)
    mons = [[k / 12.0 for k in range(1, int(12 * T) + 1)]] * n_opt
    return hostlogic.build_barrier_batch(np.full(n_opt, 100.0), K, np.full(n_opt, T), sig, r, b, np.ones(n_opt, bool),
                                         ["down-and-out"] * n_opt, list(B), [None] * n_opt, np.zeros(n_opt), mons,
                                         np.full(n_opt, n), np.full(n_opt, m))

def timeit(name, batch, reps=3):
    dev = DeviceBatch(batch)
    dev.solve(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dev.solve(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ns = batch.node_steps()
    print(f"{name}: {batch.n_solves} solves, {best:.2f} ms -> {ns / best * 1e3:.3e} node-steps/s", flush=True)

if __name__ == "__main__":
    print("lib", _lib.lib_path())
    timeit("am 8192x2000 T=0.2 (dt 1e-4), 148", american(148, 8192, 2000, 0.2))
    timeit("am 4096x2000 T=0.2, 148", american(148, 4096, 2000, 0.2))
    timeit("am 2048x2000 T=1, 592", american(592, 2048, 2000, 1.0))
    timeit("am 1000x1000 T=1, 1184", american(1184, 1000, 1000, 1.0))
    timeit("bar 2133x500 T=1, 1184", barrier(1184, 2133, 500, 1.0))
    timeit("bar 8192x2000 T=0.2, 148", barrier(148, 8192, 2000, 0.2))
