"""Dev tool: CTA-per-solve kernel timed per volatility band (stiffness of the time step), one band per launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import hostlogic, DeviceBatch, _lib
from tools.time_cta3 import timeit

def american(n_opt, n, m, T, sig_lo, sig_hi, seed=4):
    rng = np.random.default_rng(20251018 + seed)
    K = rng.uniform(180, 280, n_opt); sig = rng.uniform(sig_lo, sig_hi, n_opt)
    r = np.log1p(rng.uniform(0.03, 0.10, n_opt)); call = rng.uniform(size=n_opt) < 0.5
    return hostlogic.build_american_batch(np.full(n_opt, 229.74), K, T, sig, r, r, call, n, m, 2, False)

if __name__ == "__main__":
    print("lib", _lib.lib_path())
    for lo, hi in ((0.15, 0.20), (0.25, 0.30), (0.33, 0.36), (0.40, 0.45)):
        timeit(f"am 8192x2000 T=0.2 sigma {lo}-{hi}", american(148, 8192, 2000, 0.2, lo, hi))
    for lo, hi in ((0.15, 0.20), (0.25, 0.30), (0.40, 0.45)):
        timeit(f"am 2048x400 T=1 sigma {lo}-{hi}", american(592, 2048, 400, 1.0, lo, hi))
