"""Dev tool: per-step latency of one warp vs resident warps per SM (batch sizes 148*k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, barrier
from finite_difference_b200 import DeviceBatch
def run(name, b, steps):
    dev = DeviceBatch(b); dev.solve(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dev.solve(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {b.n_solves} solves {best*1e3:.1f} us -> {best*1e-3/steps*1.965e9:.0f} cycles/step (if one wave)", flush=True)
for k in (1, 2, 4, 8):
    run(f"american 400x400 k={k}", american(148 * k, 400, 400, False), 400)
for k in (1, 4, 8):
    run(f"american 500x500 k={k}", american(148 * k, 500, 500, True), 500)
for k in (1, 4, 8, 12):
    run(f"barrier 400x400 k={k}", barrier(148 * k, 400, 400), 400)
