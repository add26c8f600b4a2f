"""Dev tool: per-phase cycle breakdown of one warp-step (library built with -DFDM_PROFILE_CYCLES, path in FDM_B200_LIB).
The kernel prints the averages of the solve in block 0; run with 1 solve (solo latency) and a full wave (loaded)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, barrier
from finite_difference_b200 import DeviceBatch
for n_opt in (1, 148 * 4, 148 * 8, 148 * 8 * 4):
    for nm, b in (("am500", american(n_opt, 500, 500, True)), ("am400", american(n_opt, 400, 400, False)), ("bar400", barrier(n_opt, 400, 400))):
        print(f"--- {nm} n_solves={n_opt}", flush=True)
        dev = DeviceBatch(b); dev.solve(); torch.cuda.synchronize()
