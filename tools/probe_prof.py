"""Dev tool: run with FDM_B200_LIB=build/variants/libfdm_prof.so (built with -DFDM_LEAN_PROFILE -DFDM_CTA_PROFILE)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, american_div, barrier, cfg4
from finite_difference_b200 import DeviceBatch
which = sys.argv[1:] or ["lean"]
cases = []
if "lean" in which:
    cases += [("no div", american(20000, 500, 500, False)), ("2 div", american_div(20000, 500, 500)), ("b76 cfg3", american(20000, 500, 500, True)),
              ("bar400", barrier(20000, 400, 400))]
if "cta" in which:
    cases += [("dbfp-like barrier 2133x500", barrier(1184, 2132, 500)), ("cfg4 8192x2500 divs", cfg4(296, 8192, 2500, divs=True)),
              ("am 1000x1000", american(2000, 1000, 1000, False)), ("am 4096x1000", american(600, 4096, 1000, False))]
for name, b in cases:
    print("==", name, flush=True)
    d = DeviceBatch(b); d.solve(); torch.cuda.synchronize()
