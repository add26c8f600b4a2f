"""Dev tool: run with FDM_B200_LIB=build/variants/libfdm_prof.so (built with -DFDM_LEAN_PROFILE)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, american_div
from finite_difference_b200 import DeviceBatch
for name, b in (("no div", american(20000, 500, 500, False)), ("2 div", american_div(20000, 500, 500)), ("b76 cfg3", american(20000, 500, 500, True))):
    print("==", name, flush=True)
    d = DeviceBatch(b); d.solve(); torch.cuda.synchronize()
