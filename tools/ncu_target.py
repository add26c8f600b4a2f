"""Dev tool: one launch of a chosen workload for ncu captures.  usage: ncu_target.py {am400|am500|bar400} [n_solves]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, barrier
from finite_difference_b200 import DeviceBatch
which = sys.argv[1] if len(sys.argv) > 1 else "am400"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 11840
b = {"am400": lambda: american(n, 400, 400, False), "am500": lambda: american(n, 500, 500, True),
     "bar400": lambda: barrier(n, 400, 400)}[which]()
dev = DeviceBatch(b)
for _ in range(2):
    dev.solve()
torch.cuda.synchronize()
print("done", which, n, float(dev.price.sum()))
