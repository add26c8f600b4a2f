"""One warm-up launch and one profiled launch of a named shape (for ncu):
   python tools/ncu_shapes.py {am500|am500div|am400|bar400|cfg4|dbfp} [n_solves]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import DeviceBatch
from tools.time_variants import american, american_div, barrier
from tools.time_cta import cfg4
from tools import time_cta3
shape = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
batch = {"am500": lambda: american(n or 20000, 500, 500, True), "am500div": lambda: american_div(n or 5000, 500, 500), "am400": lambda: american(n or 20000, 400, 400, False),
         "bar400": lambda: barrier(n or 20000, 400, 400), "cfg4": lambda: cfg4(n or 148, 8192, 1000, divs=False),
         "dbfp": lambda: barrier(n or 1184, 2132, 500),
         "bar8192": lambda: time_cta3.barrier(n or 148, 8192, 2000, 0.2), "am8192": lambda: time_cta3.american(n or 148, 8192, 2000, 0.2)}[shape]()
dev = DeviceBatch(batch)
dev.solve(); torch.cuda.synchronize()
dev.solve(); torch.cuda.synchronize()
print("done", shape, batch.n_solves, float(dev.price.sum()))
