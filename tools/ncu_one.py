"""Dev tool: one launch of the cfg3-shaped American batch (for ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import DeviceBatch
from tools.time_variants import american
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = DeviceBatch(american(n, 500, 500, True))
dev.solve(); torch.cuda.synchronize()
dev.solve(); torch.cuda.synchronize()
print("done", float(dev.price.sum()))
