"""Dev tool: one launch of a cfg4-shaped batch (8192 x 1000) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import DeviceBatch
from tools.time_cta import cfg4
dev = DeviceBatch(cfg4(148, 8192, 1000, divs=False))
dev.solve(); torch.cuda.synchronize()
dev.solve(); torch.cuda.synchronize()
print("done", float(dev.price.sum()))
