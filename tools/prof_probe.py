import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.time_variants import american, barrier
from finite_difference_b200 import DeviceBatch
for nm, b in (("am400", american(int(sys.argv[1]), 400, 400, False)), ("am500", american(int(sys.argv[1]), 500, 500, True)), ("bar400", barrier(int(sys.argv[1]), 400, 400))):
    dev = DeviceBatch(b); dev.solve(); torch.cuda.synchronize()
