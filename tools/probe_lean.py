"""Dev tool: time the lean-kernel shapes (device-resident, best of 3) with the library in FDM_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.time_variants import american, barrier, american_div, timeit
from finite_difference_b200 import _lib
print("lib", os.path.basename(_lib.lib_path()))
def srt(b):
    key = b.sigma ** 2 / b.dx ** 2 * b.seg_dt[b.seg_begin[:-1]]
    return b.take(np.argsort(key, kind="stable"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 59200
timeit("american b76 500x500 (cfg3 shape)", srt(american(n, 500, 500, True)))
timeit("american eq 400x400", srt(american(n, 400, 400, False)))
timeit("barrier 400x400", srt(barrier(n, 400, 400)))
timeit("american eq 500x500 2 div", srt(american_div(n // 2, 500, 500)))
from tools.time_variants import barrier_T
timeit("barrier 2133x500", srt(barrier(2368, 2132, 500)))
timeit("barrier 8192x1000", srt(barrier_T(592, 8192, 1000, 1.0)))
