"""Dev tool: time the CTA-kernel shapes (device-resident, best of 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_variants import american, barrier, cfg4, timeit
timeit("dbfp-like barrier 2133x500", barrier(2368, 2132, 500))
timeit("cfg4 8192x2500 2 divs", cfg4(592, 8192, 2500, divs=True))
timeit("am 1000x1000", american(4000, 1000, 1000, False))
timeit("am 2048x1000", american(2000, 2048, 1000, False))
timeit("am 4096x1000", american(1184, 4096, 1000, False))
timeit("barrier 8192x1000", barrier(592, 8192, 1000))
