"""Dev tool: CTA kernel on cfg4's real stiffness (8192 x 10000) and the barrier runner's grid (2133 x 500)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_cta import cfg4, timeit
from tools.time_variants import barrier
from finite_difference_b200 import _lib
print("lib", _lib.lib_path())
timeit("cfg4 8192x10000 nodiv, 148 solves", cfg4(148, 8192, 10000, divs=False), reps=1)
timeit("cfg4 8192x10000 divs, 148 solves", cfg4(148, 8192, 10000), reps=1)
timeit("american 2048x2000, 592 solves", cfg4(592, 2048, 2000, divs=False), reps=1)
timeit("barrier 2132x500, 1184 solves", barrier(1184, 2132, 500), reps=2)
