"""Dev tool: where the time of dividend solves goes: phases (setup_phase) vs the spline jump."""
import os, sys, copy
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_variants import american, american_div, timeit
from finite_difference_b200 import hostlogic
print("NO_LEAN" if os.environ.get("FDM_B200_NO_LEAN") else "LEAN")
n = 20000
timeit("eq 500x500 T=1 no div (2 phases)", american(n, 500, 500, False))
b = american(n, 500, 500, False); b.rannacher[:] = 0
timeit("   rannacher 0 (1 phase)", b)
d = american_div(n, 500, 500)
timeit("eq 500x500 T=1 2 div", d)
z = copy.deepcopy(d); z.seg_div[:] = 0.0
timeit("   same 3 segments, no jump", z)
z2 = copy.deepcopy(z); z2.seg_restart[:] = 0
timeit("   3 segments, no jump, no Rannacher restarts (3 phases)", z2)
z3 = copy.deepcopy(d); z3.seg_restart[:] = 0
timeit("   3 segments, jumps, no Rannacher restarts", z3)
