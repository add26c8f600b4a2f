"""Dev tool: time the main shapes with the library in FDM_B200_LIB (see tools/time_variants.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_variants import american, barrier, timeit
from finite_difference_b200 import _lib
print("lib", _lib.lib_path(), "NO_LEAN" if os.environ.get("FDM_B200_NO_LEAN") else "")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
timeit("american b76 500x500", american(n, 500, 500, True))
timeit("american eq 400x400", american(n, 400, 400, False))
timeit("barrier 400x400", barrier(n, 400, 400))
timeit("american eq 300x300", american(n, 300, 300, False))
timeit("barrier 300x544", barrier(n, 300, 544))
timeit("american eq 320x320", american(n, 320, 320, False))
