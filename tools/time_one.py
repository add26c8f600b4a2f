"""Dev tool: time only the cfg3-shaped American batch (40000 solves) with the library in FDM_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_variants import american, timeit
from finite_difference_b200 import _lib
print("lib", _lib.lib_path())
timeit("american b76 500x500", american(int(sys.argv[1]) if len(sys.argv) > 1 else 40000, 500, 500, True))
