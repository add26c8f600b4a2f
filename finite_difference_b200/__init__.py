"""finite_difference_b200 -- B200-native batched finite-difference option pricing.

Python host code builds structure-of-arrays descriptors of 1-D log-space theta/Crank-Nicolson
PDE solves and hands them through a C ABI (ctypes, ``include/fdm_b200.h``) to hand-written
sm_100a CUDA kernels.  There is no CPU fallback.
"""
from . import _lib
from .batch import DeviceBatch, SolveBatch, SolveResults, solve_host

__all__ = ["SolveBatch", "SolveResults", "DeviceBatch", "solve_host", "_lib"]
__version__ = "0.1.0"
