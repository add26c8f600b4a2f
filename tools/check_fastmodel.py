"""Dev tool: rounding distance between the kernel's arithmetic model (tools/fastmodel.c) and the oracle."""
import ctypes as C, sys, time, math
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle
L = C.CDLL("tools/libfastmodel.so")
L.fastmodel_solve.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
    C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
    C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]

def american(spot, strike, T, sigma, r, b, is_call, model, n, M, ran=2):
    t0 = time.time()
    o = oracle.american_solve(spot, strike, T, sigma, r, b, is_call, model, n, M, ran)
    t1 = time.time()
    V = np.empty(n + 1); lo = C.c_int(); hi = C.c_int()
    x_min = math.log(o.S_min)
    L.fastmodel_solve(x_min, o.dx, n, sigma, r, b, model, int(is_call), 1, ran, M, T, spot, strike, 1,
                      0, 0.0, 0.0, None, 0, V.ctypes.data_as(C.POINTER(C.c_double)), C.byref(lo), C.byref(hi))
    i = o.spot_idx
    rel = abs(V[i] - o.V[i]) / abs(o.V[i])
    scale = np.max(np.abs(o.V))
    print(f"n={n} M={M} sig={sigma} call={is_call} model={model}: price {o.V[i]:.12g} rel {rel:.2e} "
          f"maxabs/scale {np.max(np.abs(V - o.V))/scale:.2e} ex oracle=({o.ex_lo},{o.ex_hi}) fast=({lo.value},{hi.value}) oracle {t1-t0:.2f}s")

for (n, M) in [(400, 400), (500, 500), (500, 1000), (1000, 1000), (2048, 2000)]:
    for sig in (0.15, 0.3, 0.45):
        american(100, 100, 1.0, sig, 0.05, 0.05, False, 0, n, M)
    american(229.74, 250, 31 / 365, 0.3, 0.06, 0.06, True, 1, n, M)
    american(229.74, 200, 31 / 365, 0.2, 0.06, 0.06, False, 1, n, M)
if len(sys.argv) > 1:
    for sig in (0.15, 0.3, 0.45):
        american(229.74, 230, 1.0, sig, 0.06, 0.06, False, 0, 8192, 10000)
        american(229.74, 230, 1.0, sig, 0.06, 0.06, True, 0, 8192, 10000)
