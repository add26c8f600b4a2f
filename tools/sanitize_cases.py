"""The smallest case of every kernel family through the C ABI, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Families: fdm_lean_kernel (American / linear, with a dividend jump and a Rannacher restart), fdm_warp_kernel (sizes
outside the lean classes, dividend jump, knock-out), the lean -> warp hand-back (a degenerate operator the lean kernel
refuses), fdm_cta_kernel (4 / 5 / 8 / 16 warps, dividend jump, knock-out), the book fan-out / epilogue kernels
(American with dividends, barrier) and the device-pointer entries."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finite_difference_b200 import hostlogic, solve_host
from finite_difference_b200.batch import SolveBatch
from finite_difference_b200.book import AmericanBook, BarrierBook

def am(n_space, n_time, divs=None, call=False, b76=False, sigma=0.3):
    return hostlogic.build_american_batch([100.0], [95.0 if call else 105.0], [1.0], [sigma], [0.05], [0.03], [call], [n_space],
                                          [n_time], [2], [b76], 4.5, div_taus=[divs or []])

def bar(n_space, n_time, kind="down-and-out"):
    mon = [k / 4.0 for k in range(1, 5)]
    return hostlogic.build_barrier_batch([100.0], [100.0], [1.0], [0.25], [0.05], [0.03], [True], [kind],
                                         [85.0 if "down" in kind else None], [125.0 if "up" in kind else None], [1.0], [mon],
                                         [n_space], [n_time])

divs = [(0.3, 2.0), (0.7, 1.5)]
parts = [am(300, 12, divs, call=True), am(400, 12, divs), am(500, 12, None, b76=True), bar(400, 12), bar(300, 12, "up-and-out"),   # lean
         am(64, 12, divs), am(200, 12, divs, call=True), am(350, 12), bar(129, 12), bar(460, 12),                                   # warp
         am(400, 6, None, sigma=1e-9),                                                                                             # lean hands back (alpha dt ~ 0)
         am(1000, 6, divs), am(2100, 6, divs, call=True), am(3000, 6), am(8192, 6, divs), bar(2133, 6), bar(8193, 6, "up-and-out")]  # cta
batch = SolveBatch.concat(parts)
res = solve_host(batch, dump=True, trace=True)
print("solve_host:", batch.n_solves, "solves, status", res.status.tolist(), "launches", res.n_launches)
assert np.all(np.isfinite(res.price)), res.price
bk = AmericanBook([100.0] * 3, [90.0, 100.0, 110.0], 1.0, [0.2, 0.3, 0.4], 0.05, 0.03, [True, False, True], 300, 16, 2,
                  div_taus=[divs, [], [(0.5, 1.0)]])
out = bk.run_fused("scenario")
print("american book:", out["price_log2"], out["status"])
run = bk.device_run("greeks_log2"); run.upload(); run.launch(); print("device book:", run.results()["price"])
bb = BarrierBook([100.0] * 2, [100.0, 105.0], 1.0, [0.2, 0.3], 0.05, 0.03, "call", ["down-and-out", "up-and-in"], [85.0, None],
                 [None, 130.0], 0.5, [[0.25, 0.5, 0.75]] * 2, 300, 16)
print("barrier book:", bb.run_fused("greeks")["price"], bb.run("greeks")["price"])
print("sanitize cases done")
