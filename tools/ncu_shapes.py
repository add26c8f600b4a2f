"""One warm-up launch and one profiled launch of a named shape (for ncu):
   python tools/ncu_shapes.py {am500|am500div|am400|bar400|cfg4|dbfp|bar8192|am8192|p2} [n_solves]
Batches are sorted by stiffness like the books' launch order (book.launch_order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from finite_difference_b200 import DeviceBatch
from tools.time_variants import american, american_div, american_T, barrier, barrier_T, cfg4
shape = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
batch = {"am500": lambda: american(n or 20000, 500, 500, True), "am500div": lambda: american_div(n or 5000, 500, 500), "am400": lambda: american(n or 20000, 400, 400, False),
         "bar400": lambda: barrier(n or 20000, 400, 400), "cfg4": lambda: cfg4(n or 148, 8192, 1000, divs=False),
         "dbfp": lambda: barrier(n or 1184, 2132, 500),      # 2133-node grids: the 5-warp x 14-slot class
         "bar8192": lambda: barrier_T(n or 148, 8192, 2000, 0.2), "am8192": lambda: american_T(n or 148, 8192, 2000, 0.2),
         "p2": None}[shape]
import numpy as np
if shape == "p2":
    import dataclasses
    from datetime import date, timedelta
    from finite_difference_b200 import rowvar
    val, mat = date(2025, 7, 28), date(2026, 7, 28)
    base = rowvar.march_inputs(100.0, 100.0, val, mat, 0.25, "call", "down-and-out", 90.0, None,
                               [val + timedelta(days=30 * k) for k in range(1, 13)], 0.05, None, 600, 600, 600)
    b = rowvar.RowVarBatch([dataclasses.replace(base, sigma=float(s)) for s in np.linspace(0.12, 0.45, n or 2048)])
    rowvar.solve_rowvar_host(b)
    r = rowvar.solve_rowvar_host(b)
    print("done p2", b.n_solves, float(r.price.sum()))
    sys.exit(0)
batch = batch()
key = batch.sigma ** 2 / batch.dx ** 2 * batch.seg_dt[batch.seg_begin[:-1]]
batch = batch.take(np.argsort(key, kind="stable"))
dev = DeviceBatch(batch)
dev.solve(); torch.cuda.synchronize()
dev.solve(); torch.cuda.synchronize()
print("done", shape, batch.n_solves, float(dev.price.sum()))
