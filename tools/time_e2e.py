"""Dev tool: wall-clock breakdown of the host-facing book paths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.workload(100000)
book = bench.make_book(w, np.arange(100000))
for name, fn in (("run_fused", lambda: book.run_fused("scenario")), ("run", lambda: book.run("scenario"))):
    for rep in range(3):
        t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
        print(f"{name} rep {rep}: wall {dt*1e3:.1f} ms, kernel {out['kernel_ms']:.1f} ms, launches {out['n_launches']}", flush=True)
