"""Dev tool: GPU-vs-oracle error on 8192-node grids as a function of stiffness alpha*dt."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finite_difference_b200 import hostlogic, solve_host
from oracle import oracle

for sigma in (0.15, 0.30, 0.45):
    for m in (100, 400, 1600, 10000):
        for divs in ([], [(0.3, 3.0), (0.88, 9.0)]):
            kw = dict(spot=229.74, strike=200.0, T=1.0, sigma=sigma, r=0.09, b=0.09, is_call=False, model=0,
                      n_space=8192, n_time=m, rannacher=2, s_max_mult=4.5, divs_tau=divs)
            b = hostlogic.build_american_batch([kw["spot"]], [kw["strike"]], [1.0], [sigma], [0.09], [0.09], [False],
                                               [8192], [m], [2], [False], 4.5, div_taus=[divs])
            res = solve_host(b, dump=True)
            o = oracle.american_solve(**kw)
            V = res.V[0, :8193]
            scale = np.max(np.abs(o.V))
            rel = np.max(np.abs(V - o.V) / (np.abs(o.V) + 1e-2 * scale))
            dx = float(b.dx[0]); adt = 0.5 * sigma * sigma / dx / dx * (1.0 / m)
            print(f"sigma {sigma} M {m:6d} divs {len(divs)} alpha*dt {adt:9.1f}  price rel {abs(res.price[0]-o.price)/o.price:.2e} "
                  f"max node err/(|V|+1%scale) {rel:.2e}  ex {res.idx[0,2:].tolist()} vs {[o.ex_lo,o.ex_hi]} ms {res.kernel_ms:.1f}", flush=True)
