"""Which of the pasted ``_solve_grid`` variants of ``fd_discrete_barrier_option.py:425-2421`` can run at all?
TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).

    python oracle/probe_fragments.py            # writes tests/golden/fragments.json

The file stops parsing at line 1262; lines 425-2421 are ~15 fragments pasted from other classes (SURVEY.md a-15,
f-2).  Every fragment that defines a solver (``_solve_grid`` / ``solve_grid`` / ``_run_backward``) is cut out by its
line range, dedented, compiled with the imports the file lacks (``Tuple``, ``math``), and grafted as ``_solve_grid``
onto the two classes it could belong to:

* ``DiscreteBarrierCrankNicolsonLog`` -- the importable dataclass of the same file (lines 1-423);
* ``DiscreteBarrierFDMPricer`` -- the date/curve-aware class of ``discrete_barrier_fdm_pricer.py`` (ref_loader.load_dbfp),
  whose attribute names (``num_space_nodes``, ``time_to_expiry``, ``carry_rate_nacc`` ...) most fragments use.

Then ``price()`` / ``_solve_grid(True)`` is called on a standard down-and-out call.  The record per (fragment, host
class) holds: compiled?, ran?, the exception text if not, and the value vector's interpolated price if it did.
A fragment counts as a parity target only if it runs on a host class WITHOUT any repair.
"""
from __future__ import annotations

import datetime as dt
import json
import math
import os
import sys
import textwrap
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fragments.json")

# (name, first line, last line) -- 1-based, inclusive; what SURVEY.md a-15 lists plus the other solver bodies
FRAGMENTS = [
    ("run_backward_428", 428, 620),        # _cn_subinterval_log + _run_backward (monitoring-interval marching)
    ("solve_grid_rannacher_912", 912, 1007), ("solve_grid_1009", 1009, 1131), ("solve_grid_1133", 1133, 1236),
    ("solve_grid_1261", 1261, 1322), ("solve_grid_1460", 1460, 1649), ("solve_grid_rannacher_1841", 1841, 1964),
    ("solve_grid_2238", 2238, 2346),
]
PRELUDE = "import math\nfrom typing import Dict, List, Optional, Tuple\nimport numpy as np\n"


def _fragment_source(lines, a, b):
    return textwrap.dedent("".join(lines[a - 1:b]))


def _dataclass_instance():
    mod = ref_loader.load_barrier("fd_discrete_barrier_option")
    return mod.DiscreteBarrierCrankNicolsonLog(S0=100.0, K=100.0, T=1.0, sigma=0.2, r_disc=0.05, b_carry=0.05, option_type="call",
                                               barrier_type="down-and-out", lower_barrier=90.0, rebate=0.0,
                                               monitor_times=[k / 12.0 for k in range(1, 13)], N_space=200, N_time=200)


def _dbfp_instance():
    import pandas as pd
    mod = ref_loader.load_dbfp()
    utils = ref_loader.load_utils()
    df = utils.create_rate_df(0.05)
    df["Date"] = pd.to_datetime(df["Date"], format="%Y/%m/%d").dt.strftime("%Y-%m-%d")
    val, mat = dt.date(2025, 7, 28), dt.date(2026, 1, 28)
    mons = [val + dt.timedelta(days=30 * k) for k in range(1, 6)]
    return mod.DiscreteBarrierFDMPricer(
        spot=100.0, strike=100.0, valuation_date=val, maturity_date=mat, sigma=0.2, option_type="call",
        barrier_type="down-and-out", lower_barrier=90.0, upper_barrier=None, already_in=False, already_hit=False,
        monitor_dates=mons, discount_curve=df, forward_curve=df.copy(), dividend_schedule=[], trade_id=1, direction="long",
        quantity=1, underlying_spot_days=0, option_days=0, option_settlement_days=0, rebate_amount=0.0, rebate_at_hit=False,
        contract_multiplier=1, use_one_sided_greeks_near_barrier=False, num_space_nodes=200, num_time_steps=120,
        grid_type="uniform", rannacher_steps=2, restart_on_monitoring=False, mollify_final=False, mollify_band_nodes=2,
        day_count="ACT/365")


def probe():
    path = os.path.join(ref_loader.REFERENCE_DIR, "fd_discrete_barrier_option.py")
    with open(path, "r", encoding="utf-8") as fh:
        lines = fh.readlines()
    records = []
    for name, a, b in FRAGMENTS:
        rec = {"fragment": name, "lines": [a, b], "hosts": {}}
        src = _fragment_source(lines, a, b)
        ns = {}
        try:
            exec(compile(PRELUDE + src, f"{path}:{a}-{b}", "exec"), ns)
            fns = [v for k, v in ns.items() if callable(v) and getattr(v, "__code__", None) is not None
                   and v.__code__.co_filename.startswith(path) and "self" in v.__code__.co_varnames[:1]]
            rec["compiles"] = True
            rec["functions"] = [f.__name__ for f in fns]
        except Exception as exc:
            rec["compiles"] = False
            rec["error"] = f"{type(exc).__name__}: {exc}"
            records.append(rec)
            continue
        solver = next((f for f in fns if "solve_grid" in f.__name__ or f.__name__ == "_run_backward"), None)
        if solver is None:
            rec["error"] = "no solver function with a self argument in the fragment"
            records.append(rec)
            continue
        for host_name, make in (("DiscreteBarrierCrankNicolsonLog", _dataclass_instance), ("DiscreteBarrierFDMPricer", _dbfp_instance)):
            h = {}
            try:
                obj = make()
                for f in fns:                          # helpers of the fragment ride along
                    setattr(type(obj), f.__name__, f)
                if solver.__name__ == "_run_backward":         # own signature (s_nodes, apply_barrier); needs a built grid
                    obj.configure_grid()
                    if getattr(obj, "s_nodes", None) is None:
                        obj._build_log_grid()
                    V = obj._run_backward(obj.s_nodes, True)[0]
                else:
                    setattr(type(obj), "_solve_grid", solver)
                    V = obj._solve_grid(True)
                h["ran"] = True
                h["n"] = len(V)
                h["finite"] = bool(all(math.isfinite(float(v)) for v in V))
                h["V_mid"] = float(V[len(V) // 2])
            except Exception as exc:
                h["ran"] = False
                h["error"] = f"{type(exc).__name__}: {exc}"
                h["where"] = traceback.format_exc().strip().splitlines()[-3].strip()[:160]
            rec["hosts"][host_name] = h
        records.append(rec)
    return records


if __name__ == "__main__":
    if not ref_loader.reference_available():
        raise SystemExit("needs /root/reference")
    recs = probe()
    with open(OUT, "w") as fh:
        json.dump({"source": "fd_discrete_barrier_option.py", "generated_by": "oracle/probe_fragments.py", "records": recs}, fh, indent=1)
    for r in recs:
        print(r["fragment"], "compiles" if r.get("compiles") else "NO COMPILE: " + r.get("error", ""))
        for hn, h in r.get("hosts", {}).items():
            print("    ", hn, "ran, finite=%s, V_mid=%s" % (h.get("finite"), h.get("V_mid")) if h.get("ran") else "FAILED: " + h["error"])
