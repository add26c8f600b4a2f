#!/usr/bin/env python
"""bench.py -- option-node-steps/s of the batched FD engine on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config cfg3|cfg4|cfg5|cfg2|am400|am500div|dbfp] [--scaling weak|strong] [--options M]

Default workload (BASELINE.json configs[2], SURVEY.md 8d "cfg3"): an American Black-76 scenario sweep in the shape of
run_american_scenarios.py -- per option ``price_log2()`` + ``greeks_log2()`` = 8 PDE solves on a 500-node grid (5000
time steps, 2.5e6 node-steps per option).  ``--scaling weak`` (default): 100 000 options PER GPU; ``--scaling
strong``: the configuration's stated book (cfg3 100 000, cfg4 16 384, cfg5 1 000 000 options) sharded over the N
GPUs.  No communication during the solve, one result collective per step.

One "step" = one pass of the hot path over the rank's shard.  Three timed legs, the same kernels in each:

* ``roofline``: fan-out + solves + epilogue alone (CUDA events on the launching stream around
  ``DeviceBookRun.launch``) -- algorithmic flops (22 per American node-step, 13 per European / barrier node-step,
  SURVEY.md 8d) against the FP64 peak measured live by ``fdm_fp64_peak``.
* ``value``: per-option inputs resident in HBM; launch + the result collective (N > 1), CUDA events, max over ranks.
* ``e2e``: the public API with HOST buffers -- at N = 1 ``AmericanBook.run_fused`` / ``BarrierBook.run_fused``
  (``fdm_american_book_host``: H2D, kernels, D2H inside the call), at N > 1 ``shard.ShardedBookRun.step`` (H2D of
  the shard's inputs from pinned memory, kernels, the same collective as ``value``, D2H of the whole book's results
  on rank 0).

Parity is gated on what is timed: the results left in HBM by the ``value`` leg and the host results of the ``e2e`` leg
are both compared with the CPU oracle on a sample (prices <= 1e-10 relative, exercise intervals identical); at N > 1
rank 0 checks the gathered rows of a sample that spans every rank's shard.

``other_configs`` (N = 1): the remaining BASELINE configurations and kernel families, each with its own clocks,
roofline and parity.  ``cpu_baseline``: the C oracle port on all host cores plus, when ``baseline/_ref`` is staged,
the UNMODIFIED Python reference on a small sample (``python_reference``).
``--impl reference``: the CPU arm alone, same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "option-node-steps/sec (FP64 CN, batched)"
UNIT = "node-steps/s"
FLOPS_AMERICAN = 22.0        # American IT step
FLOPS_LINEAR = 13.0          # European / barrier CN step
SEED = 20251018
F0 = 229.74
T_RUNNER = 31.0 / 365.0


# --------------------------------------------------------------------------------------
# workloads (synthetic inputs of SURVEY.md 8d; seed = 20251018 + config index)
# --------------------------------------------------------------------------------------

def _american_draws(n, seed):
    """Draw order K, sigma, NACA rate, call/put, then dividend cash amounts."""
    rng = np.random.default_rng(seed)
    K = rng.uniform(180.0, 280.0, n)
    sigma = rng.uniform(0.15, 0.45, n)
    naca = rng.uniform(0.03, 0.10, n)
    is_call = rng.uniform(size=n) < 0.5
    cash = rng.uniform(2.0, 9.0, (n, 2))
    return dict(K=K, sigma=sigma, naca=naca, r=np.log1p(naca), is_call=is_call, cash=cash)


def _barrier_draws(n, seed):
    rng = np.random.default_rng(seed)
    K = rng.uniform(80, 120, n)
    sigma = rng.uniform(0.10, 0.50, n)
    r = rng.uniform(0.01, 0.08, n)
    q = rng.uniform(0.0, 0.03, n)
    T = rng.choice([0.5, 1.0, 2.0], n)
    B = 100.0 * rng.uniform(0.70, 0.95, n)
    return dict(K=K, sigma=sigma, r=r, b=r - q, T=T, B=B)


def workload(n_options: int, seed: int = SEED + 3):
    """cfg3 draws (kept under this name: the round-1 records and the golden generator use it)."""
    return _american_draws(n_options, seed)


class Part:
    """One book of a configuration: how to build it for a selection of options and how the oracle prices it."""

    def __init__(self, kind, what, draws, build, oracle, flops, node_steps_per_option):
        self.kind, self.what, self.draws, self.build, self.oracle, self.flops = kind, what, draws, build, oracle, flops
        self.node_steps_per_option = node_steps_per_option


def _am_book(d, sel, n_space, n_time, black76, T, divs):
    from finite_difference_b200.book import AmericanBook
    n = len(sel)
    div_taus = None
    if divs:
        # two cash dividends at valuation + 46 d and + 256 d -> tau = T - t, ascending in tau
        begin = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
        tau = np.tile([T - 256.0 / 365.0, T - 46.0 / 365.0], n)
        cash = d["cash"][sel][:, ::-1].reshape(-1)
        div_taus = (begin, tau, np.ascontiguousarray(cash))
    return AmericanBook(np.full(n, F0), d["K"][sel], T, d["sigma"][sel], d["r"][sel], d["r"][sel], d["is_call"][sel],
                        n_space, n_time, 2, black76=black76, div_taus=div_taus)


def _am_oracle(d, sel, n_space, n_time, black76, T, divs, what, threads):
    """The reference's fan-out through the C oracle (one OpenMP batch) and its epilogue, independent of book.py:
    returns (seconds, node_steps, dict of per-option arrays)."""
    from oracle import oracle
    n = len(sel)
    h = 0.01
    fine1 = 2 * n_time if black76 else 2 * n_space                 # price_log2's fine solve, AM:671 / B76:545
    roles = {"price_log": [(n_time, 0.0)],
             "greeks_log2": [(n_time, 0.0), (2 * n_time, 0.0), (n_time, h), (n_time, -h), (n_time, 2 * h), (n_time, -2 * h)],
             "scenario": [(n_time, 0.0), (fine1, 0.0), (n_time, 0.0), (2 * n_time, 0.0), (n_time, h), (n_time, -h),
                          (n_time, 2 * h), (n_time, -2 * h)]}[what]
    params = []
    for nt, bump in roles:
        for i in sel:
            dv = [(T - 256.0 / 365.0, float(d["cash"][i, 1])), (T - 46.0 / 365.0, float(d["cash"][i, 0]))] if divs else []
            params.append(dict(spot=F0, strike=float(d["K"][i]), T=T, sigma=float(d["sigma"][i]) + bump, r=float(d["r"][i]),
                               b=float(d["r"][i]), is_call=bool(d["is_call"][i]), model=1 if black76 else 0,
                               n_space=n_space, n_time=nt, rannacher=2, divs_tau=dv))
    t0 = time.perf_counter()
    price, lo, hi = oracle.american_batch_prices(params, threads)
    secs = time.perf_counter() - t0
    P = price.reshape(len(roles), n)
    out = {"ex_lo": lo[:n], "ex_hi": hi[:n]}
    if what == "price_log":
        out["price_log2"] = P[0]
    else:
        k0 = 0
        if what == "scenario":
            out["price_log2"] = (4.0 * P[1] - P[0]) / 3.0
            k0 = 2
        out["price"] = (4.0 * P[k0 + 1] - P[k0]) / 3.0
        D_h = (P[k0 + 2] - P[k0 + 3]) / (2.0 * h)
        D_2h = (P[k0 + 4] - P[k0 + 5]) / (4.0 * h)
        out["vega"] = (4.0 * D_h - D_2h) / 3.0 / 100.0
    return secs, n_space * sum(nt for nt, _ in roles) * n, out


def _bar_book(d, sel, what):
    from finite_difference_b200.book import BarrierBook
    n = len(sel)
    mons = [[k / 12.0 for k in range(1, int(round(12 * t)) + 1)] for t in d["T"][sel]]
    return BarrierBook(np.full(n, 100.0), d["K"][sel], d["T"][sel], d["sigma"][sel], d["r"][sel], d["b"][sel], "call",
                       "down-and-out", list(d["B"][sel]), None, 0.0, mons, 400, 400)


def _bar_oracle(d, sel, what, threads):
    from oracle import oracle
    n = len(sel)
    dv = 1e-3
    bumps = (0.0,) if what == "solve" else (0.0, dv, -dv)
    params = []
    for bump in bumps:
        for i in sel:
            T = float(d["T"][i])
            params.append(dict(S0=100.0, K=float(d["K"][i]), T=T, sigma=float(d["sigma"][i]) + bump, r_disc=float(d["r"][i]),
                               b_carry=float(d["b"][i]), is_call=True, barrier_type="down-and-out",
                               lower_barrier=float(d["B"][i]), upper_barrier=None, rebate=0.0,
                               monitor_times=[k / 12.0 for k in range(1, int(round(12 * T)) + 1)], N_space=400, N_time=400))
    t0 = time.perf_counter()
    price, delta, gamma = oracle.barrier_batch_prices(params, threads)
    secs = time.perf_counter() - t0
    out = {"price": price[:n], "delta": delta[:n], "gamma": gamma[:n]}
    if what == "greeks":
        out["vega"] = (price[n:2 * n] - price[2 * n:3 * n]) / (2.0 * dv)
    return secs, 400 * 400 * len(bumps) * n, out


def american_part(seed, n_space, n_time, black76, T, divs, what):
    spo = {"price_log": n_time, "greeks_log2": 7 * n_time,
           "scenario": 8 * n_time + (2 * n_time if black76 else 2 * n_space)}[what] * n_space
    return Part("american", what, lambda n: _american_draws(n, seed),
                lambda d, sel: _am_book(d, sel, n_space, n_time, black76, T, divs),
                lambda d, sel, th: _am_oracle(d, sel, n_space, n_time, black76, T, divs, what, th), FLOPS_AMERICAN, spo)


def barrier_part(seed, what):
    return Part("barrier", what, lambda n: _barrier_draws(n, seed), lambda d, sel: _bar_book(d, sel, what),
                lambda d, sel, th: _bar_oracle(d, sel, what, th), FLOPS_LINEAR, 400 * 400 * (3 if what == "greeks" else 1))


class Config:
    def __init__(self, name, desc, parts, weak_options, strong_options, cpu_sample, shares=None):
        self.name, self.desc, self.parts = name, desc, parts
        self.weak_options, self.strong_options, self.cpu_sample = weak_options, strong_options, cpu_sample
        self.shares = shares or [1.0 / len(parts)] * len(parts)       # split of the option count over the parts


def make_config(name: str, steps_cfg4: int = 10000) -> Config:
    if name == "cfg3":
        return Config("cfg3", "cfg3: American Black-76 scenario sweep (run_american_scenarios shape: price_log2 + greeks_log2 = "
                      "8 solves/option, 500 nodes x 5000 steps/option, T=31/365)",
                      [american_part(SEED + 3, 500, 500, True, T_RUNNER, False, "scenario")], 100_000, 100_000, 48)
    if name == "cfg4":
        return Config("cfg4", f"cfg4: high-resolution grids, American equity with two cash dividends, 8192 nodes x {steps_cfg4} "
                      "steps, price_log (1 solve/option), T=1", [american_part(SEED + 4, 8192, steps_cfg4, False, 1.0, True, "price_log")],
                      2048, 16384, 1)
    if name == "cfg5":
        return Config("cfg5", "cfg5: Greeks by bump-and-reprice on a mixed book, 50 % American equity (greeks_log2: 5 grids "
                      "sigma, +-h, +-2h at N plus the 2N Richardson grid = 6 solves, 400 x 400, T=1) + 50 % discrete "
                      "down-and-out barrier calls (greeks: 3 grids sigma, +-1e-3, 400 x 400, monthly monitoring)",
                      [american_part(SEED + 5, 400, 400, False, 1.0, False, "greeks_log2"), barrier_part(SEED + 5, "greeks")],
                      125_000, 1_000_000, 8)
    if name == "cfg2":
        return Config("cfg2", "cfg2: 10k discrete down-and-out barrier calls, CN with monthly monitoring, 400 x 400, greeks() = "
                      "3 solves/option", [barrier_part(SEED + 2, "greeks")], 10_000, 10_000, 16)
    if name == "cfg2price":
        return Config("cfg2price", "cfg2: 10k discrete down-and-out barrier calls, CN with monthly monitoring, 400 x 400, one "
                      "knock-out solve/option", [barrier_part(SEED + 2, "solve")], 10_000, 10_000, 16)
    if name == "am400":
        return Config("am400", "headline: 100k-option American CN batch, one solve each, 400 nodes x 400 steps (equity, no "
                      "dividends, T=1)", [american_part(SEED + 1, 400, 400, False, 1.0, False, "price_log")], 100_000, 100_000, 64)
    if name == "am500div":
        return Config("am500div", "runner default with dividends: American equity, 500 x 500, two cash dividends, price_log2 + "
                      "greeks_log2 = 8 solves/option, T=1", [american_part(SEED + 6, 500, 500, False, 1.0, True, "scenario")],
                      12_500, 12_500, 8)
    raise SystemExit(f"unknown --config {name}")


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def split_counts(n, shares):
    c = [int(round(n * s)) for s in shares]
    c[-1] = n - sum(c[:-1])
    return c


def compare(ref: dict, got: dict) -> dict:
    """Prices relative to max(|ref|, 1e-2) (absolute floor 1e-12 for tiny prices), exercise intervals
    exact; sensitivities reported, bounded loosely (differences of prices divided by the bump)."""
    out = {"n_options": int(len(next(iter(ref.values()))))}
    ok = True
    for k in ("price_log2", "price"):
        if k in ref and k in got:
            e = float(np.max(np.abs(got[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-2)))
            out["max_rel_err_" + k] = e
            ok &= e <= 1e-10
    for k, tol in (("delta", 1e-8), ("gamma", 1e-7), ("vega", 1e-6)):
        if k in ref and k in got:
            e = float(np.max(np.abs(got[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-3)))
            out["max_rel_err_" + k] = e
            ok &= e <= tol
    if "ex_lo" in ref and "ex_lo" in got:
        same = bool(np.array_equal(got["ex_lo"], ref["ex_lo"]) and np.array_equal(got["ex_hi"], ref["ex_hi"]))
        out["exercise_intervals_identical"] = same
        ok &= same
    out["tolerance"] = 1e-10
    out["ok"] = bool(ok)
    return out


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms (NVML) while running."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# CPU arm
# --------------------------------------------------------------------------------------

def cpu_arm(cfg: Config, n_sample: int, threads: int):
    """The configuration's fan-out through the C oracle port on `threads` host cores, over the first `n_sample`
    options of every part.  Returns (seconds, node_steps, [per-part reference dicts], [per-part selections])."""
    secs = ns = 0.0
    refs, sels = [], []
    for part, share in zip(cfg.parts, cfg.shares):
        k = max(1, int(round(n_sample * share)))
        d = part.draws(k)
        sel = np.arange(k)
        s, n, ref = part.oracle(d, sel, threads)
        secs += s
        ns += n
        refs.append(ref)
        sels.append(sel)
    return secs, ns, refs, sels


def python_reference_leg(threads: int, n_sample: int = 16):
    """cfg3 sample through the UNMODIFIED Python reference staged under baseline/_ref (None if not staged)."""
    try:
        from oracle import pyref
        if not pyref.available():
            return None, None
        n = max(1, min(n_sample, 2 * threads))
        w = workload(n)
        out = pyref.time_cfg3(F0, w["K"], w["sigma"], w["naca"], w["is_call"], 500, 500, threads)
        info = {"value": out["node_steps"] / out["seconds"], "unit": UNIT, "cores": int(out["procs"]),
                "per_core": out["node_steps"] / out["cpu_seconds"],
                "sample": f"first {n} options of cfg3 x 8 solves through the unmodified fd_american_black76.AmericanFwdFDMPricer "
                          f"(price_log2 + greeks_log2) from baseline/_ref, one process per core ({out['seconds']:.1f} s wall)"}
        return info, out
    except Exception as exc:  # the reference arm must never take the bench down
        return {"unavailable": f"{type(exc).__name__}: {exc}"}, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    cfg = make_config(args.config, args.cfg4_steps)
    threads = host_threads()
    n_sample = args.cpu_options or max(cfg.cpu_sample, cfg.cpu_sample * threads)
    for _ in range(args.warmup):
        cpu_arm(cfg, max(len(cfg.parts), min(n_sample, threads)), threads)
    total_t = total_ns = 0.0
    for _ in range(args.steps):
        s, n, _, _ = cpu_arm(cfg, n_sample, threads)
        total_t += s
        total_ns += n
    value = total_ns / total_t
    n_opt = options_of(cfg, args, world)
    sample = (f"sample: {n_sample} options per step (of {n_opt} in the configuration), the reference's fan-out through the C "
              "oracle port of its Python arithmetic, OpenMP over solves")
    pyinfo, _ = python_reference_leg(threads) if cfg.name == "cfg3" else (None, None)
    cpu = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    if pyinfo:
        cpu["python_reference"] = pyinfo
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args, world, n_opt, sample=n_sample),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def options_of(cfg, args, world):
    if args.options:
        return args.options * (world if args.scaling == "weak" else 1)
    return cfg.weak_options * world if args.scaling == "weak" else cfg.strong_options


def config_dict(cfg, args, world, n_total, collective="none", sample=None):
    d = {"workload": cfg.desc, "options_total": int(n_total),
         "options_per_gpu": int(-(-n_total // world)),
         "node_steps_per_option": [int(p.node_steps_per_option) for p in cfg.parts],
         "sharding": (f"by option over {world} GPUs, no communication during the solve; result collective per step: "
                      f"{collective} of the packed per-option results (value and e2e legs alike)") if world > 1 else "single GPU",
         "l2": "flushed (256 MiB memset) between timed steps; per-option inputs are read once per solve"}
    if sample is not None:
        d["cpu_sample_options_per_step"] = int(sample)
    return d


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

class ShardRuns:
    """This rank's books of one configuration, resident in HBM, with their result gathers."""

    def __init__(self, cfg, n_total, rank, world, dev, dst=0):
        from finite_difference_b200 import shard
        self.cfg, self.n_total, self.world, self.rank = cfg, n_total, world, rank
        counts = split_counts(n_total, cfg.shares)
        self.parts = []
        for part, cnt in zip(cfg.parts, counts):
            d = part.draws(cnt)
            mine = shard.partition(np.full(cnt, float(part.node_steps_per_option)), world)[rank]
            book = part.build(d, mine)
            sr = shard.ShardedBookRun(book, mine, cnt, part.what, dst=dst, device=dev)
            self.parts.append(dict(part=part, draws=d, mine=mine, book=book, sr=sr, count=cnt))
        self.node_steps = sum(p["sr"].run.node_steps for p in self.parts)
        self.flops = sum(p["sr"].run.node_steps * p["part"].flops for p in self.parts)
        self.n_solves = sum(p["sr"].run.n_solves for p in self.parts)
        self.h2d = sum(p["sr"].run.h2d_bytes for p in self.parts)
        self.collective = self.parts[0]["sr"].gather.collective

    def launch(self):
        return sum(p["sr"].run.launch() for p in self.parts)

    def collect(self, d2h):
        outs = []
        for p in self.parts:
            g = p["sr"].gather
            if self.world > 1 or d2h:
                outs.append(g(p["sr"].run.packed) if d2h else g.exchange(p["sr"].run.packed))
            else:
                outs.append(None)
        return outs

    def e2e_step(self):
        return [p["sr"].step() for p in self.parts]


def time_config(cfg, args, rank, world, local, n_total, steps, warmup, with_cpu, threads):
    """All legs of one configuration.  Returns (line dict or None on ranks > 0, ok)."""
    import torch
    import torch.distributed as dist
    from finite_difference_b200 import _lib
    from finite_difference_b200.book import DeviceBookRun

    dev = torch.device("cuda", local)
    runs = ShardRuns(cfg, n_total, rank, world, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    peak_before, _ = _lib.fp64_peak_tflops()
    for _ in range(max(warmup, 3)):
        runs.launch()
        runs.collect(d2h=False)
    torch.cuda.synchronize(dev)

    # ---- value leg: K steps, CUDA events on the launching stream ------------------------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(steps)] for _ in range(3)]
    launches = 0
    barrier()
    torch.cuda.synchronize(dev)
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xff)                     # L2 flush, outside the per-step events
            ev[0][i].record(stream)
            launches += runs.launch()
            ev[1][i].record(stream)
            runs.collect(d2h=False)
            ev[2][i].record(stream)
        torch.cuda.synchronize(dev)
        t_wall = time.perf_counter() - t_wall0
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in zip(ev[0], ev[2]))
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(ev[0], ev[1]))
    peak_after, _ = _lib.fp64_peak_tflops()
    dev_results = [p["sr"].run.results() for p in runs.parts]      # what the timed leg left in HBM (caller's option order)

    t = torch.tensor([step_ms, kern_ms, t_wall * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(runs.node_steps), float(runs.n_solves)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    step_ms_max, kern_ms_max, wall_ms_max = (float(x) for x in t.tolist())
    total_node_steps, total_solves = (float(x) for x in tot.tolist())
    value = total_node_steps * steps / (step_ms_max * 1e-3)

    # ---- end to end through the public host API -------------------------------------------------------------
    def e2e_once():
        if world == 1:
            return [p["book"].run_fused(p["part"].what) for p in runs.parts]
        return runs.e2e_step()

    e2e_out = e2e_once()                               # warm-up (NCCL sets up its channels on first use)
    barrier()
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_out = e2e_once()
    torch.cuda.synchronize(dev)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_node_steps * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    if world == 1:
        h2d = sum(o["h2d_bytes"] for o in e2e_out)
        d2h = sum(o["d2h_bytes"] for o in e2e_out)
        e2e_path = ("AmericanBook / BarrierBook .run_fused -> fdm_american_book_host / fdm_barrier_book_host: per-option scalars "
                    "H2D, fan-out + solves + Richardson / bump epilogue on the device, per-option results D2H")
    else:
        h2d = runs.h2d
        d2h = sum(p["sr"].gather.d2h_bytes for p in runs.parts) if rank == 0 else 0
        e2e_path = ("shard.ShardedBookRun.step: H2D of the shard's per-option inputs from pinned memory, fdm_american_book / "
                    f"fdm_barrier_book on the device, {runs.collective} of the packed results, rows put in book order on the "
                    "device, D2H of the whole book's results on rank 0")

    # ---- parity of what was timed ---------------------------------------------------------------------------
    parity, cpu, ok = None, None, True
    if with_cpu and rank == 0:
        from oracle import oracle
        oracle.build()
        parity = {"value_leg": [], "e2e_leg": []}
        n_sample = args.cpu_options or max(cfg.cpu_sample, cfg.cpu_sample * threads)
        secs = ns_cpu = 0.0
        for p, dres, eres in zip(runs.parts, dev_results, e2e_out):
            k = min(p["count"], max(1, int(round(n_sample * p["count"] / n_total))))
            sel = np.arange(k)                                   # the first k options of the part: they span every rank
            s, n, ref = p["part"].oracle(p["draws"], sel, threads)
            secs += s
            ns_cpu += n
            # value leg: this rank's rows of the sample, as left in HBM by the timed steps
            loc = np.nonzero(p["mine"] < k)[0]
            gsel = p["mine"][loc]
            parity["value_leg"].append(compare({a: b[gsel] for a, b in ref.items()}, {a: b[loc] for a, b in dres.items()}))
            # e2e leg: rank 0 holds the whole book in book order (at N > 1 the gathered rows of every rank)
            if rank == 0 and eres is not None:
                parity["e2e_leg"].append(compare(ref, {a: np.asarray(b)[:k] for a, b in eres.items() if isinstance(b, np.ndarray)}))
                if world == 1:                                   # same kernels, same inputs: the two legs agree bit for bit
                    same = all(np.array_equal(eres[a], dres[a]) for a in dres if isinstance(eres.get(a), np.ndarray))
                    parity["e2e_equals_value_leg_bitwise"] = bool(same and parity.get("e2e_equals_value_leg_bitwise", True))
        ok = all(x["ok"] for x in parity["value_leg"] + parity["e2e_leg"])
        parity["ok"] = ok
        cpu = {"value": ns_cpu / secs, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_sample} options of the workload ({secs:.1f} s wall), every solve of their fan-out through the C "
                         "oracle port of the reference's Python arithmetic, OpenMP over solves"}
    flag = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    ok = float(flag.item()) == 0.0
    if rank != 0:
        return None, ok

    # ---- roofline (rank 0's kernels) ---------------------------------------------------------------------------
    peak = max(peak_before, peak_after)
    kern_s = kern_ms * 1e-3 / steps
    achieved = runs.flops / kern_s / 1e12
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if cfg.name in tj.get("configs", {}):
            traffic = tj["configs"][cfg.name]["dram_bytes_per_solve"] * runs.n_solves
        elif cfg.name == "cfg3":
            traffic = tj["dram_bytes_per_solve"] * runs.n_solves
    except Exception:
        pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = runs.h2d + sum(p["sr"].run.d2h_bytes for p in runs.parts) + runs.n_solves * (104 + 52)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "fdm_fp64_peak register-resident DFMA micro-benchmark, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry; profiles/fp64_peak.json keeps a record with clocks)",
                "flops_per_node_step": runs.flops / runs.node_steps,
                "kernel_ms_per_step": kern_ms / steps,
                "launches_per_step": launches / steps,
                "what": "fan-out kernel + solve kernels + epilogue kernel of one step (the solve kernels are > 99.9 % of it)",
                "hbm": {"algorithmic_bytes_per_step": int(alg_bytes), "achieved_gbs": alg_bytes / kern_s / 1e9,
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "frac": alg_bytes / kern_s / 1e9 / hbm_peak}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": step_ms_max / steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args, world, n_total, runs.collective),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": float(e2e_ms.item()) / e2e_steps, "path": e2e_path},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "wall_ms_timed_region": wall_ms_max,
            "solves_per_step": int(total_solves),
            "fp64_peak": {"tflops_before": peak_before, "tflops_after": peak_after}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
        line["parity"] = parity
    del runs, flush
    torch.cuda.empty_cache()
    return line, ok


def dbfp_config(args, local, threads):
    """The barrier scenario runner's own grid (run_config_scenarios shape): DiscreteBarrierFDMPricer rows, 2133 space
    nodes x 500 steps, 2 solves per row (sigma, sigma + 1e-4): fdm_cta_kernel, 5-warp class."""
    import datetime as dt
    import pandas as pd
    import torch
    from finite_difference_b200 import DeviceBatch, _lib, hostlogic, run_config_scenarios as rcs
    from oracle import oracle
    n = 2048
    rng = np.random.default_rng(SEED + 6)
    val, mat = dt.date(2025, 7, 28), dt.date(2025, 8, 28)
    mons = [d for d in (val + dt.timedelta(days=i) for i in range(32)) if d.weekday() < 5]
    down = rng.uniform(size=n) < 0.5
    frame = pd.DataFrame(dict(
        scenario_name=[f"s{i}" for i in range(n)], S0=229.74, K=rng.uniform(180, 280, n), sigma=rng.uniform(0.15, 0.45, n),
        rate=rng.choice(np.round(np.linspace(0.03, 0.10, 15), 4), n), barrier_type=np.where(down, "down-and-out", "up-and-out"),
        upper_barrier=np.where(down, np.nan, rng.uniform(240, 330, n)), lower_barrier=np.where(down, rng.uniform(150, 220, n), np.nan),
        FA_price=np.nan, FA_delta=np.nan, FA_gamma=np.nan, FA_vega=np.nan))
    base = dict(valuation=val, maturity=mat, monitor_dates=mons, opt_type="put", num_time_steps=500)
    cache = {}
    pricers = [rcs._build_pricer(r["S0"], r["K"], r["sigma"], r["rate"], r["barrier_type"], rcs._opt(r["upper_barrier"]),
                                 rcs._opt(r["lower_barrier"]), _curve_cache=cache, **base) for _, r in frame.iterrows()]
    a_ = []
    for p in pricers:
        a_.append(p._solve_args(True))
        a_.append(p._solve_args(True, sigma=p.sigma + 1e-4))
    batch = hostlogic.build_dbfp_batch(**{k: [a[k] for a in a_] for k in a_[0]})
    dev = DeviceBatch(batch)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dev.solve()
    torch.cuda.synchronize()
    steps = 3
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(steps)] for _ in range(2)]
    launches = 0
    peak, _ = _lib.fp64_peak_tflops()
    with ClockSampler(local) as clocks:
        for i in range(steps):
            flush.fill_(i)
            ev[0][i].record()
            launches += dev.solve()
            ev[1][i].record()
        torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in zip(*ev)) / steps
    res = dev.results()
    worst = 0.0
    for i in range(0, 32, 2):
        a = a_[i]
        o = oracle.dbfp_solve(a["S_min"], a["S_max"], a["S0_eff"], a["S0_eff"], a["K"], a["T"], a["sigma"], a["r"], a["b"], a["q"],
                              a["is_call"], a["ko_type"], a["lower_barrier"], a["upper_barrier"], a["rebate_amount"],
                              a["rebate_at_hit"], a["monitor_times"], a["num_space_nodes"], a["num_time_steps"],
                              a["rannacher_steps"])
        worst = max(worst, abs(res.price[i] - o.price) / max(abs(o.price), 1e-2))
    rcs.run_scenarios_frame(frame.iloc[:64], base)            # warm-up
    t0 = time.perf_counter()
    rcs.run_scenarios_frame(frame, base)
    e2e_s = time.perf_counter() - t0
    ns = batch.node_steps()
    return {"workload": f"dbfp: barrier scenario sweep (run_config_scenarios shape, DiscreteBarrierFDMPricer): {n} rows x 2 solves, "
                        "2133 space nodes x 500 steps, weekday monitoring, Rannacher 2",
            "value": ns / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "steps": steps, "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": {"bound": "fp64", "flops_per_node_step": FLOPS_LINEAR, "achieved": ns * FLOPS_LINEAR / ms / 1e9, "peak": peak,
                         "unit": "TFLOP/s", "frac": ns * FLOPS_LINEAR / ms / 1e9 / peak},
            "e2e": {"value": ns / e2e_s, "unit": UNIT, "seconds": e2e_s,
                    "path": "run_config_scenarios.run_scenarios_frame -> BarrierScenarioBook (numpy over rows) -> "
                            "fdm_solve_batch_host -> result frame"},
            "parity": {"n_solves": 16, "max_rel_err_price": worst, "tolerance": 1e-10, "ok": bool(worst <= 1e-10)}}


def p2_config(args, local, threads):
    """the row-varying S-space march of DiscreteBarrierFDMPricer2 (fdm_rowvar_kernel), 601 nodes x 600
    steps, every step implicit (the regime in which the reference prices), down-and-out calls with monthly monitoring."""
    import dataclasses
    from datetime import date, timedelta
    import torch
    from finite_difference_b200 import _lib, rowvar
    from oracle import oracle
    rng = np.random.default_rng(SEED + 7)
    val, mat = date(2025, 7, 28), date(2026, 7, 28)
    mons = [val + timedelta(days=30 * k) for k in range(1, 13)]
    contracts, marches = [], []
    n_vol = 256                                                 # 32 contracts x 256 volatilities = 8192 marches (4.6 waves of 12 warps per SM)
    for i in range(32):
        kw = dict(spot=100.0, strike=float(rng.uniform(85, 115)), valuation_date=val, maturity_date=mat, volatility=0.25,
                  option_type="call", barrier_type="down-and-out", lower_barrier=float(rng.uniform(70, 92)), monitoring_dates=mons,
                  flat_rate_nacc=float(rng.uniform(0.01, 0.08)), num_space_nodes=600, num_time_steps=600, rannacher_steps=600)
        contracts.append(kw)
        base = rowvar.march_inputs(**kw)
        marches += [dataclasses.replace(base, sigma=float(sg)) for sg in np.linspace(0.12, 0.45, n_vol)]
    batch = rowvar.RowVarBatch(marches)
    peak, _ = _lib.fp64_peak_tflops()
    rowvar.solve_rowvar_host(batch)
    with ClockSampler(local) as clocks:
        runs = [rowvar.solve_rowvar_host(batch) for _ in range(3)]
    ms = min(r.kernel_ms for r in runs)
    t0 = time.perf_counter()
    res = rowvar.solve_rowvar_host(batch)
    e2e_s = time.perf_counter() - t0
    worst = 0.0
    for i in range(0, 32, 4):                                   # volatility 0.12 + the construction volatility's own grid
        kw = dict(contracts[i])
        prep = oracle.pricer2_prepare(**kw)
        prep["sigma"] = float(np.linspace(0.12, 0.45, n_vol)[0])
        V = oracle.pricer2_march(prep, True)
        ref = oracle.p2_interp_linear(prep["S_eff"], [float(x) for x in prep["s_nodes"]], [float(x) for x in V])
        worst = max(worst, abs(res.price[i * n_vol] - ref) / max(abs(ref), 1e-2))
    ns = batch.node_steps()
    return {"workload": "p2: row-varying S-space march (DiscreteBarrierFDMPricer2._solve_pde_backward), 8192 down-and-out calls, "
                        "601 nodes x 600 implicit steps, monthly monitoring",
            "value": ns / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "steps": 3, "gpu_launches": res.n_launches,
            "clocks": clocks.summary(),
            "roofline": {"bound": "fp64", "flops_per_node_step": FLOPS_LINEAR, "achieved": ns * FLOPS_LINEAR / ms / 1e9, "peak": peak,
                         "unit": "TFLOP/s", "frac": ns * FLOPS_LINEAR / ms / 1e9 / peak,
                         "note": "13 flops per node-step: right-hand side 5 + Thomas 8; the reference also rebuilds the matrix every step"},
            "e2e": {"value": ns / e2e_s, "unit": UNIT, "seconds": e2e_s,
                    "path": "rowvar.solve_rowvar_host -> fdm_rowvar_solve_host (grids and payoffs H2D, march, prices D2H)"},
            "parity": {"n_solves": 8, "max_rel_err_price": worst, "tolerance": 1e-10, "ok": bool(worst <= 1e-10)}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from finite_difference_b200 import _lib, shard

    rank, world, local = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (finite_difference_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    _lib.load()
    threads = host_threads()
    cfg = make_config(args.config, args.cfg4_steps)
    n_total = options_of(cfg, args, world)
    # at N > 1 the oracle runs on rank 0 only (the other ranks wait in the flag all-reduce)
    line, ok = time_config(cfg, args, rank, world, local, n_total, args.steps, args.warmup,
                           with_cpu=not args.no_cpu, threads=threads)
    if rank == 0 and line is not None:
        if not ok:
            emit({"error": "parity gate failed", "parity": line.get("parity"), "config": line["config"]})
            return 1
        if world == 1 and not args.no_cpu and cfg.name == "cfg3":
            pyinfo, pyout = python_reference_leg(threads)
            if pyinfo:
                line["cpu_baseline"]["python_reference"] = pyinfo
            if pyout is not None:
                # the same options through the GPU path, against the reference's own Python outputs
                n = len(pyout["price"])
                w = workload(n)
                from finite_difference_b200.book import AmericanBook
                bk = AmericanBook(np.full(n, F0), w["K"], pyout["T"], w["sigma"], pyout["r"], pyout["r"], w["is_call"], 500, 500, 2,
                                  black76=True)
                got = bk.run_fused("scenario")
                pp = compare({k: pyout[k] for k in ("price_log2", "price", "delta", "gamma", "vega")}, got)
                line["parity"]["vs_python_reference"] = pp
                if not pp["ok"]:
                    emit({"error": "parity gate failed (python reference)", "parity": pp})
                    return 1
        if world == 1 and not args.no_others and cfg.name == "cfg3":
            others = {}
            sub = argparse.Namespace(**vars(args))
            sub.options, sub.scaling, sub.cpu_options = 0, "weak", 0
            for name, opts, kw in (("cfg2", 10_000, {}), ("cfg2price", 10_000, {}), ("am400", 100_000, {}),
                                   ("am500div", 12_500, {}), ("cfg4", 592, {"steps": 2500}), ("cfg5", 125_000, {})):
                c2 = make_config(name, kw.get("steps", 1000))
                sub.options = opts
                try:
                    l2, ok2 = time_config(c2, sub, 0, 1, local, opts, 3, 3, True, threads)
                    keep = {k: l2[k] for k in ("value", "unit", "ms_per_step", "steps", "clocks", "e2e", "gpu_launches", "roofline",
                                               "parity", "solves_per_step")}
                    keep["workload"] = c2.desc
                    keep["options"] = opts
                    others[name] = keep
                    ok &= ok2
                except Exception as exc:
                    others[name] = {"error": f"{type(exc).__name__}: {exc}"}
                    ok = False
            try:
                others["dbfp"] = dbfp_config(args, local, threads)
                ok &= others["dbfp"]["parity"]["ok"]
            except Exception as exc:
                others["dbfp"] = {"error": f"{type(exc).__name__}: {exc}"}
                ok = False
            try:
                others["p2"] = p2_config(args, local, threads)
                ok &= others["p2"]["parity"]["ok"]
            except Exception as exc:
                others["p2"] = {"error": f"{type(exc).__name__}: {exc}"}
                ok = False
            line["other_configs"] = others
            if not ok:
                emit({"error": "parity gate failed in other_configs", "other_configs": others})
                return 1
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


def main():
    # the JSON line must be the only thing on stdout (NCCL prints its version banner there): everything
    # else goes to stderr, the line is written to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", help="cfg3 (default) | cfg4 | cfg5 | cfg2 | cfg2price | am400 | am500div")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --options per GPU; strong: the configuration's stated book sharded over the GPUs")
    ap.add_argument("--options", type=int, default=0, help="options per GPU (weak) or in total (strong); 0 = the configuration's")
    ap.add_argument("--cfg4-steps", type=int, default=10000, help="time steps of --config cfg4 (10 000 in BASELINE.json)")
    ap.add_argument("--cpu-options", type=int, default=0, help="options in the CPU sample (default: about 20 core-seconds of CPU work)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-others", action="store_true", help="skip other_configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
