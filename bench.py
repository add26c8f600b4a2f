#!/usr/bin/env python
"""bench.py -- option-node-steps/s of the batched FD engine on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--options M]

Workload (BASELINE.json configs[2], SURVEY.md 8d "cfg3"): an American Black-76 scenario sweep in
the shape of run_american_scenarios.py -- per option ``price_log2()`` + ``greeks_log2()`` = 8 PDE
solves on a 500-node grid (5000 time steps, 2.5e6 node-steps per option), 100 000 options per
GPU (weak scaling: every rank prices its own shard of an N x 100k-option book, no communication
during the solve, one result all-gather per step).

One "step" = one pass of the hot path over the rank's shard.

* ``value``: device-resident descriptors (already in HBM), CUDA events on the launching stream.
* ``e2e``: the same sweep through the public host API (``AmericanBook.run_fused`` ->
  ``fdm_american_book_host``): numpy option arrays in host memory -> H2D of the per-option
  scalars -> fan-out, solves and Richardson / vega epilogue on the device -> D2H of the results.
* ``roofline``: FP64 (DFMA pipe) -- algorithmic flops (22 per American node-step)
  over the kernel's event time against the FP64 peak measured live by ``fdm_fp64_peak``.
* ``cpu_baseline``: the C oracle (a port of the reference's Python arithmetic) on the host
  cores over a bounded sample of the same workload; the same call is the parity gate.
* ``--impl reference``: the CPU arm alone (oracle port, all host threads), same metric/config.
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
FLOPS_PER_NODE_STEP_AMERICAN = 22.0
SEED = 20251018 + 3
N_SPACE = N_TIME = 500
T_EXP = 31.0 / 365.0
F0 = 229.74
SOLVES_PER_OPTION = 8


def workload(n_options: int, seed: int = SEED):
    """cfg3 draw order: K, sigma, NACA rate, call/put."""
    rng = np.random.default_rng(seed)
    K = rng.uniform(180.0, 280.0, n_options)
    sigma = rng.uniform(0.15, 0.45, n_options)
    naca = rng.uniform(0.03, 0.10, n_options)
    is_call = rng.uniform(size=n_options) < 0.5
    r = np.log1p(naca)          # flat NACA curve -> NACC forward rate, fd_american_black76.py:186-207
    return dict(K=K, sigma=sigma, r=r, is_call=is_call)


def make_book(w, sel):
    from finite_difference_b200.book import AmericanBook
    return AmericanBook(np.full(len(sel), F0), w["K"][sel], T_EXP, w["sigma"][sel], w["r"][sel], w["r"][sel],
                        w["is_call"][sel], N_SPACE, N_TIME, 2, black76=True)


def config_dict(n_opt_per_gpu, world):
    return {"workload": "cfg3: American Black-76 scenario sweep (run_american_scenarios shape: price_log2 + "
                        "greeks_log2 = 8 solves/option, 500 nodes x 5000 steps/option, T=31/365)",
            "options_per_gpu": int(n_opt_per_gpu), "solves_per_option": SOLVES_PER_OPTION,
            "node_steps_per_option": 2_500_000,
            "sharding": f"by option over {world} GPU(s), result all-gather per step" if world > 1 else "single GPU",
            "l2": "flushed (256 MiB memset) between timed steps; descriptors are read once per solve"}


# --------------------------------------------------------------------------------------
# CPU arm (oracle port)
# --------------------------------------------------------------------------------------

def oracle_scenario(w, sel, n_threads=0):
    """The 8 solves of each selected option through the C oracle.  Returns (seconds, node_steps,
    base prices, ex_lo, ex_hi)."""
    from oracle import oracle
    n = len(sel)
    K, sig, r, call = w["K"][sel], w["sigma"][sel], w["r"][sel], w["is_call"][sel]
    h = 0.01
    sig_all = np.concatenate([sig, sig, sig, sig, sig + h, sig - h, sig + 2 * h, sig - 2 * h])
    nt_all = np.concatenate([np.full(n, N_TIME), np.full(n, 2 * N_TIME), np.full(n, N_TIME), np.full(n, 2 * N_TIME),
                             np.full(4 * n, N_TIME)])
    tile = lambda a: np.tile(a, SOLVES_PER_OPTION)
    t0 = time.perf_counter()
    price, lo, hi = oracle.american_batch_prices_arrays(F0, tile(K), T_EXP, sig_all, tile(r), tile(r), tile(call), 1,
                                                        N_SPACE, nt_all, 2, 4.5, n_threads)
    dt = time.perf_counter() - t0
    return dt, int(np.sum(nt_all) * N_SPACE), price, lo, hi


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    threads = host_threads()
    n_sample = args.cpu_options or max(48, 48 * threads)
    w = workload(n_sample)
    sel = np.arange(n_sample)
    for _ in range(args.warmup):
        oracle_scenario(w, sel[: max(threads, 8)], threads)
    total_t, total_ns = 0.0, 0
    for _ in range(args.steps):
        dt, ns, *_ = oracle_scenario(w, sel, threads)
        total_t += dt
        total_ns += ns
    value = total_ns / total_t
    sample = f"{n_sample} options x 8 solves per step (of {args.options} per GPU), C oracle port of the reference, OpenMP"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.options, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------

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
# GPU arm
# --------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from finite_difference_b200 import DeviceBatch, _lib, shard

    rank, world, local = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (finite_difference_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()

    # ---- the book and this rank's shard --------------------------------------------------
    n_total = args.options * world
    w = workload(n_total)
    cost = np.full(n_total, 2.5e6)
    mine = shard.partition(cost, world)[rank]
    book = make_book(w, mine)
    batch, lay = book.fanout("scenario")
    node_steps = batch.node_steps()
    dbatch = DeviceBatch(batch, device=dev)
    # per-solve outputs the epilogue needs: price, delta, gamma (+ exercise interval)
    gather_in = gather_out = None
    if world > 1:
        gather_in = torch.empty((batch.n_solves, 3), dtype=torch.float64, device=dev)
        gather_out = torch.empty((world * batch.n_solves, 3), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        n = dbatch.solve(stream)
        if world > 1:
            torch.stack((dbatch.price, dbatch.delta, dbatch.gamma), dim=1, out=gather_in)
            dist.all_gather_into_tensor(gather_out, gather_in)
        return n

    def barrier():
        if world > 1:
            dist.barrier()

    peak_before, _ = _lib.fp64_peak_tflops()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, CUDA events on the launching stream -------------------------
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches = 0
    barrier()
    torch.cuda.synchronize(dev)
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xff)                     # L2 flush, outside the per-step events
            e0[i].record(stream)
            launches += dbatch.solve(stream)
            k0[i].record(stream)
            if world > 1:
                torch.stack((dbatch.price, dbatch.delta, dbatch.gamma), dim=1, out=gather_in)
                dist.all_gather_into_tensor(gather_out, gather_in)
            e1[i].record(stream)
        torch.cuda.synchronize(dev)
        t_wall = time.perf_counter() - t_wall0
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(e0, k0))
    peak_after, _ = _lib.fp64_peak_tflops()

    t = torch.tensor([step_ms, kern_ms, t_wall * 1e3], dtype=torch.float64, device=dev)
    ns = torch.tensor([float(node_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ns, op=dist.ReduceOp.SUM)
    step_ms_max, kern_ms_max, wall_ms_max = (float(x) for x in t.tolist())
    total_node_steps = float(ns.item())
    value = total_node_steps * args.steps / (step_ms_max * 1e-3)

    # ---- end to end through the public host API ----------------------------------------------
    res = book.run_fused("scenario")                 # warm-up (including the result gather: NCCL sets up its
    if world > 1:                                    # point-to-point channels on first use)
        shard.gather_results({k: res[k] for k in ("price_log2", "price", "delta", "gamma", "vega", "theta")},
                             mine, n_total, device=dev, dst=0)
    barrier()
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t_it = time.perf_counter()
        out = book.run_fused("scenario")
        if os.environ.get("FDM_BENCH_DEBUG"):
            print(f"[debug] e2e iteration wall {1e3 * (time.perf_counter() - t_it):.1f} ms kernel {out['kernel_ms']:.1f} ms",
                  file=sys.stderr, flush=True)
        if world > 1:
            # the sweep's one collective: per-option results to rank 0 (the other ranks keep their shards)
            shard.gather_results({k: out[k] for k in ("price_log2", "price", "delta", "gamma", "vega", "theta")},
                                 mine, n_total, device=dev, dst=0)
    torch.cuda.synchronize(dev)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_node_steps * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    h2d, d2h = out["h2d_bytes"], out["d2h_bytes"]                      # per-option scalars in, per-option results out
    desc_bytes = sum(getattr(batch, f).nbytes for f in
                     ("x_min", "dx", "sigma", "r", "b", "spot", "strike", "n_space", "flags", "rannacher", "seg_begin",
                      "seg_steps", "seg_dt", "seg_tau0", "seg_restart", "seg_div"))
    res_bytes = batch.n_solves * (3 * 8 + 2 * 32 + 16 + 4)

    if rank != 0:
        return 0

    # ---- roofline (rank 0's kernel) --------------------------------------------------------------
    peak = max(peak_before, peak_after)
    kern_s = kern_ms * 1e-3 / args.steps
    achieved = node_steps * FLOPS_PER_NODE_STEP_AMERICAN / kern_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_solve"] * batch.n_solves
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "fdm_fp64_peak register-resident DFMA micro-benchmark, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_node_step": FLOPS_PER_NODE_STEP_AMERICAN, "kernel": "fdm_lean_kernel<16,true,false>",
                "kernel_ms_per_launch": kern_ms / args.steps,
                "launches_per_step": "fdm_lean_kernel<16,true,false> (all 800k solves) + fdm_warp_kernel<16,true,false> in "
                                     "only_rest mode (takes what the lean kernel hands back: nothing here, exits at once)",
                "hbm": {"algorithmic_bytes_per_launch": desc_bytes + res_bytes,
                        "achieved_gbs": (desc_bytes + res_bytes) / kern_s / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "frac": (desc_bytes + res_bytes) / kern_s / 1e9 / hbm_peak}}

    # ---- CPU baseline + parity gate (rank 0, N = 1 only) ----------------------------------------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle
        oracle.build()
        threads = host_threads()
        n_sample = args.cpu_options or max(48, 48 * threads)
        n_sample = min(n_sample, len(mine))
        sel = np.arange(n_sample)
        dt, ns_cpu, p_cpu, lo_cpu, hi_cpu = oracle_scenario(w, mine[sel], threads)
        cpu = {"value": ns_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_sample} options x 8 solves of the workload ({dt:.1f} s wall), C oracle port of "
                         "the reference's Python arithmetic, OpenMP"}
        # parity: base-N solve prices and exercise intervals of the sampled options, GPU (host API) vs oracle
        full = book.fanout("scenario")[0]
        from finite_difference_b200 import solve_host
        gres = solve_host(full.take(np.concatenate([k * len(mine) + sel for k in range(SOLVES_PER_OPTION)])))
        rel = np.abs(gres.price - p_cpu) / np.maximum(np.abs(p_cpu), 1e-2)
        same_ex = bool(np.array_equal(gres.idx[:, 2], lo_cpu) and np.array_equal(gres.idx[:, 3], hi_cpu))
        parity = {"n_solves": int(len(p_cpu)), "max_rel_err_price": float(rel.max()), "tolerance": 1e-10,
                  "exercise_intervals_identical": same_ex}
        if not (rel.max() <= 1e-10 and same_ex):
            emit({"error": "parity gate failed", "parity": parity})
            return 1

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.options, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": float(e2e_ms.item()) / e2e_steps,
                    "path": "AmericanBook.run_fused('scenario') -> fdm_american_book_host: per-option scalars H2D, "
                            "fan-out + solves + Richardson/bump epilogue on the device, per-option results D2H"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "wall_ms_timed_region": wall_ms_max,
            "options_total": int(n_total), "solves_per_step": int(batch.n_solves * world),
            "sanity": {"price_log2_mean": float(np.mean(res["price_log2"])), "status_max": int(res["status"].max())}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
        line["parity"] = parity
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--options", type=int, default=100_000, help="options per GPU (8 solves each)")
    ap.add_argument("--cpu-options", type=int, default=0, help="options in the CPU sample (default 48 x threads: about 20 core-seconds of CPU work)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
