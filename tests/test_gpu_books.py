"""GPU: the device-side book fan-outs (ABI 3) -- dividend books through ``fdm_american_book_host``, the
device-pointer entries ``fdm_american_book`` / ``fdm_barrier_book``, the barrier book, the "not solved" status and the
sharded run -- against the host-built path, the reference goldens and the oracle.

Bars: a fused book builds its grid bounds with the device's exp/log (<= 1 ulp from libm), so prices agree with the
host-built batch to 1e-12 relative and with the reference to 1e-10; where descriptors are identical (barrier books,
device entry vs host entry) results are bit-identical."""
import ctypes as C

import numpy as np
import pytest

import _golden
from finite_difference_b200 import _lib, hostlogic
from finite_difference_b200.batch import DeviceBatch, SolveBatch
from finite_difference_b200.book import AmericanBook, BarrierBook
from oracle import oracle

pytestmark = pytest.mark.gpu

AM = _golden.american_cases()
DIV = [c for c in AM if c["out"]["div_taus"] and _golden.am_model(c) == 0]


def _rel(a, b, floor=1e-2):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), floor))


def _div_book(n=96, seed=7, n_space=500, n_time=500):
    rng = np.random.default_rng(seed)
    K = rng.uniform(180, 280, n)
    sig = rng.uniform(0.15, 0.45, n)
    r = np.log1p(rng.uniform(0.03, 0.10, n))
    call = rng.uniform(size=n) < 0.5
    T = 1.0
    cash = rng.uniform(2, 9, (n, 2))
    divs = [[(T - 256.0 / 365.0, float(cash[i, 1])), (T - 46.0 / 365.0, float(cash[i, 0]))] if i % 3 else
            ([] if i % 2 else [(0.5, float(cash[i, 0]))]) for i in range(n)]
    return AmericanBook(np.full(n, 229.74), K, T, sig, r, r, call, n_space, n_time, 2, div_taus=divs), divs


def test_there_are_dividend_goldens():
    assert len(DIV) >= 3


@pytest.mark.parametrize("i", range(len(DIV)), ids=[c["id"] for c in DIV])
def test_fused_dividend_book_vs_reference_golden(i):
    """price_log of the reference (golden, dividends included) through the device-side fan-out."""
    c, o = DIV[i], DIV[i]["out"]
    kw = _golden.am_oracle_kwargs(c)
    bk = AmericanBook([kw["spot"]], [kw["strike"]], [kw["T"]], [kw["sigma"]], [kw["r"]], [kw["b"]], [kw["is_call"]],
                      kw["n_space"], kw["n_time"], kw["rannacher"], s_max_mult=kw["s_max_mult"], div_taus=[kw["divs_tau"]])
    res = bk.run_fused("price_log")
    assert res["status"][0] == 0
    assert abs(res["price_log2"][0] - o["price_log"]) <= 1e-10 * abs(o["price_log"])
    assert (int(res["ex_lo"][0]), int(res["ex_hi"][0])) == (o["ex_lo"], o["ex_hi"])


def test_fused_dividend_book_matches_host_built_batch_and_oracle():
    bk, divs = _div_book()
    fused = bk.run_fused("scenario")
    host = bk.run("scenario")
    assert int(fused["status"].max()) == 0 and fused["node_steps"] == host["node_steps"] == bk.solve_steps("scenario")
    for k in ("price_log2", "price", "delta", "gamma", "vega", "theta"):
        tol = 1e-12 if k in ("price_log2", "price") else 1e-7
        assert _rel(fused[k], host[k], 1e-2 if k != "gamma" else 1e-4) <= tol, k
    assert np.array_equal(fused["ex_lo"], host["ex_lo"]) and np.array_equal(fused["ex_hi"], host["ex_hi"])
    # the base solve against the oracle (bit-exact to the reference)
    pl = bk.run_fused("price_log")
    for i in range(0, bk.n_opt, 7):
        o = oracle.american_solve(229.74, float(bk.strike[i]), 1.0, float(bk.sigma[i]), float(bk.r[i]), float(bk.b[i]),
                                  bool(bk.is_call[i]), 0, 500, 500, 2, 4.5, divs[i])
        assert abs(pl["price_log2"][i] - o.price) <= 1e-10 * max(abs(o.price), 1e-2)
        assert (int(pl["ex_lo"][i]), int(pl["ex_hi"][i])) == (o.ex_lo, o.ex_hi)


def test_dividend_csr_triple_equals_lists():
    bk, divs = _div_book(40)
    csr = bk.div_csr()
    bk2 = AmericanBook(bk.spot, bk.strike, bk.T, bk.sigma, bk.r, bk.b, bk.is_call, 500, 500, 2, div_taus=csr)
    a, b = bk.run_fused("price_log2"), bk2.run_fused("price_log2")
    assert np.array_equal(a["price_log2"], b["price_log2"])
    assert np.array_equal(bk2.run("price_log")["price"], bk.run("price_log")["price"])


def test_fused_book_rejects_what_the_reference_rejects():
    bk, _ = _div_book(8)
    bad = AmericanBook(bk.spot, bk.strike, bk.T, bk.sigma, bk.r, bk.b, bk.is_call, 200, 200, 2,
                       div_taus=[[(0.4, 1.0), (0.4, 2.0)]] + [[]] * 7)            # two dividends on one tau: dt = 0
    with pytest.raises(_lib.FdmError):
        bad.run_fused("price_log")
    neg = AmericanBook(bk.spot, bk.strike, bk.T, np.full(8, 0.015), bk.r, bk.b, bk.is_call, 200, 200, 2)
    with pytest.raises(_lib.FdmError):
        neg.run_fused("greeks_log2")                                             # sigma - 2h <= 0
    # the device entry cannot raise: the option is flagged, status FDM_STATUS_NOT_SOLVED, outputs NaN
    sig = np.full(8, 0.3)
    sig[3] = 0.015
    mixed = AmericanBook(bk.spot, bk.strike, bk.T, sig, bk.r, bk.b, bk.is_call, 200, 200, 2)
    run = mixed.device_run("greeks_log2")
    run.launch()
    out = run.results()
    assert out["status"][3] == _lib.STATUS_NOT_SOLVED and np.isnan(out["price"][3]) and np.isnan(out["vega"][3])
    ok = np.arange(8) != 3
    assert np.all(out["status"][ok] == 0) and np.all(np.isfinite(out["price"][ok]))
    bad_run = bad.device_run("price_log")
    bad_run.launch()
    out = bad_run.results()
    assert out["status"][0] == _lib.STATUS_NOT_SOLVED and np.all(out["status"][1:] == 0)


@pytest.mark.parametrize("what", ["price_log", "price_log2", "greeks_log2", "scenario"])
def test_device_book_entry_is_bit_identical_to_the_host_entry(what):
    bk, _ = _div_book(64, seed=11, n_space=400, n_time=400)
    host = bk.run_fused(what)
    run = bk.device_run(what)
    run.upload()
    n_launch = run.launch()
    dev = run.results()
    assert n_launch >= 3 and run.node_steps == host["node_steps"]
    for k, v in host.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(dev[k], v), k
    # a wider packed buffer (the multi-GPU gather layout) changes nothing
    run2 = bk.device_run(what, width=bk.n_opt + 6)
    run2.launch()
    dev2 = run2.results()
    for k in dev:
        assert np.array_equal(dev2[k], dev[k]), k


def _barrier_book(n=60, seed=3, auto=False):
    rng = np.random.default_rng(seed)
    K = rng.uniform(80, 120, n)
    sig = rng.uniform(0.10, 0.50, n)
    r = rng.uniform(0.01, 0.08, n)
    q = rng.uniform(0.0, 0.03, n)
    T = rng.choice([0.5, 1.0, 2.0], n)
    kinds = np.array(["down-and-out", "up-and-out", "down-and-in", "up-and-in", "none"])[np.arange(n) % 5]
    lower = [100.0 * rng.uniform(0.70, 0.95) if "down" in k else None for k in kinds]
    upper = [100.0 * rng.uniform(1.05, 1.40) if "up" in k else None for k in kinds]
    mons = [[k / 12.0 for k in range(1, int(round(12 * t)) + 1)] for t in T]
    call = np.where(np.arange(n) % 2 == 0, "call", "put")
    return BarrierBook(np.full(n, 100.0), K, T, sig, r, r - q, list(call), list(kinds), lower, upper,
                       rng.uniform(0.0, 2.0, n), mons, None if auto else 400, None if auto else 400)


@pytest.mark.parametrize("auto", [False, True], ids=["400x400", "auto-grid"])
def test_barrier_book_fused_is_bit_identical_to_the_host_built_batch(auto):
    bk = _barrier_book(auto=auto)
    for what in ("greeks", "solve"):
        a, b = bk.run_fused(what), bk.run(what)
        assert int(a["status"].max()) == 0 and a["node_steps"] == b["node_steps"] == bk.solve_steps(what)
        for k in ("price", "delta", "gamma", "vega", "theta"):
            if k in b:
                assert np.array_equal(a[k], b[k]), (what, k)
    run = bk.device_run("greeks")
    run.launch()
    dev = run.results()
    ko = np.array(["out" in k for k in bk.barrier_type])
    ref = bk.run("greeks")
    for k in ("price", "delta", "gamma", "vega", "theta"):
        assert np.array_equal(dev[k][ko], ref[k][ko]), k


def test_barrier_book_fused_vs_reference_goldens():
    """The 36 barrier-dataclass goldens (prices, delta, gamma, vega, theta of the reference's greeks())."""
    cases = [c for c in _golden.barrier_cases() if c["module"] == "discrete_barrier_fdm_pricer_cn" and c["out"].get("greeks")]
    if not cases:
        pytest.skip("no _cn greeks goldens")
    bk = BarrierBook([c["S0"] for c in cases], [c["K"] for c in cases], [c["T"] for c in cases], [c["sigma"] for c in cases],
                     [c["r_disc"] for c in cases], [c["b_carry"] for c in cases], [c["option_type"] for c in cases],
                     [c["barrier_type"] for c in cases], [c["lower_barrier"] for c in cases],
                     [c["upper_barrier"] for c in cases], [c["rebate"] for c in cases], [c["monitor_times"] for c in cases],
                     [c["out"]["N_space"] for c in cases], [c["out"]["N_time"] for c in cases])
    res = bk.run_fused("greeks")
    for i, c in enumerate(cases):
        g = c["out"]["greeks"]
        assert abs(res["price"][i] - g["price"]) <= 1e-10 * max(abs(g["price"]), 1e-2), c["id"]
        assert abs(res["delta"][i] - g["delta"]) <= 1e-9 * max(abs(g["delta"]), 1e-4), c["id"]
        assert abs(res["vega"][i] - g["vega"]) <= 1e-6 * max(abs(g["vega"]), 1e-2), c["id"]


def test_skipped_solves_are_visible_in_status():
    """A solve no launched kernel takes (here: an 8000-node grid with n_space_max declared as 513) must not look
    solved: status[] is pre-set to FDM_STATUS_NOT_SOLVED on the device path."""
    b = hostlogic.build_american_batch([100.0, 100.0], [100.0, 95.0], [1.0], [0.2], [0.05], [0.05], [False],
                                       [400, 8000], [50])
    dev = DeviceBatch(b)
    dev._struct.n_space_max = 513
    dev._struct.reserved = 0
    dev.solve()
    res = dev.results()
    assert res.status[0] == 0 and res.status[1] == _lib.STATUS_NOT_SOLVED
    dev = DeviceBatch(b)
    dev.solve()
    assert np.array_equal(dev.results().status, [0, 0])


def test_sharded_run_world_one_equals_run_fused():
    from finite_difference_b200 import shard
    bk, _ = _div_book(33, seed=5, n_space=400, n_time=400)
    ref = bk.run_fused("scenario")
    sr = shard.ShardedBookRun(bk, np.arange(bk.n_opt), bk.n_opt, "scenario")
    out = sr.step()
    for k in ("price_log2", "price", "delta", "gamma", "vega", "theta", "ex_lo", "ex_hi", "status"):
        assert np.array_equal(out[k], ref[k]), k


def test_two_host_threads_on_one_device():
    """The host entries keep one workspace per device behind its own lock: concurrent callers serialise, results
    stay correct."""
    import threading
    bk, _ = _div_book(48, seed=2, n_space=300, n_time=300)
    ref = bk.run_fused("price_log2")["price_log2"]
    got = [None, None]

    def work(j):
        got[j] = bk.run_fused("price_log2")["price_log2"]

    th = [threading.Thread(target=work, args=(j,)) for j in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref)


@pytest.mark.parametrize("n_space", [289, 300, 321, 385, 400, 417, 481, 500, 513])
def test_lean_kernel_dividend_solves_vs_oracle(n_space):
    """Dividend solves in the lean kernel's size classes (the spline jump runs inside fdm_lean_kernel, its scratch
    overlaying the warp's tables): every node of the value vector and the exercise interval at EVERY step against
    the oracle, for 1-3 dividends, calls (Rannacher restart after each jump) and puts."""
    from finite_difference_b200 import solve_host
    rng = np.random.default_rng(n_space)
    cases = []
    for i in range(6):
        nd = 1 + i % 3
        taus = np.sort(rng.uniform(0.08, 0.92, nd))
        divs = [(float(t), float(rng.uniform(0.5, 6.0))) for t in taus]
        cases.append(dict(spot=100.0, strike=float(rng.uniform(85, 115)), T=1.0, sigma=float(rng.uniform(0.15, 0.5)),
                          r=float(rng.uniform(0.01, 0.09)), is_call=bool(i % 2), divs=divs, n_time=int(rng.integers(60, 140))))
    parts = [hostlogic.build_american_batch([c["spot"]], [c["strike"]], [c["T"]], [c["sigma"]], [c["r"]], [c["r"]],
                                            [c["is_call"]], [n_space], [c["n_time"]], [2], [False], 4.5, div_taus=[c["divs"]])
             for c in cases]
    batch = SolveBatch.concat(parts)
    res = solve_host(batch, dump=True, trace=True)
    assert np.all(res.status == 0)
    for i, c in enumerate(cases):
        o = oracle.american_solve(c["spot"], c["strike"], c["T"], c["sigma"], c["r"], c["r"], c["is_call"], 0, n_space,
                                  c["n_time"], 2, 4.5, c["divs"], trace=True)
        V = res.V[i, : n_space + 1]
        tol = 1e-10 * np.abs(o.V) + 1e-12 * np.max(np.abs(o.V))
        assert np.all(np.abs(V - o.V) <= tol), (i, float(np.max(np.abs(V - o.V))))
        assert abs(res.price[i] - o.price) <= 1e-10 * max(abs(o.price), 1e-2)
        steps = o.ex_trace.shape[0]
        assert np.array_equal(res.ex_trace[i, :steps], o.ex_trace), i
