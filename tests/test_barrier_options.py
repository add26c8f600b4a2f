"""Engine options of the barrier dataclass beyond the importable reference: Rannacher start-up,
discounted rebate, double-out, Richardson over N and N/2 steps, one-sided vega.  No runnable reference exists for
them (the variants in fd_discrete_barrier_option.py:425-2421 do not parse), so the CUDA path is held to the extended
CPU oracle ``fdo_barrier_solve_ext`` -- parity unpinned, stated as such -- at the same 1e-10 bar."""
import ctypes as C
import math

import numpy as np
import pytest

from finite_difference_b200 import hostlogic
from oracle import oracle

MON = [k / 12.0 for k in range(1, 13)]
BASE = dict(S0=100.0, K=100.0, T=1.0, sigma=0.2, r_disc=0.05, b_carry=0.03, is_call=True)


def _oracle(bt, lo, up, reb, n, m, **opt):
    return oracle.barrier_solve(BASE["S0"], BASE["K"], BASE["T"], BASE["sigma"], BASE["r_disc"], BASE["b_carry"], True, bt,
                                lo, up, reb, MON, n, m, bt != "none", **opt)


def test_extended_oracle_reduces_to_the_pinned_path():
    """With no option set, fdo_barrier_solve_ext is bit-for-bit the pinned fdo_barrier_solve."""
    a = _oracle("down-and-out", 90.0, None, 1.5, 300, 320)
    mon = np.ascontiguousarray(MON)
    p = oracle._bar_struct(100.0, 100.0, 1.0, 0.2, 0.05, 0.03, True, "down-and-out", 90.0, None, 1.5, mon, 300, 320, True)
    V = np.empty(301); s = np.empty(301); flags = np.zeros(321, dtype=np.int32); info = oracle._FdoBarrierInfo()
    rc = oracle.lib().fdo_barrier_solve_ext(C.byref(p), 0, 0, oracle._dptr(V), oracle._dptr(s),
                                            flags.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(info))
    assert rc == 0 and np.array_equal(V, a.V) and info.price == a.price


def test_extended_oracle_options_behave():
    plain = _oracle("down-and-out", 90.0, None, 2.0, 300, 320)
    disc = _oracle("down-and-out", 90.0, None, 2.0, 300, 320, rebate_discounted=True)
    assert disc.price < plain.price                      # a rebate paid later is worth less
    dbl = _oracle("double-out", 85.0, 130.0, 0.0, 300, 320)
    dn = _oracle("down-and-out", 85.0, None, 0.0, 300, 320)
    assert 0.0 < dbl.price < dn.price                    # an extra barrier only removes value
    ran = _oracle("down-and-out", 90.0, None, 0.0, 300, 320, rannacher=2)
    assert abs(ran.price - _oracle("down-and-out", 90.0, None, 0.0, 300, 320).price) < 5e-3


def test_descriptor_options():
    b = hostlogic.build_barrier_batch([100.0, 100.0], [100.0, 100.0], [1.0, 1.0], [0.2, 0.2], [0.05, 0.05], [0.03, 0.03],
                                      [True, False], ["double-out", "down-and-out"], [80.0, 90.0], [130.0, None],
                                      [1.0, 2.0], [MON, MON], [400, 400], [400, 200], rannacher=2, rebate_discounted=True)
    assert list(b.flags) == [0x01 | 0x10 | 0x20, 0x10]
    assert list(b.barrier) == [80.0, 90.0] and list(b.barrier_up) == [130.0, 0.0]
    assert list(b.rannacher) == [2, 2] and list(b.seg_restart) == [1, 1]
    k0 = int(b.mon_step[0])
    assert b.mon_rebate[0] == 1.0 * math.exp(-0.05 * (k0 * (1.0 / 400)))
    first_of_second = int(b.mon_begin[1])
    k1 = int(b.mon_step[first_of_second])
    assert b.mon_rebate[first_of_second] == 2.0 * np.exp(-0.05 * (k1 * (1.0 / 200)))
    # double-out with a missing level: no projection, like BAR:203 / :209 for the single barriers
    c = hostlogic.build_barrier_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.03], [True], ["double-out"], [80.0],
                                      [None], [0.0], [MON], [400], [400])
    assert int(c.flags[0]) == 0x01


def test_strict_dropins_still_reject_double_out():
    from finite_difference_b200 import discrete_barrier_fdm_pricer_cn, fd_discrete_barrier_option
    for mod in (discrete_barrier_fdm_pricer_cn, fd_discrete_barrier_option):
        with pytest.raises(ValueError, match="Unsupported barrier_type"):
            mod.DiscreteBarrierCrankNicolsonLog(100.0, 100.0, 1.0, 0.2, 0.05, 0.05, "call", "double-out")._check_type()


GPU_CASES = [
    ("double_out", "double-out", 85.0, 130.0, 0.0, {}),
    ("double_out_rebate_disc", "double-out", 85.0, 130.0, 1.5, {"rebate_discounted": True}),
    ("doc_rannacher2", "down-and-out", 90.0, None, 0.0, {"rannacher": 2}),
    ("uoc_rannacher3_rebate_disc", "up-and-out", None, 125.0, 2.0, {"rannacher": 3, "rebate_discounted": True}),
    ("vanilla_rannacher2", "none", None, None, 0.0, {"rannacher": 2}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(400, 400), (300, 544), (1200, 300)])
@pytest.mark.parametrize("case", GPU_CASES, ids=[c[0] for c in GPU_CASES])
def test_gpu_options_vs_extended_oracle(case, n, m):
    from finite_difference_b200.barrier_options import DiscreteBarrierCrankNicolsonLogEx as Ex
    _, bt, lo, up, reb, opt = case
    o = _oracle(bt, lo, up, reb, n, m, **opt)
    pr = Ex(100.0, 100.0, 1.0, 0.2, 0.05, 0.03, "call", bt, lower_barrier=lo, upper_barrier=up, rebate=reb,
            monitor_times=MON, N_space=n, N_time=m, rannacher_steps=opt.get("rannacher", 0),
            rebate_discounted=opt.get("rebate_discounted", False))
    V = np.array(pr._solve_grid(apply_KO=bt != "none"))
    scale = np.max(np.abs(o.V))
    assert np.all(np.abs(V - o.V) <= 1e-10 * np.abs(o.V) + 1e-12 * scale)
    assert abs(pr.price() - o.price) <= 1e-10 * abs(o.price)


@pytest.mark.gpu
def test_gpu_richardson_one_sided_vega_and_double_in():
    from finite_difference_b200.barrier_options import DiscreteBarrierCrankNicolsonLogEx as Ex
    kw = dict(lower_barrier=90.0, rebate=0.0, monitor_times=MON, N_space=400, N_time=400, rannacher_steps=2)
    pr = Ex(100.0, 100.0, 1.0, 0.2, 0.05, 0.03, "call", "down-and-out", **kw)
    p_n = _oracle("down-and-out", 90.0, None, 0.0, 400, 400, rannacher=2).price
    p_h = _oracle("down-and-out", 90.0, None, 0.0, 400, 200, rannacher=2).price
    assert abs(pr.price_richardson() - (4.0 * p_n - p_h) / 3.0) <= 1e-10 * p_n
    g = pr.greeks_one_sided(1e-3)
    up = oracle.barrier_solve(100.0, 100.0, 1.0, 0.201, 0.05, 0.03, True, "down-and-out", 90.0, None, 0.0, MON, 400, 400,
                              True, rannacher=2)
    assert abs(g["price"] - p_n) <= 1e-10 * p_n
    assert abs(g["vega"] - (up.price - p_n) / 1e-3) <= 2e-7 * abs(g["vega"])
    # double-in = vanilla PDE - double-out PDE
    di = Ex(100.0, 100.0, 1.0, 0.2, 0.05, 0.03, "call", "double-in", lower_barrier=85.0, upper_barrier=130.0,
            monitor_times=MON, N_space=400, N_time=400)
    van = _oracle("none", 85.0, 130.0, 0.0, 400, 400).price          # same grid: the bounds follow the barrier levels
    dbl = _oracle("double-out", 85.0, 130.0, 0.0, 400, 400).price
    assert abs(di.price() - (van - dbl)) <= 1e-10 * van
