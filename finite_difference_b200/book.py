"""Books of options: the batch API above the PDE solves.

The reference prices one option object at a time; every public method of its pricer classes is
a small fan-out of ``_solve_grid`` calls followed by a scalar epilogue:

* ``price_log`` / ``price_log2`` / ``greeks_log2`` -- fd_american_option_pricer.py:659-751,
  fd_american_equity.py:913-1068, fd_american_black76.py:533-625 (1 / 2 / 6 solves),
* ``price`` / ``greeks`` of ``DiscreteBarrierCrankNicolsonLog`` --
  discrete_barrier_fdm_pricer_cn.py:429-537 (3 solves; knock-in by parity),
* the scenario runner calls ``price_log2()`` then ``greeks_log2()`` per CSV row
  (run_american_scenarios.py:144-172): 8 solves per option.

Here the fan-out of a whole book becomes ONE batch of solves (one C-ABI call), and the
epilogues are numpy expressions over the per-solve outputs, written in the reference's
operation order.  Every executed solve is kept (the runner's repeated N / 2N solves are not
de-duplicated unless ``dedupe=True``), so the node-step count equals the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np

from . import hostlogic
from .batch import DeviceBatch, SolveBatch, SolveResults, solve_host

__all__ = ["AmericanBook", "BarrierBook", "DeviceBookRun", "unpack_columns", "norm_cdf", "lapack_delta_gamma"]


def _f64(v, n=None):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
    if a.ndim == 0:
        a = a.reshape(1)
    if n is not None and a.shape[0] != n:
        a = np.ascontiguousarray(np.broadcast_to(a, (n,)))
    return a


def lapack_delta_gamma(stencil_s, stencil_v, spot_idx, n_space):
    """``_local_cubic_delta_gamma`` (fd_american_option_pricer.py:632-653) over a batch: solve
    ``[z^3 z^2 z 1] coef = V`` with ``z = s - S_snapped`` per row; delta = coef[2], gamma = 2 coef[1]."""
    i_near = np.clip(spot_idx, 1, np.asarray(n_space) - 2)
    pos = (spot_idx - i_near + 1).astype(np.int64)
    Sx = np.take_along_axis(stencil_s, pos[:, None], axis=1)
    z = stencil_s - Sx
    M = np.stack([z ** 3, z ** 2, z, np.ones_like(z)], axis=2)
    coef = np.linalg.solve(M, stencil_v[:, :, None])[:, :, 0]
    return coef[:, 2], 2.0 * coef[:, 1]


# --------------------------------------------------------------------------------------
# American books
# --------------------------------------------------------------------------------------

@dataclass
class _Layout:
    """Which solves of the flat batch belong to which role.  The flat batch is role-major:
    solves [k*n_opt, (k+1)*n_opt) hold role k of every option."""
    roles: Sequence[str]
    n_opt: int

    def sl(self, role: str) -> slice:
        k = list(self.roles).index(role)
        return slice(k * self.n_opt, (k + 1) * self.n_opt)


class AmericanBook:
    """A book of American options priced by the theta/CN + Ikonen-Toivanen solver.

    Parameters mirror the scalars the reference constructors derive
    (fd_american_option_pricer.py:117-149): ``T`` = time_to_expiry, ``r`` = discount_rate_nacc,
    ``b`` = carry_rate_nacc.  ``black76=True`` selects ``AmericanFwdFDMPricer`` semantics
    (``spot`` is the forward).  ``richardson_2n_space=True`` reproduces the two equity classes'
    ``price_log2``, which uses ``2*num_space_nodes`` time steps for the fine solve
    (fd_american_option_pricer.py:671, fd_american_equity.py:950); Black-76 uses
    ``2*num_time_steps`` (fd_american_black76.py:545).
    ``div_taus``: per option, [(tau_div, cash)] ascending in tau as ``_div_times_tau`` returns
    (fd_american_option_pricer.py:304-315).
    """

    def __init__(self, spot, strike, T, sigma, r, b, is_call, num_space_nodes=400, num_time_steps=400,
                 rannacher_steps=2, black76=False, s_max_mult=4.5, div_taus: Optional[Sequence] = None):
        self.spot = _f64(spot)
        n = self.n_opt = len(self.spot)
        self.strike, self.T, self.sigma, self.r, self.b = (_f64(v, n) for v in (strike, T, sigma, r, b))
        self.is_call = np.ascontiguousarray(np.broadcast_to(np.asarray(is_call, dtype=bool), (n,)))
        self.n_space = np.maximum(np.broadcast_to(np.asarray(num_space_nodes, dtype=np.int64), (n,)), 3)   # :142
        self.n_time = np.maximum(np.broadcast_to(np.asarray(num_time_steps, dtype=np.int64), (n,)), 4)     # :143
        self.rannacher = np.maximum(np.broadcast_to(np.asarray(rannacher_steps, dtype=np.int64), (n,)), 0)
        self.black76 = np.ascontiguousarray(np.broadcast_to(np.asarray(black76, dtype=bool), (n,)))
        self.s_max_mult = float(s_max_mult)
        self.div_taus = div_taus
        self._div_csr = None

    def div_csr(self):
        """Dividends as CSR arrays ``(begin[n+1] int32, tau float64, cash float64)`` for the fused entries, or
        ``None`` for a dividend-free book.  ``div_taus`` may already be such a triple (vectorised books)."""
        if self._div_csr is None:
            d = self.div_taus
            if d is None:
                return None
            if isinstance(d, tuple) and len(d) == 3 and isinstance(d[0], np.ndarray):
                begin, tau, cash = d
            else:
                counts = np.fromiter((len(x) if (x is not None and not self.black76[i]) else 0 for i, x in enumerate(d)),
                                     dtype=np.int64, count=self.n_opt)
                begin = np.concatenate([[0], np.cumsum(counts)])
                flat = [tc for i, x in enumerate(d) if (x is not None and not self.black76[i]) for tc in x]
                arr = np.asarray(flat, dtype=np.float64).reshape(-1, 2)
                tau, cash = arr[:, 0], arr[:, 1]
            if int(begin[-1]) == 0:
                return None
            self._div_csr = (np.ascontiguousarray(begin, dtype=np.int32), np.ascontiguousarray(tau, dtype=np.float64),
                             np.ascontiguousarray(cash, dtype=np.float64))
        return self._div_csr

    def _div_lists(self):
        """Per-option ``[(tau, cash)]`` lists for the host-side descriptor builder."""
        d = self.div_taus
        if isinstance(d, tuple) and len(d) == 3 and isinstance(d[0], np.ndarray):
            begin, tau, cash = d
            return [list(zip(tau[begin[i]:begin[i + 1]].tolist(), cash[begin[i]:begin[i + 1]].tolist()))
                    for i in range(self.n_opt)]
        return d

    # ---- fan-out -----------------------------------------------------------------------
    def _solves(self, n_time, sigma) -> SolveBatch:
        return hostlogic.build_american_batch(
            self.spot, self.strike, self.T, sigma, self.r, self.b, self.is_call, self.n_space, n_time,
            self.rannacher, self.black76, self.s_max_mult, div_taus=self._div_lists())

    def _fine_steps_price_log2(self):
        return np.where(self.black76, 2 * self.n_time, 2 * self.n_space)

    def fanout(self, what: str, dv_sigma: float = 0.01, use_richardson: bool = True, dedupe: bool = False):
        """Flat batch of every ``_solve_grid`` call behind ``what``:

        ``"price_log"`` 1 solve, ``"price_log2"`` 2, ``"greeks_log2"`` 6 (3 without Richardson),
        ``"scenario"`` = ``price_log2`` then ``greeks_log2`` as run_american_scenarios.py:144-172
        (8 solves; ``dedupe=True`` drops the repeated N solve and, where the two fine step
        counts agree, the repeated 2N solve).
        """
        h = float(dv_sigma)
        s0 = self.sigma
        parts, roles = [], []

        def add(role, n_time, sigma):
            roles.append(role)
            parts.append(self._solves(n_time, sigma))

        if what == "price_log":
            add("N", self.n_time, s0)
        if what in ("price_log2", "scenario"):
            add("p_N", self.n_time, s0)
            if use_richardson:
                add("p_2N", self._fine_steps_price_log2(), s0)
        if what in ("greeks_log2", "scenario"):
            same_n = dedupe and what == "scenario"
            if not same_n:
                add("g_N", self.n_time, s0)
            if use_richardson:
                same_2n = same_n and bool(np.all(self._fine_steps_price_log2() == 2 * self.n_time))
                if not same_2n:
                    add("g_2N", 2 * self.n_time, s0)                         # :700
                add("up_h", self.n_time, s0 + h)                             # :717-723
                add("dn_h", self.n_time, s0 - h)
                add("up_2h", self.n_time, s0 + 2.0 * h)
                add("dn_2h", self.n_time, s0 - 2.0 * h)
            else:
                add("up_h", self.n_time, s0 + h)                             # :727-729
                add("dn_h", self.n_time, s0 - h)
        if not parts:
            raise ValueError(f"unknown fan-out {what!r}")
        return SolveBatch.concat(parts), _Layout(roles, self.n_opt)

    # ---- epilogues -----------------------------------------------------------------------
    def _combine(self, what, lay: _Layout, price, delta, gamma, dv_sigma, use_richardson) -> Dict[str, np.ndarray]:
        h = float(dv_sigma)
        out: Dict[str, np.ndarray] = {}
        roles = list(lay.roles)
        if what == "price_log":
            out["price"] = price[lay.sl("N")]
            out["delta"], out["gamma"] = delta[lay.sl("N")], gamma[lay.sl("N")]
            return out
        if what in ("price_log2", "scenario"):
            pN = price[lay.sl("p_N")]
            out["price_log2"] = (4.0 * price[lay.sl("p_2N")] - pN) / 3.0 if use_richardson else pN   # :670-672
        if what in ("greeks_log2", "scenario"):
            rN = "g_N" if "g_N" in roles else "p_N"
            r2 = "g_2N" if "g_2N" in roles else "p_2N"
            if use_richardson:
                p = (4.0 * price[lay.sl(r2)] - price[lay.sl(rN)]) / 3.0      # :704-706
                d = (4.0 * delta[lay.sl(r2)] - delta[lay.sl(rN)]) / 3.0
                g = (4.0 * gamma[lay.sl(r2)] - gamma[lay.sl(rN)]) / 3.0
                D_h = (price[lay.sl("up_h")] - price[lay.sl("dn_h")]) / (2.0 * h)
                D_2h = (price[lay.sl("up_2h")] - price[lay.sl("dn_2h")]) / (4.0 * h)
                dV = (4.0 * D_h - D_2h) / 3.0                                # :725
            else:
                p, d, g = price[lay.sl(rN)], delta[lay.sl(rN)], gamma[lay.sl(rN)]
                dV = (price[lay.sl("up_h")] - price[lay.sl("dn_h")]) / (2.0 * h)
            s0, S0 = self.sigma, self.spot
            # :737-743 (q = 0) and fd_american_black76.py:614-617; the unsnapped spot is used
            carry = np.where(self.black76, 0.0, (self.b - 0.0) * S0 * d)
            theta = -(0.5 * s0 * s0 * S0 * S0 * g + carry - self.r * p)
            out.update(price=p, delta=d, gamma=g, vega=dV / 100.0, theta=theta)
        return out

    # ---- public ----------------------------------------------------------------------------
    _WHAT = {"price_log": 0, "price_log2": 1, "greeks_log2": 2, "scenario": 3}

    def launch_order(self) -> Optional[np.ndarray]:
        """Order in which the fused entries lay the options out on the device: ascending stiffness ``alpha dt`` of
        the base solve.  The kernels compile their time loop once per carry form (how far a row's influence reaches,
        a function of ``alpha dt``), CTAs start in index order, and an SM's instruction cache holds about one such
        loop -- a book sorted this way keeps neighbouring solves on the same loop (mixed-volatility books: +10..20 %).
        Results are returned in the caller's order.  ``None``: nothing to gain (small or uniform book)."""
        if self.n_opt < 64:
            return None
        S_min, S_max = hostlogic.american_domain_vec(self.spot, self.strike, self.T, self.sigma, self.s_max_mult)
        dx = (np.log(S_max) - np.log(S_min)) / self.n_space
        key = self.sigma * self.sigma / (dx * dx) * (self.T / self.n_time)
        if float(key.max()) <= 1.0001 * float(key.min()):
            return None
        return np.argsort(key, kind="stable")

    def _launch_inputs(self):
        """Per-option input arrays in launch order, built once per book (``launch_order`` and the gathers by it cost
        more than the H2D copies of a 100 000-option book)."""
        if getattr(self, "_launch", None) is None:
            from . import _lib
            order = self.launch_order()

            def take(a):
                return np.ascontiguousarray(a if order is None else a[order])

            arrs = {k: take(np.asarray(getattr(self, k), dtype=np.float64)) for k in ("spot", "strike", "T", "sigma", "r", "b")}
            arrs["flags"] = take((np.where(self.is_call, _lib.FLAG_CALL, 0)
                                  + np.where(self.black76, _lib.FLAG_BLACK76, 0)).astype(np.int32))
            for k in ("n_space", "n_time", "rannacher"):
                arrs[k] = take(np.asarray(getattr(self, k)).astype(np.int32))
            csr = self.div_csr()
            if csr is not None and order is not None:
                begin, tau, cash = csr
                lens = (begin[1:] - begin[:-1])[order].astype(np.int64)
                nb = np.concatenate([[0], np.cumsum(lens)])
                pos = np.repeat(begin[:-1][order].astype(np.int64) - nb[:-1], lens) + np.arange(int(nb[-1]))
                csr = (nb.astype(np.int32), np.ascontiguousarray(tau[pos]), np.ascontiguousarray(cash[pos]))
            ns = arrs["n_space"]
            fl = np.full(self.n_opt, _lib.FLAG_AMERICAN, dtype=np.int32)
            hints = int(_lib.load().fdm_launch_hints(ns.ctypes.data_as(_lib._pi), fl.ctypes.data_as(_lib._pi), self.n_opt,
                                                     None, 0)) & ~4
            self._launch = dict(order=None if order is None else np.ascontiguousarray(order, dtype=np.int32), arrs=arrs,
                                csr=csr, hints=hints)
        return self._launch

    def _book_struct(self, what, dv_sigma, use_richardson, ptr):
        """``(FdmAmericanBook, out_index)`` with pointers produced by ``ptr(name, numpy array)`` (host or device).  The
        options are laid out in launch order (:meth:`launch_order`); ``out_index`` (int32 array or ``None``) goes into
        ``FdmBookOut.out_index`` so that the results come back in the caller's order."""
        from . import _lib
        li = self._launch_inputs()
        n = self.n_opt
        bk = _lib.FdmAmericanBook()
        bk.n_options = n
        for k, v in li["arrs"].items():
            setattr(bk, k, ptr(k, v))
        bk.s_max_mult, bk.dv_sigma = self.s_max_mult, float(dv_sigma)
        bk.what, bk.use_richardson = self._WHAT[what], int(bool(use_richardson))
        if li["csr"] is not None:
            bk.div_begin, bk.div_tau, bk.div_cash = (ptr(k, v) for k, v in zip(("div_begin", "div_tau", "div_cash"), li["csr"]))
            bk.n_div_total = int(li["csr"][0][-1])
        bk.n_space_max = int(self.n_space.max()) if n else 3
        bk.hints = li["hints"]
        return bk, li["order"]

    def solve_steps(self, what: str = "scenario", use_richardson: bool = True) -> int:
        """Metric unit of one pass: sum over the executed solves of num_space_nodes x time steps."""
        ns, nt = self.n_space.astype(np.int64), self.n_time.astype(np.int64)
        fine1 = np.where(self.black76, 2 * nt, 2 * ns)
        total = 0
        if what == "price_log":
            total += nt
        if what in ("price_log2", "scenario"):
            total = total + nt + (fine1 if use_richardson else 0)
        if what in ("greeks_log2", "scenario"):
            total = total + nt + ((2 * nt + 4 * nt) if use_richardson else 2 * nt)
        return int(np.sum(ns * total))

    def run_fused(self, what: str = "scenario", dv_sigma: float = 0.01, use_richardson: bool = True) -> Dict[str, np.ndarray]:
        """One ``fdm_american_book_host`` call: per-option scalars (and the dividend tables) go to the device
        (~56 B per option + 16 B per dividend), the fan-out into solves -- time segments between dividends
        included -- the solves and the Richardson / bump epilogue all run there, and only the per-option results
        come back.  The grid bounds are evaluated with the device's exp/log (<= 1 ulp from libm), so prices agree
        with :meth:`run` to ~1e-14 relative."""
        import ctypes as C
        from . import _lib
        L = _lib.load()
        n = self.n_opt
        keep = []

        def host_ptr(_name, arr):
            keep.append(arr)
            return arr.ctypes.data

        bk, out_index = self._book_struct(what, dv_sigma, use_richardson, host_ptr)
        res = {k: np.empty(n) for k in ("price_log2", "price", "delta", "gamma", "vega", "theta")}
        res.update({k: np.empty(n, dtype=np.int32) for k in ("ex_lo", "ex_hi", "status")})
        out = _lib.FdmBookOut()
        for k, v in res.items():
            setattr(out, k, v.ctypes.data)
        if out_index is not None:
            out.out_index = out_index.ctypes.data
        ns, nl, ms = C.c_int64(0), C.c_int32(0), C.c_float(0.0)
        _lib.check(L.fdm_american_book_host(C.byref(bk), C.byref(out), C.byref(ns), C.byref(nl), C.byref(ms)))
        if what in ("price_log", "price_log2"):
            for k in (("vega", "theta") if what == "price_log" else ("price", "delta", "gamma", "vega", "theta")):
                res.pop(k)
        if what == "greeks_log2":
            res.pop("price_log2")
        n_div = int(bk.n_div_total)
        res.update(node_steps=int(ns.value), n_solves=n * int(L.fdm_book_solves_per_option(bk.what, bk.use_richardson)),
                   kernel_ms=float(ms.value), n_launches=int(nl.value),
                   h2d_bytes=n * (6 * 8 + 4 * 4) + (n_div * 16 + (n + 1) * 4 if n_div else 0) + (n * 4 if out_index is not None else 0),
                   d2h_bytes=n * (6 * 8 + 3 * 4))
        return res

    def device_run(self, what: str = "scenario", dv_sigma: float = 0.01, use_richardson: bool = True, device=None,
                   width: Optional[int] = None) -> "DeviceBookRun":
        """The book resident in HBM: :class:`DeviceBookRun` (``fdm_american_book``, no host copies in ``launch``)."""
        return DeviceBookRun(self, "american", what, dict(dv_sigma=dv_sigma, use_richardson=use_richardson), device, width)

    def run(self, what: str = "scenario", dv_sigma: float = 0.01, use_richardson: bool = True,
            dedupe: bool = False, lapack_greeks: bool = False) -> Dict[str, np.ndarray]:
        """Host-buffer path: one ``fdm_solve_batch_host`` call for the whole fan-out.

        ``lapack_greeks=True`` recomputes delta / gamma on the host from the exported 4-point
        stencils with the reference's own method, a LAPACK solve of the cubic's Vandermonde
        system (fd_american_option_pricer.py:642-653), instead of the in-kernel divided differences
        (the two agree to ~1e-9 relative; the system is ill-conditioned on fine grids)."""
        batch, lay = self.fanout(what, dv_sigma, use_richardson, dedupe)
        res = solve_host(batch)
        delta, gamma = res.delta, res.gamma
        if lapack_greeks:
            delta, gamma = lapack_delta_gamma(res.stencil_s, res.stencil_v, res.idx[:, 0], batch.n_space)
        out = self._combine(what, lay, res.price, delta, gamma, dv_sigma, use_richardson)
        first = lay.sl(lay.roles[0])
        out["ex_lo"], out["ex_hi"] = res.idx[first, 2], res.idx[first, 3]
        out["status"] = np.maximum.reduce([res.status[lay.sl(r)] for r in lay.roles])
        out["node_steps"] = batch.node_steps()
        out["n_solves"] = batch.n_solves
        out["kernel_ms"] = res.kernel_ms
        out["n_launches"] = res.n_launches
        return out

    def price_log(self):
        return self.run("price_log")["price"]

    def price_log2(self, use_richardson: bool = True):
        return self.run("price_log2", use_richardson=use_richardson)["price_log2"]

    def greeks_log2(self, dv_sigma: float = 0.01, use_richardson: bool = True):
        o = self.run("greeks_log2", dv_sigma, use_richardson)
        return {k: o[k] for k in ("price", "delta", "gamma", "vega", "theta")}


# --------------------------------------------------------------------------------------
# Barrier books (DiscreteBarrierCrankNicolsonLog)
# --------------------------------------------------------------------------------------

def norm_cdf(x):
    """discrete_barrier_fdm_pricer_cn.py:16-18."""
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


_OUT_OF = {"down-and-in": "down-and-out", "up-and-in": "up-and-out"}


class BarrierBook:
    """A book of discretely monitored barrier options (``DiscreteBarrierCrankNicolsonLog``,
    discrete_barrier_fdm_pricer_cn.py:25-537).  ``N_space`` / ``N_time`` ``None`` = the
    reference's auto-configuration (:92-118), frozen at the base volatility like the
    reference's first ``configure_grid`` call.
    """

    def __init__(self, S0, K, T, sigma, r_disc, b_carry, option_type, barrier_type, lower_barrier=None,
                 upper_barrier=None, rebate=0.0, monitor_times=None, N_space=None, N_time=None):
        self.S0 = _f64(S0)
        n = self.n_opt = len(self.S0)
        self.K, self.T, self.sigma, self.r_disc, self.b_carry, self.rebate = (
            _f64(v, n) for v in (K, T, sigma, r_disc, b_carry, rebate))

        def strs(v):
            return [v] * n if isinstance(v, str) else list(v)

        def opt(v):
            if v is None:
                return [None] * n
            if np.isscalar(v):
                return [float(v)] * n
            return [None if (x is None or (isinstance(x, float) and math.isnan(x))) else float(x) for x in v]

        self.option_type = [s.lower() for s in strs(option_type)]
        self.barrier_type = [s.lower() for s in strs(barrier_type)]
        for bt in self.barrier_type:
            if bt not in ("none", "down-and-out", "up-and-out", "down-and-in", "up-and-in"):
                raise ValueError(f"Unsupported barrier_type: {bt}")        # :502
        self.lower, self.upper = opt(lower_barrier), opt(upper_barrier)
        if monitor_times is None:
            monitor_times = [[] for _ in range(n)]
        elif len(monitor_times) and np.isscalar(monitor_times[0]):
            monitor_times = [list(monitor_times)] * n
        self.monitor_times = [list(m) if m is not None else [] for m in monitor_times]
        self.N_space, self.N_time = self._configure(N_space, N_time)
        self.is_call = np.array([o == "call" for o in self.option_type])

    def _configure(self, N_space, N_time):
        n = self.n_opt
        if N_space is not None and N_time is not None:
            return (np.broadcast_to(np.asarray(N_space, dtype=np.int64), (n,)).copy(),
                    np.broadcast_to(np.asarray(N_time, dtype=np.int64), (n,)).copy())
        ns = np.empty(n, dtype=np.int64)
        nt = np.empty(n, dtype=np.int64)
        Nsp = np.broadcast_to(np.asarray(-1 if N_space is None else N_space, dtype=np.int64), (n,))
        Nti = np.broadcast_to(np.asarray(-1 if N_time is None else N_time, dtype=np.int64), (n,))
        if n > 4096:
            return self._configure_vec(Nsp, Nti)
        for i in range(n):
            _, _, ns[i], nt[i] = hostlogic.barrier_configure(
                float(self.S0[i]), float(self.K[i]), float(self.T[i]), float(self.sigma[i]), self.lower[i],
                self.upper[i], self.monitor_times[i], None if Nsp[i] < 0 else int(Nsp[i]),
                None if Nti[i] < 0 else int(Nti[i]))
        return ns, nt

    def _configure_vec(self, Nsp, Nti):
        """``configure_grid`` (discrete_barrier_fdm_pricer_cn.py:59-118) over all options at once (large books: numpy's
        log / sqrt can differ from ``math`` by an ulp, which matters only where a ceil() sits exactly on an integer)."""
        n = self.n_opt
        low = np.array([np.nan if v is None else v for v in self.lower], dtype=np.float64)
        up = np.array([np.nan if v is None else v for v in self.upper], dtype=np.float64)
        if np.any(~(self.T > 0)):
            raise ValueError("T must be positive")
        if np.any(~(self.sigma > 0)):
            raise ValueError("sigma must be positive")
        if np.any(~(self.S0 > 0)):
            raise ValueError("S0 must be positive")
        s_low = np.fmin(np.fmin(self.S0, self.K), np.fmin(np.where(low > 0, low, np.nan), np.where(up > 0, up, np.nan)))
        s_high = np.fmax(np.fmax(self.S0, self.K), np.fmax(np.where(low > 0, low, np.nan), np.where(up > 0, up, np.nan)))
        S_min, S_max = np.maximum(1e-8, s_low / 4.0), s_high * 4.0
        bad = S_min >= S_max
        S_min, S_max = np.where(bad, self.S0 / 5.0, S_min), np.where(bad, self.S0 * 5.0, S_max)
        x_range = np.log(S_max) - np.log(S_min)
        dx_target = self.sigma * np.sqrt(self.T) / 12
        dx_target = np.where(dx_target <= 0.0, x_range / 300.0, dx_target)
        ns = np.where(Nsp < 0, np.maximum(np.ceil(x_range / dx_target).astype(np.int64), 300), Nsp)
        dx = x_range / ns
        nt_opt = np.ceil(0.5 * self.sigma * self.sigma * self.T / (0.4 * dx * dx)).astype(np.int64)
        n_valid = np.array([sum(1 for t in (m or []) if 0.0 < t < T) for m, T in zip(self.monitor_times, self.T)], dtype=np.int64)
        nt = np.where(Nti < 0, np.maximum(np.maximum(nt_opt, ns), 10 * (n_valid + 1)), Nti)
        return ns.astype(np.int64), nt.astype(np.int64)

    def _solves(self, sigma, apply_ko: bool = True) -> SolveBatch:
        ko = [_OUT_OF.get(bt, bt) if apply_ko else "none" for bt in self.barrier_type]
        return hostlogic.build_barrier_batch(self.S0, self.K, self.T, sigma, self.r_disc, self.b_carry,
                                             self.is_call, ko, self.lower, self.upper, self.rebate,
                                             self.monitor_times, self.N_space, self.N_time)

    def fanout(self, what: str = "greeks", dv_sigma: float = 1e-3):
        """``"solve"`` = one KO solve per option; ``"greeks"`` = base, sigma+dv, sigma-dv
        (``_pde_price_and_greeks``, discrete_barrier_fdm_pricer_cn.py:429-466)."""
        if what == "solve":
            return self._solves(self.sigma), _Layout(["base"], self.n_opt)
        if what != "greeks":
            raise ValueError(f"unknown fan-out {what!r}")
        parts = [self._solves(self.sigma), self._solves(self.sigma + dv_sigma), self._solves(self.sigma - dv_sigma)]
        return SolveBatch.concat(parts), _Layout(["base", "up", "dn"], self.n_opt)

    def _vanilla_bs(self):
        """``_vanilla_bs_price_and_greeks`` (discrete_barrier_fdm_pricer_cn.py:359-423), per option."""
        n = self.n_opt
        out = {k: np.empty(n) for k in ("price", "delta", "gamma", "vega", "theta")}
        for i in range(n):
            S, K, T, sig, r, b = (float(v[i]) for v in (self.S0, self.K, self.T, self.sigma, self.r_disc, self.b_carry))
            sqrtT = math.sqrt(T)
            d1 = (math.log(S / K) + (b + 0.5 * sig * sig) * T) / (sig * sqrtT)
            d2 = d1 - sig * sqrtT
            df_r = math.exp(-r * T)
            df_b = math.exp((b - r) * T)
            pdf = math.exp(-0.5 * d1 * d1) / math.sqrt(2.0 * math.pi)
            if self.is_call[i]:
                price = S * df_b * norm_cdf(d1) - K * df_r * norm_cdf(d2)
                delta = df_b * norm_cdf(d1)
            else:
                price = K * df_r * norm_cdf(-d2) - S * df_b * norm_cdf(-d1)
                delta = df_b * (norm_cdf(d1) - 1.0)
            gamma = df_b * pdf / (S * sig * sqrtT)
            vega = S * df_b * pdf * sqrtT
            theta = -(0.5 * sig * sig * S * S * gamma + b * S * delta - r * price)
            for k, v in zip(("price", "delta", "gamma", "vega", "theta"), (price, delta, gamma, vega, theta)):
                out[k][i] = v
        return out

    def run(self, what: str = "greeks", dv_sigma: float = 1e-3) -> Dict[str, np.ndarray]:
        """``_cn.py`` semantics (discrete_barrier_fdm_pricer_cn.py:472-537): ``none`` -> closed-form BS,
        knock-out -> PDE, knock-in -> BS minus knock-out, component-wise."""
        batch, lay = self.fanout(what, dv_sigma)
        res = solve_host(batch)
        base = lay.sl("base")
        p, d, g = res.price[base].copy(), res.delta[base].copy(), res.gamma[base].copy()
        out = dict(price=p, delta=d, gamma=g, node_steps=batch.node_steps(), n_solves=batch.n_solves,
                   kernel_ms=res.kernel_ms, n_launches=res.n_launches, status=res.status[base])
        if what == "greeks":
            vega = (res.price[lay.sl("up")] - res.price[lay.sl("dn")]) / (2.0 * dv_sigma)       # :457
            theta = -(0.5 * self.sigma * self.sigma * self.S0 * self.S0 * g
                      + self.b_carry * self.S0 * d - self.r_disc * p)                            # :460-464
            out.update(vega=vega, theta=theta)
        return self._parity_legs(out)

    # ---- fused paths: fan-out (sigma bumps), solves and vega / theta epilogue on the device ------------------
    _WHAT = {"solve": 0, "greeks": 1}

    def _table(self) -> SolveBatch:
        """Per-option descriptors at the base volatility (grid, knock-out flags and levels, monitor steps): none of
        them depends on the volatility bump, so the three solves of ``greeks`` share them."""
        if getattr(self, "_tab", None) is None:
            self._tab = self._solves(self.sigma)
        return self._tab

    def launch_order(self) -> Optional[np.ndarray]:
        """Ascending stiffness ``alpha dt``, as :meth:`AmericanBook.launch_order`."""
        if self.n_opt < 64:
            return None
        t = self._table()
        key = t.sigma * t.sigma / (t.dx * t.dx) * (self.T / self.N_time)
        if float(key.max()) <= 1.0001 * float(key.min()):
            return None
        return np.argsort(key, kind="stable")

    def _launch_inputs(self):
        """Per-option input arrays in launch order, built once per book."""
        if getattr(self, "_launch", None) is None:
            from . import _lib
            t = self._table()
            order = self.launch_order()

            def take(a):
                return np.ascontiguousarray(a if order is None else a[order])

            arrs = {name: take(np.asarray(arr, dtype=np.float64)) for name, arr in
                    (("S0", t.spot), ("K", t.strike), ("T", self.T), ("sigma", t.sigma), ("r_disc", t.r), ("b_carry", t.b),
                     ("x_min", t.x_min), ("dx", t.dx), ("barrier", t.barrier), ("rebate", t.rebate))}
            if t.barrier_up is not None:
                arrs["barrier_up"] = take(t.barrier_up)
            arrs["flags"] = take(t.flags)
            arrs["n_space"] = take(t.n_space)
            arrs["n_time"] = take(np.asarray(self.N_time).astype(np.int32))
            mon = None
            if t.mon_step is not None and len(t.mon_step):
                if order is None:
                    mon = (t.mon_begin, t.mon_step)
                else:                                          # CSR rebuilt in launch order
                    mb, me = t.mon_begin[:-1], t.mon_begin[1:]
                    lens = (me - mb)[order].astype(np.int64)
                    nb = np.concatenate([[0], np.cumsum(lens)])
                    pos = np.repeat(mb[order].astype(np.int64) - nb[:-1], lens) + np.arange(int(nb[-1]))
                    mon = (nb.astype(np.int32), np.ascontiguousarray(t.mon_step[pos]))
            hints = int(_lib.load().fdm_launch_hints(arrs["n_space"].ctypes.data_as(_lib._pi),
                                                     arrs["flags"].ctypes.data_as(_lib._pi), self.n_opt, None, 0)) & ~4
            self._launch = dict(order=None if order is None else np.ascontiguousarray(order, dtype=np.int32), arrs=arrs,
                                mon=mon, hints=hints)
        return self._launch

    def _book_struct(self, what, dv_sigma, ptr):
        """``(FdmBarrierBook, out_index)``; see :meth:`AmericanBook._book_struct`."""
        from . import _lib
        li = self._launch_inputs()
        n = self.n_opt
        bk = _lib.FdmBarrierBook()
        bk.n_options = n
        for k, v in li["arrs"].items():
            setattr(bk, k, ptr(k, v))
        if li["mon"] is not None:
            bk.mon_begin = ptr("mon_begin", li["mon"][0])
            bk.mon_step = ptr("mon_step", li["mon"][1])
            bk.n_mon_total = len(li["mon"][1])
        bk.dv_sigma, bk.what = float(dv_sigma), self._WHAT[what]
        bk.n_space_max = int(self.N_space.max()) if n else 3
        bk.hints = li["hints"]
        return bk, li["order"]

    def solve_steps(self, what: str = "greeks") -> int:
        k = 3 if what == "greeks" else 1
        return int(np.sum(self.N_space.astype(np.int64) * self.N_time.astype(np.int64))) * k

    def _parity_legs(self, out):
        """``none`` -> closed-form BS, knock-in -> BS minus knock-out (discrete_barrier_fdm_pricer_cn.py:489-500, :522-535)."""
        masks = getattr(self, "_parity_masks", None)
        if masks is None:                                       # once per book (0.8 ms on 10 k rows: as much as the H2D + D2H)
            kinds = np.array(self.barrier_type)
            none = kinds == "none"
            masks = self._parity_masks = (none, np.isin(kinds, ("down-and-in", "up-and-in")))
        none, ki = masks
        if none.any() or ki.any():
            bs = self._vanilla_bs()
            for k in ("price", "delta", "gamma", "vega", "theta"):
                if k in out:
                    out[k] = np.where(none, bs[k], np.where(ki, bs[k] - out[k], out[k]))
        return out

    def run_fused(self, what: str = "greeks", dv_sigma: float = 1e-3) -> Dict[str, np.ndarray]:
        """One ``fdm_barrier_book_host`` call: ~100 B per option (+ 4 B per monitor date) to the device; the base
        and bumped solves are expanded there and share the grid and the monitor table; price, delta, gamma, vega
        and theta per option come back.  Same values as :meth:`run` (same descriptors, same kernels, the epilogue
        in the reference's operation order)."""
        import ctypes as C
        from . import _lib
        L = _lib.load()
        n = self.n_opt
        keep = []

        def host_ptr(_name, arr):
            keep.append(arr)
            return arr.ctypes.data

        bk, out_index = self._book_struct(what, dv_sigma, host_ptr)
        res = {k: np.empty(n) for k in ("price", "delta", "gamma", "vega", "theta")}
        res["status"] = np.empty(n, dtype=np.int32)
        out = _lib.FdmBookOut()
        for k, v in res.items():
            setattr(out, k, v.ctypes.data)
        if out_index is not None:
            out.out_index = out_index.ctypes.data
        ns, nl, ms = C.c_int64(0), C.c_int32(0), C.c_float(0.0)
        _lib.check(L.fdm_barrier_book_host(C.byref(bk), C.byref(out), C.byref(ns), C.byref(nl), C.byref(ms)))
        if what == "solve":
            res.pop("vega")
        res = self._parity_legs(res)
        res.update(node_steps=int(ns.value), n_solves=n * (3 if what == "greeks" else 1), kernel_ms=float(ms.value),
                   n_launches=int(nl.value), h2d_bytes=n * (10 * 8 + 3 * 4) + int(bk.n_mon_total) * 4 + (n + 1) * 4,
                   d2h_bytes=n * (5 * 8 + 4))
        return res

    def device_run(self, what: str = "greeks", dv_sigma: float = 1e-3, device=None, width: Optional[int] = None):
        """The book resident in HBM: :class:`DeviceBookRun` (``fdm_barrier_book``); knock-out legs only."""
        return DeviceBookRun(self, "barrier", what, dict(dv_sigma=dv_sigma), device, width)

    def price(self):
        return self.run("greeks")["price"]          # the reference's price() runs all three solves, :486-487

    def greeks(self, dv_sigma: float = 1e-3):
        o = self.run("greeks", dv_sigma)
        return {k: o[k] for k in ("price", "delta", "gamma", "vega", "theta")}


# --------------------------------------------------------------------------------------
# Books resident in HBM
# --------------------------------------------------------------------------------------

class DeviceBookRun:
    """A book whose per-option inputs, workspace and results live in HBM (torch tensors); ``launch`` queues fan-out,
    solves and epilogue on a stream through the device-pointer entries ``fdm_american_book`` / ``fdm_barrier_book``
    with no host copy and no synchronisation.

    The fan-out lays the options out in the book's ``launch_order()`` (ascending stiffness) and the epilogue scatters
    every result back to its caller's row (``FdmBookOut.out_index``), so ``cols`` are in the caller's order.
    The results are views of ONE byte buffer ``packed`` (``width`` rows per column, ``width >= n``): six float64
    columns ``price_log2, price, delta, gamma, vega, theta`` followed by three int32 columns ``ex_lo, ex_hi,
    status`` -- the unit of the multi-GPU result gather (:func:`finite_difference_b200.shard.gather_packed`).
    ``upload()`` refreshes the inputs from pinned host staging buffers (the H2D leg of an end-to-end step).
    """

    F64_COLS = ("price_log2", "price", "delta", "gamma", "vega", "theta")
    I32_COLS = ("ex_lo", "ex_hi", "status")

    def __init__(self, book, kind: str, what: str, opts: dict, device=None, width: Optional[int] = None):
        import torch
        from . import _lib
        if not torch.cuda.is_available():
            raise _lib.FdmError(-4, "no CUDA device; finite_difference_b200 has no CPU fallback")
        self.torch, self.kind, self.book, self.what = torch, kind, book, what
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n = n = book.n_opt
        self.width = width = int(width) if width is not None else n
        if width < n:
            raise ValueError("width must be >= the number of options")
        self._host, self._dev = {}, {}

        def dev_ptr(name, arr):
            h = torch.from_numpy(np.array(arr, order="C", copy=True)).pin_memory()
            d = h.to(self.device, non_blocking=False)
            self._host[name], self._dev[name] = h, d
            return d.data_ptr()

        L = _lib.load()
        if kind == "american":
            self._struct, out_index = book._book_struct(what, opts["dv_sigma"], opts["use_richardson"], dev_ptr)
            need = L.fdm_american_book_workspace(n, int(self._struct.n_div_total), self._struct.what, self._struct.use_richardson)
            self._entry = L.fdm_american_book
            self.node_steps = book.solve_steps(what, opts["use_richardson"])
            self.n_solves = n * int(L.fdm_book_solves_per_option(self._struct.what, self._struct.use_richardson))
        else:
            self._struct, out_index = book._book_struct(what, opts["dv_sigma"], dev_ptr)
            need = L.fdm_barrier_book_workspace(n, self._struct.what)
            self._entry = L.fdm_barrier_book
            self.node_steps = book.solve_steps(what)
            self.n_solves = n * int(L.fdm_barrier_book_solves_per_option(self._struct.what))
        # the fan-out lays the options out in launch order; out_index sends every result back to its caller's row
        self._out_index = None if out_index is None else dev_ptr("out_index", out_index)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self._host.values())
        self.workspace = torch.empty(max(int(need), 256), dtype=torch.uint8, device=self.device)
        self.row_bytes = 8 * len(self.F64_COLS) + 4 * len(self.I32_COLS)
        self.packed = torch.zeros(width * self.row_bytes, dtype=torch.uint8, device=self.device)
        self.cols = {}
        self._out = _lib.FdmBookOut()
        for j, name in enumerate(self.F64_COLS):
            v = self.packed[j * width * 8:(j + 1) * width * 8].view(torch.float64)
            self.cols[name] = v[:n]
            setattr(self._out, name, v.data_ptr())
        base = 8 * len(self.F64_COLS) * width
        for j, name in enumerate(self.I32_COLS):
            v = self.packed[base + j * width * 4: base + (j + 1) * width * 4].view(torch.int32)
            self.cols[name] = v[:n]
            setattr(self._out, name, v.data_ptr())
        if self._out_index is not None:
            self._out.out_index = self._out_index
        self.d2h_bytes = n * self.row_bytes
        self.n_launches = 0

    def upload(self, stream=None):
        """H2D of every per-option input array from its pinned staging buffer (asynchronous on ``stream``)."""
        torch = self.torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            for name, h in self._host.items():
                self._dev[name].copy_(h, non_blocking=True)

    def launch(self, stream=None) -> int:
        import ctypes as C
        from . import _lib
        torch = self.torch
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            nl = C.c_int32(0)
            _lib.check(self._entry(C.byref(self._struct), C.byref(self._out), C.c_void_p(self.workspace.data_ptr()),
                                   C.c_size_t(self.workspace.numel()), C.c_void_p(st.cuda_stream), C.byref(nl)))
        self.n_launches = int(nl.value)
        return self.n_launches

    def results(self) -> Dict[str, np.ndarray]:
        """Synchronise and copy the result columns to the host (the caller's option order)."""
        self.torch.cuda.synchronize(self.device)
        return {k: v.cpu().numpy() for k, v in self.cols.items()}


def unpack_columns(buf: np.ndarray, width: int, n: int) -> Dict[str, np.ndarray]:
    """Columns of one ``DeviceBookRun.packed`` byte buffer (host copy)."""
    out = {}
    nf = len(DeviceBookRun.F64_COLS)
    for j, name in enumerate(DeviceBookRun.F64_COLS):
        out[name] = buf[j * width * 8:(j + 1) * width * 8].view(np.float64)[:n]
    base = 8 * nf * width
    for j, name in enumerate(DeviceBookRun.I32_COLS):
        out[name] = buf[base + j * width * 4: base + (j + 1) * width * 4].view(np.int32)[:n]
    return out
