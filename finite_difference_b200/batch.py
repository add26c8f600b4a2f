"""Batches of PDE solves: descriptor arrays (host, numpy) and the two ways to run them.

* :func:`solve_host`   -- HOST buffers through ``fdm_solve_batch_host`` (H2D, kernels, D2H inside
  the call).  This is what the drop-in pricer classes and the scenario runners use.
* :class:`DeviceBatch` -- descriptors resident in HBM as torch tensors, launched through
  ``fdm_solve_batch`` on the current torch stream; results stay on the device.

One "solve" is one ``_solve_grid`` call of the reference
(fd_american_option_pricer.py:554-602, discrete_barrier_fdm_pricer_cn.py:219-302).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (FLAG_AMERICAN, FLAG_BLACK76, FLAG_CALL, FLAG_KO_DOWN, FLAG_KO_UP, FLAG_SNAP,
                   MAX_N_SPACE, WARP_MAX_N_SPACE, FdmBatch, FdmOut)

_F64 = ("x_min", "dx", "sigma", "r", "b", "spot", "strike", "barrier", "rebate")
_I32 = ("n_space", "flags", "rannacher")


@dataclass
class SolveBatch:
    """Structure-of-arrays description of ``n`` PDE solves (see include/fdm_b200.h)."""

    x_min: np.ndarray
    dx: np.ndarray
    sigma: np.ndarray
    r: np.ndarray
    b: np.ndarray
    spot: np.ndarray
    strike: np.ndarray
    n_space: np.ndarray
    flags: np.ndarray
    rannacher: np.ndarray
    seg_begin: np.ndarray          # CSR, n+1
    seg_steps: np.ndarray
    seg_dt: np.ndarray
    seg_tau0: np.ndarray
    seg_restart: np.ndarray
    seg_div: np.ndarray
    barrier: Optional[np.ndarray] = None
    rebate: Optional[np.ndarray] = None
    mon_begin: Optional[np.ndarray] = None   # CSR, n+1
    mon_step: Optional[np.ndarray] = None
    # ABI 2 (optional): carry rate inside the Dirichlet values, upper level of a double-out, per-monitor rebate value
    b_bdry: Optional[np.ndarray] = None
    barrier_up: Optional[np.ndarray] = None
    mon_rebate: Optional[np.ndarray] = None

    def __post_init__(self):
        n = len(self.x_min)
        for name in _F64:
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, np.ascontiguousarray(v, dtype=np.float64))
        for name in _I32:
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.int32))
        self.seg_begin = np.ascontiguousarray(self.seg_begin, dtype=np.int32)
        self.seg_steps = np.ascontiguousarray(self.seg_steps, dtype=np.int32)
        self.seg_dt = np.ascontiguousarray(self.seg_dt, dtype=np.float64)
        self.seg_tau0 = np.ascontiguousarray(self.seg_tau0, dtype=np.float64)
        self.seg_restart = np.ascontiguousarray(self.seg_restart, dtype=np.int32)
        self.seg_div = np.ascontiguousarray(self.seg_div, dtype=np.float64)
        if self.mon_begin is not None:
            self.mon_begin = np.ascontiguousarray(self.mon_begin, dtype=np.int32)
            self.mon_step = np.ascontiguousarray(self.mon_step, dtype=np.int32)
        if (self.barrier is None) != (self.rebate is None):
            raise ValueError("barrier and rebate must be given together")
        for name in ("b_bdry", "barrier_up"):
            v = getattr(self, name)
            if v is not None:
                v = np.ascontiguousarray(v, dtype=np.float64)
                if len(v) != n:
                    raise ValueError(f"{name}: expected {n} entries, got {len(v)}")
                setattr(self, name, v)
        if self.mon_rebate is not None:
            self.mon_rebate = np.ascontiguousarray(self.mon_rebate, dtype=np.float64)
            if self.mon_step is None or len(self.mon_rebate) != len(self.mon_step):
                raise ValueError("mon_rebate needs one entry per mon_step entry")
        for name in _F64 + _I32:
            v = getattr(self, name)
            if v is not None and len(v) != n:
                raise ValueError(f"{name}: expected {n} entries, got {len(v)}")
        if len(self.seg_begin) != n + 1:
            raise ValueError("seg_begin must have n_solves+1 entries")
        if n:
            if int(self.n_space.min()) < 3:
                raise ValueError("n_space must be >= 3")
            if int(self.seg_steps.min(initial=1)) < 1:
                raise ValueError("seg_steps must be >= 1")
            if not np.all(np.diff(self.seg_begin) >= 1):
                raise ValueError("every solve needs at least one time segment")
            if not np.all(self.seg_dt > 0.0):
                # two dividends mapping to the same tau give dt == 0 -> ZeroDivisionError in the
                # reference (fd_american_option_pricer.py:498); rejected here
                raise ValueError("every segment needs dt > 0")

    # ------------------------------------------------------------------
    @property
    def n_solves(self) -> int:
        return int(len(self.x_min))

    @property
    def n_space_max(self) -> int:
        return int(self.n_space.max()) if self.n_solves else 0

    def steps_per_solve(self) -> np.ndarray:
        """Total marched time steps of each solve."""
        csum = np.concatenate([[0], np.cumsum(self.seg_steps, dtype=np.int64)])
        return csum[self.seg_begin[1:]] - csum[self.seg_begin[:-1]]

    def node_steps(self) -> int:
        """Metric unit: sum over solves of num_space_nodes x time steps."""
        return int(np.sum(self.n_space.astype(np.int64) * self.steps_per_solve()))

    def family_mask(self) -> int:
        """Launch hints (FdmBatch.reserved): kernel families and grid-size classes present, fdm_launch_hints()."""
        if self.n_solves == 0:
            return 0
        n_space = np.ascontiguousarray(self.n_space, dtype=np.int32)
        flags = np.ascontiguousarray(self.flags, dtype=np.int32)
        seg_div = np.ascontiguousarray(self.seg_div, dtype=np.float64)
        return int(_lib.load().fdm_launch_hints(
            n_space.ctypes.data_as(_lib._pi), flags.ctypes.data_as(_lib._pi), int(self.n_solves),
            seg_div.ctypes.data_as(_lib._pd), int(seg_div.size)))

    def take(self, sel) -> "SolveBatch":
        """Sub-batch of the selected solves (index array or boolean mask), CSR tables rebuilt."""
        sel = np.asarray(sel)
        if sel.dtype == bool:
            sel = np.nonzero(sel)[0]
        sel = sel.astype(np.int64)

        def csr(begin, *tables):
            lens = (begin[1:] - begin[:-1])[sel]
            nb = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            if len(sel) == 0 or int(lens.sum()) == 0:
                return nb, tuple(t[:0] for t in tables)
            pos = np.repeat(begin[:-1][sel].astype(np.int64) - nb[:-1], lens) + np.arange(int(lens.sum()))
            return nb, tuple(t[pos] for t in tables)

        seg_begin, (ss, sd, st, sr, sv) = csr(self.seg_begin, self.seg_steps, self.seg_dt,
                                              self.seg_tau0, self.seg_restart, self.seg_div)
        mon_begin = mon_step = mon_rebate = None
        if self.mon_begin is not None:
            if self.mon_rebate is not None:
                mon_begin, (mon_step, mon_rebate) = csr(self.mon_begin, self.mon_step, self.mon_rebate)
            else:
                mon_begin, (mon_step,) = csr(self.mon_begin, self.mon_step)
        return SolveBatch(
            x_min=self.x_min[sel], dx=self.dx[sel], sigma=self.sigma[sel], r=self.r[sel], b=self.b[sel],
            spot=self.spot[sel], strike=self.strike[sel], n_space=self.n_space[sel],
            flags=self.flags[sel], rannacher=self.rannacher[sel], seg_begin=seg_begin,
            seg_steps=ss, seg_dt=sd, seg_tau0=st, seg_restart=sr, seg_div=sv,
            barrier=None if self.barrier is None else self.barrier[sel],
            rebate=None if self.rebate is None else self.rebate[sel],
            mon_begin=mon_begin, mon_step=mon_step,
            b_bdry=None if self.b_bdry is None else self.b_bdry[sel],
            barrier_up=None if self.barrier_up is None else self.barrier_up[sel], mon_rebate=mon_rebate)

    @staticmethod
    def concat(parts: Sequence["SolveBatch"]) -> "SolveBatch":
        parts = [p for p in parts if p.n_solves]
        if not parts:
            raise ValueError("nothing to concatenate")
        has_ko = any(p.barrier is not None for p in parts)
        has_mon = any(p.mon_begin is not None for p in parts)

        def cat(name):
            return np.concatenate([getattr(p, name) for p in parts])

        def cat_opt(name):
            return np.concatenate([getattr(p, name) if getattr(p, name) is not None
                                   else np.zeros(p.n_solves) for p in parts])

        def cat_csr(begin_name, n_of):
            out, off = [np.zeros(1, dtype=np.int64)], 0
            for p in parts:
                bg = getattr(p, begin_name)
                if bg is None:
                    bg = np.zeros(p.n_solves + 1, dtype=np.int32)
                out.append(bg[1:].astype(np.int64) + off)
                off += int(bg[-1])
            return np.concatenate(out).astype(np.int32)

        mon_begin = mon_step = mon_rebate = None
        if has_mon:
            mon_begin = cat_csr("mon_begin", None)
            mon_step = np.concatenate([p.mon_step if p.mon_step is not None else np.zeros(0, np.int32)
                                       for p in parts])
            if any(p.mon_rebate is not None for p in parts):
                # parts without per-monitor values repeat their per-solve rebate
                def per_mon(p):
                    if p.mon_step is None:
                        return np.zeros(0)
                    if p.mon_rebate is not None:
                        return p.mon_rebate
                    reb = p.rebate if p.rebate is not None else np.zeros(p.n_solves)
                    return np.repeat(reb, np.diff(p.mon_begin))
                mon_rebate = np.concatenate([per_mon(p) for p in parts])
        b_bdry = None
        if any(p.b_bdry is not None for p in parts):
            b_bdry = np.concatenate([p.b_bdry if p.b_bdry is not None else p.b for p in parts])
        barrier_up = None
        if any(p.barrier_up is not None for p in parts):
            barrier_up = np.concatenate([p.barrier_up if p.barrier_up is not None
                                         else (p.barrier if p.barrier is not None else np.zeros(p.n_solves))
                                         for p in parts])
        return SolveBatch(
            x_min=cat("x_min"), dx=cat("dx"), sigma=cat("sigma"), r=cat("r"), b=cat("b"),
            spot=cat("spot"), strike=cat("strike"), n_space=cat("n_space"), flags=cat("flags"),
            rannacher=cat("rannacher"), seg_begin=cat_csr("seg_begin", None),
            seg_steps=cat("seg_steps"), seg_dt=cat("seg_dt"), seg_tau0=cat("seg_tau0"),
            seg_restart=cat("seg_restart"), seg_div=cat("seg_div"),
            barrier=cat_opt("barrier") if has_ko else None,
            rebate=cat_opt("rebate") if has_ko else None,
            mon_begin=mon_begin, mon_step=mon_step, b_bdry=b_bdry, barrier_up=barrier_up, mon_rebate=mon_rebate)

    # ------------------------------------------------------------------
    def _struct(self, ptr) -> FdmBatch:
        """FdmBatch with pointers produced by ``ptr(name, array)``."""
        s = FdmBatch()
        s.n_solves = self.n_solves
        s.n_space_max = self.n_space_max
        s.n_seg_total = len(self.seg_steps)
        s.n_mon_total = 0 if self.mon_step is None else len(self.mon_step)
        s.reserved = self.family_mask()
        for name in _F64 + _I32 + ("seg_begin", "seg_steps", "seg_dt", "seg_tau0", "seg_restart",
                                   "seg_div", "mon_begin", "mon_step", "b_bdry", "barrier_up", "mon_rebate"):
            arr = getattr(self, name)
            setattr(s, name, None if arr is None else ptr(name, arr))
        s.mon_end = None                      # ABI 3: shared monitor tables are used by the device-side book fan-out only
        return s


@dataclass
class SolveResults:
    price: np.ndarray
    delta: np.ndarray
    gamma: np.ndarray
    stencil_v: np.ndarray     # [n,4]
    stencil_s: np.ndarray     # [n,4]
    idx: np.ndarray           # [n,4]: spot/lo node, strike/centre node, ex_lo, ex_hi
    status: np.ndarray
    V: Optional[np.ndarray] = None          # [n, stride] when dump=True
    ex_trace: Optional[np.ndarray] = None   # [n, steps, 2] when trace=True
    kernel_ms: float = 0.0
    n_launches: int = 0


def _check_sizes(batch: SolveBatch):
    if batch.n_space_max > MAX_N_SPACE:
        raise _lib.FdmError(-3, f"num_space_nodes={batch.n_space_max} exceeds the largest compiled kernel "
                                f"variant ({MAX_N_SPACE})")


def solve_host(batch: SolveBatch, dump: bool = False, trace: bool = False) -> SolveResults:
    """Run every solve of ``batch`` on the current CUDA device from host buffers."""
    L = _lib.load()
    n = batch.n_solves
    res = SolveResults(
        price=np.empty(n), delta=np.empty(n), gamma=np.empty(n), stencil_v=np.empty((n, 4)),
        stencil_s=np.empty((n, 4)), idx=np.empty((n, 4), dtype=np.int32), status=np.empty(n, dtype=np.int32))
    if n == 0:
        return res
    _check_sizes(batch)
    keep = []

    def host_ptr(_name, arr):
        keep.append(arr)
        return arr.ctypes.data

    sb = batch._struct(host_ptr)
    out = FdmOut()
    for name in ("price", "delta", "gamma", "stencil_v", "stencil_s", "idx", "status"):
        setattr(out, name, getattr(res, name).ctypes.data)
    if dump:
        stride = batch.n_space_max + 1
        res.V = np.zeros((n, stride))
        out.v_dump = res.V.ctypes.data
        out.v_dump_stride = stride
    if trace:
        steps = int(batch.steps_per_solve().max())
        res.ex_trace = np.full((n, steps, 2), -1, dtype=np.int32)
        out.ex_trace = res.ex_trace.ctypes.data
        out.ex_trace_stride = 2 * steps
    nl = C.c_int32(0)
    ms = C.c_float(0.0)
    _lib.check(L.fdm_solve_batch_host(C.byref(sb), C.byref(out), C.byref(nl), C.byref(ms)))
    res.kernel_ms = float(ms.value)
    res.n_launches = int(nl.value)
    return res


class DeviceBatch:
    """Descriptors and result buffers resident in HBM (torch tensors); launch with :meth:`solve`."""

    def __init__(self, batch: SolveBatch, device=None, dump: bool = False):
        import torch

        if not torch.cuda.is_available():
            raise _lib.FdmError(-4, "no CUDA device; finite_difference_b200 has no CPU fallback")
        _check_sizes(batch)
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.batch = batch
        self.n = batch.n_solves
        self._t = {}

        def dev_ptr(name, arr):
            t = torch.from_numpy(arr if arr.flags.writeable else arr.copy()).to(self.device, non_blocking=False)
            self._t[name] = t
            return t.data_ptr()

        self._struct = batch._struct(dev_ptr)
        n = self.n
        f64 = dict(dtype=torch.float64, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.price = torch.empty(n, **f64)
        self.delta = torch.empty(n, **f64)
        self.gamma = torch.empty(n, **f64)
        self.stencil_v = torch.empty((n, 4), **f64)
        self.stencil_s = torch.empty((n, 4), **f64)
        self.idx = torch.empty((n, 4), **i32)
        self.status = torch.empty(n, **i32)
        self.V = None
        self._out = FdmOut()
        for name in ("price", "delta", "gamma", "stencil_v", "stencil_s", "idx", "status"):
            setattr(self._out, name, getattr(self, name).data_ptr())
        if dump:
            stride = batch.n_space_max + 1
            self.V = torch.zeros((n, stride), **f64)
            self._out.v_dump = self.V.data_ptr()
            self._out.v_dump_stride = stride
        self.n_launches = 0

    def solve(self, stream=None) -> int:
        """Asynchronous launch on ``stream`` (default: the current torch stream)."""
        torch = self.torch
        L = _lib.load()
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            nl = C.c_int32(0)
            _lib.check(L.fdm_solve_batch(C.byref(self._struct), C.byref(self._out),
                                         C.c_void_p(st.cuda_stream), C.byref(nl)))
        self.n_launches = int(nl.value)
        return self.n_launches

    def results(self) -> SolveResults:
        self.torch.cuda.synchronize(self.device)
        return SolveResults(
            price=self.price.cpu().numpy(), delta=self.delta.cpu().numpy(), gamma=self.gamma.cpu().numpy(),
            stencil_v=self.stencil_v.cpu().numpy(), stencil_s=self.stencil_s.cpu().numpy(),
            idx=self.idx.cpu().numpy(), status=self.status.cpu().numpy(),
            V=None if self.V is None else self.V.cpu().numpy(), n_launches=self.n_launches)


__all__ = ["SolveBatch", "SolveResults", "DeviceBatch", "solve_host", "FLAG_CALL", "FLAG_AMERICAN",
           "FLAG_BLACK76", "FLAG_SNAP", "FLAG_KO_DOWN", "FLAG_KO_UP", "MAX_N_SPACE"]
