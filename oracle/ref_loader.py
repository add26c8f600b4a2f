"""Loads the UNMODIFIED reference pricers from a read-only checkout.  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/gen_golden.py`` (in the build container, where ``/root/reference`` exists) to
produce the golden vectors under ``tests/golden/``.  Nothing that runs on the GPU box imports
this module: ``/root/reference`` does not exist there.

Two shims are needed:

* ``workalendar`` is not installed.  The American files only call
  ``SouthAfrica().add_working_days(date, n)`` (fd_american_option_pricer.py:103-115) and every
  reference driver passes lag 0, so a stub that returns the date unchanged for lag 0 is exact.
* ``discrete_barrier_fdm_pricer_cn.py`` and ``fd_discrete_barrier_option.py`` do not parse past
  lines 537 / 423; the importable prefixes hold the complete dataclass.
* ``discrete_barrier_fdm_pricer.py`` (class ``DiscreteBarrierFDMPricer``, what ``run_config_scenarios.py``
  drives) does not parse either: a stray triple quote at line 746 comments out lines 747-881.  Its own lines
  1-745 + 883-1027 form a complete class body (constructor, grid, ``_solve_grid``, ``_interp_price``, Black-76
  vanilla leg, ``_pde_price_and_greeks3``, ``price_log2``, ``greeks_log2``).  The constructor calls
  ``self._build_stock_price_grid()`` (:167), which no line of the file defines; its result is only used for
  ``self.grid_spacing`` (:168), which nothing reads.  ``load_dbfp`` therefore execs that splice UNMODIFIED and
  adds one attribute, ``_build_stock_price_grid = lambda self: [0.0, 0.0]``.  This is the smallest repair that
  lets the reference run; it is recorded here and in DESIGN.md as such.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("FDM_REFERENCE_DIR", "/root/reference")

_BARRIER_PREFIX = {
    "discrete_barrier_fdm_pricer_cn": 537,
    "fd_discrete_barrier_option": 423,
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "fd_american_option_pricer.py"))


def _install_workalendar_stub() -> None:
    if "workalendar.africa" in sys.modules:
        return
    pkg = types.ModuleType("workalendar")
    africa = types.ModuleType("workalendar.africa")

    class SouthAfrica:  # noqa: D401 - stub
        def add_working_days(self, day, delta=0):
            if delta != 0:
                raise NotImplementedError("stub calendar: only lag 0 is exact")
            return day

    africa.SouthAfrica = SouthAfrica
    pkg.africa = africa
    sys.modules["workalendar"] = pkg
    sys.modules["workalendar.africa"] = africa


def load_american(module_name: str):
    """module_name in {fd_american_option_pricer, fd_american_equity, fd_american_black76}."""
    _install_workalendar_stub()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module(module_name)


def load_barrier(module_name: str):
    """Exec the parsable prefix of a barrier file into a synthetic module."""
    n_lines = _BARRIER_PREFIX[module_name]
    path = os.path.join(REFERENCE_DIR, module_name + ".py")
    with open(path, "r", encoding="utf-8") as fh:
        src = "".join(fh.readlines()[:n_lines])
    mod = types.ModuleType("ref_" + module_name)
    mod.__file__ = path
    sys.modules[mod.__name__] = mod      # dataclasses looks the module up by name
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_utils():
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module("utils")


_DBFP_SPLICE = ((0, 745), (882, 1027))      # 0-based [start, end) line ranges of discrete_barrier_fdm_pricer.py


def load_dbfp():
    """DiscreteBarrierFDMPricer from the two parsable line ranges of discrete_barrier_fdm_pricer.py."""
    _install_workalendar_stub()
    path = os.path.join(REFERENCE_DIR, "discrete_barrier_fdm_pricer.py")
    with open(path, "r", encoding="utf-8") as fh:
        lines = fh.readlines()
    src = "".join("".join(lines[a:b]) for a, b in _DBFP_SPLICE)
    mod = types.ModuleType("ref_discrete_barrier_fdm_pricer")
    mod.__file__ = path
    sys.modules[mod.__name__] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    mod.DiscreteBarrierFDMPricer._build_stock_price_grid = lambda self: [0.0, 0.0]   # undefined in the reference
    return mod
