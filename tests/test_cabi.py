"""The C-ABI library builds, loads and exports every symbol include/fdm_b200.h declares.
No compute calls here (no GPU in the CPU tier); the error path without a device is checked."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from finite_difference_b200 import _lib, build as fdm_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fdm_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+)?(?:int|char|size_t|void)\s*\*?\s*(fdm_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_binding_exports():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_builds_and_exports_every_symbol():
    path = fdm_build.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    for name in declared_functions():
        assert hasattr(L, name), name
    assert _lib.load().fdm_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # FdmBatch: int64 + 4 x int32 + 24 pointers (ABI 3); FdmOut: 7 pointers + (ptr, int64) x 2
    assert C.sizeof(_lib.FdmBatch) == 8 + 16 + 24 * 8
    assert C.sizeof(_lib.FdmOut) == 7 * 8 + 4 * 8
    # and the compiled library agrees with the ctypes declarations
    L = _lib.load()
    for which, st in enumerate((_lib.FdmBatch, _lib.FdmOut, _lib.FdmAmericanBook, _lib.FdmBookOut, _lib.FdmBarrierBook, _lib.FdmRowVarBatch,
                                _lib.FdmRowVarOut)):
        assert L.fdm_sizeof(which) == C.sizeof(st), st.__name__


def test_invalid_arguments_are_reported_not_thrown():
    L = _lib.load()
    assert L.fdm_solve_batch(None, None, None, None) == -1
    assert b"NULL" in L.fdm_last_error()


def test_host_entry_rejects_n_space_outside_the_declared_bound():
    """A solve no kernel would take (n_space beyond n_space_max, or < 3) must not come back as FDM_OK with
    untouched outputs: the host entry scans the arrays and refuses (checked before any device call)."""
    from finite_difference_b200 import hostlogic
    L = _lib.load()
    b = hostlogic.build_american_batch([100.0, 100.0], [100.0, 90.0], [1.0], [0.2], [0.05], [0.05], [False], [50, 60], [50])
    keep = []

    def host_ptr(_name, arr):
        keep.append(arr)
        return arr.ctypes.data

    sb = b._struct(host_ptr)
    out = _lib.FdmOut()
    price = np.empty(2)
    out.price = price.ctypes.data
    sb.n_space_max = 55                                   # the second solve has 60 nodes
    assert L.fdm_solve_batch_host(C.byref(sb), C.byref(out), None, None) == -1
    assert b"n_space" in L.fdm_last_error()
    sb.n_space_max = 60
    b.seg_dt[1] = 0.0
    assert L.fdm_solve_batch_host(C.byref(sb), C.byref(out), None, None) == -1
    assert b"seg_dt" in L.fdm_last_error()


def test_book_entries_validate_on_the_host():
    L = _lib.load()
    bk, out = _lib.FdmAmericanBook(), _lib.FdmBookOut()
    bk.n_options = 1
    assert L.fdm_american_book_host(C.byref(bk), C.byref(out), None, None, None) == -1
    assert L.fdm_american_book(C.byref(bk), C.byref(out), None, 0, None, None) == -1
    bb = _lib.FdmBarrierBook()
    bb.n_options = 1
    assert L.fdm_barrier_book_host(C.byref(bb), C.byref(out), None, None, None) == -1
    assert L.fdm_barrier_book(C.byref(bb), C.byref(out), None, 0, None, None) == -1
    assert L.fdm_american_book_workspace(1000, 0, 3, 1) > 1000 * 8 * 8 * 7
    assert L.fdm_barrier_book_workspace(1000, 1) > 1000 * 3 * 8 * 7
    assert L.fdm_barrier_book_solves_per_option(1) == 3


def test_no_device_fails_loudly():
    """Without a CUDA device every solve entry returns an error: there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from finite_difference_b200 import hostlogic, solve_host
    b = hostlogic.build_american_batch([100.0], [100.0], [1.0], [0.2], [0.05], [0.05], [False], [50], [50])
    with pytest.raises(_lib.FdmError) as ei:
        solve_host(b)
    assert ei.value.code in (-4, -2)
    from finite_difference_b200 import DeviceBatch
    with pytest.raises(_lib.FdmError):
        DeviceBatch(b)
