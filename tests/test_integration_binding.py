"""INTEGRATION.md section B, executed verbatim: the ctypes stub a reference maintainer would drop next to
``fd_american_option_pricer.py`` is cut out of the document, exec'd, and assigned to ``AmericanFDMPricer._solve_grid``
of the UNMODIFIED reference class (from /root/reference in the build container, from the staged copy under
``baseline/_ref/`` on the GPU box).  ``price_log`` / ``price_log2`` of the patched reference class must then
reproduce the reference's own numbers (goldens) to 1e-10."""
import datetime as dt
import os
import re

import pytest

import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", doc, flags=re.S)
    src = next(b for b in blocks if "def solve_grid_on_gpu" in b)
    assert "/path/to/finite_difference_b200/libfdm_b200.so" in src
    return src.replace("/path/to/finite_difference_b200/libfdm_b200.so",
                       os.path.join(ROOT, "finite_difference_b200", "libfdm_b200.so"))


def test_stub_is_present_and_declares_the_current_abi():
    from finite_difference_b200 import _lib
    src = _stub_source()
    assert f"L.fdm_version() == {_lib.ABI_VERSION}" in src
    names = re.search(r'for n in \("x_min".*?\)\]', src, flags=re.S).group(0)
    for field, _ in _lib.FdmBatch._fields_[5:]:
        assert f'"{field}"' in names, field


@pytest.mark.gpu
def test_stub_drives_the_unmodified_reference_class():
    from oracle import pyref, ref_loader
    if not pyref.available():
        pytest.skip("no reference checkout and no staged copy under baseline/_ref")
    ref_loader.REFERENCE_DIR = pyref.reference_dir()
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md#B", "exec"), ns)
    mod = ref_loader.load_american("fd_american_option_pricer")
    original = mod.AmericanFDMPricer._solve_grid
    mod.AmericanFDMPricer._solve_grid = ns["solve_grid_on_gpu"]
    try:
        import pandas as pd
        utils = ref_loader.load_utils()
        checked = 0
        for c in _golden.american_cases():
            if c["module"] != "fd_american_option_pricer" or not c["forward_curve"]:
                continue
            curve = utils.create_rate_df(c["naca"])
            curve["Date"] = pd.to_datetime(curve["Date"], format="%Y/%m/%d").dt.strftime("%Y-%m-%d")
            pr = mod.AmericanFDMPricer(
                spot=c["spot"], strike=c["strike"], valuation_date=dt.date.fromisoformat(c["valuation"]),
                maturity_date=dt.date.fromisoformat(c["maturity"]), sigma=c["sigma"], option_type=c["option_type"],
                discount_curve=curve, forward_curve=curve.copy(),
                dividend_schedule=[(dt.date.fromisoformat(d), a) for d, a in c["dividends"]],
                day_count=c["day_count"], num_space_nodes=c["num_space_nodes"], num_time_steps=c["num_time_steps"],
                rannacher_steps=c["rannacher_steps"], s_max_mult=c["s_max_mult"])
            p = pr.price_log()
            assert abs(p - c["out"]["price_log"]) <= 1e-10 * abs(c["out"]["price_log"]), c["id"]
            if c["out"].get("price_log2") is not None:
                p2 = pr.price_log2()
                assert abs(p2 - c["out"]["price_log2"]) <= 1e-10 * abs(c["out"]["price_log2"]), c["id"]
            checked += 1
        assert checked >= 5
    finally:
        mod.AmericanFDMPricer._solve_grid = original
