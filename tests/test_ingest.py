"""vectorised host ingestion (dates + daily NACA curve -> T, r, b, dividend tables) against the scalar
drop-in class, which follows the reference's constructor operation for operation."""
import datetime as dt

import numpy as np
import pytest

from finite_difference_b200 import ingest, utils
from finite_difference_b200.calendar_sa import SouthAfrica
from finite_difference_b200.fd_american_black76 import AmericanFwdFDMPricer
from finite_difference_b200.fd_american_equity import AmericanFDMPricer

VAL, MAT = dt.date(2025, 7, 28), dt.date(2026, 7, 28)
DIVS = [(dt.date(2025, 9, 12), 7.63), (dt.date(2026, 4, 10), 8.0115), (dt.date(2027, 1, 1), 3.0), (dt.date(2025, 7, 28), 1.0)]


def _curve():
    c = utils.create_rate_df_iso(0.07)
    c["NACA"] = 0.05 + 1e-5 * np.arange(len(c))            # not flat
    return c


@pytest.mark.parametrize("day_count", ["ACT/365", "ACT/365F", "ACT/360", "ACT/364", "30/360", "BOND"])
@pytest.mark.parametrize("lags", [(0, 0, 0), (2, 3, 1)])
def test_book_from_dates_matches_the_scalar_constructor(day_count, lags):
    c = _curve()
    mats = [MAT, dt.date(2026, 1, 30), dt.date(2027, 3, 31)]
    bk = ingest.american_book_from_dates([229.74, 100.0, 50.0], [220.0, 90.0, 55.0], VAL, mats, [0.26, 0.3, 0.2],
                                         ["put", "call", "put"], c, c, DIVS, underlying_spot_days=lags[0], option_days=lags[1],
                                         option_settlement_days=lags[2], day_count=day_count, num_space_nodes=300, num_time_steps=200)
    begin, tau, cash = bk.div_csr()
    for i, (S, K, sig, ot) in enumerate(((229.74, 220.0, 0.26, "put"), (100.0, 90.0, 0.3, "call"), (50.0, 55.0, 0.2, "put"))):
        pr = AmericanFDMPricer(S, K, VAL, mats[i], sig, ot, c, c, DIVS, underlying_spot_days=lags[0], option_days=lags[1],
                               option_settlement_days=lags[2], day_count=day_count)
        assert bk.T[i] == pr.time_to_expiry
        assert abs(bk.r[i] - pr.discount_rate_nacc) <= 4e-16 and abs(bk.b[i] - pr.carry_rate_nacc) <= 4e-16
        assert list(zip(tau[begin[i]:begin[i + 1]].tolist(), cash[begin[i]:begin[i + 1]].tolist())) == pr._div_times_tau()
        assert bool(bk.is_call[i]) == (ot == "call")


def test_black76_book_ignores_dividends_and_carry():
    c = _curve()
    bk = ingest.american_book_from_dates([229.74], [220.0], VAL, MAT, 0.26, "call", c, c, DIVS, black76=True)
    pr = AmericanFwdFDMPricer(229.74, 220.0, VAL, MAT, 0.26, "call", c, c)
    assert bk.div_csr() is None and bk.T[0] == pr.time_to_expiry and abs(bk.r[0] - pr.discount_rate_nacc) <= 4e-16
    assert bk.b[0] == bk.r[0]


def test_missing_curve_date_and_bad_inputs_raise_like_the_reference():
    c = _curve()
    with pytest.raises(ValueError, match="Discount factor not found"):
        ingest.american_book_from_dates([1.0], [1.0], VAL, dt.date(2031, 1, 1), 0.2, "put", c)
    with pytest.raises(ValueError, match="maturity_date"):
        ingest.american_book_from_dates([1.0], [1.0], VAL, VAL, 0.2, "put", c)
    with pytest.raises(ValueError, match="option_type"):
        ingest.american_book_from_dates([1.0], [1.0], VAL, MAT, 0.2, "straddle", c)


def test_first_duplicate_curve_row_wins():
    import pandas as pd
    c = pd.DataFrame({"Date": ["2025-07-28", "2025-07-29", "2025-07-29", "2025-07-30"], "NACA": [0.01, 0.02, 0.09, 0.03]})
    assert ingest.DailyCurve(c).naca([dt.date(2025, 7, 29)])[0] == 0.02


def test_sunday_holiday_substitution_cascades():
    """25 December on a Sunday (2016, 2022): the Monday is already the Day of Goodwill, so the substitute is the 27th."""
    cal = SouthAfrica()
    for year in (2016, 2022):
        assert dt.date(year, 12, 27) in cal.holidays(year)
        assert not cal.is_working_day(dt.date(year, 12, 27))
    assert dt.date(2023, 12, 27) not in cal.holidays(2023)
    assert dt.date(2023, 1, 2) in cal.holidays(2023)                       # 1 January 2023 was a Sunday
    got = ingest.add_working_days([dt.date(2022, 12, 23)], 1)[0]
    assert got == np.datetime64(cal.add_working_days(dt.date(2022, 12, 23), 1))
    assert got == np.datetime64("2022-12-28")


def test_hundred_thousand_rows_in_a_fraction_of_a_second():
    import time
    n = 100_000
    c = _curve()
    t0 = time.perf_counter()
    bk = ingest.american_book_from_dates(np.full(n, 229.74), np.linspace(180, 280, n), VAL, MAT, 0.26, "put", c, c, DIVS)
    assert time.perf_counter() - t0 < 2.0 and bk.n_opt == n and bk.div_csr()[0][-1] == 2 * n
