"""Drop-in for the reference module ``utils`` (utils.py:17-104): rate conversions and the flat
daily NACA curve the scenario runner builds per row (host only)."""
from __future__ import annotations

import datetime as dt

import numpy as np
import pandas as pd


def nacc_to_naca(nacc_rate: float) -> float:
    """NACA = exp(NACC) - 1 (utils.py:17-41)."""
    return float(np.exp(nacc_rate) - 1.0)


def naca_to_nacc(naca_rate: float) -> float:
    """NACC = ln(1 + NACA) (utils.py:44-68)."""
    return float(np.log(1.0 + naca_rate))


CURVE_START = dt.date(2025, 7, 28)
CURVE_END = dt.date(2028, 9, 28)


def create_rate_df(rate: float) -> pd.DataFrame:
    """Flat NACA curve, daily from 2025-07-28 to 2028-09-28, dates as "YYYY/MM/DD" (utils.py:71-104)."""
    date_range = pd.date_range(start=CURVE_START, end=CURVE_END, freq="D")
    return pd.DataFrame({"Date": date_range.strftime("%Y/%m/%d"), "NACA": rate})


def create_rate_df_iso(rate: float) -> pd.DataFrame:
    """The same curve with ISO dates, the format the pricers look dates up in
    (what run_american_scenarios.py:135-142 converts to before building a pricer)."""
    date_range = pd.date_range(start=CURVE_START, end=CURVE_END, freq="D")
    return pd.DataFrame({"Date": date_range.strftime("%Y-%m-%d"), "NACA": rate})
