"""Vectorised host ingestion for books of American options: dates and daily NACA curves -> (T, r, b) and dividend tables
for a whole scenario file at once.

The reference derives these per pricer object: year fractions from the day count (fd_american_option_pricer.py:166-188),
discount factors ``(1 + NACA(date)) ** -tau`` with the NACA rate found by scanning the curve DataFrame for the ISO date
string (:190-197), the NACC forward rate between two dates (:199-203), settlement lags in South-African business days
(:101-115) and the dividend times ``tau = T - t_pay`` (:304-315) -- about 2 ms per option in pandas.  Here every step
is a numpy expression over all rows: dates are ``datetime64[D]``, the curve is one dense array indexed by day ordinal
(first occurrence of a date wins, a missing date raises ``ValueError`` like the reference), business-day lags use
``numpy.busday_offset`` with the holiday rules of ``calendar_sa``.  numpy's ``power`` / ``log`` can differ from
``math`` by an ulp, so ``r`` and ``b`` agree with the scalar classes to ~1e-16 relative, not bit for bit.
"""
from __future__ import annotations

import datetime as _dt
from typing import Optional, Sequence

import numpy as np

from .book import AmericanBook
from .calendar_sa import SouthAfrica

__all__ = ["DailyCurve", "year_fraction", "add_working_days", "american_book_from_dates"]

_D = "datetime64[D]"


def _days(dates) -> np.ndarray:
    """Anything date-like (``datetime.date``, ISO strings, ``datetime64``; scalar or sequence) -> ``datetime64[D]`` array."""
    a = np.asarray(dates)
    if a.dtype.kind == "M":
        return a.astype(_D)
    return np.asarray(a, dtype=_D)


class DailyCurve:
    """A daily NACA curve as a dense array over day ordinals.  ``curve``: DataFrame (or mapping) with columns ``Date``
    (ISO ``YYYY-MM-DD`` strings, as the pricers look dates up) and ``NACA``."""

    def __init__(self, curve):
        dates = _days(np.asarray(curve["Date"], dtype=str))
        rates = np.asarray(curve["NACA"], dtype=np.float64)
        self.first = dates.min()
        span = int((dates.max() - self.first).astype(np.int64)) + 1
        self._rate = np.full(span, np.nan)
        idx = (dates - self.first).astype(np.int64)
        # first match wins (the reference takes .iloc[0] of the matching rows): assign in reverse order
        self._rate[idx[::-1]] = rates[::-1]

    def naca(self, dates) -> np.ndarray:
        d = _days(dates)
        idx = (d - self.first).astype(np.int64)
        bad = (idx < 0) | (idx >= len(self._rate))
        out = self._rate[np.clip(idx, 0, len(self._rate) - 1)]
        bad |= np.isnan(out)
        if bad.any():
            first_bad = d.reshape(-1)[np.nonzero(bad.reshape(-1))[0][0]]
            raise ValueError(f"Discount factor not found for date: {first_bad}")      # fd_american_option_pricer.py:193-194
        return out


def year_fraction(day_count: str, start, end) -> np.ndarray:
    """``_year_fraction`` (fd_american_option_pricer.py:176-188) over arrays; ``day_count`` as the constructor normalises it."""
    dc = day_count.upper().replace("F", "")
    s, e = np.broadcast_arrays(_days(start), _days(end))
    days = (e - s).astype(np.int64).astype(np.float64)
    if dc in ("30/360", "BOND", "US30/360"):
        sy, sm, sd = _ymd(s)
        ey, em, ed = _ymd(e)
        d1 = np.minimum(sd, 30)
        d2 = np.where(d1 == 30, np.minimum(ed, 30), ed)
        out = ((ey - sy) * 360 + (em - sm) * 30 + (d2 - d1)) / 360.0
    else:
        den = {"ACT/365": 365.0, "ACT/360": 360.0, "ACT/364": 364.0}.get(dc, 365.0)
        out = days / den
    return np.where(e <= s, 0.0, out)


def _ymd(d: np.ndarray):
    y = d.astype("datetime64[Y]").astype(np.int64) + 1970
    m = d.astype("datetime64[M]").astype(np.int64) % 12 + 1
    day = (d - d.astype("datetime64[M]")).astype(np.int64) + 1
    return y, m, day


def add_working_days(dates, lag: int) -> np.ndarray:
    """``SouthAfrica().add_working_days(date, lag)`` over an array (lag 0: unchanged, like workalendar)."""
    d = _days(dates)
    if lag == 0:
        return d
    cal = SouthAfrica()
    years = range(int(d.min().astype("datetime64[Y]").astype(np.int64)) + 1969, int(d.max().astype("datetime64[Y]").astype(np.int64)) + 1973)
    hol = np.array(sorted(h for y in years for h in cal.holidays(y)), dtype=_D)
    # workalendar counts working days strictly after the start date, wherever the start falls
    start = np.busday_offset(d, 0, roll="backward" if lag > 0 else "forward", holidays=hol)
    return np.busday_offset(start, lag, holidays=hol)


def american_book_from_dates(spot, strike, valuation_date, maturity_date, sigma, option_type, discount_curve,
                             forward_curve=None, dividend_schedule: Optional[Sequence] = None, underlying_spot_days: int = 0,
                             option_days: int = 0, option_settlement_days: int = 0, day_count: str = "ACT/365",
                             num_space_nodes=400, num_time_steps=400, rannacher_steps=2, s_max_mult: float = 4.5,
                             black76: bool = False) -> AmericanBook:
    """The book of N pricer objects built from the same arguments as ``AmericanFDMPricer`` / ``AmericanFwdFDMPricer``
    (fd_american_option_pricer.py:43-160), every argument a scalar or an array of N.  ``discount_curve``: a
    :class:`DailyCurve` or a curve DataFrame shared by all rows.  ``dividend_schedule``: ``[(pay_date, amount)]`` shared by
    all rows (the runner's ``divs``).  ``option_type``: "call" / "put" (scalar or array)."""
    spot = np.atleast_1d(np.asarray(spot, dtype=np.float64))
    n = len(spot)
    val = np.broadcast_to(_days(valuation_date), (n,))
    mat = np.broadcast_to(_days(maturity_date), (n,))
    if np.any(mat <= val):
        raise ValueError("maturity_date must be after valuation_date.")
    ot = np.char.lower(np.broadcast_to(np.asarray(option_type, dtype=str), (n,)))
    if not np.all((ot == "call") | (ot == "put")):
        raise ValueError("option_type must be 'call' or 'put'.")
    curve = discount_curve if isinstance(discount_curve, DailyCurve) else DailyCurve(discount_curve)
    T = year_fraction(day_count, val, mat)
    if np.any(T <= 0.0):
        raise ValueError("time_to_expiry must be positive.")

    def discount_factor(d):                                   # :190-197
        return (1.0 + curve.naca(d)) ** (-year_fraction(day_count, val, d))

    def forward_rate(start, end):                             # :199-203
        tau = year_fraction(day_count, start, end)
        return -np.log(discount_factor(end) / discount_factor(start)) / np.maximum(1e-12, tau)

    r = forward_rate(add_working_days(val, option_days), add_working_days(mat, option_settlement_days))
    if forward_curve is not None and not black76:             # the reference reads the DISCOUNT curve here, :131-134
        b = forward_rate(add_working_days(val, underlying_spot_days), add_working_days(mat, underlying_spot_days))
    else:
        b = r
    div = None
    if dividend_schedule and not black76:                     # :304-315
        pay = _days([d for d, _ in dividend_schedule])
        amt = np.asarray([float(a) for _, a in dividend_schedule], dtype=np.float64)
        order = np.argsort(pay, kind="stable")
        pay, amt = pay[order], amt[order]
        t_rel = year_fraction(day_count, val[:, None], pay[None, :])
        keep = (val[:, None] < pay[None, :]) & (pay[None, :] < mat[:, None]) & (t_rel > 0.0) & (t_rel < T[:, None])
        tau = np.where(keep, T[:, None] - t_rel, np.inf)
        cash = np.broadcast_to(amt[None, :], tau.shape)
        srt = np.argsort(tau, axis=1, kind="stable")          # ascending tau
        tau = np.take_along_axis(tau, srt, axis=1)
        cash = np.take_along_axis(cash, srt, axis=1)
        ok = np.isfinite(tau)
        counts = ok.sum(axis=1)
        if counts.sum():
            div = (np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), np.ascontiguousarray(tau[ok]),
                   np.ascontiguousarray(cash[ok]))
    return AmericanBook(spot, strike, T, sigma, r, b, ot == "call", num_space_nodes, num_time_steps, rannacher_steps,
                        black76=black76, s_max_mult=s_max_mult, div_taus=div)
