"""South-African business-day calendar for the pricers' settlement lags.

The reference delegates to the third-party package ``workalendar`` (``requirements.txt:2``,
unpinned, not vendored): ``SouthAfrica().add_working_days(date, n)``
(fd_american_option_pricer.py:101-115).  Every reference driver passes lag 0
(run_american_scenarios.py:299-301), for which the date is returned unchanged -- that case is
exact.  For non-zero lags the holiday rules below restate the South African Public Holidays Act
(fixed dates, Good Friday, Family Day, Sunday -> Monday substitution, since 1995); parity with
``workalendar`` for those lags is UNPINNED (the package is absent from this image).
"""
from __future__ import annotations

import datetime as _dt

_FIXED = ((1, 1), (3, 21), (4, 27), (5, 1), (6, 16), (8, 9), (9, 24), (12, 16), (12, 25), (12, 26))


def _easter_sunday(year: int) -> _dt.date:
    """Anonymous Gregorian algorithm."""
    a = year % 19
    b, c = divmod(year, 100)
    d, e = divmod(b, 4)
    f = (b + 8) // 25
    g = (b - f + 1) // 3
    h = (19 * a + b - d - g + 15) % 30
    i, k = divmod(c, 4)
    l = (32 + 2 * e + 2 * i - h - k) % 7
    m = (a + 11 * h + 22 * l) // 451
    month, day = divmod(h + l - 7 * m + 114, 31)
    return _dt.date(year, month, day + 1)


class SouthAfrica:
    """Minimal stand-in for ``workalendar.africa.SouthAfrica`` (only what the pricers call)."""

    def holidays(self, year: int):
        days = {_dt.date(year, m, d) for m, d in _FIXED}
        easter = _easter_sunday(year)
        days.add(easter - _dt.timedelta(days=2))     # Good Friday
        days.add(easter + _dt.timedelta(days=1))     # Family Day
        for h in sorted(days):
            if h.weekday() == 6:                     # a holiday on a Sunday is observed on the next day that is not
                sub = h + _dt.timedelta(days=1)      # itself a holiday: 25 Dec on a Sunday (2016, 2022) -> 27 Dec, because
                while sub in days:                   # the Monday is the Day of Goodwill
                    sub += _dt.timedelta(days=1)
                days.add(sub)
        return days

    def is_working_day(self, day: _dt.date) -> bool:
        return day.weekday() < 5 and day not in self.holidays(day.year)

    def add_working_days(self, day: _dt.date, delta: int = 0) -> _dt.date:
        """``delta`` working days after (before, if negative) ``day``; ``delta == 0`` returns ``day``."""
        if delta == 0:
            return day
        step = _dt.timedelta(days=1 if delta > 0 else -1)
        left = abs(int(delta))
        while left:
            day = day + step
            if self.is_working_day(day):
                left -= 1
        return day
