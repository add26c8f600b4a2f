"""Drop-in for the reference's American scenario runner (run_american_scenarios.py).

Same function names, arguments, CSV schema and result keys (run_american_scenarios.py:46-73,
:174-204, :209-277).  Differences, all deliberate:

* the reference runner cannot run: it parses the forward-curve dates with ``format="%Y/%m/%m"``
  (:140-142), which raises ``re.error`` on any pandas.  The curves here are built with ISO dates
  directly, which is what the line intends.
* ``run_all_american_scenarios`` prices the whole CSV -- with or without cash dividends -- as ONE fused call on
  the GPU (8 PDE solves per row: ``price_log2()`` + ``greeks_log2()``, :144-172) instead of a serial Python loop,
  builds the result frame column-wise, and can shard the rows over the GPUs of a box (``torchrun``; see
  ``shard.py``).  ``run_american_scenario`` keeps the one-row path through the drop-in pricer class.
* books with per-row dates or a real (non-flat) daily curve: ``ingest.american_book_from_dates``.
"""
from __future__ import annotations

import datetime as dt
import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from . import shard
from .book import AmericanBook
from .fd_american_equity import AmericanFDMPricer
from .utils import create_rate_df_iso, naca_to_nacc

_FA = ("price", "delta", "gamma", "vega")


def _percentage_diff(model_val: float, fa_val: Optional[float]) -> float:
    """run_american_scenarios.py:24-43."""
    if fa_val is None or np.isnan(fa_val) or fa_val == 0.0:
        return np.nan
    return abs(model_val - fa_val) / abs(fa_val) * 100.0


def _result_row(scenario_name, S0, K, sigma, rate, model: Dict[str, float], fa: Dict[str, Optional[float]]) -> Dict[str, Any]:
    """Result dictionary in the reference's key order (run_american_scenarios.py:174-204)."""
    row: Dict[str, Any] = {"scenario_name": scenario_name, "S0": S0, "K": K, "sigma": sigma, "rate": rate}
    for name in _FA:
        m, f = model[name], fa[name]
        row[f"model_{name}"] = m
        row[f"FA_{name}"] = f if f is not None else np.nan
        row[f"{name}_diff"] = abs(m - f) if f is not None else np.nan
        row[f"{name}_pct_diff"] = _percentage_diff(m, f)
    return row


def run_american_scenario(scenario_name: str, S0: float, K: float, sigma: float, rate: float,
                          FA_price: Optional[float], FA_delta: Optional[float], FA_gamma: Optional[float],
                          FA_vega: Optional[float], valuation: dt.date, maturity: dt.date, opt_type: str = "call",
                          trade_number: int = 201871103, quantity: int = 1000, contract_size: int = 1,
                          position: str = "long", divs: Optional[list] = None, underlying_spot_days: int = 0,
                          option_days: int = 0, option_settlement_days: int = 0, day_count: str = "ACT/365",
                          grid_type: str = "uniform", num_space_nodes: int = 500, num_time_steps: int = 500,
                          rannacher_steps: int = 2) -> Dict[str, Any]:
    """One scenario through the drop-in pricer class (run_american_scenarios.py:46-206)."""
    if divs is None:
        divs = []
    _nacc_rate = naca_to_nacc(rate)  # noqa: F841  (kept by the reference "for symmetry", :130)
    discount_curve = create_rate_df_iso(rate)
    forward_curve = create_rate_df_iso(rate)
    pricer = AmericanFDMPricer(
        spot=S0, strike=K, valuation_date=valuation, maturity_date=maturity, sigma=sigma, option_type=opt_type,
        discount_curve=discount_curve, forward_curve=forward_curve, dividend_schedule=divs, trade_id=trade_number,
        direction=position, quantity=quantity, contract_multiplier=contract_size,
        underlying_spot_days=underlying_spot_days, option_days=option_days,
        option_settlement_days=option_settlement_days, day_count=day_count, grid_type=grid_type,
        num_space_nodes=num_space_nodes, num_time_steps=num_time_steps, rannacher_steps=rannacher_steps)
    model_price = pricer.price_log2()
    greeks = pricer.greeks_log2()
    model = {"price": model_price, "delta": greeks["delta"], "gamma": greeks["gamma"], "vega": greeks["vega"]}
    return _result_row(scenario_name, S0, K, sigma, rate, model,
                       {"price": FA_price, "delta": FA_delta, "gamma": FA_gamma, "vega": FA_vega})


def scenario_book(S0: Sequence[float], K: Sequence[float], sigma: Sequence[float], rate: Sequence[float],
                  valuation: dt.date, maturity: dt.date, opt_type: str = "call", divs: Optional[list] = None,
                  underlying_spot_days: int = 0, option_days: int = 0, option_settlement_days: int = 0,
                  day_count: str = "ACT/365", num_space_nodes: int = 500, num_time_steps: int = 500,
                  rannacher_steps: int = 2, **_ignored) -> AmericanBook:
    """The book behind a scenario CSV.  Dates -> (T, r, b) exactly as the pricer constructor derives them from
    a flat NACA curve (fd_american_equity.py:215-245): one probe pricer supplies the year fractions, the flat
    curve gives ``r = b = -ln((1+NACA)^-t_far / (1+NACA)^-t_near) / tau`` per row."""
    rate = np.asarray(rate, dtype=np.float64)
    probe = AmericanFDMPricer(
        spot=1.0, strike=1.0, valuation_date=valuation, maturity_date=maturity, sigma=0.2, option_type=opt_type,
        discount_curve=create_rate_df_iso(0.0), forward_curve=None, dividend_schedule=divs or [],
        underlying_spot_days=underlying_spot_days, option_days=option_days,
        option_settlement_days=option_settlement_days, day_count=day_count)
    yf = probe._year_fraction

    def fwd_rate(start, end):
        t_far, t_near, tau = yf(valuation, end), yf(valuation, start), yf(start, end)
        if len(rate) > 2048:                                    # numpy power / log: within an ulp of the scalar path
            return -np.log((1.0 + rate) ** (-t_far) / (1.0 + rate) ** (-t_near)) / max(1e-12, tau)
        return np.array([-math.log((1.0 + x) ** (-t_far) / (1.0 + x) ** (-t_near)) / max(1e-12, tau) for x in rate])

    r = fwd_rate(probe.discount_start_date, probe.discount_end_date)
    b = fwd_rate(probe.carry_start_date, probe.carry_end_date)      # forward_curve is always given by the runner
    div_taus = probe._div_times_tau()
    n = len(rate)
    csr = None
    if div_taus:                                                # the same schedule for every row, as CSR arrays
        k = len(div_taus)
        csr = (np.arange(0, (n + 1) * k, k, dtype=np.int32), np.tile([t for t, _ in div_taus], n),
               np.tile([c for _, c in div_taus], n))
    return AmericanBook(S0, K, probe.time_to_expiry, sigma, r, b, opt_type.lower() == "call", num_space_nodes,
                        num_time_steps, rannacher_steps, black76=False, div_taus=csr)


def run_all_american_scenarios(config_csv_path: str, output_csv_path: str, base_params: Dict[str, Any],
                               distributed: bool = False, lapack_greeks: bool = False,
                               verbose: Optional[bool] = None) -> pd.DataFrame:
    """Read the configuration CSV, price every row in one GPU batch, write the results CSV
    (run_american_scenarios.py:209-277; columns scenario_name,S0,K,sigma,rate,FA_price,FA_delta,FA_gamma,FA_vega).

    The whole file -- dividends in ``base_params["divs"]`` included -- is ONE fused call (``AmericanBook.run_fused``:
    fan-out of the 8 solves per row, time segments between dividends, Richardson / bump epilogue on the device).
    ``lapack_greeks=True`` takes the host-buffer path instead and recomputes delta / gamma with the reference's LAPACK
    solve of the four-point cubic (fd_american_option_pricer.py:642-653; the in-kernel divided differences agree
    with it to ~1e-9 relative).  ``distributed=True`` under ``torchrun``: rows are sharded over the ranks, results
    gathered, rank 0 writes the CSV.  ``verbose``: one line per scenario like the reference (default: files up to
    1000 rows).
    """
    config_df = pd.read_csv(config_csv_path)
    n = len(config_df)
    rank, world = 0, 1
    mine = np.arange(n)
    if distributed:
        rank, world, _ = shard.init_distributed()
        mine = shard.partition(np.ones(n), world)[rank]
    sub = config_df.iloc[mine]
    if len(sub):
        book = scenario_book(sub["S0"].to_numpy(float), sub["K"].to_numpy(float), sub["sigma"].to_numpy(float),
                             sub["rate"].to_numpy(float), **base_params)
        out = book.run("scenario", lapack_greeks=True) if lapack_greeks else book.run_fused("scenario")
        cols = {"price": out["price_log2"], "delta": out["delta"], "gamma": out["gamma"], "vega": out["vega"]}
    else:
        cols = {k: np.zeros(0) for k in _FA}
    if distributed and world > 1:
        cols = shard.gather_results(cols, mine, n)
    results_df = pd.DataFrame()
    if rank == 0:
        # the reference's result dictionary (:174-204), one column at a time
        data: Dict[str, Any] = {k: config_df[k].to_numpy() for k in ("scenario_name", "S0", "K", "sigma", "rate")}
        for name in _FA:
            m = np.asarray(cols[name], dtype=np.float64)
            f = config_df[f"FA_{name}"].to_numpy(dtype=np.float64)
            diff = np.abs(m - f)
            with np.errstate(divide="ignore", invalid="ignore"):
                pct = np.where(np.isnan(f) | (f == 0.0), np.nan, diff / np.abs(f) * 100.0)
            data[f"model_{name}"], data[f"FA_{name}"], data[f"{name}_diff"], data[f"{name}_pct_diff"] = m, f, diff, pct
        results_df = pd.DataFrame(data)
        if verbose if verbose is not None else n <= 1000:
            for row in results_df.itertuples(index=False):
                print(f"{row.scenario_name}:  Price %Diff: {row.price_pct_diff:.4f}%, Delta %Diff: {row.delta_pct_diff:.4f}%, "
                      f"Gamma %Diff: {row.gamma_pct_diff:.4f}%, Vega %Diff: {row.vega_pct_diff:.4f}%")
        results_df.to_csv(output_csv_path, index=False)
        print(f"\nAmerican option results saved to {output_csv_path}")
    return results_df


if __name__ == "__main__":
    base_parameters: Dict[str, Any] = {
        "valuation": dt.date(2025, 7, 28), "maturity": dt.date(2025, 8, 28), "opt_type": "put",
        "trade_number": 201871200, "quantity": 1000, "contract_size": 1, "position": "long", "divs": [],
        "underlying_spot_days": 0, "option_days": 0, "option_settlement_days": 0, "day_count": "ACT/365",
        "grid_type": "uniform", "num_space_nodes": 500, "num_time_steps": 500, "rannacher_steps": 2,
    }
    df = run_all_american_scenarios("american_config_scenarios.csv", "american_scenario_results.csv", base_parameters)
    print(df.to_string(index=False))
