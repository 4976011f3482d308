"""Host helpers on or next to the per-tick path (reference utils.py): dates, seasonality, mortality table, timers."""

from __future__ import annotations

import calendar
import datetime as dt
from collections import defaultdict
from time import perf_counter_ns

import numpy as np

__all__ = ["date", "daterange", "get_doy", "get_seasonality", "create_cumulative_deaths", "TimingStats"]


def date(value):
    """'YYYY-MM-DD' or date -> datetime.date (reference utils.py:137-143)."""
    if isinstance(value, dt.datetime):
        return value.date()
    if isinstance(value, dt.date):
        return value
    return dt.datetime.strptime(value, "%Y-%m-%d").date()  # noqa: DTZ007 - calendar dates, no time of day


def daterange(start_date, days):
    """``days`` consecutive dates from ``start_date`` as an object array (reference utils.py:146-161)."""
    start = date(start_date)
    return np.array([start + dt.timedelta(days=i) for i in range(int(days))])


def get_doy(sim) -> int:
    return sim.datevec[sim.t].timetuple().tm_yday


def get_seasonality(sim) -> float:
    """1 + A cos(2 pi (doy - peak) / days_in_year), leap-year aware (reference utils.py:616-625)."""
    day = sim.datevec[sim.t]
    days_in_year = 366 if calendar.isleap(day.year) else 365
    return 1 + sim.pars["seasonal_amplitude"] * np.cos(2 * np.pi * (get_doy(sim) - sim.pars["seasonal_peak_doy"]) / days_in_year)


def create_cumulative_deaths(total_population, max_age_years):
    """Back-loaded mortality: yearly hazard 1e-4 * 2^(age/10) (reference utils.py:795-816)."""
    ages = np.arange(max_age_years + 1)
    hazard = 0.0001 * 2.0 ** (ages / 10)
    return np.cumsum(hazard * total_population).astype(int)


class TimingStats:
    """Nested stopwatches keyed like the reference's (utils.py:819-945): ``with stats.start("X.step()")``.

    The per-tick kernels are asynchronous; a stopwatch therefore measures host enqueue time unless
    ``sync`` is set, in which case the CUDA stream is synchronised before the clock stops."""

    def __init__(self, sync: bool = False):
        self.stats = defaultdict(int)
        self.depth = 0
        self.sync = sync

    class _Stopwatch:
        def __init__(self, key, owner):
            self.key, self.owner = key, owner

        def __enter__(self):
            self.t0 = perf_counter_ns()
            return self

        def __exit__(self, *exc):
            if self.owner.sync:
                import torch

                if torch.cuda.is_available():
                    torch.cuda.current_stream().synchronize()
            self.owner.stats[self.key] += perf_counter_ns() - self.t0
            self.owner.depth -= 1

    def start(self, label):
        key = " " * (4 * self.depth) + label
        self.depth += 1
        self.stats[key] += 0
        return TimingStats._Stopwatch(key, self)

    def log(self, logger):
        if not self.stats:
            return
        width = max(map(len, self.stats))
        for label, ns in self.stats.items():
            logger.info(f"{label:<{width}} : {round(ns / 1000):11,} µsecs")
