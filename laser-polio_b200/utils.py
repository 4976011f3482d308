"""Host helpers on or next to the per-tick path (reference utils.py): dates, seasonality, mortality table, timers."""

from __future__ import annotations

import calendar
import datetime as dt
from collections import defaultdict
from time import perf_counter_ns

import numpy as np

__all__ = ["date", "daterange", "get_doy", "get_seasonality", "create_cumulative_deaths", "TimingStats", "save_sim_results",
           "add_temporal_groupings", "add_regional_groupings", "results_long_table"]


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


# --------------------------------------------------------------------------- results long table
RESULT_COLUMNS = (("S", "S"), ("E", "E"), ("I", "I"), ("R", "R"), ("P", "paralyzed"), ("births", "births"), ("deaths", "deaths"),
                  ("new_exposed", "new_exposed"), ("potentially_paralyzed", "potentially_paralyzed"),
                  ("new_potentially_paralyzed", "new_potentially_paralyzed"), ("new_paralyzed", "new_paralyzed"))


def results_long_table(sim) -> dict:
    """The columns of the reference's long results table (utils.py:716-747), one row per (timestep, node), time-major: the
    ``[nt, nodes]`` results arrays are already laid out that way (on the device too), so every column is a flat view."""
    nt, nodes = int(sim.nt), len(sim.nodes)
    lookup = getattr(sim.pars, "node_lookup", None) or {}
    names = np.array([lookup.get(n, {}).get("dot_name", "UNKNOWN") for n in range(nodes)], dtype=object)
    data = {"timestep": np.repeat(np.arange(nt), nodes), "date": np.repeat(np.asarray(sim.datevec), nodes),
            "node": np.tile(np.arange(nodes), nt), "dot_name": np.tile(names, nt)}
    for column, attr in RESULT_COLUMNS:
        arr = getattr(sim.results, attr, None)
        data[column] = np.zeros(nt * nodes, np.int32) if arr is None else np.asarray(arr).reshape(nt * nodes)
    return data


def add_temporal_groupings(df, time_config):
    """``time_period`` column from ``{"bins": [dates], "labels": [...]}`` (reference utils.py:1040-1073): left-closed periods."""
    import pandas as pd

    df = df.copy()
    if "bins" in time_config and "labels" in time_config:
        edges = [pd.Timestamp.min, *[pd.Timestamp(d) for d in time_config["bins"]], pd.Timestamp.max]
        df["time_period"] = pd.cut(df["date"], bins=edges, labels=time_config["labels"], right=False)
    return df


root = None  # like the reference's ``lp.root``: where data/regions.yaml lives (set by the caller; None = current directory)


def add_regional_groupings(df, region_groupings=None, grouping_level="adm0", regions_yaml_path=None):
    """``adm0`` / ``adm1`` / ``adm01`` columns from ``dot_name`` ("AFRO:COUNTRY:STATE:LGA") and a ``region`` column = the
    chosen admin level, overridden by the named groups of regions.yaml for the countries in ``region_groupings``
    (reference utils.py:1076-1159; calibration targets aggregate by it, calib/targets.py).  A group's patterns match the
    adm01 value when grouping at adm01 and the pattern has a colon, the adm1 value at adm1, else any dot_name containing
    the pattern (case-insensitive).  Unknown countries and a missing YAML keep the admin level, with a warning."""
    from pathlib import Path

    df = df.copy()
    parts = df["dot_name"].astype(str).str.split(":", expand=True)
    df["adm0"], df["adm1"] = parts[1], parts[2]
    df["adm01"] = df["adm0"] + ":" + df["adm1"]
    if grouping_level not in {"adm0", "adm01", "dot_name"}:
        raise ValueError(f"Invalid grouping_level: {grouping_level}. Must be one of 'adm0', 'adm01', or 'dot_name'.")
    df["region"] = df[grouping_level]
    if not region_groupings:
        return df
    path = Path(regions_yaml_path) if regions_yaml_path else Path(root or ".") / "data" / "regions.yaml"
    try:
        import yaml

        with open(path) as fh:
            groups_by_country = yaml.safe_load(fh)
    except (FileNotFoundError, ImportError):
        print(f"Warning: Could not load {path}, using adm0 for all countries")
        return df
    names = df["dot_name"].astype(str)
    for country in region_groupings:
        key = country.upper()
        if key not in groups_by_country:
            print(f"Warning: Custom regions not found for {country} in regions.yaml, using adm0")
            continue
        in_country = df["adm0"] == key
        for group, patterns in groups_by_country[key].items():
            hit = np.zeros(len(df), dtype=bool)
            for pattern in patterns:
                if grouping_level == "adm01" and ":" in pattern:
                    hit |= (df["adm01"] == pattern).to_numpy()
                else:
                    hit |= names.str.contains(pattern, case=False, na=False).to_numpy()
            df.loc[in_country.to_numpy() & hit, "region"] = group
    return df


def save_sim_results(sim, filename="simulation_results.h5", summary_config=None):
    """Reference ``save_sim_results`` (utils.py:690-786): the long table as a DataFrame with the temporal and regional
    groupings of ``summary_config`` applied, written as HDF5 (key "results") when pandas can (PyTables present) and the name
    ends in .h5, else as CSV (next to it when PyTables is missing)."""
    from pathlib import Path

    import pandas as pd

    df = pd.DataFrame(results_long_table(sim))
    df["dot_name"] = df["dot_name"].astype("category")
    if summary_config is not None:
        df["date"] = pd.to_datetime(df["date"])
        if "time_periods" in summary_config:
            df = add_temporal_groupings(df, summary_config["time_periods"])
        groupings, level = summary_config.get("region_groupings"), summary_config.get("grouping_level")
        if groupings is not None:
            df = add_regional_groupings(df, groupings, level if level is not None else "adm0")
        else:  # the reference always adds the admin columns; without a level the region is the node itself
            df = add_regional_groupings(df, grouping_level=level if level is not None else "dot_name")
    path = Path(filename)
    if path.suffix == ".h5":
        try:
            out = df.copy()
            out["date"] = pd.to_datetime(out["date"])
            out.to_hdf(path, key="results", mode="w", format="table", complevel=5)
            return df
        except ImportError:  # PyTables is not installed: keep the data, as CSV
            path = path.with_suffix(".csv")
    df.to_csv(path, index=False)
    return df
