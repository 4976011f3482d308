"""Minimal re-provision of the ``laser-core`` (~=0.6) surface the hot path sits on.

The reference imports these from the third-party ``laser_core`` package
(reference model.py:13-23, pars.py:4, run_sim.py:11-12), whose source is not part
of the reference checkout and which is not installable offline.  Only the pieces
the per-tick path and its constructors touch are provided, restated from the
published behaviour and from the reference's own call sites:

* ``PropertySet``   attribute/item bag with ``+=`` (add), ``<<=`` (override), ``|=`` (both)
* ``LaserFrame``    structure-of-arrays agent table: ``count``/``capacity``, ``add_*_property``, ``add(n)``
* ``seed``          seeds numpy's global stream (host-side draws) -- reference model.py:82
* ``calc_capacity`` population * (1 + cbr/1000/365) ** nticks -- reference model.py:129-131
* migration: ``gravity``, ``radiation``, ``row_normalizer``, ``distance`` -- reference model.py:1233-1258
* demographics: ``KaplanMeierEstimator``, ``AliasedDistribution``, ``load_pyramid_csv`` -- model.py:1558-1611, 1725

Exact laser-core numerics for radiation / KaplanMeier are *unpinned* (no source,
no golden vectors in the reference); the reference's tests only pin identities
(deaths land on ``date_of_death`` days, population bookkeeping), which tests/ reproduce.

``LaserFrame`` columns are page-locked (pinned) numpy arrays when CUDA is available,
so the H2D / D2H synchronisation points of ``SEIR_ABM.run()`` run at full PCIe rate.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "PropertySet", "LaserFrame", "seed", "calc_capacity", "gravity", "radiation", "row_normalizer", "distance",
    "KaplanMeierEstimator", "AliasedDistribution", "load_pyramid_csv",
]


# --------------------------------------------------------------------------- PropertySet
class PropertySet:
    def __init__(self, *bags):
        for bag in bags:
            items = bag.items() if hasattr(bag, "items") else bag
            for k, v in items:
                setattr(self, k, v)

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, PropertySet) else v) for k, v in self.__dict__.items()}

    def items(self):
        return self.__dict__.items()

    def keys(self):
        return self.__dict__.keys()

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    def __contains__(self, key):
        return key in self.__dict__

    def __len__(self):
        return len(self.__dict__)

    def __iter__(self):
        return iter(self.__dict__)

    def __eq__(self, other):
        return isinstance(other, PropertySet) and self.to_dict() == other.to_dict()

    def __repr__(self):
        return f"PropertySet({self.to_dict()!r})"

    @staticmethod
    def _items(other):
        return other.items() if hasattr(other, "items") else dict(other).items()

    def __iadd__(self, other):  # add NEW keys only
        for k, v in self._items(other):
            if k in self.__dict__:
                raise ValueError(f"PropertySet +=: key '{k}' already exists")
            self.__dict__[k] = v
        return self

    def __add__(self, other):
        out = PropertySet(self)
        out += other
        return out

    def __ilshift__(self, other):  # override EXISTING keys only
        for k, v in self._items(other):
            if k not in self.__dict__:
                raise ValueError(f"PropertySet <<=: key '{k}' does not exist")
            self.__dict__[k] = v
        return self

    def __lshift__(self, other):
        out = PropertySet(self)
        out <<= other
        return out

    def __ior__(self, other):  # add or override
        for k, v in self._items(other):
            self.__dict__[k] = v
        return self

    def __or__(self, other):
        out = PropertySet(self)
        out |= other
        return out


# --------------------------------------------------------------------------- LaserFrame
def _alloc(shape, dtype, fill):
    """numpy array, page-locked when a CUDA device is present (falls back to pageable memory)."""
    arr = None
    try:
        import torch

        if torch.cuda.is_available():
            tdt = getattr(torch, np.dtype(dtype).name, None)
            if tdt is not None:
                arr = torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    except Exception:  # noqa: BLE001 - pinning is an optimisation only
        arr = None
    if arr is None:
        arr = np.empty(shape, dtype=dtype)
    arr[...] = fill
    return arr


class LaserFrame:
    def __init__(self, capacity: int, initial_count: int = -1, **kwargs):
        capacity = int(capacity)
        if capacity <= 0:
            raise ValueError(f"Capacity must be positive, got {capacity}")
        initial_count = capacity if initial_count == -1 else int(initial_count)
        if not (0 <= initial_count <= capacity):
            raise ValueError(f"Initial count {initial_count} must be in [0, capacity={capacity}]")
        self._capacity = capacity
        self._count = initial_count
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def count(self) -> int:
        return self._count

    @property
    def capacity(self) -> int:
        return self._capacity

    def __len__(self) -> int:
        return self._count

    def add_scalar_property(self, name: str, dtype=np.uint32, default=0) -> None:
        setattr(self, name, _alloc(self._capacity, dtype, default))

    def add_vector_property(self, name: str, length: int, dtype=np.uint32, default=0) -> None:
        setattr(self, name, _alloc((int(length), self._capacity), dtype, default))

    def add_array_property(self, name: str, shape, dtype=np.uint32, default=0) -> None:
        setattr(self, name, _alloc(tuple(shape), dtype, default))

    def add(self, count: int):
        count = int(count)
        if self._count + count > self._capacity:
            raise ValueError(f"frame.add() exceeds capacity ({self._count=} + {count=} > {self._capacity=})")
        start = self._count
        self._count += count
        return start, self._count

    def columns(self) -> dict:
        """name -> 1-D per-agent numpy column (length == capacity)."""
        return {k: v for k, v in self.__dict__.items()
                if isinstance(v, np.ndarray) and v.ndim == 1 and v.shape[0] == self._capacity and not k.startswith("_")}


# --------------------------------------------------------------------------- misc
def seed(value: int):
    """Seed the host-side global numpy stream (reference model.py:82 -> laser_core.random.seed)."""
    np.random.seed(int(value) & 0xFFFFFFFF)
    return np.random.default_rng(int(value))


def calc_capacity(population, nticks, cbr, verbose: bool = False):
    daily_rate = (cbr / 1000.0) / 365.0
    return np.uint64(population * (1.0 + daily_rate) ** nticks)


# --------------------------------------------------------------------------- migration
def distance(lat1, lon1, lat2, lon2):
    """Haversine great-circle distance in km (scalars or broadcastable arrays)."""
    lat1, lon1, lat2, lon2 = (np.radians(np.asarray(x, dtype=np.float64)) for x in (lat1, lon1, lat2, lon2))
    a = np.sin((lat2 - lat1) / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2) ** 2
    return 6371.0 * 2 * np.arcsin(np.sqrt(a))


def gravity(pops, distances, k, a, b, c, **kwargs):
    """network[i, j] = k * p_i^a * p_j^b / d_ij^c, zero diagonal."""
    pops = np.asarray(pops, dtype=np.float64)
    d = np.array(distances, dtype=np.float64, copy=True)
    np.fill_diagonal(d, 1.0)
    net = k * (pops[:, None] ** a) * (pops[None, :] ** b) * (d ** (-float(c)))
    np.fill_diagonal(net, 0.0)
    return net


def radiation(pops, distances, k, include_home, **kwargs):
    """Radiation model: T_ij = k * p_i p_j / ((p_i + s_ij)(p_i + p_j + s_ij)), s_ij = population within d_ij of i
    (excluding i and j; including i when ``include_home``)."""
    pops = np.asarray(pops, dtype=np.float64)
    d = np.asarray(distances, dtype=np.float64)
    n = len(pops)
    net = np.zeros((n, n), dtype=np.float64)
    order = np.argsort(d, axis=1, kind="stable")
    for i in range(n):
        idx = order[i]
        sorted_pops = pops[idx]
        cum = np.cumsum(sorted_pops)
        # ties in distance share the same radius: use the cumulative sum up to the last node at that distance
        dist_sorted = d[i, idx]
        last = np.searchsorted(dist_sorted, dist_sorted, side="right") - 1
        within = cum[last] - sorted_pops  # population within the radius, excluding j itself
        if not include_home:
            within = within - pops[i]
        s = np.maximum(within, 0.0)
        pi = pops[i]
        val = k * pi * sorted_pops / ((pi + s) * (pi + sorted_pops + s))
        net[i, idx] = val
        net[i, i] = 0.0
    return net


def row_normalizer(network, max_rowsum):
    """Rescale only the rows whose sum exceeds ``max_rowsum`` so that they sum to it."""
    net = np.array(network, dtype=np.float64, copy=True)
    rs = net.sum(axis=1)
    over = rs > max_rowsum
    net[over] *= (max_rowsum / rs[over])[:, None]
    return net


# --------------------------------------------------------------------------- demographics
class AliasedDistribution:
    """Categorical sampler proportional to ``counts`` (inverse-CDF form; same distribution as the alias method)."""

    def __init__(self, counts):
        c = np.asarray(counts, dtype=np.float64)
        self._cdf = np.cumsum(c)
        self.total = float(self._cdf[-1])

    def sample(self, count: int = 1):
        u = np.random.random(int(count)) * self.total
        return np.minimum(np.searchsorted(self._cdf, u, side="right"), len(self._cdf) - 1).astype(np.int32)


def load_pyramid_csv(path):
    """Rows 'lo-hi,M,F' (last row 'lo+,M,F') -> array [[lo, hi, M, F]]."""
    rows = []
    with open(path) as fh:
        header = fh.readline()
        if not header.lower().startswith("age"):
            raise ValueError(f"{path}: expected header 'Age,M,F'")
        for line in fh:
            line = line.strip()
            if not line:
                continue
            age, m, f = line.split(",")[:3]
            if "+" in age:
                lo = hi = int(age.replace("+", ""))
            else:
                lo, hi = (int(x) for x in age.split("-"))
            rows.append((lo, hi, int(m), int(f)))
    return np.array(rows, dtype=np.int32)


class KaplanMeierEstimator:
    """Inverse-CDF draw of age at death from a cumulative-deaths-by-year table, conditional on current age."""

    def __init__(self, source):
        self._cd = np.insert(np.asarray(source, dtype=np.int64), 0, 0)  # cd[y] = deaths before age y

    def predict_year_of_death(self, ages_years, max_year: int = 100):
        ages_years = np.minimum(np.asarray(ages_years, dtype=np.int64), max_year)
        total = self._cd[max_year + 1]
        already = self._cd[ages_years]
        draw = already + 1 + np.floor(np.random.random(ages_years.shape) * np.maximum(total - already, 1)).astype(np.int64)
        yod = np.searchsorted(self._cd, draw, side="left") - 1
        return np.clip(yod, ages_years, max_year)

    def predict_age_at_death(self, ages_days, max_year: int = 100):
        ages_days = np.asarray(ages_days, dtype=np.int64)
        age_years = ages_days // 365
        yod = self.predict_year_of_death(age_years, max_year)
        u = np.random.random(ages_days.shape)
        doy_rest = ages_days % 365
        # same year as the current age: a later day of that year; otherwise any day of the year
        same = yod == np.minimum(age_years, max_year)
        doy = np.where(same, doy_rest + 1 + np.floor(u * np.maximum(364 - doy_rest, 1)), np.floor(u * 365)).astype(np.int64)
        return (yod * 365 + doy).astype(np.int32)
