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

    # ------------------------------------------------------------------ snapshots
    # laser-core's LaserFrame.save_snapshot / load_snapshot as the reference uses them (run_sim.py:388-409, 431, 457): the
    # live prefix [0, count) of every per-agent column, the recovered-by-node results array and the parameters that
    # produced the table; on load the frame gets room for the births of the run to come.  laser-core writes HDF5 through
    # h5py, which is not installed here: the container is HDF5 when h5py can be imported and the path ends in .h5 / .hdf5,
    # otherwise a numpy .npz archive with the same entries (a path given as "init_pop.h5" is honoured as the file name).
    _SNAPSHOT_SCALARS = (int, float, str, bool, np.integer, np.floating, np.bool_)

    def save_snapshot(self, path, results_r=None, pars=None) -> None:
        import json

        cols = {k: np.ascontiguousarray(v[: self._count]) for k, v in self.columns().items()}
        meta = {"count": int(self._count), "capacity": int(self._capacity)}
        if pars is not None:
            src = pars.to_dict() if hasattr(pars, "to_dict") else dict(pars)
            keep = {}
            for k, v in src.items():
                if isinstance(v, self._SNAPSHOT_SCALARS):
                    keep[k] = v.item() if isinstance(v, np.generic) else v
                elif isinstance(v, (list, tuple, np.ndarray)) and np.asarray(v).dtype.kind in "iuf" and np.asarray(v).size <= 100_000:
                    keep[k] = np.asarray(v).tolist()
            meta["pars"] = keep
        path = str(path)
        if path.endswith((".h5", ".hdf5")):
            try:
                import h5py

                # layout as recalled from laser-core ~0.6's LaserFrame.save_snapshot (not in the checkout: unverified):
                # group "people" with attrs count / capacity and one dataset per column (live prefix), dataset "recovered",
                # group "pars" with one attribute per parameter; plus our own "meta" attribute for an exact round trip
                with h5py.File(path, "w") as f:
                    g = f.create_group("people")
                    g.attrs["count"], g.attrs["capacity"] = meta["count"], meta["capacity"]
                    for k, v in cols.items():
                        g.create_dataset(k, data=v)
                    if results_r is not None:
                        f.create_dataset("recovered", data=np.asarray(results_r))
                    if "pars" in meta:
                        pg = f.create_group("pars")
                        for k, v in meta["pars"].items():
                            pg.attrs[k] = v
                    f.attrs["meta"] = json.dumps(meta)
                return
            except ImportError:
                pass
        arrays = {f"people/{k}": v for k, v in cols.items()}
        if results_r is not None:
            arrays["recovered"] = np.asarray(results_r)
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        with open(path, "wb") as fh:  # np.savez would append ".npz" to a path ending in ".h5"
            np.savez(fh, **arrays)

    @classmethod
    def load_snapshot(cls, path, n_ppl=None, cbr=None, nt=None):
        """-> (frame, results_r or None, pars dict or None).  With ``n_ppl`` / ``cbr`` / ``nt`` the capacity is sized for the
        births of an ``nt``-day run (``calc_capacity`` plus the reference's 4 / sqrt(births) safety margin, model.py:124-137);
        otherwise it is the live count."""
        import json

        path = str(path)
        cols, recovered, meta = {}, None, None
        loaded = False
        if path.endswith((".h5", ".hdf5")):
            try:
                import h5py

                if h5py.is_hdf5(path):
                    with h5py.File(path, "r") as f:
                        cols = {k: f["people"][k][...] for k in f["people"]}
                        recovered = f["recovered"][...] if "recovered" in f else None
                        if "meta" in f.attrs:
                            meta = json.loads(f.attrs["meta"])
                        else:  # a file written by laser-core itself
                            meta = {"count": int(f["people"].attrs["count"]), "capacity": int(f["people"].attrs["capacity"])}
                            if "pars" in f:
                                meta["pars"] = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in f["pars"].attrs.items()}
                    loaded = True
            except ImportError:
                pass
        if not loaded:
            with np.load(path, allow_pickle=False) as z:
                cols = {k[len("people/"):]: z[k] for k in z.files if k.startswith("people/")}
                recovered = z["recovered"] if "recovered" in z.files else None
                meta = json.loads(bytes(z["meta"]).decode())
        count = int(meta["count"])
        capacity = count
        if n_ppl is not None and cbr is not None and nt is not None:
            total = float(np.sum(n_ppl))
            rate = float(np.mean(np.atleast_1d(cbr)))
            births = float(calc_capacity(total, int(nt), rate)) - total
            fudge = 1 + 4 / np.sqrt(births) if births > 0 else 1
            capacity = max(count, int(fudge * (count + max(births, 0.0))))
        frame = cls(capacity=capacity, initial_count=count)
        for name, arr in cols.items():
            frame.add_scalar_property(name, dtype=arr.dtype, default=0)
            getattr(frame, name)[:count] = arr
        frame._unborn_from = count  # slots [count, capacity) hold defaults: DiseaseState_ABM.init_from_file draws them (model.py:506-524)
        return frame, recovered, meta.get("pars")

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
