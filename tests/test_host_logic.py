"""CPU-only tests: host-side mirror of the reference interface, and that the C-ABI library loads and exports
every symbol include/lpk.h declares (no compute calls: there is no GPU here)."""

import ctypes
import re
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lp():
    import laser_polio_b200 as lp

    return lp


@pytest.fixture(scope="module")
def pyramid(tmp_path_factory):
    path = tmp_path_factory.mktemp("data") / "pyramid.csv"
    rows = ["Age,M,F"] + [f"{5 * k}-{5 * k + 4},{int(1.7e7 * np.exp(-0.16 * k))},{int(1.6e7 * np.exp(-0.16 * k))}" for k in range(20)]
    rows.append("100+,300,500")
    path.write_text("\n".join(rows) + "\n")
    return str(path)


def test_c_abi_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    header = (ROOT / "include" / "lpk.h").read_text()
    declared = set(re.findall(r"\b(lpk_[a-z0-9_]+)\s*\(", header))
    assert {"lpk_get_deaths", "lpk_disease_state_step", "lpk_fast_ri", "lpk_fast_sia", "lpk_tx_step_prep", "lpk_tx_node_math",
            "lpk_tx_infect", "lpk_count_seirp"} <= declared
    lib = ctypes.CDLL(str(ROOT / "laser-polio_b200" / "liblpk.so"))
    for name in declared:
        assert hasattr(lib, name), name
    lib.lpk_version.restype = ctypes.c_int
    assert lib.lpk_version() >= 1
    from laser_polio_b200 import _lpk

    assert set(_lpk.EXPORTS) <= declared


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "laser-polio_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    for path in (ROOT / "laser-polio_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path


def test_run_fails_loudly_without_a_gpu(lp, pyramid):
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    sim = lp.SEIR_ABM(lp.PropertySet({"init_pop": np.array([50, 50]), "age_pyramid_path": pyramid, "verbose": 0, "seed": 1}))
    sim.components = [lp.DiseaseState_ABM, lp.Transmission_ABM]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.run()


def test_propertyset_operators(lp):
    p = lp.PropertySet({"a": 1, "b": 2})
    assert p.a == 1 and p["b"] == 2 and "a" in p and len(p) == 2 and p.to_dict() == {"a": 1, "b": 2}
    p += {"c": 3}  # tests/test_interventions.py:32
    assert p.c == 3
    with pytest.raises(ValueError):
        p += {"a": 9}
    p <<= {"a": 5}  # reference model.py:69-71
    assert p.a == 5
    with pytest.raises(ValueError):
        p <<= {"zzz": 1}
    p |= {"a": 6, "d": 4}
    assert p.a == 6 and p.d == 4


def test_laserframe_semantics(lp):
    f = lp.LaserFrame(capacity=100, initial_count=10)
    assert (f.count, f.capacity, len(f)) == (10, 100, 10)
    f.add_scalar_property("x", dtype=np.int8, default=-1)
    f.add_array_property("r", shape=(3, 2), dtype=np.int32)
    assert f.x.shape == (100,) and f.x.dtype == np.int8 and np.all(f.x == -1) and f.r.shape == (3, 2)
    assert f.add(5) == (10, 15) and f.count == 15
    with pytest.raises(ValueError):
        f.add(1000)
    assert set(f.columns()) == {"x"}


def test_seir_abm_init_matches_reference_contract(lp, pyramid):
    """tests/test_seir_abm_init.py:22-48"""
    pars = lp.PropertySet({"init_pop": np.array([1000, 500, 20]), "cbr": np.array([30, 25, 10]), "r0_scalars": np.ones(3),
                           "age_pyramid_path": pyramid, "init_prev": np.array([0.0, 0.1, 5]), "init_immun": 0.2,
                           "distances": np.ones((3, 3)) - np.eye(3), "verbose": 0, "seed": 3, "missed_frac": 0.1,
                           "vx_prob_ri": 0.5, "not_a_par": 1})
    sim = lp.SEIR_ABM(pars)
    assert sim.people.count == 1520 == len(sim.people) and sim.people.capacity > sim.people.count
    assert np.array_equal(np.bincount(sim.people.node_id[:1520]), [1000, 500, 20])
    assert np.all(np.diff(sim.people.node_id[:1520]) >= 0)  # initial population is node-contiguous
    assert np.all(sim.people.disease_state[1520:] == -1) and np.all(sim.people.node_id[1520:] == -1)
    assert sim.people.chronically_missed.sum() == 152
    assert "not_a_par" not in sim.pars
    sim.components = [lp.Transmission_ABM, lp.SIA_ABM, lp.RI_ABM, lp.DiseaseState_ABM, lp.VitalDynamics_ABM]
    assert [type(i).__name__ for i in sim.instances] == lp.default_run_order
    assert sim.results.S.shape == (31, 3) and sim.results.E_by_strain.shape == (31, 3, 3) and sim.results.S.dtype == np.int32
    assert np.array_equal(sim.results.pop[0], [1000, 500, 20])
    st = sim.people.disease_state[:1520]
    assert np.sum(st[1000:1500] == 2) == 50 and np.sum(st[1500:] == 2) == 5 and np.sum(st[:1000] == 2) == 0
    for col, dt in (("exposure_timer", np.int8), ("infection_timer", np.int8), ("paralysis_timer", np.int8),
                    ("acq_risk_multiplier", np.float32), ("daily_infectivity", np.float32), ("date_of_birth", np.int32),
                    ("date_of_death", np.int32), ("ri_timer", np.int16), ("strain", np.int8), ("node_id", np.int16)):
        assert getattr(sim.people, col).dtype == dt, col
    p = sim.people
    assert np.all(p.paralysis_timer <= p.infection_timer) and np.all(p.paralysis_timer >= 0)
    assert abs(p.acq_risk_multiplier.mean() - 1.0) < 0.2 and abs(p.daily_infectivity.mean() - 14 / 24) < 0.1
    assert np.all(p.date_of_birth[:1520] < 0) and np.all(p.date_of_death[:1520] >= 0)
    age = -p.date_of_birth[:1520]
    young = age < 30000  # the int16 column wraps for the very old, in the reference as well (model.py:1891-1894)
    assert np.all((p.ri_timer[:1520][young] + age[young] >= 42) & (p.ri_timer[:1520][young] + age[young] <= 98))
    tx = sim.instances[-1]
    assert tx.network.shape == (3, 3) and np.all(tx.network.sum(axis=1) <= sim.pars.max_migr_frac + 1e-12)


def test_bad_parameters_raise_like_the_reference(lp, pyramid):
    with pytest.raises(ValueError):  # model.py:170
        lp.SEIR_ABM(lp.PropertySet({"init_pop": np.array([10, 0]), "verbose": 0, "seed": 1}))
    sim = lp.SEIR_ABM(lp.PropertySet({"init_pop": np.array([10, 10]), "init_prev": [0.1], "age_pyramid_path": pyramid, "verbose": 0, "seed": 1}))
    with pytest.raises(ValueError):  # model.py:687
        sim.components = [lp.DiseaseState_ABM]
    sim = lp.SEIR_ABM(lp.PropertySet({"init_pop": np.array([10, 10]), "migration_method": "teleport", "age_pyramid_path": pyramid, "verbose": 0, "seed": 1}))
    with pytest.raises(ValueError):  # model.py:1256
        sim.components = [lp.Transmission_ABM]
    with pytest.raises(ValueError):  # distributions.py:39
        lp.Distribution("lognormal_int", mean=1, sigma=1)


def test_seasonality(lp):
    """tests/test_seasonality.py: 1 + A cos(2 pi (doy - peak) / days_in_year), leap-year aware."""
    def season(day, amp=0.2, peak=180):
        sim = SimpleNamespace(t=0, datevec=[lp.date(day)], pars={"seasonal_amplitude": amp, "seasonal_peak_doy": peak})
        return lp.get_seasonality(sim)

    assert abs(season("2021-06-29") - 1.2) < 1e-10  # doy 180
    vals = [season(str(d)) for d in lp.daterange("2021-01-01", 365)]
    assert abs(max(vals) - 1.2) < 1e-10 and abs(min(vals) - 0.8) < 1e-4
    assert abs(season("2021-06-19") - season("2021-07-09")) < 1e-10  # symmetric around the peak
    assert season("2021-03-01", amp=0.0) == 1.0
    assert abs(season("2020-12-31", peak=366) - 1.2) < 1e-10  # leap year: 366 days


def test_migration_and_demographics(lp):
    from laser_polio_b200 import core

    pops = np.array([1000.0, 2000.0, 500.0])
    d = np.array([[0, 10, 20], [10, 0, 5], [20, 5, 0.0]])
    g = core.gravity(pops, d, 2.0, 1, 1, 2.0)
    assert np.isclose(g[0, 1], 2.0 * 1000 * 2000 / 100) and np.all(np.diag(g) == 0)  # k p_i^a p_j^b / d^c
    r = core.radiation(pops, d, 1.0, include_home=False)
    assert np.all(np.diag(r) == 0) and np.all(r >= 0) and r[0, 1] > r[0, 2]
    n = core.row_normalizer(np.array([[0, 0.5, 0.5], [0.01, 0, 0.01], [0.2, 0.2, 0]]), 0.1)
    assert np.allclose(n.sum(axis=1), [0.1, 0.02, 0.1])  # only rows above the cap are rescaled
    assert abs(core.distance(0, 0, 0, 1) - 111.19) < 0.1
    np.random.seed(0)
    km = core.KaplanMeierEstimator(lp.create_cumulative_deaths(100000, 100))
    ages = np.random.randint(0, 80 * 365, 20000)
    aad = km.predict_age_at_death(ages)
    assert np.all(aad > ages) and aad.max() <= 101 * 365
    assert int(core.calc_capacity(1000, 365, 36.5)) == int(1000 * (1 + 0.0001) ** 365)
    s = core.AliasedDistribution(np.array([1, 0, 3])).sample(40000)
    assert set(np.unique(s)) == {0, 2} and abs((s == 2).mean() - 0.75) < 0.02


def test_distributions(lp):
    np.random.seed(1)
    assert np.all(lp.constant(value=4)(10) == 4)
    assert abs(lp.poisson(lam=3)(50000).mean() - 3) < 0.05
    assert abs(lp.gamma(shape=4.51, scale=5.32)(50000).mean() - 4.51 * 5.32) < 0.3
    x = lp.lognormal(mean=12.5, sigma=3.5)(100000)
    assert abs(x.mean() - 12.5) < 0.1 and abs(x.std() - 3.5) < 0.1
    assert abs(lp.normal(mean=3, std=1)(50000).mean() - 3) < 0.02
    u = lp.uniform(min=2, max=10)(10000)
    assert u.min() == 2 and u.max() == 9
    assert lp.exponential(scale=2.0)(10).shape == (10,)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The fused-tick argument structs are filled from Python: their layout must equal the C compiler's."""
    import subprocess

    from laser_polio_b200 import _lpk

    src = tmp_path / "sz.cpp"
    src.write_text(
        '#include <cstdio>\n#include <cstddef>\n#include "lpk.h"\nint main(){'
        'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lpk_rng), sizeof(lpk_people), sizeof(lpk_tick_args), sizeof(lpk_node_args),'
        "offsetof(lpk_tick_args, strain_r0_scalars), offsetof(lpk_tick_args, ri_step), offsetof(lpk_node_args, counts),"
        "offsetof(lpk_people, capacity)); printf(\"%zu %zu\\n\", sizeof(lpk_births_args), offsetof(lpk_births_args, tile_node));}\n"
    )
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/g++", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_lpk.Rng), ctypes.sizeof(_lpk.People), ctypes.sizeof(_lpk.TickArgs), ctypes.sizeof(_lpk.NodeArgs),
            _lpk.TickArgs.strain_r0_scalars.offset, _lpk.TickArgs.ri_step.offset, _lpk.NodeArgs.counts.offset, _lpk.People.capacity.offset,
            ctypes.sizeof(_lpk.BirthsArgs), _lpk.BirthsArgs.tile_node.offset]
    assert got == want
