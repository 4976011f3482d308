"""CPU: LaserFrame.save_snapshot / load_snapshot (reference run_sim.py:388-409, 431, 457 and the round trip
of the reference's tests/test_init_pop.py:171-175)."""

import numpy as np

import laser_polio_b200 as lp


def make_frame():
    f = lp.LaserFrame(capacity=1000, initial_count=600)
    rng = np.random.default_rng(4)
    f.add_scalar_property("disease_state", dtype=np.int8, default=-1)
    f.add_scalar_property("node_id", dtype=np.int16, default=-1)
    f.add_scalar_property("acq_risk_multiplier", dtype=np.float32, default=1.0)
    f.add_scalar_property("date_of_death", dtype=np.int32, default=0)
    f.disease_state[:600] = rng.integers(0, 4, 600)
    f.node_id[:600] = np.repeat(np.arange(3), 200)
    f.acq_risk_multiplier[:600] = rng.random(600)
    f.date_of_death[:600] = rng.integers(1, 30000, 600)
    return f


def test_round_trip(tmp_path):
    f = make_frame()
    pars = lp.PropertySet({"r0": 14.0, "seed": 3, "init_pop": np.array([200, 200, 200]), "cbr": np.array([30.0, 30.0, 30.0]), "dur": 365,
                           "dur_inf": lp.gamma(shape=4.51, scale=5.32), "node_lookup": {0: {"lat": 1.0}}})
    R = np.arange(12, dtype=np.int32).reshape(4, 3)
    for name in ("init_pop.h5", "init_pop.npz"):  # the reference's file name is honoured whatever the container
        path = tmp_path / name
        f.save_snapshot(path, R, pars)
        g, R2, p2 = lp.LaserFrame.load_snapshot(path, n_ppl=pars.init_pop, cbr=pars.cbr, nt=pars.dur + 10)
        births = float(lp.calc_capacity(600, 375, 30.0)) - 600  # room for the births of the run, with the reference's margin
        assert g.count == 600 and g.capacity == int((1 + 4 / np.sqrt(births)) * (600 + births)) and 17 < births < 20
        for col in ("disease_state", "node_id", "acq_risk_multiplier", "date_of_death"):
            a, b = getattr(f, col), getattr(g, col)
            assert a.dtype == b.dtype and np.array_equal(a[:600], b[:600]), col
        assert np.array_equal(R2, R)
        assert p2["r0"] == 14.0 and p2["seed"] == 3 and p2["init_pop"] == [200, 200, 200] and "dur_inf" not in p2 and "node_lookup" not in p2
    g, R2, p2 = lp.LaserFrame.load_snapshot(tmp_path / "init_pop.npz")
    assert g.capacity == g.count == 600
    f.save_snapshot(tmp_path / "bare.h5")
    g, R2, p2 = lp.LaserFrame.load_snapshot(tmp_path / "bare.h5")
    assert R2 is None and p2 is None and g.count == 600


def test_every_people_column_round_trips_through_a_full_sim(tmp_path):
    """The reference's tests/test_init_pop.py:171-175 on the CPU: a sim built with every stock component, its people frame
    saved and reloaded, every column identical over the live prefix (construction is host-side; nothing runs)."""
    rows = ["Age,M,F"] + [f"{5 * k}-{5 * k + 4},{int(1.7e7 * np.exp(-0.16 * k))},{int(1.6e7 * np.exp(-0.16 * k))}" for k in range(20)]
    (tmp_path / "pyramid.csv").write_text("\n".join(rows + ["100+,300,500"]) + "\n")
    pars = lp.PropertySet({
        "start_date": lp.date("2020-01-01"), "dur": 40, "init_pop": np.array([3000, 2000, 1000]), "cbr": np.array([35.0, 30.0, 25.0]),
        "r0_scalars": np.ones(3), "age_pyramid_path": str(tmp_path / "pyramid.csv"), "init_immun": 0.2, "init_prev": 0.01, "seed": 5,
        "vx_prob_ri": 0.3, "vx_prob_ipv": 0.3, "distances": np.array([[0, 40, 70], [40, 0, 50], [70, 50, 0.0]]), "verbose": 0,
        "stop_if_no_cases": False, "vx_prob_sia": [0.5, 0.5, 0.5],
        "sia_schedule": [{"date": "2020-01-10", "nodes": [0, 1], "age_range": (0, 1825), "vaccinetype": "nOPV2"}]})
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    path = tmp_path / "init_pop.h5"
    sim.people.save_snapshot(path, sim.results.R[:], sim.pars)
    people, R, loaded = lp.LaserFrame.load_snapshot(path, n_ppl=pars["init_pop"], cbr=pars["cbr"], nt=pars["dur"] + 10)
    n = sim.people.count
    assert people.count == n and set(people.columns()) == set(sim.people.columns()) and len(people.columns()) >= 15
    for name, col in sim.people.columns().items():
        got = getattr(people, name)
        assert got.dtype == col.dtype and np.array_equal(got[:n], col[:n]), name
    assert np.array_equal(R, sim.results.R) and loaded["r0"] == sim.pars.r0
