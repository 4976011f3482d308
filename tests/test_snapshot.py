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
