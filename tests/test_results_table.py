"""CPU: the long results table (reference utils.py:690-786) from a stand-in sim."""

import types

import numpy as np

import laser_polio_b200 as lp


def test_long_table_layout_and_groupings(tmp_path):
    nt, nodes = 5, 3
    rng = np.random.default_rng(0)
    res = types.SimpleNamespace(**{k: rng.integers(0, 100, (nt, nodes)).astype(np.int32) for k in
                                   ("S", "E", "I", "R", "paralyzed", "births", "deaths", "new_exposed", "potentially_paralyzed",
                                    "new_potentially_paralyzed", "new_paralyzed")})
    pars = lp.PropertySet({"node_lookup": {0: {"dot_name": "A:B:C"}, 2: {"dot_name": "A:B:D"}}})
    sim = types.SimpleNamespace(nt=nt, nodes=np.arange(nodes), results=res, pars=pars, datevec=lp.daterange(lp.date("2020-06-29"), nt))
    df = lp.save_sim_results(sim, tmp_path / "simulation_results.h5",
                             summary_config={"time_periods": {"bins": ["2020-07-01"], "labels": ["early", "late"]}})
    assert len(df) == nt * nodes
    assert list(df.columns[:4]) == ["timestep", "date", "node", "dot_name"] and "P" in df.columns
    row = df[(df.timestep == 3) & (df.node == 2)].iloc[0]  # time-major: row t * nodes + n
    assert row.name == 3 * nodes + 2 and row.S == res.S[3, 2] and row.P == res.paralyzed[3, 2] and row.dot_name == "A:B:D"
    assert df[df.node == 1].dot_name.iloc[0] == "UNKNOWN"
    assert list(df[df.node == 0].time_period) == ["early", "early", "late", "late", "late"]  # left-closed at 2020-07-01
    assert (tmp_path / "simulation_results.h5").exists() or (tmp_path / "simulation_results.csv").exists()
