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


def _golden_module():
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("make_golden_table", Path(__file__).resolve().parent / "golden" / "make_golden_table.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_long_table_equals_the_references_table(tmp_path):
    """save_sim_results with temporal + regional groupings (region_groupings at adm01, a dot_name pattern group, a country
    missing from regions.yaml) against the table the REFERENCE's own function produced for the same stand-in sim
    (tests/golden/results_table_ref.csv, made by tests/golden/make_golden_table.py); where the reference checkout is present
    the reference function is also run live and compared frame to frame, dtypes included."""
    import pandas as pd
    from pandas.testing import assert_frame_equal

    from laser_polio_b200 import utils

    g = _golden_module()
    g.write_regions(tmp_path)
    utils.root = tmp_path
    try:
        ours = lp.save_sim_results(g.stand_in_sim(lp), tmp_path / "ours.csv", summary_config=g.SUMMARY)
    finally:
        utils.root = None
    ref = pd.read_csv(g.OUT / "results_table_ref.csv", parse_dates=["date"])
    assert list(ours.columns) == list(ref.columns)
    assert len(ours) == len(ref)
    for col in ref.columns:
        a = ours[col].astype(str) if col in ("dot_name", "time_period", "adm0", "adm1", "adm01", "region") else ours[col]
        b = ref[col].astype(str) if a.dtype == object else ref[col]
        assert (a.to_numpy() == b.to_numpy()).all(), col
    assert set(ours.region) == {"NW_NGA", "S_NGA", "NIGERIA:BORNO", "COAST", "NIGER:MARADI"}
    written = pd.read_csv(tmp_path / "ours.csv")
    assert len(written) == len(ours) and list(written.columns) == list(ours.columns)
    if g.REF_UTILS.exists():  # build container: the reference function itself, live
        live = g.reference_functions(tmp_path)["save_sim_results"](g.stand_in_sim(lp), str(tmp_path / "ref.csv"), summary_config=g.SUMMARY)
        assert_frame_equal(ours, live, check_dtype=True, check_categorical=True)


def test_regional_groupings_defaults_and_errors(tmp_path):
    import pandas as pd
    import pytest

    df = pd.DataFrame({"dot_name": ["AFRO:NIGERIA:KANO:DALA", "AFRO:BENIN:LITTORAL:COTONOU"]})
    out = lp.add_regional_groupings(df)
    assert list(out.region) == ["NIGERIA", "BENIN"] and list(out.adm01) == ["NIGERIA:KANO", "BENIN:LITTORAL"]
    assert list(lp.add_regional_groupings(df, grouping_level="dot_name").region) == list(df.dot_name)
    with pytest.raises(ValueError):
        lp.add_regional_groupings(df, grouping_level="adm2")
    # no regions.yaml: warns and keeps the admin level
    assert list(lp.add_regional_groupings(df, ["NIGERIA"], regions_yaml_path=tmp_path / "missing.yaml").region) == ["NIGERIA", "BENIN"]
