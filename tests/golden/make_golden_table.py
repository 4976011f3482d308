"""Golden long results table from the REFERENCE's own save_sim_results / add_temporal_groupings / add_regional_groupings
(utils.py:690-786, 1040-1159), AST-loaded read-only from /root/reference (build container only) and run on a stand-in sim
with seeded results arrays.  tests/test_results_table.py rebuilds the same stand-in and holds lp.save_sim_results to the
stored table value for value (and, where the reference checkout is present, to the reference function run live).

    python tests/golden/make_golden_table.py
"""

import ast
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd
import yaml

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF_UTILS = Path("/root/reference/src/laser_polio/utils.py")
OUT = Path(__file__).resolve().parent

REGIONS_YAML = {"NIGERIA": {"NW_NGA": ["NIGERIA:JIGAWA", "NIGERIA:KANO"], "S_NGA": ["NIGERIA:LAGOS"]}, "BENIN": {"COAST": ["ATLANTIQUE", "littoral"]}}
DOT_NAMES = ["AFRO:NIGERIA:JIGAWA:AUYO", "AFRO:NIGERIA:KANO:DALA", "AFRO:NIGERIA:LAGOS:IKEJA", "AFRO:NIGERIA:BORNO:BAMA",
             "AFRO:BENIN:ATLANTIQUE:ABOMEY_CALAVI", "AFRO:BENIN:LITTORAL:COTONOU", "AFRO:NIGER:MARADI:TESSAOUA"]
SUMMARY = {"time_periods": {"bins": ["2020-07-02", "2020-07-05"], "labels": ["a", "b", "c"]}, "region_groupings": ["NIGERIA", "BENIN", "TOGO"],
           "grouping_level": "adm01"}
COLUMNS = ("S", "E", "I", "R", "paralyzed", "births", "deaths", "new_exposed", "potentially_paralyzed", "new_potentially_paralyzed",
           "new_paralyzed")


def stand_in_sim(lp):
    nt, nodes = 8, len(DOT_NAMES)
    rng = np.random.default_rng(42)
    res = types.SimpleNamespace(**{k: rng.integers(0, 1000, (nt, nodes)).astype(np.int32) for k in COLUMNS})
    pars = lp.PropertySet({"node_lookup": {n: {"dot_name": d} for n, d in enumerate(DOT_NAMES)}})
    return types.SimpleNamespace(nt=nt, nodes=np.arange(nodes), results=res, pars=pars, datevec=lp.daterange(lp.date("2020-06-29"), nt))


def reference_functions(data_root):
    """save_sim_results & co. of the reference, exec'd with the module globals they use (``lp.root`` -> data_root)."""
    src = REF_UTILS.read_text()
    tree = ast.parse(src)
    ns = {"np": np, "pd": pd, "yaml": yaml, "Path": Path, "lp": types.SimpleNamespace(root=Path(data_root))}
    lines = src.splitlines()
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("save_sim_results", "add_temporal_groupings", "add_regional_groupings"):
            exec(compile("\n".join(lines[node.lineno - 1:node.end_lineno]), f"<reference:{node.name}>", "exec"), ns)  # noqa: S102
    return ns


def write_regions(data_root):
    (Path(data_root) / "data").mkdir(parents=True, exist_ok=True)
    (Path(data_root) / "data" / "regions.yaml").write_text(yaml.safe_dump(REGIONS_YAML))


if __name__ == "__main__":
    import tempfile

    import laser_polio_b200 as lp

    if not REF_UTILS.exists():
        raise SystemExit("needs the reference checkout at /root/reference")
    with tempfile.TemporaryDirectory() as tmp:
        write_regions(tmp)
        df = reference_functions(tmp)["save_sim_results"](stand_in_sim(lp), str(Path(tmp) / "out.csv"), summary_config=SUMMARY)
    df.to_csv(OUT / "results_table_ref.csv", index=False)
    print("wrote results_table_ref.csv", df.shape, list(df.columns))
