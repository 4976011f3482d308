"""Definitions shared by tests/golden/make_golden.py --ensemble-full (reference side, build container only) and
tests/test_ensemble_full.py (device / oracle side): the configuration, the starting table and the node-level inputs of the
full-feature ensemble, all pure functions of fixed seeds."""

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import laser_polio_b200.synth as synth  # noqa: E402

AGENT_COLS = list(synth.COLUMNS)

# ---- gate 3, second shape: every stage of the tick on, 60 nodes ---------------------------------------------------------
FULL = {"n_agents": 120_000, "capacity": 126_000, "n_nodes": 60, "n_strains": 3, "ticks": 35, "seeds": 64, "p_paralysis": 0.05,
        "vd_step": 7, "ri_step": 14, "sia_tick": 10, "sia_nodes": 40, "sia_age": (0, 5 * 365), "sia_vaccine": "nOPV2", "sia_eff": 0.7 * 0.8,
        "cbr": 60.0, "zi": 0.3, "disp": 2.0}


def full_table():
    """The table and node-level inputs both sides of the full-feature ensemble start from (pure function of fixed seeds)."""
    c = FULL
    p0 = synth.synth_population(c["n_agents"], c["n_nodes"], seed=177, capacity=c["capacity"], f_exposed=0.0, f_infected=0.0,
                                f_recovered=0.2, f_dead=0.0)
    n = p0["count"]
    rs = np.random.default_rng(15)
    first = np.where((p0["node_id"][:n] < 10) & (p0["disease_state"][:n] == 0))[0]  # infections in nodes 0-9 only: the other 50
    p0["disease_state"][rs.choice(first, 500, replace=False)] = 2                       # are reached through the network
    p0["ipv_protected"][:] = 0
    p0["ipv_protected"][:n] = (rs.random(n) < 0.3).astype(np.int8)
    xy = rs.uniform(0, 600.0, (c["n_nodes"], 2))
    d = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1))
    W = 0.02 * np.exp(-d / 150.0)
    np.fill_diagonal(W, 0.0)
    W *= 0.08 / W.sum(axis=1, keepdims=True)
    node = {
        "network": W, "r0_scalars": rs.uniform(0.8, 1.3, c["n_nodes"]), "vx_prob_ri": rs.uniform(0.4, 0.9, c["n_nodes"]),
        "vx_prob_ipv": rs.uniform(0.4, 0.9, c["n_nodes"]), "vx_prob_sia": rs.uniform(0.4, 0.9, c["n_nodes"]).astype(np.float32),
        "pop0": np.bincount(p0["node_id"][:n], minlength=c["n_nodes"]).astype(np.int32),
    }
    return p0, node


