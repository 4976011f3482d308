#!/usr/bin/env python
"""The reference's OWN numba kernels as the timed CPU baseline (SURVEY 8d "CPU baseline timing").

Runs only where /root/reference is reachable (the build container): the hot-path functions are AST-loaded read-only by
oracle/ref_loader.py and called in the reference's order for a daily tick -- get_deaths* + disease_state_step + fast_ri* +
fast_sia* + tx_step_prep + node block (restated: it is a method body) + tx_infect_nb + count_SEIRP (* on their schedule) --
on the synthetic Nigeria-shape table, 1 JIT warm-up tick, then the timed ticks.  Prints one JSON line; the number is
recorded in BASELINE.md with the core count (the GPU box has no reference checkout, so bench.py's CPU leg there is the
C + OpenMP port of the same kernels, oracle/lp_oracle.c).

    python tools/numba_reference_leg.py [agents] [ticks]
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numba as nb  # noqa: E402

import laser_polio_b200.synth as synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import ref_loader  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 14
nodes, ns, seed = min(774, max(14, n // 20_000)), 3, 5
ref = ref_loader.load(inject_uniforms=False)
p = synth.synth_population(n, nodes, seed=seed)
rng = np.random.default_rng(seed)
W = rng.random((nodes, nodes)) * (0.1 / nodes)
np.fill_diagonal(W, 0.0)
srs = np.array([1.0, 0.25, 0.125])
r0s = rng.uniform(0.5, 1.5, nodes)
pr, pi = rng.uniform(0.3, 0.8, nodes), rng.uniform(0.3, 0.8, nodes)
vx = rng.uniform(0.4, 0.9, nodes).astype(np.float32)
targeted = (rng.random(nodes) < 0.6).astype(np.uint8)
pop = np.bincount(p["node_id"][:n], minlength=nodes).astype(np.int32)
si, sp = np.zeros(n, np.int32), np.zeros(n, np.float32)
nt = nb.get_num_threads()
stage = {}


def timed(name, fn, on):
    t0 = time.perf_counter()
    out = fn()
    if on:
        stage[name] = stage.get(name, 0.0) + time.perf_counter() - t0
    return out


total = 0.0
first = 14 - 1  # warm-up tick 13, then ticks 14 .. : a vital-dynamics + RI tick is inside the window
for k in range(1 + ticks):
    t, on = first + k, k >= 1
    t0 = time.perf_counter()
    if t % 7 == 0 or k == 0:
        tl, dying = np.zeros((nt, nodes), np.int32), np.zeros(nodes, np.int32)
        timed("get_deaths", lambda: ref["get_deaths"](np.int32(nodes), np.int32(n), p["disease_state"], p["node_id"], p["date_of_death"], np.int32(t), tl, dying), on)
    a, b = np.zeros(nodes, np.int32), np.zeros(nodes, np.int32)
    timed("disease_state_step", lambda: ref["disease_state_step"](p["node_id"], nodes, p["disease_state"], p["strain"], n, p["exposure_timer"],
                                                                    p["infection_timer"], p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"],
                                                                    p["paralysis_timer"], nb.float32(1 / 2000), a, b), on)
    if t % 14 == 0 or k == 0:
        l1, l2, l3 = (np.zeros((nt, nodes), np.int32) for _ in range(3))
        timed("fast_ri", lambda: ref["fast_ri"](14, p["node_id"], p["disease_state"], p["strain"], p["ipv_protected"], p["ri_timer"], t, pr, pi, n,
                                                l1, l2, l3, p["chronically_missed"], np.int8(1)), on)
    if t % 44 == 20 or k <= 1:
        l1, l2 = np.zeros((nt, nodes), np.int32), np.zeros((nt, nodes), np.int32)
        timed("fast_sia", lambda: ref["fast_sia"](p["node_id"], p["disease_state"], p["strain"], p["date_of_birth"], t, vx, 0.56, n, targeted, 0, 5 * 365,
                                                  l1, l2, p["chronically_missed"], np.int8(2)), on)
    beta, expo, sus = timed("tx_step_prep", lambda: ref["tx_step_prep"](nodes, n, ns, p["strain"][:n], srs, p["disease_state"][:n], p["node_id"][:n],
                                                                          p["daily_infectivity"][:n], p["acq_risk_multiplier"][:n]), on)
    beta_pre, prob = timed("node_math", lambda: orc.tx_foi(beta, W, 1.05, r0s, pop), on)
    want, _ = timed("node_math", lambda: orc.tx_draw_counts_ref(beta_pre, prob, expo, 0.0, 1000, rs=np.random), on)
    timed("tx_infect_nb", lambda: ref["tx_infect_nb"](nodes, n, ns, sus, p["node_id"][:n], p["strain"][:n], p["disease_state"][:n], si, sp,
                                                      p["acq_risk_multiplier"][:n], prob, want), on)
    timed("count_SEIRP", lambda: ref["count_SEIRP"](p["node_id"], p["disease_state"], p["strain"], p["potentially_paralyzed"], p["paralyzed"], nodes, ns, n), on)
    if on:
        total += time.perf_counter() - t0
print(json.dumps({"impl": "reference numba kernels (AST-loaded, unmodified)", "agent_days_per_s": n * ticks / total, "agents": n, "nodes": nodes,
                  "ticks": ticks, "numba_threads": nt, "cpu_count": os.cpu_count(), "seconds": round(total, 3),
                  "stage_seconds": {k: round(v, 3) for k, v in stage.items()}}))
