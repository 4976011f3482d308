#!/usr/bin/env python
"""Fixed cost of a fused day: per-kernel CUDA-event times on tables from 1e5 to 1e7 agents (few large nodes, so that every tile
is node-uniform).  usage: tools/diag_fixed_cost.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from laser_polio_b200 import kernels as K  # noqa: E402
from laser_polio_b200 import synth  # noqa: E402

for n, nodes in ((100_000, 4), (1_000_000, 4), (3_000_000, 8), (10_000_000, 32), (10_000_000, 774)):
    sim, _ = synth.synth_sim(n, nodes, 80, seed=1)
    sim.to_device()
    sim.run_ticks(6)
    K.STATS.reset()
    K.STATS.timing = True
    sim.run_ticks(40)
    st = K.STATS.summary()
    K.STATS.timing = False
    p = np.array(K.STATS.times["tick_pass"])
    print(f"{n:>9d} agents {nodes:4d} nodes: pass median {np.median(p) * 1e3:7.1f} us  min {p.min() * 1e3:7.1f} us   node {st['tick_node'][1] * 1e3:6.1f} us")
    sim.to_host()
