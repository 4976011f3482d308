#!/usr/bin/env python
"""Pass time over a long run (newborn cohorts accumulate behind the initial population and run through the general path):
mean tick_pass per 50-tick window.  usage: tools/diag_longrun.py [agents] [ticks]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
import laser_polio_b200 as lp  # noqa: E402
from laser_polio_b200 import kernels as K  # noqa: E402

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 220_000_000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 730
sim = bench.build_sim(lp, agents, 774, ticks + 60, seed=20261017, device="cuda:0")
sim.to_device()
for _ in range(3):
    sim.step_tick(sim.t)
K.STATS.reset()
K.STATS.timing = True
win = 50
for w0 in range(0, ticks, win):
    K.STATS.events = {}
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(win):
        sim.step_tick(sim.t)
    b.record()
    torch.cuda.synchronize()
    ev = K.STATS.events.get("tick_pass", [])
    plain = sorted(x.elapsed_time(y) for x, y in ev)
    print(f"ticks {sim.t - win:4d}-{sim.t - 1:4d}: {a.elapsed_time(b) / win:.3f} ms/tick wall, tick_pass mean {sum(plain) / len(plain):.3f} median {plain[len(plain) // 2]:.3f} ms, "
          f"agents {int(sim.dev.counts[1].item())}", flush=True)
