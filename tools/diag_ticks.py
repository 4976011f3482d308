#!/usr/bin/env python
"""Per-tick diagnostics of the fused engine on the bench workload: host time of step_tick() and (device-synchronised)
CUDA-event time of every liblpk call, tick by tick.  usage: tools/diag_ticks.py [agents] [ticks]"""
import sys
import time
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
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sim = bench.build_sim(lp, agents, 774, 2 * ticks + 60, seed=20261017, device="cuda:0")
sim.to_device()
for _ in range(3):
    sim.step_tick(sim.t)
torch.cuda.synchronize()
K.STATS.reset()
K.STATS.timing = True
print("tick host_us  device_ms  kernels")
for _ in range(ticks):
    t = sim.t
    K.STATS.events = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.step_tick(t)
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    dev = time.perf_counter() - t0
    ks = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in K.STATS.events.items()}
    print(f"{t:4d} {host * 1e6:8.0f} {dev * 1e3:9.3f}  {ks}")
# free-running (no per-tick sync): wall per tick
K.STATS.timing = False
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(ticks):
    sim.step_tick(sim.t)
host = time.perf_counter() - t0
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"free-running {ticks} ticks: host {host / ticks * 1e3:.3f} ms/tick, wall {wall / ticks * 1e3:.3f} ms/tick")
