#!/usr/bin/env python
"""Where a compaction's time goes at the Nigeria size: tools/diag_compact.py [agents]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from laser_polio_b200 import synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 220_000_000
sim, _ = synth.synth_sim(n, 774, 400, seed=1)
sim.to_device()
sim.run_ticks(30)
eng, dev = sim._engine, sim.dev


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms")
    return out


timed("drain", eng.drain)
timed("DeviceState.compact", dev.compact)
timed("rebuild_tiles", lambda: eng.rebuild_tiles(0))
timed("template", eng._template)
timed("rebase_tallies + hot_build", lambda: eng.rebase_tallies(sim.t))
# inside compact
count = int(dev.counts[1].item())
cap = dev.cap_eff
key = timed("key", lambda: torch.where(dev.cols["disease_state"][:cap] >= 0, dev.cols["node_id"][:cap], torch.full_like(dev.cols["node_id"][:cap], 775)))
perm = timed("sort int16 stable", lambda: torch.sort(key, stable=True).indices)
timed("gather 1-byte column", lambda: dev.cols["strain"][:cap][perm])
timed("gather 4-byte column", lambda: dev.cols["date_of_birth"][:cap][perm])
p32 = perm.to(torch.int32)
timed("gather 4-byte column (int32 index)", lambda: dev.cols["date_of_birth"][:cap][p32])
timed("assign 4-byte column", lambda: dev.cols["date_of_birth"][:cap].copy_(dev.cols["date_of_birth"][:cap][perm]))
