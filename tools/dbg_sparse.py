"""Pass time on a table with (almost) no exposed / infectious agents: the low-prevalence regime of real runs."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, __graft_entry__ as entry
entry.build()
import laser_polio_b200 as lp
from laser_polio_b200 import kernels as K
sim = bench.build_sim(lp, 220_000_000, 774, 100, seed=20261017, device="cuda:0")
st = sim.people.disease_state[: sim.people.count]
st[(st == 1) | (st == 2)] = 3          # everybody exposed / infectious recovers ...
st[:: 1_000_003][:] = 2                # ... except ~220 scattered infectious agents
sim.to_device()
for _ in range(3):
    sim.step_tick(sim.t)
K.STATS.reset(); K.STATS.timing = True
for _ in range(8):
    t = sim.t
    K.STATS.events = {}
    sim.step_tick(t)
    torch.cuda.synchronize()
    print(t, {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in K.STATS.events.items()}, flush=True)
