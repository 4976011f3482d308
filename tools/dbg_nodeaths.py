import sys, torch
sys.path.insert(0, '/root/repo')
import bench, __graft_entry__ as entry
entry.build()
import laser_polio_b200 as lp
from laser_polio_b200 import kernels as K
sim = bench.build_sim(lp, 220_000_000, 774, 100, seed=20261017, device="cuda:0")
if len(sys.argv) > 1 and sys.argv[1] == "nodeaths":
    sim.people.date_of_death[: sim.people.count] += 1_000_000
sim.to_device()
for _ in range(3):
    sim.step_tick(sim.t)
K.STATS.reset(); K.STATS.timing = True
for _ in range(12):
    t = sim.t
    K.STATS.events = {}
    sim.step_tick(t)
    torch.cuda.synchronize()
    print(t, {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in K.STATS.events.items()}, flush=True)
