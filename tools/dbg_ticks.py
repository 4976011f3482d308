import sys, json, time
sys.path.insert(0, "/root/repo")
import torch
import bench
import laser_polio_b200 as lp
from laser_polio_b200 import kernels as K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 220_000_000
sim = bench.build_sim(lp, n, 774, 120, seed=1, device="cuda:0")
sim.to_device()
K.STATS.reset(); K.STATS.timing = True
ts = []
for t in range(0, 32):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_tick(sim.t)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms per tick:", [round(x, 2) for x in ts])
for k, v in K.STATS.events.items():
    print(k, [round(a.elapsed_time(b), 2) for a, b in v])
st = sim.dev.cols["disease_state"][: sim.people.count]
print("state mix", [(s, int((st == s).sum())) for s in (-1, 0, 1, 2, 3)], "count", sim.people.count)
