#!/usr/bin/env python
"""State mix of the bench workload over time: per-tick fractions S/E/I/R, share of strain 0 among E/I, exposures per day.
usage: tools/diag_mix.py [agents] [ticks]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
import laser_polio_b200 as lp  # noqa: E402

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 220_000_000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 140
sim = bench.build_sim(lp, agents, 774, 2 * ticks + 60, seed=20261017, device="cuda:0")
sim.to_device()
for _ in range(ticks + 3):
    sim.step_tick(sim.t)
sim.to_host()
r = sim.results
n = float(agents)
print("tick   S      E      I      R    E0/E   I0/I  new_exposed/day  sia_prot  ri_prot")
for t in list(range(2, ticks, 8)):
    E, I = r.E[t].sum(), r.I[t].sum()
    print(f"{t:4d} {r.S[t].sum()/n:6.3f} {E/n:6.4f} {I/n:6.4f} {r.R[t].sum()/n:6.3f} {r.E_by_strain[t,:,0].sum()/max(E,1):6.3f} "
          f"{r.I_by_strain[t,:,0].sum()/max(I,1):6.3f} {r.new_exposed[t].sum()/n:10.5f} {r.sia_protected[t].sum()/n:9.5f} {r.ri_protected[t].sum()/n:8.5f}")
