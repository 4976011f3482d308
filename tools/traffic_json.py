#!/usr/bin/env python
"""profiles/r2_traffic.json (what bench.py reports as roofline.traffic) from the ncu summaries it cites, so that the figure is
regenerated with the captures instead of typed in:   tools/traffic_json.py [agents_at_capture=220500000]
Launch order of the captures (tools/measure_round2.sh): *_sia_vd = ticks 20 (campaign), 21 (vital dynamics), 22 (plain);
*_plain = ticks 40, 41 (plain), 42 (vital dynamics + RI)."""
import csv
import json
import sys
from pathlib import Path

P = Path(__file__).resolve().parents[1] / "profiles"
agents = int(float(sys.argv[1])) if len(sys.argv) > 1 else 220_500_000
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(name):
    rows = list(csv.reader(open(P / name)))
    h, units = rows[0], rows[1]
    r_i, w_i = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    return [float(r[r_i]) * UNIT[units[r_i]] + float(r[w_i]) * UNIT[units[w_i]] for r in rows[2:]]


a, b = launches("r2_fused_v35_sia_vd_summary.csv"), launches("r2_fused_v35_plain_summary.csv")
per_launch = {"sia": a[0], "vd": a[1], "plain": (b[0] + b[1]) / 2, "vdri": b[2]}
out = {"source": "ncu --set full --clock-control none captures of k_tick_pass inside bench.py at 2.2e8 agents (Nigeria shape), this round: "
                 "profiles/r2_fused_v35_plain_summary.csv (ticks 40-42) and profiles/r2_fused_v35_sia_vd_summary.csv (ticks 20-22); "
                 "dram__bytes_read.sum + dram__bytes_write.sum per launch / agents in the table (tools/traffic_json.py)",
       "agents_at_capture": agents, "bytes_per_launch": per_launch,
       "bytes_per_agent": {k: round(v / agents, 3) for k, v in per_launch.items()}}
out["bytes_per_agent"]["siavd"] = out["bytes_per_agent"]["siavdri"] = out["bytes_per_agent"]["sia"]
(P / "r2_traffic.json").write_text(json.dumps(out, indent=1))
print(out["bytes_per_agent"])
