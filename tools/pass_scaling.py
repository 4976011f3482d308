#!/usr/bin/env python
"""Pass time against table size on one GPU (fixed cost vs per-agent cost): tools/pass_scaling.py [agents ...]"""
import json
import subprocess
import sys

sizes = [int(float(x)) for x in sys.argv[1:]] or [27_500_000, 55_000_000, 110_000_000, 220_000_000]
for n in sizes:
    out = subprocess.run([sys.executable, "bench.py", "--agents", str(n), "--steps", "60", "--warmup", "3", "--cpu-agents", "200000", "--cpu-ticks", "2"],
                         capture_output=True, text=True)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    if not line:
        print(n, "FAILED", out.stderr[-500:])
        continue
    d = json.loads(line[-1])
    r = d["roofline"]
    print(f"{n:>11d} ms/step {d['ms_per_step']:.4f} pass {r['mean_ms']:.4f} classes "
          + " ".join(f"{k}:{v['mean_ms']:.4f}" for k, v in r["per_day_class"].items()) + f" share {r['kernel_share_of_step']} verified {d['verified']}")
