#!/usr/bin/env python
"""The synthetic scaling sweep of BASELINE.json's configs on one GPU: full-feature days (vital dynamics, RI, campaigns,
transmission, census) at 1e7 .. 3e8 agents x 1000 / 10000 nodes, plus the Zamfara shape (14 nodes, 5e6 agents).
100 days after 10 warm-up each; one line per run.    usage: tools/sweep.py > profiles/r2_sweep.txt"""
import json
import subprocess
import sys

runs = [("zamfara", 5_000_000, 14)] + [(f"sweep", n, m) for n in (10_000_000, 30_000_000, 100_000_000, 300_000_000) for m in (1000, 10000)]
print("# shape agents nodes | agent-days/s  ms/day  pass_ms  node_ms | moved-bytes frac of copy peak (plain day) | e2e agent-days/s | verified")
for name, n, m in runs:
    cmd = [sys.executable, "bench.py", "--agents", str(n), "--nodes", str(m), "--steps", "100", "--warmup", "10", "--cpu-agents", "200000", "--cpu-ticks", "2"]
    if name == "zamfara":
        cmd += ["--config", "zamfara"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired:
        print(name, n, m, "TIMEOUT", flush=True)
        continue
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    if not line:
        print(name, n, m, "FAILED", out.stderr[-300:].replace("\n", " "), flush=True)
        continue
    d = json.loads(line[-1])
    r = d["roofline"]
    share = r["kernel_share_of_step"]
    plain = r["per_day_class"].get("plain", {})
    print(f"{name:8s} {n:>11d} {m:>6d} | {d['value']:.3e}  {d['ms_per_step']:.4f}  {r['mean_ms']:.4f}  {share.get('tick_node', 0) * d['ms_per_step']:.4f} | "
          f"{plain.get('moved_frac_of_peak')} | {d['e2e']['value']:.3e} | {d['verified']}", flush=True)
