#!/usr/bin/env python
"""Top source lines of ONE captured launch of an .ncu-rep by stall samples and by executed warp-instructions.
usage: tools/ncu_lines.py prof.ncu-rep [launch_index=0] [top=30]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, k, seen = None, -1, []
inst, stall, text = collections.Counter(), collections.Counter(), {}
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "Function Name":
        if r[1] not in seen:
            seen.append(r[1])
        k = seen.index(r[1])
        continue
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif k == which and len(r) > 8 and r[0].isdigit():
        try:
            key = (cur, int(r[0]))
            inst[key] += int(r[7]); stall[key] += int(r[4]); text[key] = r[1][:110]
        except ValueError:
            pass
ti, ts = max(sum(inst.values()), 1), max(sum(stall.values()), 1)
print(f"# launch {which}: {ti} warp-instructions, {ts} stall samples")
print("# by stall samples: %stall %inst file:line source")
order = inst if (len(sys.argv) > 4 and sys.argv[4] == "inst") else stall
for key, v in order.most_common(top):
    v = stall[key]
    print(f"{100 * v / ts:5.1f} {100 * inst[key] / ti:5.1f}  {key[0]}:{key[1]:<4d} {text[key]}")
