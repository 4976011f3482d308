#!/usr/bin/env python
"""Top source lines of an .ncu-rep by stall samples: tools/ncu_stalls.py rep [n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, inst, stall, text = None, collections.Counter(), collections.Counter(), {}
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit():
        try:
            key = (cur, int(r[0])); inst[key] += int(r[7]); stall[key] += int(r[4]); text[key] = r[1][:110]
        except ValueError:
            pass
ti, ts = max(sum(inst.values()), 1), max(sum(stall.values()), 1)
print(f"# {rep}: warp-instructions {ti}, stall samples {ts}\n# %stall %inst file:line source")
for k, v in stall.most_common(n):
    print(f"{100 * v / ts:5.1f} {100 * inst[k] / ti:5.1f}  {k[0]}:{k[1]:<4d} {text[k]}")
