#!/usr/bin/env python
"""Per-source-line executed warp-instructions of one kernel variant minus another's (averaged over its launches), from
`ncu -i rep --page source --csv --print-source cuda,sass` output.  usage: tools/ncu_diff.py src.csv '<substr of variant A>' '<substr of variant B>' [top]"""
import collections
import csv
import sys

rows = csv.reader(open(sys.argv[1]))
a_key, b_key = sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
agg, launches, fp, cur = {}, collections.Counter(), None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fp = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        cur = agg.setdefault(r[1], collections.Counter())
        if fp == "lpk_tick.cu":
            launches[r[1]] += 1
    elif cur is not None and len(r) > 8 and r[0].isdigit():
        try:
            cur[(fp, int(r[0]), r[1][:100])] += int(r[7])
        except ValueError:
            pass
A = next(v for k, v in agg.items() if a_key in k)
B = next(v for k, v in agg.items() if b_key in k)
na = max(launches[next(k for k in agg if a_key in k)], 1)
nb = max(launches[next(k for k in agg if b_key in k)], 1)
diff = collections.Counter({k: A.get(k, 0) / na - B.get(k, 0) / nb for k in set(A) | set(B)})
print(f"# A: {sum(A.values()) / na / 1e6:.1f} M warp-instructions per launch ({na} launches), B: {sum(B.values()) / nb / 1e6:.1f} M ({nb}); A - B by line:")
for k, v in diff.most_common(top):
    print(f"{v / 1e6:8.2f}M  {k[0]}:{k[1]} {k[2]}")
