#!/usr/bin/env python
"""Print the SASS of the first captured kernel of an .ncu-rep with per-instruction execution counts, in address order,
keeping only instructions executed at least `frac` x (the most-executed instruction's count): i.e. the hot loop as the
machine ran it.  Columns: offset, warp-instructions executed, avg active threads, stall samples, SASS.
usage: tools/ncu_sass_hot.py prof.ncu-rep [frac=0.05] [launch_index=0]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, rows, k = None, [], -1
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        k += 1
        if k == which:
            kern = r[1]
        continue
    if k != which or not r or not r[0].startswith("0x"):
        continue
    rows.append((int(r[0], 16), r[1].strip(), int(r[2]), int(r[5]), float(r[8] or 0)))
base = rows[0][0]
top = max(x[3] for x in rows)
tot = sum(x[3] for x in rows)
ts = max(sum(x[2] for x in rows), 1)
print(f"# {kern}: {tot} warp-instructions, {ts} stall samples; showing instructions executed >= {frac} x {top}")
shown = 0
for a, s, st, n, thr in rows:
    if n >= frac * top:
        shown += n
        print(f"{a - base:6x} {n:10d} {thr:5.1f} {100 * st / ts:5.1f}%  {s}")
print(f"# shown instructions cover {100 * shown / tot:.1f}% of executed warp-instructions")
