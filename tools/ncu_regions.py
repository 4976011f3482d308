#!/usr/bin/env python
"""Aggregate an .ncu-rep's per-source-line samples into regions of lpk_tick.cu (function bodies) and helper files.
usage: tools/ncu_regions.py report.ncu-rep [launch filter ignored]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# function regions of lpk_tick.cu from the source itself
import pathlib
tick = pathlib.Path(__file__).resolve().parent.parent / "laser-polio_b200/csrc/lpk_tick.cu"
regions = []  # (start_line, name)
for n, line in enumerate(tick.read_text().splitlines(), 1):
    m = re.match(r"^(?:template.*\n)?(?:static |__device__ |__global__ |extern \"C\" )+.*?(\w+)\(", line)
    if m and not line.startswith(" "):
        regions.append((n, m.group(1)))
    m2 = re.match(r"^\s+auto (\w+) = \[&\]", line)
    if m2:
        regions.append((n, "pass::" + m2.group(1)))
def region(cur, ln):
    if cur != "lpk_tick.cu":
        return cur
    name = "?"
    for s, nm in regions:
        if s <= ln: name = nm
        else: break
    return name
cur, hdr = None, None
inst, stall, kinds = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif len(r) > 8 and r[0].isdigit():
        try:
            key = region(cur, int(r[0]))
            inst[key] += int(r[7]); stall[key] += int(r[4])
            for i, h in enumerate(hdr):
                if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit():
                    kinds[key][h] += int(r[i])
        except ValueError:
            pass
ti, ts = max(sum(inst.values()), 1), max(sum(stall.values()), 1)
print(f"# {rep}: warp-inst {ti}, samples {ts}")
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:25]:
    top = ", ".join(f"{h[6:]} {100*c/v:.0f}%" for h, c in kinds[k].most_common(4)) if v else ""
    print(f"{100*inst[k]/ti:5.1f}%inst {100*v/ts:5.1f}%stall  {k:28s} {top}")
