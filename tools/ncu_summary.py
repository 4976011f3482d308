#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the two small text files kept under profiles/:
   <out>_summary.csv    selected raw metrics per captured launch
   <out>_hot_lines.txt  executed warp-instructions and stall samples per CUDA source line (top 40)
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed_op_global_red.sum")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [i for i, h in enumerate(hdr) if h == "Kernel Name" or h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio"))]
with open(out + "_summary.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([hdr[i] for i in cols])
    w.writerow([units[i] for i in cols])
    for r in rows[2:]:
        w.writerow([r[i][:80] for i in cols])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, inst, stall, text = None, collections.Counter(), collections.Counter(), {}
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit():
        try:
            key = (cur, int(r[0]))
            inst[key] += int(r[7]); stall[key] += int(r[4]); text[key] = r[1][:120]
        except ValueError:
            pass
ti, ts = max(sum(inst.values()), 1), max(sum(stall.values()), 1)
with open(out + "_hot_lines.txt", "w") as fh:
    fh.write(f"# {rep}: executed warp-instructions {ti}, stall samples {ts} (all captured launches)\n# %inst %stall file:line source\n")
    for k, v in inst.most_common(40):
        fh.write(f"{100 * v / ti:5.1f} {100 * stall[k] / ts:5.1f}  {k[0]}:{k[1]:<4d} {text[k]}\n")
print("wrote", out + "_summary.csv", out + "_hot_lines.txt")
