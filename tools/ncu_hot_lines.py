#!/usr/bin/env python
"""Top source lines of one kernel of a .ncu-rep by warp-stall samples and by executed instructions
(needs -lineinfo + --import-source on).  Launches matching the name are summed.
usage: ncu_hot_lines.py <report> <kernel-name> [top_n] [--by-inst]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
by_inst = "--by-inst" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, 0, ""])       # (file, line) -> samples, warp insts, thread insts, source
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= 6 and r[0].isdigit():
        def col(name):
            try:
                return int(r[hdr.index(name)])
            except (ValueError, IndexError):
                return 0
        a = agg[(cur_file, int(r[0]))]
        a[0] += col("# Samples"); a[1] += col("Instructions Executed"); a[2] += col("Thread Instructions Executed"); a[3] = r[1].strip()
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
items = sorted(agg.items(), key=lambda kv: kv[1][1 if by_inst else 0], reverse=True)
print(f"{kern}: {tot_s} stall samples, {tot_i} warp instructions")
print("samples%  inst%  thr/inst  file:line  source")
for (f, ln), a in items[:top]:
    print(f"{100.0 * a[0] / tot_s:5.1f}%  {100.0 * a[1] / tot_i:5.1f}%  {a[2] / max(a[1], 1):5.1f}  {f}:{ln:<5} {a[3][:130]}")
