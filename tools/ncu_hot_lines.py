#!/usr/bin/env python
"""Top source lines by warp-stall samples for one kernel of a .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_hot_lines.py <report> <kernel-name> [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, lines, total = None, [], 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 6 and r[0].isdigit():
        try:
            s = int(r[4])
        except ValueError:
            continue
        lines.append((s, cur_file, int(r[0]), r[1].strip()))
        total += s
lines.sort(reverse=True)
print(f"{kern}: {total} stall samples")
for s, f, ln, src in lines[:top]:
    print(f"{100.0 * s / max(total, 1):5.1f}%  {f}:{ln:<5} {src[:150]}")
