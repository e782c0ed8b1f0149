#!/usr/bin/env python
"""What does a random gather of a whole 128-byte LINE cost on this part?  (First experiment of the minimizer-ordered table,
DESIGN.md section 11: the design only pays if the request-rate ceiling counts lines, not 32-byte sectors or lanes.)

Modes of kmat_gather_bench (km_gather_kernel<MODE>, lmat_b200/csrc/kmat_db.cu); n_gathers counts LANES in every mode:
  32   one random sector per lane, LDG.256 (today's bucket probe)                     -> lanes/s = lines/s
  201  4 consecutive lanes share a random line, each reads its own sector (LDG.256)  -> lines/s = lanes/s / 4
  202  4 consecutive lanes share a random line, all read the SAME sector              -> lines/s = lanes/s / 4
  204  8 consecutive lanes share a random line, 16 bytes each (LDG.128)               -> lines/s = lanes/s / 8
  203  every lane reads all 4 sectors of its own random line (4 x LDG.256)            -> lines/s = lanes/s
  205  4 consecutive lanes share a random line, every lane reads all 4 sectors        -> lines/s = lanes/s / 4
Reading: if 201/202/205 reach ~4x the lane rate of mode 32 the ceiling counts lines after coalescing and a table in
which ~4 neighbouring k-mers of a read share a line cuts the probe kernel's time by up to that factor; 205 vs 201 says
whether each lane may search the whole line or the lanes of a group have to split it and exchange by shuffle."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmat_b200 import api  # noqa: E402

LANES_PER_LINE = {32: 1, 201: 4, 202: 4, 204: 8, 203: 1, 205: 4}
span_gib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for mode in (32, 201, 202, 204, 203, 205):
    try:
        g, _ = api.gather_bench(0, span_gib << 30, mode, 1 << 29, 3)
        print(json.dumps({"mode": mode, "span_gib": span_gib, "G_lanes_per_s": round(g / 1e9, 2),
                          "G_lines_per_s": round(g / 1e9 / LANES_PER_LINE[mode], 2)}), flush=True)
    except Exception as e:
        print(json.dumps({"mode": mode, "error": str(e)[:200]}), flush=True)
