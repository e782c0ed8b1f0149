#!/usr/bin/env python
"""Run every load flavour of the random-gather probe once (for an ncu pass that attributes DRAM bytes and L2
sector requests to one random access).  argv[1] = cudaLimitMaxL2FetchGranularity to request (0 = leave)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmat_b200 import api  # noqa: E402

gran = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eff = api.lib().kmat_set_l2_fetch_granularity(0, gran)
for mode in (8, 16, 32, 101, 102, 103, 104, 105, 106):
    g, s = api.gather_bench(0, 16 << 30, mode, 1 << 28, 2)
    print(json.dumps({"l2_fetch_granularity": eff, "mode": mode, "gathers_per_s": g}), flush=True)
