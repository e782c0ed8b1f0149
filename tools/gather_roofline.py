#!/usr/bin/env python
"""Measure the random-access HBM roofline on this GPU: uniform random loads, one per
32-byte sector, over spans from L2-resident to tens of GiB, at 8/32 bytes per access and for each
cudaLimitMaxL2FetchGranularity setting.  Prints JSON lines and writes the list to argv[1]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmat_b200 import api  # noqa: E402

out = []
for gran in (0, 32, 64, 128):
    eff = api.lib().kmat_set_l2_fetch_granularity(0, gran)
    for span_gib in (1, 16, 64):
        for ab in (8, 32):
            g, s = api.gather_bench(0, int(span_gib * (1 << 30)), ab, 1 << 29, 5)
            out.append({"l2_fetch_granularity_req": gran, "l2_fetch_granularity": eff, "span_gib": span_gib, "access_bytes": ab,
                        "gathers_per_s": g, "sector_GBps": s})
            print(json.dumps(out[-1]), flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gather_roofline.json", "w"), indent=1)
