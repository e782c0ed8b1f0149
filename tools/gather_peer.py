#!/usr/bin/env python
"""Random 32-byte-sector gathers against ANOTHER GPU's memory (peer access over NVLink) for several load flavours:
the request-rate ceiling of the direct sharded mode's remote bucket reads.  Needs >= 2 GPUs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmat_b200 import api  # noqa: E402

NAMES = {8: "ld.global.u64", 32: "ld.global.nc.L1::no_allocate.L2::64B.v4.u64 (table load)", 105: "ld.global.L1::no_allocate.L2::64B.u64",
         107: "ld.global.nc.L1::no_allocate.v4.u64", 108: "ld.relaxed.sys.global.v2.u64", 109: "ld.volatile.global.v2.u64",
         102: "ld.global.cv.u64", 110: "cp.async.bulk 32 B global->smem (TMA)"}
span = int(sys.argv[1]) << 30 if len(sys.argv) > 1 else 8 << 30
for mem in (1, 0):
    for mode in (32, 107, 8, 105, 108, 109, 102, 110):
        try:
            g, s = api.gather_bench_peer(0, mem, span, mode, 1 << 27, 2)
            print(json.dumps({"exec_device": 0, "mem_device": mem, "mode": mode, "load": NAMES[mode], "G_gathers_per_s": round(g / 1e9, 2), "sector_GBps": round(s, 1)}), flush=True)
        except Exception as e:
            print(json.dumps({"mem_device": mem, "mode": mode, "error": str(e)[:200]}), flush=True)
