#!/usr/bin/env python
"""Throughput of the null-model generator (rand_read_label on the GPU): device-drawn random reads labeled
against the C2 table with rkmer semantics and folded into the per-(taxid, GC bucket) accumulators.
usage: null_bench.py [reads (default 10 M)] [read_len (150)] [genomes (2000)]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lmat_b200 import api, synth  # noqa: E402
from lmat_b200 import fixtures as fx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
G = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
dev = "cuda:0"
wd = tempfile.mkdtemp(prefix="kmat_null_")
tax, m16, anc_tid, anc_sid = synth.make_taxonomy_c2(20240, G)
paths = fx.write_taxonomy_files(tax, wd)
codes = synth.make_genomes_gpu(20240, tax, G, 500000, dev)
tbl = synth.build_table_gpu(codes, anc_sid)
n_kmers = tbl.n
db = synth.upload_table(tbl, 0)
del tbl, codes
torch.cuda.empty_cache()
inputs = api.Inputs(tree=paths["tree"], depth=paths["depth"], rank=paths["rank"], map16=paths["map16"])
ctx = api.Ctx(db, inputs, api.default_opts(rkmer_mode=1))
ctx.set_stats(False)
l0 = api.lib().kmat_launch_count()
ctx.null_random(1, 0, min(n, 1 << 20), L)            # warm-up
ctx.null_reset()
torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.null_random(20260, 0, n, L)                      # returns when the accumulators are final (stream synchronised)
dt = time.perf_counter() - t0
t, m, c, nerr = ctx.null_table()
print(json.dumps({"metric": "null_model_reads_per_s", "value": n / dt, "unit": "reads/s", "n_gpus": 1, "reads": n, "read_len": L,
                  "bases_per_s": n * L / dt, "seconds": dt, "db_kmers": int(n_kmers), "rows": int(len(t)), "reads_hit_root": int(c[t == 1].sum()) if len(t) else 0,
                  "reads_error": int(nerr), "gpu_launches": int(api.lib().kmat_launch_count() - l0),
                  "note": "reads drawn on the device (km_randgen_kernel), K1+K2+K3 with rkmer semantics, km_nullacc_kernel; wall clock around kmat_null_random"}))
