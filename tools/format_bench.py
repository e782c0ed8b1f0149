#!/usr/bin/env python
"""K5 / run-length packing in numbers: kmat_label_batch_text and kmat_label_batch_packed_rl on a synthetic workload (200 genomes,
2 M x 150 bp reads), once each after a warm-up; run under `ncu --metrics gpu__time_duration.sum -k regex:'km_(format|pack_lists|compact|score)_kernel'`
for the kernels' durations.  Prints the share of reads the device formatter left to the host and the bytes per read."""
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from lmat_b200 import api, synth
    from lmat_b200 import fixtures as fx
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    dev = "cuda:0"
    wd = tempfile.mkdtemp(prefix="kmat_fmt_")
    tax, m16, anc_tid, anc_sid = synth.make_taxonomy_c2(20240, 200)
    P = fx.write_taxonomy_files(tax, wd)
    null_lst = synth.write_null_models_for(tax, wd)
    codes = synth.make_genomes_gpu(20240, tax, 200, 500000, dev)
    tbl = synth.build_table_gpu(codes, anc_sid)
    db = synth.upload_table(tbl, 0)
    inputs = api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=null_lst, lmat_dir=wd)
    ctx = api.Ctx(db, inputs, api.default_opts(min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, want_lineage=0))
    reads = synth.make_reads_gpu(20241, codes, n_reads, 150).cpu().numpy()
    n, L = reads.shape
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    blob = np.ascontiguousarray(reads).reshape(-1)
    for _ in range(2):
        res, cands, lin, tails, on_host = ctx.label_text(blob=blob, offs=offs, prn_all=True)
    text_bytes = sum(len(t) for t in tails if t is not None)
    for _ in range(2):
        r3, c3, l3, n_words = ctx.label_packed_rl(blob=blob, offs=offs)
    print(json.dumps({"reads": n, "left_to_host": on_host, "left_to_host_frac": on_host / n, "text_bytes_per_read": text_bytes / max(1, n - on_host),
                      "pairs_per_read": len(cands) / n, "rl_words_per_read": n_words / n, "rl_bytes_per_pair": 4.0 * n_words / max(1, len(c3))}))


if __name__ == "__main__":
    main()
