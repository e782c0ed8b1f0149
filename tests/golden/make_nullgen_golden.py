#!/usr/bin/env python
"""Regenerate the rand_read_label goldens (tests/golden/nullgen.*) from the UNMODIFIED reference (build container only).

The reference seeds rand() with time(0) (src/rand_read_label.cpp:412); oracle/_ref/libfixedtime.so (an LD_PRELOAD time()
shim, oracle/standins/fixed_time.c) pins that seed, `-t 1` keeps the draws on one thread.  DB: scenarios.build_nullgen_inputs
genomes through the reference chain kmerPrefixCounter -> tax_histo -> make_db_table; the logical table is dumped into
nullgen.table.npz; each run of scenarios.NULLGEN_RUNS is the reference's `rand_read_label -w rank -f map -g N -i L -e depth
-p -t 1 -d db -c tree -o out [-h cut -r numrank]` (flags of bin/gen_rand_mod.sh:137) -> nullgen.<tag>.rand_lst."""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from oracle import refchain as rc  # noqa: E402


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_nullgen_golden"
    shutil.rmtree(wd, ignore_errors=True)
    inp = S.build_nullgen_inputs(wd)
    P = inp["paths"]
    db, _ = rc.build_db_from_genomes(P["genomes"], P["tree"], S.K, os.path.join(wd, "ref.db"), wd, map16=P["map16"])
    img = op.RefDbImage(db)
    kmers, offs, ids = img.dump()
    np.savez_compressed(os.path.join(HERE, "nullgen.table.npz"), kmers=kmers, offs=offs, ids=ids.astype(np.uint16),
                        kmer_len=np.int32(img.kmer_len), tid_bytes=np.int32(2))
    print("table:", len(kmers), "k-mers,", int((np.diff(offs) > 1).sum()), "lists")
    man = {"inputs": {k: sha(P[k]) for k in ("tree", "depth", "rank", "map16", "numrank", "genomes")}, "time": S.NULLGEN_TIME}
    for tag, run in S.NULLGEN_RUNS.items():
        ofb = os.path.join(wd, f"rrl_{tag}")
        out = rc.rand_read_label(db, ofb, P["depth"], P["tree"], run["n_reads"], run["read_len"], map16=P["map16"], rank=P["rank"],
                                 prune=run["prune"], numrank=P["numrank"] if run["prune"] else None, fixed_time=S.NULLGEN_TIME,
                                 log=os.path.join(wd, f"rrl_{tag}.log"))
        shutil.copy(ofb + ".rand_lst", os.path.join(HERE, f"nullgen.{tag}.rand_lst"))
        print(tag, sum(1 for _ in open(ofb + ".rand_lst")), "rows")
    json.dump(man, open(os.path.join(HERE, "nullgen.manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
