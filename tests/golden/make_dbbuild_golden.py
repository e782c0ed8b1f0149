#!/usr/bin/env python
"""Regenerate tests/golden/dbbuild/ from the UNMODIFIED reference (build container only: needs /root/reference
and oracle/_ref/).  For the "small" scenario it runs kmerPrefixCounter -> tax_histo (the four prefix files that are
make_db_table's input; committed gzip'ed) and then the reference make_db_table once per option set below, dumping the
logical table of each reference-built DB (ascending k-mers, CSR offsets, stored ids) as <variant>.npz.
tests/test_dbbuild_cpu.py feeds the same tax_histo files to kmat_table_build and compares record for record."""
import gzip
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "dbbuild")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from oracle import refchain as rc  # noqa: E402

VARIANTS = {
    "plain": {},
    "prune2": {"prune": 2, "numrank": True},            # "small" has lists of up to 3 ids: cut-offs 2 and 1 prune
    "prune2_nomap": {"prune": 2},                       # -g without -m: lists longer than 2 are cut to stored id 1
    "human_adaptor": {"human": True, "adaptor": True},
    "human_prune1": {"human": True, "adaptor": True, "prune": 1, "numrank": True},
}


def kmer_str(v, k):
    return "".join("ACGT"[(int(v) >> (2 * (k - 1 - j))) & 3] for j in range(k))


def main():
    scratch = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_dbbuild_golden"
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(OUT, exist_ok=True)
    inp = S.build_inputs("small", scratch)
    P = inp["paths"]
    k = S.K
    kdbs = rc.kmer_prefix_counter(P["genomes"], k, os.path.join(scratch, "kdb"), scratch)
    ths = [rc.tax_histo(kdb, P["tree"], os.path.join(scratch, f"th.{i}.bin"), scratch) for i, kdb in enumerate(kdbs)]
    for i, t in enumerate(ths):
        with open(t, "rb") as f, gzip.GzipFile(os.path.join(OUT, f"th.{i}.bin.gz"), "wb", mtime=0) as g:
            g.write(f.read())
    # the plain table gives the k-mers to derive the human / adaptor feeds from
    db = rc.make_db_table(ths, os.path.join(scratch, "plain.db"), k, 2, scratch, map16=P["map16"])
    kmers, offs, ids = op.RefDbImage(db).dump()
    rng = np.random.default_rng(77)
    present = rng.choice(kmers, 400, replace=False)
    absent = rng.integers(0, 1 << (2 * k), 600, dtype=np.uint64)
    absent = absent[~np.isin(absent, kmers)]
    # k-mers right around the file boundaries (first / last record of each prefix file) exercise the per-file re-read
    edges = []
    for q in range(1, 4):
        lo = kmers[kmers < (np.uint64(q) << np.uint64(2 * k - 2))].max()
        hi = kmers[kmers >= (np.uint64(q) << np.uint64(2 * k - 2))].min()
        edges += [lo, lo + np.uint64(1), hi - np.uint64(1), hi, hi + np.uint64(1)]
    human = np.unique(np.concatenate([present, absent, np.array(edges, dtype=np.uint64), kmers[-3:] + np.uint64(5)]))
    with open(os.path.join(OUT, "human_kmers.txt"), "w") as f:
        for v in human:
            f.write(kmer_str(v, k) + "\n")
    adaptor = np.unique(np.concatenate([rng.choice(kmers, 60, replace=False), rng.choice(human, 40, replace=False)]))
    with open(os.path.join(OUT, "adaptor_kmers.txt"), "w") as f:
        for v in rng.permutation(adaptor):
            f.write(kmer_str(v, k) + "\n")
    # the id map of the human / adaptor variants must know 9606 and 32630 (the reference asserts otherwise)
    m = open(P["map16"]).read().rstrip("\n").split("\n")
    nxt = max(int(x.split()[1]) for x in m) + 1
    with open(os.path.join(OUT, "map16_human.txt"), "w") as f:
        f.write("\n".join(m) + f"\n9606 {nxt}\n32630 {nxt + 1}\n")
    shutil.copy(P["map16"], os.path.join(OUT, "map16.txt"))
    shutil.copy(P["numrank"], os.path.join(OUT, "numrank.txt"))
    with open(os.path.join(OUT, "numrank_human.txt"), "w") as f:
        f.write(open(P["numrank"]).read().rstrip("\n") + "\n9606 15\n32630 1\n")
    manifest = {}
    for name, v in VARIANTS.items():
        lst = os.path.join(scratch, f"{name}.inputs")
        with open(lst, "w") as f:
            f.write("\n".join(ths) + "\n")
        out_db = os.path.join(scratch, f"{name}.db")
        cmd = [os.path.join(rc.REF_BIN, "make_db_table"), "-i", lst, "-l", "-o", out_db, "-k", str(k), "-s", "2",
               "-f", os.path.join(OUT, "map16_human.txt" if v.get("human") or v.get("adaptor") else "map16.txt")]
        if v.get("prune"):
            cmd += ["-g", str(v["prune"])]
        if v.get("numrank"):
            cmd += ["-m", os.path.join(OUT, "numrank_human.txt" if v.get("human") else "numrank.txt")]
        if v.get("human"):
            # -c: room for the human-only k-mers; without it the reference overruns its record array into the lists
            cmd += ["-j", os.path.join(OUT, "human_kmers.txt"), "-c", "2000"]
        if v.get("adaptor"):
            cmd += ["-u", os.path.join(OUT, "adaptor_kmers.txt")]
        log = rc._run(cmd, log=os.path.join(scratch, f"{name}.mdt.log"))
        km, of, idv = op.RefDbImage(out_db).dump()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kmers=km, offs=of, ids=idv.astype(np.uint32))
        stats = {ln.split(":")[0].strip(): ln.split(":")[1].strip() for ln in log.split("\n")
                 if ln.split(":")[0].strip() in ("singletons", "doubles", "kmers reduced", "kmers cut to 1", "new human k-mers", "new human + other k-mers")}
        manifest[name] = {"n_kmers": int(len(km)), "n_ids": int(len(idv)), "reference_counters": stats}
        print(name, manifest[name])
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
