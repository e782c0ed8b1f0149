#!/usr/bin/env python
"""Regenerate the gene_label goldens (tests/golden/small.gene*) from the UNMODIFIED reference (build container only).
Gene DB: scenarios.build_gene_table("small") written as a tax_histo-format file and built by the reference
make_db_table32 (DBTID_T = uint32_t, no -f map); input: the reference read_label output of the run_rl option set
(tests/golden/small.run_rl.out.gz); run: reference gene_label -l <list> -d <gene db> -g <annotation.gz> -t 1."""
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from lmat_b200 import fixtures as fx  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from oracle import refchain as rc  # noqa: E402


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_gene_golden"
    shutil.rmtree(wd, ignore_errors=True)
    os.makedirs(wd)
    kmers, offs, gids, annot = S.build_gene_table("small")
    th = os.path.join(wd, "genes.th.bin")
    fx.write_tax_histo(th, S.K, kmers, offs, gids)
    db = os.path.join(wd, "genes.db")
    rc._run([os.path.join(rc.REF_BIN, "make_db_table32"), "-i", th, "-o", db, "-k", str(S.K), "-s", "2"], log=os.path.join(wd, "mdt.log"))
    km2, of2, id2 = op.RefDbImage(db, tid_bytes=4).dump()
    assert np.array_equal(km2, kmers) and np.array_equal(of2, offs) and np.array_equal(id2, gids)
    np.savez_compressed(os.path.join(HERE, "small.genetable.npz"), kmers=km2, offs=of2, ids=id2.astype(np.uint32), kmer_len=np.int32(S.K), tid_bytes=np.int32(4))
    rl = os.path.join(wd, "rl0.out")
    with gzip.open(os.path.join(HERE, "small.run_rl.out.gz"), "rb") as f, open(rl, "wb") as o:
        o.write(f.read())
    lst = os.path.join(wd, "rl.lst")
    open(lst, "w").write(rl + "\n")
    ann = os.path.join(wd, "annot.txt.gz")
    with gzip.GzipFile(ann, "wb", mtime=0) as g:
        g.write(("\n".join(annot) + "\n").encode())
    for tag, extra in (("gene", ["-x", "0", "-q", "0", "-b", "0"]), ("gene_thr", ["-x", "0.3", "-q", "40", "-b", "0.5"])):
        ofb = os.path.join(wd, f"{tag}_")
        rc._run([os.path.join(rc.REF_BIN, "gene_label"), "-l", lst, "-d", db, "-o", ofb, "-g", ann, "-t", "1"] + extra, log=os.path.join(wd, f"{tag}.log"))
        with open(ofb + "0.out", "rb") as f, gzip.GzipFile(os.path.join(HERE, f"small.{tag}.out.gz"), "wb", mtime=0) as g:
            g.write(f.read())
        for fn in os.listdir(wd):
            if fn.startswith(f"{tag}_.") and "genesummary" in fn:
                shutil.copy(os.path.join(wd, fn), os.path.join(HERE, "small." + fn.replace(f"{tag}_.", f"{tag}.")))
                print(fn, os.path.getsize(os.path.join(wd, fn)))
        print(tag, sum(1 for _ in open(ofb + "0.out")), "lines")


if __name__ == "__main__":
    main()
