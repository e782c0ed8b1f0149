#!/usr/bin/env python
"""tests/golden/<scenario>.run_rl_verbose.out.gz: the .out file the UNMODIFIED reference read_label writes under -p -y.
-y prints debug traces to stdout (not reproduced by the GPU path) and ALSO changes the -p list of the .out line: candidates
with a negative score are printed too (read_label.cpp:901).  Run in the build container after `make -C oracle ref`."""
import gzip
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from oracle import refchain as rc  # noqa: E402


def main():
    scratch = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_golden_verbose"
    for name in ("small", "lists"):
        wd = os.path.join(scratch, name)
        shutil.rmtree(wd, ignore_errors=True)
        inp = S.build_inputs(name, wd)
        P = inp["paths"]
        db, _ = rc.build_db_from_genomes(P["genomes"], P["tree"], S.K, os.path.join(wd, "ref.db"), wd, map16=P["map16"])
        o = S.OPTION_SETS["run_rl"]
        ofb = os.path.join(wd, "rl_verbose_")
        rc.read_label(db, P["reads"], ofb, P["depth"], P["tree"], threads=1, map16=P["map16"], rank=P["rank"], names=P["names"],
                      null_lst=P["null_lst"], lmat_dir=wd, min_score=o["min_score"], min_kmer=o["min_kmer"], hbias=o["hbias"],
                      sdiff=o["sdiff"], prn_all=True, verbose=True)
        data = open(ofb + "0.out", "rb").read()
        plain = gzip.open(os.path.join(HERE, f"{name}.run_rl.out.gz")).read()
        with gzip.GzipFile(os.path.join(HERE, f"{name}.run_rl_verbose.out.gz"), "wb", mtime=0) as f:
            f.write(data)
        print(name, len(data), "bytes;", sum(a != b for a, b in zip(data.split(b"\n"), plain.split(b"\n"))), "lines differ from the plain -p golden")


if __name__ == "__main__":
    main()
