"""CPU, build container only: what the UNMODIFIED reference does at the two hard limits of the GPU path.

Reads beyond 65,554 bases: libkmat refuses them (KMAT_ST_ERROR / UNSUPPORTED, never a wrong label).  The reference does print a
line -- but `cand_kmer_cnt` is a uint16_t (read_label.cpp:698), so the k-mer count it divides every hit count by, and prints
as the third number of the line, has wrapped modulo 65,536: its scores for such a read are arithmetic on a wrapped count.  This
test pins that observation, so the divergence is a documented property of the reference, not a guess."""
import os

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op
from oracle import refchain as rc


def test_reference_kmer_count_wraps_beyond_65535(tmp_path):
    if not (rc.have_ref("make_db_table") and rc.have_ref("read_label") and rc.have_ref("kmerPrefixCounter") and rc.have_ref("tax_histo")):
        pytest.skip("oracle/_ref binaries not present")
    wd = str(tmp_path)
    inp = S.build_inputs("small", wd)
    P = inp["paths"]
    db, _ = rc.build_db_from_genomes(P["genomes"], P["tree"], S.K, os.path.join(wd, "ref.db"), wd, map16=P["map16"])
    # one 70,000-base read: the scenario's genomes back to back (24 x 4000 bases), so nearly every k-mer hits
    read = "".join(fx.codes_to_str(g) for g in inp["genomes"].values())[:70000]
    assert len(read) == 70000
    fa = os.path.join(wd, "long.fa")
    with open(fa, "w") as f:
        f.write(">long\n" + read + "\n")
    ofb = os.path.join(wd, "rl_long_")
    rc.read_label(db, fa, ofb, P["depth"], P["tree"], threads=1, map16=P["map16"], rank=P["rank"], names=P["names"], null_lst=None,
                  lmat_dir=wd, min_kmer=30, hbias=0, sdiff=1.0, prn_all=True)
    line = open(ofb + "0.out").read().split("\n")[0].split("\t")
    assert line[0] == "long" and len(line) >= 4
    printed = int(line[2].split()[2])                       # "<log_avg> <stdev> <cand_kmer_cnt>"
    valid, _, km, fl = op.encode_read(read, S.K)
    first = int((fl == 1).sum())                            # positions whose label_vec[pos].first >= 0 (first occurrences)
    assert first > 65535, "the read was meant to exceed the 16-bit count"
    assert printed == first % 65536 and printed != first    # the reference's own count has wrapped
