#!/usr/bin/env python
"""Golden files of BASELINE configs[0]: the reference's own bundled data through the UNMODIFIED reference chain.

  DB     /root/reference/src/kmerdb/examples/tests/data/test.fa (five adenovirus genomes, make_db_table.cpp:111-112) re-headed
         as >1001 .. >1005 for kmerPrefixCounter -> tax_histo -> make_db_table under the taxonomy of scenarios.c1_taxonomy()
  reads  /root/reference/example/example.tgz : simple_list.1000.fna, verbatim (1000 real reads, wrapped FASTA, headers with
         spaces) and 400 reads simulated from the five genomes (1 % substitutions, half of them reverse-complemented, every
         50th with an N at offset 70; numpy PCG64 seed 7)
  out    reference read_label -t 1, option sets run_rl and defaults, for both read files (+ .fastsummary / .nomatchsum)

Run in the build container: python tests/golden/make_golden_c1.py.  Writes tests/golden/c1.*"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from oracle import refchain as rc  # noqa: E402

REF = "/root/reference"


def gz_write(path, data):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(data)


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_golden_c1"
    shutil.rmtree(wd, ignore_errors=True)
    os.makedirs(wd)
    # reads of the tarball, verbatim
    with tarfile.open(os.path.join(REF, "example", "example.tgz")) as tf:
        data = tf.extractfile("simple_list.1000.fna").read()
    gz_write(os.path.join(HERE, "c1.reads_example.fna.gz"), data)
    # genomes
    seqs = []
    for ln in open(os.path.join(REF, "src", "kmerdb", "examples", "tests", "data", "test.fa")):
        ln = ln.strip()
        if ln and not ln.startswith(">"):
            seqs.append(ln)
    assert len(seqs) == len(S.C1_GENOME_TIDS)
    gfa = os.path.join(wd, "genomes.fa")
    with open(gfa, "w") as f:
        for tid, s in zip(S.C1_GENOME_TIDS, seqs):
            f.write(f">{tid}\n{s}\n")
    # simulated reads
    rng = np.random.Generator(np.random.PCG64(7))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    lines = []
    for i in range(400):
        g = int(rng.integers(0, len(seqs)))
        p = int(rng.integers(0, len(seqs[g]) - 150))
        r = list(seqs[g][p:p + 150].upper())
        for j in range(150):
            if rng.random() < 0.01:
                r[j] = "ACGT"[int(rng.integers(0, 4))]
        if rng.random() < 0.5:
            r = [comp.get(c, "N") for c in reversed(r)]
        if i % 50 == 49:
            r[70] = "N"
        lines.append(f">sim{i} genome={S.C1_GENOME_TIDS[g]} pos={p}\n{''.join(r)}\n")
    gz_write(os.path.join(HERE, "c1.reads_sim.fa.gz"), "".join(lines).encode())
    inp = S.build_c1_inputs(wd)
    P = inp["paths"]
    db, _ = rc.build_db_from_genomes(gfa, P["tree"], S.K, os.path.join(wd, "ref.db"), wd, map16=P["map16"])
    img = op.RefDbImage(db)
    kmers, offs, ids = img.dump()
    np.savez_compressed(os.path.join(HERE, "c1.table.npz"), kmers=kmers, offs=offs, ids=ids.astype(np.uint16), kmer_len=np.int32(img.kmer_len), tid_bytes=np.int32(2))
    man = {"n_kmers": int(len(kmers)), "singletons": int((np.diff(offs.astype(np.int64)) == 1).sum()), "outputs": {},
           "inputs": {k: hashlib.sha256(open(P[k], "rb").read()).hexdigest() for k in ("tree", "depth", "rank", "map16", "names", "null_lst")}}
    for rk in ("reads_example", "reads_sim"):
        for oname in ("run_rl", "defaults"):
            o = S.OPTION_SETS[oname]
            ofb = os.path.join(wd, f"rl_{rk}_{oname}_")
            rc.read_label(db, P[rk], ofb, P["depth"], P["tree"], threads=1, map16=P["map16"], rank=P["rank"], names=P["names"], null_lst=P["null_lst"],
                          lmat_dir=wd, min_score=o["min_score"], min_kmer=o["min_kmer"], hbias=o["hbias"], sdiff=o["sdiff"], prn_all=o["prn_all"])
            data = open(ofb + "0.out", "rb").read()
            gz_write(os.path.join(HERE, f"c1.{rk}.{oname}.out.gz"), data)
            man["outputs"][f"{rk}.{oname}"] = hashlib.sha256(data).hexdigest()
            if oname == "run_rl":
                for suf in ("fastsummary", "nomatchsum"):
                    shutil.copy(f"{ofb}.{o['min_score']:g}.{o['min_kmer']}.{suf}", os.path.join(HERE, f"c1.{rk}.run_rl.{suf}"))
    img.close()
    with open(os.path.join(HERE, "c1.manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print(man["n_kmers"], "k-mers,", man["singletons"], "singletons")


if __name__ == "__main__":
    main()
