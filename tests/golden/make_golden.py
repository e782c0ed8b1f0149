#!/usr/bin/env python
"""Regenerate tests/golden/ from the UNMODIFIED reference (run in the build container, where
/root/reference exists and `make -C oracle ref` has produced oracle/_ref/).

For every scenario of tests/scenarios.py it
  1. writes the seeded text inputs (taxonomy, reads, null models, ...) to a scratch directory,
  2. builds the DB with the reference chain kmerPrefixCounter -> tax_histo -> make_db_table,
  3. dumps the logical table (ascending k-mers, CSR offsets, stored 16-bit ids) by walking the
     reference-built image the way SortedDb::begin_/next do            -> <scenario>.table.npz
  4. runs the reference read_label (-t 1) for every option set          -> <scenario>.<opts>.out.gz
     plus .fastsummary / .nomatchsum of the run_rl option set,
  5. records sha256 of every regenerated text input in manifest.json so a drift of the seeded
     generators is detected instead of silently invalidating the goldens.
Nothing here is read at run time on the GPU box except the files it writes.
"""
import gzip
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

GOLDEN_SCENARIOS = ("small", "lists")
INPUT_KEYS = ("tree", "depth", "rank", "map16", "numrank", "names", "reads", "reads_wrapped", "reads_fq", "null_lst",
              "plasmids")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def ref_flags(o):
    extra = []
    if o.get("permissive"):
        extra.append("-s")
    if o.get("phix_off"):
        extra.append("-h")
    if o.get("hide_read"):
        extra.append("-a")
    if o.get("min_fnd"):
        extra += ["-z", str(o["min_fnd"])]
    return extra


def run_reference(db, P, wd, oname, o, reads=None, fastq=False, tag=None):
    ofb = os.path.join(wd, f"rl_{tag or oname}_")
    rc.read_label(db, reads or P["reads"], ofb, P["depth"], P["tree"], threads=1, map16=P["map16"], rank=P["rank"],
                  names=P["names"], null_lst=P["null_lst"] if o["null"] else None, lmat_dir=wd,
                  min_score=o["min_score"], min_kmer=o["min_kmer"], hbias=o["hbias"], sdiff=o["sdiff"],
                  prn_all=o["prn_all"], prune=o.get("prune"), numrank=P["numrank"] if o.get("prune") else None,
                  plasmids=P["plasmids"] if o.get("plasmids") else None, extra=ref_flags(o), fastq=fastq)
    return ofb


def main():
    scratch = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_golden"
    manifest = {}
    for name in GOLDEN_SCENARIOS:
        wd = os.path.join(scratch, name)
        shutil.rmtree(wd, ignore_errors=True)
        inp = S.build_inputs(name, wd)
        P = inp["paths"]
        db, _ = rc.build_db_from_genomes(P["genomes"], P["tree"], S.K, os.path.join(wd, "ref.db"), wd, map16=P["map16"])
        img = op.RefDbImage(db)
        kmers, offs, ids = img.dump()
        np.savez_compressed(os.path.join(HERE, f"{name}.table.npz"), kmers=kmers, offs=offs, ids=ids.astype(np.uint16),
                            kmer_len=np.int32(img.kmer_len), tid_bytes=np.int32(2))
        entry = {"inputs": {k: sha(P[k]) for k in INPUT_KEYS}, "n_kmers": int(len(kmers)), "n_reads": len(inp["seqs"]),
                 "outputs": {}}
        for k in os.listdir(wd):
            if k.startswith("null.") and k.endswith(".gz"):
                entry["inputs"][k] = hashlib.sha256(gzip.open(os.path.join(wd, k)).read()).hexdigest()
        for oname, o in S.OPTION_SETS.items():
            ofb = run_reference(db, P, wd, oname, o)
            data = open(ofb + "0.out", "rb").read()
            with gzip.GzipFile(os.path.join(HERE, f"{name}.{oname}.out.gz"), "wb", mtime=0) as f:
                f.write(data)
            entry["outputs"][oname] = hashlib.sha256(data).hexdigest()
            if oname == "run_rl":
                for suffix in (f".{o['min_score']:g}.{o['min_kmer']}.fastsummary", f".{o['min_score']:g}.{o['min_kmer']}.nomatchsum"):
                    shutil.copy(ofb + suffix, os.path.join(HERE, f"{name}.run_rl{suffix[suffix.rfind('.'):]}"))
        # input-format variants: wrapped FASTA and FASTQ (header-pairing quirk) under the run_rl options
        for tag, reads, fq in (("wrapped", P["reads_wrapped"], False), ("fastq", P["reads_fq"], True)):
            ofb = run_reference(db, P, wd, "run_rl", S.OPTION_SETS["run_rl"], reads=reads, fastq=fq, tag=tag)
            data = open(ofb + "0.out", "rb").read()
            with gzip.GzipFile(os.path.join(HERE, f"{name}.{tag}.out.gz"), "wb", mtime=0) as f:
                f.write(data)
            entry["outputs"][tag] = hashlib.sha256(data).hexdigest()
        manifest[name] = entry
        img.close()
        print(name, "k-mers", entry["n_kmers"], "reads", entry["n_reads"])
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
