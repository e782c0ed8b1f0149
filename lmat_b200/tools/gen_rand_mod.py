#!/usr/bin/env python3
"""Generate the GC-binned null models read_label loads with -n: the job of the reference's bin/gen_rand_mod.sh, with the
GPU rand_read_label (lmat_b200/bin/rand_read_label) and the Python 3 roll-up (merge_cnts.py) in place of the reference's.

For every read length L: rand_read_label draws num_bases / L random reads (ten GC buckets), labels them against the
database and writes <odir>/<db>.<L>.<num_bases>.rl_output.rand_lst; merge_cnts rolls sparsely observed taxids up to their
genus-or-higher ancestors -> <odir>/null.bin.10.<...>.rand_lst(.gz); finally <odir>/<db>.null_lst.txt lists
"<L - k + 1> <file>" per model (the key loadRandHits / closest() select a model by: the read's k-mer count, :566-571).

usage: gen_rand_mod.py --db_file=DB (--read_len=L | --read_range=BEG:END:STEP) [--num_bases=N] [--min_sample_size=100]
                       [--tax_histo_cnt=FILE] [--odir=.] [--threads=1] --tax_dir=DIR | (--conv= --depth= --taxtree= --rankinfo=)
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from lmat_b200.tools import merge_cnts  # noqa: E402


def parse(argv):
    o = dict(db_file="", read_len=0, read_range="", num_bases=10_000_000_000, min_sample_size=100, tax_histo_cnt="", odir=".", threads=1,
             tax_dir="", conv="", depth="", taxtree="", rankinfo="", k=20, binsize=10, debug=False)
    for a in argv:
        if a == "--debug":
            o["debug"] = True
            continue
        if not a.startswith("--") or "=" not in a:
            raise SystemExit(f"Unrecognized argument [{a}]\n{__doc__}")
        k, v = a[2:].split("=", 1)
        if k not in o:
            raise SystemExit(f"Unrecognized argument [{a}]\n{__doc__}")
        o[k] = type(o[k])(v) if not isinstance(o[k], bool) else True
    if o["tax_dir"]:                       # the file names gen_rand_mod.sh:104-110 expects in an LMAT runtime-inputs directory
        d = o["tax_dir"]
        o["conv"] = o["conv"] or os.path.join(d, "m9.32To16.map")
        o["depth"] = o["depth"] or os.path.join(d, "depth_for_ncbi_taxonomy.segment.pruned.dat")
        o["taxtree"] = o["taxtree"] or os.path.join(d, "ncbi_taxonomy.segment.dat.nohl")
        o["rankinfo"] = o["rankinfo"] or os.path.join(d, "ncbi_taxid_to_rank.txt")
        o["tax_histo_cnt"] = o["tax_histo_cnt"] or os.path.join(d, "tcnt.m9.20.tax_histo")
    if not o["db_file"] or (not o["read_len"] and not o["read_range"]):
        raise SystemExit(__doc__)
    return o


def main(argv):
    o = parse(argv)
    exe = os.environ.get("KMAT_RAND_READ_LABEL", os.path.join(os.path.dirname(HERE), "bin", "rand_read_label"))
    if o["read_len"]:
        lens = [o["read_len"]]
    else:
        beg, end, step = (int(x) for x in o["read_range"].split(":"))
        lens = list(range(beg, end + 1, step))
    os.makedirs(o["odir"], exist_ok=True)
    dbname = os.path.basename(o["db_file"])
    entries = []
    for L in lens:
        num_reads = (o["num_bases"] // L) // o["threads"]                       # gen_rand_mod.sh:126-127
        print(f"Create null model. Read_length={L} Reads per thread={num_reads} Total reads={o['num_bases']}")
        oname = f"{dbname}.{L}.{o['num_bases']}.rl_output"
        ofile = os.path.join(o["odir"], oname)
        cmd = [exe, "-w", o["rankinfo"], "-f", o["conv"], "-g", str(num_reads), "-i", str(L), "-e", o["depth"], "-p", "-t", str(o["threads"]),
               "-d", o["db_file"], "-c", o["taxtree"], "-o", ofile]                 # gen_rand_mod.sh:137
        with open(ofile + ".log", "w") as log:
            rc = subprocess.run(cmd, stdout=log, stderr=subprocess.STDOUT).returncode
        sfile = ofile + ".rand_lst"
        if rc != 0 or not os.path.exists(sfile):
            print(f"warning no {sfile} found (rand_read_label exit status {rc}, see {ofile}.log)")
            continue
        out = os.path.join(o["odir"], f"null.bin.{o['binsize']}.{oname}.rand_lst")
        merge_cnts.roll_up(sfile, o["taxtree"], o["rankinfo"], o["min_sample_size"], o["tax_histo_cnt"], out, o["binsize"])
        if not o["debug"]:
            with open(out, "rb") as src, gzip.open(out + ".gz", "wb") as dst:
                shutil.copyfileobj(src, dst)
            os.unlink(out)
            out += ".gz"
        entries.append((L - o["k"] + 1, os.path.basename(out)))                 # gen_rand_mod.sh:152: "$t=$1-19"
    lst = os.path.join(o["odir"], f"{dbname}.null_lst.txt")
    with open(lst, "w") as f:
        for kc, name in sorted(entries):
            f.write(f"{kc} {name}\n")
    print(f"Wrote {lst}; copy it and the null.bin.* files to $LMAT_DIR and pass it to read_label with -n")
    return 0 if entries else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
