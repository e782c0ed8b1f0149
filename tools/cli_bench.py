#!/usr/bin/env python
"""File-to-file throughput of the drop-in host binary (lmat_b200/bin/read_label): synthetic DB saved as a flat .kmat
image + a FASTA of synthetic reads in /dev/shm -> <ofbase><t>.out / .fastsummary / .nomatchsum.  Reports the binary's
own "Total query time" (read_label.cpp:1868-1869 equivalent: excludes DB load / table upload) and wall time.
usage: cli_bench.py [--genomes 200] [--reads 4000000] [--threads 8] [--keep]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--genome-len", type=int, default=500000)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE passed to the binary")
    a = ap.parse_args()
    import numpy as np
    import torch
    from lmat_b200 import api, build
    from lmat_b200 import fixtures as fx
    from lmat_b200 import synth
    lib, exe = build.build_all()
    wd = tempfile.mkdtemp(prefix="kmat_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        dev = "cuda:0"
        tax, m16, anc_tid, anc_sid = synth.make_taxonomy_c2(20240, a.genomes)
        P = fx.write_taxonomy_files(tax, wd)
        null_lst = synth.write_null_models_for(tax, wd)
        codes = synth.make_genomes_gpu(20240, tax, a.genomes, a.genome_len, dev)
        tbl = synth.build_table_gpu(codes, anc_sid)
        sid2tid = np.zeros(65536, dtype=np.uint32)
        for t, s in m16.items():
            sid2tid[s] = t
        kmers, offs, tids, sids = synth.table_to_host(tbl, sid2tid)
        db = os.path.join(wd, "synth.kmat")
        api.Table.from_arrays(kmers, offs, sids, 20, 2).save(db)
        n_kmers = len(kmers)
        del tbl, kmers, offs, tids, sids
        reads = synth.make_reads_gpu(20241, codes, a.reads, a.read_len).cpu().numpy()
        del codes
        torch.cuda.empty_cache()
        fa = os.path.join(wd, "reads.fa")
        t0 = time.time()
        # ">r<i>\n<seq>\n" records written with numpy (a Python loop over 10^7 reads would take minutes)
        n, L = reads.shape
        hdr = np.char.add(">r", np.arange(n).astype(str)).astype("S")
        w = hdr.dtype.itemsize
        rec = np.full((n, w + 1 + L + 1), ord("\n"), dtype=np.uint8)
        rec[:, :w] = np.frombuffer(hdr.tobytes(), dtype=np.uint8).reshape(n, w)
        # headers are right-padded with NULs by numpy: move the newline right behind the text by padding with spaces
        rec[:, :w][rec[:, :w] == 0] = ord(" ")
        rec[:, w + 1:w + 1 + L] = reads
        rec.tofile(fa)
        fa_bytes = os.path.getsize(fa)
        print(f"wrote {fa} ({fa_bytes / 1e9:.2f} GB) in {time.time() - t0:.1f}s; db {os.path.getsize(db) / 1e9:.2f} GB, {n_kmers} k-mers", file=sys.stderr)
        args = [exe, "-f", P["map16"], "-u", P["names"], "-w", P["rank"], "-x", "0", "-j", "30", "-l", "0", "-b", "1.0", "-n", null_lst, "-e", P["depth"],
                "-p", "-t", str(a.threads), "-i", fa, "-d", db, "-c", P["tree"], "-o", os.path.join(wd, "out")]
        env = dict(os.environ, LMAT_DIR=wd)
        for kv in a.env:
            k, v = kv.split("=", 1)
            env[k] = v
        best = None
        for it in range(a.repeat):
            for f in os.listdir(wd):
                if f.startswith("out"):
                    os.remove(os.path.join(wd, f))
            t0 = time.time()
            p = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
            wall = time.time() - t0
            if p.returncode != 0:
                print(p.stdout[-2000:], p.stderr[-2000:], file=sys.stderr)
                raise SystemExit(f"read_label failed rc={p.returncode}")
            for ln in p.stderr.splitlines():
                if ln.startswith("[kmat"):
                    print(ln, file=sys.stderr)
            q = float(re.search(r"Total query time: ([0-9.eE+-]+) sec", p.stdout).group(1))
            up = float(re.search(r"Table upload time: ([0-9.eE+-]+) sec", p.stdout).group(1))
            out_bytes = sum(os.path.getsize(os.path.join(wd, f)) for f in os.listdir(wd) if re.match(r"out\d+\.out$", f))
            r = {"query_s": q, "wall_s": wall, "upload_s": up, "out_bytes": out_bytes}
            if best is None or q < best["query_s"]:
                best = r
        line = {"metric": "reads_per_s (file to file, read_label host binary)", "value": a.reads / best["query_s"], "unit": "reads/s",
                "reads": a.reads, "read_len": a.read_len, "db_kmers": int(n_kmers), "threads": a.threads, "gpus": api.device_count(),
                "query_s": best["query_s"], "wall_s": best["wall_s"], "table_upload_s": best["upload_s"],
                "fasta_GBps": fa_bytes / best["query_s"] / 1e9, "out_GBps": best["out_bytes"] / best["query_s"] / 1e9,
                "host_cores": os.cpu_count(), "env": a.env}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
