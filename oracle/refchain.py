"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference binaries built into oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It shells out to the reference's own tools (kmerPrefixCounter -> tax_histo ->
make_db_table -> read_label; SURVEY.md section 8(c) recipe) and never feeds anything back into the
product path.
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref")


def have_ref(tool: str = "read_label") -> bool:
    return os.access(os.path.join(REF_BIN, tool), os.X_OK)


def _run(cmd, log=None, env=None, cwd=None, check=True, timeout=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=e, cwd=cwd, timeout=timeout)
    if log:
        with open(log, "wb") as f:
            f.write(p.stdout)
    if check and p.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed rc={p.returncode}\n{p.stdout.decode(errors='replace')[-2000:]}")
    return p.stdout.decode(errors="replace")


def kmer_prefix_counter(fasta, k, out_prefix, workdir):
    """kmerPrefixCounter for the four 1-base prefixes (kmerPrefixCounter.cpp:18-27)."""
    outs = []
    for pfx in range(4):
        _run([os.path.join(REF_BIN, "kmerPrefixCounter"), "-i", fasta, "-k", str(k), "-o", out_prefix, "-l", "1",
              "-f", str(pfx)], log=os.path.join(workdir, f"kpc.{pfx}.log"))
        outs.append(f"{out_prefix}.{pfx}")
    return outs


def tax_histo(kdb, tree, out, workdir):
    _run([os.path.join(REF_BIN, "tax_histo"), "-o", out, "-d", kdb, "-t", tree, "-f", "32"],
         log=os.path.join(workdir, os.path.basename(out) + ".log"))
    return out


def make_db_table(th_files, out_db, k, size_gib, workdir, map16=None, prune=None, numrank=None, tid_bits=16):
    """make_db_table.cpp:150-249: -i list -l -o db -k K -s GiB [-f map] [-g N -m ranks]."""
    lst = os.path.join(workdir, os.path.basename(out_db) + ".inputs")
    with open(lst, "w") as f:
        for t in th_files:
            f.write(t + "\n")
    tool = "make_db_table" + ("32" if tid_bits == 32 else "")
    cmd = [os.path.join(REF_BIN, tool), "-i", lst, "-l", "-o", out_db, "-k", str(k), "-s", str(size_gib)]
    if map16 and tid_bits == 16:
        cmd += ["-f", map16]
    if prune:
        cmd += ["-g", str(prune), "-m", numrank]
    if os.path.exists(out_db):
        os.unlink(out_db)
    _run(cmd, log=os.path.join(workdir, os.path.basename(out_db) + ".mdt.log"))
    return out_db


def build_db_from_genomes(fasta, tree, k, out_db, workdir, **kw):
    kdbs = kmer_prefix_counter(fasta, k, os.path.join(workdir, "kdb"), workdir)
    ths = [tax_histo(kdb, tree, os.path.join(workdir, f"th.{i}.bin"), workdir) for i, kdb in enumerate(kdbs)]
    size = kw.pop("size_gib", 2)
    return make_db_table(ths, out_db, k, size, workdir, **kw), ths


def read_label(db, reads, ofbase, depth, tree, threads=1, map16=None, rank=None, names=None, null_lst=None,
               lmat_dir=None, min_score=0, min_kmer=30, hbias=0, sdiff=1.0, prn_all=True, fastq=False,
               prune=None, numrank=None, plasmids=None, extra=(), tid_bits=16, verbose=False, log=None,
               timeout=None):
    """Invoke the reference read_label with the flag set of bin/run_rl.sh:243.  Returns (stdout, query_time)."""
    tool = "read_label" + ("32" if tid_bits == 32 else "")
    cmd = [os.path.join(REF_BIN, tool)]
    if map16 and tid_bits == 16:
        cmd += ["-f", map16]
    if prune:
        cmd += ["-g", str(prune)]
    if numrank:
        cmd += ["-m", numrank]
    if names:
        cmd += ["-u", names]
    if rank:
        cmd += ["-w", rank]
    cmd += ["-x", str(min_score), "-j", str(min_kmer), "-l", str(hbias), "-b", str(sdiff)]
    if null_lst:
        cmd += ["-n", null_lst]
    cmd += ["-e", depth]
    if prn_all:
        cmd += ["-p"]
    if plasmids:
        cmd += ["-r", plasmids]
    if verbose:
        cmd += ["-y"]
    cmd += ["-t", str(threads), "-i", reads, "-d", db, "-c", tree, "-o", ofbase]
    if fastq:
        cmd += ["-q"]
    cmd += list(extra)
    env = {"LMAT_DIR": lmat_dir} if lmat_dir else None
    out = _run(cmd, log=log, env=env, timeout=timeout)
    m = re.search(r"Total query time: ([0-9.eE+-]+) sec", out)
    return out, (float(m.group(1)) if m else None)


def collect_out_lines(ofbase, threads):
    """All per-read lines of <ofbase><t>.out, as a list (thread order, then file order)."""
    lines = []
    for t in range(threads):
        p = f"{ofbase}{t}.out"
        if os.path.exists(p):
            with open(p) as f:
                lines += f.read().split("\n")
    return [ln for ln in lines if ln]


def rand_read_label(db, ofbase, depth, tree, n_reads, read_len, threads=1, map16=None, rank=None, prune=None, numrank=None,
                    fixed_time=None, log=None, timeout=None):
    """Invoke the reference rand_read_label with the flags of bin/gen_rand_mod.sh:137 (-w rank -f map -g N -i L -e depth
    -p -t T -d db -c tree -o out; -h cut -r numeric ranks for the run-time pruning).  fixed_time: preload the time() shim
    so that srand(time(0)) (rand_read_label.cpp:412) gets this seed.  Writes <ofbase>.rand_lst."""
    cmd = [os.path.join(REF_BIN, "rand_read_label")]
    if rank:
        cmd += ["-w", rank]
    if map16:
        cmd += ["-f", map16]
    cmd += ["-g", str(n_reads), "-i", str(read_len), "-e", depth, "-p", "-t", str(threads), "-d", db, "-c", tree, "-o", ofbase]
    if prune:
        cmd += ["-h", str(prune), "-r", numrank]
    env = None
    if fixed_time is not None:
        env = {"LD_PRELOAD": os.path.join(REF_BIN, "libfixedtime.so"), "KMAT_FIXED_TIME": str(fixed_time)}
    return _run(cmd, log=log, env=env, timeout=timeout)
