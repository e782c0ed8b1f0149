"""CPU: the minimizer-ordered first level of the k-mer table (lmat_b200/csrc/kmat_mzr.h).  tests/mzr_check.cpp is compiled
against the header and checks, on the host: both hashes are bijections with working inverses; the minimizer, its offset and
its strand flag computed from a read's forward k-mer (either strand) equal those of the canonical k-mer, ties included; the
table key maps back to the k-mer and names one owner shard whose line range holds it, for several geometries; the shard
ranges tile the line space; a two-level table built with the sector / overflow-flag rules answers every stored k-mer with
its payload and misses every absent one; the sliding-minimum form the probe kernel uses (one hash per base) equals the
definition for every window count 2..8; and the k-mers of a 150 bp read make far fewer requests than there are k-mers."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_minimizer_layout_properties(tmp_path):
    exe = str(tmp_path / "mzr_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "mzr_check.cpp")], check=True)
    p = subprocess.run([exe, "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    st = json.loads(p.stdout)
    assert st["ties"] > 1000 and st["absent_checked"] > 100000
    assert st["read_line_requests"] + st["read_second_level"] < 0.45 * st["read_kmers"], st      # ~48 requests for ~131 k-mers
