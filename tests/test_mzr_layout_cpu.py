"""CPU: the building blocks of the minimizer-ordered table layout (lmat_b200/csrc/kmat_mzr.h -- a layout study for the
probe kernel, not used by libkmat yet).  tests/mzr_check.cpp is compiled against the header and checks, on the host:
the minimizer order is a bijection with a working inverse; the minimizer, its offset and its strand flag computed from a
read's forward k-mer (either strand) equal those of the canonical k-mer, ties included; (line, key) maps back to the
k-mer for several geometries; a table built with the line/slot rules answers every stored k-mer with its payload and
misses every absent one; and the k-mers of a 150 bp read touch far fewer lines than there are k-mers."""
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
    assert st["read_lines"] < 0.4 * st["read_kmers"], st           # ~39 lines for ~131 k-mers at 1.6 k-mers per line
