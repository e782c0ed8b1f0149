"""CPU: the null-model roll-up (lmat_b200/tools/merge_cnts.py, a Python 3 restatement of the reference's Python 2
bin/merge_cnts.py) and the gen_rand_mod driver.  Pinned against tests/golden/rollup/: the outputs of the reference script
itself, executed unmodified under oracle/py2run.py (Python 2 semantics emulated; tests/golden/make_golden_rollup.py).  Also
checked: the Python 2 comparison semantics the script depends on, the structural contract of the output (what loadRandHits,
read_label.cpp:512-678, needs), and that both loaders of this repository -- the oracle's and libkmat's -- accept the models."""
import gzip
import os
import stat
import sys

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api
from lmat_b200.tools import gen_rand_mod, merge_cnts
from oracle import oracle_py as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case,min_obs,thc,inp", [("counts_min2", 2, "counts.txt", "in.rand_lst"), ("counts_min5", 5, "counts.txt", "in.rand_lst"),
                                                  ("nocounts_min2", 2, "missing.txt", "in.nocounts.rand_lst")])
def test_rollup_equals_the_reference_script(case, min_obs, thc, inp, tmp_path):
    """Same lines as /root/reference/bin/merge_cnts.py on the hand-made NCBI-like inputs (E. coli / Shigella defaults,
    eukaryote genus default, human, "other sequences", plasmid-range ids, bins below min_obs, missing count table).  The
    reference emits them in Python 2 dict order, this tool in ascending taxid order (documented deviation; the consumer keys
    every line by taxid), so the bodies are compared as sorted lists."""
    G = os.path.join(GOLDEN, "rollup")
    out = str(tmp_path / "out.txt")
    n = merge_cnts.roll_up(os.path.join(G, inp), os.path.join(G, "tax.dat"), os.path.join(G, "rank.txt"), min_obs, os.path.join(G, thc), out, 10)
    mine = open(out).read().split("\n")
    want = open(os.path.join(G, f"out.{case}.txt")).read().split("\n")
    assert mine[0] == want[0] == "10" and n == len(want) - 2
    assert sorted(mine[1:]) == sorted(want[1:])
    if case != "nocounts_min2":
        assert any(" genus-561 " in ln for ln in want) and any(ln.startswith("32630 ") for ln in want)


def test_py2_harness_semantics():
    """oracle/py2run.py: the three Python 2 behaviours the reference script depends on."""
    from oracle import py2run
    assert py2run.py2_div(7, 2) == 3 and py2run.py2_div(-7, 2) == -4 and py2run.py2_div(7.0, 2) == 3.5
    assert py2run.py2_cmp("Gt", "0.1", 5) and not py2run.py2_cmp("GtE", 0.9, "0.1") and py2run.py2_cmp("Lt", 3, 4) and py2run.py2_cmp("Gt", "b", "a")
    d = py2run.Py2Dict()
    for k in range(20, 0, -1):
        d[k] = k
    assert d.keys() == list(range(1, 21))                       # small ints land in their own slots of the 32-slot table
    d = py2run.Py2Dict()
    for k in (8, 16, 0):
        d[k] = 1
    assert d.keys() == [8, 16, 0]                               # all hash to slot 0 of 8: 8 stays, 16 -> slot 1 (i*5+1+perturb), 0 -> later probe


def test_python2_mixed_comparisons():
    assert merge_cnts.py2_gt("0.5", 0) and merge_cnts.py2_gt("0", 1e9) and not merge_cnts.py2_gt(0.9, "0.1")
    assert merge_cnts.py2_gt("0.5", "0.25") and not merge_cnts.py2_gt("0.0833", "0.5")       # strings: lexicographic
    assert not merge_cnts.py2_ge(0.99, "0.1") and merge_cnts.py2_ge("x", 3) and merge_cnts.py2_ge(3, 3)


@pytest.fixture(scope="module")
def rolled(tmp_path_factory):
    wd = str(tmp_path_factory.mktemp("rollup"))
    inp = S.build_nullgen_inputs(wd)
    P = inp["paths"]
    # the k-mer count table (tcnt.*.tax_histo): every taxid of the tree, leaves "own" more k-mers than internal nodes
    thc = os.path.join(wd, "tcnt.tax_histo")
    with open(thc, "w") as f:
        for t in inp["tax"].tids():
            f.write(f"{t} {5000 if t >= 100000 else 300}\n")
    out = os.path.join(wd, "null.bin.10.x.rand_lst")
    n = merge_cnts.roll_up(os.path.join(GOLDEN, "nullgen.rl150.rand_lst"), P["tree"], P["rank"], 2, thc, out, 10)
    return inp, P, wd, out, n


def test_rollup_output_is_what_loadrandhits_parses(rolled):
    inp, P, wd, out, n = rolled
    lines = open(out).read().split("\n")
    assert lines[0] == "10" and lines[-1] == ""
    body = lines[1:-1]
    tids = [int(ln.split()[0]) for ln in body]
    assert n == len(body) and len(set(tids)) == len(tids)
    assert set(tids) == set(inp["tax"].tids())                      # every taxid of the count table gets a line (562 is not in this tree)
    for ln in body:
        t = ln.split()
        assert len(t) == 2 + 3 * 10
        cls, anc = t[1].rsplit("-", 1)
        assert cls in ("genus", "family", "order", "class", "phylum", "kingdom", "no_rank", "life", "species") and int(anc) in inp["tax"].parent
        for b in range(10):
            int(t[2 + 3 * b]); float(t[3 + 3 * b]); int(t[4 + 3 * b])
    root = [ln for ln in body if ln.startswith("1 ")][0].split()
    assert all(root[3 + 3 * b] == "1.0" for b in range(10))         # tid 1: cut-off 1.0 in every bin (:296-297)
    # a strain's cut-offs come from the strains observed under its genus: values of the .rand_lst, never invented
    seen = set()
    for ln in open(os.path.join(GOLDEN, "nullgen.rl150.rand_lst")):
        seen.update(ln.split()[1::2])
    strain = [ln for ln in body if ln.startswith("100005 ")][0].split()
    assert all(strain[3 + 3 * b] in seen | {"1.0", "0"} for b in range(10))


def test_generated_models_load_in_oracle_and_libkmat(rolled):
    inp, P, wd, out, n = rolled
    with open(out, "rb") as src, gzip.open(os.path.join(wd, "null.131.rand_lst.gz"), "wb") as dst:
        dst.write(src.read())
    lst = os.path.join(wd, "null_lst.txt")
    open(lst, "w").write("131 null.131.rand_lst.gz\n")
    orc = op.Oracle()
    orc.load_files(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=lst, lmat_dir=wd)      # raises on a parse failure
    api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=lst, lmat_dir=wd)


def test_gen_rand_mod_driver_with_a_stub_generator(rolled, tmp_path, monkeypatch):
    """The driver's own logic (read counts per thread, file names, gzip, the null_lst index), with rand_read_label replaced
    by a stub that hands back the reference's .rand_lst -- the real binary needs a GPU (tests/test_null_model.py)."""
    inp, P, wd, out, n = rolled
    stub = tmp_path / "rand_read_label_stub"
    stub.write_text(f"#!{sys.executable}\nimport sys, shutil\na = sys.argv\nassert a[a.index('-i') + 1] in ('150', '170') and int(a[a.index('-g') + 1]) > 0\n"
                    f"shutil.copy({os.path.join(GOLDEN, 'nullgen.rl150.rand_lst')!r}, a[a.index('-o') + 1] + '.rand_lst')\n")
    stub.chmod(stub.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("KMAT_RAND_READ_LABEL", str(stub))
    od = str(tmp_path / "models")
    rc = gen_rand_mod.main([f"--db_file={wd}/my.db", "--read_range=150:170:20", "--num_bases=3000000", "--threads=2", f"--odir={od}", "--min_sample_size=2",
                            f"--conv={P['map16']}", f"--depth={P['depth']}", f"--taxtree={P['tree']}", f"--rankinfo={P['rank']}", f"--tax_histo_cnt={wd}/tcnt.tax_histo"])
    assert rc == 0
    idx = open(os.path.join(od, "my.db.null_lst.txt")).read().split("\n")
    assert idx[0] == "131 null.bin.10.my.db.150.3000000.rl_output.rand_lst.gz" and idx[1] == "151 null.bin.10.my.db.170.3000000.rl_output.rand_lst.gz"
    assert gzip.open(os.path.join(od, idx[0].split()[1])).read().decode() == open(out).read()
    api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=os.path.join(od, "my.db.null_lst.txt"), lmat_dir=od)
