"""gene_label (src/gene_label.cpp:217-301, 378-713).

CPU: the oracle's restatement against the output of the UNMODIFIED reference gene_label on a 32-bit gene DB
(tests/golden/small.gene.out.gz, made by tests/golden/make_gene_golden.py).
GPU (-m gpu): kmat_gene_batch against the oracle, and the gene_label drop-in binary against the reference's .out and
.genesummary files byte for byte."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api, build
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gene_table():
    t = np.load(os.path.join(GOLDEN, "small.genetable.npz"))
    return t["kmers"], t["offs"], t["ids"].astype(np.uint32)


@pytest.fixture(scope="module")
def gene_oracle(gene_table):
    sd = op.SortedDbArrays(*gene_table, 20, 4)
    return op.Oracle(cdb=sd.cdb(), keep=sd)


def rl_records():
    """(hdr, read, taxid, tax_score) of the lines gene_label processes, parsed like gene_label.cpp:586-623."""
    out = []
    for line in gzip.open(os.path.join(GOLDEN, "small.run_rl.out.gz")).read().decode("latin-1").split("\n"):
        c = line.split("\t")
        if len(c) < 5:
            continue
        st = c[2].split()
        try:
            if float(st[2]) == -1:
                continue
            call = c[4].split()
            tid, ts, mt = int(call[0]), np.float32(call[1]), call[2]
        except (ValueError, IndexError):
            continue                      # the record glued to the silent-NoMatch read (no newline quirk): UB in the reference
        out.append((c[0], c[1], 0 if mt[0] in "NR" else tid, ts))
    return out


def fmt(hdr, read, tid, ts, count, cnt, gene):
    g = lambda v: "%g" % np.float32(v)
    return f"{hdr}\t{read}\t{tid} {g(ts)}\t\t-1 {count} {cnt}\t{gene} {g(np.float32(count) / np.float32(cnt))} GL"


def test_gene_table_fixture_is_reproducible(gene_table):
    kmers, offs, ids, _ = S.build_gene_table("small")
    assert np.array_equal(kmers, gene_table[0]) and np.array_equal(offs, gene_table[1]) and np.array_equal(ids, gene_table[2])
    assert (np.diff(offs.astype(np.int64)) > 1).sum() > 100 and ids.min() >= S.GENE_ID0


def test_oracle_gene_label_equals_reference(gene_oracle):
    want = {w.split("\t")[0]: w for w in gzip.open(os.path.join(GOLDEN, "small.gene.out.gz")).read().decode("latin-1").split("\n") if w}
    n_ok = 0
    for hdr, read, tid, ts in rl_records():
        n, cnt, gene, count = gene_oracle.gene_label([read])[0]
        if n == 0:
            assert hdr not in want
            continue
        assert want[hdr] == fmt(hdr, read, tid, ts, count, cnt, gene)
        n_ok += 1
    assert n_ok >= len(want) - 1 and n_ok > 300


@pytest.mark.gpu
def test_gene_batch_equals_oracle(gene_table, gene_oracle):
    db = api.Db.upload(api.Table.from_arrays(*gene_table, 20, 4))
    reads = [r for _, r, _, _ in rl_records()]
    inp = S.build_inputs("small", "/tmp/kmat_gene_inputs")
    names = list(inp["genomes"])
    rng = np.random.default_rng(3)
    for L in (0, 5, 19, 20, 21, 150, 161, 256, 257, 700, 3900):                        # every K1 variant, several genes per read
        g = fx.codes_to_str(inp["genomes"][names[int(rng.integers(0, len(names)))]])
        reads += [g[:L], (g[100:100 + L // 2] * 2)[:L], g[:L].lower()]
    got = db.gene_label(reads)
    want = gene_oracle.gene_label(reads)
    assert (got["status"] >= 0).all()
    for i, (n, cnt, gene, count) in enumerate(want):
        assert got["n_genes"][i] == n and got["valid_kmers"][i] == cnt, i
        assert got["status"][i] == (1 if n else 0)
        if n:
            assert got["gene"][i] == gene and got["count"][i] == count, i
            assert got["score"][i] == np.float32(count) / np.float32(cnt)
    assert (got["n_genes"] > 1).sum() > 20


@pytest.mark.gpu
def test_gene_batch_reads_with_many_genes(gene_table, gene_oracle):
    """Reads stitched from many genomes hit more genes than the 64 register slots of km_gene_kernel: the big kernel
    (up to 1024 genes in shared memory) must give the oracle's top gene, count and score."""
    db = api.Db.upload(api.Table.from_arrays(*gene_table, 20, 4))
    inp = S.build_inputs("small", "/tmp/kmat_gene_inputs_big")
    gstr = [fx.codes_to_str(g) for g in inp["genomes"].values()]
    rng = np.random.default_rng(5)
    reads = []
    for L, piece in [(150, 21), (2000, 21), (2000, 30), (6000, 25), (30000, 40)]:
        for rep in range(4):
            parts = []
            while sum(map(len, parts)) < L:
                gs = gstr[int(rng.integers(0, len(gstr)))]
                a = int(rng.integers(0, len(gs) - piece))
                parts.append(gs[a:a + piece])
            reads.append("".join(parts)[:L])
    reads.append(gstr[0][:150])
    got = db.gene_label(reads)
    want = gene_oracle.gene_label(reads)
    assert max(n for n, *_ in want) > 64
    assert (got["status"] >= 0).all(), got["status"]
    for i, (n, cnt, gene, count) in enumerate(want):
        assert got["n_genes"][i] == n and got["valid_kmers"][i] == cnt, i
        assert got["status"][i] == (1 if n else 0)
        if n:
            assert got["gene"][i] == gene and got["count"][i] == count, (i, n)
            assert got["score"][i] == np.float32(count) / np.float32(cnt)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,extra", [("gene", ["-x", "0", "-q", "0", "-b", "0"]), ("gene_thr", ["-x", "0.3", "-q", "40", "-b", "0.5"])])
def test_gene_label_cli_equals_reference(gene_table, tmp_path, tag, extra):
    build.build_all()
    db = str(tmp_path / "genes.kmat")
    api.Table.from_arrays(*gene_table, 20, 4).save(db)
    rl = str(tmp_path / "rl0.out")
    open(rl, "wb").write(gzip.open(os.path.join(GOLDEN, "small.run_rl.out.gz")).read())
    lst = str(tmp_path / "rl.lst")
    open(lst, "w").write(rl + "\n")
    annot = S.build_gene_table("small")[3]
    ann = str(tmp_path / "annot.txt.gz")
    with gzip.open(ann, "wb") as f:
        f.write(("\n".join(annot) + "\n").encode())
    ofb = str(tmp_path / f"{tag}_")
    p = subprocess.run([build.GL_BIN, "-l", lst, "-d", db, "-o", ofb, "-g", ann, "-t", "1"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    want = gzip.open(os.path.join(GOLDEN, f"small.{tag}.out.gz")).read().decode("latin-1").split("\n")
    got = open(ofb + "0.out", encoding="latin-1").read().split("\n")
    # the record glued to the silent-NoMatch read (reference quirk) parses through uninitialised floats in
    # the reference; every other line must be identical
    glued = {w.split("\t")[0] for w in want if w.startswith("period25")} | {g.split("\t")[0] for g in got if g.startswith("period25")}
    assert [w for w in want if w.split("\t")[0] not in glued] == [g for g in got if g.split("\t")[0] not in glued]
    for fn in os.listdir(GOLDEN):
        if fn.startswith(f"small.{tag}.") and "genesummary" in fn:
            mine = ofb + fn[len(f"small.{tag}"):]
            assert os.path.exists(mine), mine
            a, b = open(os.path.join(GOLDEN, fn)).read(), open(mine).read()
            assert a == b, fn
