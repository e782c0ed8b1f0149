"""BASELINE configs[0]: the reference's bundled data (tests/golden/c1.*, made by tests/golden/make_golden_c1.py with the
UNMODIFIED reference chain): a DB of the five genomes of src/kmerdb/examples/tests/data/test.fa, the 1000 real reads of
example/example.tgz (wrapped FASTA, headers with spaces) and 400 reads simulated from the genomes.  CPU: the oracle and the
host reader against the reference's outputs.  GPU: kmat_label_batch and the drop-in binary, byte for byte."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api, build
from oracle import oracle_py as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class C1:
    def __init__(self, workdir):
        self.inp = S.build_c1_inputs(workdir)
        self.paths = self.inp["paths"]
        self.workdir = workdir
        man = json.load(open(os.path.join(GOLDEN, "c1.manifest.json")))
        import hashlib
        for k, want in man["inputs"].items():
            assert hashlib.sha256(open(self.paths[k], "rb").read()).hexdigest() == want, f"seeded input {k} drifted from the c1 manifest"
        t = np.load(os.path.join(GOLDEN, "c1.table.npz"))
        self.kmers, self.offs, self.ids = t["kmers"], t["offs"], t["ids"].astype(np.uint32)
        assert len(self.kmers) == man["n_kmers"] == 147121            # what the reference chain builds from test.fa

    def golden(self, name):
        return gzip.open(os.path.join(GOLDEN, f"c1.{name}.out.gz")).read().decode("latin-1")


@pytest.fixture(scope="module")
def c1(tmp_path_factory):
    return C1(str(tmp_path_factory.mktemp("c1")))


def _opts(o):
    return dict(min_kmer=o["min_kmer"], hbias=o["hbias"], sdiff=o["sdiff"], min_score=o["min_score"])


@pytest.mark.parametrize("reads", ["reads_example", "reads_sim"])
@pytest.mark.parametrize("oname", ["run_rl", "defaults"])
def test_oracle_reproduces_reference_on_bundled_data(c1, reads, oname):
    o = S.OPTION_SETS[oname]
    sd = op.SortedDbArrays(c1.kmers, c1.offs, c1.ids, 20, 2)
    orc = op.Oracle(cdb=sd.cdb(), keep=sd)
    orc.set_opts(prn_all=int(o["prn_all"]), **_opts(o))
    P = c1.paths
    orc.load_files(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=P["null_lst"], lmat_dir=c1.workdir)
    hdrs, seqs = op.read_fasta_like_reference(P[reads])
    assert len(seqs) == (1000 if reads == "reads_example" else 400)
    res, _, _ = orc.label(seqs)
    assert op.assemble_lines(hdrs, seqs, orc.tails(res)) == c1.golden(f"{reads}.{oname}")


def test_host_reader_on_the_real_example_file(c1):
    """kmat_reader on simple_list.1000.fna: headers with spaces and '|', 80-column wrapped sequences, reads of 36-250 bases"""
    hdrs, seqs = op.read_fasta_like_reference(c1.paths["reads_example"])
    for threads in (1, 4):
        got_h, got_s = api.read_file(c1.paths["reads_example"], threads=threads, max_reads=300)
        assert got_h == hdrs and got_s == seqs
    assert any(" " in h for h in hdrs) and max(len(s) for s in seqs) > 200


@pytest.mark.gpu
@pytest.mark.parametrize("reads", ["reads_example", "reads_sim"])
@pytest.mark.parametrize("oname", ["run_rl", "defaults"])
def test_gpu_labels_bundled_data_like_the_reference(c1, reads, oname):
    o = S.OPTION_SETS[oname]
    db = api.Db.upload(api.Table.from_arrays(c1.kmers, c1.offs, c1.ids, 20, 2))
    P = c1.paths
    inp = api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=P["null_lst"], lmat_dir=c1.workdir)
    ctx = api.Ctx(db, inp, api.default_opts(want_lineage=0 if o["prn_all"] else 1, **_opts(o)))
    hdrs, seqs = op.read_fasta_like_reference(P[reads])
    res, cands, lin = ctx.label(seqs)
    assert op.assemble_lines(hdrs, seqs, ctx.tails(res, cands, lin, prn_all=o["prn_all"])) == c1.golden(f"{reads}.{oname}")


@pytest.mark.gpu
@pytest.mark.parametrize("reads", ["reads_example", "reads_sim"])
def test_cli_on_bundled_data(c1, reads, tmp_path):
    """The drop-in binary with the flag set of bin/run_rl.sh:243, -t 1: .out, .fastsummary and .nomatchsum byte-identical"""
    lib, exe = build.build_all()
    dbp = str(tmp_path / "c1.kmat")
    api.Table.from_arrays(c1.kmers, c1.offs, c1.ids, 20, 2).save(dbp)
    P = c1.paths
    ofb = str(tmp_path / "rl_")
    cmd = [exe, "-f", P["map16"], "-u", P["names"], "-w", P["rank"], "-x", "0", "-j", "30", "-l", "0", "-b", "1.0", "-n", P["null_lst"], "-e", P["depth"], "-p",
           "-t", "1", "-i", P[reads], "-d", dbp, "-c", P["tree"], "-o", ofb]
    e = dict(os.environ)
    e["LMAT_DIR"] = c1.workdir
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=600)
    assert p.returncode == 0, p.stderr
    assert open(ofb + "0.out", encoding="latin-1").read() == c1.golden(f"{reads}.run_rl")
    assert open(ofb + ".0.30.fastsummary").read() == open(os.path.join(GOLDEN, f"c1.{reads}.run_rl.fastsummary")).read()
    assert open(ofb + ".0.30.nomatchsum").read() == open(os.path.join(GOLDEN, f"c1.{reads}.run_rl.nomatchsum")).read()
