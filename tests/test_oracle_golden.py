"""CPU: the plain-C oracle restatement reproduces the committed reference read_label outputs
(tests/golden/*.out.gz, produced by the unmodified reference -- see tests/golden/make_golden.py)
byte for byte, for every option set."""
import numpy as np
import pytest

import scenarios as S
from conftest import oracle_for
from oracle import oracle_py as op


@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
@pytest.mark.parametrize("opts", list(S.OPTION_SETS))
def test_oracle_reproduces_reference_out(scen, opts, request):
    g = request.getfixturevalue(scen)
    orc = oracle_for(g, opts)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, _, _ = orc.label(seqs)
    assert not (res["err"] != 0).any()
    mine = op.assemble_lines(hdrs, seqs, orc.tails(res), prn_read=not S.OPTION_SETS[opts].get("hide_read"))
    assert mine == g.golden_out(opts)


@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
def test_oracle_reproduces_reference_out_under_verbose(scen, request):
    """-p -y: the candidates with a negative score are printed too (read_label.cpp:901; golden from the unmodified reference,
    tests/golden/make_golden_verbose.py)."""
    g = request.getfixturevalue(scen)
    orc = oracle_for(g, "run_rl")
    orc.set_opts(prn_all=2)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, _, _ = orc.label(seqs)
    mine = op.assemble_lines(hdrs, seqs, orc.tails(res))
    assert mine == g.golden_out("run_rl_verbose") and mine != g.golden_out("run_rl")


@pytest.mark.parametrize("tag,key,fastq", [("wrapped", "reads_wrapped", False), ("fastq", "reads_fq", True)])
def test_reader_restatement_matches_reference(golden_small, tag, key, fastq):
    g = golden_small
    orc = oracle_for(g, "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths[key], fastq=fastq)
    res, _, _ = orc.label(seqs)
    assert op.assemble_lines(hdrs, seqs, orc.tails(res)) == g.golden_out(tag)


def test_encoder_known_answers():
    """2-bit MSB-first packing and canonical min(fwd, rc) (read_label.cpp:978-1009, kencode.hpp:76-82)."""
    v, b, km, fl = op.encode_read("A" * 20, 20)
    assert v == 1 and km[0] == 0 and fl[0] == 1          # AAAA.. = 0 < TTTT.. = 2^40-1
    v, b, km, fl = op.encode_read("T" * 20, 20)
    assert km[0] == 0                                      # reverse complement of T*20 is A*20
    v, b, km, fl = op.encode_read("ACGT" * 5, 20)          # its own reverse complement
    want = int("".join("{:02b}".format("ACGT".index(c)) for c in "ACGT" * 5), 2)
    assert km[0] == want
    v, b, km, fl = op.encode_read("ACGTN" + "ACGT" * 6, 20)
    assert v == 5 and fl[:5].tolist() == [0, 0, 0, 0, 0]   # the N resets the run
    v, b, km, fl = op.encode_read("acgt" * 10, 20)
    assert v == 21 and (fl == 2).sum() == 18               # period-4 repeat: 3 distinct canonical k-mers, rest duplicates
    v, b, km, fl = op.encode_read("GC" * 30, 20)
    assert b == 10                                         # GC = 100 % -> bin 10 (read_label.cpp:1205-1206)


def test_encoder_matches_numpy_restatement():
    from lmat_b200 import fixtures as fx
    rng = np.random.default_rng(5)
    for _ in range(50):
        codes = rng.integers(0, 4, size=int(rng.integers(20, 300))).astype(np.uint8)
        want = fx.canonical_kmers(codes, 20)
        v, b, km, fl = op.encode_read(fx.codes_to_str(codes), 20)
        assert v == len(want) and np.array_equal(km, want)
