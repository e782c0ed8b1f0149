"""GPU (-m gpu): the two-level table (minimizer-ordered lines + bucket table behind the sector overflow flags,
lmat_b200/csrc/kmat_mzr.h) for every k-mer length class -- k below 17 and above 23 (no first level), 17 .. 23 (2 to 8
minimizer windows) -- roomy and tight: every stored k-mer is found with exactly its list, absent k-mers miss; the fast
probe kernel's keys (one hash per base + a sliding minimum over the lanes) find what the definition-based lookup finds;
and a k = 18 scenario labels exactly like the oracle (the reference supports k = 18 and 20, SortedDb.hpp:188-200)."""
import os

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


def _scenario(workdir, k, seed=41, n_leaves=30, genome_len=3000):
    tax = fx.make_taxonomy(seed, n_leaves, specials=True)
    paths = fx.write_taxonomy_files(tax, workdir)
    genomes = fx.make_genomes(seed + 1, tax, genome_len, share_frac=0.3, conserved_rank="order", conserved_len=400)
    kmers, offs, tids = fx.build_kmer_table(genomes, tax, k)
    m16 = fx.map16(tax)
    ids = np.array([m16[int(t)] for t in tids], dtype=np.uint32)
    hdrs, seqs = fx.simulate_reads(seed + 2, genomes, 500, 150, n_rate=0.003, lower_frac=0.1, len_jitter=40)
    paths["null_lst"] = fx.write_null_models(seed + 3, tax, workdir)
    return tax, paths, genomes, np.asarray(kmers, dtype=np.uint64), np.asarray(offs, dtype=np.uint64), ids, hdrs, seqs


def _first_occurrence_stats(seqs, k, table):
    """(lookups, hits) the probe kernel must count: first occurrences of valid canonical k-mers per read"""
    look = hits = 0
    for s in seqs:
        v, b, km, fl = op.encode_read(s, k)
        first = km[fl == 1]
        look += len(first)
        hits += int(np.isin(first, table).sum())
    return look, hits


@pytest.mark.parametrize("k", [12, 16, 17, 18, 19, 20, 21, 22, 23, 25])
def test_lookup_every_k(k, tmp_path, monkeypatch):
    tax, paths, genomes, kmers, offs, ids, hdrs, seqs = _scenario(str(tmp_path), k, seed=40 + k)
    t = api.Table.from_arrays(kmers, offs, ids, k, 2)
    rng = np.random.default_rng(k)
    absent = rng.integers(0, 1 << (2 * k), 60000, dtype=np.uint64)
    absent = absent[~np.isin(absent, kmers)]
    near = (kmers[::5] ^ np.uint64(1))
    near = near[~np.isin(near, kmers)]                  # one base off a stored k-mer: same minimizer, same line
    q = np.concatenate([kmers, absent, near])
    want_offs = np.concatenate([offs, np.full(len(absent) + len(near), offs[-1], dtype=np.uint64)])
    for tight in (False, True):
        if tight:
            monkeypatch.setenv("KMAT_TEST_TIGHT_TABLE", "1")
        db = api.Db.upload(t)
        monkeypatch.delenv("KMAT_TEST_TIGHT_TABLE", raising=False)
        assert db.size == len(kmers)
        if tight and 17 <= k <= 23:
            assert db.overflow > 0, "the tight table was meant to push k-mers into the second level"
        g_offs, g_ids = db.lookup(q)
        assert np.array_equal(g_offs, want_offs) and np.array_equal(g_ids, ids)
    # the single-level table (what k outside 17..23 uses) on a k that has a first level: same answers
    monkeypatch.setenv("KMAT_NO_LINE_LEVEL", "1")
    db1 = api.Db.upload(t)
    monkeypatch.delenv("KMAT_NO_LINE_LEVEL")
    assert db1.overflow == 0
    g_offs, g_ids = db1.lookup(q)
    assert np.array_equal(g_offs, want_offs) and np.array_equal(g_ids, ids)


@pytest.mark.parametrize("k", [17, 18, 20, 21, 23, 24])
@pytest.mark.parametrize("tight", [False, True])
def test_fast_kernel_finds_what_the_definition_finds(k, tight, tmp_path, monkeypatch):
    """kmat_label_batch on reads <= 160 / <= 256 bases runs km_encode_probe_fast_kernel, whose table keys come from the
    sliding minimum; its lookup / hit counters must equal the first occurrences found in the table by plain set arithmetic."""
    tax, paths, genomes, kmers, offs, ids, hdrs, seqs = _scenario(str(tmp_path), k, seed=70 + k)
    if tight:
        monkeypatch.setenv("KMAT_TEST_TIGHT_TABLE", "1")
    db = api.Db.upload(api.Table.from_arrays(kmers, offs, ids, k, 2))
    monkeypatch.delenv("KMAT_TEST_TIGHT_TABLE", raising=False)
    inp = api.Inputs(tree=paths["tree"], depth=paths["depth"], rank=paths["rank"], map16=paths["map16"], null_lst=paths["null_lst"], lmat_dir=str(tmp_path))
    ctx = api.Ctx(db, inp, api.default_opts(min_kmer=30, hbias=0.0, sdiff=1.0))
    for batch in ([s[:150] for s in seqs], [s + s[:60] for s in seqs[:200]]):          # <= 160 bases (5 chunks) and <= 256 bases (8 chunks)
        assert max(len(s) for s in batch) <= 256
        res, cands, lin = ctx.label(batch)
        assert (res["status"] != 6).all()
        st = ctx.stats()
        look, hits = _first_occurrence_stats(batch, k, kmers)
        assert (st.lookups, st.hits) == (look, hits)


def test_k18_labels_match_oracle(tmp_path):
    k = 18
    tax, paths, genomes, kmers, offs, ids, hdrs, seqs = _scenario(str(tmp_path), k, seed=118, n_leaves=40)
    db = api.Db.upload(api.Table.from_arrays(kmers, offs, ids, k, 2))
    inp = api.Inputs(tree=paths["tree"], depth=paths["depth"], rank=paths["rank"], map16=paths["map16"], null_lst=paths["null_lst"], lmat_dir=str(tmp_path))
    ctx = api.Ctx(db, inp, api.default_opts(min_kmer=30, hbias=0.0, sdiff=1.0))
    sd = op.SortedDbArrays(kmers, offs, ids, k, 2)
    orc = op.Oracle(cdb=sd.cdb(), keep=sd)
    orc.set_opts(min_kmer=30, hbias=0.0, sdiff=1.0, prn_all=1)
    orc.load_files(tree=paths["tree"], depth=paths["depth"], rank=paths["rank"], map16=paths["map16"], null_lst=paths["null_lst"], lmat_dir=str(tmp_path))
    long_reads = [fx.codes_to_str(next(iter(genomes.values())))[:2500], "ACGT" * 200]      # the CTA-per-read and any-length kernels
    for batch in (seqs, seqs[:50] + long_reads):
        res, cands, lin = ctx.label(batch)
        ores, _, _ = orc.label(batch)
        assert np.array_equal(res["valid_kmers"], ores["valid_kmers"])
        assert ctx.tails(res, cands, lin, prn_all=True) == orc.tails(ores)
