"""GPU (-m gpu): reads with more candidate taxids than the shared-memory kernel holds (512).  The reference has no bound
(std::set / std::map per read, read_label.cpp:698-726); libkmat hands such reads from km_cand_kernel (64 candidates, registers)
to km_cand_big_kernel (512, shared memory) to km_cand_huge_kernel / km_score_huge_kernel (16,384, global memory, hashed).
A wide taxonomy (800 leaves, ~1,650 nodes) and chimeric reads stitched from hundreds of genomes; the checker is the oracle
(pinned to the unmodified reference by tests/test_oracle_golden.py)."""
import types

import numpy as np
import pytest

import scenarios as S
from conftest import oracle_for
from lmat_b200 import api
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op
from test_gpu_parity import make_ctx

WIDE = dict(seed=47, n_leaves=800, genome_len=1200, share=0.2, cons_rank="order", cons_len=300)


def wide_scenario(workdir):
    """A GoldenScenario look-alike (the attributes make_ctx / oracle_for read) over a table computed here."""
    tax = fx.make_taxonomy(WIDE["seed"], WIDE["n_leaves"], specials=True)
    paths = fx.write_taxonomy_files(tax, workdir)
    genomes = fx.make_genomes(WIDE["seed"] + 1, tax, WIDE["genome_len"], share_frac=WIDE["share"], conserved_rank=WIDE["cons_rank"],
                              conserved_len=WIDE["cons_len"])
    kmers, offs, tids = fx.build_kmer_table(genomes, tax, S.K)
    m16 = fx.map16(tax)
    g = types.SimpleNamespace()
    g.name, g.workdir, g.kmer_len, g.tid_bytes = "wide", workdir, S.K, 2
    g.kmers, g.offs, g.ids = kmers, offs, np.array([m16[int(t)] for t in tids], dtype=np.uint32)
    paths["null_lst"] = fx.write_null_models(WIDE["seed"] + 3, tax, workdir)
    paths["plasmids"] = None
    g.paths = paths
    g.genomes = genomes
    return g


def chimeric_reads(genomes, seed):
    rng = np.random.default_rng(seed)
    gstr = [fx.codes_to_str(c) for c in genomes.values()]
    seqs = []
    for L, piece, reps in [(150, 21, 3), (9000, 21, 3), (12000, 22, 3), (12000, 30, 2), (30000, 21, 1)]:
        for _ in range(reps):
            parts = []
            while sum(map(len, parts)) < L:
                gs = gstr[int(rng.integers(0, len(gstr)))]
                a = int(rng.integers(0, len(gs) - piece))
                parts.append(gs[a:a + piece])
            seqs.append("".join(parts)[:L])
    seqs += [gstr[0][:150], gstr[1][100:400], gstr[2][:1000]]         # ordinary reads in the same batch
    return seqs


@pytest.mark.gpu
@pytest.mark.parametrize("opts", ["run_rl", "permissive", "prune3", "defaults"])
def test_reads_with_more_than_512_candidates_match_oracle(tmp_path, opts):
    g = wide_scenario(str(tmp_path))
    seqs = chimeric_reads(g.genomes, 29)
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes), 0)
    ctx = make_ctx(g, db, opts)
    orc = oracle_for(g, opts)
    ores, ocands, _ = orc.label(seqs)
    assert (ores["n_cand"] > 512).sum() >= 4, ores["n_cand"].tolist()            # the last-resort kernels
    if opts != "prune3":
        assert ((ores["n_cand"] > 64) & (ores["n_cand"] <= 512)).any(), ores["n_cand"].tolist()   # and the shared-memory ones
    res, cands, lin = ctx.label(seqs)
    assert (res["status"] != 6).all(), res["err"][res["status"] == 6]
    assert np.array_equal(res["n_cand"], ores["n_cand"])
    assert ctx.tails(res, cands, lin, prn_all=S.OPTION_SETS[opts]["prn_all"]) == orc.tails(ores)
    # the same reads through the device-resident entry point in a second pass: the queues and scratch slots are reused
    res2, cands2, lin2 = ctx.label(seqs[::-1])
    assert ctx.tails(res2, cands2, lin2, prn_all=S.OPTION_SETS[opts]["prn_all"]) == orc.tails(ores)[::-1]
