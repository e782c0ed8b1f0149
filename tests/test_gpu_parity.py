"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed
reference outputs.  Bit-exact bar: k-mer encoding, table hits and taxid lists, and the per-read output
lines byte for byte (float scores included -- the device reproduces the reference's float operation
order and glibc's logf; no tolerance is needed or used)."""
import numpy as np
import pytest

import scenarios as S
from conftest import oracle_for
from lmat_b200 import api
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dbs(golden_small, golden_lists):
    out = {}
    for g in (golden_small, golden_lists):
        t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
        out[g.name] = api.Db.upload(t)
    return out


def make_ctx(g, db, opts_name):
    o = S.OPTION_SETS[opts_name]
    P = g.paths
    inp = api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"],
                     numrank=P["numrank"] if o.get("prune") else None, plasmids=P["plasmids"] if o.get("plasmids") else None,
                     null_lst=P["null_lst"] if o["null"] else None, lmat_dir=g.workdir)
    opts = api.default_opts(min_kmer=o["min_kmer"], hbias=o["hbias"], sdiff=o["sdiff"], min_score=o["min_score"],
                            permissive=int(bool(o.get("permissive"))), phix_screen=0 if o.get("phix_off") else 1,
                            min_fnd_kmer=o.get("min_fnd", 1), max_count=o.get("prune", 65535), want_lineage=0 if o["prn_all"] else 1)
    return api.Ctx(db, inp, opts)


@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
def test_lookup_matches_oracle_and_table(scen, request, dbs):
    g = request.getfixturevalue(scen)
    db = dbs[g.name]
    assert db.size == len(g.kmers)
    rng = np.random.default_rng(3)
    q = np.concatenate([g.kmers, rng.integers(0, 1 << 40, 50000, dtype=np.uint64), g.kmers[::7] ^ np.uint64(1)])
    offs, ids = db.lookup(q)
    sd = op.SortedDbArrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    o_offs, o_ids = op.Oracle(cdb=sd.cdb(), keep=sd).lookup(q)
    assert np.array_equal(offs, o_offs) and np.array_equal(ids, o_ids)
    # every k-mer of the table is found with exactly its list, in list order
    n = len(g.kmers)
    assert np.array_equal(offs[:n + 1], g.offs) and np.array_equal(ids[:int(g.offs[-1])], g.ids)


def test_encode_matches_oracle(golden_small, dbs):
    g = golden_small
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs + ["", "A", "ACGT" * 100, "N" * 50, "acgtnACGT" * 30, "GC" * 40, fx.codes_to_str(np.arange(700) % 4)]
    kmers, flags, valid, bins, offs = dbs[g.name].encode(seqs)
    for r, s in enumerate(seqs):
        v, b, km, fl = op.encode_read(s, 20)
        o = int(offs[r])
        n = max(len(s) - 19, 0)
        assert valid[r] == v, (r, s[:30])
        assert np.array_equal(flags[o:o + n], fl), r
        assert np.array_equal(kmers[o:o + n][fl > 0], km[fl > 0]), r
        if v > 0:
            assert bins[r] == b, r


@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
@pytest.mark.parametrize("opts", list(S.OPTION_SETS))
def test_labels_match_reference_golden(scen, opts, request, dbs):
    """kmat_label_batch -> kmat_format_tail reproduces the reference read_label .out byte for byte."""
    g = request.getfixturevalue(scen)
    o = S.OPTION_SETS[opts]
    ctx = make_ctx(g, dbs[g.name], opts)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, cands, lin = ctx.label(seqs)
    assert (res["status"] != 6).all(), f"{(res['status'] == 6).sum()} reads hit an unsupported path"
    mine = op.assemble_lines(hdrs, seqs, ctx.tails(res, cands, lin, prn_all=o["prn_all"]), prn_read=not o.get("hide_read"))
    want = g.golden_out(opts)
    if mine != want:
        a, b = mine.split("\n"), want.split("\n")
        bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
        raise AssertionError(f"{len(bad)} lines differ; first: {a[bad[0]][-200:]!r} vs {b[bad[0]][-200:]!r}")


def test_labels_match_oracle_fields(golden_lists, dbs):
    """Field-level comparison with the oracle (status, counts, tid, score bits, candidate arrays)."""
    g = golden_lists
    ctx = make_ctx(g, dbs[g.name], "run_rl")
    orc = oracle_for(g, "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, cands, lin = ctx.label(seqs)
    ores, ocands, olin = orc.label(seqs)
    for f in ("status", "n1", "n2", "valid_kmers", "cand_kmer_cnt", "match", "tid", "n_cand"):
        sel = slice(None) if f in ("status", "valid_kmers") else (ores["status"] >= 2)
        assert np.array_equal(res[f][sel], ores[f][sel]), f
    lab = ores["status"] >= 4
    for f in ("score", "log_avg", "stdev"):
        assert np.array_equal(res[f][lab].view(np.uint32), ores[f][lab].view(np.uint32)), f
    for i in np.nonzero(ores["status"] == 5)[0]:
        a = cands[int(res["cand_off"][i]):int(res["cand_off"][i]) + int(res["n_cand"][i])]
        b = ocands[int(ores["cand_off"][i]):int(ores["cand_off"][i]) + int(ores["n_cand"][i])]
        assert np.array_equal(a["tid"], b["tid"]) and np.array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32)), i


def test_batch_split_invariance(golden_small, dbs):
    """Size-independent property: labels do not depend on how reads are batched."""
    g = golden_small
    ctx = make_ctx(g, dbs[g.name], "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, cands, lin = ctx.label(seqs)
    whole = ctx.tails(res, cands, lin)
    parts = []
    for a in range(0, len(seqs), 37):
        r, c, l = ctx.label(seqs[a:a + 37])
        parts += ctx.tails(r, c, l)
    assert parts == whole


@pytest.mark.parametrize("opts", ["run_rl", "permissive", "prune3"])
def test_long_and_ragged_reads_match_oracle(golden_lists, dbs, opts):
    """Variable-length batch (BASELINE config 5 shape): reads from 0 to 12 kbp in one call exercise the three
    position-mask paths of the candidate kernel (<= 160 positions in registers, <= 320, global scratch) and the
    probe kernel's global dedup sets; compared line for line with the oracle."""
    g = golden_lists
    inp = S.build_inputs("lists", g.workdir + "/long_" + opts)
    rng = np.random.default_rng(11)
    names = list(inp["genomes"])
    seqs = []
    for L in [0, 5, 19, 20, 21, 150, 179, 180, 181, 250, 339, 340, 341, 700, 2999, 6000, 12000]:
        for rep in range(3):
            parts = []
            while sum(map(len, parts)) < L:
                gsel = fx.codes_to_str(inp["genomes"][names[int(rng.integers(0, 4))]])      # a few related genomes only
                a = int(rng.integers(0, max(1, len(gsel) - 500)))
                parts.append(gsel[a:a + int(rng.integers(100, 2500))])
            s = "".join(parts)[:L]
            if L > 100 and rep == 1:
                s = s[:60] + "N" + s[61:]
            if L > 100 and rep == 2:
                s = s.lower()
            seqs.append(s)
    ctx = make_ctx(g, dbs[g.name], opts)
    orc = oracle_for(g, opts)
    res, cands, lin = ctx.label(seqs)
    ores, _, _ = orc.label(seqs)
    assert (res["status"] != 6).all()
    assert ctx.tails(res, cands, lin, prn_all=True) == orc.tails(ores)


@pytest.mark.parametrize("opts", ["run_rl", "permissive", "prune3", "defaults"])
def test_reads_with_more_than_64_candidates_match_oracle(golden_lists, dbs, opts):
    """Chimeric reads stitched from many genomes carry more candidate taxids than the warp kernel's 64 register slots:
    they take the big kernels (km_cand_big_kernel + km_score_big_kernel, up to 512 candidates) and must come out exactly
    like the oracle's -- short reads (the <= 160-position kernel), medium and long ones (global position masks)."""
    g = golden_lists
    inp = S.build_inputs("lists", g.workdir + "/big_" + opts)
    rng = np.random.default_rng(23)
    names = list(inp["genomes"])
    gstr = [fx.codes_to_str(inp["genomes"][n]) for n in names]
    seqs = []
    for L, piece in [(150, 21), (150, 24), (160, 22), (300, 23), (900, 25), (4000, 30), (12000, 40)]:
        for rep in range(6):
            parts = []
            while sum(map(len, parts)) < L:
                gs = gstr[int(rng.integers(0, len(gstr)))]
                a = int(rng.integers(0, len(gs) - piece))
                parts.append(gs[a:a + piece])
            seqs.append("".join(parts)[:L])
    seqs += [gstr[0][:150], gstr[1][100:400]]              # ordinary reads in the same batch
    ctx = make_ctx(g, dbs[g.name], opts)
    orc = oracle_for(g, opts)
    res, cands, lin = ctx.label(seqs)
    ores, ocands, _ = orc.label(seqs)
    big = ores["n_cand"] > 64
    assert big.sum() >= 12, (int(big.sum()), ores["n_cand"].tolist())
    if opts != "prune3":
        assert (ores["n_cand"][:18] > 64).any()            # also among the 150 / 160-base reads (the register-mask kernel)
    assert (res["status"] != 6).all(), res["err"][res["status"] == 6]
    assert np.array_equal(res["n_cand"], ores["n_cand"])
    assert ctx.tails(res, cands, lin, prn_all=S.OPTION_SETS[opts]["prn_all"]) == orc.tails(ores)


def _sample_reads(g, workdir, lengths, seed, reps=4):
    inp = S.build_inputs("lists", workdir)
    rng = np.random.default_rng(seed)
    names = list(inp["genomes"])
    seqs = []
    for L in lengths:
        for rep in range(reps):
            gsel = fx.codes_to_str(inp["genomes"][names[int(rng.integers(0, len(names)))]])
            a = int(rng.integers(0, max(1, len(gsel) - 400)))
            s = gsel[a:a + L]
            if L > 60 and rep == 1:
                s = s[:40] + "N" + s[41:]
            if L > 60 and rep == 2:
                s = (s[:L // 2] + s[:L // 2])[:L]          # repeated half: duplicate k-mers across chunks
            if rep == 3:
                s = s.lower()
            seqs.append(s)
    return seqs


@pytest.mark.parametrize("lengths", [[0, 5, 19, 20, 21, 52, 100, 150, 159, 160], [30, 161, 200, 255, 256]], ids=["nch5", "nch8"])
@pytest.mark.parametrize("variant", ["fast", "streaming"])
def test_short_read_probe_kernels_match_oracle(golden_lists, dbs, lengths, variant, monkeypatch):
    """The register-resident encode+probe kernel (<= 160 / <= 256 bases) and the streaming one give the oracle's lines."""
    g = golden_lists
    if variant == "streaming":
        monkeypatch.setenv("KMAT_NO_FAST_PROBE", "1")
    seqs = _sample_reads(g, g.workdir + f"/short_{max(lengths)}_{variant}", lengths, 5)
    ctx = make_ctx(g, dbs[g.name], "run_rl")
    orc = oracle_for(g, "run_rl")
    res, cands, lin = ctx.label(seqs)
    ores, _, _ = orc.label(seqs)
    assert (res["status"] != 6).all()
    assert np.array_equal(res["valid_kmers"], ores["valid_kmers"])
    assert ctx.tails(res, cands, lin, prn_all=True) == orc.tails(ores)
    st = ctx.stats()
    assert st.lookups > 0 and st.hits > 0


def test_tight_table_uses_displacement_and_stash(golden_lists, monkeypatch):
    """A nearly full table (test knob) pushes keys through every displacement and into the overflow stash: lookups
    and labels must not change."""
    g = golden_lists
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    roomy = api.Db.upload(t)
    monkeypatch.setenv("KMAT_TEST_TIGHT_TABLE", "1")
    tight = api.Db.upload(t)
    monkeypatch.delenv("KMAT_TEST_TIGHT_TABLE")
    assert tight.bytes < roomy.bytes
    rng = np.random.default_rng(9)
    q = np.concatenate([g.kmers, rng.integers(0, 1 << 40, 50000, dtype=np.uint64)])
    a_offs, a_ids = roomy.lookup(q)
    b_offs, b_ids = tight.lookup(q)
    assert np.array_equal(a_offs, b_offs) and np.array_equal(a_ids, b_ids)
    n = len(g.kmers)
    assert np.array_equal(b_offs[:n + 1], g.offs) and np.array_equal(b_ids[:int(g.offs[-1])], g.ids)
    ctx = make_ctx(g, tight, "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, cands, lin = ctx.label(seqs)
    mine = op.assemble_lines(hdrs, seqs, ctx.tails(res, cands, lin, prn_all=True))
    assert mine == g.golden_out("run_rl")
    assert ctx.stats().probe_extra_buckets > 0


def test_pipelined_pass_matches_serial(golden_lists, dbs):
    """kmat_ctx_set_pipeline: overlapping the probe kernel of one sub-batch with the candidate / scoring kernels of the
    previous one (two streams) must not change a single result; candidates are addressed through cand_off."""
    g = golden_lists
    ctx = make_ctx(g, dbs[g.name], "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs * 3
    ctx.set_pipeline(1)
    res, cands, lin = ctx.label(seqs)
    serial = ctx.tails(res, cands, lin, prn_all=True)
    for sub in (2, 5, 16):
        ctx.set_pipeline(sub)
        r2, c2, l2 = ctx.label(seqs)
        assert ctx.tails(r2, c2, l2, prn_all=True) == serial, sub
        assert np.array_equal(r2["status"], res["status"]) and np.array_equal(r2["n_cand"], res["n_cand"])


@pytest.mark.parametrize("long_kernel", [True, False])
def test_long_reads_with_internal_repeats_match_oracle(golden_lists, dbs, long_kernel, monkeypatch):
    """10 kbp-class reads full of internal repeats, N runs and case changes: the CTA-per-read probe kernel (dedup set in
    shared memory, lowest position wins) and the any-length kernel (global set, chunks in order) must both reproduce the
    oracle's first-occurrence semantics."""
    if not long_kernel:
        monkeypatch.setenv("KMAT_NO_LONG_PROBE", "1")
    g = golden_lists
    inp = S.build_inputs("lists", g.workdir + "/rep")
    rng = np.random.default_rng(31)
    gstr = [fx.codes_to_str(x) for x in inp["genomes"].values()]
    seqs = []
    for L in [257, 300, 1000, 2999, 5000, 9999, 10000, 11999, 12000, 12001, 20000]:
        for rep in range(4):
            gs = gstr[int(rng.integers(0, 6))]
            parts = []
            while sum(map(len, parts)) < L:
                a = int(rng.integers(0, len(gs) - 400))
                piece = gs[a:a + int(rng.integers(25, 400))]
                parts.append(piece)
                if rng.random() < 0.5:
                    parts.append(parts[int(rng.integers(0, len(parts)))])          # an earlier piece again
            s = "".join(parts)[:L]
            if rep == 1:
                s = s[:100] + "N" * 3 + s[103:L // 2] + "n" + s[L // 2 + 1:]
            if rep == 2:
                s = s.lower()
            if rep == 3:
                s = s[:L // 2] + s[:L - L // 2]                                     # the first half twice
            seqs.append(s)
    ctx = make_ctx(g, dbs[g.name], "run_rl")
    orc = oracle_for(g, "run_rl")
    res, cands, lin = ctx.label(seqs)
    ores, _, _ = orc.label(seqs)
    assert (res["status"] != 6).all()
    for f in ("valid_kmers", "cand_kmer_cnt", "bin_sel", "n_cand"):
        sel = ores["status"] >= 2
        assert np.array_equal(res[f][sel], ores[f][sel]), (f, np.nonzero(res[f][sel] != ores[f][sel])[0][:5])
    assert ctx.tails(res, cands, lin, prn_all=True) == orc.tails(ores)
