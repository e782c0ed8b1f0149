"""The compact interface (include/kmat.h: kmat_pack_reads / kmat_label_batch_packed / kmat_result_expand): 2-bit packed reads
in, 32-byte results out.  CPU: the host packer against a plain restatement, any thread count.  GPU: the same output lines as
kmat_label_batch -- and therefore as the reference -- for every option set that changes what the one pair list holds."""
import ctypes as C

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api
from oracle import oracle_py as op


def _pack(blob, threads):
    L = api.lib()
    total = len(blob)
    codes = np.zeros(max(1, L.kmat_pack_words(total)), dtype=np.uint32)
    inv = np.zeros(total + 1, dtype=np.uint64)
    n_inv = C.c_uint64()
    assert L.kmat_pack_reads(blob, total, threads, codes.ctypes.data, inv.ctypes.data, len(inv), C.byref(n_inv)) == 0
    return codes, inv[:n_inv.value]


def test_pack_reads_matches_restatement():
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNnRY-*\x00\xff", dtype=np.uint8)
    p = np.array([20, 20, 20, 20, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1], dtype=float)
    for total in (0, 1, 15, 16, 17, 1000, 70001, 300000):
        blob = alphabet[rng.choice(len(alphabet), size=total, p=p / p.sum())].tobytes()
        want_codes = np.zeros(max(1, (total + 15) // 16), dtype=np.uint32)
        want_inv = []
        lut = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2, ord("T"): 3, ord("t"): 3}
        for i, ch in enumerate(blob):
            if ch in lut:
                want_codes[i // 16] |= np.uint32(lut[ch] << (2 * (i % 16)))
            else:
                want_inv.append(i)
        for threads in (1, 3, 8):
            codes, inv = _pack(blob, threads)
            assert np.array_equal(codes[:len(want_codes)], want_codes) and inv.tolist() == want_inv
    L = api.lib()                                              # too small an invalid-base buffer: the count needed comes back
    n_inv = C.c_uint64()
    codes = np.zeros(4, dtype=np.uint32)
    inv = np.zeros(1, dtype=np.uint64)
    assert L.kmat_pack_reads(b"NNNNACGT", 8, 1, codes.ctypes.data, inv.ctypes.data, 1, C.byref(n_inv)) == -10 and n_inv.value == 4


def test_result_expand_restores_the_line_integers():
    L = api.lib()
    r32 = np.zeros(1, dtype=api.RESULT32_DTYPE)
    out = np.zeros(1, dtype=api.RESULT_DTYPE)
    for status, want in ((0, (12, 20, 0)), (1, (17, 30, 17)), (2, (150, 20, 17))):       # ReadTooShort (len), ReadTooShort (valid), NoDbHits
        r32["flags"] = status | (3 << 3) | (7 << 6) | (5 << 10)
        r32["valid_kmers"] = 17
        r32["list_off"], r32["n_list"] = 99, 4
        L.kmat_result_expand(r32.ctypes.data, 12 if status == 0 else 150, 20, 30, 0, out.ctypes.data)
        assert (int(out["n1"][0]), int(out["n2"][0]), int(out["valid_kmers"][0])) == want
        assert int(out["status"][0]) == status and int(out["match"][0]) == 3 and int(out["bin_sel"][0]) == 7 and int(out["err"][0]) == -5
        assert int(out["cand_off"][0]) == 99 and int(out["n_cand"][0]) == 4 and int(out["n_lin"][0]) == 0
    L.kmat_result_expand(r32.ctypes.data, 150, 20, 30, 1, out.ctypes.data)
    assert int(out["lin_off"][0]) == 99 and int(out["n_lin"][0]) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
@pytest.mark.parametrize("opts", ["run_rl", "defaults", "permissive", "prune3", "quirk", "nonull"])
def test_packed_labels_equal_ascii_labels(scen, opts, request):
    from test_gpu_parity import make_ctx
    g = request.getfixturevalue(scen)
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes))
    ctx = make_ctx(g, db, opts)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs + ["", "A", "N" * 50, "acgtnACGT" * 30, seqs[0] * 4, seqs[1][:33]]          # empty, short, all-N, mixed case, long (CTA-per-read kernel)
    hdrs = hdrs + [f"x{i}" for i in range(6)]
    prn_all = S.OPTION_SETS[opts]["prn_all"]
    res, cands, lin = ctx.label(seqs)
    want = ctx.tails(res, cands, lin, prn_all=prn_all)
    r2, c2, l2 = ctx.label_packed(seqs)
    got = ctx.tails(r2, c2, l2, prn_all=prn_all)
    assert got == want
    assert np.array_equal(res["valid_kmers"], r2["valid_kmers"]) and np.array_equal(res["status"], r2["status"])
    if scen == "golden_small" and opts in ("run_rl",):
        assert op.assemble_lines(hdrs[:-6], seqs[:-6], got[:-6]) == g.golden_out(opts)
    # the run-length form of the lists (kmat_label_batch_packed_rl + kmat_list_decode): the same lines from fewer words
    r3, c3, l3, n_words = ctx.label_packed_rl(seqs)
    assert ctx.tails(r3, c3, l3, prn_all=prn_all) == want
    n_pairs = len(c3) + len(l3)                         # the pairs the records refer to (the plain interface also ships the orphaned
    ref2 = int((r2["n_lin"] if S.OPTION_SETS[opts]["prn_all"] is False else r2["n_cand"]).astype(np.int64).sum())   # pairs of PhiX / silent reads)
    assert n_pairs == ref2 <= len(c2) + len(l2) and (n_pairs == 0 or n_pairs < n_words < 2 * n_pairs)


def test_list_decode():
    """kmat_list_decode: a taxid word with bit 31 set is followed by the score of it and of the unflagged taxids after it."""
    f = lambda x: int(np.array([x], dtype=np.float32).view(np.uint32)[0])
    words = np.array([0x80000000 | 562, f(1.5), 561, 543, 0x80000000 | 1224, f(-0.25), 0x80000000 | 2, f(1.5), 1], dtype=np.uint32)
    out = np.zeros(6, dtype=api.PAIR_DTYPE)
    used = api.lib().kmat_list_decode(words.ctypes.data, 6, out.ctypes.data)
    assert used == 9
    assert out["tid"].tolist() == [562, 561, 543, 1224, 2, 1] and out["score"].tolist() == [1.5, 1.5, 1.5, -0.25, 1.5, 1.5]
    assert api.lib().kmat_list_decode(words.ctypes.data, 0, out.ctypes.data) == 0
