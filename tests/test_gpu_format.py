"""GPU (-m gpu): K5, the output-line tails formatted on the device (kmat_label_batch_text, kmat_format.cuh), against the host
formatter kmat_format_tail -- which tests/test_abi_cpu.py pins to printf("%g") and tests/test_cli.py / the goldens to the
reference's own lines.  Every read either carries the same bytes or is handed back to the host formatter."""
import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api
from lmat_b200 import fixtures as fx
from oracle import oracle_py as op
from test_gpu_parity import make_ctx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
@pytest.mark.parametrize("opts", ["run_rl", "defaults", "tight", "permissive", "prune3", "nophix_hide", "quirk", "nonull"])
def test_device_tails_equal_host_tails(request, scen, opts):
    g = request.getfixturevalue(scen)
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes), 0)
    ctx = make_ctx(g, db, opts)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    prn_all = S.OPTION_SETS[opts]["prn_all"]
    res, cands, lin = ctx.label(seqs)
    want = ctx.tails(res, cands, lin, prn_all=prn_all)
    res2, cands2, lin2, tails, on_host = ctx.label_text(seqs, prn_all=prn_all)
    for f in res.dtype.names:                                        # cand_off / lin_off depend on the order the warps took their room
        if f not in ("cand_off", "lin_off"):
            assert np.array_equal(res[f], res2[f]), f
    assert on_host <= len(seqs) // 20, on_host                       # the device formats (nearly) everything
    for i, t in enumerate(tails):
        if t is not None:
            assert t == want[i], (i, t, want[i])
    # and the assembled file is the reference's golden output
    if opts in ("run_rl",):
        full = [t if t is not None else want[i] for i, t in enumerate(tails)]
        assert op.assemble_lines(hdrs, seqs, full) == g.golden_out(opts)
        # -p together with -y (prn_all = 2): candidates with a negative score are printed too (read_label.cpp:901)
        res3, cands3, lin3, tails3, _ = ctx.label_text(seqs, prn_all=2)
        want3 = ctx.tails(res3, cands3, lin3, prn_all=2)
        full3 = [t if t is not None else want3[i] for i, t in enumerate(tails3)]
        assert all(t is None or t == want3[i] for i, t in enumerate(tails3))
        assert op.assemble_lines(hdrs, seqs, full3) == g.golden_out("run_rl_verbose")


def test_device_tails_many_reads_and_small_buffers(golden_lists, monkeypatch):
    """Several chunks per call (first and last a quarter of the regular size), a text buffer that runs out and one of zero
    bytes: what does not fit is left to the host, nothing is cut."""
    monkeypatch.setenv("KMAT_CHUNK_READS", "7000")
    g = golden_lists
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes), 0)
    ctx = make_ctx(g, db, "run_rl")
    inp = S.build_inputs("lists", g.workdir + "/fmt_many")
    hdrs, seqs = fx.simulate_reads(91, inp["genomes"], 60000, 150, n_rate=0.002, lower_frac=0.05, len_jitter=60)
    res, cands, lin = ctx.label(seqs)
    want = ctx.tails(res, cands, lin, prn_all=True)
    for cap in (None, 1 << 20, 0):
        _, _, _, tails, on_host = ctx.label_text(seqs, prn_all=True, text_cap=cap)
        if cap is None:
            assert on_host < len(seqs) // 50
        if cap == 0:
            assert on_host == len(seqs)
        if cap == 1 << 20:
            assert 0 < on_host < len(seqs)
        assert all(t is None or t == want[i] for i, t in enumerate(tails))


def test_device_float_text_equals_printf():
    """kf_g (kmat_format.cuh) against printf("%g") -- Python's % operator on the promoted double, which is what ostream << float
    prints -- on random floats of every binade it takes (2^-70 .. 999999), the boundaries of its two ranges, ties and zeros; and
    that what it does not take is handed back, never printed wrong."""
    rng = np.random.default_rng(3)
    bits = np.concatenate([
        rng.integers((127 - 70) << 23, 0x497423F0, size=3_000_000, dtype=np.int64),                 # the whole supported range, uniform over bit patterns
        rng.integers(0x38D1B718 - 5000, 0x38D1B718 + 5000, size=10_000, dtype=np.int64),            # around 1e-4
        rng.integers(0x497423F0 - 5000, 0x497423F0 + 5000, size=10_000, dtype=np.int64),            # around 999999
        rng.integers((127 - 70) << 23, ((127 - 70) << 23) + 5000, size=5_000, dtype=np.int64),      # the smallest values taken
        rng.integers(0, (127 - 70) << 23, size=5_000, dtype=np.int64),                              # below: left to the host
        np.array([0, 0x7F800000, 0x7FC00000, 0x3F800000, 0x3DCCCCCD, 0x42C80000], dtype=np.int64),  # 0, inf, nan, 1, 0.1, 100
    ]).astype(np.uint32)
    bits[::2] |= np.uint32(0x80000000)                                                              # every other one negative
    vals = bits.view(np.float32)
    # decimal ties: k + 0.5 at the sixth digit, exactly representable
    ties = np.array([100000.5, 100001.5, 12345.25, 12345.75, 0.5, 0.25, 0.125, 2.5, 1234.5], dtype=np.float32)
    vals = np.concatenate([vals, ties, -ties])
    out = np.zeros((len(vals), 16), dtype=np.uint8)
    rc = api.lib().kmat_test_format_floats(0, vals.ctypes.data, len(vals), out.ctypes.data)
    assert rc == 0
    on_host = 0
    raw = out.tobytes()
    for i, v in enumerate(vals.tolist()):
        t = raw[16 * i:16 * i + 16]
        if t[0] == 0xFF:
            on_host += 1
            a = abs(v)
            assert not (2.0 ** -70 <= a < 999999.0) or a != a, v              # only what is outside the two ranges
            continue
        assert t.rstrip(b"\0").decode() == "%g" % v, (v, t)
    assert on_host < 30_000
