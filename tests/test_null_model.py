"""Null-model generation: rand_read_label (src/rand_read_label.cpp + src/rkmer.hpp).

CPU: the glibc rand() restatement against libc itself; the oracle's rand_read_label restatement against the .rand_lst
files the UNMODIFIED reference wrote under a fixed time(0) seed (tests/golden/make_nullgen_golden.py); the host-side
merge / writer (kmat_null_write).
GPU (-m gpu): kmat_null_batch over the reference's own reads against those goldens; the rand_read_label drop-in binary
(KMAT_RAND_COMPAT=glibc) byte for byte; the device read generator against its restatement here, and kmat_null_random
against the oracle on a table built from the device-drawn reads."""
import ctypes as C
import json
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
def nullgen(tmp_path_factory):
    wd = str(tmp_path_factory.mktemp("nullgen"))
    inp = S.build_nullgen_inputs(wd)
    man = json.load(open(os.path.join(GOLDEN, "nullgen.manifest.json")))
    import hashlib
    for k, want in man["inputs"].items():
        assert hashlib.sha256(open(inp["paths"][k], "rb").read()).hexdigest() == want, f"seeded nullgen input {k} drifted"
    assert man["time"] == S.NULLGEN_TIME
    t = np.load(os.path.join(GOLDEN, "nullgen.table.npz"))
    inp["table"] = (t["kmers"], t["offs"], t["ids"].astype(np.uint32), int(t["kmer_len"]), int(t["tid_bytes"]))
    return inp


def golden_text(tag):
    return open(os.path.join(GOLDEN, f"nullgen.{tag}.rand_lst")).read()


def make_oracle(table, P, prune):
    sd = op.SortedDbArrays(*table)
    orc = op.Oracle(cdb=sd.cdb(), keep=sd)
    orc.set_opts(max_count=prune or 65535)
    orc.load_files(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], numrank=P["numrank"] if prune else None)
    return orc


def make_ctx(table, P, prune):
    db = api.Db.upload(api.Table.from_arrays(*table), 0)
    inp = api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], numrank=P["numrank"] if prune else None)
    return api.Ctx(db, inp, api.default_opts(rkmer_mode=1, max_count=prune or 65535))


# ------------------------------------------------------------------------------------------------ CPU
def test_glibc_rand_restatement_equals_libc():
    libc = C.CDLL(None)
    L = op.lib()
    for seed in (0, 1, 42, S.NULLGEN_TIME, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF):
        g = op.GlibcRand()
        L.kmo_srand(C.byref(g), C.c_uint(seed))
        libc.srand(C.c_uint(seed))
        assert [L.kmo_rand(C.byref(g)) for _ in range(3000)] == [libc.rand() for _ in range(3000)], seed


def test_gen_rand_reads_shape():
    reads = op.gen_rand_reads(S.NULLGEN_TIME, 200, 150)
    for i, r in enumerate(reads):
        assert len(r) == 150 and set(r) <= set("acgt")
        gc = sum(ch in "gc" for ch in r)
        b = i % 10
        # num_gc = (unsigned)((float)(gc_draw / 100.0) * 150) with gc_draw in [10 b, 10 b + 9]
        lo = int(np.float32(np.float32(b * 10 / 100.0) * np.float32(150)))
        hi = int(np.float32(np.float32((b * 10 + 9) / 100.0) * np.float32(150)))
        assert lo <= gc <= hi, (i, gc)
    # the generator is a stream: a later start is a different run, the same start the same run
    assert op.gen_rand_reads(S.NULLGEN_TIME, 50, 150) == reads[:50]


@pytest.mark.parametrize("tag", list(S.NULLGEN_RUNS))
def test_oracle_equals_reference_rand_lst(nullgen, tag):
    run = S.NULLGEN_RUNS[tag]
    orc = make_oracle(nullgen["table"], nullgen["paths"], run["prune"])
    reads = op.gen_rand_reads(S.NULLGEN_TIME, run["n_reads"], run["read_len"])
    # in two batches: the accumulators persist, the bucket follows the run index
    orc.null_batch(reads[:237], 0)
    orc.null_batch(reads[237:], 237)
    t, m, c = orc.null_table()
    assert op.format_rand_lst(t, m, c) == golden_text(tag)
    # the fixture does what it is there for: fractional hits, human ids kept apart (no collapse in rkmer.hpp)
    assert len(t) > 100 and {9606, 63221, 741158} <= set(t.tolist())
    assert ((m > 0) & (m < 1)).sum() > 500


def test_null_write_merges_by_max_and_sum(tmp_path):
    a = (np.array([5, 9, 100], np.uint32), np.full((3, 10), 0.25, np.float32), np.full((3, 10), 2, np.uint64))
    b = (np.array([9, 7], np.uint32), np.full((2, 10), 0.5, np.float32), np.full((2, 10), 3, np.uint64))
    b[1][0, 3] = 0.125
    out = str(tmp_path / "m.rand_lst")
    api.null_write(out, [a, b])
    rows = {int(ln.split()[0]): ln.split()[1:] for ln in open(out)}
    assert list(rows) == [5, 7, 9, 100]
    assert rows[5] == ["0.25", "2"] * 10 and rows[7] == ["0.5", "3"] * 10
    want9 = ["0.5", "5"] * 10
    want9[6] = "0.25"
    assert rows[9] == want9
    api.null_write(out, [])
    assert open(out).read() == ""


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(S.NULLGEN_RUNS))
def test_null_batch_equals_reference_rand_lst(nullgen, tag, tmp_path):
    run = S.NULLGEN_RUNS[tag]
    ctx = make_ctx(nullgen["table"], nullgen["paths"], run["prune"])
    reads = op.gen_rand_reads(S.NULLGEN_TIME, run["n_reads"], run["read_len"])
    ctx.null_batch(reads[:311], first_index=0)
    ctx.null_batch(reads[311:], first_index=311)
    t, m, c, nerr = ctx.null_table()
    assert nerr == 0
    out = str(tmp_path / "x.rand_lst")
    api.null_write(out, [(t, m, c)])
    assert open(out).read() == golden_text(tag)
    # reset really clears
    ctx.null_reset()
    assert len(ctx.null_table()[0]) == 0
    # a label call on an rkmer ctx is refused instead of silently using the wrong semantics
    with pytest.raises(api.KmatError):
        ctx.label(reads[:2])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(S.NULLGEN_RUNS))
def test_rand_read_label_cli_equals_reference(nullgen, tag, tmp_path):
    build.build_all()
    run, P = S.NULLGEN_RUNS[tag], nullgen["paths"]
    db = str(tmp_path / "null.kmat")
    api.Table.from_arrays(*nullgen["table"]).save(db)
    ofb = str(tmp_path / f"rrl_{tag}")
    cmd = [build.RRL_BIN, "-w", P["rank"], "-f", P["map16"], "-g", str(run["n_reads"]), "-i", str(run["read_len"]), "-e", P["depth"], "-p",
           "-t", "1", "-d", db, "-c", P["tree"], "-o", ofb]
    if run["prune"]:
        cmd += ["-h", str(run["prune"]), "-r", P["numrank"]]
    env = dict(os.environ, KMAT_RAND_COMPAT="glibc", KMAT_RAND_SEED=str(S.NULLGEN_TIME), KMAT_DEVICES="0")
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert p.returncode == 0, p.stderr + p.stdout
    assert open(ofb + ".rand_lst").read() == golden_text(tag)


def py_draw_reads(seed, first, n, rl):
    """The device generator of kmat_null.cuh restated with numpy uint64 arithmetic."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(z):
        z = z.astype(np.uint64)
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        idx = np.arange(first, first + n, dtype=np.uint64)
        key = mix(np.uint64(seed) ^ mix(idx + np.uint64(0x9E3779B97F4A7C15)))
        G = np.uint64(0x9E3779B97F4A7C15)
        d0 = mix(key)                                                   # kn_draw(key, 0)
        gc_draw = (idx % np.uint64(10)) * np.uint64(10) + (((d0 >> np.uint64(32)) * np.uint64(10)) >> np.uint64(32))
        rem_gc = (np.float32(1) * (gc_draw.astype(np.float64) / 100.0).astype(np.float32) * np.float32(rl)).astype(np.uint32).astype(np.int64)
        out = np.zeros((n, rl), dtype=np.uint8)
        for i in range(rl):
            d = mix(key + np.uint64(1 + i) * G)
            pick = (((d >> np.uint64(32)) * np.uint64(rl - i)) >> np.uint64(32)).astype(np.int64) < rem_gc
            coin = (d & np.uint64(1)).astype(bool)
            out[:, i] = np.where(pick, np.where(coin, ord("g"), ord("c")), np.where(coin, ord("a"), ord("t")))
            rem_gc = rem_gc - pick
    return [bytes(r) for r in out]


@pytest.mark.gpu
def test_device_generator_equals_restatement():
    for seed, first, n, rl in ((1, 0, 777, 150), (2**63 + 12345, 999_999_999_990, 300, 64), (7, 5, 40, 1000), (9, 0, 3, 70000)):
        got = api.null_draw_reads(seed, first, n, rl)
        assert got == py_draw_reads(seed, first, n, rl), (seed, first, n, rl)
    # batching does not change the reads: index-keyed
    a = api.null_draw_reads(11, 0, 500, 100)
    assert a[123:400] == api.null_draw_reads(11, 123, 277, 100)
    # GC content per bucket as genRandRead defines it, and both letters of each class present
    for i, r in enumerate(a):
        gc = sum(ch in b"gc" for ch in r)
        b = i % 10
        assert int(np.float32(np.float32(b * 10 / 100.0) * np.float32(100))) <= gc <= int(np.float32(np.float32((b * 10 + 9) / 100.0) * np.float32(100)))
    txt = b"".join(a)
    frac = [txt.count(ch) / len(txt) for ch in (b"a", b"c", b"g", b"t")]
    assert abs(frac[0] - frac[3]) < 0.01 and abs(frac[1] - frac[2]) < 0.01


@pytest.mark.gpu
def test_null_random_equals_oracle_on_device_drawn_reads(tmp_path):
    """A table built from mutated fragments of the reads the DEVICE draws, so that they hit it; the accumulated table of
    kmat_null_random must equal the oracle's over the same reads (fetched with the test hook)."""
    seed, n, rl = 424242, 4000, 120
    reads = api.null_draw_reads(seed, 0, n, rl)
    tax = fx.make_taxonomy(57, 40, specials=True)
    P = fx.write_taxonomy_files(tax, str(tmp_path))
    rng = fx.rng_for(58)
    code = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"acgt"):
        code[ch] = i
    leaves = list(tax.leaves)
    frags = {t: [fx.random_codes(rng, 200, 0.5)] for t in leaves}
    reads_t1 = api.null_draw_reads((seed + 0x632BE59BD9B4E019) & (2**64 - 1), 0, 500, rl)      # the CLI's second "thread"
    for r in reads[:1500] + reads_t1:
        codes = code[np.frombuffer(r, dtype=np.uint8)]
        li = int(rng.integers(0, len(leaves)))
        for o in {leaves[li], leaves[(li + int(rng.integers(0, 3))) % len(leaves)]}:
            a = int(rng.integers(0, rl - 40))
            seg = codes[a:a + int(rng.integers(40, rl - a + 1))].copy()
            flip = rng.random(len(seg)) < 0.01
            seg[flip] = (seg[flip] + rng.integers(1, 4, size=int(flip.sum()))) % 4
            frags[o].append(seg)
    genomes = {t: np.concatenate(v) for t, v in frags.items()}
    kmers, offs, tids = fx.build_kmer_table(genomes, tax, 20)
    m16 = fx.map16(tax)
    ids = np.array([m16[int(t)] for t in tids], dtype=np.uint32)
    table = (kmers, offs, ids, 20, 2)
    orc = make_oracle(table, P, None)
    orc.null_batch(reads, 0)
    want = op.format_rand_lst(*orc.null_table())
    assert want.count("\n") > 50
    ctx = make_ctx(table, P, None)
    ctx.null_random(seed, 0, 1700, rl)               # two calls: the accumulators persist, the key is the run index
    ctx.null_random(seed, 1700, n - 1700, rl)
    t, m, c, nerr = ctx.null_table()
    assert nerr == 0
    out = str(tmp_path / "dev.rand_lst")
    api.null_write(out, [(t, m, c)])
    assert open(out).read() == want
    # the binary's default mode: -t 2 -g N = two key spaces (seed + C * thread), buckets restart per thread
    build.build_all()
    db = str(tmp_path / "dev.kmat")
    api.Table.from_arrays(*table).save(db)
    ofb = str(tmp_path / "cli")
    env = dict(os.environ, KMAT_RAND_SEED=str(seed), KMAT_DEVICES="0")
    p = subprocess.run([build.RRL_BIN, "-w", P["rank"], "-f", P["map16"], "-g", "1503", "-i", str(rl), "-e", P["depth"], "-p", "-t", "2", "-d", db,
                        "-c", P["tree"], "-o", ofb], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert p.returncode == 0, p.stderr + p.stdout
    orc.null_reset()
    orc.null_batch(api.null_draw_reads(seed, 0, 1503, rl), 0)
    orc.null_batch(api.null_draw_reads((seed + 0x632BE59BD9B4E019) & (2**64 - 1), 0, 1503, rl), 0)
    assert open(ofb + ".rand_lst").read() == op.format_rand_lst(*orc.null_table())
