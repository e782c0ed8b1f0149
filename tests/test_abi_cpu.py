"""CPU: the C-ABI library loads, exports every symbol include/kmat.h declares, and its host-side
pieces (table ingest, input parsers, formatter) behave -- no compute calls, no GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from lmat_b200 import api
from oracle import oracle_py as op


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "kmat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kmat_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kmat.h but not exported by libkmat.so"
    assert set(names) == set(api.EXPORTS)
    assert L.kmat_abi_version() == 1


def test_compute_call_without_gpu_fails_loudly():
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    t = api.Table.from_arrays(np.array([5, 9], dtype=np.uint64), np.array([0, 1, 2], dtype=np.uint64), np.array([3, 4], dtype=np.uint32))
    with pytest.raises(api.KmatError) as e:
        api.Db.upload(t)
    assert e.value.code == -2          # KMAT_ERR_NO_DEVICE: there is no CPU fallback


def test_every_device_entry_point_fails_loudly_without_gpu(tmp_path):
    """The entry points added for gene_label, null models, content_summ, the peer gather probe: argument errors are
    reported as such, and a valid call without a device is KMAT_ERR_NO_DEVICE -- never a silent CPU result."""
    import ctypes as C
    L = api.lib()
    with pytest.raises(api.KmatError) as e:
        api.KmerCov([8, 21])                       # k > 20 does not fit the 40-bit k-mer field
    assert e.value.code == -7
    with pytest.raises(api.KmatError) as e:
        api.KmerCov(list(range(1, 10)))            # more than 8 k values
    assert e.value.code == -1
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.KmatError) as e:
        api.KmerCov([8, 20])
    assert e.value.code == -2
    with pytest.raises(api.KmatError) as e:
        api.null_draw_reads(1, 0, 4, 50)
    assert e.value.code == -2
    with pytest.raises(api.KmatError) as e:
        api.gather_bench_peer(0, 1, 1 << 20, 32, 1 << 10, 1)
    assert e.value.code == -2
    free_b, total_b = C.c_uint64(), C.c_uint64()
    assert L.kmat_device_memory(0, C.byref(free_b), C.byref(total_b)) == -2
    # the host binaries refuse to run too
    from lmat_b200 import build
    build.build_all()
    import subprocess
    for exe, args in ((build.RRL_BIN, ["-d", "x", "-o", str(tmp_path / "o"), "-t", "1", "-i", "50", "-g", "10", "-f", "m"]),
                      (build.GL_BIN, ["-d", "x", "-o", str(tmp_path / "o"), "-l", "lst"]),
                      (build.CS_BIN, ["-f", "lst", "-l", "fs", "-c", "tree", "-o", str(tmp_path / "o")])):
        p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode != 0, exe


def test_table_from_sorteddb_layout_roundtrip(golden_small):
    """Walk the reference's SortedDb memory layout (rebuilt by oracle_py.SortedDbArrays from the golden
    dump) through kmat_table_from_sorteddb and get the dump back."""
    g = golden_small
    sd = op.SortedDbArrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    t = api.Table.from_sorteddb(sd.top_tier, sd.kmer_table, sd.storage, g.kmer_len, g.tid_bytes)
    k, o, i = t.arrays()
    assert t.size == len(g.kmers) and t.kmer_length == 20
    assert np.array_equal(k, g.kmers) and np.array_equal(o, g.offs) and np.array_equal(i, g.ids)


def test_flat_image_save_open(golden_small, tmp_path):
    g = golden_small
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    p = str(tmp_path / "t.kmat")
    t.save(p)
    t2 = api.Table.open(p)
    k, o, i = t2.arrays()
    assert np.array_equal(k, g.kmers) and np.array_equal(o, g.offs) and np.array_equal(i, g.ids)
    with pytest.raises(api.KmatError):
        api.Table.open(str(tmp_path / "missing.db"))
    bad = tmp_path / "bad.db"
    bad.write_bytes(b"x" * 200)
    with pytest.raises(api.KmatError) as e:
        api.Table.open(str(bad))
    assert e.value.code == -6


def test_table_rejects_unsorted():
    with pytest.raises(api.KmatError):
        api.Table.from_arrays(np.array([9, 5], dtype=np.uint64), np.array([0, 1, 2], dtype=np.uint64), np.array([3, 4], dtype=np.uint32))


def test_inputs_parse_and_bad_tree(golden_small, tmp_path):
    P = golden_small.paths
    api.Inputs(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], numrank=P["numrank"], plasmids=P["plasmids"],
               null_lst=P["null_lst"], lmat_dir=golden_small.workdir)
    with pytest.raises(api.KmatError) as e:
        api.Inputs(tree=str(tmp_path / "nope.tree"))
    assert e.value.code == -5


def test_format_tail_matches_oracle_formatter(golden_small):
    """The product's line formatter and the oracle's agree on every result the oracle produces."""
    from conftest import oracle_for
    g = golden_small
    for opts, prn in (("run_rl", True), ("defaults", False)):
        orc = oracle_for(g, opts)
        hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
        res, cands, lin = orc.label(seqs)
        want = orc.tails(res)
        r2 = np.zeros(len(res), dtype=api.RESULT_DTYPE)
        for f in api.RESULT_DTYPE.names:
            r2[f] = res[f]
        buf = C.create_string_buffer(1 << 20)
        for i in range(len(res)):
            n = api.lib().kmat_format_tail(r2[i:i + 1].ctypes.data, cands.ctypes.data if len(cands) else None,
                                           lin.ctypes.data if len(lin) else None, int(prn), buf, len(buf))
            assert n >= 0 and buf.raw[:n].decode() == want[i]


def test_format_tail_numbers_match_printf_g():
    """The fast float formatter behind kmat_format_tail equals printf("%g") of the promoted double (what
    ostream << float prints, read_label.cpp:894-937) on random bit patterns, score-like values and rounding boundaries."""
    rng = np.random.default_rng(12)
    vals = [rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32).view(np.float32),           # any bit pattern
            (rng.random(200000) * 30 - 15).astype(np.float32), rng.random(200000).astype(np.float32),
            np.log(rng.integers(1, 131, 100000) / 131.0 / 0.0301).astype(np.float32),
            (rng.integers(0, 10 ** 6, 100000) / 10 ** rng.integers(0, 9, 100000)).astype(np.float32),     # short decimals
            ((rng.integers(0, 10 ** 6, 100000) + 0.5) / 10 ** rng.integers(0, 9, 100000)).astype(np.float32),   # near ties
            np.array([0.0, -0.0, 1e-4, 9.99995e-5, 999999.0, 999999.5, 1e6, 0.1, 1.0, 100000.0, 0.000123456, 1e-38, 3e38, np.inf, -np.inf], dtype=np.float32)]
    v = np.concatenate(vals)
    v = v[~np.isnan(v)]
    res = np.zeros(1, dtype=api.RESULT_DTYPE)
    res["status"], res["match"], res["n_cand"], res["cand_off"] = 5, 0, 16, 0
    cands = np.zeros(16, dtype=api.PAIR_DTYPE)
    buf = C.create_string_buffer(4096)
    for a in range(0, len(v) - 18, 18):
        blk = v[a:a + 18]
        res["log_avg"], res["stdev"], res["cand_kmer_cnt"], res["tid"], res["score"] = blk[0], blk[1], 131, 9606, abs(blk[2])
        cands["tid"] = rng.integers(0, 2 ** 32, 16, dtype=np.uint64) >> rng.integers(0, 32, 16, dtype=np.uint64)     # every digit count
        cands["score"] = np.abs(blk[2:18])
        cands["score"][5] = cands["score"][4]                       # repeated score: the memoised text
        n = api.lib().kmat_format_tail(res.ctypes.data, cands.ctypes.data, None, 1, buf, len(buf))
        want = "%g %g 131\t" % (float(blk[0]), float(blk[1])) + "".join(" %d %g" % (int(t), float(s)) for t, s in zip(cands["tid"][::-1], cands["score"][::-1]))
        want += "\t9606 %g DirectMatch\n" % float(abs(blk[2]))
        assert n >= 0 and buf.raw[:n].decode() == want


def test_format_tail_negative_scores_and_verbose_mode():
    """-p prints a candidate only when its score is >= 0; together with -y (prn_all = 2) all of them (read_label.cpp:901); a list
    without a printable candidate prints "-1 -1"."""
    res = np.zeros(1, dtype=api.RESULT_DTYPE)
    res["status"], res["match"], res["n_cand"], res["cand_off"] = 5, 0, 4, 0
    res["log_avg"], res["stdev"], res["cand_kmer_cnt"], res["tid"], res["score"] = 1.5, 0.25, 100, 562, 2.0
    cands = np.zeros(4, dtype=api.PAIR_DTYPE)
    cands["tid"] = [10, 20, 30, 40]
    cands["score"] = [-0.5, 0.0, -3.25, 2.0]
    buf = C.create_string_buffer(4096)
    def fmt(mode):
        n = api.lib().kmat_format_tail(res.ctypes.data, cands.ctypes.data, None, mode, buf, len(buf))
        assert n >= 0
        return buf.raw[:n].decode()
    assert fmt(1) == "1.5 0.25 100\t 40 2 20 0\t562 2 DirectMatch\n"
    assert fmt(2) == "1.5 0.25 100\t 40 2 30 -3.25 20 0 10 -0.5\t562 2 DirectMatch\n"
    assert fmt(0) == "1.5 0.25 100\t562 2 DirectMatch\n"
    cands["score"] = [-0.5, -1.0, -3.25, -2.0]
    assert fmt(1) == "1.5 0.25 100\t-1 -1\t562 2 DirectMatch\n"


def test_flat_image_damaged_header_is_rejected(golden_small, tmp_path):
    """A .kmat image whose header counts are damaged (including values that would wrap the size computation around)
    or whose offsets do not end at n_ids is KMAT_ERR_FORMAT, not a crash."""
    import struct
    g = golden_small
    p = str(tmp_path / "t.kmat")
    api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes).save(p)
    good = open(p, "rb").read()
    for field_off, val in ((24, 1 << 62), (24, (1 << 64) - 1), (32, (1 << 64) - 1), (32, 1 << 62), (32, len(g.ids) - 1), (24, len(g.kmers) - 1),
                           (12, 0), (12, 200), (16, 3)):
        d = bytearray(good)
        d[field_off:field_off + (8 if field_off >= 24 else 4)] = struct.pack("<Q" if field_off >= 24 else "<I", val)
        open(p, "wb").write(bytes(d))
        with pytest.raises(api.KmatError) as e:
            api.Table.open(p)
        assert e.value.code == -6, (field_off, val)
    open(p, "wb").write(good[:len(good) // 2])
    with pytest.raises(api.KmatError):
        api.Table.open(p)


def test_fast_float_formatter_equals_printf_over_its_whole_range(tmp_path):
    """km_fmt_g is integer-exact by construction; this compares it with printf("%g") on every 7th float of its fast range
    (40 M values; KMAT_EXHAUSTIVE=1: all 279 M, which is how the change was accepted)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fmt_exhaustive")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "fmt_exhaustive.cpp"),
                    "-o", exe, "-lz", "-lpthread"], check=True)
    p = subprocess.run([exe, "1" if os.environ.get("KMAT_EXHAUSTIVE") else "7"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "mismatches 0" in p.stdout


def test_format_tail_never_writes_past_its_capacity():
    """Random results (any status, any float bit pattern, up to 300 candidates) into buffers of awkward sizes: the call
    either fits (n <= cap) or reports KMAT_ERR_OVERFLOW / an error, and the bytes behind `cap` stay untouched (the
    formatter places digits with fixed-size copies that need slack, which the capacity check has to account for)."""
    L = api.lib()
    rng = np.random.default_rng(5)
    for it in range(4000):
        nc = int(rng.choice([0, 1, 2, 5, 17, 64, 300]))
        nl = int(rng.choice([0, 1, 3, 40]))
        res = np.zeros(1, dtype=api.RESULT_DTYPE)
        res["status"], res["match"] = rng.integers(0, 8), rng.integers(0, 6)
        for f in ("n1", "n2", "valid_kmers", "cand_kmer_cnt", "bin_sel"):
            res[f] = rng.integers(-2 ** 31, 2 ** 31)
        res["tid"], res["n_cand"], res["n_lin"] = rng.integers(0, 2 ** 32), nc, nl
        res["score"], res["log_avg"], res["stdev"] = rng.integers(0, 2 ** 32, 3, dtype=np.uint64).astype(np.uint32).view(np.float32)
        cands, lin = np.zeros(max(nc, 1), dtype=api.PAIR_DTYPE), np.zeros(max(nl, 1), dtype=api.PAIR_DTYPE)
        cands["tid"], lin["tid"] = rng.integers(0, 2 ** 32, len(cands)), rng.integers(0, 2 ** 32, len(lin))
        mode = it % 3
        for arr in (cands, lin):
            n = len(arr)
            arr["score"] = (rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32) if mode == 0
                            else (rng.random(n) * 10).astype(np.float32) if mode == 1 else np.repeat(np.float32(rng.random()), n))
        full = 160 + 40 * max(nc, nl)
        cap = int(rng.choice([0, 1, 50, 159, 160, 161, 200, full, full - 1, 100000]))
        buf = np.full(cap + 64, 0xA5, dtype=np.uint8)
        n = L.kmat_format_tail(res.ctypes.data, cands.ctypes.data, lin.ctypes.data, int(rng.integers(0, 2)), C.cast(buf.ctypes.data, C.c_char_p), cap)
        assert n <= cap
        assert np.all(buf[cap:] == 0xA5), (it, cap, n)
