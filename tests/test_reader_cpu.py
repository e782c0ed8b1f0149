"""CPU: the C host parser (kmat_reader_*, replaces read_label.cpp:1651-1732) against the oracle's reader
restatement, which tests/test_oracle_golden.py pins to the reference's own outputs for FASTA, wrapped FASTA
and FASTQ (the previous-record-header quirk included)."""
import os

import pytest

from lmat_b200 import api
from oracle import oracle_py as op


@pytest.mark.parametrize("key,fastq", [("reads", False), ("reads_wrapped", False), ("reads_fq", True)])
@pytest.mark.parametrize("max_reads", [1, 7, 1 << 20])
def test_reader_matches_oracle_reader(golden_small, key, fastq, max_reads):
    path = golden_small.paths[key]
    want = op.read_fasta_like_reference(path, fastq=fastq)
    got = api.read_file(path, fastq=fastq, max_reads=max_reads)
    assert got[0] == want[0] and got[1] == want[1]


CASES = {
    "empty": b"",
    "no_trailing_newline": b">a\nACGTACGT\n>b\nGGGG",
    "blank_and_single_char_lines": b">a\nACGT\n\nA\nTTTT\n>b\n\n>c\nCC\n",
    "no_header": b"ACGTACGTAC\nGGGT\n>x\nAAAA\n",
    "crlf": b">a\r\nACGT\r\n>b\r\nGG\r\n",
    "empty_header": b">\nACGT\n>\nGGCC\n",
    "fasta_read_as_fastq": b">h1\nACGT\n>h2\nGGGG\n",
    "fastq_multi": b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\nTT\n-r2\nIIIIII\n@r3\nAA\n+\nII",
    "fastq_no_quality_at_end": b"@r1\nACGT\n+\n",
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("fastq", [False, True])
def test_reader_edge_cases(tmp_path, name, fastq):
    p = os.path.join(tmp_path, name + ".txt")
    open(p, "wb").write(CASES[name])
    want = op.read_fasta_like_reference(p, fastq=fastq)
    for mr in (1, 3, 1000):
        got = api.read_file(p, fastq=fastq, max_reads=mr)
        assert got[0] == want[0] and got[1] == want[1], (name, fastq, mr)


def test_reader_long_line_and_chunk_boundaries(tmp_path):
    # one 40 MB sequence line (larger than the reader's chunk) between normal records
    p = os.path.join(tmp_path, "big.fa")
    with open(p, "wb") as f:
        f.write(b">a\nACGT\n>big\n" + b"ACGT" * (10 << 20) + b"\n>c\nGG\nTT\n")
    hdrs, seqs = api.read_file(p, max_reads=2)
    assert hdrs == ["a", "big", "c"] and [len(s) for s in seqs] == [4, 40 << 20, 4] and seqs[2] == "GGTT"


def test_reader_missing_file():
    with pytest.raises(api.KmatError):
        api.read_file("/nonexistent/reads.fa")


@pytest.mark.parametrize("seg_bytes", [1, 7, 64, 4096])
def test_parallel_reader_equals_sequential(tmp_path, golden_small, seg_bytes, monkeypatch):
    """kmat_reader_open_mt: FASTA cut at header lines into segments parsed by several threads gives the same
    (header, read) sequence as the sequential state machine -- wrapped lines, blank lines, empty headers
    ("unknown_hdr:<global ordinal>"), consecutive headers, no trailing newline, data before the first header."""
    monkeypatch.setenv("KMAT_READER_SEG_BYTES", str(seg_bytes))
    files = {k: open(golden_small.paths[k], "rb").read() for k in ("reads", "reads_wrapped")}
    files["mixed"] = (b"ACGTACGTAA\nCCGG\n>h1\nACGT\nTTGA\n\nA\n>\nGGGTTT\n>h3\n>h4\nAC\n>\n>\nTTTTT\nGG\n>h7 with spaces\tand tab\r\nACGT\r\n" * 40
                      + b">last\nACGTACGTT")
    files["headers_only"] = b">a\n>b\n>c\n"
    files["one_record"] = b">only\nACGT\n"
    for name, data in files.items():
        p = os.path.join(tmp_path, f"{name}_{seg_bytes}.fa")
        open(p, "wb").write(data)
        want = api.read_file(p, threads=1)
        for th in (2, 5):
            got = api.read_file(p, threads=th)
            assert got == want, (name, th)
        assert want == tuple(op.read_fasta_like_reference(p)) or list(want) == list(op.read_fasta_like_reference(p))


@pytest.mark.parametrize("seg_bytes", [1, 9, 80, 4096])
def test_parallel_fastq_reader_equals_sequential(tmp_path, golden_small, seg_bytes, monkeypatch):
    """kmat_reader_open_mt on FASTQ: segments start at the '@' line of a record (an '@' line followed by sequence lines, a
    '+' / '-' line, the quality line and then another '@' line or the end), each read keeps the reference's pairing with
    the PREVIOUS record's header across segment boundaries, quality lines starting with '@' or '+' never start a segment."""
    monkeypatch.setenv("KMAT_READER_SEG_BYTES", str(seg_bytes))
    import numpy as np
    rng = np.random.default_rng(4)
    recs = []
    for i in range(300):
        L = int(rng.integers(1, 90))
        seq = "".join("ACGTN"[x] for x in rng.integers(0, 5, L))
        q0 = "@+-I#"[int(rng.integers(0, 5))]                       # quality lines that look like headers / separators
        qual = q0 + "".join(chr(int(x)) for x in rng.integers(33, 74, L - 1))
        hdr = "" if i % 37 == 5 else f"read{i} len={L}"
        if i % 11 == 3 and L > 10:                                  # sequence wrapped over two lines
            recs.append(f"@{hdr}\n{seq[:L // 2]}\n{seq[L // 2:]}\n+{hdr}\n{qual}\n")
        else:
            recs.append(f"@{hdr}\n{seq}\n{'+-'[i % 2]}\n{qual}\n")
    files = {"synthetic": "".join(recs).encode(), "golden": open(golden_small.paths["reads_fq"], "rb").read(),
             "no_final_newline": "".join(recs)[:-1].encode(), "one": b"@r\nACGT\n+\nIIII\n"}
    for name, data in files.items():
        p = os.path.join(tmp_path, f"{name}_{seg_bytes}.fq")
        open(p, "wb").write(data)
        want = api.read_file(p, fastq=True, threads=1)
        assert list(want) == list(op.read_fasta_like_reference(p, fastq=True)), name
        for th in (2, 6):
            got = api.read_file(p, fastq=True, threads=th)
            assert got == want, (name, th)
    assert sum(h.startswith("unknown_hdr:") for h in api.read_file(os.path.join(tmp_path, f"synthetic_{seg_bytes}.fq"), fastq=True, threads=3)[0]) >= 8


def _fuzz_inputs(seed, n_cases):
    """Byte soup over the characters the state machine looks at, and FASTQ-shaped records with occasional damage
    (missing '+' lines, missing quality lines, blank lines, quality lines starting with '@' or '+')."""
    import random
    rng = random.Random(seed)
    alpha = [b">", b"@", b"+", b"-", b"\n", b"\n", b"\n", b"A", b"C", b"G", b"T", b"N", b"\r", b" ", b"\t", b"I", b"x"]
    for i in range(n_cases):
        if i % 2 == 0:
            n = rng.choice([0, 1, 2, 5, 20, 80, 300])
            yield b"".join(rng.choice(alpha) * rng.choice([1, 1, 1, 2, 7]) for _ in range(n))
        else:
            out = []
            for r in range(rng.choice([1, 3, 10, 40])):
                seq = "".join(rng.choice("ACGTN") for _ in range(rng.choice([1, 4, 30])))
                qual = rng.choice(["I", "@", "+", "-", ">"]) + "I" * (len(seq) - 1)
                lines = ["@r%d" % r if rng.random() < 0.9 else "@", seq, rng.choice(["+", "+", "-", "+r"]), qual]
                dmg = rng.random()
                if dmg < 0.05: del lines[2]
                elif dmg < 0.10: del lines[3]
                elif dmg < 0.15: lines.insert(rng.randrange(4), "")
                elif dmg < 0.20: lines.insert(2, seq[:3])
                elif dmg < 0.25: del lines[0]
                out += lines
            yield ("\n".join(out) + ("\n" if rng.random() < 0.7 else "")).encode()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reader_fuzz_sequential_and_parallel_equal_oracle(tmp_path, seed, monkeypatch):
    """Arbitrary input, FASTA and FASTQ mode: the sequential reader and the parallel one (any segment size) give the oracle
    reader's (header, read) sequence.  Malformed FASTQ is where the parallel reader's cut rule is not enough and its
    after-the-fact check (open_end -> sequential from that segment on) has to take over."""
    p = os.path.join(tmp_path, "fz.txt")
    cases = [b"N\n@\n\n+\nG"] + list(_fuzz_inputs(seed, 60))
    for data in cases:
        open(p, "wb").write(data)
        for fastq in (False, True):
            want = op.read_fasta_like_reference(p, fastq=fastq)
            monkeypatch.delenv("KMAT_READER_SEG_BYTES", raising=False)
            for mr in (1, 1000):
                got = api.read_file(p, fastq=fastq, max_reads=mr)
                assert got[0] == want[0] and got[1] == want[1], (data, fastq, mr)
            for seg in ("1", "13", "200"):
                monkeypatch.setenv("KMAT_READER_SEG_BYTES", seg)
                got = api.read_file(p, fastq=fastq, threads=3)
                assert got[0] == want[0] and got[1] == want[1], (data, fastq, seg)


def test_parallel_reader_respects_batch_limits(tmp_path):
    """Parallel mode parses 8 MB segments; kmat_reader_next still hands out at most max_reads reads / max_bases bases per
    batch (a batch always holds at least one read), in file order, with the same records as the sequential reader."""
    import ctypes as C
    import numpy as np
    from lmat_b200 import fixtures as fx
    rng = np.random.default_rng(11)
    hdrs = [f"r{i} x" for i in range(3000)]
    seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(30, 300)))) for _ in hdrs]
    p = str(tmp_path / "r.fa")
    fx.write_fasta(p, hdrs, seqs)
    L = api.lib()
    for max_reads, max_bases in ((17, 1 << 30), (1 << 20, 2000), (1, 1)):
        r = C.c_void_p()
        assert L.kmat_reader_open_mt(p.encode(), 0, 4, C.byref(r)) == 0
        b = C.c_void_p(L.kmat_read_batch_new())
        got, sizes = [], []
        while True:
            n = L.kmat_reader_next(r, max_reads, max_bases, b)
            assert n >= 0
            if n == 0:
                break
            bp, op_, hp, hop = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            nn, first = C.c_uint32(), C.c_uint64()
            assert L.kmat_read_batch_view(b, C.byref(bp), C.byref(op_), C.byref(hp), C.byref(hop), C.byref(nn), C.byref(first)) == 0
            offs = np.ctypeslib.as_array(C.cast(op_, C.POINTER(C.c_uint64)), shape=(n + 1,))
            assert first.value == len(got) + 1 and nn.value == n
            assert n <= max_reads and (n == 1 or int(offs[n]) <= max_bases)
            bases = C.string_at(bp, int(offs[n]))
            got += [bases[int(offs[i]):int(offs[i + 1])].decode() for i in range(n)]
            sizes.append(n)
        L.kmat_read_batch_free(b)
        L.kmat_reader_close(r)
        assert got == seqs and len(sizes) > 1


@pytest.mark.parametrize("threads,max_reads", [(1, 1 << 20), (1, 7), (4, 1 << 20), (4, 50)])
def test_pinned_batches_hand_out_the_same_reads(golden_small, threads, max_reads, monkeypatch):
    """kmat_read_batch_new_pinned (what the read_label binary uses): same reads whether the page-locked buffer exists (GPU box)
    or pinning fails and the batch silently stays pageable (here)."""
    monkeypatch.setenv("KMAT_READER_SEG_BYTES", "4096")
    for key, fastq in (("reads", False), ("reads_wrapped", False), ("reads_fq", True)):
        path = golden_small.paths[key]
        want = op.read_fasta_like_reference(path, fastq=fastq)
        got = api.read_file(path, fastq=fastq, max_reads=max_reads, threads=threads, pinned=True)
        assert got[0] == want[0] and got[1] == want[1], (key, threads, max_reads)


@pytest.mark.gpu
def test_pinned_batches_on_the_gpu_box(golden_small, monkeypatch):
    monkeypatch.setenv("KMAT_READER_SEG_BYTES", "4096")
    for threads, max_reads in ((1, 7), (4, 50), (4, 1 << 20)):
        for key, fastq in (("reads", False), ("reads_fq", True)):
            path = golden_small.paths[key]
            want = op.read_fasta_like_reference(path, fastq=fastq)
            got = api.read_file(path, fastq=fastq, max_reads=max_reads, threads=threads, pinned=True)
            assert got[0] == want[0] and got[1] == want[1], (key, threads, max_reads)
