"""TEST INFRASTRUCTURE: ctypes front-end of oracle/libkmat_oracle.so plus independent (pure Python)
parsers of the reference's run-time text inputs and of the KMPERM01 stand-in DB image.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmat_oracle.so")

ST_NAMES = ["SHORT_LEN", "SHORT_VALID", "NODBHITS", "SILENT", "PHIX", "LABELED"]
MATCH_NAMES = ["DirectMatch", "MultiMatch", "PartialMultiMatch", "NoMatch", "LCA_ERROR"]


class Pair(C.Structure):
    _fields_ = [("tid", C.c_uint32), ("score", C.c_float)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32), ("valid_kmers", C.c_int32),
                ("cand_kmer_cnt", C.c_int32), ("match", C.c_int32), ("tid", C.c_uint32), ("score", C.c_float),
                ("log_avg", C.c_float), ("stdev", C.c_float), ("n_cand", C.c_uint32), ("n_lin", C.c_uint32),
                ("cand_off", C.c_uint64), ("lin_off", C.c_uint64), ("bin_sel", C.c_int32), ("err", C.c_int32)]


RESULT_DTYPE = np.dtype([("status", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("valid_kmers", "<i4"),
                         ("cand_kmer_cnt", "<i4"), ("match", "<i4"), ("tid", "<u4"), ("score", "<f4"),
                         ("log_avg", "<f4"), ("stdev", "<f4"), ("n_cand", "<u4"), ("n_lin", "<u4"),
                         ("cand_off", "<u8"), ("lin_off", "<u8"), ("bin_sel", "<i4"), ("err", "<i4")])
PAIR_DTYPE = np.dtype([("tid", "<u4"), ("score", "<f4")])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class Opts(C.Structure):
    _fields_ = [("min_kmer", C.c_int), ("min_fnd_kmer", C.c_int), ("sdiff", C.c_float), ("hbias", C.c_float),
                ("min_score", C.c_float), ("max_count", C.c_int), ("permissive", C.c_int), ("phix_screen", C.c_int),
                ("prn_all", C.c_int), ("prn_read", C.c_int), ("rkmer", C.c_int)]


class GlibcRand(C.Structure):
    _fields_ = [("st", C.c_int32 * 31), ("f", C.c_int), ("r", C.c_int)]


class Db(C.Structure):
    _fields_ = [("top_tier", C.c_void_p), ("kmer_table", C.c_void_p), ("storage", C.c_void_p),
                ("kmer_len", C.c_int), ("tid_bytes", C.c_int)]


def build_lib():
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "kmat_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_lib())
        L.kmo_ctx_new.restype = C.c_void_p
        L.kmo_ctx_free.argtypes = [C.c_void_p]
        L.kmo_default_opts.argtypes = [C.POINTER(Opts)]
        L.kmo_set_opts.argtypes = [C.c_void_p, C.POINTER(Opts)]
        L.kmo_set_db.argtypes = [C.c_void_p, C.POINTER(Db)]
        for name in ("kmo_set_tree", "kmo_set_depth", "kmo_set_conv", "kmo_set_prune_ranks"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.kmo_set_ranks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.kmo_set_plasmids.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.kmo_load_null_models.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.kmo_label_batch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.kmo_cands.restype = C.POINTER(Pair)
        L.kmo_cands.argtypes = [C.c_void_p]
        L.kmo_lineage.restype = C.POINTER(Pair)
        L.kmo_lineage.argtypes = [C.c_void_p]
        L.kmo_lookup_batch.restype = C.c_int64
        L.kmo_lookup_batch.argtypes = [C.POINTER(Db), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.kmo_encode_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.kmo_format_tail.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        L.kmo_gene_label_read.argtypes = [C.POINTER(Db), C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.kmo_srand.argtypes = [C.POINTER(GlibcRand), C.c_uint]
        L.kmo_rand.argtypes = [C.POINTER(GlibcRand)]
        L.kmo_gen_rand_reads.argtypes = [C.POINTER(GlibcRand), C.c_uint64, C.c_uint64, C.c_int, C.c_char_p]
        L.kmo_null_batch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint64]
        L.kmo_null_rows.restype = C.c_uint32
        L.kmo_null_rows.argtypes = [C.c_void_p]
        L.kmo_null_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kmo_null_reset.argtypes = [C.c_void_p]
        L.kmo_logf.restype = C.c_float
        L.kmo_logf.argtypes = [C.c_float]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------------
# independent parsers of the reference's text inputs
# ------------------------------------------------------------------------------------------------
def parse_tree(path):
    """TaxTree ctor (TaxTree.hpp:24-57) + TaxNode::read (TaxNode.hpp:131-147): token stream
    ``id nchild child*nchild parent`` then the rest of that line and one name line."""
    with open(path, "rb") as f:
        data = f.read().decode("latin-1")
    lines = data.split("\n")
    body = lines[3:]               # two header lines + the count line
    tids, parents = [], []
    i = 0
    while i < len(body):
        toks = body[i].split()
        if not toks:
            i += 1
            continue
        tid, n = int(toks[0]), int(toks[1])
        parents.append(int(toks[2 + n]))
        tids.append(tid)
        i += 2                     # the name line
    return np.array(tids, dtype=np.uint32), np.array(parents, dtype=np.uint32)


def parse_pairs(path, dtype=np.uint32):
    a, b = [], []
    with open(path) as f:
        for ln in f:
            t = ln.split()
            if len(t) >= 2:
                a.append(int(t[0]))
                b.append(int(t[1]))
    return np.array(a, dtype=dtype), np.array(b, dtype=dtype)


def parse_ranks(path):
    tids, codes = [], []
    with open(path) as f:
        for ln in f:
            t = ln.split()
            if len(t) >= 2:
                tids.append(int(t[0]))
                codes.append(1 if t[1] == "strain" else 2 if t[1] == "species" else 0)
    return np.array(tids, dtype=np.uint32), np.array(codes, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------
# KMPERM01 stand-in image (oracle/standins/jemalloc/pallocator.h) -> SortedDb arrays
# ------------------------------------------------------------------------------------------------
class RefDbImage:
    """mmap of a DB written by oracle/_ref/make_db_table; exposes the SortedDb members
    (SortedDb.hpp:453-481, x86-64 layout) as numpy views."""

    def __init__(self, path, tid_bytes=2):
        self.f = open(path, "rb")
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        if self.mm[:8] != b"KMPERM01":
            raise ValueError("not a KMPERM01 image")
        base, size, brk, nreg = struct.unpack_from("<QQQQ", self.mm, 8)
        (n0,) = struct.unpack_from("<Q", self.mm, 40)
        (obj,) = struct.unpack_from("<Q", self.mm, 48)
        o = obj - base
        idx_config, = struct.unpack_from("<i", self.mm, o)
        m_n_kmers, = struct.unpack_from("<Q", self.mm, o + 8)
        klen, = struct.unpack_from("<B", self.mm, o + 16)
        p_storage, p_table, p_tt, n_rec = struct.unpack_from("<QQQQ", self.mm, o + 24)
        self.idx_config, self.kmer_len, self.n_kmers, self.tid_bytes = idx_config, klen, n_rec, tid_bytes
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        tt_count = 1 << 27
        self.top_tier = self.buf[p_tt - base:p_tt - base + tt_count * 8].view("<u8")
        self.kmer_table = self.buf[p_table - base:p_table - base + n_rec * 8]
        self.storage = self.buf[p_storage - base:brk]
        self.bits_2nd = 13 if klen == 20 else 9

    def cdb(self):
        return Db(self.top_tier.ctypes.data, self.kmer_table.ctypes.data, self.storage.ctypes.data, self.kmer_len,
                  self.tid_bytes)

    def dump(self):
        """(kmers ascending, offs, stored ids) -- the logical table, walked like SortedDb::begin_/next."""
        tt = np.asarray(self.top_tier)
        nz = np.nonzero(tt)[0]
        counts = (tt[nz] >> np.uint64(48)).astype(np.int64)
        starts = (tt[nz] & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
        rec = self.kmer_table.view(np.dtype([("lsb", "<u2"), ("page", "<u2"), ("off", "<u4")]))
        kmers = np.zeros(self.n_kmers, dtype=np.uint64)
        for p, s, c in zip(nz, starts, counts):
            kmers[s:s + c] = (np.uint64(p) << np.uint64(self.bits_2nd)) | rec["lsb"][s:s + c].astype(np.uint64)
        offs = np.zeros(self.n_kmers + 1, dtype=np.uint64)
        ids = []
        st = self.storage
        PAGE = 4294701056
        w = self.tid_bytes
        for i in range(self.n_kmers):
            page = int(rec["page"][i]) & 0xFF
            off = int(rec["off"][i])
            if page == 255:
                ids.append(np.array([off & (0xFFFF if w == 2 else 0xFFFFFFFF)], dtype=np.uint32))
            else:
                a = page * PAGE + off
                if int(kmers[i]) % 4096 == 0:
                    a += 8
                n = int(st[a]) | (int(st[a + 1]) << 8)
                ids.append(st[a + 2:a + 2 + n * w].view("<u2" if w == 2 else "<u4").astype(np.uint32))
            offs[i + 1] = offs[i] + len(ids[-1])
        return kmers, offs, (np.concatenate(ids) if ids else np.zeros(0, dtype=np.uint32))

    def close(self):
        self.top_tier = self.kmer_table = self.storage = self.buf = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.f.close()


class SortedDbArrays:
    """The reference's SortedDb memory layout rebuilt in numpy from a logical table dump, following
    SortedDb<tid_T>::add_data (SortedDb.cpp:84-751, no pruning / human / adaptor feeds): top tier entry
    = count<<48 | first record (:291-292 reader, :268-276 writer), 8-byte records sorted by k-mer,
    singleton taxid inline with page 255 (:492-517), otherwise the list is appended to page 0 as
    [u64 kmer if kmer % 4096 == 0][u16 count][count x tid_T] (:412-470).  Lets the oracle run where
    /root/reference (and so oracle/_ref/make_db_table) is absent, e.g. on the GPU box."""

    def __init__(self, kmers, offs, ids, kmer_len=20, tid_bytes=2):
        kmers = np.asarray(kmers, dtype=np.uint64)
        offs = np.asarray(offs, dtype=np.int64)
        ids = np.asarray(ids, dtype=np.uint32)
        assert np.all(kmers[1:] > kmers[:-1]), "k-mers must be strictly ascending (SortedDb.cpp:164-167)"
        self.kmer_len, self.tid_bytes, self.n_kmers = kmer_len, tid_bytes, len(kmers)
        bits = 13 if kmer_len == 20 else 9
        self.bits_2nd = bits
        n = len(kmers)
        self.top_tier = np.zeros(1 << 27, dtype=np.uint64)
        pfx = (kmers >> np.uint64(bits)).astype(np.int64)
        upfx, first, cnt = np.unique(pfx, return_index=True, return_counts=True)
        self.top_tier[upfx] = (cnt.astype(np.uint64) << np.uint64(48)) | first.astype(np.uint64)
        rec = np.zeros(n, dtype=np.dtype([("lsb", "<u2"), ("page", "<u2"), ("off", "<u4")]))
        rec["lsb"] = (kmers & np.uint64((1 << bits) - 1)).astype(np.uint16)
        lens = np.diff(offs)
        single = lens == 1
        rec["page"][single] = 255
        rec["off"][single] = ids[offs[:-1][single]]
        multi = np.nonzero(~single)[0]
        echo = (kmers[multi] % np.uint64(4096) == 0).astype(np.int64) * 8
        sizes = echo + 2 + lens[multi] * tid_bytes
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if len(multi) else np.zeros(0, np.int64)
        total = int(sizes.sum()) if len(multi) else 0
        assert total < 4294701056, "single page only in this rebuild"
        st = np.zeros(total + 16, dtype=np.uint8)
        rec["page"][multi] = 0
        rec["off"][multi] = starts
        tdt = "<u2" if tid_bytes == 2 else "<u4"
        for i, s, e in zip(multi, starts, echo):
            a = int(s)
            if e:
                st[a:a + 8] = np.frombuffer(np.uint64(kmers[i]).tobytes(), dtype=np.uint8)
                a += 8
            c = int(lens[i])
            st[a:a + 2] = np.frombuffer(np.uint16(c).tobytes(), dtype=np.uint8)
            st[a + 2:a + 2 + c * tid_bytes] = ids[offs[i]:offs[i + 1]].astype(tdt).view(np.uint8)
        self.kmer_table = rec.view(np.uint8)
        self.storage = st

    def cdb(self):
        return Db(self.top_tier.ctypes.data, self.kmer_table.ctypes.data, self.storage.ctypes.data, self.kmer_len,
                  self.tid_bytes)


# ------------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, db: RefDbImage | None = None, cdb: Db | None = None, keep=None):
        self.L = lib()
        self.ctx = C.c_void_p(self.L.kmo_ctx_new())
        self._keep = [db, keep]
        self.db = db
        self.cdb = cdb if cdb is not None else (db.cdb() if db is not None else None)
        if self.cdb is not None:
            self.L.kmo_set_db(self.ctx, C.byref(self.cdb))
        self.opts = Opts()
        self.L.kmo_default_opts(C.byref(self.opts))

    def __del__(self):
        try:
            self.L.kmo_ctx_free(self.ctx)
        except Exception:
            pass

    def set_opts(self, **kw):
        for k, v in kw.items():
            setattr(self.opts, k, v)
        self.L.kmo_set_opts(self.ctx, C.byref(self.opts))

    def _pair(self, fn, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        fn(self.ctx, len(a), a.ctypes.data, b.ctypes.data)

    def load_files(self, tree=None, depth=None, rank=None, map16=None, numrank=None, plasmids=None, null_lst=None,
                   lmat_dir=None):
        if tree:
            self._pair(self.L.kmo_set_tree, *parse_tree(tree))
        if depth:
            self._pair(self.L.kmo_set_depth, *parse_pairs(depth))
        if rank:
            t, c = parse_ranks(rank)
            self.L.kmo_set_ranks(self.ctx, len(t), t.ctypes.data, c.ctypes.data)
        if map16:
            t32, t16 = parse_pairs(map16)
            self._pair(self.L.kmo_set_conv, t16, t32)
        if numrank:
            self._pair(self.L.kmo_set_prune_ranks, *parse_pairs(numrank))
        if plasmids:
            with open(plasmids) as f:
                t = np.array([int(x) for x in f.read().split()], dtype=np.uint32)
            self.L.kmo_set_plasmids(self.ctx, len(t), t.ctypes.data)
        if null_lst:
            rc = self.L.kmo_load_null_models(self.ctx, null_lst.encode(), lmat_dir.encode() if lmat_dir else None)
            if rc < 0:
                raise RuntimeError(f"kmo_load_null_models rc={rc}")

    def label(self, seqs):
        """seqs: list of str/bytes.  Returns (results structured array, cands, lineage)."""
        bs = [s.encode() if isinstance(s, str) else s for s in seqs]
        offs = np.zeros(len(bs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in bs])
        blob = b"".join(bs)
        res = np.zeros(len(bs), dtype=RESULT_DTYPE)
        self.L.kmo_label_batch(self.ctx, blob, offs.ctypes.data, len(bs), res.ctypes.data)
        nc = int((res["cand_off"][-1] + res["n_cand"][-1])) if len(bs) else 0
        nl = int((res["lin_off"][-1] + res["n_lin"][-1])) if len(bs) else 0
        cands = np.ctypeslib.as_array(self.L.kmo_cands(self.ctx), shape=(max(nc, 1),)).view(PAIR_DTYPE)[:nc].copy() if nc else np.zeros(0, PAIR_DTYPE)
        lin = np.ctypeslib.as_array(self.L.kmo_lineage(self.ctx), shape=(max(nl, 1),)).view(PAIR_DTYPE)[:nl].copy() if nl else np.zeros(0, PAIR_DTYPE)
        return res, cands, lin

    def tails(self, res):
        """The text the reference writes after 'hdr\\tread\\t' for each result (cands must still be the
        ones of the last label() call)."""
        out = []
        buf = C.create_string_buffer(1 << 20)
        for i in range(len(res)):
            n = self.L.kmo_format_tail(self.ctx, res[i:i + 1].ctypes.data, buf, len(buf))
            out.append(buf.raw[:n].decode())
        return out

    def null_batch(self, seqs, first_index=0):
        """rand_read_label.cpp proc_line/construct_labels over reads whose GC bucket is (first_index + i) % 10;
        accumulates into the ctx."""
        bs = [s.encode() if isinstance(s, str) else s for s in seqs]
        offs = np.zeros(len(bs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in bs])
        rc = self.L.kmo_null_batch(self.ctx, b"".join(bs), offs.ctypes.data, len(bs), first_index)
        if rc:
            raise RuntimeError(f"kmo_null_batch rc={rc}")

    def null_table(self):
        """(tids ascending, max fraction [rows,10] f32, read count [rows,10] i32)"""
        n = self.L.kmo_null_rows(self.ctx)
        t = np.zeros(n, dtype=np.uint32)
        m = np.zeros((n, 10), dtype=np.float32)
        c = np.zeros((n, 10), dtype=np.int32)
        if n:
            self.L.kmo_null_get(self.ctx, t.ctypes.data, m.ctypes.data, c.ctypes.data)
        return t, m, c

    def null_reset(self):
        self.L.kmo_null_reset(self.ctx)

    def gene_label(self, seqs):
        """gene_label.cpp:217-301 per read -> list of (n_genes, valid_cnt, gene, count)."""
        out = []
        v, g, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        for s_ in seqs:
            b = s_ if isinstance(s_, bytes) else s_.encode("latin-1")
            n = self.L.kmo_gene_label_read(C.byref(self.cdb), b, len(b), C.byref(v), C.byref(g), C.byref(c))
            out.append((n, v.value, g.value, c.value))
        return out

    def lookup(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        offs = np.zeros(len(kmers) + 1, dtype=np.uint64)
        cap = max(1024, 64 * len(kmers))
        while True:
            ids = np.zeros(cap, dtype=np.uint32)
            n = self.L.kmo_lookup_batch(C.byref(self.cdb), kmers.ctypes.data, len(kmers), offs.ctypes.data, ids.ctypes.data, cap)
            if n == -1:
                cap *= 4
                continue
            if n < 0:
                raise RuntimeError(f"kmo_lookup_batch rc={n}")
            return offs, ids[:n]


def read_fasta_like_reference(path, fastq=False):
    """(headers, reads) exactly as read_label's single-producer parser pairs them
    (read_label.cpp:1651-1732): FASTA lines of length <= 1 are ignored, wrapped lines are joined, a read
    is pushed when the next '>' (or EOF) arrives; in FASTQ mode the '+'/'-' line pushes the read paired
    with the PREVIOUS record's header (the first record gets an empty header), and the quality line is
    skipped.  Empty header -> ``unknown_hdr:<n>`` with n = 1-based pop count."""
    out = []
    with open(path, "rb") as f:
        lines = f.read().decode("latin-1").split("\n")
    if lines and lines[-1] == "":
        lines.pop()                     # getline does not yield a final empty line after the last newline
    read_buff, hdr_buff, last_hdr = "", "", ""
    i = 0
    n = len(lines)
    finished = False
    while not finished:
        if i < n:
            line = lines[i]
            i += 1
        else:
            finished = True
            line = ""
        c0 = line[0] if line else "\0"
        if c0 == ">" or (fastq and c0 == "@"):
            last_hdr = hdr_buff
            hdr_buff = line[1:]
        if c0 != ">" and len(line) > 1 and not fastq:
            read_buff += line
            line = ""
            c0 = "\0"
        if fastq and c0 not in "@+-":
            read_buff += line
            line = ""
            c0 = "\0"
        if ((c0 == ">" or finished) or (fastq and c0 in "+-")) and len(read_buff) > 0:
            out.append((hdr_buff if finished else last_hdr, read_buff))
            read_buff = ""
            if fastq:
                i += 1                  # skip the quality line
    hdrs, seqs = [], []
    for k, (h, s) in enumerate(out):
        hdrs.append(h if h else f"unknown_hdr:{k + 1}")
        seqs.append(s)
    return hdrs, seqs


def encode_read(seq, k):
    L = lib()
    b = seq.encode() if isinstance(seq, str) else seq
    n = max(len(b) - k + 1, 0)
    km = np.zeros(max(n, 1), dtype=np.uint64)
    fl = np.zeros(max(n, 1), dtype=np.uint8)
    bs = C.c_int(0)
    valid = L.kmo_encode_read(b, len(b), k, km.ctypes.data, fl.ctypes.data, C.byref(bs))
    return valid, bs.value, km[:n], fl[:n]


def assemble_lines(hdrs, seqs, tails, prn_read=True):
    """Re-create the .out byte stream of one thread: 'hdr\\tread\\t' + tail (the tail may be empty --
    the silent-NoMatch quirk -- so records can run together exactly as in the reference)."""
    parts = []
    for h, s, t in zip(hdrs, seqs, tails):
        parts.append(f"{h}\t{s if prn_read else 'X'}\t{t}")
    return "".join(parts)


def gen_rand_reads(seed, n_reads, read_len, first=0):
    """The reads the reference rand_read_label draws on one thread after srand(seed) (rand_read_label.cpp:85-103,
    :687-699): list of n_reads lower-case strings."""
    g = GlibcRand()
    lib().kmo_srand(C.byref(g), C.c_uint(seed & 0xFFFFFFFF))
    buf = C.create_string_buffer(n_reads * read_len + 1)
    lib().kmo_gen_rand_reads(C.byref(g), first, n_reads, read_len, buf)
    raw = buf.raw[:n_reads * read_len].decode()
    return [raw[i * read_len:(i + 1) * read_len] for i in range(n_reads)]


def format_rand_lst(tids, mx, cnt):
    """The .rand_lst text (rand_read_label.cpp:745-754): 'tid' then ' max cnt' per GC bucket; floats through
    ostream<<float == %g."""
    out = []
    for i in range(len(tids)):
        out.append(str(int(tids[i])) + "".join(" %g %d" % (float(mx[i, b]), int(cnt[i, b])) for b in range(mx.shape[1])) + "\n")
    return "".join(out)
