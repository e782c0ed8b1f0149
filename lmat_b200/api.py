"""ctypes binding of libkmat's C ABI (include/kmat.h) -- the host-side mirror used by tests and bench.py.

Nothing here computes: every method is a thin call through the C ABI into the CUDA library.  If the
shared library is missing it is built in-tree with nvcc (lmat_b200.build); if no GPU is present the
compute calls raise KmatError (KMAT_ERR_NO_DEVICE) -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

ST_NAMES = ["SHORT_LEN", "SHORT_VALID", "NODBHITS", "SILENT", "PHIX", "LABELED", "ERROR"]
MATCH_NAMES = ["DirectMatch", "MultiMatch", "PartialMultiMatch", "NoMatch", "LCA_ERROR"]

RESULT_DTYPE = np.dtype([("status", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("valid_kmers", "<i4"),
                         ("cand_kmer_cnt", "<i4"), ("match", "<i4"), ("tid", "<u4"), ("score", "<f4"),
                         ("log_avg", "<f4"), ("stdev", "<f4"), ("n_cand", "<u4"), ("n_lin", "<u4"),
                         ("cand_off", "<u8"), ("lin_off", "<u8"), ("bin_sel", "<i4"), ("err", "<i4")])
PAIR_DTYPE = np.dtype([("tid", "<u4"), ("score", "<f4")])
RESULT32_DTYPE = np.dtype([("tid", "<u4"), ("score", "<f4"), ("log_avg", "<f4"), ("stdev", "<f4"), ("list_off", "<u4"), ("n_list", "<u2"),
                           ("valid_kmers", "<u2"), ("cand_kmer_cnt", "<u2"), ("flags", "<u2"), ("n_cand", "<u4")])
GENE_DTYPE = np.dtype([("status", "<i4"), ("valid_kmers", "<u4"), ("n_genes", "<u4"), ("gene", "<u4"), ("count", "<u4"), ("score", "<f4")])


class Opts(C.Structure):
    _fields_ = [("min_kmer", C.c_int32), ("min_fnd_kmer", C.c_int32), ("sdiff", C.c_float), ("hbias", C.c_float),
                ("min_score", C.c_float), ("max_count", C.c_int32), ("permissive", C.c_int32),
                ("phix_screen", C.c_int32), ("want_lineage", C.c_int32), ("rkmer_mode", C.c_int32)]


class BatchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("lookups", "hits", "list_hits", "list_ids", "probe_extra_buckets",
                                           "algorithmic_bytes", "reads_fast", "reads_slow", "reads_error")]


class KmatError(RuntimeError):
    def __init__(self, code, detail):
        super().__init__(f"libkmat error {code}: {detail}")
        self.code = code


EXPORTS = [
    "kmat_strerror", "kmat_last_error", "kmat_abi_version", "kmat_device_count", "kmat_device_memory", "kmat_table_from_sorteddb",
    "kmat_table_from_arrays", "kmat_table_open", "kmat_table_save", "kmat_table_size", "kmat_table_kmer_length",
    "kmat_table_tid_bytes", "kmat_table_build", "kmat_build_opts_default", "kmat_table_view", "kmat_table_free", "kmat_db_upload", "kmat_db_build_device",
    "kmat_shard_of", "kmat_table_device_bytes", "kmat_db_size", "kmat_db_bytes", "kmat_db_overflow", "kmat_db_kmer_length", "kmat_db_device", "kmat_db_free",
    "kmat_lookup_batch", "kmat_encode_batch", "kmat_inputs_load", "kmat_inputs_free", "kmat_opts_default",
    "kmat_ctx_create", "kmat_ctx_set_opts", "kmat_ctx_destroy", "kmat_label_batch", "kmat_label_batch_device",
    "kmat_pack_words", "kmat_pack_reads", "kmat_label_batch_packed", "kmat_result_expand",
    "kmat_ctx_sync", "kmat_ctx_last_stats", "kmat_ctx_set_stats", "kmat_ctx_set_pipeline", "kmat_gene_batch", "kmat_shard_encode", "kmat_shard_serve", "kmat_shard_finish", "kmat_ctx_device_results", "kmat_ctx_last_kernel_ms", "kmat_launch_count", "kmat_format_tail", "kmat_gather_bench", "kmat_gather_bench_peer",
    "kmat_set_l2_fetch_granularity", "kmat_reader_open", "kmat_reader_open_mt", "kmat_reader_close", "kmat_read_batch_new", "kmat_read_batch_new_pinned", "kmat_label_batch_text", "kmat_test_format_floats", "kmat_label_batch_packed_rl", "kmat_list_decode", "kmat_read_batch_free",
    "kmat_reader_next", "kmat_read_batch_view", "kmat_tally_class", "kmat_host_alloc", "kmat_host_free",
    "kmat_ctx_peer_export", "kmat_ctx_peer_attach", "kmat_comm_unique_id", "kmat_comm_init", "kmat_comm_free", "kmat_shard_label_device", "kmat_shard_label_batch",
    "kmat_kcov_create", "kmat_kcov_add", "kmat_kcov_finish", "kmat_kcov_query", "kmat_kcov_free",
    "kmat_null_reset", "kmat_null_batch", "kmat_null_random", "kmat_null_draw_reads", "kmat_null_fetch", "kmat_null_write",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("KMAT_LIB") or _build.build_lib()      # KMAT_LIB: an instrumented build (tools/asan_host.sh)
    L = C.CDLL(path)
    vp, u64p = C.c_void_p, C.c_void_p
    L.kmat_strerror.restype = C.c_char_p
    L.kmat_strerror.argtypes = [C.c_int]
    L.kmat_last_error.restype = C.c_char_p
    L.kmat_device_memory.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kmat_table_from_sorteddb.argtypes = [vp, C.c_uint64, C.c_int, vp, C.c_uint64, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmat_table_from_arrays.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmat_table_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.kmat_table_save.argtypes = [vp, C.c_char_p]
    L.kmat_table_size.restype = C.c_uint64
    L.kmat_table_size.argtypes = [vp]
    L.kmat_table_kmer_length.argtypes = [vp]
    L.kmat_table_tid_bytes.argtypes = [vp]
    L.kmat_table_view.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.kmat_table_free.argtypes = [vp]
    L.kmat_db_upload.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmat_db_build_device.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmat_shard_of.restype = C.c_uint32
    L.kmat_shard_of.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.kmat_db_size.restype = C.c_uint64
    L.kmat_db_size.argtypes = [vp]
    L.kmat_table_device_bytes.restype = C.c_uint64
    L.kmat_table_device_bytes.argtypes = [vp, C.c_int]
    L.kmat_db_bytes.restype = C.c_uint64
    L.kmat_db_bytes.argtypes = [vp]
    L.kmat_db_overflow.restype = C.c_uint64
    L.kmat_db_overflow.argtypes = [vp]
    L.kmat_db_kmer_length.argtypes = [vp]
    L.kmat_db_device.argtypes = [vp]
    L.kmat_db_free.argtypes = [vp]
    L.kmat_lookup_batch.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_encode_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, vp]
    L.kmat_inputs_load.argtypes = [C.c_char_p] * 8 + [C.POINTER(vp)]
    L.kmat_inputs_free.argtypes = [vp]
    L.kmat_opts_default.argtypes = [C.POINTER(Opts)]
    L.kmat_ctx_create.argtypes = [vp, vp, C.POINTER(Opts), C.POINTER(vp)]
    L.kmat_ctx_set_opts.argtypes = [vp, C.POINTER(Opts)]
    L.kmat_ctx_destroy.argtypes = [vp]
    L.kmat_label_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_label_batch_packed_rl.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_list_decode.argtypes = [vp, C.c_uint32, vp]
    L.kmat_list_decode.restype = C.c_uint32
    L.kmat_test_format_floats.argtypes = [C.c_int, vp, C.c_uint32, vp]
    L.kmat_label_batch_text.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vp, C.c_uint64, C.POINTER(C.c_uint64),
                                        C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64), vp]
    L.kmat_label_batch_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, vp, vp]
    L.kmat_comm_unique_id.argtypes = [vp]
    L.kmat_comm_init.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.kmat_comm_free.argtypes = [vp]
    L.kmat_comm_free.restype = None
    L.kmat_shard_label_device.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.kmat_shard_label_batch.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_pack_words.restype = C.c_uint64
    L.kmat_pack_words.argtypes = [C.c_uint64]
    L.kmat_pack_reads.argtypes = [vp, C.c_uint64, C.c_int, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_label_batch_packed.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmat_result_expand.restype = None
    L.kmat_result_expand.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, C.c_int, vp]
    L.kmat_ctx_sync.argtypes = [vp]
    L.kmat_ctx_last_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.kmat_ctx_set_stats.argtypes = [vp, C.c_int]
    L.kmat_ctx_set_pipeline.argtypes = [vp, C.c_int]
    L.kmat_gene_batch.argtypes = [vp, C.c_char_p, vp, C.c_uint32, vp]
    L.kmat_shard_encode.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(vp), vp, vp]
    L.kmat_shard_serve.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), vp, vp]
    L.kmat_shard_finish.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    L.kmat_ctx_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.kmat_ctx_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.kmat_launch_count.restype = C.c_uint64
    L.kmat_format_tail.argtypes = [vp, vp, vp, C.c_int, C.c_char_p, C.c_size_t]
    L.kmat_gather_bench.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.kmat_gather_bench_peer.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.kmat_set_l2_fetch_granularity.argtypes = [C.c_int, C.c_int]
    L.kmat_reader_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.kmat_reader_open_mt.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmat_reader_close.argtypes = [vp]
    L.kmat_read_batch_new.restype = vp
    L.kmat_read_batch_new_pinned.restype = vp
    L.kmat_read_batch_free.argtypes = [vp]
    L.kmat_reader_next.restype = C.c_int64
    L.kmat_reader_next.argtypes = [vp, C.c_uint32, C.c_uint64, vp]
    L.kmat_read_batch_view.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.kmat_tally_class.argtypes = [vp, C.c_float, C.c_int32]
    L.kmat_host_alloc.restype = vp
    L.kmat_host_alloc.argtypes = [C.c_size_t]
    L.kmat_host_free.argtypes = [vp]
    L.kmat_ctx_peer_export.argtypes = [vp, vp]
    L.kmat_ctx_peer_attach.argtypes = [vp, C.c_int, vp]
    L.kmat_kcov_create.argtypes = [C.c_int, vp, C.c_int, C.POINTER(vp)]
    L.kmat_kcov_add.argtypes = [vp, vp, vp, vp, C.c_uint32]
    L.kmat_kcov_finish.argtypes = [vp]
    L.kmat_kcov_query.argtypes = [vp, C.c_int, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.kmat_kcov_free.argtypes = [vp]
    L.kmat_null_reset.argtypes = [vp]
    L.kmat_null_batch.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64]
    L.kmat_null_random.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    L.kmat_null_draw_reads.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp]
    L.kmat_null_fetch.argtypes = [vp, vp, vp, vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.kmat_null_write.argtypes = [C.c_char_p, C.c_int, vp, vp, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        L = lib()
        raise KmatError(rc, f"{L.kmat_strerror(rc).decode()}: {L.kmat_last_error().decode(errors='replace')}")
    return rc


def device_count() -> int:
    return lib().kmat_device_count()


def _b(s):
    return None if s is None else os.fsencode(s)


class BuildOpts(C.Structure):
    _fields_ = [("kmer_length", C.c_int32), ("tax_histo_format", C.c_int32), ("tid_cutoff", C.c_int32), ("reserved", C.c_int32),
                ("stopper", C.c_uint64), ("map16", C.c_char_p), ("numrank", C.c_char_p), ("human_kmers", C.c_char_p),
                ("adaptor_kmers", C.c_char_p)]


class Table:
    """Host logical table (SortedDb contents flattened)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_arrays(cls, kmers, offs, ids, kmer_len=20, tid_bytes=2):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        h = C.c_void_p()
        _check(lib().kmat_table_from_arrays(kmers.ctypes.data, offs.ctypes.data, ids.ctypes.data, len(kmers), kmer_len, tid_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def from_sorteddb(cls, top_tier, kmer_table, storage, kmer_len=20, tid_bytes=2):
        """top_tier: uint64 array; kmer_table: raw bytes of kmer_record[]; storage: raw bytes (reference layout)."""
        h = C.c_void_p()
        bits = 13 if kmer_len == 20 else 9
        _check(lib().kmat_table_from_sorteddb(top_tier.ctypes.data, len(top_tier), bits, kmer_table.ctypes.data, len(kmer_table) // 8,
                                              storage.ctypes.data, len(storage), kmer_len, tid_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def open(cls, path, tid_bytes=2):
        h = C.c_void_p()
        _check(lib().kmat_table_open(_b(path), tid_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def build(cls, files, kmer_len=20, map16=None, tid_cutoff=0, numrank=None, human_kmers=None, adaptor_kmers=None,
              tax_histo_format=True, stopper=0):
        """kmat_table_build: the table make_db_table would build from these tax_histo files (-f -g -m -j -u -h -q)."""
        o = BuildOpts()
        lib().kmat_build_opts_default(C.byref(o))
        o.kmer_length, o.tax_histo_format, o.tid_cutoff, o.stopper = kmer_len, int(tax_histo_format), tid_cutoff, stopper
        o.map16, o.numrank, o.human_kmers, o.adaptor_kmers = (_b(x) if x else None for x in (map16, numrank, human_kmers, adaptor_kmers))
        arr = (C.c_char_p * len(files))(*[_b(f) for f in files])
        h = C.c_void_p()
        _check(lib().kmat_table_build(arr, len(files), C.byref(o), C.byref(h)))
        return cls(h)

    def save(self, path):
        _check(lib().kmat_table_save(self.h, _b(path)))

    @property
    def size(self):
        return lib().kmat_table_size(self.h)

    @property
    def kmer_length(self):
        return lib().kmat_table_kmer_length(self.h)

    def arrays(self):
        k, o, i = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n_ids = C.c_uint64()
        _check(lib().kmat_table_view(self.h, C.byref(k), C.byref(o), C.byref(i), C.byref(n_ids)))
        n = self.size
        kmers = np.ctypeslib.as_array(C.cast(k, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
        offs = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        ids = np.ctypeslib.as_array(C.cast(i, C.POINTER(C.c_uint32)), shape=(max(n_ids.value, 1),))[:n_ids.value].copy()
        return kmers, offs, ids

    def __del__(self):
        try:
            lib().kmat_table_free(self.h)
        except Exception:
            pass


class Db:
    """Device-resident hash table."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep

    @classmethod
    def upload(cls, table: Table, device=0, shard_index=0, shard_count=1):
        h = C.c_void_p()
        _check(lib().kmat_db_upload(table.h, device, shard_index, shard_count, C.byref(h)))
        return cls(h)

    @classmethod
    def build_device(cls, device, kmer_len, tid_bytes, n, d_kmers_ptr, d_payload_ptr, d_pool_ptr, pool_words, n_stored_ids=0,
                     shard_index=0, shard_count=1):
        h = C.c_void_p()
        _check(lib().kmat_db_build_device(device, kmer_len, tid_bytes, n, d_kmers_ptr, d_payload_ptr, d_pool_ptr, pool_words, n_stored_ids,
                                          shard_index, shard_count, C.byref(h)))
        return cls(h)

    @property
    def size(self):
        return lib().kmat_db_size(self.h)

    @property
    def bytes(self):
        return lib().kmat_db_bytes(self.h)

    @property
    def overflow(self):
        """k-mers of a two-level table that live in its second level (their first-level sector was full)"""
        return lib().kmat_db_overflow(self.h)

    @property
    def kmer_length(self):
        return lib().kmat_db_kmer_length(self.h)

    def gene_label(self, seqs):
        """kmat_gene_batch over a list of reads -> numpy array of GENE_DTYPE."""
        bases, offs = pack_reads(seqs)
        out = np.zeros(len(seqs), dtype=GENE_DTYPE)
        _check(lib().kmat_gene_batch(self.h, bases, offs.ctypes.data, len(seqs), out.ctypes.data))
        return out

    def lookup(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        offs = np.zeros(len(kmers) + 1, dtype=np.uint64)
        n_ids = C.c_uint64()
        cap = max(1024, 4 * len(kmers))
        while True:
            ids = np.zeros(cap, dtype=np.uint32)
            rc = lib().kmat_lookup_batch(self.h, kmers.ctypes.data, len(kmers), offs.ctypes.data, ids.ctypes.data, cap, C.byref(n_ids))
            if rc == -10:
                cap = int(n_ids.value)
                continue
            _check(rc)
            return offs, ids[:n_ids.value]

    def encode(self, seqs):
        blob, offs = pack_reads(seqs)
        total = int(offs[-1])
        kmers = np.zeros(total + 1, dtype=np.uint64)
        flags = np.zeros(total + 1, dtype=np.uint8)
        valid = np.zeros(len(seqs), dtype=np.int32)
        bins = np.zeros(len(seqs), dtype=np.int32)
        _check(lib().kmat_encode_batch(self.h, blob, offs.ctypes.data, len(seqs), kmers.ctypes.data, flags.ctypes.data, valid.ctypes.data, bins.ctypes.data))
        return kmers, flags, valid, bins, offs

    def __del__(self):
        try:
            lib().kmat_db_free(self.h)
        except Exception:
            pass


class Inputs:
    def __init__(self, tree=None, depth=None, rank=None, map16=None, numrank=None, plasmids=None, null_lst=None, lmat_dir=None):
        self.h = C.c_void_p()
        _check(lib().kmat_inputs_load(_b(tree), _b(depth), _b(rank), _b(map16), _b(numrank), _b(plasmids), _b(null_lst), _b(lmat_dir), C.byref(self.h)))

    def __del__(self):
        try:
            lib().kmat_inputs_free(self.h)
        except Exception:
            pass


def pack_reads(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    return b"".join(bs), offs


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().kmat_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Ctx:
    def __init__(self, db: Db, inputs: Inputs, opts: Opts | None = None):
        self.db, self.inputs = db, inputs
        self.opts = opts or default_opts()
        self.h = C.c_void_p()
        _check(lib().kmat_ctx_create(db.h, inputs.h, C.byref(self.opts), C.byref(self.h)))

    def set_opts(self, **kw):
        for k, v in kw.items():
            setattr(self.opts, k, v)
        _check(lib().kmat_ctx_set_opts(self.h, C.byref(self.opts)))

    def label(self, seqs=None, blob=None, offs=None):
        """Label a batch through kmat_label_batch (host buffers in, host buffers out)."""
        if seqs is not None:
            blob, offs = pack_reads(seqs)
        n = len(offs) - 1
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = max(4096, 32 * n)
        n_c, n_l = C.c_uint64(), C.c_uint64()
        while True:
            cands = np.zeros(cap, dtype=PAIR_DTYPE)
            lin = np.zeros(cap if self.opts.want_lineage else 1, dtype=PAIR_DTYPE)
            ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
            rc = lib().kmat_label_batch(self.h, ptr, offs.ctypes.data, n, res.ctypes.data, cands.ctypes.data, len(cands), C.byref(n_c),
                                        lin.ctypes.data if self.opts.want_lineage else None, len(lin), C.byref(n_l))
            if rc == -10:
                cap = int(max(n_c.value, n_l.value)) + 16
                continue
            _check(rc)
            return res, cands[:n_c.value], lin[:n_l.value]

    def label_text(self, seqs=None, blob=None, offs=None, prn_all=True, text_cap=None):
        """kmat_label_batch_text: label() plus the output-line tails formatted on the device (K5).  Returns (results, candidates,
        lineage, tails, n_on_host): tails[i] is the device text, or None where the device left read i to kmat_format_tail."""
        if seqs is not None:
            blob, offs = pack_reads(seqs)
        n = len(offs) - 1
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = max(4096, 32 * n)
        n_c, n_l, n_t = C.c_uint64(), C.c_uint64(), C.c_uint64()
        text = np.zeros(max(1, 256 * n + 4096 if text_cap is None else text_cap), dtype=np.uint8)
        ref = np.zeros(max(1, n), dtype=np.uint64)
        while True:
            cands = np.zeros(cap, dtype=PAIR_DTYPE)
            lin = np.zeros(cap if self.opts.want_lineage else 1, dtype=PAIR_DTYPE)
            ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
            rc = lib().kmat_label_batch_text(self.h, ptr, offs.ctypes.data, n, res.ctypes.data, cands.ctypes.data, len(cands), C.byref(n_c),
                                             lin.ctypes.data if self.opts.want_lineage else None, len(lin), C.byref(n_l), int(prn_all),
                                             text.ctypes.data, 0 if text_cap == 0 else len(text), C.byref(n_t), ref.ctypes.data)
            if rc == -10:
                cap = int(max(n_c.value, n_l.value)) + 16
                continue
            _check(rc)
            break
        raw = text.tobytes()
        tails, on_host = [], 0
        for i in range(n):
            r = int(ref[i])
            if r == 0xFFFFFFFFFFFFFFFF:
                tails.append(None)
                on_host += 1
            else:
                o, ln = r >> 24, r & 0xFFFFFF
                assert o + ln <= n_t.value
                tails.append(raw[o:o + ln].decode())
        return res, cands[:n_c.value], lin[:n_l.value], tails, on_host

    def label_packed(self, seqs=None, blob=None, offs=None, threads=4):
        """The compact interface: kmat_pack_reads on the host, kmat_label_batch_packed, 32-byte results expanded back to the
        64-byte records (so that tails() applies).  Returns (results, candidates, lineage) like label()."""
        if seqs is not None:
            blob, offs = pack_reads(seqs)
        n = len(offs) - 1
        total = int(offs[n])
        ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
        codes = np.zeros(max(1, lib().kmat_pack_words(total)), dtype=np.uint32)
        n_inv = C.c_uint64()
        inv = np.zeros(1024, dtype=np.uint64)
        rc = lib().kmat_pack_reads(ptr, total, threads, codes.ctypes.data, inv.ctypes.data, len(inv), C.byref(n_inv))
        if rc == -10:
            inv = np.zeros(n_inv.value, dtype=np.uint64)
            rc = lib().kmat_pack_reads(ptr, total, threads, codes.ctypes.data, inv.ctypes.data, len(inv), C.byref(n_inv))
        _check(rc)
        res32 = np.zeros(n, dtype=RESULT32_DTYPE)
        cap = max(4096, 32 * n)
        n_l = C.c_uint64()
        while True:
            lst = np.zeros(cap, dtype=PAIR_DTYPE)
            rc = lib().kmat_label_batch_packed(self.h, codes.ctypes.data, inv.ctypes.data, n_inv.value, offs.ctypes.data, n, res32.ctypes.data,
                                               lst.ctypes.data, len(lst), C.byref(n_l))
            if rc == -10:
                cap = int(n_l.value) + 16
                continue
            _check(rc)
            break
        res = np.zeros(n, dtype=RESULT_DTYPE)
        k = self.db.kmer_length
        for i in range(n):
            lib().kmat_result_expand(res32[i:i + 1].ctypes.data, int(offs[i + 1] - offs[i]), k, self.opts.min_kmer, self.opts.want_lineage, res[i:i + 1].ctypes.data)
        lst = lst[:n_l.value]
        empty = np.zeros(0, dtype=PAIR_DTYPE)
        return (res, empty, lst) if self.opts.want_lineage else (res, lst, empty)

    def label_packed_rl(self, seqs=None, blob=None, offs=None, threads=4):
        """kmat_label_batch_packed_rl: the compact interface with run-length lists, decoded back (kmat_list_decode) into the
        pair arrays label_packed() returns.  Returns (results, candidates, lineage, words used)."""
        if seqs is not None:
            blob, offs = pack_reads(seqs)
        n = len(offs) - 1
        total = int(offs[n])
        ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
        codes = np.zeros(max(1, lib().kmat_pack_words(total)), dtype=np.uint32)
        n_inv = C.c_uint64()
        inv = np.zeros(1024, dtype=np.uint64)
        rc = lib().kmat_pack_reads(ptr, total, threads, codes.ctypes.data, inv.ctypes.data, len(inv), C.byref(n_inv))
        if rc == -10:
            inv = np.zeros(n_inv.value, dtype=np.uint64)
            rc = lib().kmat_pack_reads(ptr, total, threads, codes.ctypes.data, inv.ctypes.data, len(inv), C.byref(n_inv))
        _check(rc)
        res32 = np.zeros(n, dtype=RESULT32_DTYPE)
        cap = max(4096, 16 * n)
        n_w = C.c_uint64()
        while True:
            words = np.zeros(cap, dtype=np.uint32)
            rc = lib().kmat_label_batch_packed_rl(self.h, codes.ctypes.data, inv.ctypes.data, n_inv.value, offs.ctypes.data, n, res32.ctypes.data,
                                                  words.ctypes.data, len(words), C.byref(n_w))
            if rc == -10:
                cap = int(n_w.value) + 16
                continue
            _check(rc)
            break
        res = np.zeros(n, dtype=RESULT_DTYPE)
        k = self.db.kmer_length
        lst = np.zeros(max(1, int(res32["n_list"].astype(np.int64).sum())), dtype=PAIR_DTYPE)
        at = 0
        for i in range(n):
            nl = int(res32["n_list"][i])
            used = lib().kmat_list_decode(words[int(res32["list_off"][i]):].ctypes.data, nl, lst[at:].ctypes.data) if nl else 0
            assert int(res32["list_off"][i]) + used <= n_w.value
            r32 = res32[i:i + 1].copy()
            r32["list_off"] = at
            lib().kmat_result_expand(r32.ctypes.data, int(offs[i + 1] - offs[i]), k, self.opts.min_kmer, self.opts.want_lineage, res[i:i + 1].ctypes.data)
            at += nl
        lst = lst[:at]
        empty = np.zeros(0, dtype=PAIR_DTYPE)
        return ((res, empty, lst) if self.opts.want_lineage else (res, lst, empty)) + (int(n_w.value),)

    def tails(self, res, cands, lin, prn_all=True):
        out = []
        buf = C.create_string_buffer(1 << 20)
        cp = cands.ctypes.data if len(cands) else None
        lp = lin.ctypes.data if len(lin) else None
        for i in range(len(res)):
            n = lib().kmat_format_tail(res[i:i + 1].ctypes.data, cp, lp, int(prn_all), buf, len(buf))
            _check(n)
            out.append(buf.raw[:n].decode())
        return out

    # ---- DB-sharded mode, direct variant: probes go to the owner shard's memory (CUDA IPC / peer access over NVLink)
    def peer_export(self):
        """-> 512-byte blob describing this rank's shard (numpy uint8), to be all-gathered."""
        blob = np.zeros(512, dtype=np.uint8)
        _check(lib().kmat_ctx_peer_export(self.h, blob.ctypes.data))
        return blob

    def peer_attach(self, blobs):
        """blobs: the exports of all shards, indexed by shard ([n_shards, 512] uint8)."""
        b = np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.uint8) for x in blobs]))
        _check(lib().kmat_ctx_peer_attach(self.h, len(b), b.ctypes.data))

    # ---- null-model generation (rand_read_label); the ctx must have been created with rkmer_mode = 1
    def null_reset(self):
        _check(lib().kmat_null_reset(self.h))

    def null_batch(self, seqs=None, blob=None, offs=None, first_index=0):
        if seqs is not None:
            blob, offs = pack_reads(seqs)
        ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
        _check(lib().kmat_null_batch(self.h, ptr, offs.ctypes.data, len(offs) - 1, first_index))

    def null_random(self, seed, first_index, n_reads, read_len):
        _check(lib().kmat_null_random(self.h, seed, first_index, n_reads, read_len))

    def null_table(self):
        """(tids ascending, max fraction [rows, 10] f32, read counts [rows, 10] u64, reads the kernels could not process)"""
        n, err = C.c_uint32(), C.c_uint64()
        rc = lib().kmat_null_fetch(self.h, None, None, None, 0, C.byref(n), C.byref(err))
        if rc not in (0, -10):
            _check(rc)
        t = np.zeros(n.value, dtype=np.uint32)
        m = np.zeros((n.value, 10), dtype=np.float32)
        c = np.zeros((n.value, 10), dtype=np.uint64)
        _check(lib().kmat_null_fetch(self.h, t.ctypes.data, m.ctypes.data, c.ctypes.data, n.value, C.byref(n), C.byref(err)))
        return t, m, c, err.value

    def label_device(self, d_bases_ptr, d_offs_ptr, n_reads, total_bases, max_read_len, d_out_ptr=None, stream=None):
        """kmat_label_batch_device: inputs already in HBM, results stay on the device; asynchronous."""
        _check(lib().kmat_label_batch_device(self.h, d_bases_ptr, d_offs_ptr, n_reads, total_bases, max_read_len, d_out_ptr, stream))

    # ---- DB-sharded mode: the three device phases of one round (the exchange between them is lmat_b200.sharded's)
    def shard_encode(self, d_bases_ptr, d_offs_ptr, n_reads, total_bases, max_read_len, n_shards, stream=None):
        """-> (device pointer of the queries grouped by owner, counts[n_shards])"""
        q = C.c_void_p()
        counts = np.zeros(n_shards, dtype=np.uint64)
        _check(lib().kmat_shard_encode(self.h, d_bases_ptr, d_offs_ptr, n_reads, total_bases, max_read_len, n_shards, C.byref(q), counts.ctypes.data, stream))
        return q.value or 0, counts

    def shard_serve(self, d_queries_ptr, counts, stream=None):
        """-> (device pointer of the hit words, device pointer of the list payload, payload_counts[n_shards])"""
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        rep, pay = C.c_void_p(), C.c_void_p()
        pc = np.zeros(len(counts), dtype=np.uint64)
        _check(lib().kmat_shard_serve(self.h, d_queries_ptr, counts.ctypes.data, len(counts), C.byref(rep), C.byref(pay), pc.ctypes.data, stream))
        return rep.value or 0, pay.value or 0, pc

    def shard_finish(self, d_reply_ptr, d_payload_ptr, payload_counts, d_out_ptr=None, stream=None):
        pc = np.ascontiguousarray(payload_counts, dtype=np.uint64)
        _check(lib().kmat_shard_finish(self.h, d_reply_ptr or None, d_payload_ptr or None, pc.ctypes.data, len(pc), d_out_ptr, stream))

    def device_results(self):
        """-> (device pointer of the results of the last pass run with d_out=None, device pointer of the candidate pairs, n_cands)"""
        o, cd, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(lib().kmat_ctx_device_results(self.h, C.byref(o), C.byref(cd), C.byref(n)))
        return o.value or 0, cd.value or 0, n.value

    def kernel_ms(self):
        a, b, d = C.c_float(), C.c_float(), C.c_float()
        _check(lib().kmat_ctx_last_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(d)))
        return a.value, b.value, d.value

    def set_pipeline(self, sub_batches):
        _check(lib().kmat_ctx_set_pipeline(self.h, int(sub_batches)))

    def set_stats(self, enable):
        _check(lib().kmat_ctx_set_stats(self.h, int(enable)))

    def sync(self):
        _check(lib().kmat_ctx_sync(self.h))

    def stats(self) -> BatchStats:
        s = BatchStats()
        _check(lib().kmat_ctx_last_stats(self.h, C.byref(s)))
        return s

    def __del__(self):
        try:
            lib().kmat_ctx_destroy(self.h)
        except Exception:
            pass


class Comm:
    """NCCL communicator of the DB-sharded mode, owned by libkmat (kmat_comm_*).  rank r's Ctx must sit on shard r."""

    @staticmethod
    def _preload():
        # In a Python process torch's own NCCL (same soname, newer) must be the one in the address space before libkmat
        # dlopen()s "libnccl.so.2": a system copy loaded first would shadow it for torch's later import.
        try:
            import torch  # noqa: F401
        except Exception:
            pass

    def __init__(self, device, rank, world, unique_id):
        Comm._preload()
        self.h = C.c_void_p()
        self.rank, self.world = rank, world
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert uid.size == 128
        _check(lib().kmat_comm_init(device, rank, world, uid.ctypes.data, C.byref(self.h)))

    @staticmethod
    def unique_id():
        Comm._preload()
        uid = np.zeros(128, dtype=np.uint8)
        _check(lib().kmat_comm_unique_id(uid.ctypes.data))
        return uid.tobytes()

    def label_device(self, ctx, d_bases_ptr, h_offs, d_offs_ptr, n_reads, d_out_ptr, round_reads=0, stream=None):
        """COLLECTIVE: kmat_shard_label_device.  h_offs: numpy uint64 [n_reads + 1].  Returns (lookups, served, payload_words, rounds)."""
        st = np.zeros(4, dtype=np.uint64)
        ho = np.ascontiguousarray(h_offs, dtype=np.uint64)
        _check(lib().kmat_shard_label_device(ctx.h, self.h, d_bases_ptr, ho.ctypes.data, d_offs_ptr, n_reads, round_reads, d_out_ptr, st.ctypes.data, stream))
        return tuple(int(x) for x in st)

    def label(self, ctx, seqs):
        """COLLECTIVE: kmat_shard_label_batch (host buffers).  Returns (results, candidates, lineage) like Ctx.label."""
        blob, offs = pack_reads(seqs)
        n = len(offs) - 1
        res = np.zeros(max(1, n), dtype=RESULT_DTYPE)
        cap = max(4096, 32 * n)
        n_c, n_l = C.c_uint64(), C.c_uint64()
        cands = np.zeros(cap, dtype=PAIR_DTYPE)
        lin = np.zeros(cap if ctx.opts.want_lineage else 1, dtype=PAIR_DTYPE)
        _check(lib().kmat_shard_label_batch(ctx.h, self.h, blob, offs.ctypes.data, n, res.ctypes.data, cands.ctypes.data, len(cands), C.byref(n_c),
                                            lin.ctypes.data if ctx.opts.want_lineage else None, len(lin), C.byref(n_l)))
        return res[:n], cands[:n_c.value], lin[:n_l.value]

    def __del__(self):
        try:
            if self.h:
                lib().kmat_comm_free(self.h)
                self.h = None
        except Exception:
            pass


def gather_bench(device=0, span_bytes=1 << 30, access_bytes=8, n_gathers=1 << 28, iters=5):
    g, s = C.c_double(), C.c_double()
    _check(lib().kmat_gather_bench(device, span_bytes, access_bytes, n_gathers, iters, C.byref(g), C.byref(s)))
    return g.value, s.value


def gather_bench_peer(device, mem_device, span_bytes=1 << 30, access_bytes=32, n_gathers=1 << 27, iters=3):
    g, s = C.c_double(), C.c_double()
    _check(lib().kmat_gather_bench_peer(device, mem_device, span_bytes, access_bytes, n_gathers, iters, C.byref(g), C.byref(s)))
    return g.value, s.value


def read_file(path, fastq=False, max_reads=1 << 20, max_bases=1 << 28, threads=1, pinned=False):
    """(headers, reads) through kmat_reader_* -- the host parser that replaces read_label.cpp:1651-1732.
    pinned: the batch hands its bases out from page-locked memory (what the read_label binary does)."""
    L = lib()
    r, hdrs, seqs = C.c_void_p(), [], []
    _check(L.kmat_reader_open_mt(_b(path), int(fastq), int(threads), C.byref(r)))
    b = C.c_void_p(L.kmat_read_batch_new_pinned() if pinned else L.kmat_read_batch_new())
    try:
        while True:
            n = L.kmat_reader_next(r, max_reads, max_bases, b)
            _check(n)
            if n == 0:
                break
            bp, op_, hp, hop = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            nn, first = C.c_uint32(), C.c_uint64()
            _check(L.kmat_read_batch_view(b, C.byref(bp), C.byref(op_), C.byref(hp), C.byref(hop), C.byref(nn), C.byref(first)))
            offs = np.ctypeslib.as_array(C.cast(op_, C.POINTER(C.c_uint64)), shape=(n + 1,))
            hoffs = np.ctypeslib.as_array(C.cast(hop, C.POINTER(C.c_uint64)), shape=(n + 1,))
            bases = C.string_at(bp, int(offs[n]))
            hd = C.string_at(hp, int(hoffs[n]))
            for i in range(n):
                seqs.append(bases[int(offs[i]):int(offs[i + 1])].decode("latin-1"))
                hdrs.append(hd[int(hoffs[i]):int(hoffs[i + 1])].decode("latin-1"))
    finally:
        L.kmat_read_batch_free(b)
        L.kmat_reader_close(r)
    return hdrs, seqs


def null_draw_reads(seed, first_index, n_reads, read_len, device=0):
    """The reads kmat_null_random draws on the device for run indices first_index.. (test hook): list of bytes."""
    buf = np.zeros(n_reads * read_len, dtype=np.uint8)
    _check(lib().kmat_null_draw_reads(device, seed, first_index, n_reads, read_len, buf.ctypes.data))
    raw = buf.tobytes()
    return [raw[i * read_len:(i + 1) * read_len] for i in range(n_reads)]


def null_write(path, sets):
    """kmat_null_write: merge (tids, max, counts) row sets (max / sum) and write the .rand_lst text."""
    n = len(sets)
    keep = [(np.ascontiguousarray(t, np.uint32), np.ascontiguousarray(m, np.float32), np.ascontiguousarray(c, np.uint64)) for t, m, c in sets]
    tp = (C.c_void_p * n)(*[k[0].ctypes.data for k in keep])
    mp = (C.c_void_p * n)(*[k[1].ctypes.data for k in keep])
    cp = (C.c_void_p * n)(*[k[2].ctypes.data for k in keep])
    nr = np.array([len(k[0]) for k in keep], dtype=np.uint32)
    _check(lib().kmat_null_write(_b(path), n, tp, mp, cp, nr.ctypes.data))


class KmerCov:
    """kmat_kcov_*: per (k, group) distinct canonical k-mers of the added reads, each counted once per read."""

    def __init__(self, k_sizes, device=0):
        self.k = np.ascontiguousarray(k_sizes, dtype=np.int32)
        self.h = C.c_void_p()
        _check(lib().kmat_kcov_create(device, self.k.ctypes.data, len(self.k), C.byref(self.h)))

    def __del__(self):
        try:
            lib().kmat_kcov_free(self.h)
        except Exception:
            pass

    def add(self, seqs, groups):
        blob, offs = pack_reads(seqs)
        g = np.ascontiguousarray(groups, dtype=np.uint32)
        ptr = blob if isinstance(blob, (bytes, bytearray)) else blob.ctypes.data
        _check(lib().kmat_kcov_add(self.h, ptr, offs.ctypes.data, g.ctypes.data, len(g)))

    def finish(self):
        _check(lib().kmat_kcov_finish(self.h))

    def query(self, k_index, group):
        """-> (distinct, total, {count: number of k-mers})"""
        d, t, n = C.c_uint64(), C.c_uint64(), C.c_uint32()
        _check(lib().kmat_kcov_query(self.h, k_index, group, C.byref(d), C.byref(t), None, None, 0, C.byref(n)))
        hc = np.zeros(max(1, n.value), dtype=np.uint32)
        hn = np.zeros(max(1, n.value), dtype=np.uint64)
        _check(lib().kmat_kcov_query(self.h, k_index, group, C.byref(d), C.byref(t), hc.ctypes.data, hn.ctypes.data, len(hc), C.byref(n)))
        return d.value, t.value, {int(hc[i]): int(hn[i]) for i in range(n.value)}
