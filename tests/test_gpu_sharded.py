"""GPU (-m gpu): DB-sharded mode (SURVEY 8(e) mode B).  The table is split into `world` shards by kmat_shard_of; the
ranks run as threads of this process on one GPU (LocalExchange), each with its own shard, context and reads, and go
through the same rounds (encode -> all-to-all -> serve -> all-to-all -> finish) as the one-process-per-GPU NCCL path.
The labels must equal the replicated table's, which test_gpu_parity.py pins to the reference's own output."""
import threading

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api, sharded
from oracle import oracle_py as op
from test_gpu_parity import make_ctx

pytestmark = pytest.mark.gpu


def run_ranks(world, fn):
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            out[r] = fn(r)
        except BaseException as e:       # noqa: BLE001 - re-raised in the main thread
            err[r] = e
            raise
    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    for e in err:
        if e is not None:
            raise e
    assert all(o is not None for o in out), "a rank did not finish"
    return out


@pytest.mark.parametrize("world,opts,slots", [(2, "run_rl", 1), (2, "permissive", 1), (2, "prune3", 2), (3, "run_rl", 2), (3, "permissive", 2), (3, "prune3", 1), (8, "run_rl", 2)])
def test_sharded_labels_equal_replicated(golden_lists, world, opts, slots):
    import torch
    g = golden_lists
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    full = api.Db.upload(t)
    shards = [api.Db.upload(t, 0, r, world) for r in range(world)]
    assert sum(db.size for db in shards) == full.size == len(g.kmers)
    assert all(0 < db.size < full.size for db in shards)
    for r in range(world):                               # kmat_shard_of is the partition
        own = np.array([api.lib().kmat_shard_of(int(k), g.kmer_len, world) == r for k in g.kmers[:2000]])
        offs, ids = shards[r].lookup(g.kmers[:2000])
        assert np.array_equal(np.diff(offs.astype(np.int64)) > 0, own)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs + ["", "ACGT", "N" * 40, seqs[0][:25]]
    ref_ctx = make_ctx(g, full, opts)
    res, cands, lin = ref_ctx.label(seqs)
    want = ref_ctx.tails(res, cands, lin, prn_all=True)
    # uneven split: rank 0 gets half of the reads, so the others keep serving after their own reads are done
    cut = [0, len(seqs) // 2] + [len(seqs) // 2 + (len(seqs) - len(seqs) // 2) * (i + 1) // (world - 1) for i in range(world - 1)]
    ctxs = [make_ctx(g, shards[r], opts) for r in range(world)]
    ctxs2 = [make_ctx(g, shards[r], opts) for r in range(world)] if slots == 2 else None      # two-slot pipeline of the driver
    grp = sharded.LocalGroup(world)

    def rank(r):
        ex = sharded.LocalExchange(grp, r, sync=lambda: torch.cuda.synchronize())
        mine = seqs[cut[r]:cut[r + 1]]
        rr, cc, lab = sharded.label_sequences(ctxs[r], ex, "cuda:0", mine, world, round_reads=97, ctx2=ctxs2[r] if ctxs2 else None)
        return ctxs[r].tails(rr, cc, np.zeros(0, dtype=api.PAIR_DTYPE), prn_all=True), lab

    outs = run_ranks(world, rank)
    got = [t for tails, _ in outs for t in tails]
    assert got == want
    labs = [lab for _, lab in outs]
    assert sum(l.lookups for l in labs) == sum(l.served for l in labs) > 0
    assert len({l.rounds for l in labs}) == 1            # every rank took part in every round


def test_sharded_single_rank_is_a_plain_pass(golden_small):
    """world = 1: the exchange is the identity; exercises the three phases against the replicated path."""
    import torch
    g = golden_small
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes))
    ctx = make_ctx(g, db, "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    res, cands, lin = ctx.label(seqs)
    want = ctx.tails(res, cands, lin, prn_all=True)
    grp = sharded.LocalGroup(1)
    rr, cc, lab = sharded.label_sequences(ctx, sharded.LocalExchange(grp, 0, sync=lambda: torch.cuda.synchronize()), "cuda:0", seqs, 1)
    assert ctx.tails(rr, cc, np.zeros(0, dtype=api.PAIR_DTYPE), prn_all=True) == want
    mine = op.assemble_lines(hdrs, seqs, ctx.tails(rr, cc, np.zeros(0, dtype=api.PAIR_DTYPE), prn_all=True))
    assert mine == g.golden_out("run_rl")


@pytest.mark.parametrize("lists", ["replicate", "fetch"])
@pytest.mark.parametrize("world,opts", [(2, "run_rl"), (3, "permissive"), (3, "prune3"), (8, "run_rl")])
def test_direct_sharded_labels_equal_replicated(golden_lists, world, opts, lists, monkeypatch):
    """Direct variant: every virtual rank maps all shards (same process, same device: plain pointers) and labels its reads
    with the ordinary kmat_label_batch; the probe kernel sends each gather to the owner shard's arrays."""
    monkeypatch.setenv("KMAT_PEER_LISTS", lists)          # list pools copied to every rank / records fetched from their owners per pass
    monkeypatch.setenv("KMAT_TEST_TIGHT_TABLE", "1")      # nearly full shards: displaced keys and the per-shard stash are exercised too
    g = golden_lists
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    shards = [api.Db.upload(t, 0, r, world) for r in range(world)]
    monkeypatch.delenv("KMAT_TEST_TIGHT_TABLE")
    full = api.Db.upload(t)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs + ["", "ACGT", "N" * 40, seqs[0][:25], seqs[1] * 3, seqs[2] * 9]        # every K1 / K3 variant
    ref_ctx = make_ctx(g, full, opts)
    res, cands, lin = ref_ctx.label(seqs)
    want = ref_ctx.tails(res, cands, lin, prn_all=True)
    ctxs = [make_ctx(g, shards[r], opts) for r in range(world)]
    with pytest.raises(api.KmatError):                    # a shard alone does not answer for the whole table
        ctxs[0].label(seqs[:3])
    sharded.attach_peers_local(ctxs)
    for r in range(world):
        mine = seqs[r::world]
        rr, cc, ll = ctxs[r].label(mine)
        assert ctxs[r].tails(rr, cc, ll, prn_all=True) == want[r::world], r
    # mismatched options are refused
    other = make_ctx(g, shards[0], "prune3" if opts != "prune3" else "run_rl")
    with pytest.raises(api.KmatError):
        other.peer_attach([c.peer_export() for c in ctxs])


def test_direct_sharded_record_buffer_overflow_is_reported_and_recovers(golden_lists, monkeypatch):
    """The local list-record buffer of the direct variant: hits it has no room for are dropped AND reported
    (KMAT_ERR_OVERFLOW after the internal retries), never mislabeled; the buffer doubles until the batch fits."""
    g = golden_lists
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    world = 2
    shards = [api.Db.upload(t, 0, r, world) for r in range(world)]
    full = api.Db.upload(t)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    ref_ctx = make_ctx(g, full, "run_rl")
    res, cands, lin = ref_ctx.label(seqs)
    want = ref_ctx.tails(res, cands, lin, prn_all=True)
    ctxs = [make_ctx(g, shards[r], "run_rl") for r in range(world)]
    monkeypatch.setenv("KMAT_PEER_LISTS", "fetch")
    sharded.attach_peers_local(ctxs)
    monkeypatch.setenv("KMAT_TEST_PEER_RECS", "16")
    import ctypes as C
    blob, offs = api.pack_reads(seqs)
    out = np.zeros(len(seqs), dtype=api.RESULT_DTYPE)
    cands = np.zeros(64 * len(seqs), dtype=api.PAIR_DTYPE)
    n_c = C.c_uint64()
    rc = api.lib().kmat_label_batch(ctxs[0].h, blob, offs.ctypes.data, len(seqs), out.ctypes.data, cands.ctypes.data, len(cands), C.byref(n_c), None, 0, None)
    assert rc == -10, rc                                   # 16, 32, 64 words: still too small after the internal retries
    assert b"list-record buffer" in api.lib().kmat_last_error()
    rr, cc, ll = ctxs[0].label(seqs)                       # the Python wrapper retries on KMAT_ERR_OVERFLOW; the buffer doubles each time
    assert ctxs[0].tails(rr, cc, ll, prn_all=True) == want


def _label_through_comm(ctx, comm, device, seqs, round_reads):
    """One rank's reads through kmat_shard_label_device; returns the result records (numpy) of this rank."""
    import torch
    lens = np.array([len(x) for x in seqs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    flat = b"".join(x.encode("latin-1") for x in seqs) or b"\0"
    with torch.cuda.device(device):
        bases = torch.frombuffer(bytearray(flat), dtype=torch.uint8).to(device)
        d_offs = torch.as_tensor(offs.astype(np.int64), device=device)
        d_out = torch.zeros(max(1, len(seqs)) * api.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        st = comm.label_device(ctx, bases.data_ptr(), offs, d_offs.data_ptr(), len(seqs), d_out.data_ptr(), round_reads=round_reads)
        ctx.sync()
        torch.cuda.synchronize(device)
        res = d_out.cpu().numpy().view(api.RESULT_DTYPE)[:len(seqs)].copy()
    return res, st


def test_nccl_exchange_single_rank(golden_lists):
    """kmat_comm_* / kmat_shard_label_device with a world of one: NCCL send/recv to self; labels of the replicated path."""
    g = golden_lists
    db = api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes))
    ctx = make_ctx(g, db, "run_rl")
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    seqs = seqs + ["", "ACGT", "N" * 40]
    want, _, _ = ctx.label(seqs)
    comm = api.Comm(0, 0, 1, api.Comm.unique_id())
    res, st = _label_through_comm(ctx, comm, "cuda:0", seqs, round_reads=97)
    for f in ("status", "tid", "match", "valid_kmers", "cand_kmer_cnt", "n_cand"):
        assert np.array_equal(res[f], want[f]), f
    assert np.array_equal(res["score"].view(np.uint32), want["score"].view(np.uint32))
    assert st[0] == st[1] > 0 and st[3] == (len(seqs) + 96) // 97


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_exchange_between_gpus(golden_lists, world):
    """One thread per GPU (what the read_label binary does): shard r on device r, NCCL send/recv between the devices."""
    import torch
    if api.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    g = golden_lists
    t = api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    full = api.Db.upload(t)
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    ref_ctx = make_ctx(g, full, "run_rl")
    want, _, _ = ref_ctx.label(seqs)
    cut = [len(seqs) * i // world for i in range(world + 1)]
    cut[1] = len(seqs) // 2 if world > 2 else cut[1]           # uneven: the others keep serving
    cut = sorted(cut)
    uid = api.Comm.unique_id()
    shards = [api.Db.upload(t, r, r, world) for r in range(world)]
    ctxs = [make_ctx(g, shards[r], "run_rl") for r in range(world)]

    def rank(r):
        comm = api.Comm(r, r, world, uid)
        return _label_through_comm(ctxs[r], comm, f"cuda:{r}", seqs[cut[r]:cut[r + 1]], round_reads=101)

    outs = run_ranks(world, rank)
    res = np.concatenate([o[0] for o in outs])
    for f in ("status", "tid", "match", "valid_kmers", "cand_kmer_cnt", "n_cand"):
        assert np.array_equal(res[f], want[f]), f
    assert np.array_equal(res["score"].view(np.uint32), want["score"].view(np.uint32))
    assert sum(o[1][0] for o in outs) == sum(o[1][1] for o in outs) > 0


def _device_arrays(g, dev="cuda:0"):
    """The golden table as the device arrays kmat_db_build_device takes: k-mers, payloads (stored id, or LIST | pool word offset)
    and ONE list pool of [u16 count][u16 id]* records -- the form the bench's generator hands to every shard."""
    import torch
    assert g.tid_bytes == 2
    offs = g.offs.astype(np.int64)
    cnt = np.diff(offs)
    n = len(g.kmers)
    payload = np.zeros(n, dtype=np.int64)
    single = cnt == 1
    payload[single] = g.ids[offs[:-1][single]]
    pool, at = [], 0
    for i in np.nonzero(~single)[0]:
        c = int(cnt[i])
        rec = np.zeros(((1 + c) + 1) // 2 * 2, dtype=np.uint16)
        rec[0] = c
        rec[1:1 + c] = g.ids[offs[i]:offs[i] + c]
        payload[i] = (1 << 31) | at
        pool.append(rec)
        at += len(rec) // 2
    pool16 = np.concatenate(pool + [np.zeros(2, dtype=np.uint16)])
    pay32 = np.where(payload >= (1 << 31), payload - (1 << 32), payload).astype(np.int32)
    return (torch.as_tensor(g.kmers.astype(np.int64), device=dev), torch.as_tensor(pay32, device=dev),
            torch.as_tensor(pool16.view(np.int16), device=dev), at)


@pytest.mark.parametrize("world,opts", [(2, "run_rl"), (3, "permissive"), (3, "prune3"), (8, "run_rl")])
def test_direct_sharded_shared_pool_labels_equal_replicated(golden_lists, world, opts):
    """Shards built with kmat_db_build_device from the WHOLE table's device arrays (bench.py, C4) share one list pool and each
    resolves only the lists of its own k-mers; kmat_ctx_peer_attach merges the resolved pools.  (An earlier version aliased
    the local pool without the merge: list hits owned by another shard then read unresolved records.)"""
    import torch
    g = golden_lists
    kmers, pay32, pool16, pool_words = _device_arrays(g)
    assert pool_words > 0
    torch.cuda.synchronize()
    build = lambda r, w: api.Db.build_device(0, g.kmer_len, 2, len(g.kmers), kmers.data_ptr(), pay32.data_ptr(), pool16.data_ptr(), pool_words,
                                             shard_index=r, shard_count=w)
    full = build(0, 1)
    shards = [build(r, world) for r in range(world)]
    hdrs, seqs = op.read_fasta_like_reference(g.paths["reads"])
    ref_ctx = make_ctx(g, full, opts)
    res, cands, lin = ref_ctx.label(seqs)
    want = ref_ctx.tails(res, cands, lin, prn_all=True)
    # the device-built table answers like the uploaded one, which test_gpu_parity.py pins to the reference
    up_ctx = make_ctx(g, api.Db.upload(api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)), opts)
    r2, c2, l2 = up_ctx.label(seqs)
    assert up_ctx.tails(r2, c2, l2, prn_all=True) == want
    ctxs = [make_ctx(g, shards[r], opts) for r in range(world)]
    sharded.attach_peers_local(ctxs)
    for r in range(world):
        mine = seqs[r::world]
        rr, cc, ll = ctxs[r].label(mine)
        assert ctxs[r].tails(rr, cc, ll, prn_all=True) == want[r::world], r
    with pytest.raises(api.KmatError):                    # the merged pool cannot be rebuilt by one rank alone
        ctxs[0].set_opts(max_count=2)
