"""DB-sharded read labeling (SURVEY.md 8(e) mode B): the k-mer table is partitioned over the ranks by kmat_shard_of,
reads stay on their home rank, query k-mers travel.

One round over (a chunk of) a rank's reads:

    home   ctx.shard_encode   encode + dedup, first-occurrence k-mers grouped by owner       [CUDA, libkmat]
    ------ all-to-all: per-owner counts, then the k-mers (8 B each) ------------------------ [exchange]
    owner  ctx.shard_serve    probe the local shard -> hit words + packed list records       [CUDA, libkmat]
    ------ all-to-all: per-source payload sizes, hit words (4 B each), list records --------- [exchange]
    home   ctx.shard_finish   hit words back to the read positions, candidate + scoring kernels

`ShardedLabeler` is the per-rank driver.  The exchange is pluggable: `DistExchange` = torch.distributed
all_to_all_single (NCCL over NVLink between one-process-per-GPU ranks; gloo in the CPU tests), `LocalExchange` =
several ranks as threads of one process (tests on a single GPU; a single-process multi-GPU host).  The device phases
are pluggable too (`CudaPhases` wraps the C ABI); the CPU tests drive the same protocol with a numpy stand-in.
"""
from __future__ import annotations

import threading

import numpy as np


# ------------------------------------------------------------------------------------------------
# exchanges
# ------------------------------------------------------------------------------------------------
class DistExchange:
    """torch.distributed backend: all_to_all_single over the default (or a given) process group."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def counts(self, send_counts):
        import torch
        s = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=self.device)
        r = torch.empty_like(s)
        self.dist.all_to_all_single(r, s, group=self.group)
        return r.cpu().numpy().astype(np.uint64)

    def all_to_all(self, send, send_counts, recv_counts):
        import torch
        recv = torch.empty(int(np.sum(recv_counts)), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_counts],
                                    input_split_sizes=[int(x) for x in send_counts], group=self.group)
        return recv

    def any(self, flag):
        import torch
        t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())

    def max_int(self, v):
        import torch
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(t.item())


class LocalGroup:
    """Shared state of `world` ranks running as threads of one process."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class LocalExchange:
    def __init__(self, group: LocalGroup, rank, sync=None):
        self.g, self.rank, self.world = group, rank, group.world
        self.sync = sync or (lambda: None)          # makes this rank's device work visible (torch.cuda.synchronize)

    def counts(self, send_counts):
        self.g.slots[self.rank] = np.asarray(send_counts, dtype=np.uint64).copy()
        self.g.barrier.wait()
        recv = np.array([self.g.slots[s][self.rank] for s in range(self.world)], dtype=np.uint64)
        self.g.barrier.wait()
        return recv

    def all_to_all(self, send, send_counts, recv_counts):
        import torch
        offs = np.concatenate([[0], np.cumsum(np.asarray(send_counts, dtype=np.int64))])
        self.sync()
        self.g.slots[self.rank] = (send, offs)
        self.g.barrier.wait()
        parts = []
        for s in range(self.world):
            t, o = self.g.slots[s]
            part = t[int(o[self.rank]):int(o[self.rank + 1])]
            assert part.numel() == int(recv_counts[s])
            parts.append(part.to(send.device))
        recv = torch.cat(parts) if parts else send[:0]
        self.sync()
        self.g.barrier.wait()
        return recv

    def any(self, flag):
        self.g.slots[self.rank] = bool(flag)
        self.g.barrier.wait()
        r = any(self.g.slots)
        self.g.barrier.wait()
        return r

    def max_int(self, v):
        self.g.slots[self.rank] = int(v)
        self.g.barrier.wait()
        r = max(self.g.slots)
        self.g.barrier.wait()
        return r


# ------------------------------------------------------------------------------------------------
# device phases over the C ABI
# ------------------------------------------------------------------------------------------------
class _DevMem:
    """Library-owned device memory as something torch.as_tensor() accepts."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


def _wrap(ptr, count, dtype, device):
    import torch
    if not count:
        return torch.empty(0, dtype=dtype, device=device)
    itemsize = torch.empty(0, dtype=dtype).element_size()
    return torch.as_tensor(_DevMem(ptr, count * itemsize), device=device).view(dtype)


class CudaPhases:
    """The three libkmat phases of a round for one rank (api.Ctx over this rank's table shard)."""

    def __init__(self, ctx, device, n_shards, stream=None, torch_stream=None):
        """stream: raw cudaStream_t handle the library launches on (None = the ctx's own).  torch_stream: a torch.cuda.Stream
        instead -- the phases AND this slot's exchanges then run on it (ShardedLabeler's two-slot pipeline)."""
        self.ctx, self.device, self.n_shards, self.stream = ctx, device, n_shards, stream
        self.tstream = torch_stream
        if torch_stream is not None:
            self.stream = torch_stream.cuda_stream

    def on_stream(self):
        import contextlib
        import torch
        return torch.cuda.stream(self.tstream) if self.tstream is not None else contextlib.nullcontext()

    def empty_round(self):
        """Arguments of a round without reads (this rank still serves the others' queries)."""
        import torch
        if not hasattr(self, "_zero"):
            self._zero = torch.zeros(2, dtype=torch.int64, device=self.device)
        return (None, self._zero.data_ptr(), 0, 0, 0, None)

    def encode(self, bases_ptr, offs_ptr, n_reads, total_bases, max_len):
        import torch
        q, counts = self.ctx.shard_encode(bases_ptr, offs_ptr, n_reads, total_bases, max_len, self.n_shards, self.stream)
        return _wrap(q, int(counts.sum()), torch.int64, self.device), counts

    def serve(self, queries, counts):
        import torch
        rep, pay, pc = self.ctx.shard_serve(queries.data_ptr() if queries.numel() else None, counts, self.stream)
        return _wrap(rep, int(np.sum(counts)), torch.int32, self.device), _wrap(pay, int(pc.sum()), torch.int32, self.device), pc

    def finish(self, reply, payload, payload_counts, out_ptr):
        self.ctx.shard_finish(reply.data_ptr() if reply.numel() else None, payload.data_ptr() if payload.numel() else None,
                              payload_counts, out_ptr, self.stream)


# ------------------------------------------------------------------------------------------------
# per-rank driver
# ------------------------------------------------------------------------------------------------
class ShardedLabeler:
    """Runs the rounds of one rank.  Every rank of the group must call run() the same number of times; inside, ranks
    keep exchanging (with empty contributions once their own reads are done) until every rank has finished."""

    def __init__(self, phases, exchange, round_reads=1 << 20, round_bases=(1 << 32) - (1 << 20), phases2=None):
        """phases2: a second set of phases over a second context of the same shard.  Rounds then alternate between the two
        slots and the encode phase of round i+1 is queued (on the other slot's stream) before round i's query exchange is
        waited for, its finish phase overlaps the next round's exchange: see _run_pipelined."""
        self.ph, self.ex, self.ph2 = phases, exchange, phases2
        self.round_reads, self.round_bases = int(round_reads), int(round_bases)
        self.lookups = 0            # first-occurrence k-mers this rank sent out (its unique lookups)
        self.served = 0             # queries this rank answered
        self.payload_words = 0
        self.rounds = 0
        self.timing = None          # set to {} to collect per-phase device time (ms, summed over rounds)
        self._marks = []

    def plan(self, offs_host):
        """Cut reads [0, n) into rounds bounded by round_reads and round_bases; offs_host = n+1 absolute offsets."""
        n = len(offs_host) - 1
        out, r0 = [], 0
        while r0 < n:
            r1 = min(n, r0 + self.round_reads)
            if int(offs_host[r1]) - int(offs_host[r0]) > self.round_bases:
                r1 = int(np.searchsorted(offs_host, int(offs_host[r0]) + self.round_bases, side="right")) - 1
                r1 = max(r1, r0 + 1)
            out.append((r0, r1))
            r0 = r1
        return out

    def run(self, rounds, round_args, on_round=None):
        """rounds: list of (r0, r1); round_args(r0, r1) -> (*encode_args, out): what the phases' encode() takes (CudaPhases:
        bases_ptr, offs_ptr, n_reads, total_bases, max_len with offsets local to the chunk) followed by what finish()
        gets as its output argument.  on_round(r0, r1) is called after each finished round of this rank."""
        if self.ph2 is not None:
            return self._run_pipelined(rounds, round_args, on_round)
        i = 0
        while True:
            mine = i < len(rounds)
            if not self.ex.any(mine):
                break
            if mine:
                r0, r1 = rounds[i]
                args = round_args(r0, r1)
            else:
                r0 = r1 = 0
                args = self.ph.empty_round()
            self._round(args)
            if mine and on_round:
                on_round(r0, r1)
            i += 1
            self.rounds += 1
        self._collect_timing()

    def _run_pipelined(self, rounds, round_args, on_round):
        """Two slots (contexts + streams), rounds alternate.  All ranks issue the same sequence of collectives: per round
        counts, queries | counts, hit words, list records -- the encode of the next round carries none."""
        import contextlib
        ph = [self.ph, self.ph2]
        n_rounds = self.ex.max_int(len(rounds))              # one collective instead of one per round
        cur = None
        if getattr(ph[0], "tstream", None) is not None:
            import torch
            cur = torch.cuda.current_stream()
            for p in ph:
                p.tstream.wait_stream(cur)

        def ctx(p):
            return p.on_stream() if hasattr(p, "on_stream") else contextlib.nullcontext()

        def encode(i):
            p = ph[i % 2]
            a = round_args(*rounds[i]) if i < len(rounds) else p.empty_round()
            with ctx(p):
                q, counts = p.encode(*a[:-1])
            return a, q, counts

        nxt = encode(0) if n_rounds else None
        for i in range(n_rounds):
            p = ph[i % 2]
            args, send_q, send_counts = nxt
            self.lookups += int(np.sum(send_counts))
            with ctx(p):
                recv_counts = self.ex.counts(send_counts)
                recv_q = self.ex.all_to_all(send_q, send_counts, recv_counts)
            if i + 1 < n_rounds:
                nxt = encode(i + 1)                          # on the other slot's stream: runs while the queries travel
            with ctx(p):
                reply, payload, pay_counts = p.serve(recv_q, recv_counts)
                self.served += int(np.sum(recv_counts))
                self.payload_words += int(np.sum(pay_counts))
                my_pay_counts = self.ex.counts(pay_counts)
                my_reply = self.ex.all_to_all(reply, recv_counts, send_counts)
                my_payload = self.ex.all_to_all(payload, pay_counts, my_pay_counts)
                p.finish(my_reply, my_payload, my_pay_counts, args[-1])      # asynchronous: overlaps the next round's exchange
            if i < len(rounds) and on_round:
                on_round(*rounds[i], i % 2)
            self.rounds += 1
        if cur is not None:
            for p in ph:
                cur.wait_stream(p.tstream)

    def _mark(self, name):
        """Phase timing (opt-in: self.timing = {} before run()): a CUDA event on the current stream after each phase."""
        if self.timing is None:
            return
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._marks.append((name, e))

    def _collect_timing(self):
        if self.timing is None or not self._marks:
            return
        import torch
        torch.cuda.synchronize()
        prev = None
        for name, e in self._marks:
            if prev is not None and name != "start":
                self.timing[name] = self.timing.get(name, 0.0) + prev.elapsed_time(e)
            prev = e
        self._marks = []

    def _round(self, args):
        import contextlib
        with (self.ph.on_stream() if hasattr(self.ph, "on_stream") else contextlib.nullcontext()):     # phases AND exchanges on the phases' stream
            self._round_on_stream(args)

    def _round_on_stream(self, args):
        out_ptr = args[-1]
        self._mark("start")
        send_q, send_counts = self.ph.encode(*args[:-1])
        self._mark("encode")
        self.lookups += int(np.sum(send_counts))
        recv_counts = self.ex.counts(send_counts)                          # how many queries each source sends me
        recv_q = self.ex.all_to_all(send_q, send_counts, recv_counts)
        self._mark("exchange_queries")
        reply, payload, pay_counts = self.ph.serve(recv_q, recv_counts)    # pay_counts[s]: list words for source s
        self._mark("serve")
        self.served += int(np.sum(recv_counts))
        self.payload_words += int(np.sum(pay_counts))
        my_pay_counts = self.ex.counts(pay_counts)                         # list words each owner sends me
        my_reply = self.ex.all_to_all(reply, recv_counts, send_counts)     # hit words, in the order of send_q
        my_payload = self.ex.all_to_all(payload, pay_counts, my_pay_counts)
        self._mark("exchange_replies")
        self.ph.finish(my_reply, my_payload, my_pay_counts, out_ptr)
        self._mark("finish")


def label_sequences(ctx, exchange, device, seqs, n_shards, round_reads=1 << 20, ctx2=None):
    """Convenience (tests, small inputs): label a list of reads through the sharded rounds of this rank and return
    (results, candidates) like api.Ctx.label -- numpy arrays, cand_off indexing the returned candidate array."""
    import torch
    from . import api
    lens = np.array([len(x) for x in seqs], dtype=np.int64)
    offs_host = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = b"".join(x if isinstance(x, bytes) else x.encode("latin-1") for x in seqs)
    bases = torch.frombuffer(bytearray(flat if flat else b"\0"), dtype=torch.uint8).to(device)
    if ctx2 is None:
        lab = ShardedLabeler(CudaPhases(ctx, device, n_shards), exchange, round_reads=round_reads)
    else:                                   # two-slot pipeline: each slot its own context and torch stream
        lab = ShardedLabeler(CudaPhases(ctx, device, n_shards, torch_stream=torch.cuda.Stream(device)), exchange, round_reads=round_reads,
                             phases2=CudaPhases(ctx2, device, n_shards, torch_stream=torch.cuda.Stream(device)))
    ctxs = [ctx, ctx2]
    res_parts, cand_parts, keep = [], [], []
    state = {"cands": 0}

    def round_args(r0, r1):
        o = torch.as_tensor(offs_host[r0:r1 + 1] - offs_host[r0], device=device)
        keep.append(o)
        return (bases.data_ptr() + int(offs_host[r0]), o.data_ptr(), r1 - r0, int(offs_host[r1] - offs_host[r0]),
                int(lens[r0:r1].max()) if r1 > r0 else 0, None)

    def on_round(r0, r1, slot=0):
        torch.cuda.synchronize(device)
        optr, cptr, n_c = ctxs[slot].device_results()
        res = _wrap(optr, (r1 - r0) * api.RESULT_DTYPE.itemsize, torch.uint8, device).cpu().numpy().view(api.RESULT_DTYPE).copy()
        cands = _wrap(cptr, n_c * api.PAIR_DTYPE.itemsize, torch.uint8, device).cpu().numpy().view(api.PAIR_DTYPE).copy()
        res["cand_off"] += np.uint64(state["cands"])
        state["cands"] += len(cands)
        res_parts.append(res)
        cand_parts.append(cands)

    lab.run(lab.plan(offs_host), round_args, on_round)
    res = np.concatenate(res_parts) if res_parts else np.zeros(0, dtype=api.RESULT_DTYPE)
    cands = np.concatenate(cand_parts) if cand_parts else np.zeros(0, dtype=api.PAIR_DTYPE)
    return res, cands, lab


def attach_peers(ctx, group=None, device=None):
    """Direct variant of the DB-sharded mode for one-process-per-GPU ranks: all-gather the shard descriptions
    (kmat_ctx_peer_export) over torch.distributed and map every peer's table into this rank (kmat_ctx_peer_attach).
    Afterwards ctx.label / ctx.label_device work against the whole table with no exchange rounds."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.from_numpy(ctx.peer_export())
    if device is not None:
        mine = mine.to(device)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine, group=group)
    ctx.peer_attach([b.cpu().numpy() for b in allb])
    dist.barrier(group=group)


def attach_peers_local(ctxs):
    """The same inside one process (several GPUs driven by one host, or virtual ranks on one GPU in the tests)."""
    blobs = [c.peer_export() for c in ctxs]
    for c in ctxs:
        c.peer_attach(blobs)
