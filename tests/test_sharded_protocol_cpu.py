"""CPU: the DB-sharded round protocol of lmat_b200.sharded over torch.distributed (gloo, world_size 2 and 3).

The device phases are replaced by a numpy stand-in built on the oracle (test infrastructure): encode = the oracle's
first-occurrence canonical k-mers grouped by kmat_shard_of, serve = dictionary lookup in the rank's shard of the golden
table, finish = taxid lists back at their (read, position).  What is under test is the driver and the exchange: counts,
split sizes, ordering, payload rebasing, ranks with fewer rounds than others.  The CUDA phases are covered on the GPU
by tests/test_gpu_sharded.py through the same driver."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MISS = -2


class NumpyPhases:
    def __init__(self, rank, world, kmers, offs, ids, k):
        import ctypes as C
        from lmat_b200 import api
        self.rank, self.world, self.k = rank, world, k
        L = api.lib()                                          # the C-ABI library loads without a GPU; kmat_shard_of is host code
        self.owner = lambda km: L.kmat_shard_of(C.c_uint64(int(km)), k, world)
        self.shard = {int(km): ids[int(offs[i]):int(offs[i + 1])] for i, km in enumerate(kmers) if self.owner(km) == rank}
        self.origin = None

    def empty_round(self):
        return ([], None)

    def encode(self, reads):
        import torch
        from oracle import oracle_py as op
        per_owner = [[] for _ in range(self.world)]
        for ri, s in enumerate(reads):
            v, b, km, fl = op.encode_read(s, self.k)
            for p in np.nonzero(fl == 1)[0]:
                per_owner[self.owner(km[p])].append((int(km[p]), ri, int(p)))
        self.origin = [(ri, p) for lst in per_owner for (_, ri, p) in lst]
        q = np.array([x for lst in per_owner for (x, _, _) in lst], dtype=np.int64)
        self.counts = np.array([len(lst) for lst in per_owner], dtype=np.uint64)
        return torch.from_numpy(q), self.counts

    def serve(self, queries, counts):
        import torch
        q = queries.numpy()
        reply = np.full(len(q), MISS, dtype=np.int32)
        payload, pay_counts, i = [], np.zeros(self.world, dtype=np.uint64), 0
        for s in range(self.world):
            seg = []
            for _ in range(int(counts[s])):
                lst = self.shard.get(int(q[i]))
                if lst is not None:
                    reply[i] = len(seg)                        # offset inside the source's payload segment
                    seg += [len(lst)] + [int(t) for t in lst]
                i += 1
            pay_counts[s] = len(seg)
            payload += seg
        return torch.from_numpy(reply), torch.from_numpy(np.array(payload, dtype=np.int32)), pay_counts

    def finish(self, reply, payload, pay_counts, out):
        reply, payload = reply.numpy(), payload.numpy()
        base = np.concatenate([[0], np.cumsum(pay_counts.astype(np.int64))])
        qstart = np.concatenate([[0], np.cumsum(self.counts.astype(np.int64))])
        for o in range(self.world):
            for i in range(int(qstart[o]), int(qstart[o + 1])):
                if reply[i] == MISS:
                    continue
                at = int(base[o]) + int(reply[i])
                n = int(payload[at])
                out[self.origin[i]] = [int(t) for t in payload[at + 1:at + 1 + n]]


def _worker(rank, world, port, q, pipelined=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from lmat_b200 import sharded
    from oracle import oracle_py as op
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = np.load(os.path.join(ROOT, "tests", "golden", "small.table.npz"))
        kmers, offs, ids, k = t["kmers"], t["offs"], t["ids"].astype(np.uint32), int(t["kmer_len"])
        rng = np.random.default_rng(100 + rank)
        # reads cut out of the table's own k-mers (hits) plus noise; rank r gets 3 + 2 r rounds of 5 reads
        def read():
            parts = []
            for _ in range(4):
                km = int(kmers[rng.integers(0, len(kmers))])
                parts.append("".join("ACGT"[(km >> (2 * (k - 1 - j))) & 3] for j in range(k)))
            return "".join(parts) + "N" + "".join("ACGT"[x] for x in rng.integers(0, 4, 30))
        reads = [read() for _ in range(5 * (3 + 2 * rank))]
        ph = NumpyPhases(rank, world, kmers, offs, ids, k)
        ph2 = NumpyPhases(rank, world, kmers, offs, ids, k) if pipelined else None
        lab = sharded.ShardedLabeler(ph, sharded.DistExchange("cpu"), round_reads=5, phases2=ph2)
        got = {}
        out_rounds = [{}, {}]
        turn = [0]
        lens = np.array([len(r) for r in reads])
        rounds = lab.plan(np.concatenate([[0], np.cumsum(lens)]))

        def round_args(r0, r1):
            o = out_rounds[turn[0] % 2]                # the pipelined driver asks for round i+1 before round i has finished
            turn[0] += 1
            o.clear()
            return (reads[r0:r1], o)

        def on_round(r0, r1, slot=None):
            o = out_rounds[slot] if slot is not None else out_rounds[(turn[0] - 1) % 2]
            for (ri, p), lst in o.items():
                got[(r0 + ri, p)] = lst
        lab.run(rounds, round_args, on_round)
        full = {int(km): [int(x) for x in ids[int(offs[i]):int(offs[i + 1])]] for i, km in enumerate(kmers)}
        want = {}
        for ri, s in enumerate(reads):
            v, b, km, fl = op.encode_read(s, k)
            for p in np.nonzero(fl == 1)[0]:
                if int(km[p]) in full:
                    want[(ri, int(p))] = full[int(km[p])]
        q.put((rank, got == want, len(want), lab.rounds, lab.lookups, lab.served))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,pipelined", [(2, False), (3, False), (2, True), (3, True)])
def test_sharded_rounds_over_gloo(world, pipelined):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world + (10 if pipelined else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, pipelined)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, *_ in res), res
    assert all(n > 0 for _, _, n, *_ in res)
    assert len({r[3] for r in res}) == 1 and res[0][3] == 3 + 2 * (world - 1)      # all ranks ran the longest rank's rounds
    assert sum(r[4] for r in res) == sum(r[5] for r in res)                        # every query sent was served
