"""GPU-side synthetic workload generator for bench.py and the full-size property tests.

Bench/test infrastructure, not the product path: torch is used here only to fabricate inputs (random
genomes, their canonical 20-mers, the k-mer -> taxid-list table content, Illumina-like reads) at the
sizes BASELINE.json names (2,000 genomes x 500 kbp ~ 1e9 distinct 20-mers,
10 M x 150 bp reads) without a multi-minute host round trip.  The table content is handed to the library
through the C ABI (kmat_db_build_device); the reads go through kmat_label_batch like any user's reads.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import api
from . import fixtures as fx

LEVELS_C2 = ("kingdom", "phylum", "order", "family", "genus", "species", "strain")   # leaf depth 7


class Workload:
    pass


def make_taxonomy_c2(seed, n_genomes):
    tax = fx.make_taxonomy(seed, n_genomes, levels=LEVELS_C2)
    m16 = fx.map16(tax)
    depth = len(LEVELS_C2)
    anc_tid = np.zeros((n_genomes, depth + 1), dtype=np.int64)     # [g, d] = taxid of the ancestor at depth d (d = depth: itself)
    for gi, tid in enumerate(tax.leaves):
        path = [tid] + tax.path_to_root(tid)                      # leaf ... root
        assert len(path) == depth + 1
        anc_tid[gi] = path[::-1]
    anc_sid = np.vectorize(m16.get)(anc_tid).astype(np.int64)
    return tax, m16, anc_tid, anc_sid


def make_genomes_gpu(seed, tax, n_genomes, genome_len, device, share_frac=0.10, mut=0.02, gc_range=(0.3, 0.7)):
    """uint8 codes [G, L] on the device; 10 % of each genome copied with 2 % substitutions from the previous
    genome under the same parent."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    G, L = n_genomes, genome_len
    codes = torch.empty((G, L), dtype=torch.uint8, device=device)
    gc = torch.rand(G, generator=g, device=device) * (gc_range[1] - gc_range[0]) + gc_range[0]
    step = max(1, (1 << 28) // L)
    for a in range(0, G, step):
        b = min(G, a + step)
        u = torch.rand((b - a, L), generator=g, device=device)
        t0 = ((1 - gc[a:b]) / 2)[:, None]
        t2 = (0.5 + gc[a:b] / 2)[:, None]
        codes[a:b] = ((u > t0).to(torch.uint8) + (u > 0.5).to(torch.uint8) + (u > t2).to(torch.uint8))
    n = int(L * share_frac)
    if n > 0:
        rng = fx.rng_for(seed + 1)
        last = {}
        for gi, tid in enumerate(tax.leaves[:G]):
            par = tax.parent[tid]
            if par in last:
                s0, d0 = int(rng.integers(0, L - n + 1)), int(rng.integers(0, L - n + 1))
                seg = codes[last[par], s0:s0 + n].clone()
                flip = torch.rand(n, generator=g, device=device) < mut
                add = torch.randint(1, 4, (n,), generator=g, device=device, dtype=torch.uint8)
                seg = torch.where(flip, (seg + add) % 4, seg)
                codes[gi, d0:d0 + n] = seg
            last[par] = gi
    return codes


def canonical_kmers_gpu(codes, k):
    """int64 canonical k-mers [G, L-k+1] (same packing / min rule as read_label.cpp:992-1009)."""
    G, L = codes.shape
    n = L - k + 1
    fwd = torch.zeros((G, n), dtype=torch.int64, device=codes.device)
    rev = torch.zeros((G, n), dtype=torch.int64, device=codes.device)
    for j in range(k):
        c = codes[:, j:j + n].to(torch.int64)
        fwd = (fwd << 2) | c
        rev |= (3 - c) << (2 * j)
        del c
    return torch.minimum(fwd, rev)


def build_table_gpu(codes, anc_sid, k=20, chunk_genomes=None):
    """Logical table content on the device: distinct canonical k-mers (ascending), payload words and the list
    pool in libkmat's layout (kmat_db_build_device contract).  The taxid list of a k-mer shared by several
    genomes is the set tax_histo stores: every genome's taxid plus all nodes up to and including their LCA
    (TaxTree.hpp:159-260); list order here: deepest nodes first, then by stored id."""
    dev = codes.device
    G, L = codes.shape
    D = anc_sid.shape[1] - 1
    gbits = max(1, int(np.ceil(np.log2(max(G, 2)))))
    assert 2 * k + gbits <= 63
    keys = []
    step = chunk_genomes or max(1, (1 << 27) // L)
    for a in range(0, G, step):
        b = min(G, a + step)
        km = canonical_kmers_gpu(codes[a:b], k)
        gid = torch.arange(a, b, device=dev, dtype=torch.int64)[:, None]
        keys.append(((km << gbits) | gid).reshape(-1))
        del km
    key = torch.cat(keys)
    del keys
    anc = torch.as_tensor(anc_sid, device=dev)
    LIMIT = (1 << 31) - 1024                        # torch.sort / unique_consecutive take at most INT_MAX elements
    if key.numel() <= LIMIT:
        return _table_from_keys(key, anc, D, gbits, dev)
    # larger tables (C4: 2.6 G (k-mer, genome) pairs): split by k-mer range -- canonical k-mers of random genomes are close to
    # uniform below 2^(2k-1) -- build every part on its own and concatenate (k-mers stay ascending; list offsets are rebased)
    n_parts = int(key.numel() // (LIMIT // 2)) + 1
    top = 1 << (2 * k + gbits)
    parts, pool_base = [], 0
    for pi in range(n_parts):
        lo, hi = top * pi // (2 * n_parts), (top * (pi + 1) // (2 * n_parts) if pi + 1 < n_parts else top)
        sub = key[(key >= lo) & (key < hi)]
        w = _table_from_keys(sub, anc, D, gbits, dev)
        is_list = w.payload_i64 >= (1 << 31)
        w.payload_i64 = torch.where(is_list, w.payload_i64 + pool_base, w.payload_i64)
        pool_base += w.pool_words
        parts.append(w)
    del key
    out = Workload()
    out.kmers = torch.cat([w.kmers for w in parts])
    out.payload_i64 = torch.cat([w.payload_i64 for w in parts])
    out.pool16 = torch.cat([w.pool16[:2 * w.pool_words] for w in parts] + [torch.zeros(2, dtype=torch.int16, device=dev)])
    out.pool_words = pool_base
    assert pool_base < (1 << 31)
    out.n = int(out.kmers.numel())
    out.counts = None
    out.single = torch.cat([w.single for w in parts])
    out.lists = None                               # table_logical / table_to_host are for the small CPU samples only
    return out


def _table_from_keys(key, anc, D, gbits, dev):
    """The table content of a set of (k-mer << gbits | genome) keys (see build_table_gpu)."""
    key = torch.sort(key).values
    key = torch.unique_consecutive(key)
    kmer_all = key >> gbits
    gid_all = key & ((1 << gbits) - 1)
    del key
    kmers, counts = torch.unique_consecutive(kmer_all, return_counts=True)
    n = kmers.numel()
    starts = torch.cumsum(counts, 0) - counts
    payload = torch.zeros(n, dtype=torch.int64, device=dev)
    single = counts == 1
    payload[single] = anc[gid_all[starts[single]], D]
    multi_idx = torch.nonzero(~single).squeeze(1)
    S = multi_idx.numel()
    list_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    pool16 = torch.zeros(2, dtype=torch.int16, device=dev)
    pool_words = 0
    lists = None
    if S > 0:
        # members of the multi-genome k-mers
        mc = counts[multi_idx]
        seg = torch.repeat_interleave(torch.arange(S, device=dev), mc)
        mstart = starts[multi_idx]
        within = torch.arange(seg.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(mc, 0) - mc, mc)
        mg = gid_all[mstart[seg] + within]
        A = anc[mg]                                                     # [M, D+1] stored ids by depth
        lo = torch.full((S, D + 1), 1 << 40, dtype=torch.int64, device=dev).scatter_reduce(0, seg[:, None].expand(-1, D + 1), A, "amin")
        hi = torch.full((S, D + 1), -1, dtype=torch.int64, device=dev).scatter_reduce(0, seg[:, None].expand(-1, D + 1), A, "amax")
        agree = (lo == hi).to(torch.int64)
        lca_depth = torch.cumprod(agree, 1).sum(1) - 1                   # deepest level on which all members agree
        # (segment, depth, sid) triples: every member's nodes strictly below the LCA, plus the LCA itself
        dgrid = torch.arange(D + 1, device=dev)[None, :]
        take = dgrid > lca_depth[seg][:, None]
        segs = seg[:, None].expand(-1, D + 1)[take]
        deps = dgrid.expand(seg.numel(), -1)[take]
        sids = A[take]
        lca_sid = lo[torch.arange(S, device=dev), lca_depth]
        segs = torch.cat([segs, torch.arange(S, device=dev)])
        deps = torch.cat([deps, lca_depth])
        sids = torch.cat([sids, lca_sid])
        tkey = torch.unique((segs << 24) | ((D - deps) << 16) | sids)     # sorted: segment, deepest first, then sid
        lseg = tkey >> 24
        lsid = tkey & 0xFFFF
        ln = torch.bincount(lseg, minlength=S)
        lstart = torch.cumsum(ln, 0) - ln
        # pool placement: record of w words; records of <= 8 words are rounded up to a power of two so that they
        # never straddle a 32-byte sector, longer ones to a multiple of 8; larger classes first keeps alignment
        words = (2 + 2 * ln + 3) // 4
        alloc = torch.where(words <= 1, torch.ones_like(words), torch.where(words <= 2, torch.full_like(words, 2),
                torch.where(words <= 4, torch.full_like(words, 4), ((words + 7) // 8) * 8)))
        order = torch.argsort(alloc, descending=True, stable=True)
        off_sorted = torch.cumsum(alloc[order], 0) - alloc[order]
        off = torch.empty_like(off_sorted)
        off[order] = off_sorted
        pool_words = int(alloc.sum().item())
        assert pool_words < (1 << 31)
        pool16 = torch.zeros(2 * pool_words + 2, dtype=torch.int16, device=dev)
        pool16[2 * off] = ln.to(torch.int16)
        rank = torch.arange(lsid.numel(), device=dev) - lstart[lseg]
        pool16[2 * off[lseg] + 1 + rank] = lsid.to(torch.int16)
        payload[multi_idx] = (1 << 31) | off
        lists = dict(multi_idx=multi_idx, ln=ln, lstart=lstart, lsid=lsid)
    out = Workload()
    out.kmers = kmers
    out.payload_i64 = payload
    out.pool16 = pool16
    out.pool_words = pool_words
    out.n = n
    out.counts = counts
    out.lists = lists
    out.single = single
    return out


def upload_table(tbl, device_index=0, k=20, shard_index=0, shard_count=1):
    """kmat_db_build_device on the arrays of build_table_gpu (optionally one shard of the table only)."""
    pay = (tbl.payload_i64 & 0xFFFFFFFF).to(torch.int64)
    pay32 = torch.empty(tbl.n, dtype=torch.int32, device=tbl.kmers.device)
    pay32.copy_(torch.where(pay >= (1 << 31), pay - (1 << 32), pay).to(torch.int32))
    kmers = tbl.kmers.contiguous()
    # everything the library does not read goes back to the driver before the table is allocated (a 1.7 G k-mer table takes
    # 69-137 GB; torch's caching allocator would otherwise sit on the generator's temporaries)
    n_lists = int((~tbl.single).sum().item())
    tbl.n_lists = n_lists
    tbl.payload_i64 = None
    tbl.counts = None
    tbl.single = None
    tbl.lists = None
    del pay
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    db = api.Db.build_device(device_index, k, 2, tbl.n, kmers.data_ptr(), pay32.data_ptr(), tbl.pool16.data_ptr(), tbl.pool_words,
                             shard_index=shard_index, shard_count=shard_count)
    torch.cuda.synchronize()
    return db


def table_to_host(tbl, sid_to_tid):
    """(kmers uint64, offs uint64, tids uint32) numpy arrays of the logical table, taxids (not stored ids)."""
    kmers = tbl.kmers.cpu().numpy().astype(np.uint64)
    n = tbl.n
    lens = torch.ones(n, dtype=torch.int64, device=tbl.kmers.device)
    if tbl.lists is not None:
        lens[tbl.lists["multi_idx"]] = tbl.lists["ln"]
    offs = torch.zeros(n + 1, dtype=torch.int64, device=tbl.kmers.device)
    offs[1:] = torch.cumsum(lens, 0)
    ids = torch.zeros(int(offs[-1].item()), dtype=torch.int64, device=tbl.kmers.device)
    ids[offs[:-1][tbl.single]] = tbl.payload_i64[tbl.single]
    if tbl.lists is not None:
        li = tbl.lists
        seg_of = torch.repeat_interleave(torch.arange(li["ln"].numel(), device=ids.device), li["ln"])
        rank = torch.arange(li["lsid"].numel(), device=ids.device) - li["lstart"][seg_of]
        ids[offs[:-1][li["multi_idx"]][seg_of] + rank] = li["lsid"]
    sid2tid = np.asarray(sid_to_tid, dtype=np.uint32)
    return kmers, offs.cpu().numpy().astype(np.uint64), sid2tid[ids.cpu().numpy()], ids.cpu().numpy().astype(np.uint32)


def table_logical(tbl, sid_to_tid):
    """The logical table as torch tensors on the table's device: (kmers int64 ascending, offs int64 [n + 1], tids int64 in
    list order, taxids not stored ids).  Same content as table_to_host without the host copies."""
    dev = tbl.kmers.device
    n = tbl.n
    lens = torch.ones(n, dtype=torch.int64, device=dev)
    if tbl.lists is not None:
        lens[tbl.lists["multi_idx"]] = tbl.lists["ln"]
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    ids = torch.zeros(int(offs[-1].item()), dtype=torch.int64, device=dev)
    ids[offs[:-1][tbl.single]] = tbl.payload_i64[tbl.single]
    if tbl.lists is not None:
        li = tbl.lists
        seg_of = torch.repeat_interleave(torch.arange(li["ln"].numel(), device=dev), li["ln"])
        rank = torch.arange(li["lsid"].numel(), device=dev) - li["lstart"][seg_of]
        ids[offs[:-1][li["multi_idx"]][seg_of] + rank] = li["lsid"]
    s2t = torch.as_tensor(np.asarray(sid_to_tid, dtype=np.int64), device=dev)
    return tbl.kmers, offs, s2t[ids]


def write_tax_histo_torch(path, k, kmers, offs, tids, chunk=1 << 23):
    """tax_histo binary (the format fixtures.write_tax_histo documents) from torch tensors on any device, assembled in
    chunks of `chunk` records on that device: [kmer u64][count u16][count x tid u32], eight 0xFF bytes after every 1500th
    record (the reader's sanity marker)."""
    import struct
    dev = kmers.device
    n = int(kmers.numel())
    ar8 = torch.arange(8, device=dev)
    with open(path, "wb") as f:
        f.write(struct.pack("<IQQIcI", 29, n, 0xFFFFFFFFFFFFFFFF, 999, b"N", k))
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            m = b - a
            o = offs[a:b + 1] - offs[a]
            lens = o[1:] - o[:-1]
            idx = torch.arange(a, b, device=dev)
            sanity = ((idx + 1) % 1500 == 0).to(torch.int64) * 8
            sizes = 10 + 4 * lens + sanity
            pos = torch.cumsum(sizes, 0) - sizes
            buf = torch.zeros(int(sizes.sum().item()), dtype=torch.uint8, device=dev)
            kb = kmers[a:b].contiguous().view(torch.uint8).reshape(m, 8)                   # little endian host and device
            buf[(pos[:, None] + ar8[None, :]).reshape(-1)] = kb.reshape(-1)
            cb = lens.to(torch.int16).contiguous().view(torch.uint8).reshape(m, 2)
            buf[(pos[:, None] + 8 + ar8[None, :2]).reshape(-1)] = cb.reshape(-1)
            t = tids[int(offs[a].item()):int(offs[b].item())]
            rec = torch.repeat_interleave(torch.arange(m, device=dev), lens)
            rank = torch.arange(t.numel(), device=dev) - o[:-1][rec]
            tb = t.to(torch.int32).contiguous().view(torch.uint8).reshape(-1, 4)
            buf[((pos[rec] + 10 + 4 * rank)[:, None] + ar8[None, :4]).reshape(-1)] = tb.reshape(-1)
            sp = torch.nonzero(sanity).squeeze(1)
            if sp.numel():
                spos = pos[sp] + 10 + 4 * lens[sp]
                buf[(spos[:, None] + ar8[None, :]).reshape(-1)] = 0xFF
            f.write(buf.cpu().numpy().tobytes())
            del buf


def make_reads_gpu(seed, codes, n_reads, read_len=150, novel_frac=0.10, err_lo=0.001, err_hi=0.02, n_rate=0.0005,
                   chunk=1 << 20, genomes=None):
    """ASCII reads [n_reads, read_len] uint8 on the device (SURVEY.md 8(d) read model: 90 % from genomes, 10 %
    random, substitution error rising linearly 0.1 % -> 2 %, 0.05 % N, 50 % reverse strand)."""
    dev = codes.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    G, L = codes.shape
    flat = codes.reshape(-1)
    out = torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    ascii_lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    perr = torch.linspace(err_lo, err_hi, read_len, device=dev)[None, :]
    ar = torch.arange(read_len, device=dev)[None, :]
    for a in range(0, n_reads, chunk):
        b = min(n_reads, a + chunk)
        m = b - a
        if genomes is None:
            gi = torch.randint(0, G, (m,), generator=g, device=dev)
        else:
            gi = torch.as_tensor(genomes, device=dev)[torch.randint(0, len(genomes), (m,), generator=g, device=dev)]
        pos = torch.randint(0, L - read_len + 1, (m,), generator=g, device=dev)
        c = flat[(gi * L + pos)[:, None] + ar]
        novel = torch.rand(m, generator=g, device=dev) < novel_frac
        rnd = torch.randint(0, 4, (m, read_len), generator=g, device=dev, dtype=torch.uint8)
        c = torch.where(novel[:, None], rnd, c)
        flip = torch.rand((m, read_len), generator=g, device=dev) < perr
        add = torch.randint(1, 4, (m, read_len), generator=g, device=dev, dtype=torch.uint8)
        c = torch.where(flip, (c + add) % 4, c)
        rc = torch.rand(m, generator=g, device=dev) < 0.5
        c = torch.where(rc[:, None], 3 - torch.flip(c, dims=[1]), c)
        s = ascii_lut[c.long()]
        isn = torch.rand((m, read_len), generator=g, device=dev) < n_rate
        s = torch.where(isn, torch.full_like(s, 78), s)
        out[a:b] = s
    return out


def write_tax_histo_fast(path, k, kmers, offs, tids):
    """Vectorised writer of the tax_histo binary (same format as fixtures.write_tax_histo)."""
    import struct
    n = len(kmers)
    offs = offs.astype(np.int64)
    lens = np.diff(offs)
    idx = np.arange(n, dtype=np.int64)
    sanity = ((idx + 1) % 1500 == 0).astype(np.int64) * 8
    sizes = 10 + 4 * lens + sanity
    pos = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    buf = np.zeros(int(sizes.sum()), dtype=np.uint8)
    kb = np.ascontiguousarray(kmers.astype("<u8")).view(np.uint8).reshape(n, 8)
    buf[(pos[:, None] + np.arange(8)[None, :]).reshape(-1)] = kb.reshape(-1)
    cb = np.ascontiguousarray(lens.astype("<u2")).view(np.uint8).reshape(n, 2)
    buf[(pos[:, None] + 8 + np.arange(2)[None, :]).reshape(-1)] = cb.reshape(-1)
    rec = np.repeat(idx, lens)
    rank = np.arange(len(tids), dtype=np.int64) - offs[:-1][rec]
    tb = np.ascontiguousarray(tids.astype("<u4")).view(np.uint8).reshape(len(tids), 4)
    buf[((pos[rec] + 10 + 4 * rank)[:, None] + np.arange(4)[None, :]).reshape(-1)] = tb.reshape(-1)
    sp = np.nonzero(sanity)[0]
    if len(sp):
        spos = pos[sp] + 10 + 4 * lens[sp]
        buf[(spos[:, None] + np.arange(8)[None, :]).reshape(-1)] = 0xFF
    with open(path, "wb") as f:
        f.write(struct.pack("<IQQIcI", 29, n, 0xFFFFFFFFFFFFFFFF, 999, b"N", k))
        f.write(buf.tobytes())


def write_null_models_for(tax, outdir, seed=60240):
    """10 bins x model lengths {31,56,81,106,131}, cut-offs U[0.005,0.08], seed 60240."""
    return fx.write_null_models(seed, tax, outdir, holes=False)
