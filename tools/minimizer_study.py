#!/usr/bin/env python
"""Design study (CPU, numpy): how many DRAM lines would a read touch if the k-mer table were ordered by minimizer?

Today every unique k-mer of a read costs one random 32-byte bucket gather (the probe kernel sits at the
request-rate ceiling, so only FEWER requests can make it faster).  This script builds a scaled-down synthetic DB with the
generator family of the bench (random genomes, sibling-shared segments), assigns every canonical k-mer to the line
hash(canonical minimizer) of a table of 128-byte lines (16 slots, linear probing to the next line when full) and replays
reads with the bench's error model, counting for each read the distinct lines its lookups touch (misses included: a miss
must read the home line, and every following line while they are full).

usage: minimizer_study.py [genomes=100] [genome_len=500000] [m=12] [reads=20000] [slots_per_line=16] [mean_fill=4.0]
"""
import sys
import time

import numpy as np

K = 20


def revcomp(x, k):
    y = np.zeros_like(x)
    t = (~x) & ((np.uint64(1) << np.uint64(2 * k)) - np.uint64(1))
    for i in range(k):
        y = (y << np.uint64(2)) | ((t >> np.uint64(2 * i)) & np.uint64(3))
    return y


def mix(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
        x = (x ^ (x >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
    return x ^ (x >> np.uint64(33))


def kmers_of(codes, k):
    """forward k-mers of a uint8 code array (no N here), as uint64"""
    n = len(codes) - k + 1
    out = np.zeros(n, dtype=np.uint64)
    c = codes.astype(np.uint64)
    for i in range(k):
        out = (out << np.uint64(2)) | c[i:i + n]
    return out


def canon_and_minimizer(codes, k, m):
    """per k-mer position: canonical k-mer, and the hash of its canonical minimizer (random order = smallest mix())"""
    fwd = kmers_of(codes, k)
    canon = np.minimum(fwd, revcomp(fwd, k))
    mf = kmers_of(codes, m)
    mh = mix(np.minimum(mf, revcomp(mf, m)))               # one value per m-mer position
    w = k - m + 1
    n = len(fwd)
    best = mh[:n].copy()
    for j in range(1, w):
        best = np.minimum(best, mh[j:j + n])
    return canon, best


def main():
    a = sys.argv[1:]
    G = int(a[0]) if len(a) > 0 else 100
    GL = int(a[1]) if len(a) > 1 else 500000
    m = int(a[2]) if len(a) > 2 else 12
    R = int(a[3]) if len(a) > 3 else 20000
    S = int(a[4]) if len(a) > 4 else 16
    fill = float(a[5]) if len(a) > 5 else 4.0
    rng = np.random.default_rng(20240)
    t0 = time.time()
    genomes = []
    for g in range(G):
        gc = rng.uniform(0.3, 0.7)
        p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        x = rng.choice(4, size=GL, p=p).astype(np.uint8)
        if g % 4 and genomes:                              # 10 % copied from a sibling with 2 % mutations
            n = GL // 10
            s0, d0 = int(rng.integers(0, GL - n)), int(rng.integers(0, GL - n))
            seg = genomes[g - 1][s0:s0 + n].copy()
            f = rng.random(n) < 0.02
            seg[f] = (seg[f] + rng.integers(1, 4, size=int(f.sum()))) % 4
            x[d0:d0 + n] = seg
        genomes.append(x)
    cs, ms = [], []
    for x in genomes:
        c, b = canon_and_minimizer(x, K, m)
        cs.append(c)
        ms.append(b)
    canon = np.concatenate(cs)
    minim = np.concatenate(ms)
    canon, idx = np.unique(canon, return_index=True)
    minim = minim[idx]
    n_kmers = len(canon)
    n_lines = 1 << int(np.ceil(np.log2(n_kmers / fill)))
    home = (minim >> np.uint64(20)) & np.uint64(n_lines - 1)
    occ = np.bincount(home.astype(np.int64), minlength=n_lines)
    # linear probing of whole lines: overflow of a line spills into the following ones (simulated on the occupancy array)
    load = occ.astype(np.int64)
    carry = 0
    full_run = np.zeros(n_lines, dtype=np.int64)           # lines a probe starting here must read beyond the first
    eff = np.zeros(n_lines, dtype=np.int64)
    for rnd in range(2):                                   # two sweeps handle the wrap-around
        for i in range(n_lines):
            tot = load[i] + carry if rnd == 0 or i == 0 or carry else load[i]
            if rnd == 1 and not carry:
                break
            eff[i] = min(S, tot)
            carry = max(0, tot - S)
            load[i] = eff[i] if rnd == 1 else load[i]
    # extra lines read by a probe with home line i: the run of full lines starting at i
    isfull = eff >= S
    run = np.zeros(n_lines + 1, dtype=np.int64)
    for i in range(n_lines - 1, -1, -1):
        run[i] = run[i + 1] + 1 if isfull[i] else 0
    print(f"DB: {G} genomes x {GL} bp -> {n_kmers} canonical {K}-mers; m = {m} (window {K - m + 1}), {n_lines} lines of {S} slots, mean fill {n_kmers / n_lines:.2f}; "
          f"{(occ > S).mean() * 100:.2f} % of the lines overflow, {isfull.mean() * 100:.2f} % end up full; built in {time.time() - t0:.0f} s")
    kmers_per_min = np.bincount(np.unique(minim, return_inverse=True)[1])
    print(f"k-mers per distinct minimizer: mean {kmers_per_min.mean():.1f}, p99 {np.percentile(kmers_per_min, 99):.0f}, max {kmers_per_min.max()}")
    # reads: 90 % genomic with 0.1 % -> 2 % substitutions along the read, 10 % random
    now_req, new_req, supers = [], [], []
    for r in range(R):
        if rng.random() < 0.1:
            x = rng.integers(0, 4, size=150).astype(np.uint8)
        else:
            g = genomes[int(rng.integers(0, G))]
            s0 = int(rng.integers(0, GL - 150))
            x = g[s0:s0 + 150].copy()
            f = rng.random(150) < np.linspace(0.001, 0.02, 150)
            x[f] = (x[f] + rng.integers(1, 4, size=int(f.sum()))) % 4
        c, b = canon_and_minimizer(x, K, m)
        _, first = np.unique(c, return_index=True)
        c, b = c[np.sort(first)], b[np.sort(first)]
        now_req.append(len(c))
        h = ((b >> np.uint64(20)) & np.uint64(n_lines - 1)).astype(np.int64)
        lines = set()
        for hh in np.unique(h):
            for d in range(int(run[hh]) + 1):
                lines.add((hh + d) % n_lines)
        new_req.append(len(lines))
        supers.append(1 + int((b[1:] != b[:-1]).sum()))
    now_req, new_req, supers = np.array(now_req), np.array(new_req), np.array(supers)
    print(f"{R} reads of 150 bp: unique k-mer lookups (= gathers today) {now_req.mean():.1f} per read; super-k-mers {supers.mean():.1f}; "
          f"distinct lines with the minimizer layout {new_req.mean():.1f} per read (p95 {np.percentile(new_req, 95):.0f}) -> {now_req.mean() / new_req.mean():.2f}x fewer requests")


if __name__ == "__main__":
    main()
