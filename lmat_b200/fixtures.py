"""Synthetic inputs for the read_label path: taxonomy, genomes, k-mer table input, reads, null models.

Test/bench infrastructure (numpy only).  Every generator is seeded with
``numpy.random.Generator(PCG64(seed))`` as SURVEY.md section 8(d) specifies.  The file formats
written here are the ones the reference parses:

* taxonomy tree  -- TaxTree.hpp:24-57 / TaxNode.hpp:131-147 (two '#' lines, a count line, then per
  node ``id nchild child... parent`` / ``name``; NO trailing newline)
* depth file     -- read_label.cpp:1573-1582   (``tid depth``)
* rank file      -- read_label.cpp:1560-1567   (``tid rank``)
* 32->16 map     -- read_label.cpp:1585-1602   (``tid32 tid16``)
* numeric ranks  -- read_label.cpp:1543-1559   (``tid rank_num``)
* null models    -- read_label.cpp:512-678     (list file + one model file per k-mer count)
* tax_histo file -- KmerFileMetaData.cpp:16-34 + tax_histo.cpp:257-281 (input of make_db_table)
* genome FASTA for kmerPrefixCounter -- kmerPrefixCounter.cpp:116-130 (``>tid`` / one-line sequence)
"""
from __future__ import annotations

import gzip
import os
import struct
from dataclasses import dataclass, field

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
HUMAN_TIDS = (9606, 63221, 741158)
PHIX_TIDS = (374840, 10847, 32630)
NUMERIC_RANK = {"species": 15, "genus": 14, "family": 12, "order": 10, "class": 8, "phylum": 6,
                "kingdom": 4, "superkingdom": 2}


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


# ----------------------------------------------------------------------------------------------
# taxonomy
# ----------------------------------------------------------------------------------------------
@dataclass
class Taxonomy:
    parent: dict = field(default_factory=dict)     # tid -> parent tid (root: itself)
    rank: dict = field(default_factory=dict)       # tid -> rank string
    name: dict = field(default_factory=dict)
    leaves: list = field(default_factory=list)     # genome tids (strain level or specials)

    def add(self, tid, parent, rank, name=None):
        assert tid not in self.parent, tid
        self.parent[tid] = parent
        self.rank[tid] = rank
        self.name[tid] = name or f"{rank}_{tid}"

    def path_to_root(self, tid):
        """Strict ancestors, nearest first (TaxTree.hpp:60-91)."""
        out = []
        while self.parent[tid] != tid:
            tid = self.parent[tid]
            out.append(tid)
        return out

    def depth(self, tid):
        return len(self.path_to_root(tid))

    def children(self):
        ch = {t: [] for t in self.parent}
        for t, p in self.parent.items():
            if t != p:
                ch[p].append(t)
        return ch

    def tids(self):
        return sorted(self.parent)


def make_taxonomy(seed: int, n_leaves: int, levels=("kingdom", "phylum", "class", "order", "family",
                                                    "genus", "species", "strain"),
                  top_fanout: int = 8, specials: bool = False, first_tid: int = 100) -> Taxonomy:
    """Balanced random taxonomy: root(1) -> levels[0] -> ... -> levels[-1] (= the genome tids).

    Internal nodes get small tids, leaves get tids from 100000 up.  With ``specials`` the tree also
    carries the hard-coded tids of include/tid_checks.hpp and read_label.cpp:38,69,82-104: a human
    species 9606 with siblings 63221/741158, PhiX 10847, a plasmid tid in [1e7, 1.1e7), 20999999 and
    the "bad genome" 12721, so that the golden fixtures exercise those branches.
    """
    rng = rng_for(seed)
    tax = Taxonomy()
    tax.add(1, 1, "no rank", "root")
    nlev = len(levels)
    # number of nodes per level: geometric interpolation between top_fanout and n_leaves
    counts = [max(1, int(round(top_fanout * (n_leaves / top_fanout) ** (i / (nlev - 1))))) for i in range(nlev)]
    counts[0] = min(top_fanout, n_leaves)
    counts[-1] = n_leaves
    for i in range(1, nlev):
        counts[i] = max(counts[i], counts[i - 1])
    next_tid = first_tid
    prev = [1]
    for li, lev in enumerate(levels):
        cur = []
        n = counts[li]
        # each parent gets at least one child, the rest are spread randomly
        owners = list(range(len(prev))) + list(rng.integers(0, len(prev), size=n - len(prev)))
        owners.sort()
        for o in owners:
            if li == nlev - 1:
                tid = 100000 + len(cur)
            else:
                tid = next_tid
                next_tid += 1
            tax.add(tid, prev[o], lev)
            cur.append(tid)
        prev = cur
    tax.leaves = list(prev)
    if specials:
        genus = [t for t, r in tax.rank.items() if r == "genus"]
        g0, g1, g2 = genus[0], genus[len(genus) // 2], genus[-1]
        tax.add(9606, g0, "species", "Homo sapiens")
        tax.add(63221, 9606, "subspecies", "Homo sapiens neanderthalensis")
        tax.add(741158, 9606, "subspecies", "Homo sapiens ssp. Denisova")
        tax.add(10847, g1, "species", "Enterobacteria phage phiX174")
        sp = [t for t, r in tax.rank.items() if r == "species" and tax.parent[t] == g2][0]
        tax.add(10000123, sp, "no rank", "plasmid pKMAT")
        tax.add(20999999, g2, "no rank", "synthetic filler")
        tax.add(12721, g2, "species", "HIV bad genome")
        tax.leaves += [9606, 63221, 741158, 10847, 10000123, 20999999, 12721]
    return tax


def taxonomy_tree_text(tax: Taxonomy) -> str:
    ch = tax.children()
    lines = ["#kmat synthetic taxonomy", "#format: id nchild child... parent / name", str(len(tax.parent))]
    for tid in tax.tids():
        kids = sorted(ch[tid])
        lines.append(" ".join(str(x) for x in [tid, len(kids), *kids, tax.parent[tid]]))
        lines.append(tax.name[tid])
    return "\n".join(lines)          # no trailing newline (phantom-node hazard)


def write_taxonomy_files(tax: Taxonomy, outdir: str, prefix: str = "tax") -> dict:
    os.makedirs(outdir, exist_ok=True)
    p = {k: os.path.join(outdir, f"{prefix}.{k}") for k in ("tree", "depth", "rank", "map16", "numrank", "names")}
    with open(p["tree"], "w") as f:
        f.write(taxonomy_tree_text(tax))
    with open(p["depth"], "w") as f:
        for t in tax.tids():
            f.write(f"{t} {tax.depth(t)}\n")
    with open(p["rank"], "w") as f:
        for t in tax.tids():
            f.write(f"{t} {tax.rank[t].replace(' ', '_')}\n")
    with open(p["map16"], "w") as f:          # bin/Tid16_getMapping.py:81-96: 1->1, others from 2
        f.write("1 1\n")
        for i, t in enumerate(x for x in tax.tids() if x != 1):
            f.write(f"{t} {i + 2}\n")
    with open(p["numrank"], "w") as f:        # bin/build_tid_numeric_rank_table.py:21-71
        for t in tax.tids():
            f.write(f"{t} {numeric_rank(tax, t)}\n")
    with open(p["names"], "w") as f:          # read_label.cpp:1812-1835 (-u), names for fastsummary
        for t in tax.tids():
            f.write(f"depth={tax.depth(t)},taxid={t},ktaxid={t},entries=0\t{tax.rank[t]},{tax.name[t]}\n")
    return p


def numeric_rank(tax: Taxonomy, tid: int) -> int:
    inter = False
    for t in [tid] + tax.path_to_root(tid):
        r = tax.rank[t]
        if r in NUMERIC_RANK:
            return NUMERIC_RANK[r] + (1 if inter else 0)
        inter = True
    return 1


def map16(tax: Taxonomy) -> dict:
    m = {1: 1}
    for i, t in enumerate(x for x in tax.tids() if x != 1):
        m[t] = i + 2
    return m


# ----------------------------------------------------------------------------------------------
# genomes / k-mers
# ----------------------------------------------------------------------------------------------
def random_codes(rng, n, gc):
    """n random 2-bit codes (A=0,C=1,G=2,T=3) with GC fraction gc."""
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return rng.choice(4, size=n, p=p).astype(np.uint8)


def make_genomes(seed: int, tax: Taxonomy, genome_len: int, share_frac: float = 0.10, mut: float = 0.02,
                 gc_range=(0.3, 0.7), conserved_rank: str | None = None, conserved_len: int = 0,
                 conserved_mut: float = 0.002) -> dict:
    """tid -> uint8 code array.  ``share_frac`` of each genome is copied (with ``mut`` substitutions)
    from the previous genome under the same parent, creating multi-tid k-mers.
    With ``conserved_rank`` every node of that rank also plants one ``conserved_len`` segment into all
    genomes below it, which yields long taxid lists (tens of tids) for the k-mers inside it."""
    rng = rng_for(seed)
    genomes = {}
    last_by_parent = {}
    for tid in tax.leaves:
        gc = rng.uniform(*gc_range)
        g = random_codes(rng, genome_len, gc)
        par = tax.parent[tid]
        sib = last_by_parent.get(par)
        if sib is None and par in genomes:      # child of a genome-bearing node (human subspecies, plasmid)
            sib = par
        if sib is not None and share_frac > 0:
            n = int(genome_len * share_frac)
            src0 = int(rng.integers(0, genome_len - n + 1))
            dst0 = int(rng.integers(0, genome_len - n + 1))
            seg = genomes[sib][src0:src0 + n].copy()
            flip = rng.random(n) < mut
            seg[flip] = (seg[flip] + rng.integers(1, 4, size=int(flip.sum()))) % 4
            g[dst0:dst0 + n] = seg
        genomes[tid] = g
        last_by_parent[par] = tid
    if conserved_rank and conserved_len:
        for node in [t for t in tax.tids() if tax.rank[t] == conserved_rank]:
            seg0 = random_codes(rng, conserved_len, rng.uniform(*gc_range))
            for tid in tax.leaves:
                if node in tax.path_to_root(tid) and len(genomes[tid]) >= conserved_len:
                    seg = seg0.copy()
                    flip = rng.random(conserved_len) < conserved_mut
                    seg[flip] = (seg[flip] + rng.integers(1, 4, size=int(flip.sum()))) % 4
                    d0 = int(rng.integers(0, len(genomes[tid]) - conserved_len + 1))
                    genomes[tid][d0:d0 + conserved_len] = seg
    return genomes


def codes_to_str(codes) -> str:
    return BASES[codes].tobytes().decode()


def canonical_kmers(codes: np.ndarray, k: int) -> np.ndarray:
    """All canonical k-mers of an all-ACGT code array (kencode.hpp:76-82 packing, min(fwd, rc) rule of
    kmerPrefixCounter.cpp:140-141 / read_label.cpp:1009)."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    c = codes.astype(np.uint64)
    fwd = np.zeros(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        fwd = (fwd << np.uint64(2)) | c[j:j + n]
        rev |= (np.uint64(3) - c[j:j + n]) << np.uint64(2 * j)
    return np.minimum(fwd, rev)


def lca_subtree(tax: Taxonomy, tids) -> list:
    """The tid set tax_histo stores for a k-mer (TaxTree.hpp:159-260): every input tid plus every
    node between them and their LCA, LCA included.  Deterministic order: input tids ascending, then
    the added ancestors by decreasing depth (the reference's own order is its unordered_map's)."""
    tids = sorted(set(tids))
    if len(tids) == 1:
        return tids
    paths = {t: [t] + tax.path_to_root(t) for t in tids}
    common = set(paths[tids[0]])
    for t in tids[1:]:
        common &= set(paths[t])
    lca = max(common, key=lambda x: tax.depth(x))
    out = set(tids)
    for t in tids:
        for a in paths[t]:
            out.add(a)
            if a == lca:
                break
    extra = sorted(out - set(tids), key=lambda x: (-tax.depth(x), x))
    return tids + extra


def build_kmer_table(genomes: dict, tax: Taxonomy, k: int):
    """(kmers ascending uint64, offs uint64[n+1], tids uint32) -- the logical content of a tax_histo file."""
    ks, gs = [], []
    for tid, g in genomes.items():
        km = np.unique(canonical_kmers(g, k))
        ks.append(km)
        gs.append(np.full(len(km), tid, dtype=np.uint32))
    ks = np.concatenate(ks)
    gs = np.concatenate(gs)
    order = np.lexsort((gs, ks))
    ks, gs = ks[order], gs[order]
    uniq, start, cnt = np.unique(ks, return_index=True, return_counts=True)
    offs = np.zeros(len(uniq) + 1, dtype=np.uint64)
    tids_out = []
    cache = {}
    single = cnt == 1
    # singletons are the bulk: handle them vectorised, the rest through the LCA cache
    lens = np.ones(len(uniq), dtype=np.uint64)
    multi_idx = np.nonzero(~single)[0]
    multi_sets = []
    for i in multi_idx:
        key = tuple(gs[start[i]:start[i] + cnt[i]].tolist())
        s = cache.get(key)
        if s is None:
            s = cache[key] = lca_subtree(tax, key)
        multi_sets.append(s)
        lens[i] = len(s)
    offs[1:] = np.cumsum(lens)
    tids_out = np.zeros(int(offs[-1]), dtype=np.uint32)
    tids_out[offs[:-1][single].astype(np.int64)] = gs[start[single]]
    for i, s in zip(multi_idx, multi_sets):
        o = int(offs[i])
        tids_out[o:o + len(s)] = s
    return uniq.astype(np.uint64), offs, tids_out


def write_tax_histo(path: str, k: int, kmers, offs, tids) -> None:
    """tax_histo binary: 29-byte header, then ``u64 kmer; u16 n; n x u32 tid`` with a
    ``u64 ~0`` sanity word after every 1500th record (metag_typedefs.hpp:9, tax_histo.cpp:273-277)."""
    n = len(kmers)
    with open(path, "wb") as f:
        f.write(struct.pack("<IQQIcI", 29, n, 0xFFFFFFFFFFFFFFFF, 999, b"N", k))
        buf = bytearray()
        for i in range(n):
            a, b = int(offs[i]), int(offs[i + 1])
            buf += struct.pack("<QH", int(kmers[i]), b - a)
            buf += np.asarray(tids[a:b], dtype="<u4").tobytes()
            if (i + 1) % 1500 == 0:
                buf += b"\xff" * 8
            if len(buf) > (1 << 22):
                f.write(buf)
                buf = bytearray()
        f.write(buf)


def write_kpc_fasta(path: str, genomes: dict) -> None:
    with open(path, "w") as f:
        for tid, g in genomes.items():
            f.write(f">{tid}\n{codes_to_str(g)}\n")


# ----------------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------------
def simulate_reads(seed: int, genomes: dict, n_reads: int, read_len: int = 150, novel_frac: float = 0.10,
                   err_lo: float = 0.001, err_hi: float = 0.02, n_rate: float = 0.0005, lower_frac: float = 0.0,
                   len_jitter: int = 0):
    """Illumina-like reads: 90 % sampled from genomes, 10 % random, substitution error
    rising linearly along the read, rare N, 50 % reverse strand.  Returns (headers, sequences)."""
    rng = rng_for(seed)
    tids = list(genomes)
    hdrs, seqs = [], []
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    for i in range(n_reads):
        L = read_len if not len_jitter else int(rng.integers(max(1, read_len - len_jitter), read_len + len_jitter + 1))
        if rng.random() < novel_frac:
            codes = random_codes(rng, L, rng.uniform(0.3, 0.7))
            src = "novel"
        else:
            tid = tids[int(rng.integers(0, len(tids)))]
            g = genomes[tid]
            LL = min(L, len(g))
            p = int(rng.integers(0, len(g) - LL + 1))
            codes = g[p:p + LL].copy()
            src = f"g{tid}_{p}"
        L = len(codes)
        perr = np.linspace(err_lo, err_hi, L)
        flip = rng.random(L) < perr
        codes[flip] = (codes[flip] + rng.integers(1, 4, size=int(flip.sum()))) % 4
        if rng.random() < 0.5:
            codes = comp[codes[::-1]]
        s = bytearray(BASES[codes].tobytes())
        ns = np.nonzero(rng.random(L) < n_rate)[0]
        for j in ns:
            s[j] = ord("N")
        if lower_frac and rng.random() < lower_frac:
            s = bytearray(bytes(s).lower())
        hdrs.append(f"r{i}_{src}")
        seqs.append(bytes(s).decode())
    return hdrs, seqs


def write_fasta(path: str, hdrs, seqs, wrap: int = 0) -> None:
    with open(path, "w") as f:
        for h, s in zip(hdrs, seqs):
            f.write(f">{h}\n")
            if wrap:
                for i in range(0, len(s), wrap):
                    f.write(s[i:i + wrap] + "\n")
            else:
                f.write(s + "\n")


def write_fastq(path: str, hdrs, seqs) -> None:
    with open(path, "w") as f:
        for h, s in zip(hdrs, seqs):
            f.write(f"@{h}\n{s}\n+\n{'I' * len(s)}\n")


# ----------------------------------------------------------------------------------------------
# null models
# ----------------------------------------------------------------------------------------------
CLASS_RANKS = ("genus", "family", "order", "class", "phylum", "kingdom")


def null_class(tax: Taxonomy, tid: int) -> str:
    """``<rank>-<tid>`` of the first genus-or-higher ancestor (bin/merge_cnts.py:233-258,304)."""
    for a in tax.path_to_root(tid):
        if tax.rank[a] in CLASS_RANKS:
            return f"{tax.rank[a]}-{a}"
    return "no_rank-1"


def write_null_models(seed: int, tax: Taxonomy, outdir: str, kmer_counts=(31, 56, 81, 106, 131), nbins: int = 10,
                      lo: float = 0.005, hi: float = 0.08, compress: bool = True, holes: bool = True) -> str:
    """Null-model list + files in the format loadRandHits parses (read_label.cpp:512-678).  Paths in the
    list are relative to $LMAT_DIR (= outdir).  ``holes`` adds num_obs==0 bins of both kinds
    (kmer_cnt >= 100000 -> 0.5; < 100000 -> nearest-bin fill) so those branches are exercised."""
    rng = rng_for(seed)
    os.makedirs(outdir, exist_ok=True)
    lst = os.path.join(outdir, "null_lst.txt")
    with open(lst, "w") as fl:
        for kc in kmer_counts:
            name = f"null.{kc}.rand_lst" + (".gz" if compress else "")
            fl.write(f"{kc} {name}\n")
            lines = [str(nbins)]
            for tid in tax.tids():
                cls = null_class(tax, tid)
                parts = [str(tid), cls]
                for b in range(nbins):
                    val = rng.uniform(lo, hi)
                    obs, kcnt = int(rng.integers(1, 50)), int(rng.integers(1000, 3000000))
                    if holes:
                        u = rng.random()
                        if u < 0.04:
                            obs, kcnt = 0, int(rng.integers(100000, 3000000))
                        elif u < 0.10:
                            obs, kcnt = 0, int(rng.integers(100, 99999))
                    parts += [str(obs), f"{val:.6g}", str(kcnt)]
                lines.append(" ".join(parts))
            data = ("\n".join(lines) + "\n").encode()
            path = os.path.join(outdir, name)
            if compress:
                with gzip.open(path, "wb") as f:
                    f.write(data)
            else:
                with open(path, "wb") as f:
                    f.write(data)
    return lst
