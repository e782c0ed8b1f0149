"""TEST INFRASTRUCTURE: a plain-Python restatement of the reference's content_summ (src/content_summ.cpp), used only by
tests/ to check the GPU path (lmat_b200/bin/content_summ).  Pinned: tests/test_content_summ.py compares its output with
the files the UNMODIFIED reference wrote (tests/golden/lists.cs_*, made by tests/golden/make_content_summ_golden.py).

Every step cites the reference line it follows; quirks are restated on purpose (the shadowed `kos` pointer at :503 that
drops the k-mer coverage of the first taxid of every rank, std::string::find/substr arithmetic on npos, istream extraction
writing 0 on failure).
"""
from __future__ import annotations

import struct

NPOS = (1 << 64) - 1
HUMAN = (9606, 63221, 741158)                      # include/tid_checks.hpp:15-28


def _find(s: str, ch: str, start: int = 0) -> int:
    start &= NPOS
    if start > len(s):
        return NPOS
    i = s.find(ch, start)
    return NPOS if i < 0 else i


def _substr(s: str, pos: int, n: int) -> str:
    pos &= NPOS
    n &= NPOS
    if pos > len(s):
        raise IndexError("std::out_of_range")      # the reference would terminate
    return s[pos:pos + n] if n < len(s) else s[pos:]


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def _extract_uint(tok):
    """istream >> uint32_t (num_get / strtoull semantics): optional sign, digits; failure -> 0."""
    i, n = 0, len(tok)
    neg = False
    if i < n and tok[i] in "+-":
        neg = tok[i] == "-"
        i += 1
    j = i
    while j < n and tok[j].isdigit():
        j += 1
    if j == i:
        return 0, False, tok
    v = int(tok[i:j])
    if v > 0xFFFFFFFF:
        return 0xFFFFFFFF, False, tok[j:]
    if neg:
        v = (-v) & 0xFFFFFFFF
    return v, True, tok[j:]


def _extract_float(text):
    """istream >> float on the leading token of text; failure -> 0.0.  Returns (value, ok, rest)."""
    t = text.lstrip(" \t\n\r\f\v")
    j = 0
    n = len(t)
    if j < n and t[j] in "+-":
        j += 1
    k = j
    while k < n and (t[k].isdigit() or t[k] == "."):
        k += 1
    if k < n and t[k] in "eE" and k > j:
        m = k + 1
        if m < n and t[m] in "+-":
            m += 1
        if m < n and t[m].isdigit():
            while m < n and t[m].isdigit():
                m += 1
            k = m
    try:
        return _f32(float(t[:k])), True, t[k:]
    except ValueError:
        return 0.0, False, t


def canonical_kmers_once(read: str, k: int):
    """retrieve_kmer_labels (:114-155) for one k: the distinct canonical k-mers of the read."""
    code = {"a": 0, "A": 0, "c": 1, "C": 1, "g": 2, "G": 2, "t": 3, "T": 3}
    mask = (1 << (2 * k)) - 1
    hb = (k - 1) * 2
    fwd = rev = 0
    run = 0
    seen = set()
    for ch in read:
        t = code.get(ch)
        if t is None:
            run = 0
            continue
        fwd = ((fwd << 2) | t) & mask
        rev = ((t ^ 3) << hb) | (rev >> 2)
        run += 1
        if run >= k:
            seen.add(fwd if fwd < rev else rev)
    return seen


def content_summ(tree_parent: dict, tree_name: dict, rank_table: dict, fastsummary: str, out_files: list, k_sizes: list, rank_check: set,
                 threshold: float = 0.0, skip_human: bool = False, human_reg: bool = False, plasmids: set = frozenset()):
    """Returns {suffix: text}: '' = the summary written to <ofbase>, '.<rank>_kmer_cov' = the per-rank coverage files."""
    rank_table = dict(rank_table)                                  # operator[] default-inserts "" (:367,:370,:489)

    def rank_of(t):
        return rank_table.setdefault(t, "")

    def is_plasmid(t):
        return 10000000 <= t < 11000000 or t in plasmids          # :46

    def path_to_root(t):                                           # TaxTree::getPathToRoot (TaxTree.hpp:60-91): strict ancestors
        out = []
        if t not in tree_parent:
            return out
        while tree_parent[t] != t:
            t = tree_parent[t]
            out.append(t)
            if t not in tree_parent:
                break
        return out

    want_rank = "region" if human_reg else "species"
    weighted, read_cnts, strain2spec, clst = {}, {}, {}, []
    for line in open(fastsummary, encoding="latin-1").read().split("\n"):     # :357-383 (getline into a 2024-byte buffer)
        if len(line) >= 2023:
            break
        if "\tNULL\t" in line or line == "":
            if line == "":
                continue
            continue
        toks = line.split()
        wght, ok, _ = _extract_float(toks[0]) if toks else (0.0, False, "")
        rc = _extract_uint(toks[1])[0] if len(toks) > 1 else 0
        tid = _extract_uint(toks[2])[0] if len(toks) > 2 else 0
        weighted.setdefault(tid, wght)
        read_cnts.setdefault(tid, rc)
        if rank_of(tid) == want_rank:
            strain2spec.setdefault(tid, tid)
        if not is_plasmid(tid):
            for a in path_to_root(tid):
                if rank_of(a) == want_rank:
                    strain2spec.setdefault(tid, a)
        clst.append(tid)
    thr = _f32(threshold)
    tracks = []                                                    # per "thread" (input file): [k index] -> {tid: {kmer: count}}
    for fn in out_files:
        track = [dict() for _ in k_sizes]
        tracks.append(track)
        data = open(fn, encoding="latin-1").read()
        lines = data.split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        for line in lines:                                         # :405-441
            p1 = _find(line, "\t")
            p2 = _find(line, "\t", p1 + 1)
            p3 = _find(line, "\t", p2 + 1)
            p4 = _find(line, "\t", p3 + 1)
            p5 = _find(line, "\t", p4 + 1)
            read_buff = _substr(line, p1 + 1, p2 - p1 - 1)
            tws = _substr(line, p4 + 1, p5 - p4 - 1)
            if tws[:1] in ("N", "R"):
                continue
            toks = tws.split()
            taxid, ok, rest = _extract_uint(toks[0]) if toks else (0, False, "")
            if ok and rest == "" and len(toks) > 1:
                score, _, _ = _extract_float(toks[1])
            elif ok and rest != "":
                score, _, _ = _extract_float(rest)                 # "12.5" read as taxid 12, score .5 -- not produced by read_label
            else:
                score = 0.0
            if taxid in HUMAN and skip_human:
                continue
            if score < thr:
                continue
            use_tid = taxid
            if taxid in strain2spec and not is_plasmid(taxid):
                use_tid = strain2spec[taxid]
            rnk = rank_table[use_tid] if use_tid in rank_table else "undef"
            if rnk in rank_check or is_plasmid(taxid):
                for ki, k in enumerate(k_sizes):
                    d = track[ki].setdefault(use_tid, {})
                    for km in canonical_kmers_once(read_buff, k):
                        d[km] = d.get(km, 0) + 1
    # ---- the tree of called taxids (:443-462)
    seen, child = set(), {}
    for tid in clst:
        node = tid
        for p in path_to_root(tid):
            if node not in seen:
                seen.add(node)
                child.setdefault(p, []).append(node)
            node = p
    out = {"": "Name\tTaxID\tReads\tWReads\n"}
    tab = {}
    open_l = [1]
    rank_files = {}
    while open_l:                                                  # :470-522
        tid = open_l.pop(0)
        chk = tab.get(tid, "") + "\t"
        tab.setdefault(tid, "")
        for c in child.setdefault(tid, []):
            tab[c] = chk
            open_l.insert(0, c)
        tot = read_cnts.setdefault(tid, 0)
        wrdc = 0.0
        if tot > 0:
            wrdc = weighted.setdefault(tid, 0.0)
            rank = rank_of(tid)
            if rank != "no_rank":
                if is_plasmid(tid):
                    rank = "plasmid"
                kos = None
                if rank in rank_files:
                    kos = rank
                else:
                    rank_files[rank] = ""                          # the new stream lands in a shadowing local (:503): kos stays NULL
                if kos is not None and tot > 1:
                    rank_files[rank] += _kmer_cov(tracks, tid, k_sizes)
        out[""] += tab[tid] + tree_name.get(tid, "") + "\t" + str(tid) + "\t" + str(tot) + "\t" + ("%g" % wrdc) + "\n"
    for r, txt in rank_files.items():
        out["." + r + "_kmer_cov"] = txt
    return out


def _kmer_cov(tracks, tid, k_sizes):                               # compKmerCov (:527-571)
    txt = ""
    for ki, k in enumerate(k_sizes):
        merged = {}
        total = 0
        for track in tracks:
            for km, c in track[ki].get(tid, {}).items():
                total += c
                merged[km] = merged.get(km, 0) + c
        hist = {}
        for c in merged.values():
            hist[c] = hist.get(c, 0) + 1
        txt += f"taxid={tid} distinct_kmer_cnt={len(merged)} k_size={k} tot_kmer_cnt={total}\n"
        for c in sorted(hist):
            txt += f"{tid} {k} {c} {hist[c]}\n"
    return txt


def parse_tree_with_names(path):
    """(parent, name) dicts from the -c taxonomy file (TaxTree ctor, TaxTree.hpp:24-57 / TaxNode::read :131-147)."""
    lines = open(path, encoding="latin-1").read().split("\n")
    body = lines[3:]
    parent, name = {}, {}
    i = 0
    while i < len(body):
        toks = body[i].split()
        if not toks:
            i += 1
            continue
        tid, n = int(toks[0]), int(toks[1])
        parent[tid] = int(toks[2 + n])
        name[tid] = body[i + 1] if i + 1 < len(body) else ""
        i += 2
    return parent, name


def parse_rank_table(path):
    """ifs >> tid >> rank until extraction fails (:326-331)."""
    out = {}
    toks = open(path, encoding="latin-1").read().split()
    i = 0
    while i + 1 < len(toks):
        v, ok, rest = _extract_uint(toks[i])
        if not ok or rest != "":
            break
        out.setdefault(v, toks[i + 1])                             # unordered_map::insert: the first entry wins
        i += 2
    return out
