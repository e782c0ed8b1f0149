"""Shared fixture scenarios for the parity tests and for tests/golden/make_golden.py.

A scenario = seeded taxonomy + genomes + reads + null models written to a work directory.  The DB
itself is built by the caller: here in the build container through the unmodified reference chain
(oracle/refchain.py), on the GPU box from the committed golden table dump.
"""
from __future__ import annotations

import os

from lmat_b200 import fixtures as fx

K = 20

SCENARIOS = {
    # name: (tax seed, n_leaves, genome_len, share, conserved_rank, conserved_len, n_reads, read_len)
    "small": dict(seed=7, n_leaves=24, genome_len=4000, share=0.3, cons_rank=None, cons_len=0, n_reads=400, read_len=150),
    "lists": dict(seed=17, n_leaves=90, genome_len=3000, share=0.4, cons_rank="order", cons_len=500, n_reads=600, read_len=150),
    "mixed": dict(seed=27, n_leaves=40, genome_len=5000, share=0.3, cons_rank="family", cons_len=300, n_reads=500, read_len=120),
}

# option sets applied to read_label (reference flags) / to the oracle and product (field names)
OPTION_SETS = {
    "run_rl": dict(null=True, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True),
    "nonull": dict(null=False, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True),
    "defaults": dict(null=True, min_kmer=35, hbias=3.0, sdiff=1.0, min_score=0.0, prn_all=False),
    "tight": dict(null=True, min_kmer=30, hbias=1.5, sdiff=0.25, min_score=0.5, prn_all=True),
    "permissive": dict(null=True, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True, permissive=True),
    "prune3": dict(null=True, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True, prune=3),
    "nophix_hide": dict(null=True, min_kmer=30, hbias=0.0, sdiff=2.0, min_score=0.0, prn_all=False, phix_off=True, hide_read=True),
    "quirk": dict(null=True, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True, min_fnd=40),
    "plasmid": dict(null=False, min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, prn_all=True, plasmids=True),
}


def build_inputs(name: str, workdir: str) -> dict:
    """Write every text input of a scenario; returns paths + in-memory objects."""
    sc = SCENARIOS[name]
    os.makedirs(workdir, exist_ok=True)
    tax = fx.make_taxonomy(sc["seed"], sc["n_leaves"], specials=True)
    paths = fx.write_taxonomy_files(tax, workdir)
    genomes = fx.make_genomes(sc["seed"] + 1, tax, sc["genome_len"], share_frac=sc["share"],
                              conserved_rank=sc["cons_rank"], conserved_len=sc["cons_len"])
    paths["genomes"] = os.path.join(workdir, "genomes.fa")
    fx.write_kpc_fasta(paths["genomes"], genomes)
    hdrs, seqs = fx.simulate_reads(sc["seed"] + 2, genomes, sc["n_reads"], sc["read_len"], n_rate=0.003,
                                   lower_frac=0.1, len_jitter=40)
    # hand-made edge cases: too short, exactly k, all N, GC-only (bin_sel = 10), period-25 repeat
    # (the silent-NoMatch quirk of SURVEY.md 2.2.7), a read with no header
    g0 = fx.codes_to_str(next(iter(genomes.values())))
    extra = [("short", "ACGTACGTAC"), ("exact_k", g0[100:120]), ("all_n", "N" * 60),
             ("gc_only", "GC" * 40), ("period25", (g0[200:225] * 4)[:80]), ("", g0[300:420]),
             ("lower_n", g0[500:560].lower() + "n" + g0[561:640].lower())]
    for h, s in extra:
        hdrs.append(h)
        seqs.append(s)
    paths["reads"] = os.path.join(workdir, "reads.fa")
    fx.write_fasta(paths["reads"], hdrs, seqs)
    paths["reads_wrapped"] = os.path.join(workdir, "reads_wrapped.fa")
    fx.write_fasta(paths["reads_wrapped"], hdrs, seqs, wrap=60)
    paths["reads_fq"] = os.path.join(workdir, "reads.fq")
    fx.write_fastq(paths["reads_fq"], [h or "nohdr" for h in hdrs], seqs)
    paths["null_lst"] = fx.write_null_models(sc["seed"] + 3, tax, workdir)
    paths["plasmids"] = os.path.join(workdir, "plasmids.txt")
    with open(paths["plasmids"], "w") as f:
        f.write(f"{tax.leaves[0]}\n{tax.leaves[3]}\n")
    paths["workdir"] = workdir
    return dict(paths=paths, tax=tax, genomes=genomes, hdrs=hdrs, seqs=seqs)


# ------------------------------------------------------------------------------------------------
# gene_label scenario: a gene DB over the genomes of a read_label scenario
# ------------------------------------------------------------------------------------------------
GENE_LEN = 500
GENE_ID0 = 7_000_000            # 32-bit ids, far above the 16-bit range


def build_gene_table(name: str):
    """Cut every genome of scenario `name` into GENE_LEN-base "genes" (ids GENE_ID0 + running index) and map every
    canonical k-mer to the ascending list of genes containing it.  Returns (kmers, offs, gene ids, annotation lines);
    genes of sibling genomes share k-mers through the scenario's shared segments, so multi-gene lists occur."""
    import numpy as np
    sc = SCENARIOS[name]
    tax = fx.make_taxonomy(sc["seed"], sc["n_leaves"], specials=True)
    genomes = fx.make_genomes(sc["seed"] + 1, tax, sc["genome_len"], share_frac=sc["share"],
                              conserved_rank=sc["cons_rank"], conserved_len=sc["cons_len"])
    pairs, annot, gid = [], [], GENE_ID0
    for tid, codes in genomes.items():
        for a in range(0, len(codes) - GENE_LEN + 1, GENE_LEN):
            km = np.unique(fx.canonical_kmers(codes[a:a + GENE_LEN], K))
            pairs.append(np.stack([km, np.full(len(km), gid, dtype=np.uint64)], axis=1))
            annot.append(f"{tid}\t{gid}\tgene_{gid - GENE_ID0}_of_{tid}\thypothetical protein {a}..{a + GENE_LEN}")
            gid += 1
    allp = np.concatenate(pairs)
    order = np.lexsort((allp[:, 1], allp[:, 0]))
    allp = allp[order]
    kmers, start = np.unique(allp[:, 0], return_index=True)
    offs = np.concatenate([start, [len(allp)]]).astype(np.uint64)
    return kmers.astype(np.uint64), offs, allp[:, 1].astype(np.uint32), annot


# ------------------------------------------------------------------------------------------------
# rand_read_label scenario.  The reference draws its reads itself (rand(), seeded by time(0)); with
# the seed fixed (oracle/standins/fixed_time.c) they are the reads oracle_py.gen_rand_reads predicts, and the DB is
# built FROM mutated fragments of those reads so that "random" reads do hit it (a 20-mer of a truly random read is in a
# small DB with probability ~1e-7).
# ------------------------------------------------------------------------------------------------
NULLGEN_TIME = 1700000123            # the value time(0) returns under the shim = the srand() seed
NULLGEN_RUNS = {
    # tag: read length (-i), reads per thread (-g), pruning (-h N with -r numeric ranks)
    "rl150": dict(read_len=150, n_reads=600, prune=None),
    "rl64p": dict(read_len=64, n_reads=500, prune=3),
}
NULLGEN_TAX = dict(seed=37, n_leaves=30)


def build_nullgen_inputs(workdir: str) -> dict:
    import numpy as np
    from oracle import oracle_py as op
    os.makedirs(workdir, exist_ok=True)
    tax = fx.make_taxonomy(NULLGEN_TAX["seed"], NULLGEN_TAX["n_leaves"], specials=True)
    paths = fx.write_taxonomy_files(tax, workdir)
    rng = fx.rng_for(NULLGEN_TAX["seed"] + 1)
    code = {"a": 0, "c": 1, "g": 2, "t": 3}
    frags = {t: [fx.random_codes(rng, 300, 0.5)] for t in tax.leaves}
    leaves = list(tax.leaves)
    for tag, run in NULLGEN_RUNS.items():
        reads = op.gen_rand_reads(NULLGEN_TIME, run["n_reads"], run["read_len"])
        for r in reads:
            if rng.random() > 0.6:
                continue
            codes = np.array([code[ch] for ch in r], dtype=np.uint8)
            li = int(rng.integers(0, len(leaves)))
            owners = [leaves[li]]
            if rng.random() < 0.5:
                owners.append(leaves[(li + 1) % len(leaves)])     # usually a sibling: shared k-mers, multi-tid lists
            if rng.random() < 0.2:
                owners.append(leaves[int(rng.integers(0, len(leaves)))])
            for o in owners:
                a = int(rng.integers(0, max(1, len(codes) - 30)))
                b = int(rng.integers(min(len(codes), a + 30), len(codes) + 1))
                seg = codes[a:b].copy()
                flip = rng.random(len(seg)) < rng.uniform(0, 0.03)
                seg[flip] = (seg[flip] + rng.integers(1, 4, size=int(flip.sum()))) % 4
                frags[o].append(seg)
    genomes = {t: np.concatenate(v) for t, v in frags.items()}
    paths["genomes"] = os.path.join(workdir, "genomes.fa")
    fx.write_kpc_fasta(paths["genomes"], genomes)
    paths["workdir"] = workdir
    return dict(paths=paths, tax=tax, genomes=genomes)


# ------------------------------------------------------------------------------------------------
# content_summ scenario: the read_label output of the `lists` scenario, summarised
# ------------------------------------------------------------------------------------------------
CONTENT_SUMM_RUNS = {
    # flags of bin/run_cs.sh:148 (k-values 8,10,12,14,17; ranks plasmid,species,genus; the extra plasmid list)
    "run_cs": dict(k="8,10,12,14,17", ranks="plasmid,species,genus", threshold=None, skip_human=False, plasmids=True),
    "k20_thr": dict(k="20,9", ranks="species", threshold=0.4, skip_human=True, plasmids=False),
}


def content_summ_inputs(workdir: str, golden_dir: str) -> dict:
    """The reference read_label output of the lists scenario cut into two files (two "threads"), the list naming them
    and the .fastsummary."""
    import gzip
    raw = gzip.open(os.path.join(golden_dir, "lists.run_rl.out.gz")).read()
    cut = raw.index(b"\n", len(raw) // 2) + 1
    parts = []
    for i, blob in enumerate((raw[:cut], raw[cut:])):
        p = os.path.join(workdir, f"rl{i}.out")
        with open(p, "wb") as f:
            f.write(blob)
        parts.append(p)
    lst = os.path.join(workdir, "rl.lst")
    with open(lst, "w") as f:
        f.write("\n".join(parts) + "\n")
    fs = os.path.join(workdir, "rl.fastsummary")
    with open(os.path.join(golden_dir, "lists.run_rl.fastsummary"), "rb") as src, open(fs, "wb") as dst:
        dst.write(src.read())
    return dict(lst=lst, fastsummary=fs, parts=parts)


def content_summ_args(run: dict, P: dict, files: dict, ofbase: str) -> list:
    a = []
    if run["skip_human"]:
        a += ["-s"]
    if run["plasmids"]:
        a += ["-p", P["plasmids"]]
    a += ["-c", P["tree"], "-l", files["fastsummary"], "-k", run["k"], "-f", files["lst"], "-r", P["rank"], "-a", run["ranks"]]
    if run["threshold"] is not None:
        a += ["-v", str(run["threshold"])]
    return a + ["-o", ofbase]


# ------------------------------------------------------------------------------------------------
# BASELINE configs[0] ("c1"): the reference's bundled data.  DB: the five adenovirus genomes of
# src/kmerdb/examples/tests/data/test.fa under a small hand-made taxonomy (the tarball ships no DB); reads: the 1000 real
# reads of example/example.tgz (simple_list.1000.fna: headers with spaces, 80-column wrapped lines, 60-250 bases) and 400
# reads simulated from the genomes.  The table dump and both read files are committed under tests/golden/ (made by
# tests/golden/make_golden_c1.py from the reference tree); this function only writes the seeded text inputs.
# ------------------------------------------------------------------------------------------------
C1_GENOME_TIDS = (1001, 1002, 1003, 1004, 1005)


def c1_taxonomy():
    tax = fx.Taxonomy()
    tax.add(1, 1, "no rank", "root")
    tax.add(10, 1, "superkingdom", "Viruses")
    tax.add(100, 10, "family", "Adenoviridae")
    tax.add(101, 100, "genus", "Mastadenovirus")
    tax.add(102, 100, "genus", "Atadenovirus")
    tax.add(201, 101, "species", "Human mastadenovirus C")
    tax.add(1001, 101, "species", "Human adenovirus 52")
    tax.add(1002, 101, "species", "Simian adenovirus 7")
    tax.add(1003, 201, "strain", "Human adenovirus C serotype 5")
    tax.add(1004, 101, "species", "Porcine adenovirus 3")
    tax.add(1005, 102, "species", "Snake adenovirus")
    tax.leaves = list(C1_GENOME_TIDS)
    return tax


def build_c1_inputs(workdir: str) -> dict:
    import gzip
    import shutil
    os.makedirs(workdir, exist_ok=True)
    tax = c1_taxonomy()
    paths = fx.write_taxonomy_files(tax, workdir)
    paths["null_lst"] = fx.write_null_models(1007, tax, workdir)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for key, name in (("reads_example", "c1.reads_example.fna"), ("reads_sim", "c1.reads_sim.fa")):
        paths[key] = os.path.join(workdir, name)
        src = os.path.join(golden, name + ".gz")
        if os.path.exists(src):
            with gzip.open(src, "rb") as f, open(paths[key], "wb") as o:
                shutil.copyfileobj(f, o)
    paths["workdir"] = workdir
    return dict(paths=paths, tax=tax)
