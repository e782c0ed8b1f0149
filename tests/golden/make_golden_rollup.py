#!/usr/bin/env python
"""Golden files of the null-model roll-up: the reference's own bin/merge_cnts.py (a Python 2 script) executed UNMODIFIED
from /root/reference under oracle/py2run.py (Python 2 division / mixed-type ordering / dict iteration order emulated in
the AST) on hand-made inputs that reach the branches real NCBI data reaches: E. coli / Shigella (561, 562, 620: the
merge_hack defaults), Eukaryota (2759: the E. coli default for genus-level eukaryotes), human (9606, 63221), "other
sequences" (28384), Archaea, plasmid-range ids (>= 10^7), bins with too few observations (min_obs borrowing), a missing
k-mer count table (every taxid counts 1 k-mer).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_rollup.py
Writes tests/golden/rollup/{tax.dat,rank.txt,counts.txt,in.rand_lst,out.<case>.txt}.  tests/test_null_rollup_cpu.py compares
lmat_b200/tools/merge_cnts.py with them."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import py2run  # noqa: E402

OUT = os.path.join(HERE, "rollup")
REF = "/root/reference/bin/merge_cnts.py"

# tid: (parent, rank)
TREE = {
    1: (1, "no_rank"),
    131567: (1, "no_rank"), 2: (131567, "domain"), 1224: (2, "phylum"), 1236: (1224, "class"), 91347: (1236, "order"), 543: (91347, "family"),
    561: (543, "genus"), 562: (561, "species"), 83333: (562, "strain"), 511145: (83333, "strain"), 199310: (562, "strain"),
    620: (543, "genus"), 623: (620, "species"), 198214: (623, "strain"), 622: (620, "species"), 300267: (622, "strain"),
    570: (543, "genus"), 573: (570, "species"), 272620: (573, "strain"),
    1239: (2, "phylum"), 91061: (1239, "class"), 186826: (91061, "order"), 1300: (186826, "family"), 1301: (1300, "genus"),
    1313: (1301, "species"), 170187: (1313, "strain"), 171101: (1313, "strain"),
    2157: (131567, "domain"), 28890: (2157, "phylum"), 2172: (28890, "genus"), 2173: (2172, "species"), 420247: (2173, "strain"),
    2759: (131567, "domain"), 33208: (2759, "kingdom"), 9604: (33208, "family"), 9605: (9604, "genus"), 9606: (9605, "species"), 63221: (9606, "subspecies"),
    4751: (2759, "kingdom"), 4930: (4751, "genus"), 4932: (4930, "species"), 559292: (4932, "strain"),
    5820: (2759, "genus"), 5833: (5820, "species"), 36329: (5833, "strain"),
    10239: (1, "domain"), 10662: (10239, "genus"), 10665: (10662, "species"),
    28384: (1, "no_rank"), 81077: (28384, "no_rank"), 32630: (81077, "species"),
    10000123: (562, "strain"), 10000456: (1313, "strain"),
}
# taxids that get a line in the .rand_lst (what rand_read_label saw), their k-mer counts, and a few without observations
OBSERVED = [83333, 511145, 199310, 198214, 300267, 272620, 170187, 171101, 420247, 9606, 63221, 559292, 36329, 10665, 32630, 10000123, 562, 1313, 561]
COUNTS = {83333: 4500000, 511145: 4600000, 199310: 5200, 198214: 4400000, 300267: 4300000, 272620: 5300000, 170187: 2100000, 171101: 2000000,
          420247: 1800000, 9606: 2900000000, 63221: 3100, 559292: 12000000, 36329: 23000000, 10665: 168000, 32630: 900, 10000123: 90000,
          562: 4700000, 1313: 2050000, 561: 150, 620: 120, 623: 4400100, 573: 5300100, 4932: 12000100, 1: 7, 2: 5000, 2759: 7000, 9605: 10, 10000456: 5}
NBINS = 10


def rnd(seed):
    state = [seed & 0xFFFFFFFF]

    def nxt():
        state[0] = (1103515245 * state[0] + 12345) & 0x7FFFFFFF
        return state[0]
    return nxt


def write_inputs():
    os.makedirs(OUT, exist_ok=True)
    kids = {}
    for t, (p, r) in TREE.items():
        if t != p:
            kids.setdefault(p, []).append(t)
    with open(os.path.join(OUT, "tax.dat"), "w") as f:                   # TaxTree text format: "tid n_children child... parent" + a name line
        f.write("# hand-made tree for the roll-up goldens\n#\n#\n")
        for t, (p, r) in TREE.items():
            ch = kids.get(t, [])
            f.write(" ".join(str(x) for x in [t, len(ch)] + ch + [p]) + "\n")
            f.write(f"name of {t}\n")
    with open(os.path.join(OUT, "rank.txt"), "w") as f:
        for t, (p, r) in TREE.items():
            if t != 1:
                f.write(f"{t} {r}\n")
    with open(os.path.join(OUT, "counts.txt"), "w") as f:
        for t, c in COUNTS.items():
            f.write(f"{t} {c}\n")
    g = rnd(20261017)
    with open(os.path.join(OUT, "in.rand_lst"), "w") as f:                # "tid (max fraction, observations) x bins", as rand_read_label writes it
        for i, t in enumerate(OBSERVED):
            vals = []
            for b in range(NBINS):
                frac = (g() % 9000) / 10000.0 + 0.01
                obs = g() % 60 if (i + b) % 4 else g() % 3            # some bins below min_obs
                vals += [("%g" % frac), str(obs)]
            f.write(str(t) + " " + " ".join(vals) + "\n")
    # without a count table every bacterial / archaeal / "other" taxid is skipped (k-mer count 1 < 100000), so 561 never gets an
    # entry and the reference raises NameError (merge_hack) at the first "other sequences" taxid: that input leaves 32630 out
    with open(os.path.join(OUT, "in.nocounts.rand_lst"), "w") as f:
        for ln in open(os.path.join(OUT, "in.rand_lst")):
            if not ln.startswith("32630 "):
                f.write(ln)


CASES = {
    "counts_min2": dict(min_obs=2, thc="counts.txt"),
    "counts_min5": dict(min_obs=5, thc="counts.txt"),
    "nocounts_min2": dict(min_obs=2, thc="missing.txt", inp="in.nocounts.rand_lst"),
}


def main():
    write_inputs()
    for name, c in CASES.items():
        out = os.path.join(OUT, f"out.{name}.txt")
        rc = py2run.run_script(REF, [os.path.join(OUT, c.get("inp", "in.rand_lst")), os.path.join(OUT, "tax.dat"), os.path.join(OUT, "rank.txt"), str(c["min_obs"]),
                                     os.path.join(OUT, c["thc"]), out, str(NBINS)])
        assert rc == 0, (name, rc)
        print(name, sum(1 for _ in open(out)) - 1, "lines")


if __name__ == "__main__":
    main()
