#!/usr/bin/env python3
"""Null-model roll-up: <x>.rand_lst (rand_read_label) -> null.bin.<B>.<x>.rand_lst, the file read_label loads with -n.

A Python 3 restatement of the reference's bin/merge_cnts.py (a Python 2 script; SURVEY.md 8(f-1) "roll-up"), keeping its
Python 2 arithmetic where the result depends on it:
  * `a / b` on ints is floor division (:169,:256);
  * comparing a number with a string never raises: every number orders before every string (:174 `pcnt >= store_pcnt` is
    therefore always False -- the "replacement" branch, which would crash the script, is dead; :263 `pcnt > rval_pcnt[it]`
    is True the first time and a lexicographic string comparison afterwards).
Deviations, on purpose: taxids are processed in ascending order (the reference iterates a Python 2 dict, i.e. in hash order;
the consumer, loadRandHits, read_label.cpp:512-678, keys every line by taxid), and `merge_hack` / the E. coli defaults are
empty instead of a NameError when taxids 561 / 562 are absent from the data.
PARITY: pinned against the reference script itself -- no Python 2 interpreter exists offline, so oracle/py2run.py executes
/root/reference/bin/merge_cnts.py unmodified with Python 2 division, mixed-type ordering and dict iteration order emulated
in the AST (tests/golden/make_golden_rollup.py -> tests/golden/rollup/); tests/test_null_rollup_cpu.py compares the lines.

usage: merge_cnts.py <x.rand_lst> <taxonomy> <rank table> <min_obs> <tax_histo counts | missing> <output> <num_bins>
"""
import sys

HUMAN = (9606, 63221, 741158)
ROLL_RANKS = ("genus", "family", "order", "class", "phylum", "kingdom", "domain", "life")
MAG_DIFF = 100


def py2_ge(a, b):
    """a >= b under Python 2 ordering of mixed types (numbers < strings)."""
    sa, sb = isinstance(a, str), isinstance(b, str)
    if sa == sb:
        return a >= b
    return sa


def py2_gt(a, b):
    sa, sb = isinstance(a, str), isinstance(b, str)
    if sa == sb:
        return a > b
    return sa


def load_taxonomy(path):
    """tid -> parent from the LMAT taxonomy file (:71-95): '#' lines skip two more lines, node lines are followed by a
    name line, the parent is the last token."""
    parents = {}
    with open(path) as a:
        while True:
            line1 = a.readline()
            if not line1:
                break
            if line1[0] == "#":
                a.readline()
                a.readline()
                continue
            vals = line1.split()
            if not vals:
                continue
            tid = int(vals[0])
            a.readline()
            parents[tid] = int(vals[-1].rstrip())
    return parents


def roll_up(rand_lst, taxonomy, rank_file, min_obs, thc_file, output_file, num_bins):
    tax_hist_cnt = {}
    ignore_thc = False
    try:
        with open(thc_file) as f:
            for line in f:
                vals = line.rstrip().split()
                tax_hist_cnt.setdefault(int(vals[0]), int(vals[1]))
    except Exception:
        ignore_thc = True                                               # :58-59: every observed taxid then counts 1 k-mer
    ranks = {}
    with open(rank_file) as f:
        for line in f:
            tid, rank = line.rstrip().split()
            ranks[int(tid)] = rank
    ranks.setdefault(1, "life")
    parents = load_taxonomy(taxonomy)

    def rolls_here(tid, human_yes):
        r = ranks[tid]
        return (r == "species" and human_yes) or r in ROLL_RANKS

    store_rank_val = {}
    is_euk = {}
    with open(rand_lst) as f:                                           # :97-196
        for line in f:
            t = line.rstrip().split()
            if not t:
                continue
            tid = int(t[0])
            if ignore_thc:
                tax_hist_cnt.setdefault(tid, 1)
            if tid not in tax_hist_cnt:
                continue
            curr_tid = parents[tid]
            kmer_cnt = tax_hist_cnt[tid]
            x = tid
            while True:
                if x == 2759:
                    is_euk.setdefault(tid, 1)
                    break
                if x == parents[x]:
                    break
                x = parents[x]
            is_ignore = False
            x = tid
            while True:
                if x in (2, 2157, 28384):
                    is_ignore = True
                    break
                if x == parents[x]:
                    break
                x = parents[x]
            human_yes = tid in HUMAN
            if (not human_yes and tid >= 10000000) or (is_ignore and kmer_cnt < 100000):
                continue
            t.pop(0)
            while True:
                if rolls_here(curr_tid, human_yes):
                    if curr_tid in store_rank_val:
                        lst = store_rank_val[curr_tid]
                        for obi in range(0, num_bins, 2):               # sic: the first num_bins / 2 (value, count) pairs only
                            pcnt = float(t[obi])
                            fnd = False
                            for it in range(len(lst)):
                                obs_lst, store_kmer_cnt = lst[it]
                                for it1 in range(0, len(obs_lst), 2):
                                    chk_diff = kmer_cnt // store_kmer_cnt
                                    # `chk_diff < mag_diff and pcnt >= store_pcnt` (:174): float >= str is False in Python 2
                                    assert not (chk_diff < MAG_DIFF and py2_ge(pcnt, obs_lst[it1]))
                                    if chk_diff < MAG_DIFF:
                                        fnd = True
                                        break
                            if not fnd:
                                store_rank_val[curr_tid].append((t, kmer_cnt))
                    else:
                        store_rank_val.setdefault(curr_tid, [(t, kmer_cnt)])
                    break
                if parents[curr_tid] == curr_tid:
                    break
                curr_tid = parents[curr_tid]
    merge_hack = list(store_rank_val.get(561, []))
    if 620 in store_rank_val:
        merge_hack.extend(store_rank_val[620])
    def_euk = None
    lines = [str(num_bins)]
    qlst = [562] + sorted(tax_hist_cnt.keys())
    once = set()
    for tid in qlst:                                                    # :213-334
        if tid in once:
            continue
        once.add(tid)
        if tid not in parents:
            continue
        curr_tid = parents[tid]
        use_val = []
        tid_kcnt = tax_hist_cnt.get(tid, 0)
        human_yes = tid in HUMAN
        if tid >= 10000000 and not human_yes:
            tid_kcnt = tax_hist_cnt.get(curr_tid, 0)                    # the reference raises KeyError when the parent has no count
        is_other = False
        while True:
            if curr_tid == 28384:
                is_other = True
                break
            if rolls_here(curr_tid, human_yes):
                if curr_tid in store_rank_val:
                    use_val = store_rank_val[curr_tid]
                    if curr_tid in (561, 620):
                        use_val = merge_hack
            if use_val != []:
                break
            if parents[curr_tid] == curr_tid:
                break
            curr_tid = parents[curr_tid]
        if is_other:
            use_val = merge_hack
        if tid == 9606 and 9606 in store_rank_val:
            use_val = store_rank_val[9606]
        rval_pcnt, rval_kcnt, rval_obs = [0] * num_bins, [0] * num_bins, [0] * num_bins
        rval_pcnt1, rval_kcnt1, rval_obs1 = [1.0] * num_bins, [0] * num_bins, [0] * num_bins
        close_match = [-1] * num_bins
        fnd_match = False
        for oblst, kcnt in use_val:
            diff_pcnt = tid_kcnt // kcnt
            for it2 in range(0, len(oblst), 2):
                pcnt, obs = oblst[it2], oblst[it2 + 1]
                it = it2 // 2
                if diff_pcnt < MAG_DIFF and py2_gt(pcnt, rval_pcnt[it]):
                    rval_pcnt[it], rval_obs[it], rval_kcnt[it] = pcnt, obs, kcnt
                    fnd_match = True
                if diff_pcnt < close_match[it] or close_match[it] == -1:
                    rval_pcnt1[it], rval_obs1[it], rval_kcnt1[it] = pcnt, obs, kcnt
                    close_match[it] = diff_pcnt
        if not fnd_match:
            rval_pcnt, rval_kcnt, rval_obs = rval_pcnt1, rval_kcnt1, rval_obs1
        use_rank = "genus" if human_yes else ranks[curr_tid]
        if tid == 562:
            def_euk = (rval_pcnt, rval_obs, rval_kcnt)
        if tid in is_euk and use_rank == "genus" and def_euk is not None:
            rval_pcnt, rval_obs, rval_kcnt = def_euk
        if tid == 1:
            rval_pcnt = [1.0] * num_bins
        save_rit, save_fit = -1, -1
        for it in range(len(rval_pcnt)):                                # bins with too few observations borrow a neighbour's value
            if int(rval_obs[it]) < min_obs:
                for rit in range(it - 1, -1, -1):
                    if int(rval_obs[rit]) >= min_obs:
                        save_rit = rit
                        break
                for fit in range(it + 1, len(rval_pcnt)):
                    if int(rval_obs[fit]) >= min_obs:
                        save_fit = fit
                        break
                d1 = abs(it - save_rit) if save_rit >= 0 else num_bins + 1
                d2 = abs(it - save_fit) if save_fit >= 0 else num_bins + 1
                if d1 <= d2 and save_rit != -1:
                    rval_pcnt[it] = rval_pcnt[save_rit]
                elif save_fit != -1:
                    rval_pcnt[it] = rval_pcnt[save_fit]
        out = str(tid) + " " + str(use_rank) + "-" + str(curr_tid)
        for it in range(len(rval_pcnt)):
            out += " " + str(rval_obs[it]) + " " + str(rval_pcnt[it]) + " " + str(rval_kcnt[it])
        lines.append(out)
    with open(output_file, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(lines) - 1


def main(argv):
    if len(argv) != 8:
        print(__doc__)
        return 1
    n = roll_up(argv[1], argv[2], argv[3], int(argv[4]), argv[5], argv[6], int(argv[7]))
    print("how much", n)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
