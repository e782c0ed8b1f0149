"""content_summ (src/content_summ.cpp).

CPU: the plain-Python restatement (oracle/content_summ_py.py) against the files the UNMODIFIED reference wrote for the
`lists` scenario (tests/golden/lists.cs_*, made by tests/golden/make_content_summ_golden.py).
GPU (-m gpu): the content_summ drop-in binary against the same goldens byte for byte, and kmat_kcov_* against the
restatement on ragged reads (N runs, lower case, repeated k-mers, several batches)."""
import os
import subprocess

import numpy as np
import pytest

import scenarios as S
from lmat_b200 import api, build
from oracle import content_summ_py as cs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cs_inputs(golden_lists, tmp_path_factory):
    wd = str(tmp_path_factory.mktemp("cs"))
    files = S.content_summ_inputs(wd, GOLDEN)
    return golden_lists.paths, files, wd


def golden_files(tag):
    pre = f"lists.cs_{tag}.summ"
    return {fn[len(pre):]: open(os.path.join(GOLDEN, fn)).read() for fn in os.listdir(GOLDEN) if fn.startswith(pre)}


@pytest.mark.parametrize("tag", list(S.CONTENT_SUMM_RUNS))
def test_oracle_equals_reference(cs_inputs, tag):
    P, files, _ = cs_inputs
    run = S.CONTENT_SUMM_RUNS[tag]
    parent, name = cs.parse_tree_with_names(P["tree"])
    pl = set(int(x) for x in open(P["plasmids"]).read().split()) if run["plasmids"] else set()
    out = cs.content_summ(parent, name, cs.parse_rank_table(P["rank"]), files["fastsummary"], files["parts"], [int(x) for x in run["k"].split(",")],
                          set(run["ranks"].split(",")), threshold=run["threshold"] or 0.0, skip_human=run["skip_human"], plasmids=pl)
    want = golden_files(tag)
    assert sorted(out) == sorted(want)
    for k in want:
        assert out[k] == want[k], k
    assert len(want) >= 6 and sum(len(v) for v in want.values()) > 10000


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(S.CONTENT_SUMM_RUNS))
def test_content_summ_cli_equals_reference(cs_inputs, tag, tmp_path):
    build.build_all()
    P, files, _ = cs_inputs
    ofb = str(tmp_path / f"{tag}.summ")
    p = subprocess.run([build.CS_BIN] + S.content_summ_args(S.CONTENT_SUMM_RUNS[tag], P, files, ofb), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    want = golden_files(tag)
    got = {fn[len(f"{tag}.summ"):]: open(os.path.join(str(tmp_path), fn)).read() for fn in os.listdir(str(tmp_path)) if fn.startswith(f"{tag}.summ")}
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k] == want[k], k


@pytest.mark.gpu
def test_kcov_equals_restatement_on_ragged_reads():
    rng = np.random.default_rng(9)
    ks = [20, 8, 13, 1]
    base = "".join("ACGT"[x] for x in rng.integers(0, 4, 3000))
    reads, groups = [], []
    for i in range(600):
        L = int(rng.choice([0, 1, 7, 8, 19, 20, 21, 31, 32, 33, 64, 150, 151, 400, 2500]))
        a = int(rng.integers(0, max(1, len(base) - L)))
        s = base[a:a + L]
        if i % 5 == 1 and L > 30:
            s = s[:L // 3] + "N" + s[L // 3 + 1:2 * L // 3] + "nn" + s[2 * L // 3 + 2:]
        if i % 5 == 2:
            s = s.lower()
        if i % 5 == 3 and L > 40:
            s = (s[:L // 2] * 2)[:L]                      # the same k-mers twice in one read: counted once
        if i % 7 == 0:
            s = "A" * L
        reads.append(s)
        groups.append(0xFFFFFFFF if i % 11 == 0 else int(rng.integers(0, 5)) * 1000 + 3)     # sparse group ids, some reads skipped
    kc = api.KmerCov(ks)
    kc.add(reads[:250], groups[:250])
    kc.add(reads[250:], groups[250:])
    kc.finish()
    for ki, k in enumerate(ks):
        for g in sorted(set(groups) - {0xFFFFFFFF}):
            want = {}
            for s, gg in zip(reads, groups):
                if gg == g:
                    for km in cs.canonical_kmers_once(s, k):
                        want[km] = want.get(km, 0) + 1
            hist = {}
            for c in want.values():
                hist[c] = hist.get(c, 0) + 1
            d, t, h = kc.query(ki, g)
            assert (d, t, h) == (len(want), sum(want.values()), hist), (k, g)
    assert kc.query(0, 77) == (0, 0, {})                  # a group nothing was added for
