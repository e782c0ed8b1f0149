#!/usr/bin/env python
"""Regenerate the content_summ goldens (tests/golden/lists.cs_*.…) from the UNMODIFIED reference (build container only).
Input: the reference read_label output of the `lists` scenario (lists.run_rl.out.gz, cut into two files so that two
"threads" are merged) and its .fastsummary; run: reference content_summ with the flags of bin/run_cs.sh:148."""
import gzip
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from oracle import refchain as rc  # noqa: E402

RUNS = S.CONTENT_SUMM_RUNS


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/kmat_cs_golden"
    shutil.rmtree(wd, ignore_errors=True)
    inp = S.build_inputs("lists", wd)
    P = inp["paths"]
    files = S.content_summ_inputs(wd, HERE)
    for tag, run in RUNS.items():
        ofb = os.path.join(wd, f"{tag}.summ")
        cmd = [os.path.join(rc.REF_BIN, "content_summ")] + S.content_summ_args(run, P, files, ofb)
        rc._run(cmd, log=os.path.join(wd, f"{tag}.log"))
        for fn in sorted(os.listdir(wd)):
            if fn.startswith(f"{tag}.summ"):
                dst = os.path.join(HERE, "lists.cs_" + fn)
                shutil.copy(os.path.join(wd, fn), dst)
                print(fn, os.path.getsize(dst))


if __name__ == "__main__":
    main()
