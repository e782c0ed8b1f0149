"""Parity at bench scale: the drop-in binary (lmat_b200/bin/read_label -> libkmat, GPU) against the UNMODIFIED reference
read_label (oracle/_ref) on the bench workload's own generators -- the 2000-genome C2 taxonomy with its ~4 k stored ids in
the 16-bit map, the C2 null models, a reference-built DB of the first genomes and reads drawn with the bench's read model.
Same DB file, same reads file, same flags (bin/run_rl.sh:243); the per-read records are compared as sorted multisets (this
is bench.py's `parity_at_scale` leg at a size that finishes in well under a minute)."""
import argparse
import os
import shutil
import tempfile

import pytest

from lmat_b200 import api
from oracle import refchain as rc


@pytest.mark.gpu
def test_parity_at_bench_scale():
    import torch
    import bench
    if not (rc.have_ref("make_db_table") and rc.have_ref("read_label")):
        pytest.skip("oracle/_ref binaries not present")
    assert api.device_count() > 0
    a = argparse.Namespace(genomes=2000, genome_len=500000, cpu_genomes=48, cpu_reads=100000, read_len=150)
    wd = tempfile.mkdtemp(prefix="kmat_scale_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        sample = bench.build_cpu_sample(a, wd, "cuda:0")
        assert sample["kind"] == "reference" and sample["n_kmers"] > 20_000_000
        threads = min(16, os.cpu_count() or 1)
        bench.run_cpu_sample(sample, wd, threads)
        par = bench.parity_at_scale(sample, wd, threads)
        assert par["records_ref"] == par["records_kmat"] == sample["n_reads"], par
        assert par["lines_equal"], par
    finally:
        shutil.rmtree(wd, ignore_errors=True)
        shutil.rmtree(bench._sample_cache_dir(a), ignore_errors=True)
        torch.cuda.empty_cache()
