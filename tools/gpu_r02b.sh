#!/bin/bash
# Why is the round-1 line-table variant (exp5, -DKMAT_LINE_TABLE=1) 3.8x slower than the bucket table?  Reduced workload
# (400 genomes, 2 M reads): bench line with the extra-bucket statistic for both libraries, then ncu --set full + source of exp5.
set -u
mkdir -p gpurun_out
A="--genomes 400 --reads 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/r02b_default_small.json 2> gpurun_out/r02b_default_small.err
KMAT_LIB=$PWD/lmat_b200/variants/libkmat_exp5.so python bench.py $A > gpurun_out/r02b_exp5_small.json 2> gpurun_out/r02b_exp5_small.err
KMAT_LIB=$PWD/lmat_b200/variants/libkmat_exp5.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_encode_probe_fast -s 2 -c 1 \
    -o gpurun_out/r02b_exp5_full -f python bench.py $A > gpurun_out/r02b_exp5_full.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ("default_small", "exp5_small"):
    j = json.loads(open(f"gpurun_out/r02b_{n}.json").read().strip().splitlines()[-1])
    print(n, j["value"], j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"], j["hit_rate"])
PY
