#!/bin/bash
# line table iteration 4: cheaper key assembly; 2 vs 3 CTAs per SM for the probe kernel; full GPU suite
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02g_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -2 gpurun_out/r02g_bench.err
KMAT_LIB=$PWD/lmat_b200/variants/libkmat_c3.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02g_bench_c3.json 2> gpurun_out/r02g_bench_c3.err; tail -2 gpurun_out/r02g_bench_c3.err
python - <<'PY'
import json
for n in ("bench", "bench_c3"):
    j = json.loads(open(f"gpurun_out/r02g_{n}.json").read().strip().splitlines()[-1])
    print(n, round(j["value"]/1e6,1), j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["labels_checksum_rank0"])
PY
