#!/bin/bash
# two-level table (minimizer-ordered lines + bucket table): new table tests, the whole GPU suite, then the bench with kernel split
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_table_levels.py -m gpu -x -q > gpurun_out/r02d_levels.log 2>&1; echo "levels rc=$?"; tail -15 gpurun_out/r02d_levels.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02d_pytest_gpu.log
for dens in 2 4; do
  KMAT_LINE_DENSITY=$dens timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_d$dens.json 2> gpurun_out/r02d_bench_d$dens.err
  tail -2 gpurun_out/r02d_bench_d$dens.err
done
KMAT_NO_LINE_LEVEL=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02d_bench_noline.json 2> gpurun_out/r02d_bench_noline.err
python - <<'PY'
import json
for n in ("d2", "d4", "noline"):
    try:
        j = json.loads(open(f"gpurun_out/r02d_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"], j["labels_checksum_rank0"], j["setup_s"], (j.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
