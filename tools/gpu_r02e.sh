#!/bin/bash
# two-level table, second iteration (32-bit sliding minimum, shard range guard): tests, bench, ncu --set full of the LINE kernel
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02e_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -2 gpurun_out/r02e_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_encode_probe_fast -s 2 -c 1 \
    -o gpurun_out/r02e_probe_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02e_probe_full.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e6,1), j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"], j["labels_checksum_rank0"], (j.get("e2e") or {}).get("value"))
PY
