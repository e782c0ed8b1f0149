#!/bin/bash
# two-level table, third iteration (rolled chunk loop, batched second-level probes): tests, bench, ncu; memcheck of the direct-mode failure
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02f_pytest_gpu.log
if grep -q "direct_sharded" gpurun_out/r02f_pytest_gpu.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_sharded.py -k "direct_sharded_labels_equal_replicated and 2-run_rl-fetch" -x -q > gpurun_out/r02f_memcheck.log 2>&1
  grep -A25 "Invalid" gpurun_out/r02f_memcheck.log | head -60
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -2 gpurun_out/r02f_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_encode_probe_fast -s 2 -c 1 \
    -o gpurun_out/r02f_probe_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02f_probe_full.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e6,1), j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"], j["labels_checksum_rank0"], (j.get("e2e") or {}).get("value"))
PY
