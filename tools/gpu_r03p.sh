#!/bin/bash
# one B200: the long-read probe kernel with the stateless sliding minimum -- full GPU suite, then the long-read and the default bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r03p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03p_pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --reads 100000 --read-len 10000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r03p_bench_long.json 2> gpurun_out/r03p_bench_long.err; echo "long rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r03p_bench.json 2> gpurun_out/r03p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("bench_long", "bench"):
    try:
        j = json.loads(open(f"gpurun_out/r03p_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]), round(j["ms_per_step"], 2), j["kernels_ms"], j["labels_checksum_rank0"], j.get("reads_error"))
    except Exception as e:
        print(n, "failed", e)
PY
