#!/bin/bash
# one B200: BASELINE configs[4]-style long reads (100 k x 10 kbp) with the round's final code
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --reads 100000 --read-len 10000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03n_bench_long.json 2> gpurun_out/r03n_bench_long.err; echo "rc=$?"; tail -2 gpurun_out/r03n_bench_long.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r03n_bench_long.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["ms_per_step"], j["kernels_ms"], "e2e", (j.get("e2e") or {}).get("value"), "ascii", (j.get("e2e_ascii") or {}).get("value"), j.get("reads_error"))
PY
