#!/bin/bash
# final check of the committed code on one B200: GPU suite, smoke, the driver's two bench commands
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r03l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03l_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r03l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r03l_smoke.log
( time timeout 1500 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r03l_bench_reference.json 2> gpurun_out/r03l_bench_reference.err ) 2>&1 | grep real
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r03l_bench.json 2> gpurun_out/r03l_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r03l_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r03l_bench.json").read().strip().splitlines()[-1])
e = j["e2e"]
print(round(j["value"]/1e6,1), j["kernels_ms"], j["roofline"]["frac"], j["roofline"]["traffic"], "e2e", round(e["value"]/1e6,1), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], "plain", round(e["plain_pairs"]["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1), j.get("parity_at_scale"), j.get("cpu_baseline"), j["gpu_launches"])
j = json.loads(open("gpurun_out/r03l_bench_reference.json").read().strip().splitlines()[-1]); print("reference", j["value"])
PY
