#!/bin/bash
# the driver's own-arm command on the final code
set -u
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r03q_bench.json 2> gpurun_out/r03q_bench.err; echo "rc=$?"; tail -1 gpurun_out/r03q_bench.json | cut -c1-200
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r03q_bench.json").read().strip().splitlines()[-1])
e = j["e2e"]
print(round(j["value"]/1e6,1), j["kernels_ms"], round(j["roofline"]["frac"],3), "e2e", round(e["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1), (j.get("parity_at_scale") or {}).get("lines_equal"), (j.get("cpu_baseline") or {}).get("value"), j["clocks"])
PY
