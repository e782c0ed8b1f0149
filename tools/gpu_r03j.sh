#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_compact_io.py tests/test_gpu_many_candidates.py -m gpu -q -x > gpurun_out/r03j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r03j_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r03j_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r03j_bench.json").read().strip().splitlines()[-1])
e = j["e2e"]
print(round(j["value"]/1e6,1), "e2e rl", round(e["value"]/1e6,1), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], "plain", round(e["plain_pairs"]["value"]/1e6,1), e["plain_pairs"]["d2h_bytes_per_step"], "ascii", round(j["e2e_ascii"]["value"]/1e6,1), j["labels_checksum_rank0"], e["labels_checksum"])
PY
