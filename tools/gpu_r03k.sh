#!/bin/bash
# 8 x B200: replicated table, end to end through the run-length compact interface
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03k_n8.json 2> gpurun_out/r03k_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r03k_n8.json").read().strip().splitlines()[-1])
e = j["e2e"]
print(round(j["value"]/1e6,1), "e2e rl", round(e["value"]/1e6,1), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], "plain", round(e["plain_pairs"]["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1), j["labels_checksum_rank0"], e["labels_checksum"])
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r03k_n8.err | grep -E "Error|error" | head -3
