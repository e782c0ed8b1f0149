#!/bin/bash
# 8 x B200, C4 in direct mode (every shard holds the same list pool: nothing to copy)
set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 5 --warmup 3 --workload C4 --table-mode direct --no-e2e > gpurun_out/r02n_c4_direct.json 2> gpurun_out/r02n_c4_direct.err; echo "rc=$?"
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r02n_c4_direct.err | grep -E "Error|error" | head -3
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02n_c4_direct.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e6,1), round(j["ms_per_step"],1), j.get("labels_checksum_rank0"), j.get("reads_error"), j["config"].get("db_bytes"), j["config"].get("db_kmers"), j.get("kernels_ms"))
PY
