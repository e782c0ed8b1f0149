#!/bin/bash
# 8 x B200: where the host-buffer path loses time at N = 8 (per-chunk trace of every rank), and the N = 4 point of the scaling series
set -u
mkdir -p gpurun_out
KMAT_PIPE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r03g_n8_trace.json 2> gpurun_out/r03g_n8_trace.err; echo "n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03g_n4.json 2> gpurun_out/r03g_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for n in ("n8_trace", "n4"):
    try:
        j = json.loads(open(f"gpurun_out/r03g_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), "e2e", round(j["e2e"]["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1))
    except Exception as e:
        print(n, "failed", e)
PY
grep -A16 "pipe trace" gpurun_out/r03g_n8_trace.err | sed -n 18,34p
nproc; lscpu | grep -i "numa\|socket\|model name" | head -6
