#!/bin/bash
# 8 x B200: replicated (device + e2e), direct and exchange arms on the C2 table; C4 (table larger than one GPU) direct + exchange
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err; echo "$name rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r02u_$name.err | grep -E "Error|error" | head -3; }
run replicated
run direct --table-mode direct --no-e2e
run sharded_nccl --table-mode sharded --exchange nccl
run c4_direct --workload C4 --table-mode direct --no-e2e
run c4_sharded_nccl --workload C4 --table-mode sharded --exchange nccl
python - <<'PY'
import json
for n in ("replicated", "direct", "sharded_nccl", "c4_direct", "c4_sharded_nccl"):
    try:
        j = json.loads(open(f"gpurun_out/r02u_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), round(j["ms_per_step"],1), j.get("labels_checksum_rank0"), j.get("reads_error"), j["config"].get("db_bytes", j["config"].get("db_bytes_per_gpu")), j["config"].get("db_kmers"), "e2e", round(((j.get("e2e") or {}).get("value") or 0)/1e6,1), "ascii", round(((j.get("e2e_ascii") or {}).get("value") or 0)/1e6,1), j.get("kernels_ms"))
    except Exception as e:
        print(n, "failed", e)
PY
