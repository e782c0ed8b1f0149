#!/bin/bash
# 8 x B200, C4 (5200 genomes = 2.5 G k-mers: the table does not fit one GPU): NCCL exchange and direct peer reads
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 "$@" > gpurun_out/r02m_$name.json 2> gpurun_out/r02m_$name.err; echo "$name rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r02m_$name.err | grep -E "Error|error" | head -3; }
run c4_sharded_nccl --workload C4 --table-mode sharded --exchange nccl
run c4_direct --workload C4 --table-mode direct --no-e2e
python - <<'PY'
import json
for n in ("c4_sharded_nccl", "c4_direct"):
    try:
        j = json.loads(open(f"gpurun_out/r02m_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), round(j["ms_per_step"],1), j.get("labels_checksum_rank0"), j.get("reads_error"), j["config"].get("db_bytes", j["config"].get("db_bytes_per_gpu")), j["config"].get("db_kmers"), j.get("setup_s"), j.get("kernels_ms"))
    except Exception as e:
        print(n, "failed", e)
PY
