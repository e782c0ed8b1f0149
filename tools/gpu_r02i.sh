#!/bin/bash
# 2 GPUs: NCCL exchange inside libkmat (tests + bench, sharded nccl vs torch driver vs replicated), e2e of the replicated mode at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "nccl or sharded_labels_equal" > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02i_tests.log
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err; tail -2 gpurun_out/r02i_$name.err; }
run sharded_nccl --table-mode sharded --exchange nccl
run sharded_torch --table-mode sharded --exchange torch
run replicated
python - <<'PY'
import json
for n in ("sharded_nccl", "sharded_torch", "replicated"):
    try:
        j = json.loads(open(f"gpurun_out/r02i_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), j["ms_per_step"], j.get("labels_checksum_rank0"), j.get("reads_error"), j.get("phase_ms_rank0"), (j.get("e2e") or {}).get("value"), (j.get("e2e_ascii") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
