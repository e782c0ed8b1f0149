#!/bin/bash
# Multi-GPU pass on one box: the replicated (mode A) and DB-sharded (mode B) bench lines at N GPUs.  usage: gpu_scale.sh <tag> <N> ["modes"]
set -u
mkdir -p gpurun_out
TAG=${1:-s}; N=${2:-2}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $2 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  echo "$1 rc=$?"; tail -c 1500 gpurun_out/${TAG}_$1.json; tail -3 gpurun_out/${TAG}_$1.err
}
MODES=${3:-"replicated sharded direct"}
for m in $MODES; do
  case $m in replicated) run replicated "";; *) run $m "--table-mode $m";; esac
done
