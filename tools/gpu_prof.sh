#!/bin/bash
# ncu --set full capture of the hot kernels on a reduced table.  usage: gpu_prof.sh <tag> [kernel regex]
set -u
mkdir -p gpurun_out
TAG=${1:-p}; KRE=${2:-'km_(encode_probe|cand|score)'}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 3 -c 3 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --genomes 400 --reads 2000000 \
    > gpurun_out/${TAG}_prof_bench.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${TAG}_prof_bench.log | cut -c1-400
