#!/bin/bash
# Quick GPU pass: parity tests + the full-size bench line (no CPU baseline).  usage: gpu_quick.sh <tag> [bench args]
set -u
mkdir -p gpurun_out
TAG=${1:-q}; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
