#!/bin/bash
# 2 GPUs: the drop-in binary on two distinct devices (replicated / direct peer reads / NCCL exchange) + the NCCL tests
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_sharded.py -m gpu -q -k "two_distinct or nccl" > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02j_tests.log
