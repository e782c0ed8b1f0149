#!/bin/bash
# 2 x B200: the tests that need two GPUs (drop-in binary on two devices in every table mode, NCCL exchange between two GPUs), final code
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sharded.py tests/test_cli.py -m gpu -q -rs > gpurun_out/r03h_two_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r03h_two_gpu_tests.log
