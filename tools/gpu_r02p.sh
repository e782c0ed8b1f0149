#!/bin/bash
# one B200: the DB-sharded GPU tests incl. the shared-pool direct mode
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r02p_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02p_tests.log
