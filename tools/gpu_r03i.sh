#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_format.py tests/test_cli.py -m gpu -q -x > gpurun_out/r03i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r03i_pytest.log | cut -c1-300
