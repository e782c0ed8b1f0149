#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_format.py -m gpu -q -x > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r03a_pytest.log | cut -c1-300
timeout 1200 python -m pytest tests/test_cli.py tests/test_c1_example.py -m gpu -q -x > gpurun_out/r03a_pytest_cli.log 2>&1; echo "pytest cli rc=$?"; tail -5 gpurun_out/r03a_pytest_cli.log | cut -c1-300
timeout 1200 python tools/cli_bench.py --reads 64000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/r03a_cli.json 2> gpurun_out/r03a_cli.err; tail -3 gpurun_out/r03a_cli.err; cat gpurun_out/r03a_cli.json
