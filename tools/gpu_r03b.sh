#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_format.py tests/test_cli.py -m gpu -q -x > gpurun_out/r03b_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r03b_pytest.log | cut -c1-300
timeout 1200 python tools/cli_bench.py --reads 64000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/r03b_cli.json 2> gpurun_out/r03b_cli.err; tail -2 gpurun_out/r03b_cli.err; cat gpurun_out/r03b_cli.json
timeout 1200 python tools/cli_bench.py --reads 64000000 --threads 10 --env KMAT_CLI_TRACE=1 --env KMAT_READER_THREADS=6 > gpurun_out/r03b_cli6.json 2> gpurun_out/r03b_cli6.err; tail -2 gpurun_out/r03b_cli6.err; cat gpurun_out/r03b_cli6.json
