#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cli.py tests/test_reader_cpu.py tests/test_c1_example.py -m gpu -q -x > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest.log
timeout 1200 python tools/cli_bench.py --reads 64000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/r02z_a.json 2> gpurun_out/r02z_a.err; tail -3 gpurun_out/r02z_a.err; cat gpurun_out/r02z_a.json
