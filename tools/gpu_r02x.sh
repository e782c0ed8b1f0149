#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cli.py tests/test_reader_cpu.py tests/test_c1_example.py -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02x_pytest.log
TAG=r02x tools/gpu_cli.sh
timeout 900 python tools/cli_bench.py --reads 16000000 --threads 12 --env KMAT_CLI_TRACE=1 --env KMAT_WORKERS_PER_GPU=1 > gpurun_out/r02x_c.json 2> gpurun_out/r02x_c.err; tail -2 gpurun_out/r02x_c.err; cat gpurun_out/r02x_c.json
timeout 900 python tools/cli_bench.py --reads 16000000 --threads 12 --env KMAT_CLI_TRACE=1 --env KMAT_NO_PINNED=1 > gpurun_out/r02x_d.json 2> gpurun_out/r02x_d.err; tail -2 gpurun_out/r02x_d.err; cat gpurun_out/r02x_d.json
