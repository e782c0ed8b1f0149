#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python tools/cli_bench.py --reads 64000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/r02y_a.json 2> gpurun_out/r02y_a.err; tail -3 gpurun_out/r02y_a.err; cat gpurun_out/r02y_a.json
