#!/bin/bash
# parity-at-scale test + the default bench run with its new CPU sample (256 genomes) and parity leg; reference arm first (as the driver does)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_scale.py -m gpu -x -q > gpurun_out/r02c_scale_test.log 2>&1; echo "scale test rc=$?"; tail -15 gpurun_out/r02c_scale_test.log
( time timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err ) 2>&1 | grep real
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
