#!/bin/bash
# Round-end GPU pass: parity tests, smoke, the bench line (with the CPU baseline), launch list, one --set full capture,
# DRAM traffic of the probe kernel, the null-model generator and the long-read (C5) shape.  Outputs: gpurun_out/<tag>_*.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_reference.json
timeout 600 python tools/null_bench.py > gpurun_out/${TAG}_null_bench.json 2> gpurun_out/${TAG}_null_bench.err; echo "null rc=$?"; cat gpurun_out/${TAG}_null_bench.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --reads 100000 --read-len 10000 > gpurun_out/${TAG}_bench_long.json 2> gpurun_out/${TAG}_bench_long.err; echo "long rc=$?"; tail -c 900 gpurun_out/${TAG}_bench_long.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:km_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launchlist rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'km_(encode_probe|cand|score)_kernel' -s 3 -c 3 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --genomes 400 --reads 2000000 \
    > gpurun_out/${TAG}_prof_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:km_encode_probe -s 1 -c 3 --csv --log-file gpurun_out/${TAG}_traffic.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_traffic_bench.log 2>&1; echo "traffic rc=$?"
ls -la gpurun_out | tail -20
