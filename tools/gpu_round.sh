#!/bin/bash
# One GPU-box pass: parity tests, smoke, the full-size bench line, the ncu launch list of the same command
# and one `--set full` capture of each hot kernel on a reduced table.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:km_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launchlist rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'km_(encode_probe|cand|score)_kernel' -s 3 -c 3 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --genomes 400 --reads 2000000 \
    > gpurun_out/${TAG}_prof_bench.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
# DRAM traffic of the dominant kernel at the full bench workload (roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:km_encode_probe -s 1 -c 3 --csv --log-file gpurun_out/${TAG}_traffic.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_traffic_bench.log 2>&1; echo "traffic rc=$?"
