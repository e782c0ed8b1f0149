#!/bin/bash
# Round 2, first GPU call: (1) parity suite + smoke on the tree as round 1 left it, (2) ncu --set full of the dominant
# kernel (km_encode_probe_fast_kernel<5,...>) on the full C2 workload, (3) line-gather rates, (4) the compile-time
# variants prepared in round 1 (tools/build_variants.sh must have run here first).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_encode_probe_fast -s 2 -c 1 \
    -o gpurun_out/r02a_probe_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e \
    > gpurun_out/r02a_probe_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02a_probe_full.ncu-rep
timeout 900 tools/gpu_line_gather.sh
timeout 2400 tools/gpu_k4_packed.sh
