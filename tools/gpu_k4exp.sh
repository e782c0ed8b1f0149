#!/bin/bash
# K4 occupancy experiment: dummy dynamic shared memory caps the resident CTAs of km_score_kernel (KMAT_SCORE_SMEM)
mkdir -p gpurun_out
for sm in 0 16384 24576 32768 49152 65536 98304; do
  KMAT_SCORE_SMEM=$sm timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('score_smem', $sm, 'value', round(j['value'] / 1e6, 1), j['kernels_ms'])" | tee -a gpurun_out/${1:-k4}_exp.txt
done
