#!/bin/bash
# First GPU call of the next round (one box, ~25 min): everything that was prepared on the CPU after the last GPU
# minute of round 1 was spent, in the order of what decides the next design step.
#   1. full GPU parity suite + smoke on the current tree (host-side changes since the last GPU run: reader fallback,
#      table checks at upload, gather modes 201-205 -- the per-read kernels are unchanged)
#   2. line-gather rates (tools/line_gather.py): does the request ceiling count lines or sectors?  Decides whether the
#      minimizer-ordered table (lmat_b200/csrc/kmat_mzr.h) is worth its kernels.
#   3. the compile-time variants (tools/gpu_k4_packed.sh): exp5 = the line table (-DKMAT_LINE_TABLE=1) first, then the
#      scoring-kernel switches (packed depth, per-CTA sort by candidate count): parity tests + bench line each.
# Usage: tools/build_variants.sh (here, ~1 min) then gpurun --timeout 3600 -- tools/gpu_round3_first.sh
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 tools/gpu_line_gather.sh
timeout 2400 tools/gpu_k4_packed.sh
