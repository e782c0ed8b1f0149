#!/bin/bash
# gpurun --timeout 2400 -- tools/gpu_k4_packed.sh : compile-time experiments on the scoring kernel (lmat_b200/csrc/kmat_label.cu)
# against the default build.  Run tools/build_variants.sh first (on the CPU box): the variants are then loaded through
# KMAT_LIB with no compile time on the GPU box; without them each variant is rebuilt there.  Per variant: the parity tests
# that go through the ctypes binding, then the bench line (kernel split in "kernels_ms").
#   exp1 -DKMAT_K4_PACKED_DEPTH=1     depth carried in the sorted rank_label element (no local-memory loads in TCmp)
#   exp2 -DKMAT_K4_BLOCK_SORT=1       a CTA counting-sorts 512 queued reads by candidate count before scoring them
#   exp3 both                          exp4 both, 1024 reads per CTA
#   exp5 -DKMAT_LINE_TABLE=1           the minimizer-ordered line table (kmat_mzr.h) on today's slot format: replicated table
#                                      only (the sharded tests are skipped for it); 69 GB table for the bench workload
#   exp6 exp5 + -DKMAT_LINE_SHFL=1     the minimizer from one hash per base and a sliding minimum over the lanes
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k4_default.json 2> gpurun_out/k4_default.err
DEFS=("" "-DKMAT_K4_PACKED_DEPTH=1" "-DKMAT_K4_BLOCK_SORT=1" "-DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1" "-DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1 -DKS_SORT_ROUNDS=8" "-DKMAT_LINE_TABLE=1" "-DKMAT_LINE_TABLE=1 -DKMAT_LINE_SHFL=1")
rebuilt=0
for i in 5 6 1 2 3 4; do
    lib=$PWD/lmat_b200/variants/libkmat_exp$i.so
    if [ -f "$lib" ]; then export KMAT_LIB=$lib
    else unset KMAT_LIB; rebuilt=1; KMAT_NVCC_DEFINES="${DEFS[$i]}" python -c "from lmat_b200 import build; build.build_all(force=True)"; fi
    T="tests/test_gpu_parity.py tests/test_gpu_sharded.py"; [ $i -ge 5 ] && T="tests/test_gpu_parity.py"
    python -m pytest $T -m gpu -q > gpurun_out/k4_exp${i}_tests.log 2>&1
    echo "exp$i ${DEFS[$i]}: $(tail -1 gpurun_out/k4_exp${i}_tests.log)"
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k4_exp${i}.json 2> gpurun_out/k4_exp${i}.err
done
unset KMAT_LIB
[ $rebuilt = 1 ] && python -c "from lmat_b200 import build; build.build_all(force=True)"
python - <<'PY'
import json
for n in ("default", "exp5", "exp6", "exp1", "exp2", "exp3", "exp4"):
    try:
        j = json.loads(open(f"gpurun_out/k4_{n}.json").read().strip().splitlines()[-1])
        print(n, j.get("value"), j.get("ms_per_step"), j.get("kernels_ms"), (j.get("roofline") or {}).get("traffic"))
    except Exception as e:
        print(n, "failed", e)
PY
