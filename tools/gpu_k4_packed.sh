#!/bin/bash
# gpurun --timeout 2400 -- tools/gpu_k4_packed.sh : compile-time experiments on the scoring kernel (lmat_b200/csrc/kmat_label.cu)
# against the default build: for each define set the library is rebuilt, the parity tests run, then the bench line
# (kernel split in "kernels_ms").  The default library is restored at the end.
#   -DKMAT_K4_PACKED_DEPTH=1  depth carried in the sorted rank_label element (no local-memory loads in TCmp)
#   -DKMAT_K4_BLOCK_SORT=1    a CTA counting-sorts 512 queued reads by candidate count before scoring them
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k4_default.json 2> gpurun_out/k4_default.err
i=0
for defs in "-DKMAT_K4_PACKED_DEPTH=1" "-DKMAT_K4_BLOCK_SORT=1" "-DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1" "-DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1 -DKS_SORT_ROUNDS=8"; do
    i=$((i+1))
    KMAT_NVCC_DEFINES="$defs" python -c "from lmat_b200 import build; build.build_all(force=True)"
    python -m pytest tests -m gpu -x -q -k "parity or golden or cli" > gpurun_out/k4_exp${i}_tests.log 2>&1
    echo "$defs: $(tail -1 gpurun_out/k4_exp${i}_tests.log)"
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k4_exp${i}.json 2> gpurun_out/k4_exp${i}.err
done
python -c "from lmat_b200 import build; build.build_all(force=True)"
python - <<'PY'
import json
for n in ("default", "exp1", "exp2", "exp3", "exp4"):
    try:
        j = json.loads(open(f"gpurun_out/k4_{n}.json").read().strip().splitlines()[-1])
        print(n, j.get("value"), j.get("ms_per_step"), j.get("kernels_ms"))
    except Exception as e:
        print(n, "failed", e)
PY
