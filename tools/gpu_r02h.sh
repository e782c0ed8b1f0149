#!/bin/bash
# compact interface: tests + bench with the packed e2e leg; line density 2 vs 4
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compact_io.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02h_tests.log
for dens in 2 4; do
  KMAT_LINE_DENSITY=$dens timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_d$dens.json 2> gpurun_out/r02h_bench_d$dens.err; tail -2 gpurun_out/r02h_bench_d$dens.err
done
python - <<'PY'
import json
for n in ("d2", "d4"):
    j = json.loads(open(f"gpurun_out/r02h_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, round(j["value"]/1e6,1), j["kernels_ms"], j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"], j["labels_checksum_rank0"])
    print("   e2e", j["e2e"]); print("   e2e_ascii", j["e2e_ascii"])
PY
