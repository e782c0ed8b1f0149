#!/bin/bash
# one B200: (1) C4 replicated at density 8 = the label checksum the two DB-sharded arms must reproduce; (2) A/B of the dedup bitmap size
set -u
mkdir -p gpurun_out
KMAT_LINE_DENSITY=8 timeout 1200 python bench.py --workload C4 --table-mode replicated --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02o_c4_replicated_d8.json 2> gpurun_out/r02o_c4_replicated_d8.err; echo "c4 rc=$?"; tail -3 gpurun_out/r02o_c4_replicated_d8.err
for v in head new; do
  if [ $v = head ]; then export KMAT_LIB=$PWD/lmat_b200/build/libkmat_head.so; else unset KMAT_LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02o_ab_$v.json 2> gpurun_out/r02o_ab_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for n in ("c4_replicated_d8", "ab_head", "ab_new"):
    try:
        j = json.loads(open(f"gpurun_out/r02o_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), round(j["ms_per_step"],2), j.get("labels_checksum_rank0"), j.get("reads_error"), j["config"].get("db_bytes"), j.get("extra_buckets_per_lookup"), j.get("kernels_ms"))
    except Exception as e:
        print(n, "failed", e)
PY
