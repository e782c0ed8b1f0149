#!/bin/bash
# one B200: host-buffer pipeline after the K4 window / graduated chunk changes (trace + plain run)
set -u
mkdir -p gpurun_out
KMAT_PIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_trace.json 2> gpurun_out/r02t_trace.err; echo "trace rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_plain.json 2> gpurun_out/r02t_plain.err; echo "plain rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02t_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]/1e6,1), "e2e", round(j["e2e"]["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1), j["labels_checksum_rank0"], j["e2e"]["labels_checksum"], j["e2e_ascii"]["labels_checksum"])
    except Exception as e:
        print(f, "failed", e)
PY
grep -A14 "pipe trace" gpurun_out/r02t_trace.err | sed -n 16,32p
