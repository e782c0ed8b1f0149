#!/bin/bash
# one B200: where the host-buffer pipeline loses time -- per-chunk trace, chunk sizes
set -u
mkdir -p gpurun_out
KMAT_PIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_trace.json 2> gpurun_out/r02q_trace.err; echo "trace rc=$?"
for cr in 262144 524288 2097152 5000000; do
  KMAT_CHUNK_READS=$cr timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_chunk_$cr.json 2> gpurun_out/r02q_chunk_$cr.err; echo "$cr rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02q_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]/1e6,1), "e2e", round(j["e2e"]["value"]/1e6,1), "ascii", round(j["e2e_ascii"]["value"]/1e6,1))
    except Exception as e:
        print(f, "failed", e)
PY
grep -A14 "pipe trace" gpurun_out/r02q_trace.err | head -64
