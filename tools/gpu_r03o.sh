#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'km_(format|pack_lists|compact|score|cand|encode_probe_fast)_kernel' -c 60 --csv --log-file gpurun_out/r03o_format_launches.csv python tools/format_bench.py 1000000 > gpurun_out/r03o_format.json 2> gpurun_out/r03o_format.err; echo "rc=$?"; tail -2 gpurun_out/r03o_format.err | cut -c1-300; tail -1 gpurun_out/r03o_format.json
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/r03o_format_launches.csv") if l.startswith('"')))
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-14:]:
    print(r[ik].split("(")[0][:50], float(r[iv].replace(",", "")) / 1e6, "ms")
PY
