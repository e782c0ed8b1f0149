#!/bin/bash
# one B200: launch list of the product kernels, device-resident pass (10 M reads per launch) vs the chunks of the host-buffer path
set -u
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^km_' -c 700 --csv --log-file gpurun_out/r02r_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_under_ncu.json 2> gpurun_out/r02r.err; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02r_launches.csv")))
hdr = None; seq = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); seq.append((d["Kernel Name"].split("(")[0][:60], float(d["Metric Value"].replace(",", "")), d.get("Grid Size", "")))
print(len(seq), "launches")
for i, s in enumerate(seq[:140]): print(i, s)
PY
