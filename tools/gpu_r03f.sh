#!/bin/bash
# final validation of round 2 on one B200: GPU suite, smoke, the driver's two bench commands, ncu launch list / traffic /
# full capture of the three per-read kernels on the bench workload, the C3 workload, file-to-file throughput
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r03f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03f_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r03f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r03f_smoke.log
( time timeout 1500 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r03f_bench_reference.json 2> gpurun_out/r03f_bench_reference.err ) 2>&1 | grep real
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum --clock-control none -k regex:km_encode_probe_fast -c 3 --csv --log-file gpurun_out/r03f_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu traffic rc=$?"
python tools/ncu_traffic.py gpurun_out/r03f_traffic.csv profiles/traffic.json 10000000 2000 150 500000 | cut -c1-300
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r03f_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^km_' -c 120 --csv --log-file gpurun_out/r03f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'km_(encode_probe_fast|cand|score)_kernel' -s 3 -c 3 -o gpurun_out/r03f_k123_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu full rc=$?"
timeout 900 python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03f_bench_c3.json 2> gpurun_out/r03f_bench_c3.err; tail -2 gpurun_out/r03f_bench_c3.err
timeout 900 python tools/cli_bench.py --reads 64000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/r03f_cli_bench.json 2> gpurun_out/r03f_cli_bench.err; tail -1 gpurun_out/r03f_cli_bench.err; cut -c1-300 gpurun_out/r03f_cli_bench.json
python - <<'PY'
import json
for n in ("bench", "bench_c3"):
    try:
        j = json.loads(open(f"gpurun_out/r03f_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]/1e6,1), j["kernels_ms"], j["roofline"]["frac"], j["roofline"].get("request_rate_frac"), j["roofline"].get("traffic"), j.get("extra_buckets_per_lookup"), j["config"]["db_bytes"])
        print("   e2e", (j.get("e2e") or {}).get("value"), "ascii", (j.get("e2e_ascii") or {}).get("value"), "cpu", j.get("cpu_baseline"), "parity", j.get("parity_at_scale"))
    except Exception as e:
        print(n, "failed", e)
try:
    j = json.loads(open("gpurun_out/r03f_bench_reference.json").read().strip().splitlines()[-1]); print("reference", j["value"], j.get("cpu_baseline"))
except Exception as e:
    print("reference failed", e)
PY
