#!/bin/bash
# one B200: bench.py on variant builds of libkmat (tools/build_variant.sh).  usage: tools/gpu_variants.sh <tag> <name>... ("base" = the in-tree library)
set -u
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = base ]; then unset KMAT_LIB; else export KMAT_LIB=$PWD/lmat_b200/build/libkmat_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err; echo "$v rc=$?"
done
python - "$tag" "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for v in sys.argv[2:]:
    try:
        j = json.loads(open(f"gpurun_out/{tag}_{v}.json").read().strip().splitlines()[-1])
        k = j["kernels_ms"]
        print(f"{v:10s} dev {j['value']/1e6:7.1f}  ms {j['ms_per_step']:6.2f}  probe {k['encode_probe']:.2f} cand {k['candidates']:.2f} score {k['score']:.2f}  e2e {(j.get('e2e') or {}).get('value', 0)/1e6:7.1f} ascii {(j.get('e2e_ascii') or {}).get('value', 0)/1e6:7.1f}  chk {j.get('labels_checksum_rank0')} {(j.get('e2e') or {}).get('labels_checksum')}")
    except Exception as e:
        print(v, "failed", e)
PY
