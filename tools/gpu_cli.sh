#!/bin/bash
# one B200: file-to-file throughput of the drop-in binary with the per-stage busy times
set -u
mkdir -p gpurun_out
tag=${TAG:-cli}
timeout 900 python tools/cli_bench.py --reads 16000000 --threads 12 --env KMAT_CLI_TRACE=1 > gpurun_out/${tag}_a.json 2> gpurun_out/${tag}_a.err; tail -3 gpurun_out/${tag}_a.err; cat gpurun_out/${tag}_a.json
timeout 900 python tools/cli_bench.py --reads 16000000 --threads 12 --env KMAT_CLI_TRACE=1 --env KMAT_BATCH_READS=1000000 > gpurun_out/${tag}_b.json 2> gpurun_out/${tag}_b.err; tail -3 gpurun_out/${tag}_b.err; cat gpurun_out/${tag}_b.json
nproc; free -g | head -2; df -h /dev/shm | tail -1
