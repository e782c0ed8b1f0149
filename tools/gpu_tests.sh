#!/bin/bash
# one B200: the GPU suite (or the tests given on the command line)
set -u
mkdir -p gpurun_out
tag=${TAG:-tests}
timeout 2400 python -m pytest ${@:-tests} -m gpu -q -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest_gpu.log
