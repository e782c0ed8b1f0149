#!/bin/bash
# gpurun --timeout 1500 -- tools/gpu_k4_packed.sh : the scoring kernel with the depth packed into the sorted element
# (-DKMAT_K4_PACKED_DEPTH=1, lmat_b200/csrc/kmat_label.cu) against the default build: parity tests, then the bench line of
# each build (kernel split in "kernel_ms").  The default library is restored at the end.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/k4_default.json 2> gpurun_out/k4_default.err
touch lmat_b200/csrc/kmat_label.cu
KMAT_NVCC_DEFINES="-DKMAT_K4_PACKED_DEPTH=1" python -c "from lmat_b200 import build; build.build_all(force=True)"
python -m pytest tests -m gpu -x -q -k "parity or golden or cli" > gpurun_out/k4_packed_tests.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/k4_packed.json 2> gpurun_out/k4_packed.err
python -c "from lmat_b200 import build; build.build_all(force=True)"
tail -3 gpurun_out/k4_packed_tests.log; cat gpurun_out/k4_default.json gpurun_out/k4_packed.json
