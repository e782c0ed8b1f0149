#!/bin/bash
# Builds the compile-time experiment variants of libkmat HERE (nvcc cross-compiles without a GPU) into
# lmat_b200/variants/libkmat_<name>.so, so that the GPU box spends no time compiling: tools/gpu_k4_packed.sh picks them up
# through KMAT_LIB (lmat_b200/api.py).  *.so is git-ignored but travels with gpurun.  Usage: tools/build_variants.sh
# Remove lmat_b200/variants/ again once the experiments are read (30 MB per variant in every gpurun snapshot).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/lmat_b200/variants; mkdir -p $V/obj
python -c "import sys; sys.path.insert(0, '$ROOT'); from lmat_b200 import build; build.build_all()"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I$ROOT/include"
build_one() {   # name, defines
    local name=$1; shift
    $NVCC $FLAGS "$@" -c $ROOT/lmat_b200/csrc/kmat_label.cu -o $V/obj/kmat_label_$name.o
    $NVCC $FLAGS "$@" -c $ROOT/lmat_b200/csrc/kmat_db.cu -o $V/obj/kmat_db_$name.o
    $NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $V/libkmat_$name.so $V/obj/kmat_label_$name.o $V/obj/kmat_db_$name.o \
        $ROOT/lmat_b200/build/kmat_host.cpp.o $ROOT/lmat_b200/build/kmat_reader.cpp.o $ROOT/lmat_b200/build/kmat_build.cpp.o -lz -Xcompiler -fPIC
    echo "built $V/libkmat_$name.so ($*)"
}
build_one exp1 -DKMAT_K4_PACKED_DEPTH=1
build_one exp2 -DKMAT_K4_BLOCK_SORT=1
build_one exp3 -DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1
build_one exp4 -DKMAT_K4_PACKED_DEPTH=1 -DKMAT_K4_BLOCK_SORT=1 -DKS_SORT_ROUNDS=8
build_one exp5 -DKMAT_LINE_TABLE=1
build_one exp6 -DKMAT_LINE_TABLE=1 -DKMAT_LINE_SHFL=1
