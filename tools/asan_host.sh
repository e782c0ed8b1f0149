#!/bin/bash
# Host code of libkmat under AddressSanitizer + UBSan: the .cpp sources are rebuilt with -fsanitize=address,undefined and
# linked with the regular CUDA objects into /tmp/kmat_asan/libkmat.so; the CPU test suite (no GPU needed) then runs
# against that library through KMAT_LIB.  Usage: tools/asan_host.sh [pytest args]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/kmat_asan
mkdir -p $OUT
python -c "import sys; sys.path.insert(0, '$ROOT'); from lmat_b200 import build; build.build_all()"
OBJS=""
for s in kmat_host kmat_reader kmat_build; do
    g++ -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17 -fPIC \
        -I$ROOT/include -c $ROOT/lmat_b200/csrc/$s.cpp -o $OUT/$s.o
    OBJS="$OBJS $OUT/$s.o"
done
g++ -shared -fsanitize=address,undefined -o $OUT/libkmat.so $OBJS $ROOT/lmat_b200/build/kmat_db.cu.o $ROOT/lmat_b200/build/kmat_label.cu.o \
    -L/usr/local/cuda/lib64 -lcudart -lz -lpthread
cd $ROOT
ASAN_LIB=$(gcc -print-file-name=libasan.so)
LD_PRELOAD=$ASAN_LIB ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    KMAT_LIB=$OUT/libkmat.so python -m pytest tests -x -q -m "not gpu" "${@}"
