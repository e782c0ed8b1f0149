#!/bin/bash
# build a compile-time variant of libkmat into lmat_b200/build/libkmat_<name>.so (loaded through KMAT_LIB on the GPU box)
# usage: tools/build_variant.sh <name> "<nvcc defines>"
set -eu
name=$1; defs=$2
cd "$(dirname "$0")/.."
B=lmat_b200/build/var_$name
mkdir -p $B
for f in kmat_db.cu kmat_label.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $defs -Iinclude -c lmat_b200/csrc/$f -o $B/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lmat_b200/build/libkmat_$name.so $B/kmat_db.cu.o $B/kmat_label.cu.o lmat_b200/build/kmat_host.cpp.o lmat_b200/build/kmat_reader.cpp.o lmat_b200/build/kmat_build.cpp.o -lz -Xcompiler -fPIC
ls -la lmat_b200/build/libkmat_$name.so
