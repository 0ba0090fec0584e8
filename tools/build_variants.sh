#!/bin/bash
# usage: tools/build_variants.sh <file.cu> name1 "flags1" name2 "flags2" ...
# Builds seqhasher_b200/_variants/lib_<name>.so = the current objects with <file.cu> recompiled under extra -D flags
# (A/B timing on the GPU box: SHB_LIB=<path> python tools/ab_c3.py).
set -e
cd "$(dirname "$0")/../seqhasher_b200/csrc"
src=$1; shift
mkdir -p ../_variants
base=$(basename $src .cu)
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2 -ccbin /usr/bin/g++ $flags -c $src -o ../_variants/${base}_$name.o &&
    objs=""; for o in kernels tile long parse api; do if [ $o = $base ]; then objs="$objs ../_variants/${base}_$name.o"; else objs="$objs ../_build/$o.o"; fi; done;
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_variants/lib_$name.so $objs -cudart static && echo built $name ) &
done
wait
