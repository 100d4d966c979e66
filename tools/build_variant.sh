#!/bin/bash
# usage: tools/build_variant.sh <name> [-DPROP_...=..]...   -> ditreeonlineplanner_b200/libditree_<name>.so
# Rebuilds propagate.cu with the given macros and links it with the other (already built) objects.
set -e
cd /root/repo/ditreeonlineplanner_b200
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c csrc/propagate.cu -o build/propagate_$name.o 2>&1 | grep -A2 "rowsILb1" | grep -E "Used|spill"
objs=""; for f in ctx geom reduce probmap cond gemm denoiser; do objs="$objs build/$f.o"; done
nvcc -shared -o libditree_$name.so $objs build/propagate_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lpthread -lrt
echo built libditree_$name.so
